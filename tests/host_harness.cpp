// Test-only host build of taichi_elements_b200/csrc/mpm_math.cuh: lets the CPU
// test-suite check the exact device math (SVD, plasticity, stress) against the
// oracle without a GPU.  Never loaded by the product package.
#include "../taichi_elements_b200/csrc/mpm_math.cuh"
#include "../taichi_elements_b200/csrc/mpm_quant.cuh"
extern "C" {
void host_svd3(const float* F, int n, float* U, float* sig, float* V) {
  for (int i = 0; i < n; ++i) mpm::svd3(F + 9 * i, U + 9 * i, sig + 3 * i, V + 9 * i);
}
void host_svd2(const float* F, int n, float* U, float* sig, float* V) {
  for (int i = 0; i < n; ++i) mpm::svd2(F + 4 * i, U + 4 * i, sig + 2 * i, V + 4 * i);
}
// arrays are AoS row-major per particle
void host_particle_update(int dim, const float* consts12, int support_plasticity, float dt, int n,
                          const int* material, float* F, const float* C, float* Jp,
                          float* affine, float* mass, int* fast_taken) {
  mpm::Consts K{};
  K.dx = consts12[0]; K.inv_dx = consts12[1]; K.p_vol = consts12[2]; K.p_mass = consts12[3];
  K.mu_0 = consts12[4]; K.lambda_0 = consts12[5]; K.alpha = consts12[6]; K.sand_coef = consts12[7];
  K.water_density = consts12[8]; K.inv_dx2 = consts12[9]; K.four_inv_dx = consts12[10];
  K.support_plasticity = support_plasticity;
  K.g2p2g = (int)consts12[11] & 1;      // fused-mode variant
  K.clamp_F = ((int)consts12[11] >> 1) & 1;
  for (int i = 0; i < n; ++i) {
    if (fast_taken) {     // which particles avoid the SVD (the 3D P2G kernel runs the others in a second pass)
      float Fn[9], a9[9], m, jp = Jp[i];
      if (dim == 2) { mpm::trial_F<2>(K, dt, material[i], F + 4 * i, C + 4 * i, jp, Fn);
                      fast_taken[i] = mpm::particle_update_fast<2>(K, dt, material[i], Fn, C + 4 * i, jp, a9, m); }
      else { mpm::trial_F<3>(K, dt, material[i], F + 9 * i, C + 9 * i, jp, Fn);
             fast_taken[i] = mpm::particle_update_fast<3>(K, dt, material[i], Fn, C + 9 * i, jp, a9, m); }
    }
    if (dim == 2)
      mpm::particle_update<2>(K, dt, material[i], F + 4 * i, C + 4 * i, Jp[i], affine + 4 * i, mass[i]);
    else
      mpm::particle_update<3>(K, dt, material[i], F + 9 * i, C + 9 * i, Jp[i], affine + 9 * i, mass[i]);
  }
}
// quantised storage codecs (csrc/mpm_quant.cuh): encode -> packed words -> decode; kind 0 = x, 1 = v, 2 = F
void host_quant_round(int kind, int n, const float* in, float* out, unsigned* words) {
  for (int i = 0; i < n; ++i) {
    if (kind == 0) { mpm::encode_x3(in + 3 * i, words + 2 * i); mpm::decode_x3(words + 2 * i, out + 3 * i); }
    else if (kind == 1) { mpm::encode_v3(in + 3 * i, words + 2 * i); mpm::decode_v3(words + 2 * i, out + 3 * i); }
    else { mpm::encode_F9(in + 9 * i, words + 5 * i); mpm::decode_F9(words + 5 * i, out + 9 * i); }
  }
}
}
