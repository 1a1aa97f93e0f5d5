// Per-particle constitutive math of the MLS-MPM P2G step, written once as
// __host__ __device__ so the same code is unit-tested on the CPU (tests/ build
// a host harness from this header) and runs inside the sm_100a kernels.
//
// Follows /root/reference/engine/mpm_solver.py:506-574 (F update, hardening,
// SVD, plasticity, stress, affine) and :321-342 (sand_projection).  ti.svd is
// external to the reference tree (Taichi 1.1.0); its convention -- U, V proper
// rotations, sign of det F on the last singular value -- is restated here.
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define MPM_HD __host__ __device__ __forceinline__
#else
#define MPM_HD inline
#endif

namespace mpm {

enum Material { WATER = 0, ELASTIC = 1, SNOW = 2, SAND = 3, STATIONARY = 4 };

struct Consts {
  float dx, inv_dx;
  float p_vol, p_mass;
  float mu_0, lambda_0;
  float alpha;            // sand friction coefficient (:208-210)
  float sand_coef;        // (dim*lambda_0 + 2 mu_0) / (2 mu_0)  (:336-337)
  float water_density;
  float inv_dx2;          // inv_dx**2 as one f32 constant (:569)
  float four_inv_dx;      // 4*inv_dx as one f32 constant (:721)
  int support_plasticity;
  int g2p2g;              // P2G half of the fused kernel (:405-483): see particle_update
  int clamp_F;            // quant: F clamped to +-F_bound in g2p2g (:99, 415-416)
  float v_allowed_cfl;    // dx * g2p2g_allowed_cfl, 0 = no grid-velocity clamp (:589, 596-598)
};

// Approximate (MUFU-based, ~1-2 ulp) division / reciprocal square root on the
// device: IEEE division and sqrt carry slow-path subroutines that ncu showed to
// be 13% of P2G's instructions; every use below is far inside the 1e-4 parity
// budget.  On the host (test harness) they are the exact operations.
MPM_HD float rsqrt_(float x) {
#if defined(__CUDA_ARCH__)
  return rsqrtf(x);
#else
  return 1.0f / sqrtf(x);
#endif
}
MPM_HD float fdiv_(float a, float b) {
#if defined(__CUDA_ARCH__)
  return __fdividef(a, b);
#else
  return a / b;
#endif
}
MPM_HD float sqrt_pos_(float x) {   // x >= 0, normal range
#if defined(__CUDA_ARCH__)
  return x * rsqrtf(fmaxf(x, 1e-37f));
#else
  return sqrtf(x);
#endif
}

// ---------------------------------------------------------------------------
// Fixed-point rounding of the quantised storage (quant=True, mpm_quant.cuh; ref :106-114): used here because the
// fused kernel's `self.F[p] = new_F` (:416) rounds F to its 16-bit grid BEFORE the constitutive model reads it back.
// ---------------------------------------------------------------------------
static constexpr float QX_MAX = 2.0f;      // :107
static constexpr int QX_BITS = 21;
static constexpr float QF_MAX = 4.1f;      // F_bound + 0.1 (:99, 113)
static constexpr int QF_BITS = 16;
static constexpr int QV_FRAC = 19, QV_EXP = 7;

MPM_HD int q_round(float t) { return (int)(t + (t < 0.0f ? -0.5f : 0.5f)); }       // round half away from zero
MPM_HD int q_fixed(float x, float max_value, int bits) {
  const float inv_scale = (float)(1 << (bits - 1)) / max_value;
  const int lim = (1 << (bits - 1)) - 1;
  const float t = x * inv_scale;
  if (!(t > -(float)lim)) return -lim;                                             // (also NaN)
  if (t > (float)lim) return lim;
  return q_round(t);
}
MPM_HD float dq_fixed(int q, float max_value, int bits) { return (float)q * (max_value / (float)(1 << (bits - 1))); }

MPM_HD void round_F9(float* F) {
#pragma unroll
  for (int i = 0; i < 9; ++i) F[i] = dq_fixed(q_fixed(F[i], QF_MAX, QF_BITS), QF_MAX, QF_BITS);
}

// ---------------------------------------------------------------------------
// 2x2 SVD, Taichi's closed form.  Row-major 2x2.
// ---------------------------------------------------------------------------
MPM_HD void svd2(const float* F, float* U, float* sig, float* V) {
  float a = F[0] + F[3];
  float b = F[2] - F[1];
  float s = 1.0f / sqrtf(a * a + b * b);
  float rc = a * s, rs = b * s;
  // R = [[rc,-rs],[rs,rc]],  S = R^T F (symmetric)
  float S00 = rc * F[0] + rs * F[2];
  float S01 = rc * F[1] + rs * F[3];
  float S11 = -rs * F[1] + rc * F[3];
  float c = 1.0f, sn = 0.0f, s1 = S00, s2 = S11;
  if (fabsf(S01) >= 1e-5f) {
    float tau = 0.5f * (S00 - S11);
    float w = sqrtf(tau * tau + S01 * S01);
    float t = tau > 0.0f ? S01 / (tau + w) : S01 / (tau - w);
    c = 1.0f / sqrtf(t * t + 1.0f);
    sn = -t * c;
    float c2 = c * c, sn2 = sn * sn, cs2 = 2.0f * c * sn * S01;
    s1 = c2 * S00 - cs2 + sn2 * S11;
    s2 = sn2 * S00 + cs2 + c2 * S11;
  }
  float v00, v01, v10, v11;
  if (s1 < s2) {
    sig[0] = s2; sig[1] = s1;
    v00 = -sn; v01 = c; v10 = -c; v11 = -sn;
  } else {
    sig[0] = s1; sig[1] = s2;
    v00 = c; v01 = sn; v10 = -sn; v11 = c;
  }
  V[0] = v00; V[1] = v01; V[2] = v10; V[3] = v11;
  U[0] = rc * v00 - rs * v10; U[1] = rc * v01 - rs * v11;
  U[2] = rs * v00 + rc * v10; U[3] = rs * v01 + rc * v11;
}

// ---------------------------------------------------------------------------
// 3x3 SVD: cyclic Jacobi on S = F^T F (exact rotations, fixed sweep count so
// warps stay convergent), eigenvalues sorted descending, then U and sigma from
// B = F V by re-orthogonalised Gram-Schmidt with u2 = u0 x u1 so that U is a
// rotation and sigma[2] carries the sign of det F (Sifakis/Taichi convention).
// Row-major 3x3.
// ---------------------------------------------------------------------------
MPM_HD void jacobi_rot(float& app, float& aqq, float& apq, float& arp, float& arq,
                       float* V, int p, int q) {
  // rotate in the (p,q) plane so that apq -> 0; r is the third index
  float c = 1.0f, s = 0.0f;
  if (fabsf(apq) > 1e-30f) {
    float theta = fdiv_(aqq - app, 2.0f * apq);
    float at = fminf(fabsf(theta), 1e18f);   // keeps theta^2 finite (rsqrt-based sqrt maps inf to NaN)
    float t = fdiv_(1.0f, at + sqrt_pos_(at * at + 1.0f));
    t = theta < 0.0f ? -t : t;
    c = rsqrt_(t * t + 1.0f);
    s = t * c;
    float tpq = t * apq;
    app -= tpq;
    aqq += tpq;
    apq = 0.0f;
    float nrp = c * arp - s * arq;
    float nrq = s * arp + c * arq;
    arp = nrp; arq = nrq;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      float vp = V[k * 3 + p], vq = V[k * 3 + q];
      V[k * 3 + p] = c * vp - s * vq;
      V[k * 3 + q] = s * vp + c * vq;
    }
  }
}

MPM_HD void svd3(const float* F, float* U, float* sig, float* V) {
  // S = F^T F
  float a00 = F[0] * F[0] + F[3] * F[3] + F[6] * F[6];
  float a11 = F[1] * F[1] + F[4] * F[4] + F[7] * F[7];
  float a22 = F[2] * F[2] + F[5] * F[5] + F[8] * F[8];
  float a01 = F[0] * F[1] + F[3] * F[4] + F[6] * F[7];
  float a02 = F[0] * F[2] + F[3] * F[5] + F[6] * F[8];
  float a12 = F[1] * F[2] + F[4] * F[5] + F[7] * F[8];
  V[0] = 1; V[1] = 0; V[2] = 0; V[3] = 0; V[4] = 1; V[5] = 0; V[6] = 0; V[7] = 0; V[8] = 1;
#pragma unroll 1
  for (int sweep = 0; sweep < 4; ++sweep) {
    // F^T F of an MPM particle is close to diagonal: most lanes are done after two sweeps
    if (a01 * a01 + a02 * a02 + a12 * a12 <= 1e-15f * (a00 * a00 + a11 * a11 + a22 * a22)) break;
    jacobi_rot(a00, a11, a01, a02, a12, V, 0, 1);
    jacobi_rot(a00, a22, a02, a01, a12, V, 0, 2);
    jacobi_rot(a11, a22, a12, a01, a02, V, 1, 2);
  }
  // sort eigenvalues descending; a column swap is paired with a sign flip so
  // det V stays +1
#define MPM_SWAPCOL(i, j, ei, ej)                      \
  if (ei < ej) {                                       \
    float te = ei; ei = ej; ej = te;                   \
    for (int k = 0; k < 3; ++k) {                      \
      float tv = V[k * 3 + i];                         \
      V[k * 3 + i] = V[k * 3 + j];                     \
      V[k * 3 + j] = -tv;                              \
    }                                                  \
  }
  MPM_SWAPCOL(0, 1, a00, a11)
  MPM_SWAPCOL(0, 2, a00, a22)
  MPM_SWAPCOL(1, 2, a11, a22)
#undef MPM_SWAPCOL
  // B = F V
  float B[9];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j)
      B[i * 3 + j] = F[i * 3 + 0] * V[0 * 3 + j] + F[i * 3 + 1] * V[1 * 3 + j] +
                     F[i * 3 + 2] * V[2 * 3 + j];
  // u0 = b0 / |b0|
  float n0sq = B[0] * B[0] + B[3] * B[3] + B[6] * B[6];
  float n0 = sqrt_pos_(n0sq);
  float u00, u10, u20;
  if (n0 > 1e-30f) { float r = rsqrt_(n0sq); u00 = B[0] * r; u10 = B[3] * r; u20 = B[6] * r; }
  else { u00 = 1; u10 = 0; u20 = 0; }
  // u1 = normalise(b1 - (u0.b1) u0)
  float d01 = u00 * B[1] + u10 * B[4] + u20 * B[7];
  float w0 = B[1] - d01 * u00, w1 = B[4] - d01 * u10, w2 = B[7] - d01 * u20;
  float n1sq = w0 * w0 + w1 * w1 + w2 * w2;
  float n1 = sqrt_pos_(n1sq);
  float u01, u11, u21;
  if (n1 > 1e-12f * n0 && n1 > 1e-30f) { float r = rsqrt_(n1sq); u01 = w0 * r; u11 = w1 * r; u21 = w2 * r; }
  else {
    // rank <= 1: any unit vector orthogonal to u0
    float ax = fabsf(u00), ay = fabsf(u10), az = fabsf(u20);
    float ex = (ax <= ay && ax <= az) ? 1.0f : 0.0f;
    float ey = (ex == 0.0f && ay <= az) ? 1.0f : 0.0f;
    float ez = (ex == 0.0f && ey == 0.0f) ? 1.0f : 0.0f;
    float dd = ex * u00 + ey * u10 + ez * u20;
    w0 = ex - dd * u00; w1 = ey - dd * u10; w2 = ez - dd * u20;
    float r = rsqrt_(w0 * w0 + w1 * w1 + w2 * w2);
    u01 = w0 * r; u11 = w1 * r; u21 = w2 * r;
  }
  // u2 = u0 x u1
  float u02 = u10 * u21 - u20 * u11;
  float u12 = u20 * u01 - u00 * u21;
  float u22 = u00 * u11 - u10 * u01;
  U[0] = u00; U[1] = u01; U[2] = u02;
  U[3] = u10; U[4] = u11; U[5] = u12;
  U[6] = u20; U[7] = u21; U[8] = u22;
  sig[0] = n0;
  sig[1] = u01 * B[1] + u11 * B[4] + u21 * B[7];
  sig[2] = u02 * B[2] + u12 * B[5] + u22 * B[8];
}

template <int D> MPM_HD void svd(const float* F, float* U, float* sig, float* V);
template <> MPM_HD void svd<2>(const float* F, float* U, float* sig, float* V) { svd2(F, U, sig, V); }
template <> MPM_HD void svd<3>(const float* F, float* U, float* sig, float* V) { svd3(F, U, sig, V); }

template <int D> MPM_HD float det(const float* A);
template <> MPM_HD float det<2>(const float* A) { return A[0] * A[3] - A[1] * A[2]; }
template <> MPM_HD float det<3>(const float* A) {
  return A[0] * (A[4] * A[8] - A[5] * A[7]) - A[1] * (A[3] * A[8] - A[5] * A[6]) +
         A[2] * (A[3] * A[7] - A[4] * A[6]);
}


// ---------------------------------------------------------------------------
// Polar rotation R = U V^T of a 3x3 F with det F > 0 by Newton's iteration
// R <- (R + R^-T)/2 (quadratic convergence; singular values s -> (s + 1/s)/2).
// ELASTIC / STATIONARY particles need only R and J = det F (SURVEY Appendix B:
// sigma/new_sigma == 1 exactly, so Jp and F are untouched), which replaces the
// ~2500-instruction Jacobi SVD ncu measured on that path.  Returns false when F
// is (near) singular, inverted or the iteration does not settle: the caller then
// takes the SVD path, whose convention handles reflections.
// ---------------------------------------------------------------------------
MPM_HD bool polar3_newton(const float* F, float* R, float& J) {
  J = F[0] * (F[4] * F[8] - F[5] * F[7]) - F[1] * (F[3] * F[8] - F[5] * F[6]) + F[2] * (F[3] * F[7] - F[4] * F[6]);
  float nf = 0.0f;
#pragma unroll
  for (int i = 0; i < 9; ++i) { R[i] = F[i]; nf += F[i] * F[i]; }
  if (!(J > 1e-4f * nf * sqrt_pos_(nf) * 0.19245f)) return false;   // det <= 1e-4 * (||F||_F/sqrt3)^3
  bool ok = false;
#pragma unroll 1
  for (int it = 0; it < 12; ++it) {
    float c[9];
    c[0] = R[4] * R[8] - R[5] * R[7]; c[1] = R[5] * R[6] - R[3] * R[8]; c[2] = R[3] * R[7] - R[4] * R[6];
    c[3] = R[2] * R[7] - R[1] * R[8]; c[4] = R[0] * R[8] - R[2] * R[6]; c[5] = R[1] * R[6] - R[0] * R[7];
    c[6] = R[1] * R[5] - R[2] * R[4]; c[7] = R[2] * R[3] - R[0] * R[5]; c[8] = R[0] * R[4] - R[1] * R[3];
    float d = R[0] * c[0] + R[1] * c[1] + R[2] * c[2];
    float h = fdiv_(0.5f, d);
    float delta = 0.0f;
#pragma unroll
    for (int i = 0; i < 9; ++i) {
      float rn = 0.5f * R[i] + h * c[i];     // cofactor / det = R^-T
      delta = fmaxf(delta, fabsf(rn - R[i]));
      R[i] = rn;
    }
    if (delta < 3e-4f) {                      // the next step squares the error: one more, then stop
      if (ok || delta < 1e-7f) { ok = true; break; }
      ok = true;
    }
  }
  return ok;
}

// C = A * B
template <int D> MPM_HD void matmul(const float* A, const float* B, float* C) {
#pragma unroll
  for (int i = 0; i < D; ++i)
#pragma unroll
    for (int j = 0; j < D; ++j) {
      float s = 0.0f;
#pragma unroll
      for (int k = 0; k < D; ++k) s += A[i * D + k] * B[k * D + j];
      C[i * D + j] = s;
    }
}
// C = A * B^T
template <int D> MPM_HD void matmul_nt(const float* A, const float* B, float* C) {
#pragma unroll
  for (int i = 0; i < D; ++i)
#pragma unroll
    for (int j = 0; j < D; ++j) {
      float s = 0.0f;
#pragma unroll
      for (int k = 0; k < D; ++k) s += A[i * D + k] * B[j * D + k];
      C[i * D + j] = s;
    }
}
// C = U diag(s) V^T
template <int D> MPM_HD void usvt(const float* U, const float* s, const float* V, float* C) {
#pragma unroll
  for (int i = 0; i < D; ++i)
#pragma unroll
    for (int j = 0; j < D; ++j) {
      float a = 0.0f;
#pragma unroll
      for (int k = 0; k < D; ++k) a += U[i * D + k] * s[k] * V[j * D + k];
      C[i * D + j] = a;
    }
}

// C = U diag(s) U^T (symmetric)
template <int D> MPM_HD void udut(const float* U, const float* s, float* C) {
  float Us[D * D];
#pragma unroll
  for (int i = 0; i < D; ++i)
#pragma unroll
    for (int k = 0; k < D; ++k) Us[i * D + k] = U[i * D + k] * s[k];
#pragma unroll
  for (int i = 0; i < D; ++i)
#pragma unroll
    for (int j = i; j < D; ++j) {
      float a = 0.0f;
#pragma unroll
      for (int k = 0; k < D; ++k) a += Us[i * D + k] * U[j * D + k];
      C[i * D + j] = a;
      C[j * D + i] = a;
    }
}

// ---------------------------------------------------------------------------
// sand_projection (engine/mpm_solver.py:321-342).  sig in/out, Jp in/out.
// ---------------------------------------------------------------------------
template <int D> MPM_HD void sand_projection(const Consts& K, float* sig, float& Jp) {
  float eps[D], tr = 0.0f;
#pragma unroll
  for (int i = 0; i < D; ++i) { eps[i] = logf(fmaxf(fabsf(sig[i]), 1e-4f)); tr += eps[i]; }
  tr += Jp;
  float nrm = 0.0f, eh[D];
#pragma unroll
  for (int i = 0; i < D; ++i) { eh[i] = eps[i] - tr / (float)D; nrm += eh[i] * eh[i]; }
  nrm = sqrtf(nrm) + 1e-20f;
  if (tr >= 0.0f) {
    Jp = tr;
#pragma unroll
    for (int i = 0; i < D; ++i) sig[i] = 1.0f;
  } else {
    Jp = 0.0f;
    float dg = nrm + K.sand_coef * tr * K.alpha;
    float f = fmaxf(0.0f, dg) / nrm;
#pragma unroll
    for (int i = 0; i < D; ++i) sig[i] = expf(eps[i] - f * eh[i]);
  }
}

// ---------------------------------------------------------------------------
// The per-particle part of p2g (engine/mpm_solver.py:506-574), in three pieces (a second, compacted pass over the
// particles that need the full SVD was measured and dropped; the split keeps the fast cases readable):
//   trial_F               F_trial = (I + dt C) F_in                                  (:507-513)
//   particle_update_fast  every case that does NOT need the SVD; returns false otherwise
//   particle_update_svd   the general path (ti.svd, plasticity, stress)           (:525-566)
// particle_update() chains them (every P2G kernel, the host harness).
//   in : F (stored), C, Jp, material, dt
//   out: F (new stored), Jp (new), affine = stress + mass*C, mass
//
// Which cases avoid the SVD, and why the results are the reference's up to round-off:
//   WATER                only J = det F.
//   ELASTIC, STATIONARY  sigma is not clamped, Jp *= 1: only R = U V^T (Newton polar) and J = det F.
//   SNOW                 if every singular value already lies inside the clamp interval [1 - 2.5e-2, 1 + 4.5e-3]
//                        the clamp is the identity: F and Jp are unchanged and the stress is the elastic formula
//                        with the hardened mu, lambda.  The test needs no decomposition: sigma_i in (lo, hi) for all
//                        i  <=>  F^T F - lo^2 I and hi^2 I - F^T F are positive definite (Sylvester's criterion).
//                        A value within round-off of a bound may be classified either way; both branches then
//                        agree to round-off (the clamp moves it by ~1e-7).
//   SAND                 sand_projection only needs tr = sum_i log sigma_i + Jp = log det F + Jp to pick its case;
//                        for tr >= 0 ("expanding", :331-333) Sigma' = I: F' = U V^T = R, Jp' = tr, and the stress
//                        (all log sigma'_i = 0) vanishes.  At tr = 0 both cases give Sigma' = I and Jp' = 0, so a
//                        round-off difference in tr is harmless.
// ---------------------------------------------------------------------------
template <int D>
MPM_HD void trial_F(const Consts& K, float dt, int material, const float* F, const float* C, float Jp, float* Fn) {
  constexpr int DD = D * D;
  const bool fused = K.g2p2g != 0;   // [g2p2g] differences, SURVEY Appendix D-1
  float Fin[DD];
  if (material == WATER && !fused) {                         // :508-511 ([g2p2g] keeps the stored F, :414)
#pragma unroll
    for (int i = 0; i < DD; ++i) Fin[i] = 0.0f;
#pragma unroll
    for (int i = 0; i < D; ++i) Fin[i * D + i] = 1.0f;
    if (K.support_plasticity) Fin[0] = Jp;
  } else {
#pragma unroll
    for (int i = 0; i < DD; ++i) Fin[i] = F[i];
  }
  float A[DD];                                               // I + dt*C  (:513)
#pragma unroll
  for (int i = 0; i < DD; ++i) A[i] = dt * C[i];
#pragma unroll
  for (int i = 0; i < D; ++i) A[i * D + i] += 1.0f;
  matmul<D>(A, Fin, Fn);
  if (fused && K.clamp_F) {                                  // [g2p2g] :415-416
#pragma unroll
    for (int i = 0; i < DD; ++i) Fn[i] = fmaxf(-4.0f, fminf(4.0f, Fn[i]));
    if constexpr (D == 3) round_F9(Fn);                      // the store to the 16-bit field (3D packed storage)
  }
}

// hardening (:515-524): mu, lambda for this particle
MPM_HD void lame(const Consts& K, int material, float Jp, float& mu, float& la) {
  const bool fused = K.g2p2g != 0;
  float h = 1.0f;                                            // :515-521 ([g2p2g] hardens water too, :419-421)
  if (K.support_plasticity && (material != WATER || fused)) h = expf(10.0f * (1.0f - Jp));
  if (material == ELASTIC) h = 0.3f;
  mu = K.mu_0 * h;
  la = K.lambda_0 * h;
  if (material == WATER) mu = 0.0f;
}

template <int D> MPM_HD void finish_affine(const Consts& K, float dt, const float* stress, const float* C, float mass,
                                           float* affine) {
  const float scale = -dt * K.p_vol * 4.0f * K.inv_dx2;      // :569
#pragma unroll
  for (int i = 0; i < D * D; ++i) affine[i] = scale * stress[i] + mass * C[i];   // :574
}

// symmetric 3x3 (a00 a01 a02 / a11 a12 / a22) positive definite? (leading principal minors > 0)
MPM_HD bool spd3(float a00, float a01, float a02, float a11, float a12, float a22) {
  const float m2 = a00 * a11 - a01 * a01;
  const float d = a00 * (a11 * a22 - a12 * a12) - a01 * (a01 * a22 - a12 * a02) + a02 * (a01 * a12 - a11 * a02);
  return a00 > 0.0f && m2 > 0.0f && d > 0.0f;
}

// Fn: in = trial F, out = new stored F (only when true is returned)
template <int D>
MPM_HD bool particle_update_fast(const Consts& K, float dt, int material, float* Fn, const float* C, float& Jp,
                                 float* affine, float& mass) {
  constexpr int DD = D * D;
  const bool fused = K.g2p2g != 0;
  float mu, la;
  lame(K, material, Jp, mu, la);
  float stress[DD];
  mass = K.p_mass;
  if (material == WATER) {
    // mu = 0: only J = prod(sigma) = det F is needed.
    float J = det<D>(Fn);
#pragma unroll
    for (int i = 0; i < DD; ++i) Fn[i] = 0.0f;
#pragma unroll
    for (int i = 0; i < D; ++i) Fn[i * D + i] = 1.0f;
    Fn[0] = J;                                               // :537-542
    if (K.support_plasticity && !fused) Jp = J;              // [g2p2g] does not reset Jp (:440-444)
    float p = la * J * (J - 1.0f);
#pragma unroll
    for (int i = 0; i < DD; ++i) stress[i] = 0.0f;
#pragma unroll
    for (int i = 0; i < D; ++i) stress[i * D + i] = p;
    if (!fused) mass *= K.water_density;                     // :571-573 ([g2p2g]: p_mass, :472)
    finish_affine<D>(K, dt, stress, C, mass, affine);
    return true;
  }
  if constexpr (D == 3) {
    float Rp[DD], Jpol = 1.0f;
    bool elastic_like = material == ELASTIC || material == STATIONARY;
    if (material == SNOW) {
      // clamp interval of :529-531 squared, as f32 constants
      const float lo = 1.0f - 2.5e-2f, hi = 1.0f + 4.5e-3f;
      const float lo2 = lo * lo, hi2 = hi * hi;
      const float a00 = Fn[0] * Fn[0] + Fn[3] * Fn[3] + Fn[6] * Fn[6];
      const float a11 = Fn[1] * Fn[1] + Fn[4] * Fn[4] + Fn[7] * Fn[7];
      const float a22 = Fn[2] * Fn[2] + Fn[5] * Fn[5] + Fn[8] * Fn[8];
      const float a01 = Fn[0] * Fn[1] + Fn[3] * Fn[4] + Fn[6] * Fn[7];
      const float a02 = Fn[0] * Fn[2] + Fn[3] * Fn[5] + Fn[6] * Fn[8];
      const float a12 = Fn[1] * Fn[2] + Fn[4] * Fn[5] + Fn[7] * Fn[8];
      elastic_like = spd3(a00 - lo2, a01, a02, a11 - lo2, a12, a22 - lo2) &&
                     spd3(hi2 - a00, -a01, -a02, hi2 - a11, -a12, hi2 - a22);
    }
    if (elastic_like) {
      if (!polar3_newton(Fn, Rp, Jpol)) return false;        // inverted / singular: the SVD convention decides
      float T[DD];
#pragma unroll
      for (int i = 0; i < DD; ++i) T[i] = 2.0f * mu * (Fn[i] - Rp[i]);
      matmul_nt<D>(T, Fn, stress);                           // :550
      float p = la * Jpol * (Jpol - 1.0f);
#pragma unroll
      for (int i = 0; i < D; ++i) stress[i * D + i] += p;
      finish_affine<D>(K, dt, stress, C, mass, affine);
      return true;
    }
    if (material == SAND && K.support_plasticity) {
      const float J = det<D>(Fn);
      float n2 = 0.0f;
#pragma unroll
      for (int i = 0; i < DD; ++i) n2 += Fn[i] * Fn[i];
      // (J > 1e-2 and sum sigma^2 < 12 bound every sigma away from the 1e-4 floor of :326)
      if (!(J > 1e-2f) || !(n2 < 12.0f)) return false;
      const float tr = logf(J) + Jp;
      if (!(tr >= 0.0f)) return false;
      if (!polar3_newton(Fn, Rp, Jpol)) return false;
      Jp = tr;                                               // :331-333
#pragma unroll
      for (int i = 0; i < DD; ++i) { Fn[i] = Rp[i]; stress[i] = 0.0f; }
      finish_affine<D>(K, dt, stress, C, mass, affine);
      return true;
    }
  }
  return false;
}

// the general path: Fn in = trial F, out = new stored F
template <int D>
MPM_HD void particle_update_svd(const Consts& K, float dt, int material, float* Fn, const float* C, float& Jp,
                                float* affine, float& mass) {
  constexpr int DD = D * D;
  float mu, la;
  lame(K, material, Jp, mu, la);
  float stress[DD];
  mass = K.p_mass;
  float U[DD], V[DD], sig[D];
  svd<D>(Fn, U, sig, V);                                     // :525
  if (material != SAND) {
    float J = 1.0f;                                          // :527-536
    bool clamped = false;
#pragma unroll
    for (int d = 0; d < D; ++d) {
      float ns = sig[d];
      if (material == SNOW) ns = fminf(fmaxf(sig[d], 1.0f - 2.5e-2f), 1.0f + 4.5e-3f);
      if (K.support_plasticity) Jp *= sig[d] / ns;
      clamped = clamped || (ns != sig[d]);
      sig[d] = ns;
      J *= ns;
    }
    // :543-545 rebuilds F = U sig V^T for every snow particle; without a clamp that is F itself
    if (material == SNOW && clamped) usvt<D>(U, sig, V, Fn);
    // :547-551  2 mu (F - U V^T) F^T + la J (J - 1) I with F = U sig V^T:
    //   (F - R) F^T = U (sig - 1) V^T V sig U^T = U diag(sig (sig - 1)) U^T -- no V, no F - R cancellation
    const float p = la * J * (J - 1.0f);
    float dg[D];
#pragma unroll
    for (int d = 0; d < D; ++d) dg[d] = 2.0f * mu * sig[d] * (sig[d] - 1.0f) + p;
    udut<D>(U, dg, stress);
  } else if (K.support_plasticity) {                         // :553-566
    sand_projection<D>(K, sig, Jp);
    usvt<D>(U, sig, V, Fn);
    // center_i = (2 mu_0 log sig_i + lambda_0 sum log sig) / sig_i; U center V^T F^T = U diag(center sig) U^T
    float ls[D], lsum = 0.0f;
#pragma unroll
    for (int i = 0; i < D; ++i) { ls[i] = logf(sig[i]); lsum += ls[i]; }
    float dg[D];
#pragma unroll
    for (int i = 0; i < D; ++i) dg[i] = 2.0f * K.mu_0 * ls[i] + K.lambda_0 * lsum;
    udut<D>(U, dg, stress);
  } else {
#pragma unroll
    for (int i = 0; i < DD; ++i) stress[i] = 0.0f;
  }
  finish_affine<D>(K, dt, stress, C, mass, affine);
}

template <int D>
MPM_HD void particle_update(const Consts& K, float dt, int material, float* F, const float* C,
                            float& Jp, float* affine, float& mass) {
  constexpr int DD = D * D;
  float Fn[DD];
  trial_F<D>(K, dt, material, F, C, Jp, Fn);
  if (!particle_update_fast<D>(K, dt, material, Fn, C, Jp, affine, mass))
    particle_update_svd<D>(K, dt, material, Fn, C, Jp, affine, mass);
#pragma unroll
  for (int i = 0; i < DD; ++i) F[i] = Fn[i];                 // :567
}

}  // namespace mpm
