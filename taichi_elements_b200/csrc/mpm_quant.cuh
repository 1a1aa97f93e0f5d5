// quant=True: bit-packed particle storage of /root/reference/engine/mpm_solver.py:106-114, 216-262 (3D):
//   x   3 x 21-bit signed fixed point, range +-2.0                  -> one 64-bit word   (:107-108, 221-223)
//   v   3 x 19-bit fractions sharing one 7-bit exponent             -> one 64-bit word   (:110-111, 224-226)
//   F   9 x 16-bit signed fixed point, range +-(F_bound + 0.1)      -> five 32-bit words (:113-114, 229-247; the spare
//       half word holds `material` in the reference -- here the material lives in the tag word)
// The encodings are Taichi's quantised types (ti.types.quant.fixed / quant.float with shared_exponent), which are NOT part
// of the reference tree [EXT]; they are restated here from their definitions:
//   fixed(bits, max_value), signed:  scale = max_value / 2^(bits-1);  q = clamp(round_half_away(x / scale), +-(2^(bits-1) - 1));
//                                    value = q * scale
//   float(exp=7, frac=19), shared exponent over the 3 components: e = floor(log2(max_d |v_d|)), clamped to [-64, 63] and
//                                    stored biased by 64; every fraction is a 19-bit two's-complement integer
//                                    m_d = clamp(round_half_away(v_d / 2^(e - 17)), +-(2^18 - 1)); value = m_d * 2^(e - 17)
//                                    (the largest component keeps 18 significant bits, smaller ones lose low bits)
// Written once as __host__ __device__ so that tests/host_harness.cpp checks them against the NumPy restatement in
// oracle/quant_oracle.py.
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#include "mpm_math.cuh"

namespace mpm {

MPM_HD uint32_t f2u_(float f) {
#ifdef __CUDA_ARCH__
  return __float_as_uint(f);
#else
  uint32_t u; memcpy(&u, &f, 4); return u;
#endif
}
MPM_HD float u2f_(uint32_t u) {
#ifdef __CUDA_ARCH__
  return __uint_as_float(u);
#else
  float f; memcpy(&f, &u, 4); return f;
#endif
}

// ---- x: 3 x 21 bits in two words
MPM_HD void encode_x3(const float* x, uint32_t* w) {
  const unsigned long long m = (1ull << QX_BITS) - 1ull;
  unsigned long long p = 0;
#pragma unroll
  for (int d = 0; d < 3; ++d) p |= ((unsigned long long)(unsigned)q_fixed(x[d], QX_MAX, QX_BITS) & m) << (QX_BITS * d);
  w[0] = (uint32_t)p;
  w[1] = (uint32_t)(p >> 32);
}
MPM_HD void decode_x3(const uint32_t* w, float* x) {
  const unsigned long long p = (unsigned long long)w[0] | ((unsigned long long)w[1] << 32);
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    const int q = (int)((uint32_t)(p >> (QX_BITS * d)) << (32 - QX_BITS)) >> (32 - QX_BITS);   // sign-extend 21 bits
    x[d] = dq_fixed(q, QX_MAX, QX_BITS);
  }
}
MPM_HD void round_x3(float* x) { uint32_t w[2]; encode_x3(x, w); decode_x3(w, x); }

// ---- v: 3 x 19-bit fractions + a shared 7-bit exponent in two words
MPM_HD void encode_v3(const float* v, uint32_t* w) {
  const float a = fmaxf(fabsf(v[0]), fmaxf(fabsf(v[1]), fabsf(v[2])));
  int e = -64;
  if (a > 0.0f && a == a) {
    // floor(log2 a) is the biased exponent field - 127 for a normal number; zero and denormals (field 0) lie below 2^-64
    // and clamp to -64 either way
    e = (int)(f2u_(a) >> 23) - 127;
    e = e < -64 ? -64 : (e > 63 ? 63 : e);
  }
  const float inv = u2f_((uint32_t)(17 - e + 127) << 23);      // 2^(17 - e), exponent in [-46, 81]
  const int lim = (1 << (QV_FRAC - 1)) - 1;
  const unsigned long long m = (1ull << QV_FRAC) - 1ull;
  unsigned long long p = 0;
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    const float t = v[d] * inv;
    int q = !(t > -(float)lim) ? -lim : (t > (float)lim ? lim : q_round(t));
    p |= ((unsigned long long)(unsigned)q & m) << (QV_FRAC * d);
  }
  p |= (unsigned long long)(unsigned)(e + 64) << (3 * QV_FRAC);
  w[0] = (uint32_t)p;
  w[1] = (uint32_t)(p >> 32);
}
MPM_HD void decode_v3(const uint32_t* w, float* v) {
  const unsigned long long p = (unsigned long long)w[0] | ((unsigned long long)w[1] << 32);
  const int e = (int)((p >> (3 * QV_FRAC)) & 127ull) - 64;
  const float s = u2f_((uint32_t)(e - 17 + 127) << 23);        // 2^(e - 17), exponent in [-81, 46]
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    const int q = (int)((uint32_t)(p >> (QV_FRAC * d)) << (32 - QV_FRAC)) >> (32 - QV_FRAC);
    v[d] = (float)q * s;
  }
}
MPM_HD void round_v3(float* v) { uint32_t w[2]; encode_v3(v, w); decode_v3(w, v); }

// ---- F: 9 x 16 bits in five words (row-major pairs as the reference places them, :229-247)
MPM_HD void encode_F9(const float* F, uint32_t* w) {
#pragma unroll
  for (int i = 0; i < 5; ++i) {
    const uint32_t lo = (uint32_t)q_fixed(F[2 * i], QF_MAX, QF_BITS) & 0xffffu;
    const uint32_t hi = i < 4 ? ((uint32_t)q_fixed(F[2 * i + 1], QF_MAX, QF_BITS) & 0xffffu) : 0u;
    w[i] = lo | (hi << 16);
  }
}
MPM_HD void decode_F9(const uint32_t* w, float* F) {
#pragma unroll
  for (int i = 0; i < 9; ++i) {
    const uint32_t h = (i & 1) ? (w[i >> 1] >> 16) : (w[i >> 1] & 0xffffu);
    F[i] = dq_fixed((int)(short)h, QF_MAX, QF_BITS);
  }
}
}  // namespace mpm
