// Pieces shared by the fused g2p2g kernels (mpm_g2p2g.cuh: key pass and the general kernel; mpm_p2g3.cuh: the 3D
// cell-owner variant): the block table of the INPUT grid, the kernel arguments and the gather.
#pragma once
#include "mpm_kernels.cuh"

namespace mpm {

// dense block table of one substep (mpm_bin.cuh): [particle-block flags | grid-block flags] and their exclusive scan
struct GridTable {
  const int* flags;
  const int* fscan;
  int nlin;
  KeyLayout L;
};
// grid slot of an absolute leaf block in that table, or -1
template <int D> __device__ __forceinline__ int table_slot(const GridTable& T, const int* babs) {
  int rel[D];
#pragma unroll
  for (int d = 0; d < D; ++d) {
    rel[d] = babs[d] - T.L.ob[d];
    if (rel[d] < 0 || rel[d] >= T.L.eb[d]) return -1;
  }
  const int lin = (int)rel_to_key<D>(T.L, rel);
  return T.flags[T.nlin + lin] ? T.fscan[T.nlin + lin] - T.fscan[T.nlin] : -1;
}

template <int D> struct FusedArgs {
  SubstepArgs<D> s;        // src / dst state, perm, particle blocks, OUTPUT grid, status, layout, constants, dt
  const float4* grid_in;   // INPUT grid: last substep's output (velocities), null on the first substep
  GridTable tin;           // its block table
  int n_old;               // rows that existed when the input grid was scattered; rows >= n_old skip the gather
  uint32_t* keys;          // key pass: sort keys of the advected positions, by storage row
  int* flags;              // key pass: block flags of the output grid's table
  int nlin;
};

// Quadratic B-spline gather of v (and C = 4 inv_dx sum w v (x) (o - fx), :376-394) from a node tile of edge TW.
// Explicit fma order: the key pass and the fused kernel must produce the SAME v bit for bit.
template <int D, bool WITH_C>
__device__ __forceinline__ void gather_vC(const float4* __restrict__ tile, int TW, const int* l, const float* fx,
                                          float four_inv_dx, float* nv, float* nC) {
  float w[3][D];
#pragma unroll
  for (int d = 0; d < D; ++d) {      // (:379) every operation rounded on its own: no contraction may differ between callers
    const float t0 = __fsub_rn(1.5f, fx[d]), t1 = __fsub_rn(fx[d], 1.0f), t2 = __fsub_rn(fx[d], 0.5f);
    w[0][d] = __fmul_rn(0.5f, __fmul_rn(t0, t0));
    w[1][d] = __fsub_rn(0.75f, __fmul_rn(t1, t1));
    w[2][d] = __fmul_rn(0.5f, __fmul_rn(t2, t2));
  }
#pragma unroll
  for (int d = 0; d < D; ++d) nv[d] = 0.0f;
  if (WITH_C) {
#pragma unroll
    for (int i = 0; i < D * D; ++i) nC[i] = 0.0f;
  }
  if constexpr (D == 3) {
    // tensor-product evaluation: partial sums along z, then y, then x, of sum w g and of its first moments
    // (279 fused multiply-adds instead of 27 x 15); the v chain (A0 -> B00 -> nv) is the same with and without C
    float mw[3][3];
#pragma unroll
    for (int d = 0; d < 3; ++d)
#pragma unroll
      for (int i = 0; i < 3; ++i) mw[i][d] = __fmul_rn(w[i][d], __fsub_rn((float)i, fx[d]));
    float cx[3] = {0.f, 0.f, 0.f}, cy[3] = {0.f, 0.f, 0.f}, cz[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      float B00[3] = {0.f, 0.f, 0.f}, B10[3] = {0.f, 0.f, 0.f}, B01[3] = {0.f, 0.f, 0.f};
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        float A0[3] = {0.f, 0.f, 0.f}, A1[3] = {0.f, 0.f, 0.f};
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          const float4 g = tile[((l[0] + i) * TW + (l[1] + j)) * TW + (l[2] + k)];
          const float gv[3] = {g.x, g.y, g.z};
#pragma unroll
          for (int r = 0; r < 3; ++r) {
            A0[r] = __fmaf_rn(w[k][2], gv[r], A0[r]);
            if (WITH_C) A1[r] = __fmaf_rn(mw[k][2], gv[r], A1[r]);
          }
        }
#pragma unroll
        for (int r = 0; r < 3; ++r) {
          B00[r] = __fmaf_rn(w[j][1], A0[r], B00[r]);
          if (WITH_C) {
            B10[r] = __fmaf_rn(mw[j][1], A0[r], B10[r]);
            B01[r] = __fmaf_rn(w[j][1], A1[r], B01[r]);
          }
        }
      }
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        nv[r] = __fmaf_rn(w[i][0], B00[r], nv[r]);
        if (WITH_C) {
          cx[r] = __fmaf_rn(mw[i][0], B00[r], cx[r]);
          cy[r] = __fmaf_rn(w[i][0], B10[r], cy[r]);
          cz[r] = __fmaf_rn(w[i][0], B01[r], cz[r]);
        }
      }
    }
    if (WITH_C) {
#pragma unroll
      for (int r = 0; r < 3; ++r) { nC[r * 3 + 0] = cx[r]; nC[r * 3 + 1] = cy[r]; nC[r * 3 + 2] = cz[r]; }
    }
  } else {
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        const float4 g = tile[(l[0] + i) * TW + (l[1] + j)];
        const float wt = __fmul_rn(w[i][0], w[j][1]);
        const float gv[2] = {g.x, g.y};
#pragma unroll
        for (int r = 0; r < 2; ++r) nv[r] = __fmaf_rn(wt, gv[r], nv[r]);
        if (WITH_C) {
          const float dp[2] = {(float)i - fx[0], (float)j - fx[1]};
#pragma unroll
          for (int r = 0; r < 2; ++r) {
            const float wv = wt * gv[r];
#pragma unroll
            for (int c = 0; c < 2; ++c) nC[r * 2 + c] = fmaf(wv, dp[c], nC[r * 2 + c]);
          }
        }
      }
  }
  if (WITH_C) {
#pragma unroll
    for (int i = 0; i < D * D; ++i) nC[i] *= four_inv_dx;
  }
}

}  // namespace mpm
