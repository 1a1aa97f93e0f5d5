// use_g2p2g=True: the fused substep of /root/reference/engine/mpm_solver.py:363-485 (driver :773-787).
//
// Reference kernel, per particle: gather v and C from the INPUT grid (last substep's output, normalised and
// boundary-processed) at the old position, advect, recompute base / fx / weights at the NEW position, F <- (I + dt C) F
// with the C that was just gathered (a register value, never stored), constitutive model, scatter to the OUTPUT grid.
//
// Here the substep is
//   k_g2p2g_keys     light pass over last substep's particle blocks: gather v only, x' = x + dt v, sort key of x' and
//                    the block flags of the output grid (8 + 4 B per particle; no state is written)
//   binning          (mpm_bin.cuh) on those keys: perm, particle blocks and grid blocks of the OUTPUT grid
//   k_g2p2g          ONE kernel per substep over the NEW blocks: stage the (LEAF+4)^D node tile of the input grid
//                    (a particle's old base lies within one cell of its new one: the (-1, 2) range hint of :406-412),
//                    gather v and C at the old position -- v with the SAME arithmetic as the key pass, so x' and its
//                    cell are reproduced bit for bit --, advect, F update with C in registers, constitutive model,
//                    new x, v, F, Jp to the sorted slot, scatter into the output grid.  C never touches memory:
//                    68 B read + 64 B written per particle (132 B) + 12 B of the key pass.
//   k_grid_op        normalise, gravity, velocity clamp, colliders on the output grid
// Two grids and two block tables ping-pong (num_grids = 2, :152).  The differences of the reference's fused kernel
// from its split pair (Appendix D-1) are all here: particles added since the last substep skip the gather (v kept,
// C = 0, :396-399); STATIONARY particles keep x and v but use the gathered C (:401-403, 414); hardening applies to
// WATER, which keeps its stored F and does not reset Jp; no water_density; F clamped to +-4 with quant (mpm_math.cuh,
// Consts::g2p2g / clamp_F); grid velocities clamped to +-dx*cfl/dt (k_grid_op).
#pragma once
#include "mpm_p2g3.cuh"

namespace mpm {

// sort key of a position + flags of its leaf block and of the blocks its stencil touches (as k_bin_keys)
template <int D>
__device__ __forceinline__ uint32_t key_and_flags(const float* x, float inv_dx, const KeyLayout& L, int* __restrict__ flags,
                                                  int nlin, Status* st) {
  using G = Geo<D>;
  uint32_t lin = 0, cell = 0, sp = 0;
  bool bad = false;
#pragma unroll
  for (int d = 0; d < D; ++d) {
    const int g = base_index(x[d], inv_dx) + L.half;
    int rel = (g >> G::LOG_LEAF) - L.ob[d];
    if (rel < 0 || rel > L.eb[d] - 2) { bad = true; rel = min(max(rel, 0), L.eb[d] - 2); }
    lin = lin * (uint32_t)L.eb[d] + (uint32_t)rel;
    const uint32_t lc = (uint32_t)(g & (G::LEAF - 1));
    cell = (cell << G::LOG_LEAF) | lc;
    sp |= (lc >= (uint32_t)(G::LEAF - 2)) ? (1u << d) : 0u;
  }
  if (bad) { atomicOr(&st->err, ERR_BBOX); return INVALID_KEY; }
  if (flags[lin] == 0) flags[lin] = 1;
  int* gf = flags + nlin;
#pragma unroll
  for (uint32_t o = 0; o < (uint32_t)G::NO; ++o)
    if ((o & ~sp) == 0) {
      const int t = (int)lin + oct_delta_l<D>(L, (int)o);
      if (gf[t] == 0) gf[t] = 1;
    }
  return (lin << G::CB) | cell;
}

// ---- key pass over LAST substep's particle blocks (storage order = that substep's sorted order, perm = identity)
template <int D, bool Q>
__global__ void __launch_bounds__(128) k_g2p2g_keys(FusedArgs<D> a, int npb_old) {
  using G = Geo<D>;
  using P = PStore<D, Q>;
  __shared__ float4 tile[G::TN];
  pdl_enter();
  if (a.s.st->err) return;
  const int tid = threadIdx.x;
  if (npb_old < 0) npb_old = a.s.st->npb;      // inside a batch: the block count of the substep enqueued just before
  for (int b = blockIdx.x; b < npb_old; b += gridDim.x) {
    __syncthreads();
    for (int n = tid; n < G::TN; n += blockDim.x) {
      int oct, cell;
      tile_node<D>(n, oct, cell);
      const int slot = a.s.pb_nbr[b * G::NO + oct];
      tile[n] = slot >= 0 ? a.grid_in[(size_t)slot * G::CELLS + cell] : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    int org[D];
    {
      int rel[D];
      key_to_rel<D>(a.tin.L, a.s.pb_key[b], rel);
#pragma unroll
      for (int d = 0; d < D; ++d) org[d] = (rel[d] + a.tin.L.ob[d]) << G::LOG_LEAF;
    }
    __syncthreads();
    const int start = a.s.pb_start[b], end = a.s.pb_start[b + 1];
    for (int s = start + tid; s < end; s += blockDim.x) {
      float x[D], fx[D], nv[D];
      int l[D];
      P::load_x(a.s.src, s, x);
#pragma unroll
      for (int d = 0; d < D; ++d) {
        const int base = base_index(x[d], a.s.K.inv_dx);
        fx[d] = __fsub_rn(__fmul_rn(x[d], a.s.K.inv_dx), (float)base);
        l[d] = min(max(base + a.tin.L.half - org[d], 0), G::LEAF - 1);
      }
      const uint32_t mat = tag_mat(__ldg(a.s.src + P::w(P::TAG, s)));
      gather_vC<D, false>(tile, G::T, l, fx, a.s.K.four_inv_dx, nv, nullptr);
      if (mat != (uint32_t)STATIONARY) {
        P::round_v(nv);                                                                   // (packed storage: a store rounds)
#pragma unroll
        for (int d = 0; d < D; ++d) x[d] = __fadd_rn(x[d], __fmul_rn(a.s.dt, nv[d]));     // :401-403
        P::round_x(x);
      }
      a.keys[s] = key_and_flags<D>(x, a.s.K.inv_dx, a.s.L, a.flags, a.nlin, a.s.st);
    }
  }
}
// rows added since the last substep (all rows on the first one): no gather, x' = x + dt v with the seeded v (:396-399)
template <int D, bool Q>
__global__ void k_g2p2g_keys_tail(FusedArgs<D> a, int r0, int n) {
  using P = PStore<D, Q>;
  pdl_enter();
  if (a.s.st->err) return;
  for (int s = r0 + blockIdx.x * blockDim.x + threadIdx.x; s < n; s += gridDim.x * blockDim.x) {
    float x[D], v[D];
    const uint32_t mat = tag_mat(__ldg(a.s.src + P::w(P::TAG, s)));
    P::load_x(a.s.src, s, x);
    P::load_v(a.s.src, s, v);
    if (mat != (uint32_t)STATIONARY) {
#pragma unroll
      for (int d = 0; d < D; ++d) x[d] = __fadd_rn(x[d], __fmul_rn(a.s.dt, v[d]));
      P::round_x(x);
    }
    a.keys[s] = key_and_flags<D>(x, a.s.K.inv_dx, a.s.L, a.flags, a.nlin, a.s.st);
  }
}

// ---- the fused kernel, general form (2D and 3D): one CTA per NEW particle block, node tile of the output grid
// accumulated with shared-memory atomics (the cell-owner 3D variant is k_p2g3<.., G2P2G> in mpm_p2g3.cuh)
template <int D>
__global__ void __launch_bounds__(256) k_g2p2g(FusedArgs<D> a) {
  using G = Geo<D>;
  using FL = Fld<D>;
  constexpr int TW = G::LEAF + 4, TWN = D == 3 ? TW * TW * TW : TW * TW;
  __shared__ float4 gin[TWN];        // input-grid velocities around the block, one cell wider on every side
  __shared__ float4 tile[G::TN];     // output-grid accumulators
  __shared__ int s_b;
  pdl_enter();
  const SubstepArgs<D>& A = a.s;
  if (A.st->err) return;
  const int npb = A.st->npb;
  const int tid = threadIdx.x;
  float vmax = 0.0f;
  int lo[D], hi[D];
#pragma unroll
  for (int d = 0; d < D; ++d) { lo[d] = INT_MAX; hi[d] = INT_MIN; }
  for (;;) {
    __syncthreads();
    if (tid == 0) s_b = atomicAdd(&A.st->work_p2g, 1);
    __syncthreads();
    const int b = s_b;
    if (b >= npb) break;
    int org[D];   // absolute cell coordinate (offset by half) of the block origin
    {
      int rel[D];
      key_to_rel<D>(A.L, A.pb_key[b], rel);
#pragma unroll
      for (int d = 0; d < D; ++d) org[d] = (rel[d] + A.L.ob[d]) << G::LOG_LEAF;
    }
    for (int n = tid; n < G::TN; n += blockDim.x) tile[n] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int n = tid; n < TWN; n += blockDim.x) {
      float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
      if (a.grid_in) {
        int c[D], m = n, babs[D], cell = 0;
#pragma unroll
        for (int d = D - 1; d >= 0; --d) { c[d] = org[d] - 1 + m % TW; m /= TW; }
#pragma unroll
        for (int d = 0; d < D; ++d) {
          babs[d] = c[d] >> G::LOG_LEAF;
          cell = (cell << G::LOG_LEAF) | (c[d] & (G::LEAF - 1));
        }
        const int slot = table_slot<D>(a.tin, babs);
        if (slot >= 0) g = a.grid_in[(size_t)slot * G::CELLS + cell];
      }
      gin[n] = g;
    }
    __syncthreads();
    const int start = A.pb_start[b], end = A.pb_start[b + 1];
    for (int s = start + tid; s < end; s += blockDim.x) {
      const uint32_t p = A.perm[s];
      float x[D], v[D], fx[D], C[D * D];
      int l[D];
#pragma unroll
      for (int d = 0; d < D; ++d) {
        x[d] = ldf<D>(A.src, FL::X + d, p);
        v[d] = ldf<D>(A.src, FL::V + d, p);
      }
      float F[D * D];
#pragma unroll
      for (int i = 0; i < D * D; ++i) F[i] = ldf<D>(A.src, FL::F + i, p);
      float Jp = ldf<D>(A.src, FL::JP, p);
      const uint32_t tag = ldu<D>(A.src, FL::TAG, p), mat = tag_mat(tag);
      // ---- G2P half at the OLD position (:376-403)
      if ((int)p < a.n_old) {
#pragma unroll
        for (int d = 0; d < D; ++d) {
          const int base = base_index(x[d], A.K.inv_dx);
          fx[d] = __fsub_rn(__fmul_rn(x[d], A.K.inv_dx), (float)base);
          l[d] = min(max(base + A.L.half - (org[d] - 1), 0), TW - 3);
        }
        float nv[D];
        gather_vC<D, true>(gin, TW, l, fx, A.K.four_inv_dx, nv, C);
        if (mat != (uint32_t)STATIONARY) {
#pragma unroll
          for (int d = 0; d < D; ++d) v[d] = nv[d];
        }
      } else {
#pragma unroll
        for (int i = 0; i < D * D; ++i) C[i] = 0.0f;                                   // :396-399
      }
      if (mat != (uint32_t)STATIONARY) {
#pragma unroll
        for (int d = 0; d < D; ++d) x[d] = __fadd_rn(x[d], __fmul_rn(A.dt, v[d]));    // :401-403
      }
      // ---- P2G half at the NEW position (:405-483)
      float aff[D * D], mass;
      particle_update<D>(A.K, A.dt, (int)mat, F, C, Jp, aff, mass);
#pragma unroll
      for (int d = 0; d < D; ++d) {
        stf<D>(A.dst, FL::X + d, s, x[d]);
        stf<D>(A.dst, FL::V + d, s, v[d]);
        vmax = fmaxf(vmax, fabsf(v[d]));
      }
#pragma unroll
      for (int i = 0; i < D * D; ++i) stf<D>(A.dst, FL::F + i, s, F[i]);
      stf<D>(A.dst, FL::JP, s, Jp);
      stu<D>(A.dst, FL::TAG, s, tag);
      float w[3][D], mv[D];
#pragma unroll
      for (int d = 0; d < D; ++d) {
        const int base = base_index(x[d], A.K.inv_dx);
        lo[d] = min(lo[d], base); hi[d] = max(hi[d], base);
        fx[d] = x[d] * A.K.inv_dx - (float)base;
        l[d] = min(max(base + A.L.half - org[d], 0), G::LEAF - 1);
        w[0][d] = 0.5f * (1.5f - fx[d]) * (1.5f - fx[d]);
        w[1][d] = 0.75f - (fx[d] - 1.0f) * (fx[d] - 1.0f);
        w[2][d] = 0.5f * (fx[d] - 0.5f) * (fx[d] - 0.5f);
        mv[d] = mass * v[d];
      }
      if constexpr (D == 3) {
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
          for (int j = 0; j < 3; ++j)
#pragma unroll
            for (int k = 0; k < 3; ++k) {
              const float dp0 = ((float)i - fx[0]) * A.K.dx, dp1 = ((float)j - fx[1]) * A.K.dx,
                          dp2 = ((float)k - fx[2]) * A.K.dx;
              const float wt = w[i][0] * w[j][1] * w[k][2];
              float* t = reinterpret_cast<float*>(&tile[((l[0] + i) * G::T + (l[1] + j)) * G::T + (l[2] + k)]);
              atomicAdd(t + 0, wt * (mv[0] + aff[0] * dp0 + aff[1] * dp1 + aff[2] * dp2));
              atomicAdd(t + 1, wt * (mv[1] + aff[3] * dp0 + aff[4] * dp1 + aff[5] * dp2));
              atomicAdd(t + 2, wt * (mv[2] + aff[6] * dp0 + aff[7] * dp1 + aff[8] * dp2));
              atomicAdd(t + 3, wt * mass);
            }
      } else {
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
          for (int j = 0; j < 3; ++j) {
            const float dp0 = ((float)i - fx[0]) * A.K.dx, dp1 = ((float)j - fx[1]) * A.K.dx;
            const float wt = w[i][0] * w[j][1];
            float* t = reinterpret_cast<float*>(&tile[(l[0] + i) * G::T + (l[1] + j)]);
            atomicAdd(t + 0, wt * (mv[0] + aff[0] * dp0 + aff[1] * dp1));
            atomicAdd(t + 1, wt * (mv[1] + aff[2] * dp0 + aff[3] * dp1));
            atomicAdd(t + 2, wt * mass);
          }
      }
    }
    __syncthreads();
    for (int n = tid; n < G::TN; n += blockDim.x) {
      const float4 val = tile[n];
      const float m = (D == 3) ? val.w : val.z;
      if (m != 0.0f) {
        int oct, cell;
        tile_node<D>(n, oct, cell);
        const int slot = A.pb_nbr[b * G::NO + oct];
        if (slot >= 0) red_add_v4(A.grid + (size_t)slot * G::CELLS + cell, val);
      }
    }
  }
  // compute_max_velocity (:726-735) and the next bounding box, once per CTA
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    vmax = fmaxf(vmax, __shfl_xor_sync(0xffffffffu, vmax, o));
#pragma unroll
    for (int d = 0; d < D; ++d) {
      lo[d] = min(lo[d], __shfl_xor_sync(0xffffffffu, lo[d], o));
      hi[d] = max(hi[d], __shfl_xor_sync(0xffffffffu, hi[d], o));
    }
  }
  if ((tid & 31) == 0) {
    if (vmax != vmax) vmax = __int_as_float(0x7f800000);
    atomicMax(&A.st->maxv_bits, __float_as_uint(vmax));
#pragma unroll
    for (int d = 0; d < D; ++d)
      if (lo[d] <= hi[d]) { atomicMin(&A.st->bb_min[d], lo[d]); atomicMax(&A.st->bb_max[d], hi[d]); }
  }
}

// end of a fused substep: commit it (the host flips the live set by the number of committed substeps)
__global__ void k_fused_commit(Status* st) {
  pdl_enter();
  if (!st->err) {
    st->done += 1;
    if (st->maxv_bits > st->maxv_all) st->maxv_all = st->maxv_bits;
  }
}

}  // namespace mpm
