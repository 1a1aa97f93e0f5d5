// P2G, cell-owner formulation, third revision (3D, sm_100a).
//
// Same algorithm as mpm_p2g.cuh (engine/mpm_solver.py:487-584: one thread per
// particle for the constitutive update, one thread per (cell, x-slice) for the
// scatter, accumulated in registers), re-organised around what ncu showed on the
// second revision: 14 CTA barriers per block (3.4 of ~6 resident warps per
// scheduler parked at a barrier), a 64-step rank loop on two warps, and an
// FFMA-per-scalar inner loop.
//   * payload is AoS, four float4 per particle (m*v|m and the three columns of
//     dx*A, each with one component of fx in its fourth lane), 16-byte columns
//     rotated by the particle index against bank conflicts: 4 LDS.128 instead of
//     16 LDS.32, and the scatter arithmetic runs on packed pairs
//     (fma.rn.f32x2 -> FFMA2)
//   * the accumulators of a thread are flushed ONCE per block into nine private
//     4x4x6 copies (one per (x, y) stencil offset) that alias the dead payload:
//     three conflict-free rounds over the z offset, no tile to clear; a node then
//     sums the <= 9 copies that touch it and issues one REDG.E.ADD.F32x4
//   * the next block's cell ranges and the descending-count cell order (counting
//     sort on a 32-bin shared histogram) are prepared by the first warp that runs
//     out of scatter work, off the critical path (double-buffered)
// Barriers per block: 6, of which only two can see unbalanced arrivals.
#pragma once
#include <type_traits>

#include "mpm_fused.cuh"

namespace mpm {

struct P2G3 {
  static constexpr int T = 192, SL = 3, NPT = 9, PS = 4;   // threads, x-slices, nodes per thread, float4 per particle
  static constexpr int CP = 16 * 6 + 1;                    // one flush copy: [cx][cy][z] float4, +1: the three
                                                           // x-slices of a cell must not share banks
  static constexpr int COPIES = 9 * CP;                    // float4
};

template <int CHUNK> constexpr size_t p2g3_smem_bytes() {
  static_assert(CHUNK * P2G3::PS >= P2G3::COPIES, "flush copies alias the payload");
  return (size_t)CHUNK * P2G3::PS * sizeof(float4);
}

__device__ __forceinline__ float2 f2(float x) { return make_float2(x, x); }

// One warp: cell ranges of block nb (relative to its first particle) and the cells in
// descending particle-count order (a warp's trip count in the scatter loop is the largest
// count among its lanes): counting sort over min(count, 31), order inside a bin arbitrary.
__device__ __forceinline__ void p2g3_prepare(const SubstepArgs<3>& a, int nb, int npb, int lane, int* cs, int* order,
                                             int* hist) {
  if (nb >= npb) return;
  const int ns = a.pb_start[nb];
  for (int c = lane; c <= 64; c += 32) cs[c] = a.cellstart[(size_t)nb * 64 + c] - ns;
  hist[lane] = 0;
  __syncwarp();
  const int m0 = min(cs[lane + 1] - cs[lane], 31), m1 = min(cs[lane + 33] - cs[lane + 32], 31);
  const int r0 = atomicAdd(&hist[m0], 1), r1 = atomicAdd(&hist[m1], 1);
  __syncwarp();
  const int h = hist[lane];
  int incl = h;                                   // cells whose bin is >= lane
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_down_sync(0xffffffffu, incl, o);
    if (lane + o < 32) incl += t;
  }
  __syncwarp();
  hist[lane] = incl - h;                          // cells in a higher bin come first
  __syncwarp();
  order[hist[m0] + r0] = lane;
  order[hist[m1] + r1] = lane + 32;
}

// Work-queue position -> particle block.  With slabs the blocks of the last and of the first block column come
// first: their tiles hold the nodes of the grid columns shared with the neighbours, which travel (as vector
// reductions into the neighbour's halo plane over NVLink) while the interior of the slab is still being scattered.
__device__ __forceinline__ int p2g3_block_of(int q, int npb, int n_hi) {
  return q < n_hi ? npb - n_hi + q : q - n_hi;
}
// tell the neighbours that every node of the shared columns has been sent for substep epoch + 1
__device__ __forceinline__ void p2g3_publish_halo(const CommBufs& cb) {
#pragma unroll
  for (int s = 0; s < 2; ++s)
    if (cb.flag_halo[s]) st_release_sys(cb.flag_halo[s], cb.epoch + 1u);
}

// FUSED: the multi-GPU variant (fused halo); the single-device kernel carries none of its state.
__device__ __forceinline__ const SubstepArgs<3>& p2g3_args(const SubstepArgs<3>& a) { return a; }
__device__ __forceinline__ const SubstepArgs<3>& p2g3_args(const FusedArgs<3>& a) { return a.s; }

// G2P2G: the use_g2p2g fused kernel (mpm_g2p2g.cuh): the constitutive phase first gathers v and C from the INPUT grid
// at the old position (an 8^3 node tile around the block, staged once per block) and advects; C stays in registers.
template <int CHUNK, int MINB, bool FUSED, bool G2P2G = false, bool Q = false>
__global__ void __launch_bounds__(P2G3::T, MINB)
k_p2g3(typename std::conditional<G2P2G, FusedArgs<3>, SubstepArgs<3>>::type arg) {
  constexpr int D = 3;
  const SubstepArgs<3>& a = p2g3_args(arg);
  using G = Geo<3>;
  using FL = Fld<3>;
  constexpr int T = P2G3::T, CH = CHUNK, PS = P2G3::PS;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float4* pay = reinterpret_cast<float4*>(smem_raw);      // [CH][PS]; the flush copies [27][64] alias it
  __shared__ int s_cs[2][G::CELLS + 1];
  __shared__ int s_order[2][G::CELLS];
  __shared__ int s_nbr[G::NO];
  __shared__ int s_b, s_next, s_ticket, s_ticket_q;
  __shared__ int s_hist[32];
  __shared__ float4 s_gin[G2P2G ? 512 : 1];   // use_g2p2g: velocities of the INPUT grid, 8^3 nodes around the block
  float vmax = 0.0f;                            // use_g2p2g: compute_max_velocity and the next bounding box
  int bb_lo[3] = {INT_MAX, INT_MAX, INT_MAX}, bb_hi[3] = {INT_MIN, INT_MIN, INT_MIN};
  pdl_enter();
  const int tid = threadIdx.x, lane = tid & 31;
  constexpr bool fused = FUSED;
  if (a.st->err) {
    // a neighbour must never wait for a message that will not come
    if (fused && blockIdx.x == 0 && tid == 0) p2g3_publish_halo(a.cb);
    return;
  }
  const int npb = a.st->npb;
  const int n_bhi = fused ? a.st->bnd_hi : 0, n_bnd = fused ? a.st->bnd_hi + a.st->bnd_lo : 0;
  if (fused && n_bnd == 0 && blockIdx.x == 0 && tid == 0) p2g3_publish_halo(a.cb);   // nothing to send
  const size_t cap = a.cap;
  const int sl = tid % P2G3::SL;
  const float slf = (float)sl;
  // quadratic B-spline weight of this thread's x-slice as a polynomial in fx (:505)
  const float k2 = sl == 1 ? -1.0f : 0.5f, k1 = sl == 0 ? -1.5f : (sl == 1 ? 2.0f : -0.5f),
              k0 = sl == 0 ? 1.125f : (sl == 1 ? -0.25f : 0.125f);

  if (tid == 0) { s_b = atomicAdd(&a.st->work_p2g, 1); s_ticket = 0; }
  __syncthreads();
  int qpos = s_b;                                           // position in the work queue
  int b = qpos < npb ? p2g3_block_of(qpos, npb, n_bhi) : npb;
  if (tid < 32) p2g3_prepare(a, b, npb, lane, s_cs[0], s_order[0], s_hist);
  int u = 0;
  __syncthreads();
  while (b < npb) {
    if (tid == 0) {                                         // claimed one block ahead
      const int qn = atomicAdd(&a.st->work_p2g, 1);
      s_next = qn < npb ? p2g3_block_of(qn, npb, n_bhi) : npb;
      s_ticket_q = qn;
    }
    const int start = a.pb_start[b], end = a.pb_start[b + 1];
    const int cnt = end - start;
    if (tid < G::NO) s_nbr[tid] = a.pb_nbr[b * G::NO + tid];
    const bool boundary = fused && qpos < n_bnd;            // fused halo: this block's tile touches a shared column
    int org[3] = {0, 0, 0};                                 // use_g2p2g: absolute cell (offset by half) of the block origin
    if constexpr (G2P2G) {
      int rel[3];
      key_to_rel<3>(a.L, a.pb_key[b], rel);
#pragma unroll
      for (int d = 0; d < 3; ++d) org[d] = (rel[d] + a.L.ob[d]) << G::LOG_LEAF;
      for (int n = tid; n < 512; n += T) {
        float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
        if (arg.grid_in) {
          const int c[3] = {org[0] - 1 + (n >> 6), org[1] - 1 + ((n >> 3) & 7), org[2] - 1 + (n & 7)};
          const int babs[3] = {c[0] >> 2, c[1] >> 2, c[2] >> 2};
          const int slot = table_slot<3>(arg.tin, babs);
          if (slot >= 0) g = arg.grid_in[(size_t)slot * G::CELLS + (((c[0] & 3) << 4) | ((c[1] & 3) << 2) | (c[2] & 3))];
        }
        s_gin[n] = g;
      }
      __syncthreads();
    }
    const int* cs = s_cs[u];
    const int cell = s_order[u][tid / P2G3::SL];
    const int c_lo = cs[cell], c_hi = cs[cell + 1];

    // a block with more than CH particles (rare) takes several passes, each complete down to
    // the global reductions, so the accumulators never live across a constitutive update
    for (int c0 = 0; c0 < cnt; c0 += CH) {
      const int cn = min(CH, cnt - c0);
      const bool last = c0 + CH >= cnt;
      if constexpr (G2P2G) {
      // ---- phase 1 (use_g2p2g): G2P half at the old position, advection, P2G half at the new one (ref :376-483)
      constexpr int NIT = (CH + T - 1) / T;
      uint32_t pq[NIT];
#pragma unroll
      for (int k = 0; k < NIT; ++k) pq[k] = (tid + k * T < cn) ? a.perm[start + c0 + tid + k * T] : 0u;
#pragma unroll 1
      for (int q = tid; q < cn; q += T) {
        const int s = start + c0 + q;
        const uint32_t p = pq[0];
#pragma unroll
        for (int k = 0; k + 1 < NIT; ++k) pq[k] = pq[k + 1];
        using P = PStore<3, Q ? 1 : 0>;                          // f32 words, or the packed storage of quant=True
        float x[D], v[D], fx[D], C[D * D];
        P::load_x(a.src, p, x);
        P::load_v(a.src, p, v);
        float F[D * D], aff[D * D], mass;
        P::load_F(a.src, p, F);
        float Jp = __uint_as_float(__ldg(a.src + P::w(P::JP, p)));
        const uint32_t tag = __ldg(a.src + P::w(P::TAG, p)), mat = tag_mat(tag);
        if ((int)p < arg.n_old) {
          int l[D];
#pragma unroll
          for (int d = 0; d < D; ++d) {
            const int base = base_index(x[d], a.K.inv_dx);
            fx[d] = __fsub_rn(__fmul_rn(x[d], a.K.inv_dx), (float)base);
            l[d] = min(max(base + a.L.half - (org[d] - 1), 0), 5);
          }
          float nv[D];
          gather_vC<D, true>(s_gin, 8, l, fx, a.K.four_inv_dx, nv, C);
          if (mat != (uint32_t)STATIONARY) {
#pragma unroll
            for (int d = 0; d < D; ++d) v[d] = nv[d];
          }
        } else {
#pragma unroll
          for (int i = 0; i < D * D; ++i) C[i] = 0.0f;                                 // :396-399
        }
        uint32_t qx[2], qv[2];                    // packed storage: the stored words
        P::round_v(v, qv);                        // (`self.v[p] = new_v` rounds, the advection reads it back)
        if (mat != (uint32_t)STATIONARY) {
#pragma unroll
          for (int d = 0; d < D; ++d) x[d] = __fadd_rn(x[d], __fmul_rn(a.dt, v[d]));  // :401-403
        }
        P::round_x(x, qx);                        // (and the P2G half reads the stored x, :406)
        particle_update<D>(a.K, a.dt, (int)mat, F, C, Jp, aff, mass);
        P::put_x(a.dst, s, x, qx);
        P::put_v(a.dst, s, v, qv);
        P::store_F(a.dst, s, F);
        a.dst[P::w(P::JP, s)] = __float_as_uint(Jp);
        a.dst[P::w(P::TAG, s)] = tag;
#pragma unroll
        for (int d = 0; d < D; ++d) {
          vmax = fmaxf(vmax, fabsf(v[d]));
          const int nb = base_index(x[d], a.K.inv_dx);
          bb_lo[d] = min(bb_lo[d], nb); bb_hi[d] = max(bb_hi[d], nb);
          fx[d] = x[d] * a.K.inv_dx - (float)nb;                                       // :409
        }
        const float dx = a.K.dx;
        const int sw = (q >> 1) & 3;
        pay[q * PS + (0 ^ sw)] = make_float4(mass * v[0], mass * v[1], mass * v[2], mass);
        pay[q * PS + (1 ^ sw)] = make_float4(aff[0] * dx, aff[3] * dx, aff[6] * dx, fx[0]);
        pay[q * PS + (2 ^ sw)] = make_float4(aff[1] * dx, aff[4] * dx, aff[7] * dx, fx[1]);
        pay[q * PS + (3 ^ sw)] = make_float4(aff[2] * dx, aff[5] * dx, aff[8] * dx, fx[2]);
      }
      } else {
      // ---- phase 1 (single pass): constitutive update (engine/mpm_solver.py:506-574), payload to shared memory
      constexpr int NIT = (CH + T - 1) / T;
      uint32_t pq[NIT];
#pragma unroll
      for (int k = 0; k < NIT; ++k) pq[k] = (tid + k * T < cn) ? a.perm[start + c0 + tid + k * T] : 0u;
#pragma unroll 1
      for (int q = tid; q < cn; q += T) {
        const int s = start + c0 + q;
        const uint32_t p = pq[0];
#pragma unroll
        for (int k = 0; k + 1 < NIT; ++k) pq[k] = pq[k + 1];
        using P = PStore<3, Q ? 2 : 0>;                            // f32 words, or quant=True: packed x v F + f32 C
        float x[D], v[D];
        P::load_x(a.src, p, x);
        P::load_v(a.src, p, v);
        float F[D * D], C[D * D], aff[D * D], mass;
        P::load_F(a.src, p, F);
#pragma unroll
        for (int i = 0; i < D * D; ++i) C[i] = __uint_as_float(__ldg(a.src + P::w(P::C + i, p)));
        float Jp = __uint_as_float(__ldg(a.src + P::w(P::JP, p)));
        const int mat = (int)tag_mat(__ldg(a.src + P::w(P::TAG, p)));
        particle_update<D>(a.K, a.dt, mat, F, C, Jp, aff, mass);
        P::store_F(a.dst, s, F);                                   // (quant: the store rounds F to its 16-bit grid, :567)
        a.dst[P::w(P::JP, s)] = __float_as_uint(Jp);
        // x and the tag go to the sorted slot as well (raw words): G2P then streams them instead of chasing perm -> x
        // (not in the halo variant: it is at its register limit and measured 3-7 % slower with them; k_g2p<.., XS = false>)
        if constexpr (!FUSED) {
#pragma unroll
          for (int i = 0; i < P::XW; ++i) a.dst[P::w(P::X + i, s)] = __ldg(a.src + P::w(P::X + i, p));
          a.dst[P::w(P::TAG, s)] = __ldg(a.src + P::w(P::TAG, p));
        }
        float fx[D];
#pragma unroll
        for (int d = 0; d < D; ++d) fx[d] = x[d] * a.K.inv_dx - (float)base_index(x[d], a.K.inv_dx);   // :503
        const float dx = a.K.dx;                                       // dpos = (o - fx) * dx
        const int sw = (q >> 1) & 3;                                  // 16-byte columns rotated: conflict-free STS.128
        pay[q * PS + (0 ^ sw)] = make_float4(mass * v[0], mass * v[1], mass * v[2], mass);
        pay[q * PS + (1 ^ sw)] = make_float4(aff[0] * dx, aff[3] * dx, aff[6] * dx, fx[0]);
        pay[q * PS + (2 ^ sw)] = make_float4(aff[1] * dx, aff[4] * dx, aff[7] * dx, fx[1]);
        pay[q * PS + (3 ^ sw)] = make_float4(aff[2] * dx, aff[5] * dx, aff[8] * dx, fx[2]);
      }
      }
      __syncthreads();                                                  // payload ready
      // next block's particle rows towards L2 while this one computes (storage order is last
      // substep's sorted order, so rows [start, end) are almost where `perm` will point)
      if (c0 == 0 && a.pf_mode) {
        const int nb = s_next;
        if (nb < npb) {
          // every word P2G reads (x .. material) is in the first FL::TAG + 1 rows of the block's tiles
          const int ns = a.pb_start[nb], ne = a.pb_start[nb + 1];
          const int t0 = ns >> TILE_LOG, nt = ((ne - 1) >> TILE_LOG) - t0 + 1;
          constexpr int NW = PStore<3, Q ? (G2P2G ? 1 : 2) : 0>::N;     // words per particle of the storage in use
          for (int i = tid; i < nt; i += T)
            prefetch_l2_range(a.src + (size_t)(t0 + i) * NW * TILE, (uint32_t)NW * TILE * 4u);
          if (tid == T - 1) prefetch_l2_range(a.perm + ns, (uint32_t)(ne - ns) * 4u);
        }
      }
      // ---- phase 2: (cell, slice) register accumulation (:577-584), packed pairs
      float2 acc01[P2G3::NPT], acc23[P2G3::NPT];
#pragma unroll
      for (int r = 0; r < P2G3::NPT; ++r) { acc01[r] = make_float2(0.f, 0.f); acc23[r] = make_float2(0.f, 0.f); }
      {
        const int lo = max(c_lo, c0) - c0, hi = min(c_hi, c0 + cn) - c0;
        for (int q = lo; q < hi; ++q) {
          const int sw = (q >> 1) & 3;
          const float4 M = pay[q * PS + (0 ^ sw)], Ax = pay[q * PS + (1 ^ sw)], Ay = pay[q * PS + (2 ^ sw)],
                       Az = pay[q * PS + (3 ^ sw)];
          const float fx0 = Ax.w, fx1 = Ay.w, fx2 = Az.w;
          const float wi = fmaf(fmaf(k2, fx0, k1), fx0, k0);
          const float d0s = slf - fx0;
          const float2 d0 = f2(d0s);
          const float2 a01 = __ffma2_rn(make_float2(Ax.x, Ax.y), d0, make_float2(M.x, M.y));
          const float a2 = fmaf(Ax.z, d0s, M.z);
          float wy[3], wz[3];
          wy[0] = 0.5f * (1.5f - fx1) * (1.5f - fx1);
          wy[1] = 0.75f - (fx1 - 1.0f) * (fx1 - 1.0f);
          wy[2] = 0.5f * (fx1 - 0.5f) * (fx1 - 0.5f);
          wz[0] = 0.5f * (1.5f - fx2) * (1.5f - fx2);
          wz[1] = 0.75f - (fx2 - 1.0f) * (fx2 - 1.0f);
          wz[2] = 0.5f * (fx2 - 0.5f) * (fx2 - 0.5f);
          float2 d2[3];
#pragma unroll
          for (int k = 0; k < 3; ++k) d2[k] = f2((float)k - fx2);
#pragma unroll
          for (int j = 0; j < 3; ++j) {
            const float d1s = (float)j - fx1;
            const float2 d1 = f2(d1s);
            const float wij = wi * wy[j];
            const float2 b01 = __ffma2_rn(make_float2(Ay.x, Ay.y), d1, a01);
            const float b2 = fmaf(Ay.z, d1s, a2);
#pragma unroll
            for (int k = 0; k < 3; ++k) {
              const float2 wt = f2(wij * wz[k]);
              const float2 t01 = __ffma2_rn(make_float2(Az.x, Az.y), d2[k], b01);
              const float2 t2m = make_float2(fmaf(Az.z, d2[k].x, b2), M.w);        // (momentum z, mass)
              acc01[j * 3 + k] = __ffma2_rn(wt, t01, acc01[j * 3 + k]);
              acc23[j * 3 + k] = __ffma2_rn(wt, t2m, acc23[j * 3 + k]);
            }
          }
        }
      }
      // the first warp out of scatter work prepares the next block
      if (last) {
        int tk = 0;
        if (lane == 0) tk = atomicAdd(&s_ticket, 1);
        tk = __shfl_sync(0xffffffffu, tk, 0);
        if (tk == 0) p2g3_prepare(a, s_next, npb, lane, s_cs[u ^ 1], s_order[u ^ 1], s_hist);
      }
      __syncthreads();                                                  // payload dead
      // ---- phase 3: accumulators -> the nine (x-slice, y-offset) copies, z offset by rounds.
      // For a fixed z offset k the map cell -> (cx, cy, cz + k) is injective, so a round is a
      // plain store (first touch) or a private read-add-write.
      {
        const int cx = cell >> 4, cy = (cell >> 2) & 3, cz = cell & 3;
        float4* my = pay + (sl * 3) * P2G3::CP + (cx * 4 + cy) * 6 + cz;
#pragma unroll
        for (int j = 0; j < 3; ++j) {                                   // z = cz (k = 0); z = 4, 5 (k = 2)
          my[j * P2G3::CP] = make_float4(acc01[j * 3].x, acc01[j * 3].y, acc23[j * 3].x, acc23[j * 3].y);
          if (cz >= 2)
            my[j * P2G3::CP + 2] = make_float4(acc01[j * 3 + 2].x, acc01[j * 3 + 2].y, acc23[j * 3 + 2].x,
                                               acc23[j * 3 + 2].y);
        }
        if (tid == 0) { s_ticket = 0; }
        __syncthreads();
#pragma unroll
        for (int j = 0; j < 3; ++j) {                                   // z = cz + 1
          float4 t = my[j * P2G3::CP + 1];
          t.x += acc01[j * 3 + 1].x; t.y += acc01[j * 3 + 1].y; t.z += acc23[j * 3 + 1].x; t.w += acc23[j * 3 + 1].y;
          my[j * P2G3::CP + 1] = t;
        }
        __syncthreads();
        if (cz < 2) {
#pragma unroll
          for (int j = 0; j < 3; ++j) {                                 // z = cz + 2 in {2, 3}
            float4 t = my[j * P2G3::CP + 2];
            t.x += acc01[j * 3 + 2].x; t.y += acc01[j * 3 + 2].y; t.z += acc23[j * 3 + 2].x; t.w += acc23[j * 3 + 2].y;
            my[j * P2G3::CP + 2] = t;
          }
        }
      }
      __syncthreads();                                                  // copies ready
      // ---- phase 4: node = sum of the copies (i, j) with 0 <= nx - i, ny - j < 4 -> global grid
      int hb_x = 0, hb_y = 0, hb_z = 0;      // absolute block column, (y, z) position in the halo plane
      if (boundary) {
        int rel[3];
        key_to_rel<3>(a.L, a.pb_key[b], rel);
        hb_x = rel[0] + a.L.ob[0]; hb_y = rel[1]; hb_z = rel[2];
      }
#pragma unroll
      for (int pass = 0; pass < 2; ++pass) {
        const int n = tid + pass * T;
        if (n < G::TN) {
          const int nz = n % 6, ny = (n / 6) % 6, nx = n / 36;
          const float4* row = pay + (nx * 4 + ny) * 6 + nz;
          float4 val = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
          for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j)
              if ((unsigned)(nx - i) < 4u && (unsigned)(ny - j) < 4u) {
                const float4 o = row[(i * 3 + j) * P2G3::CP - i * 24 - j * 6];
                val.x += o.x; val.y += o.y; val.z += o.z; val.w += o.w;
              }
          if (val.w != 0.0f) {
            int oct, cellg;
            tile_node<D>(n, oct, cellg);
            const int slot = s_nbr[oct];
            if (slot >= 0) red_add_v4(a.grid + (size_t)slot * G::CELLS + cellg, val);
            if (boundary) {
              // a node in the first two cell layers of a grid column shared with a neighbour rank: this block's
              // partial sum also goes into its slot of that neighbour's halo plane (a plain store into peer memory
              // over NVLink -- the slot belongs to this block alone); the neighbour's grid op adds the slabs
              const int side = (hb_x == a.slab.hi - 1 && nx >= 4) ? 1 : ((hb_x == a.slab.lo && nx <= 1) ? 0 : -1);
              if (side >= 0 && a.cb.plane_out[side]) {
                float4* plane = a.cb.plane_out[side] + (size_t)(a.cb.epoch % 3u) * a.cb.plane_blocks * HALO_SLAB;
                plane[(size_t)(hb_y * a.L.eb[2] + hb_z) * HALO_SLAB + (nx & 3) * 36 + ny * 6 + nz] = val;
              }
            }
          }
        }
      }
      if (boundary && last) __threadfence_system();                     // my reductions are performed before ...
      int q_next = 0;
      if (fused) q_next = s_ticket_q;
      if (last) { b = s_next; u ^= 1; }
      __syncthreads();                                                  // copies dead, s_next consumed
      if (boundary && last && tid == 0) {                               // ... the block is counted as sent
        if (atomicAdd(&a.st->halo_done, 1) == n_bnd - 1) {
          __threadfence_system();
          p2g3_publish_halo(a.cb);
        }
      }
      if (fused && last) qpos = q_next;
    }
  }
  if constexpr (G2P2G) {      // compute_max_velocity (:726-735) and the next bounding box, once per CTA lifetime
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      vmax = fmaxf(vmax, __shfl_xor_sync(0xffffffffu, vmax, o));
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        bb_lo[d] = min(bb_lo[d], __shfl_xor_sync(0xffffffffu, bb_lo[d], o));
        bb_hi[d] = max(bb_hi[d], __shfl_xor_sync(0xffffffffu, bb_hi[d], o));
      }
    }
    if (lane == 0) {
      if (vmax != vmax) vmax = __int_as_float(0x7f800000);
      atomicMax(&a.st->maxv_bits, __float_as_uint(vmax));
#pragma unroll
      for (int d = 0; d < 3; ++d)
        if (bb_lo[d] <= bb_hi[d]) { atomicMin(&a.st->bb_min[d], bb_lo[d]); atomicMax(&a.st->bb_max[d], bb_hi[d]); }
    }
  }
}

}  // namespace mpm
