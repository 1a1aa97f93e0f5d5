// P2G, cell-owner formulation (sm_100a).
//
// Shared-memory float atomics are CAS loops on this architecture
// (ATOMS.CAST.SPIN), ~2 SM-cycles per lane: 108 of them per particle made the
// first P2G 85% of the substep.  Here nothing on the accumulation path is
// atomic:
//   phase 0  cell ranges of the block from the sorted keys
//   phase 1  one thread per particle (coalesced through `perm`): F update, SVD,
//            plasticity, stress (engine/mpm_solver.py:506-574); new F/Jp go to
//            the particle's sorted slot; the scatter payload (fx, m*v, dx*A,
//            m) is staged in shared memory, SoA
//   phase 2  one thread per (cell, x-slice i): loops over the particles of its
//            cell and accumulates its 3^(D-1) nodes x (D+1) values in
//            REGISTERS (all particles of a cell share the same 3^D nodes)
//   phase 3  conflict-free flush: for a fixed node offset the map cell ->
//            cell + offset is injective, so each round is a plain
//            read-add-write into the slice's private tile copy
//   phase 4  tile -> global grid with 128-bit vector reductions
//            (REDG.E.ADD.F32x4), one per touched node
// The scatter itself follows engine/mpm_solver.py:577-584.
#pragma once
#include "mpm_kernels.cuh"

namespace mpm {

template <int D> struct P2GCfg;
template <> struct P2GCfg<3> {
  static constexpr int SL = 3, NPT = 9, PAY = 16, THREADS = 192;
};
template <> struct P2GCfg<2> {
  static constexpr int SL = 3, NPT = 3, PAY = 9, THREADS = 768;
};

template <int D, int CHUNK> constexpr size_t p2g_smem_bytes() {
  using G = Geo<D>;
  using P = P2GCfg<D>;
  return (size_t)P::SL * G::TN * sizeof(float4) + (size_t)P::PAY * CHUNK * sizeof(float) +
         (size_t)(G::CELLS + 1) * sizeof(int);
}

// CHUNK = particles staged per pass (shared-memory payload), MINB = CTAs per SM the
// register allocation is bounded for
template <int D, int CHUNK, int MINB>
__global__ void __launch_bounds__(P2GCfg<D>::THREADS, MINB) k_p2g_cell(SubstepArgs<D> a) {
  using G = Geo<D>;
  using FL = Fld<D>;
  using P = P2GCfg<D>;
  constexpr int T = P::THREADS, CH = CHUNK;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float4* tile = reinterpret_cast<float4*>(smem_raw);                 // [SL][TN]
  float* pay = reinterpret_cast<float*>(tile + P::SL * G::TN);        // [PAY][CH]
  int* cs = reinterpret_cast<int*>(pay + P::PAY * CH);                // [CELLS+1]
  __shared__ int s_b;
  __shared__ int s_nbr[G::NO];
  if (a.st->err) return;
  const int npb = a.st->npb;
  const int tid = threadIdx.x;
  const size_t cap = a.cap;
  const int sl = tid % P::SL;
  __shared__ int s_order[G::CELLS];                                   // cells by descending particle count

  __shared__ int s_next;
  if (tid == 0) s_b = atomicAdd(&a.st->work_p2g, 1);
  __syncthreads();
  int b = s_b;
  while (b < npb) {
    if (tid == 0) s_next = atomicAdd(&a.st->work_p2g, 1);   // claimed one block ahead, for the prefetch below
    const int start = a.pb_start[b], end = a.pb_start[b + 1];
    const int cnt = end - start;
    for (int n = tid; n < P::SL * G::TN; n += T) tile[n] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (tid < G::NO) s_nbr[tid] = a.pb_nbr[b * G::NO + tid];
    // phase 0: first sorted position of every cell of the block: straight from the
    // counting sort's bucket starts, or (radix fallback) from the sorted keys
    if (a.cellstart) {
      for (int c = tid; c <= G::CELLS; c += T) cs[c] = a.cellstart[(size_t)b * G::CELLS + c] - start;
    } else {
      for (int q = tid; q < cnt; q += T) {
        const int c = (int)(a.keys[start + q] & (G::CELLS - 1));
        const int cp = q > 0 ? (int)(a.keys[start + q - 1] & (G::CELLS - 1)) : -1;
        for (int k = cp + 1; k <= c; ++k) cs[k] = q;
        if (q == cnt - 1)
          for (int k = c + 1; k <= G::CELLS; ++k) cs[k] = cnt;
      }
    }
    __syncthreads();
    // Balance phase 2: a warp's trip count is the largest cell count among its
    // lanes, so hand out cells in descending-count order (3D; the 256-cell 2D
    // leaf keeps the natural order).
    int cell = tid / P::SL;
    if constexpr (D == 3) {
      if (tid < G::CELLS) {
        const int mine = cs[tid + 1] - cs[tid];
        int rank = 0;
        for (int c = 0; c < G::CELLS; ++c) {
          const int other = cs[c + 1] - cs[c];
          rank += (other > mine) || (other == mine && c < tid);
        }
        s_order[rank] = tid;
      }
      __syncthreads();
      cell = s_order[tid / P::SL];
    }
    int cl[D];                                                        // local cell coords
#pragma unroll
    for (int d = 0; d < D; ++d) cl[d] = (cell >> (G::LOG_LEAF * (D - 1 - d))) & (G::LEAF - 1);

    for (int c0 = 0; c0 < cnt; c0 += CH) {
      const int cn = min(CH, cnt - c0);
      // ---- phase 1: constitutive update, payload to shared memory
      // all of this thread's `perm` entries first: one exposed latency per chunk, not per particle
      constexpr int NIT = (CH + T - 1) / T;
      uint32_t pq[NIT];
#pragma unroll
      for (int k = 0; k < NIT; ++k) pq[k] = (tid + k * T < cn) ? a.perm[start + c0 + tid + k * T] : 0u;
#pragma unroll 1
      for (int q = tid; q < cn; q += T) {
        const int s = start + c0 + q;
        const uint32_t p = pq[0];
#pragma unroll
        for (int k = 0; k + 1 < NIT; ++k) pq[k] = pq[k + 1];   // rotate (keeps the loop body single-copy)
        float x[D], v[D];
#pragma unroll
        for (int d = 0; d < D; ++d) {
          x[d] = ldf<D>(a.src, FL::X + d, p);
          v[d] = ldf<D>(a.src, FL::V + d, p);
        }
        float F[D * D], C[D * D], aff[D * D], mass;
#pragma unroll
        for (int i = 0; i < D * D; ++i) {
          F[i] = ldf<D>(a.src, FL::F + i, p);
          C[i] = ldf<D>(a.src, FL::C + i, p);
        }
        float Jp = ldf<D>(a.src, FL::JP, p);
        const int mat = (int)tag_mat(ldu<D>(a.src, FL::TAG, p));
        particle_update<D>(a.K, a.dt, mat, F, C, Jp, aff, mass);
#pragma unroll
        for (int i = 0; i < D * D; ++i) stf<D>(a.dst, FL::F + i, s, F[i]);
        stf<D>(a.dst, FL::JP, s, Jp);
#pragma unroll
        for (int d = 0; d < D; ++d) stf<D>(a.dst, FL::X + d, s, x[d]);   // G2P reads x and the tag at the sorted slot
        stu<D>(a.dst, FL::TAG, s, ldu<D>(a.src, FL::TAG, p));
#pragma unroll
        for (int d = 0; d < D; ++d) {
          const int base = base_index(x[d], a.K.inv_dx);
          pay[d * CH + q] = x[d] * a.K.inv_dx - (float)base;            // fx (:503)
          pay[(D + d) * CH + q] = mass * v[d];
        }
#pragma unroll
        for (int i = 0; i < D * D; ++i) pay[(2 * D + i) * CH + q] = aff[i] * a.K.dx;   // dpos = (o - fx) * dx
        pay[(2 * D + D * D) * CH + q] = mass;
      }
      __syncthreads();
      // While this block's arithmetic runs, pull the next block's particle state
      // towards L2.  Storage order is last substep's sorted order, so the rows
      // [start, end) of the next block are (almost) where `perm` will point.
      if (c0 == 0) {
        const int nb = s_next;
        if (nb < npb) {
          const int ns = a.pb_start[nb], ne = a.pb_start[nb + 1];
          const int t0 = ns >> TILE_LOG, nt = ((ne - 1) >> TILE_LOG) - t0 + 1;
          for (int i = tid; i < nt; i += T)
            prefetch_l2_range(a.src + (size_t)(t0 + i) * FL::N * TILE, (uint32_t)(FL::TAG + 1) * TILE * 4u);
          if (tid == T - 1) prefetch_l2_range(a.perm + ns, (uint32_t)(ne - ns) * 4u);
        }
      }
      // ---- phase 2: per-(cell, slice) register accumulation
      float acc[P::NPT][D + 1];
#pragma unroll
      for (int i = 0; i < P::NPT; ++i)
#pragma unroll
        for (int j = 0; j <= D; ++j) acc[i][j] = 0.0f;
      {
        const int lo = max(cs[cell], c0) - c0, hi = min(cs[cell + 1], c0 + cn) - c0;
        for (int q = lo; q < hi; ++q) {
          float fx[D], mv[D], A[D * D];
#pragma unroll
          for (int d = 0; d < D; ++d) { fx[d] = pay[d * CH + q]; mv[d] = pay[(D + d) * CH + q]; }
#pragma unroll
          for (int i = 0; i < D * D; ++i) A[i] = pay[(2 * D + i) * CH + q];
          const float mass = pay[(2 * D + D * D) * CH + q];
          float w[3][D];
#pragma unroll
          for (int d = 0; d < D; ++d) {
            w[0][d] = 0.5f * (1.5f - fx[d]) * (1.5f - fx[d]);           // :505
            w[1][d] = 0.75f - (fx[d] - 1.0f) * (fx[d] - 1.0f);
            w[2][d] = 0.5f * (fx[d] - 0.5f) * (fx[d] - 0.5f);
          }
          const float wi = sl == 0 ? w[0][0] : (sl == 1 ? w[1][0] : w[2][0]);
          const float d0 = (float)sl - fx[0];
          if constexpr (D == 3) {
            const float a0 = mv[0] + A[0] * d0, a1 = mv[1] + A[3] * d0, a2 = mv[2] + A[6] * d0;
#pragma unroll
            for (int j = 0; j < 3; ++j) {
              const float d1 = (float)j - fx[1];
              const float wij = wi * w[j][1];
              const float b0 = a0 + A[1] * d1, b1 = a1 + A[4] * d1, b2 = a2 + A[7] * d1;
#pragma unroll
              for (int k = 0; k < 3; ++k) {
                const float d2 = (float)k - fx[2];
                const float wt = wij * w[k][2];
                acc[j * 3 + k][0] += wt * (b0 + A[2] * d2);
                acc[j * 3 + k][1] += wt * (b1 + A[5] * d2);
                acc[j * 3 + k][2] += wt * (b2 + A[8] * d2);
                acc[j * 3 + k][3] += wt * mass;
              }
            }
          } else {
            const float a0 = mv[0] + A[0] * d0, a1 = mv[1] + A[2] * d0;
#pragma unroll
            for (int j = 0; j < 3; ++j) {
              const float d1 = (float)j - fx[1];
              const float wt = wi * w[j][1];
              acc[j][0] += wt * (a0 + A[1] * d1);
              acc[j][1] += wt * (a1 + A[3] * d1);
              acc[j][2] += wt * mass;
            }
          }
        }
      }
      // ---- phase 3: conflict-free flush into this slice's tile copy (the
      // barrier of each round also protects the payload before the next chunk)
      {
        float4* my = tile + sl * G::TN;
#pragma unroll
        for (int r = 0; r < P::NPT; ++r) {
          int n;
          if constexpr (D == 3) n = ((cl[0] + sl) * G::T + (cl[1] + r / 3)) * G::T + (cl[2] + r % 3);
          else n = (cl[0] + sl) * G::T + (cl[1] + r);
          float4 t = my[n];
          t.x += acc[r][0]; t.y += acc[r][1]; t.z += acc[r][2];
          if constexpr (D == 3) t.w += acc[r][3];
          my[n] = t;
          __syncthreads();
        }
      }
    }
    // ---- phase 4: tile -> global grid
    for (int n = tid; n < G::TN; n += T) {
      float4 val = tile[n];
#pragma unroll
      for (int c = 1; c < P::SL; ++c) {
        const float4 o = tile[c * G::TN + n];
        val.x += o.x; val.y += o.y; val.z += o.z; val.w += o.w;
      }
      const float m = (D == 3) ? val.w : val.z;
      if (m != 0.0f) {
        int oct, cellg;
        tile_node<D>(n, oct, cellg);
        const int slot = s_nbr[oct];
        if (slot >= 0) red_add_v4(a.grid + (size_t)slot * G::CELLS + cellg, val);
      }
    }
    b = s_next;
    __syncthreads();
  }
}

}  // namespace mpm
