// Binning (build_pid, /root/reference/engine/mpm_solver.py:344-361) as a
// single-digit radix (counting) sort on the (leaf block, cell) key.
//
// Particles are stored in the previous substep's sorted order and move less
// than a cell per substep, so a full multi-pass LSD radix sort re-derives an
// order that is already almost there.  With a dense flag table over the
// block-aligned particle box (the key layout's eb[0]*eb[1]*eb[2] entries) one
// pass suffices:
//   k_bin_keys     key per particle (4 particles per thread, 128-bit accesses); flag its leaf block and every leaf block
//                  its 3^D stencil touches (the reference's active-block set)
//   ExclusiveSum   over [particle-block flags | grid-block flags]: dense slots
//                  in key order
//   k_bin_rank     rank of the particle inside its (block, cell) bucket by a
//                  warp-aggregated atomic counter
//   ExclusiveSum   over the bucket counts: bucket starts (= block starts and
//                  the per-cell ranges P2G needs, for free)
//   k_bin_scatter  perm[start + rank] = particle
//   k_bin_finish   per particle block: start, key, the 2^D neighbour slots;
//                  grid-block keys; counts and per-substep status reset
// Bin index per particle, per-block counts and the active-block set are
// bit-identical to the oracle; the order inside a cell is arbitrary, as in the
// reference (ti.append is an atomic).  Boxes too large for the flag table fall
// back to the multi-pass radix sort in mpm_api.cu.
#pragma once
#include "mpm_kernels.cuh"

namespace mpm {

template <int D> __device__ __forceinline__ int oct_delta(const KeyLayout& L, int o) {
  if constexpr (D == 3) return ((o & 1) ? L.eb[1] * L.eb[2] : 0) + ((o & 2) ? L.eb[2] : 0) + ((o & 4) ? 1 : 0);
  else return ((o & 1) ? L.eb[1] : 0) + ((o & 2) ? 1 : 0);
}

__device__ __forceinline__ void commit_substep(Status* st) {
  if (!st->err) {
    st->done += 1;
    if (st->maxv_bits > st->maxv_all) st->maxv_all = st->maxv_bits;
  }
}

// The three per-particle passes below are pure streaming kernels whose speed is set by
// the number of bytes in flight, so every thread handles FOUR consecutive particles with
// 128-bit loads/stores (capacity is a multiple of 64, rows are 16-byte aligned).
template <int D, int QM = 0>      // QM: storage mode of PStore (0 f32 words, 2 bit-packed x of quant=True)
__global__ void k_bin_keys(const uint32_t* __restrict__ state, size_t cap, float inv_dx, KeyLayout L, Slab slab,
                           uint32_t* __restrict__ keys, int* __restrict__ flags, int nlin, int commit_prev,
                           Status* st) {
  using G = Geo<D>;
  if (commit_prev && blockIdx.x == 0 && threadIdx.x == 0) commit_substep(st);
  if (st->err) return;
  const int n = st->n_cur;
  const uint32_t nquad = ((uint32_t)n + 3u) >> 2;
  for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < nquad; t += gridDim.x * blockDim.x) {
    float4 xs[D];
    if constexpr (QM == 0) {
#pragma unroll
      for (int d = 0; d < D; ++d)
        xs[d] = __ldg(reinterpret_cast<const float4*>(state + word<D>(Fld<D>::X + d, 4u * t)));   // 4 particles of one tile row
    } else {
      using P = PStore<D, QM>;
      const uint4 q0 = __ldg(reinterpret_cast<const uint4*>(state + P::w(P::X, 4u * t)));
      const uint4 q1 = __ldg(reinterpret_cast<const uint4*>(state + P::w(P::X + 1, 4u * t)));
      const uint32_t w0[4] = {q0.x, q0.y, q0.z, q0.w}, w1[4] = {q1.x, q1.y, q1.z, q1.w};
      float xd[4][3];
#pragma unroll
      for (int j = 0; j < 4; ++j) { const uint32_t w[2] = {w0[j], w1[j]}; decode_x3(w, xd[j]); }
#pragma unroll
      for (int d = 0; d < D; ++d) xs[d] = make_float4(xd[0][d], xd[1][d], xd[2][d], xd[3][d]);
    }
    // with slabs a row can be DEAD (handed over while the cuts moved, mpm_rebalance_pack): its position no longer counts
    uint4 tg = make_uint4(0, 0, 0, 0);
    if (slab.enabled) tg = __ldg(reinterpret_cast<const uint4*>(state + PStore<D, QM>::w(PStore<D, QM>::TAG, 4u * t)));
    uint32_t out[4];
    uint32_t prev_lin = 0xFFFFFFFFu, prev_om = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const uint32_t p = 4u * t + j;
      uint32_t lin = 0, cell = 0, sp = 0;
      const uint32_t tgj = j == 0 ? tg.x : (j == 1 ? tg.y : (j == 2 ? tg.z : tg.w));
      bool bad = false, mine = p < (uint32_t)n && tag_mat(tgj) != MAT_DEAD;
#pragma unroll
      for (int d = 0; d < D; ++d) {
        const float xv = j == 0 ? xs[d].x : (j == 1 ? xs[d].y : (j == 2 ? xs[d].z : xs[d].w));
        int g = base_index(xv, inv_dx) + L.half;
        if (d == 0 && slab.enabled) { const int bx = g >> G::LOG_LEAF; mine = mine && bx >= slab.lo && bx < slab.hi; }
        int rel = (g >> G::LOG_LEAF) - L.ob[d];
        if (rel < 0 || rel > L.eb[d] - 2) { bad = true; rel = min(max(rel, 0), L.eb[d] - 2); }
        lin = lin * (uint32_t)L.eb[d] + (uint32_t)rel;
        const uint32_t lc = (uint32_t)(g & (G::LEAF - 1));
        cell = (cell << G::LOG_LEAF) | lc;
        sp |= (lc >= (uint32_t)(G::LEAF - 2)) ? (1u << d) : 0u;
      }
      // a particle whose base block left this rank's slab has already been handed to the
      // neighbour (G2P packs it): it is dropped from the local sort
      out[j] = mine ? ((lin << G::CB) | cell) : INVALID_KEY;
      if (mine && bad) { atomicOr(&st->err, ERR_BBOX); mine = false; }
      if (!mine) continue;
      // flag the leaf block and the blocks the stencil reaches; consecutive particles
      // mostly repeat the previous one's block and octants
      uint32_t om = 0;
#pragma unroll
      for (uint32_t o = 0; o < (uint32_t)G::NO; ++o)
        if ((o & ~sp) == 0) om |= 1u << o;
      if (lin == prev_lin && (om & ~prev_om) == 0) continue;
      prev_om = lin == prev_lin ? (prev_om | om) : om;
      prev_lin = lin;
      if (flags[lin] == 0) flags[lin] = 1;
      int* gf = flags + nlin;
#pragma unroll
      for (int o = 0; o < G::NO; ++o)
        if ((om >> o) & 1u) {
          const int tt = (int)lin + oct_delta<D>(L, o);
          if (gf[tt] == 0) gf[tt] = 1;
        }
    }
    reinterpret_cast<uint4*>(keys)[t] = make_uint4(out[0], out[1], out[2], out[3]);
  }
}

// ---------------------------------------------------------------------------
// Exclusive prefix sum of n ints in ONE launch (replaces cub::DeviceScan's init + scan pair, and
// takes part in the programmatic-dependent-launch chain): decoupled look-back over 8192-item
// tiles.  A tile descriptor is {epoch:30 | status:2 | value:32}; the epoch changes with every
// launch, so descriptors of earlier launches read as "not ready" and are never cleared.  The grid
// is persistent and no larger than what is co-resident; every CTA takes its tiles in increasing
// order, so the predecessors a look-back waits for are always being worked on.
// COMMIT: thread 0 first commits the previous substep (k_substep_begin folded in).
// ---------------------------------------------------------------------------
static constexpr int SCAN_T = 512, SCAN_IPT = 16, SCAN_TILE = SCAN_T * SCAN_IPT;   // 8192-item tiles: few look-back hops
__device__ __forceinline__ unsigned long long scan_ld(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void scan_st(unsigned long long* p, unsigned long long v) {
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
template <bool COMMIT>
__global__ void __launch_bounds__(SCAN_T) k_scan_excl(const int* __restrict__ in, int* __restrict__ out, int n,
                                                     unsigned long long* __restrict__ desc, uint32_t epoch,
                                                     Status* st, const int* __restrict__ n_dev, int n_mul) {
  pdl_enter();
  // optional device-side length (the per-cell tables are sized by block CAPACITY, the scan only
  // needs the blocks that exist): n = min(n, *n_dev * n_mul + 1)
  if (n_dev) n = min(n, max(*n_dev, 0) * n_mul + 1);
  if (COMMIT && blockIdx.x == 0 && threadIdx.x == 0) {
    if (!st->err) {
      st->done += 1;
      if (st->maxv_bits > st->maxv_all) st->maxv_all = st->maxv_bits;
    }
    st->err |= st->next_err;
    st->next_err = 0;
  }
  __shared__ int s_warp[SCAN_T / 32];
  __shared__ int s_excl;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int ntiles = (n + SCAN_TILE - 1) / SCAN_TILE;
  const unsigned long long tag = (unsigned long long)(epoch & 0x3fffffffu) << 34;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int base = tile * SCAN_TILE + tid * SCAN_IPT;
    int v[SCAN_IPT];
    if (base + SCAN_IPT <= n) {
#pragma unroll
      for (int q = 0; q < SCAN_IPT / 4; ++q) {
        const int4 a = __ldg(reinterpret_cast<const int4*>(in + base) + q);
        v[4 * q] = a.x; v[4 * q + 1] = a.y; v[4 * q + 2] = a.z; v[4 * q + 3] = a.w;
      }
    } else {
#pragma unroll
      for (int i = 0; i < SCAN_IPT; ++i) v[i] = base + i < n ? in[base + i] : 0;
    }
    int tsum = 0;
#pragma unroll
    for (int i = 0; i < SCAN_IPT; ++i) tsum += v[i];
    int incl = tsum;                                        // inclusive scan of the thread sums in the warp
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    if (lane == 31) s_warp[wid] = incl;
    __syncthreads();
    if (wid == 0) {
      int w = lane < SCAN_T / 32 ? s_warp[lane] : 0;
      int wi = w;
#pragma unroll
      for (int o = 1; o < SCAN_T / 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, wi, o);
        if (lane >= o) wi += t;
      }
      const int agg = __shfl_sync(0xffffffffu, wi, SCAN_T / 32 - 1);
      if (lane < SCAN_T / 32) s_warp[lane] = wi - w;        // exclusive prefix of each warp inside the tile
      int excl = 0;
      if (tile == 0) {
        if (lane == 0) scan_st(desc, tag | (2ull << 32) | (unsigned)agg);
      } else {
        if (lane == 0) scan_st(desc + tile, tag | (1ull << 32) | (unsigned)agg);
        int t = tile - 1;
        unsigned spins = 0;
        while (true) {
          const int idx = t - lane;
          unsigned long long d = tag | (2ull << 32);        // before tile 0: prefix 0
          if (idx >= 0) d = scan_ld(desc + idx);
          const bool ready = (d >> 34) == (tag >> 34) && ((d >> 32) & 3ull) != 0ull;
          if (!__all_sync(0xffffffffu, ready)) {
            if (++spins > (1u << 24)) { if (lane == 0) atomicOr(&st->err, ERR_SCAN_TIMEOUT); break; }   // never hang
            continue;
          }
          const unsigned pref = __ballot_sync(0xffffffffu, ((d >> 32) & 3ull) == 2ull);
          const int last = pref ? __ffs(pref) - 1 : 31;     // nearest tile whose inclusive prefix is known
          int val = lane <= last ? (int)(unsigned)(d & 0xffffffffull) : 0;
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) val += __shfl_xor_sync(0xffffffffu, val, o);
          excl += val;
          if (pref) break;
          t -= 32;
        }
        if (lane == 0) scan_st(desc + tile, tag | (2ull << 32) | (unsigned)(excl + agg));
      }
      if (lane == 0) s_excl = excl;
    }
    __syncthreads();
    int run = s_excl + s_warp[wid] + (incl - tsum);
    if (base + SCAN_IPT <= n) {
      int o[SCAN_IPT];
#pragma unroll
      for (int i = 0; i < SCAN_IPT; ++i) { o[i] = run; run += v[i]; }
#pragma unroll
      for (int q = 0; q < SCAN_IPT / 4; ++q)
        reinterpret_cast<int4*>(out + base)[q] = make_int4(o[4 * q], o[4 * q + 1], o[4 * q + 2], o[4 * q + 3]);
    } else {
#pragma unroll
      for (int i = 0; i < SCAN_IPT; ++i) {
        if (base + i < n) out[base + i] = run;
        run += v[i];
      }
    }
    __syncthreads();
  }
}

template <int D>
__global__ void k_bin_rank(const uint32_t* __restrict__ keys, const int* __restrict__ fscan,
                           int* __restrict__ cellcount, uint32_t* __restrict__ rank, uint32_t* __restrict__ pb_key,
                           int max_blocks, Status* st) {
  using G = Geo<D>;
  pdl_enter();
  if (st->err) return;
  const int n = st->n_cur;
  const uint32_t nquad = ((uint32_t)n + 3u) >> 2;
  // Particles are stored in last substep's (block, cell) order, so the four consecutive keys of a thread mostly form
  // one or two RUNS of equal keys: one atomic per run (adding its length) instead of one warp-wide match per particle
  // (66 -> see profiles/README.md).  Runs of the same cell in neighbouring threads simply hit the same counter twice.
  for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < nquad; t += gridDim.x * blockDim.x) {
    const uint4 kq = __ldg(reinterpret_cast<const uint4*>(keys) + t);
    const uint32_t kk[4] = {kq.x, kq.y, kq.z, kq.w};
    uint32_t idx[4], lin[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      // rows past n in the last quad hold stale keys when the keys came from G2P / the unpack pass
      const uint32_t key = 4u * t + j < (uint32_t)n ? kk[j] : INVALID_KEY;
      idx[j] = 0xFFFFFFFFu; lin[j] = 0;
      if (key != INVALID_KEY) {
        if (j > 0 && key == kk[j - 1] && idx[j - 1] != 0xFFFFFFFFu) { idx[j] = idx[j - 1]; lin[j] = lin[j - 1]; continue; }
        lin[j] = key >> G::CB;
        const int b = fscan[lin[j]];
        if (b < max_blocks) idx[j] = (uint32_t)b * G::CELLS + (key & (G::CELLS - 1));
        else atomicOr(&st->err, ERR_BLOCK_CAPACITY);
      }
    }
    uint32_t rr[4] = {0u, 0u, 0u, 0u};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (idx[j] == 0xFFFFFFFFu) continue;
      if (j > 0 && idx[j] == idx[j - 1]) { rr[j] = rr[j - 1] + 1u; continue; }     // inside a run
      int len = 1;
#pragma unroll
      for (int q = j + 1; q < 4; ++q) len += (q - j == len && idx[q] == idx[j]) ? 1 : 0;
      rr[j] = (uint32_t)atomicAdd(&cellcount[idx[j]], len);
      if (rr[j] == 0u) pb_key[idx[j] / G::CELLS] = lin[j];
    }
    reinterpret_cast<uint4*>(rank)[t] = make_uint4(rr[0], rr[1], rr[2], rr[3]);
  }
}

// Bucket starts for tables too long for one look-back chain (> 256 tiles: the prefix of k_scan_excl travels one
// 32-tile window per hop): two levels, all own kernels.  k_cell_sums: particles per particle block (one warp per
// block); k_scan_excl over those <= max_blocks sums; k_cell_starts: exclusive scan inside each block + its offset.
template <int D>
__global__ void k_cell_sums(const int* __restrict__ cellcount, const int* __restrict__ npb_dev, int* __restrict__ blocksum,
                            const Status* st) {
  using G = Geo<D>;
  pdl_enter();
  if (st->err) return;
  const int npb = *npb_dev, lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarp = (gridDim.x * blockDim.x) >> 5;
  for (int b = warp; b <= npb; b += nwarp) {
    int s = 0;
    if (b < npb)
      for (int c = lane; c < G::CELLS; c += 32) s += cellcount[(size_t)b * G::CELLS + c];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) blocksum[b] = s;                     // (entry npb = 0: the scan's last output is the total)
  }
}
template <int D>
__global__ void k_cell_starts(const int* __restrict__ cellcount, const int* __restrict__ npb_dev,
                              const int* __restrict__ blockstart, int* __restrict__ cellstart, const Status* st) {
  using G = Geo<D>;
  pdl_enter();
  if (st->err) return;
  const int npb = *npb_dev, lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarp = (gridDim.x * blockDim.x) >> 5;
  for (int b = warp; b <= npb; b += nwarp) {
    int run = blockstart[b];
    if (b == npb) { if (lane == 0) cellstart[(size_t)npb * G::CELLS] = run; continue; }
    for (int c0 = 0; c0 < G::CELLS; c0 += 32) {
      const int v = cellcount[(size_t)b * G::CELLS + c0 + lane];
      int incl = v;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
      }
      cellstart[(size_t)b * G::CELLS + c0 + lane] = run + incl - v;
      run += __shfl_sync(0xffffffffu, incl, 31);
    }
  }
}

template <int D>
__global__ void k_bin_scatter(const uint32_t* __restrict__ keys, const uint32_t* __restrict__ rank,
                              const int* __restrict__ fscan, const int* __restrict__ cellstart,
                              uint32_t* __restrict__ perm, const Status* st) {
  using G = Geo<D>;
  pdl_enter();
  if (st->err) return;
  const int n = st->n_cur;
  const uint32_t nquad = ((uint32_t)n + 3u) >> 2;
  for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < nquad; t += gridDim.x * blockDim.x) {
    const uint4 kq = __ldg(reinterpret_cast<const uint4*>(keys) + t);
    const uint4 rq = __ldg(reinterpret_cast<const uint4*>(rank) + t);
    const uint32_t kk[4] = {kq.x, kq.y, kq.z, kq.w}, rr[4] = {rq.x, rq.y, rq.z, rq.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (kk[j] == INVALID_KEY || 4u * t + j >= (uint32_t)n) continue;
      const int b = fscan[kk[j] >> G::CB];
      perm[cellstart[(size_t)b * G::CELLS + (kk[j] & (G::CELLS - 1))] + rr[j]] = 4u * t + j;
    }
  }
}

template <int D>
__global__ void k_bin_finish(const int* __restrict__ flags, const int* __restrict__ fscan, int nlin, KeyLayout L,
                             const uint32_t* __restrict__ pb_key, const int* __restrict__ cellstart,
                             int* __restrict__ pb_start, int* __restrict__ pb_nbr, uint32_t* __restrict__ gb_key,
                             int max_blocks, Status* st, Slab slab) {
  using G = Geo<D>;
  pdl_enter();
  if (st->err) return;
  const int npb = fscan[nlin];
  const int ngb = fscan[2 * nlin] - npb;
  const bool first = blockIdx.x == 0 && threadIdx.x == 0;
  if (first && max(npb, ngb) > st->need_blocks) st->need_blocks = max(npb, ngb);
  if (npb > max_blocks || ngb > max_blocks) {
    if (first) st->err |= ERR_BLOCK_CAPACITY;
    return;
  }
  if (first) {
    st->npb = npb; st->ngb = ngb; st->ngb_raw = ngb;
    const int n_live = cellstart[(size_t)npb * G::CELLS];   // particles that were ranked
    pb_start[npb] = n_live;
    st->n_live = n_live;
    st->mig_cnt[0] = st->mig_cnt[1] = 0;
    st->halo_cnt[0] = st->halo_cnt[1] = 0;
    st->work_p2g = 0; st->work_g2p = 0;
    st->maxv_bits = 0;
    st->maxgv_bits = 0;
    for (int d = 0; d < 3; ++d) { st->bb_min[d] = INT_MAX; st->bb_max[d] = INT_MIN; }
    // slabs: particle blocks of the first and of the last block column (keys are x-major, so they are the first
    // bnd_lo and the last bnd_hi entries of the block list); P2G takes them first (fused halo, mpm_comm.cuh)
    int blo = 0, bhi = 0;
    if (slab.enabled) {
      int plane = 1;
      for (int d = 1; d < D; ++d) plane *= L.eb[d];
      const long long rl = (long long)slab.lo - L.ob[0], rh = (long long)slab.hi - 1 - L.ob[0];
      if (rl >= 0 && rl < L.eb[0]) blo = fscan[(int)(rl + 1) * plane] - fscan[(int)rl * plane];
      if (rh >= 0 && rh < L.eb[0]) bhi = fscan[(int)(rh + 1) * plane] - fscan[(int)rh * plane];
      if (blo + bhi > npb) blo = npb - bhi;          // a one-column slab: the same blocks
    }
    st->bnd_lo = blo; st->bnd_hi = bhi;
    st->halo_done = 0;
  }
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < npb * G::NO; i += gridDim.x * blockDim.x) {
    const int b = i / G::NO, o = i % G::NO;
    const int t = (int)pb_key[b] + oct_delta<D>(L, o);
    int slot = -1;
    if (flags[nlin + t]) {
      slot = fscan[nlin + t] - npb;
      gb_key[slot] = (uint32_t)t;
    }
    pb_nbr[i] = slot;
    if (o == 0) pb_start[b] = cellstart[(size_t)b * G::CELLS];
  }
}

}  // namespace mpm
