// Multi-GPU slab decomposition: device side of the two neighbour exchanges
// (no counterpart in the reference, which is single-device).
//
// Rank r owns leaf-block columns [lo, hi) along x.  A particle belongs to the
// rank that owns its BASE block, so its 3^D stencil spills at most into block
// column `hi`, the first column of rank r+1.  Per substep:
//   after P2G   both ranks hold partial (momentum, mass) sums of that shared
//               column: each packs its copy (k_halo_pack), the buffers are
//               exchanged (NCCL send/recv, issued by the host between kernels)
//               and added (k_halo_add).  a + b == b + a bit for bit, so both
//               then run the grid op on identical sums and no velocities need
//               to travel back.
//   in G2P      a particle whose new base block left [lo, hi) is also written,
//               complete (all state words), to the migration buffer of that
//               side; the next substep's binning drops it locally
//               (k_bin_keys) and the neighbour appends it (k_mig_unpack).
// Buffers have a fixed capacity so no message size ever depends on a device
// value (no host synchronisation inside a batch); word 0 of each is the count.
#pragma once
#include "mpm_kernels.cuh"

namespace mpm {

// (by, bz) of a leaf block, absolute coordinates (each < 2^15), in one word
__device__ __forceinline__ uint32_t halo_key(int by, int bz) { return ((uint32_t)by << 16) | ((uint32_t)bz & 0xFFFFu); }

// one warp per active grid block: copy the blocks of column `bx_abs` to `buf`
template <int D>
__global__ void k_halo_pack(const float4* __restrict__ grid, const uint32_t* __restrict__ gb_key, KeyLayout L,
                            int bx_abs, uint32_t* __restrict__ buf, int halo_cap, int side, Status* st) {
  using G = Geo<D>;
  pdl_enter();
  if (st->err) return;
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarp = (gridDim.x * blockDim.x) >> 5;
  const int ngb = st->ngb;
  for (int g = warp; g < ngb; g += nwarp) {
    int rel[D];
    key_to_rel<D>(L, gb_key[g], rel);
    if (rel[0] + L.ob[0] != bx_abs) continue;
    int idx = 0;
    if (lane == 0) idx = atomicAdd(&st->halo_cnt[side], 1);
    idx = __shfl_sync(0xffffffffu, idx, 0);
    if (idx >= halo_cap) continue;            // counted; k_comm_headers raises the error
    if (lane == 0) buf[COMM_HEADER + idx] = halo_key(rel[1] + L.ob[1], D == 3 ? rel[D - 1] + L.ob[D - 1] : 0);
    float4* out = reinterpret_cast<float4*>(buf + COMM_HEADER + halo_cap) + (size_t)idx * G::CELLS;
    for (int c = lane; c < G::CELLS; c += 32) out[c] = grid[(size_t)g * G::CELLS + c];
    __threadfence_system();   // peer path: the records may live in the neighbour's memory
  }
}

// one warp per received block: add it to the local copy of column `bx_abs`
// (blocks this rank does not have active are not needed here and are skipped)
template <int D>
__global__ void k_halo_add(float4* __restrict__ grid, const int* __restrict__ flags, const int* __restrict__ fscan,
                           int nlin, KeyLayout L, int bx_abs, const uint32_t* __restrict__ buf, int halo_cap,
                           const Status* st) {
  using G = Geo<D>;
  pdl_enter();
  if (st->err) return;
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarp = (gridDim.x * blockDim.x) >> 5;
  const int cnt = min((int)buf[0], halo_cap);
  const int npb = st->npb;
  for (int i = warp; i < cnt; i += nwarp) {
    const uint32_t k = buf[COMM_HEADER + i];
    int rel[D];
    rel[0] = bx_abs - L.ob[0];
    rel[1] = (int)(k >> 16) - L.ob[1];
    if constexpr (D == 3) rel[2] = (int)(k & 0xFFFFu) - L.ob[2];
    bool inside = true;
#pragma unroll
    for (int d = 0; d < D; ++d) inside = inside && rel[d] >= 0 && rel[d] < L.eb[d];
    if (!inside) continue;
    const int lin = (int)rel_to_key<D>(L, rel);
    if (!flags[nlin + lin]) continue;
    const int slot = fscan[nlin + lin] - npb;
    const float4* in = reinterpret_cast<const float4*>(buf + COMM_HEADER + halo_cap) + (size_t)i * G::CELLS;
    float4* dst = grid + (size_t)slot * G::CELLS;
    for (int c = lane; c < G::CELLS; c += 32) {
      float4 a = dst[c];
      const float4 b = in[c];
      a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
      dst[c] = a;
    }
  }
}

// one thread: wait until both neighbours have published `epoch` (null = no neighbour)
__global__ void k_wait_flags(const uint32_t* f0, const uint32_t* f1, uint32_t epoch, Status* st) {
  pdl_enter();
  const unsigned long long t0 = global_ns();
  const uint32_t* f[2] = {f0, f1};
  for (int s = 0; s < 2; ++s) {
    if (!f[s]) continue;
    while ((int32_t)(ld_acquire_sys(f[s]) - epoch) < 0) {
      if (global_ns() - t0 > 4000000000ull) { st->err |= ERR_COMM_TIMEOUT; return; }   // 4 s: the peer is gone
      __nanosleep(200);
    }
  }
}

// message headers (one thread): counts, overflow detection, and (peer path) the signal.
// They run even after a device error so that a neighbour never waits for ever.
__global__ void k_halo_headers(CommBufs cb, uint32_t epoch, Status* st) {
  pdl_enter();
  for (int s = 0; s < 2; ++s) {
    if (!cb.halo[s]) continue;
    int c = st->err ? 0 : st->halo_cnt[s];
    if (c > cb.halo_cap) { st->err |= ERR_COMM_CAPACITY; c = 0; }
    cb.halo[s][0] = (uint32_t)c;
    if (cb.flag_halo[s]) { __threadfence_system(); st_release_sys(cb.flag_halo[s], epoch); }
  }
}
__global__ void k_mig_headers(CommBufs cb, uint32_t epoch, Status* st) {
  pdl_enter();
  publish_migration(cb, epoch, st);
}

// append the particles received from the -x and +x neighbours to the live set; when the last
// G2P already emitted the coming substep's sort keys and block flags (fused key pass), do the
// same for the appended rows (same arithmetic as k_bin_keys)
template <int D>
__global__ void k_mig_unpack(uint32_t* __restrict__ state, size_t cap, uint32_t* __restrict__ from_lo,
                             uint32_t* __restrict__ from_hi, int mig_cap, Status* st,
                             uint32_t* __restrict__ keys, int* __restrict__ flags, int nlin, KeyLayout L, Slab slab,
                             float inv_dx, CommBufs cb, Statics stat) {
  using G = Geo<D>;
  using FL = Fld<D>;
  constexpr int NW = FL::JP + 1;          // state words copied verbatim; the message's material / colour / id / emitter
                                          // rows become the tag and a new row of the static side arrays
  pdl_enter();
  // fused exchange: the neighbours publish the epoch of the substep whose leavers these buffers hold
  if (cb.fused) cta_wait_epochs(from_lo ? cb.wait_mig[0] : nullptr, from_hi ? cb.wait_mig[1] : nullptr, cb.epoch, st);
  bool active = *(volatile int*)&st->err == 0;
  const int c0 = (active && from_lo) ? min((int)*(volatile uint32_t*)from_lo, mig_cap) : 0;
  const int c1 = (active && from_hi) ? min((int)*(volatile uint32_t*)from_hi, mig_cap) : 0;
  const int base = st->n_cur, sbase = st->n_static;
  if (active && ((size_t)base + c0 + c1 > cap || (size_t)sbase + c0 + c1 > cap)) {
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicOr(&st->err, ERR_PARTICLE_CAPACITY);
    active = false;
  }
  const int total = active ? (c0 + c1) * NW : 0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int f = i / (c0 + c1), r = i % (c0 + c1);
    const uint32_t v = r < c0 ? from_lo[COMM_HEADER + (size_t)f * mig_cap + r]
                              : from_hi[COMM_HEADER + (size_t)f * mig_cap + (r - c0)];
    state[word<D>(f, (uint32_t)(base + r))] = v;
  }
  for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < (active ? c0 + c1 : 0); r += gridDim.x * blockDim.x) {
    const uint32_t* m = r < c0 ? from_lo + COMM_HEADER + r : from_hi + COMM_HEADER + (r - c0);
    const uint32_t sid = (uint32_t)(sbase + r);
    state[word<D>(FL::TAG, (uint32_t)(base + r))] = make_tag(m[(size_t)FL::MAT * mig_cap], sid);
    stat.color[sid] = m[(size_t)FL::COLOR * mig_cap];
    stat.gid[sid] = m[(size_t)FL::ID * mig_cap];
    stat.emit[sid] = m[(size_t)FL::EMIT * mig_cap];
  }
  const int nrows = (active && keys) ? c0 + c1 : 0;
  for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < nrows; r += gridDim.x * blockDim.x) {
    uint32_t lin = 0, cell = 0, sp = 0;
    bool bad = false, mine = true;
#pragma unroll
    for (int d = 0; d < D; ++d) {
      const uint32_t w = r < c0 ? from_lo[COMM_HEADER + (size_t)(Fld<D>::X + d) * mig_cap + r]
                                : from_hi[COMM_HEADER + (size_t)(Fld<D>::X + d) * mig_cap + (r - c0)];
      const int g = base_index(__uint_as_float(w), inv_dx) + L.half;
      if (d == 0 && slab.enabled) { const int bx = g >> G::LOG_LEAF; mine = bx >= slab.lo && bx < slab.hi; }
      int rel = (g >> G::LOG_LEAF) - L.ob[d];
      if (rel < 0 || rel > L.eb[d] - 2) { bad = true; rel = min(max(rel, 0), L.eb[d] - 2); }
      lin = lin * (uint32_t)L.eb[d] + (uint32_t)rel;
      const uint32_t lc = (uint32_t)(g & (G::LEAF - 1));
      cell = (cell << G::LOG_LEAF) | lc;
      sp |= (lc >= (uint32_t)(G::LEAF - 2)) ? (1u << d) : 0u;
    }
    keys[base + r] = mine ? ((lin << G::CB) | cell) : INVALID_KEY;
    if (mine && bad) { atomicOr(&st->next_err, ERR_BBOX); mine = false; }
    if (!mine) continue;
    flags[lin] = 1;
    int* gf = flags + nlin;
#pragma unroll
    for (uint32_t o = 0; o < (uint32_t)G::NO; ++o)
      if ((o & ~sp) == 0) gf[(int)lin + oct_delta_l<D>(L, (int)o)] = 1;
  }
  if (cb.fused) {
    // the LAST CTA to finish commits the appended rows and marks the messages consumed (no commit kernel)
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0 && atomicAdd(&st->unpack_done, 1) == (int)gridDim.x - 1) {
      st->unpack_done = 0;
      if (active) { st->n_cur = base + c0 + c1; st->n_static = sbase + c0 + c1; }
      if (from_lo) from_lo[0] = 0;
      if (from_hi) from_hi[0] = 0;
    }
  }
}
// ---- re-cutting (DistributedMPMSolver.rebalance): after the slab bounds have changed, rows whose base block column
// now lies outside [lo, hi) are moved to the neighbour on that side in rounds of at most `cap` rows per side: the row
// (all virtual words: its static attributes travel) goes into a message buffer of the migration format and its tag is
// marked DEAD so that a later round does not send it again and the next binning drops it.  With cap == 0 nothing is
// sent: rows outside the slab (leavers of the last substep, already delivered) are only marked, so that moving the cut
// over them cannot bring them back to life.
template <int D>
__global__ void k_rebalance_pack(uint32_t* __restrict__ state, Statics stat, int n, float inv_dx, int half, Slab slab,
                                 uint32_t* __restrict__ buf_lo, uint32_t* __restrict__ buf_hi, int cap,
                                 unsigned long long* __restrict__ counters /* [0..1] sent lo / hi, [2] still outside */) {
  using G = Geo<D>;
  using FL = Fld<D>;
  for (uint32_t s = blockIdx.x * blockDim.x + threadIdx.x; s < (uint32_t)n; s += gridDim.x * blockDim.x) {
    const size_t wt = word<D>(FL::TAG, s);
    const uint32_t tag = state[wt];
    if (tag_mat(tag) == MAT_DEAD) continue;
    const int bx = (base_index(__uint_as_float(state[word<D>(FL::X, s)]), inv_dx) + half) >> G::LOG_LEAF;
    const int dir = bx < slab.lo ? 0 : (bx >= slab.hi ? 1 : -1);
    if (dir < 0) continue;
    if (cap == 0) { state[wt] = make_tag(MAT_DEAD, tag_sid(tag)); continue; }
    uint32_t* buf = dir == 0 ? buf_lo : buf_hi;
    const unsigned long long idx = buf ? atomicAdd(&counters[dir], 1ull) : (unsigned long long)cap;
    if (idx >= (unsigned long long)cap) { atomicAdd(&counters[2], 1ull); continue; }
    uint32_t* m = buf + COMM_HEADER + idx;
    for (int f = 0; f < FL::NV; ++f) m[(size_t)f * cap] = vword<D>(state, stat, f, s);
    state[wt] = make_tag(MAT_DEAD, tag_sid(tag));
  }
}
__global__ void k_rebalance_headers(uint32_t* buf_lo, uint32_t* buf_hi, int cap, const unsigned long long* counters) {
  if (buf_lo) buf_lo[0] = (uint32_t)min(counters[0], (unsigned long long)cap);
  if (buf_hi) buf_hi[0] = (uint32_t)min(counters[1], (unsigned long long)cap);
}

__global__ void k_mig_commit(uint32_t* from_lo, uint32_t* from_hi, int mig_cap, Status* st) {
  pdl_enter();
  if (st->err) return;
  const int c0 = from_lo ? min((int)from_lo[0], mig_cap) : 0;
  const int c1 = from_hi ? min((int)from_hi[0], mig_cap) : 0;
  st->n_cur += c0 + c1;
  st->n_static += c0 + c1;
  if (from_lo) from_lo[0] = 0;   // consumed: a second unpack of the same message appends nothing
  if (from_hi) from_hi[0] = 0;
}

}  // namespace mpm
