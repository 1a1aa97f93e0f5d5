// Device kernels of the MLS-MPM substep (sm_100a).  One substep inside a batch =
//   k_scan_excl<1> (flags; commits the previous substep) -> k_bin_rank -> k_scan_excl<0> (cells)
//   -> k_bin_scatter -> k_bin_finish                                     (mpm_bin.cuh; keys + flags came from
//   -> k_clear_grid -> k_p2g3 (mpm_p2g3.cuh) -> k_grid_op -> k_g2p        the previous k_g2p, else k_bin_keys)
// chained with programmatic dependent launch (pdl_enter()).  Fallback for particle boxes too large for the
// flag table:
//   k_reset -> k_keys -> radix sort -> head select -> k_pb_finalize -> k_pb_masks
//   -> sort/unique of candidate grid blocks -> k_gb_finalize -> k_nbr -> k_clear_grid -> k_p2g_cell -> ...
// replacing build_pid / p2g / grid_normalization_and_gravity / grid_bounding_box
// / collide / g2p / compute_max_velocity of /root/reference/engine/mpm_solver.py
// (:344-361, 487-616, 618-687, 694-735).
#pragma once
#include "mpm_common.cuh"
#include "mpm_quant.cuh"

namespace mpm {

// Particle state layout: tiles of 32 particles, inside a tile one 128-byte row per state word:
//   word(f, p) = ((p / 32) * NF + f) * 32 + p % 32.
// A warp that reads word f of 32 consecutive particles touches one 128-byte line, as with a plain structure of arrays,
// but all words of a particle sit at CONSTANT offsets (f * 128 B) from one address: a kernel that touches 26 words of
// a particle needs one 64-bit address computation instead of 26 (the strided layout [f][capacity] of round 1 cost ~3
// integer instructions per access and kept the register file full of row pointers).
static constexpr int TILE = 32, TILE_LOG = 5;
template <int NF> __host__ __device__ __forceinline__ size_t word_nf(int f, uint32_t p) {
  return ((size_t)(p >> TILE_LOG) * NF + f) * TILE + (p & (TILE - 1));
}
__host__ __device__ __forceinline__ size_t word_rt(int nf, int f, uint32_t p) {      // run-time field count
  return ((size_t)(p >> TILE_LOG) * nf + f) * TILE + (p & (TILE - 1));
}
template <int D> __host__ __device__ __forceinline__ size_t word(int f, uint32_t p) { return word_nf<Fld<D>::N>(f, p); }
template <int D> __device__ __forceinline__ float ldf(const uint32_t* __restrict__ s, int f, uint32_t p) {
  return __uint_as_float(__ldg(s + word<D>(f, p)));
}
template <int D> __device__ __forceinline__ uint32_t ldu(const uint32_t* __restrict__ s, int f, uint32_t p) {
  return __ldg(s + word<D>(f, p));
}
template <int D> __device__ __forceinline__ void stf(uint32_t* __restrict__ s, int f, uint32_t p, float v) {
  s[word<D>(f, p)] = __float_as_uint(v);
}
template <int D> __device__ __forceinline__ void stu(uint32_t* __restrict__ s, int f, uint32_t p, uint32_t v) {
  s[word<D>(f, p)] = v;
}

// ---- particle storage accessors.  Q = 0: the f32 words of Fld<D>.  quant=True in 3D stores x, v and F bit-packed
// (mpm_quant.cuh, ref :106-114, 216-247):
//   Q = 1 (with use_g2p2g):   xq[2] vq[2] Fq[5] Jp tag        = 11 words (44 B per set instead of 104); C does not exist
//                             in that mode (a register value of the fused kernel, ref :102-103)
//   Q = 2 (split substep):    xq[2] vq[2] Fq[5] Jp tag C[9]   = 20 words (80 B); C stays f32 as in the reference (:101-102, 219-220)
struct FldQ3 { static constexpr int X = 0, V = 2, F = 4, JP = 9, TAG = 10, C = 11, N = 11, NC = 20; };
__host__ __device__ __forceinline__ int quant_words(int quant) { return quant == 2 ? FldQ3::NC : FldQ3::N; }
template <int D, int Q> struct PStore;
template <int D> struct PStore<D, 0> {
  using FL = Fld<D>;
  static constexpr int N = FL::N, JP = FL::JP, TAG = FL::TAG, C = FL::C, X = FL::X, XW = D;
  static __device__ __forceinline__ size_t w(int f, uint32_t p) { return word<D>(f, p); }
  static __device__ __forceinline__ void load_x(const uint32_t* __restrict__ s, uint32_t p, float* x) {
#pragma unroll
    for (int d = 0; d < D; ++d) x[d] = ldf<D>(s, FL::X + d, p);
  }
  static __device__ __forceinline__ void load_v(const uint32_t* __restrict__ s, uint32_t p, float* v) {
#pragma unroll
    for (int d = 0; d < D; ++d) v[d] = ldf<D>(s, FL::V + d, p);
  }
  // coherent loads: from a set the same kernel also writes (G2P reads x at the sorted slot and stores the new x there)
  static __device__ __forceinline__ void load_x_rw(const uint32_t* s, uint32_t p, float* x) {
#pragma unroll
    for (int d = 0; d < D; ++d) x[d] = __uint_as_float(s[word<D>(FL::X + d, p)]);
  }
  static __device__ __forceinline__ void load_F(const uint32_t* __restrict__ s, uint32_t p, float* F) {
#pragma unroll
    for (int i = 0; i < D * D; ++i) F[i] = ldf<D>(s, FL::F + i, p);
  }
  static __device__ __forceinline__ void store_x(uint32_t* __restrict__ s, uint32_t p, const float* x) {
#pragma unroll
    for (int d = 0; d < D; ++d) stf<D>(s, FL::X + d, p, x[d]);
  }
  static __device__ __forceinline__ void store_v(uint32_t* __restrict__ s, uint32_t p, const float* v) {
#pragma unroll
    for (int d = 0; d < D; ++d) stf<D>(s, FL::V + d, p, v[d]);
  }
  static __device__ __forceinline__ void store_F(uint32_t* __restrict__ s, uint32_t p, const float* F) {
#pragma unroll
    for (int i = 0; i < D * D; ++i) stf<D>(s, FL::F + i, p, F[i]);
  }
  // what a store keeps (the fused kernel reads x and v back after storing them, ref :401-409)
  static __device__ __forceinline__ void round_x(float*) {}
  static __device__ __forceinline__ void round_v(float*) {}
  // round + keep the encoded words for the store that follows (one encode instead of three with packed storage)
  static __device__ __forceinline__ void round_x(float*, uint32_t*) {}
  static __device__ __forceinline__ void round_v(float*, uint32_t*) {}
  static __device__ __forceinline__ void put_x(uint32_t* __restrict__ s, uint32_t p, const float* x, const uint32_t*) { store_x(s, p, x); }
  static __device__ __forceinline__ void put_v(uint32_t* __restrict__ s, uint32_t p, const float* v, const uint32_t*) { store_v(s, p, v); }
};
template <int NW> struct PStoreQ3 {
  static constexpr int N = NW, JP = FldQ3::JP, TAG = FldQ3::TAG, C = FldQ3::C, X = FldQ3::X, XW = 2;
  static __device__ __forceinline__ size_t w(int f, uint32_t p) { return word_nf<NW>(f, p); }
  static __device__ __forceinline__ void load_x(const uint32_t* __restrict__ s, uint32_t p, float* x) {
    const uint32_t q[2] = {__ldg(s + w(FldQ3::X, p)), __ldg(s + w(FldQ3::X + 1, p))};
    decode_x3(q, x);
  }
  static __device__ __forceinline__ void load_v(const uint32_t* __restrict__ s, uint32_t p, float* v) {
    const uint32_t q[2] = {__ldg(s + w(FldQ3::V, p)), __ldg(s + w(FldQ3::V + 1, p))};
    decode_v3(q, v);
  }
  static __device__ __forceinline__ void load_x_rw(const uint32_t* s, uint32_t p, float* x) {
    const uint32_t q[2] = {s[w(FldQ3::X, p)], s[w(FldQ3::X + 1, p)]};
    decode_x3(q, x);
  }
  static __device__ __forceinline__ void load_F(const uint32_t* __restrict__ s, uint32_t p, float* F) {
    uint32_t q[5];
#pragma unroll
    for (int i = 0; i < 5; ++i) q[i] = __ldg(s + w(FldQ3::F + i, p));
    decode_F9(q, F);
  }
  static __device__ __forceinline__ void store_x(uint32_t* __restrict__ s, uint32_t p, const float* x) {
    uint32_t q[2];
    encode_x3(x, q);
    s[w(FldQ3::X, p)] = q[0]; s[w(FldQ3::X + 1, p)] = q[1];
  }
  static __device__ __forceinline__ void store_v(uint32_t* __restrict__ s, uint32_t p, const float* v) {
    uint32_t q[2];
    encode_v3(v, q);
    s[w(FldQ3::V, p)] = q[0]; s[w(FldQ3::V + 1, p)] = q[1];
  }
  static __device__ __forceinline__ void store_F(uint32_t* __restrict__ s, uint32_t p, const float* F) {
    uint32_t q[5];
    encode_F9(F, q);
#pragma unroll
    for (int i = 0; i < 5; ++i) s[w(FldQ3::F + i, p)] = q[i];
  }
  static __device__ __forceinline__ void round_x(float* x) { round_x3(x); }
  static __device__ __forceinline__ void round_v(float* v) { round_v3(v); }
  static __device__ __forceinline__ void round_x(float* x, uint32_t* q) { encode_x3(x, q); decode_x3(q, x); }
  static __device__ __forceinline__ void round_v(float* v, uint32_t* q) { encode_v3(v, q); decode_v3(q, v); }
  static __device__ __forceinline__ void put_x(uint32_t* __restrict__ s, uint32_t p, const float*, const uint32_t* q) {
    s[w(FldQ3::X, p)] = q[0]; s[w(FldQ3::X + 1, p)] = q[1];
  }
  static __device__ __forceinline__ void put_v(uint32_t* __restrict__ s, uint32_t p, const float*, const uint32_t* q) {
    s[w(FldQ3::V, p)] = q[0]; s[w(FldQ3::V + 1, p)] = q[1];
  }
};
template <> struct PStore<3, 1> : PStoreQ3<FldQ3::N> {};
template <> struct PStore<3, 2> : PStoreQ3<FldQ3::NC> {};
// run-time selection for the utility kernels (seeding, boxes, read-back): quant (0, 1, 2 = Q above) != 0 only exists in 3D
template <int D> __device__ __forceinline__ void load_x_rt(const uint32_t* __restrict__ s, int quant, uint32_t p, float* x) {
  if constexpr (D == 3) {
    if (quant == 1) { PStore<3, 1>::load_x(s, p, x); return; }
    if (quant == 2) { PStore<3, 2>::load_x(s, p, x); return; }
  }
  PStore<D, 0>::load_x(s, p, x);
}
template <int D> __device__ __forceinline__ void load_v_rt(const uint32_t* __restrict__ s, int quant, uint32_t p, float* v) {
  if constexpr (D == 3) {
    if (quant == 1) { PStore<3, 1>::load_v(s, p, v); return; }
    if (quant == 2) { PStore<3, 2>::load_v(s, p, v); return; }
  }
  PStore<D, 0>::load_v(s, p, v);
}
template <int D> __device__ __forceinline__ size_t tag_word_rt(int quant, uint32_t p) {
  if constexpr (D == 3) { if (quant) return word_rt(quant_words(quant), FldQ3::TAG, p); }
  return word<D>(Fld<D>::TAG, p);
}
template <int D> __device__ __forceinline__ uint32_t load_tag_rt(const uint32_t* __restrict__ s, int quant, uint32_t p) {
  return __ldg(s + tag_word_rt<D>(quant, p));
}

// virtual word `f` (read-back ABI numbering, Fld<D>::NV words) of the particle in storage slot s
template <int D> __device__ __forceinline__ uint32_t vword(const uint32_t* __restrict__ state, const Statics& st, int f, uint32_t s,
                                                          int quant = 0) {
  using FL = Fld<D>;
  if constexpr (D == 3) {
    if (quant) {             // packed storage: decode the group the word belongs to
      const int nw = quant_words(quant);
      if (f < FL::F) {
        float t[3];
        if (f < FL::V) load_x_rt<3>(state, quant, s, t); else load_v_rt<3>(state, quant, s, t);
        return __float_as_uint(t[f < FL::V ? f : f - FL::V]);
      }
      if (f < FL::C) {
        uint32_t q[5];
        float F[9];
#pragma unroll
        for (int i = 0; i < 5; ++i) q[i] = state[word_rt(nw, FldQ3::F + i, s)];
        decode_F9(q, F);
        return __float_as_uint(F[f - FL::F]);
      }
      if (f < FL::JP) return quant == 2 ? state[word_rt(nw, FldQ3::C + (f - FL::C), s)] : 0u;   // (no C with use_g2p2g: zeros)
      if (f == FL::JP) return state[word_rt(nw, FldQ3::JP, s)];
      const uint32_t tq = state[word_rt(nw, FldQ3::TAG, s)];
      if (f == FL::MAT) return tag_mat(tq);
      const uint32_t sq = tag_sid(tq);
      return f == FL::COLOR ? st.color[sq] : (f == FL::ID ? st.gid[sq] : st.emit[sq]);
    }
  }
  if (f <= FL::JP) return state[word<D>(f, s)];
  const uint32_t tag = state[word<D>(FL::TAG, s)];
  if (f == FL::MAT) return tag_mat(tag);
  const uint32_t sid = tag_sid(tag);
  return f == FL::COLOR ? st.color[sid] : (f == FL::ID ? st.gid[sid] : st.emit[sid]);
}

// base = int(floor(x * inv_dx - 0.5)), f32 multiply then f32 subtract, never
// fused, so the bin index is bit-identical to the oracle
// (engine/mpm_solver.py:357, 497, 703).
__device__ __forceinline__ int base_index(float x, float inv_dx) {
  return (int)floorf(__fsub_rn(__fmul_rn(x, inv_dx), 0.5f));
}

__device__ __forceinline__ void red_add_v4(float4* addr, float4 v) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(v.x), "f"(v.y),
               "f"(v.z), "f"(v.w)
               : "memory");
}

// Programmatic dependent launch (sm_90+): every kernel of the substep chain starts with
// pdl_enter(): wait until the grids it depends on have completed and flushed, then allow the
// NEXT kernel of the stream to be scheduled into the SM resources this grid frees while its last
// CTAs drain (that kernel parks in its own pdl_enter()).  Both are no-ops for ordinary launches.
__device__ __forceinline__ void pdl_enter() {
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

// L2 prefetch of [p, p + bytes): one cp.async.bulk.prefetch (sm_90+) per range
__device__ __forceinline__ void prefetch_l2_range(const void* p, uint32_t bytes) {
  const unsigned long long a = (unsigned long long)__cvta_generic_to_global(p);
  const unsigned long long a0 = a & ~15ull;
  const uint32_t sz = (bytes + (uint32_t)(a - a0) + 15u) & ~15u;
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(a0), "r"(sz) : "memory");
}
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// ---- peer path: producer kernels write straight into the neighbour's receive buffers
// over NVLink; a release-store of the substep epoch tells the neighbour the data is
// complete, an acquire-spin on the local epoch word (with a time-out) gates its consumer.
__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// every thread of the CTA waits (through thread 0) until the epoch words f0 / f1 (null = no neighbour on that
// side), written by the neighbours with st_release_sys, have reached `epoch`; 4 s time-out -> error bit, never a hang
__device__ __forceinline__ void cta_wait_epochs(const uint32_t* f0, const uint32_t* f1, uint32_t epoch, Status* st) {
  if (threadIdx.x == 0) {
    const unsigned long long t0 = global_ns();
    const uint32_t* f[2] = {f0, f1};
    for (int s = 0; s < 2; ++s) {
      if (!f[s]) continue;
      while ((int32_t)(ld_acquire_sys(f[s]) - epoch) < 0) {
        if (global_ns() - t0 > 4000000000ull) { atomicOr(&st->err, ERR_COMM_TIMEOUT); break; }
        __nanosleep(100);
      }
    }
  }
  __syncthreads();
}

// migration message headers: counts, overflow detection and (peer path) the "data ready" epoch.  They are
// written even after a device error (count 0) so that a neighbour never waits for ever.
__device__ __forceinline__ void publish_migration(const CommBufs& cb, uint32_t epoch, Status* st) {
  for (int s = 0; s < 2; ++s) {
    if (!cb.mig[s]) continue;
    int c = *(volatile int*)&st->err ? 0 : *(volatile int*)&st->mig_cnt[s];
    if (c > cb.mig_cap) { atomicOr(&st->err, ERR_COMM_CAPACITY); c = 0; }
    cb.mig[s][0] = (uint32_t)c;
    if (cb.flag_mig[s]) { __threadfence_system(); st_release_sys(cb.flag_mig[s], epoch); }
  }
  if (!*(volatile int*)&st->err) st->n_cur = *(volatile int*)&st->n_live;    // rows of the set G2P just wrote
}

template <int D> __device__ __forceinline__ int oct_delta_l(const KeyLayout& L, int o) {
  if constexpr (D == 3) return ((o & 1) ? L.eb[1] * L.eb[2] : 0) + ((o & 2) ? L.eb[2] : 0) + ((o & 4) ? 1 : 0);
  else return ((o & 1) ? L.eb[1] : 0) + ((o & 2) ? 1 : 0);
}

// linear block key (relative to the layout box) <-> relative block coords
template <int D> __device__ __forceinline__ void key_to_rel(const KeyLayout& L, uint32_t lin, int* rel) {
#pragma unroll
  for (int d = D - 1; d >= 0; --d) {
    rel[d] = (int)(lin % (uint32_t)L.eb[d]);
    lin /= (uint32_t)L.eb[d];
  }
}
template <int D> __device__ __forceinline__ uint32_t rel_to_key(const KeyLayout& L, const int* rel) {
  uint32_t lin = 0;
#pragma unroll
  for (int d = 0; d < D; ++d) lin = lin * (uint32_t)L.eb[d] + (uint32_t)rel[d];
  return lin;
}

// ------------------------------------------------------------------ reset/end
__global__ void k_reset(Status* st) {
  st->npb = 0; st->ngb_raw = 0; st->ngb = 0;
  st->work_p2g = 0; st->work_g2p = 0;
  st->maxv_bits = 0;
  st->maxgv_bits = 0;
  for (int d = 0; d < 3; ++d) { st->bb_min[d] = INT_MAX; st->bb_max[d] = INT_MIN; }
}
__global__ void k_batch_begin(Status* st, int n, int n_static) { st->n_cur = n; st->n_live = n; st->n_static = n_static; }
// first kernel of a substep whose keys came from the previous G2P: commit that substep, then
// let the errors its key pass raised take effect
__global__ void k_substep_begin(Status* st) {
  pdl_enter();
  if (!st->err) {
    st->done += 1;
    if (st->maxv_bits > st->maxv_all) st->maxv_all = st->maxv_bits;
  }
  st->err |= st->next_err;
  st->next_err = 0;
}
__global__ void k_end(Status* st) {
  if (!st->err) {
    st->done += 1;
    if (st->maxv_bits > st->maxv_all) st->maxv_all = st->maxv_bits;
  }
}

// ------------------------------------------------------------------ binning
template <int D>
__global__ void k_keys(const uint32_t* __restrict__ state, size_t cap, int n, float inv_dx, KeyLayout L,
                       uint32_t* __restrict__ keys, uint32_t* __restrict__ vals, Status* st) {
  using G = Geo<D>;
  if (st->err) return;
  for (uint32_t p = blockIdx.x * blockDim.x + threadIdx.x; p < (uint32_t)n; p += gridDim.x * blockDim.x) {
    uint32_t lin = 0, cell = 0;
    bool bad = false;
#pragma unroll
    for (int d = 0; d < D; ++d) {
      int g = base_index(ldf<D>(state, Fld<D>::X + d, p), inv_dx) + L.half;
      int rel = (g >> G::LOG_LEAF) - L.ob[d];
      if (rel < 0 || rel > L.eb[d] - 2) { bad = true; rel = min(max(rel, 0), L.eb[d] - 2); }
      lin = lin * (uint32_t)L.eb[d] + (uint32_t)rel;
      cell = (cell << G::LOG_LEAF) | (uint32_t)(g & (G::LEAF - 1));
    }
    keys[p] = (lin << G::CB) | cell;
    vals[p] = p;
    if (bad) atomicOr(&st->err, ERR_BBOX);
  }
}

// particle bounding box in global signed base-cell coordinates
template <int D>
__global__ void k_bbox(const uint32_t* __restrict__ state, size_t cap, int n, float inv_dx, Status* st, int quant) {
  int lo[D], hi[D];
#pragma unroll
  for (int d = 0; d < D; ++d) { lo[d] = INT_MAX; hi[d] = INT_MIN; }
  for (uint32_t p = blockIdx.x * blockDim.x + threadIdx.x; p < (uint32_t)n; p += gridDim.x * blockDim.x) {
    float xx[D];
    load_x_rt<D>(state, quant, p, xx);
#pragma unroll
    for (int d = 0; d < D; ++d) {
      int b = base_index(xx[d], inv_dx);
      lo[d] = min(lo[d], b); hi[d] = max(hi[d], b);
    }
  }
#pragma unroll
  for (int d = 0; d < D; ++d) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      lo[d] = min(lo[d], __shfl_xor_sync(0xffffffffu, lo[d], o));
      hi[d] = max(hi[d], __shfl_xor_sync(0xffffffffu, hi[d], o));
    }
    if ((threadIdx.x & 31) == 0 && lo[d] <= hi[d]) {
      atomicMin(&st->bb_min[d], lo[d]);
      atomicMax(&st->bb_max[d], hi[d]);
    }
  }
}

struct HeadOp {
  const uint32_t* keys;
  int cb;
  __device__ __forceinline__ bool operator()(const int& s) const {
    return s == 0 || (keys[s] >> cb) != (keys[s - 1] >> cb);
  }
};

// Output iterator that drops writes past `cap` (DeviceSelect has no capacity).
struct BoundedOut {
  int* ptr;
  int cap;
  struct Ref {
    int* p;
    __host__ __device__ __forceinline__ Ref& operator=(int v) { if (p) *p = v; return *this; }
    __host__ __device__ __forceinline__ Ref& operator=(const Ref&) { return *this; }
  };
  using iterator_category = std::random_access_iterator_tag;
  using value_type = int;
  using difference_type = ptrdiff_t;
  using pointer = int*;
  using reference = Ref;
  __host__ __device__ __forceinline__ Ref operator[](ptrdiff_t i) const {
    return Ref{(i >= 0 && i < cap) ? ptr + i : nullptr};
  }
  __host__ __device__ __forceinline__ Ref operator*() const { return (*this)[0]; }
  __host__ __device__ __forceinline__ BoundedOut operator+(ptrdiff_t i) const {
    return BoundedOut{ptr + i, cap - (int)i};
  }
};

__global__ void k_pb_finalize(Status* st, int* pb_start, int n, int max_blocks) {
  int npb = st->npb;
  if (npb > st->need_blocks) st->need_blocks = npb;
  if (npb > max_blocks) { st->err |= ERR_BLOCK_CAPACITY; return; }
  pb_start[npb] = n;
}

// One warp per particle block: OR of the octant masks of its particles
// (which of the 2^D blocks {b + o} the stencils base + {0,1,2}^D touch), then
// the candidate grid-block keys.  The union of candidates is exactly the
// reference's active-block set after P2G (engine/mpm_solver.py:582-584).
template <int D>
__global__ void k_pb_masks(const uint32_t* __restrict__ keys, const int* __restrict__ pb_start,
                           KeyLayout L, int max_blocks, uint32_t* __restrict__ pb_mask,
                           uint32_t* __restrict__ cand, uint32_t* __restrict__ pb_key, Status* st) {
  using G = Geo<D>;
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarp = (gridDim.x * blockDim.x) >> 5;
  const bool err = st->err != 0;
  const int npb = err ? 0 : st->npb;
  for (int b = warp; b < max_blocks; b += nwarp) {
    uint32_t out = INVALID_KEY;
    if (b < npb) {
      int start = pb_start[b], end = pb_start[b + 1];
      uint32_t m = 0;
      for (int s = start + lane; s < end; s += 32) {
        uint32_t cell = keys[s] & (G::CELLS - 1);
        uint32_t sp = 0;
#pragma unroll
        for (int d = 0; d < D; ++d) {
          uint32_t l = (cell >> (G::LOG_LEAF * (D - 1 - d))) & (G::LEAF - 1);
          sp |= (l >= (uint32_t)(G::LEAF - 2)) ? (1u << d) : 0u;
        }
#pragma unroll
        for (uint32_t o = 0; o < (uint32_t)G::NO; ++o)
          if ((o & ~sp) == 0) m |= 1u << o;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) m |= __shfl_xor_sync(0xffffffffu, m, o);
      if (lane == 0) { pb_mask[b] = m; pb_key[b] = keys[start] >> G::CB; }
      if (lane < G::NO && ((m >> lane) & 1u)) {
        int rel[D];
        key_to_rel<D>(L, keys[start] >> G::CB, rel);
#pragma unroll
        for (int d = 0; d < D; ++d) rel[d] += (lane >> d) & 1;
        out = rel_to_key<D>(L, rel);
      }
    }
    if (lane < G::NO) cand[b * G::NO + lane] = out;
  }
}

__global__ void k_gb_finalize(Status* st, const uint32_t* gb_key, int max_blocks) {
  if (st->err) return;
  int raw = st->ngb_raw;
  int ngb = (raw > 0 && gb_key[raw - 1] == INVALID_KEY) ? raw - 1 : raw;
  if (ngb > st->need_blocks) st->need_blocks = ngb;
  if (ngb > max_blocks) { st->err |= ERR_BLOCK_CAPACITY; ngb = 0; }
  st->ngb = ngb;
}

template <int D>
__global__ void k_nbr(const uint32_t* __restrict__ keys, const int* __restrict__ pb_start,
                      const uint32_t* __restrict__ pb_mask, const uint32_t* __restrict__ gb_key,
                      KeyLayout L, int* __restrict__ pb_nbr, const Status* st) {
  using G = Geo<D>;
  if (st->err) return;
  const int npb = st->npb, ngb = st->ngb;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < npb * G::NO; i += gridDim.x * blockDim.x) {
    int b = i / G::NO, o = i % G::NO;
    int slot = -1;
    if ((pb_mask[b] >> o) & 1u) {
      int rel[D];
      key_to_rel<D>(L, keys[pb_start[b]] >> G::CB, rel);
#pragma unroll
      for (int d = 0; d < D; ++d) rel[d] += (o >> d) & 1;
      uint32_t k = rel_to_key<D>(L, rel);
      int lo = 0, hi = ngb - 1;
      while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (gb_key[mid] < k) lo = mid + 1; else hi = mid;
      }
      slot = lo;   // present by construction
    }
    pb_nbr[i] = slot;
  }
}

template <int D>
__global__ void k_clear_grid(float4* __restrict__ grid, const Status* st, int* __restrict__ z1, int n1,
                             int* __restrict__ z2, int n2, int n1_blocks_mul, float4* __restrict__ zp0,
                             float4* __restrict__ zp1, int nzp) {
  pdl_enter();
  // fused halo: the planes the neighbours fill during the NEXT substep (the ones of this substep are being
  // filled right now, the previous ones were read by the last grid op) -- see CommBufs
  {
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nzp; i += gridDim.x * blockDim.x) {
      if (zp0) zp0[i] = z;
      if (zp1) zp1[i] = z;
    }
  }
  // tables the NEXT substep of the batch expects zeroed (per-cell counts; block flags when the
  // coming G2P emits them), so that no memset node sits between the kernels of the chain
  const int gtid = blockIdx.x * blockDim.x + threadIdx.x, gsz = gridDim.x * blockDim.x;
  if (n1_blocks_mul) n1 = min(n1, st->npb * n1_blocks_mul + 1);   // only the cells of existing blocks were counted
  for (int i = gtid; i < n1; i += gsz) z1[i] = 0;
  for (int i = gtid; i < n2; i += gsz) z2[i] = 0;
  if (st->err) return;
  const size_t total = (size_t)st->ngb * Geo<D>::CELLS;
  const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x)
    grid[i] = z;
}

// tile node index -> (octant, cell inside that leaf)
template <int D> __device__ __forceinline__ void tile_node(int n, int& oct, int& cell) {
  using G = Geo<D>;
  int c[D];
#pragma unroll
  for (int d = D - 1; d >= 0; --d) { c[d] = n % G::T; n /= G::T; }
  oct = 0; cell = 0;
#pragma unroll
  for (int d = 0; d < D; ++d) {
    int hi = c[d] >= G::LEAF;
    oct |= hi << d;
    cell = (cell << G::LOG_LEAF) | (c[d] - hi * G::LEAF);
  }
}

template <int D> struct SubstepArgs {
  const uint32_t* src;   // live state set
  uint32_t* dst;         // other set
  size_t cap;
  const uint32_t* keys;  // sorted keys
  const uint32_t* perm;  // sorted position -> slot in src
  const int* pb_start;
  const int* pb_nbr;
  const uint32_t* pb_key;   // linear leaf-block key of each particle block
  const int* cellstart;     // [npb*CELLS+1] first sorted position of every cell (counting-sort path), or null
  float4* grid;
  Status* st;
  KeyLayout L;
  Consts K;
  float dt;
  uint32_t* next_keys;  // non-null: G2P also emits the next substep's sort keys and block flags
  int* next_flags;      //   (same key layout, see mpm_bin.cuh); saves the k_bin_keys pass
  int next_nlin;
  int pf_mode;    // next-block L2 prefetch: 0 off, 1 one prefetch per 128 B, 2 bulk range prefetch, 3 one per 32 B
  int n_rows;     // g2p2g: rows of the live set (rows >= pb_start[npb] were added after the binning)
  Statics stat;   // static side arrays (G2P reads them for the particles it hands to a neighbour rank)
  Slab slab;      // multi-GPU: this rank's block columns (mpm_comm.cuh)
  CommBufs cb;    // multi-GPU: migration / halo send buffers
};

// ------------------------------------------------------------------ P2G
// One CTA per particle block (dynamic queue).  The (LEAF+2)^D node tile is
// accumulated in shared memory and flushed with 128-bit vector reductions
// (REDG.E.ADD.F32x4) into the up-to-2^D leaf blocks it overlaps.
// Grid node record: (momentum[D], mass) padded to float4.
constexpr int P2G_THREADS = 256;
template <int D>
__global__ void __launch_bounds__(P2G_THREADS) k_p2g(SubstepArgs<D> a) {
  using G = Geo<D>;
  using FL = Fld<D>;
  __shared__ float4 tile[G::TN];
  __shared__ int s_b;
  __shared__ int s_nbr[G::NO];
  if (a.st->err) return;
  const int npb = a.st->npb;
  const int tid = threadIdx.x;
  const size_t cap = a.cap;
  for (;;) {
    if (tid == 0) s_b = atomicAdd(&a.st->work_p2g, 1);
    __syncthreads();
    const int b = s_b;
    if (b >= npb) break;
    const int start = a.pb_start[b], end = a.pb_start[b + 1];
    for (int n = tid; n < G::TN; n += P2G_THREADS) tile[n] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (tid < G::NO) s_nbr[tid] = a.pb_nbr[b * G::NO + tid];
    int org[D];   // absolute cell coordinate of the block origin
    {
      int rel[D];
      key_to_rel<D>(a.L, a.pb_key[b], rel);
#pragma unroll
      for (int d = 0; d < D; ++d) org[d] = (rel[d] + a.L.ob[d]) << G::LOG_LEAF;
    }
    __syncthreads();
    for (int s = start + tid; s < end; s += P2G_THREADS) {
      const uint32_t p = a.perm[s];
      float x[D], v[D], fx[D], w[3][D];
      int l[D];
#pragma unroll
      for (int d = 0; d < D; ++d) {
        x[d] = ldf<D>(a.src, FL::X + d, p);
        v[d] = ldf<D>(a.src, FL::V + d, p);
        int base = base_index(x[d], a.K.inv_dx);
        fx[d] = x[d] * a.K.inv_dx - (float)base;                 // :503
        l[d] = min(max(base + a.L.half - org[d], 0), G::LEAF - 1);
        w[0][d] = 0.5f * (1.5f - fx[d]) * (1.5f - fx[d]);        // :505
        w[1][d] = 0.75f - (fx[d] - 1.0f) * (fx[d] - 1.0f);
        w[2][d] = 0.5f * (fx[d] - 0.5f) * (fx[d] - 0.5f);
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
      float mv[D];
#pragma unroll
      for (int d = 0; d < D; ++d) mv[d] = mass * v[d];
      // scatter (:577-584)
      if constexpr (D == 3) {
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
          for (int j = 0; j < 3; ++j)
#pragma unroll
            for (int k = 0; k < 3; ++k) {
              float dp0 = ((float)i - fx[0]) * a.K.dx, dp1 = ((float)j - fx[1]) * a.K.dx,
                    dp2 = ((float)k - fx[2]) * a.K.dx;
              float wt = w[i][0] * w[j][1] * w[k][2];
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
            float dp0 = ((float)i - fx[0]) * a.K.dx, dp1 = ((float)j - fx[1]) * a.K.dx;
            float wt = w[i][0] * w[j][1];
            float* t = reinterpret_cast<float*>(&tile[(l[0] + i) * G::T + (l[1] + j)]);
            atomicAdd(t + 0, wt * (mv[0] + aff[0] * dp0 + aff[1] * dp1));
            atomicAdd(t + 1, wt * (mv[1] + aff[2] * dp0 + aff[3] * dp1));
            atomicAdd(t + 2, wt * mass);
          }
      }
    }
    __syncthreads();
    for (int n = tid; n < G::TN; n += P2G_THREADS) {
      float4 val = tile[n];
      float m = (D == 3) ? val.w : val.z;
      if (m != 0.0f) {
        int oct, cell;
        tile_node<D>(n, oct, cell);
        int slot = s_nbr[oct];
        if (slot >= 0) red_add_v4(a.grid + (size_t)slot * G::CELLS + cell, val);
      }
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------ grid op
// grid_normalization_and_gravity + grid_bounding_box + every collider, fused
// (engine/mpm_solver.py:586-616, 618-687), one thread per cell of every
// active leaf block.  Node record becomes (velocity[D], mass).
template <int D>
__global__ void k_grid_op(float4* __restrict__ grid, const uint32_t* __restrict__ gb_key, KeyLayout L,
                          const ColliderTable* __restrict__ ct, Grav grav, GridCfg cfg, float dx, float dt,
                          float v_allowed, Status* st, int* __restrict__ zero, int nzero, Slab slab, CommBufs cb) {
  using G = Geo<D>;
  pdl_enter();
  // (slab runs) the block-flag table the coming G2P fills for the next substep
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nzero; i += gridDim.x * blockDim.x) zero[i] = 0;
  // fused halo: the neighbours' partial sums of the shared columns arrive in my planes while P2G runs; they are
  // complete once both have published this substep's epoch
  const float4* pl_lo = nullptr;
  const float4* pl_hi = nullptr;
  if (cb.fused) {
    cta_wait_epochs(cb.wait_halo[0], cb.wait_halo[1], cb.epoch + 1u, st);
    const size_t off = (size_t)(cb.epoch % 3u) * cb.plane_blocks * HALO_SLAB;
    if (cb.plane_in[0]) pl_lo = cb.plane_in[0] + off;
    if (cb.plane_in[1]) pl_hi = cb.plane_in[1] + off;
  }
  if (st->err) return;
  float gvmax = 0.0f;
  const size_t total = (size_t)st->ngb * G::CELLS;
  const int ncol = ct->n;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int slot = (int)(i / G::CELLS);
    const int cell = (int)(i % G::CELLS);
    float4 rec = grid[i];
    float v[3] = {rec.x, rec.y, (D == 3) ? rec.z : 0.0f};
    int rel[D], I[3] = {0, 0, 0};
    key_to_rel<D>(L, gb_key[slot], rel);
#pragma unroll
    for (int d = 0; d < D; ++d) {
      int lc = (cell >> (G::LOG_LEAF * (D - 1 - d))) & (G::LEAF - 1);
      I[d] = ((rel[d] + L.ob[d]) << G::LOG_LEAF) + lc - L.half;
    }
    float m = (D == 3) ? rec.w : rec.z;
    if constexpr (D == 3) {
      if (cb.fused) {
        const int col = rel[0] + L.ob[0];
        const float4* pl = col == slab.lo ? pl_lo : (col == slab.hi ? pl_hi : nullptr);
        const int cx = cell >> 4, cy = (cell >> 2) & 3, cz = cell & 3;
        if (pl && cx <= 1) {
          // the neighbour's particle blocks (y - dy, z - dz) whose tiles cover this node: tile node (cy + 4 dy, cz + 4 dz)
#pragma unroll
          for (int dy = 0; dy < 2; ++dy)
#pragma unroll
            for (int dz = 0; dz < 2; ++dz) {
              if ((dy && cy > 1) || (dz && cz > 1) || rel[1] - dy < 0 || rel[2] - dz < 0) continue;
              const float4 h = pl[(size_t)((rel[1] - dy) * L.eb[2] + (rel[2] - dz)) * HALO_SLAB + cx * 36 + (cy + 4 * dy) * 6 +
                                  (cz + 4 * dz)];
              v[0] += h.x; v[1] += h.y; v[2] += h.z; m += h.w;
            }
        }
      }
    }
    if (m > 0.0f) {                                            // :591-593
      float inv = __fdiv_rn(1.0f, m);
#pragma unroll
      for (int d = 0; d < D; ++d) v[d] = __fadd_rn(__fmul_rn(inv, v[d]), __fmul_rn(dt, grav.g[d]));
    }
    if (v_allowed > 0.0f) {                                    // g2p2g grid-velocity clamp (:596-598)
#pragma unroll
      for (int d = 0; d < D; ++d) v[d] = fminf(fmaxf(v[d], -v_allowed), v_allowed);
    }
    for (int c = 0; c < ncol; ++c) {
      const ColliderDev& col = ct->c[c];
      if (col.kind == 0) {                                     // :600-616
#pragma unroll
        for (int d = 0; d < D; ++d) {
          int lo = col.unbounded ? -cfg.grid_size / 2 + cfg.padding : cfg.padding;
          int hi = col.unbounded ? cfg.grid_size / 2 - cfg.padding : cfg.res[d] - cfg.padding;
          if (I[d] < lo && v[d] < 0.0f) v[d] = 0.0f;
          if (I[d] >= hi && v[d] > 0.0f) v[d] = 0.0f;
        }
      } else if (col.kind == 1) {                              // sphere :621-640
        float off[3] = {0, 0, 0}, nsq = 0.0f;
#pragma unroll
        for (int d = 0; d < D; ++d) {
          off[d] = __fsub_rn(__fmul_rn((float)I[d], dx), col.a[d]);
          nsq = __fadd_rn(nsq, __fmul_rn(off[d], off[d]));
        }
        if (nsq < col.r2) {
          if (col.surface == 0) {
#pragma unroll
            for (int d = 0; d < D; ++d) v[d] = 0.0f;
          } else {
            float invn = __fdiv_rn(1.0f, __fadd_rn(sqrtf(nsq), 1e-5f));
            float nrm[3] = {0, 0, 0}, nc = 0.0f;
#pragma unroll
            for (int d = 0; d < D; ++d) { nrm[d] = __fmul_rn(off[d], invn); nc = __fadd_rn(nc, __fmul_rn(nrm[d], v[d])); }
            float k = (col.surface == 1) ? nc : fminf(nc, 0.0f);
#pragma unroll
            for (int d = 0; d < D; ++d) v[d] = __fsub_rn(v[d], __fmul_rn(nrm[d], k));
          }
        }
      } else {                                                 // plane :660-685
        float dotn = 0.0f;
#pragma unroll
        for (int d = 0; d < D; ++d)
          dotn = __fadd_rn(dotn, __fmul_rn(__fsub_rn(__fmul_rn((float)I[d], dx), col.a[d]), col.b[d]));
        if (dotn < 0.0f) {
          if (col.surface == 0) {
#pragma unroll
            for (int d = 0; d < D; ++d) v[d] = 0.0f;
          } else {
            float nc = 0.0f;
#pragma unroll
            for (int d = 0; d < D; ++d) nc = __fadd_rn(nc, __fmul_rn(v[d], col.b[d]));
            float k = (col.surface == 1) ? nc : fminf(nc, 0.0f);
            float nsq = 0.0f;
#pragma unroll
            for (int d = 0; d < D; ++d) { v[d] = __fsub_rn(v[d], __fmul_rn(col.b[d], k)); nsq = __fadd_rn(nsq, __fmul_rn(v[d], v[d])); }
            float norm = sqrtf(nsq);
            if (nc < 0.0f && norm > 1e-30f) {                  // :679-683
              float invn = __fdiv_rn(1.0f, norm);
              float sc = fmaxf(0.0f, __fadd_rn(norm, __fmul_rn(nc, col.friction)));
#pragma unroll
              for (int d = 0; d < D; ++d) v[d] = __fmul_rn(__fmul_rn(v[d], invn), sc);
            }
          }
        }
      }
    }
    if (D == 3) grid[i] = make_float4(v[0], v[1], v[2], m);
    else grid[i] = make_float4(v[0], v[1], m, 0.0f);
#pragma unroll
    for (int d = 0; d < D; ++d) gvmax = fmaxf(gvmax, fabsf(v[d]));
  }
  // compute_max_grid_velocity (engine/mpm_solver.py:737-746), folded into the pass
  __shared__ unsigned s_gv;
  if (threadIdx.x == 0) s_gv = 0u;
  __syncthreads();
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) gvmax = fmaxf(gvmax, __shfl_xor_sync(0xffffffffu, gvmax, o));
  if ((threadIdx.x & 31) == 0 && gvmax > 0.0f) atomicMax(&s_gv, __float_as_uint(gvmax));
  __syncthreads();
  if (threadIdx.x == 0 && s_gv) atomicMax(&st->maxgv_bits, s_gv);
}

// ------------------------------------------------------------------ G2P
// Gather from the staged velocity tile, update v, C, x (engine/mpm_solver.py:
// 694-724), write the particle to its sorted slot in the other state set, and
// fold compute_max_velocity (:726-735) and the next bounding box into the pass.
// XS: P2G left x and the tag at the sorted slot of the other set (every P2G kernel except the halo variant of k_p2g3, which
// is at its register limit and measured 3-7 % slower with the four extra stores): G2P streams them.  Otherwise it follows
// perm to the live set, as P2G does, and moves the tag itself.
template <int D, int G2P_THREADS, int G2P_MINB, bool BULK = false, int QM = 0, bool XS = true>
__global__ void __launch_bounds__(G2P_THREADS, G2P_MINB) k_g2p(SubstepArgs<D> a) {
  using G = Geo<D>;
  using FL = Fld<D>;
  using P = PStore<D, QM>;      // QM = 2: quant=True, packed x / v (a store rounds, ref :106-111, 720-724) + f32 C
  // The velocity tile of the NEXT block is fetched with cp.async into the other buffer while this
  // block's particles are gathered: one CTA barrier per block, no exposed grid-load latency.
  __shared__ float4 tile_buf[2][G::TN];
  __shared__ int s_b;
  pdl_enter();
  if (a.st->err) {
    if (a.cb.fused && blockIdx.x == 0 && threadIdx.x == 0) publish_migration(a.cb, a.cb.epoch + 1u, a.st);
    return;
  }
  const int npb = a.st->npb;
  const int tid = threadIdx.x;
  const size_t cap = a.cap;
  float vmax = 0.0f;
  int lo[D], hi[D];
#pragma unroll
  for (int d = 0; d < D; ++d) { lo[d] = INT_MAX; hi[d] = INT_MIN; }
  __shared__ int s_next[2];
  // fused key pass: which spill patterns (bit = 1 << nsp) the particles that STAY in this CTA's
  // current block show; expanded to block flags once per block instead of once per particle
  __shared__ unsigned s_seen[2];
  if (tid == 0) { s_b = atomicAdd(&a.st->work_g2p, 1); s_seen[0] = 0u; s_seen[1] = 0u; }
  __syncthreads();
  int b = s_b;
  // asynchronous copy of block blk's (LEAF+2)^D node tile (zero-filled where no grid block exists)
  auto stage_tile = [&](int blk, float4* dstt) {
    for (int n = tid; n < G::TN; n += G2P_THREADS) {
      int oct, cell;
      tile_node<D>(n, oct, cell);
      const int slot = a.pb_nbr[blk * G::NO + oct];
      const float4* src = a.grid + (slot >= 0 ? (size_t)slot * G::CELLS + cell : 0);
      const unsigned daddr = (unsigned)__cvta_generic_to_shared(dstt + n);
      const int nbytes = slot >= 0 ? 16 : 0;
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(daddr), "l"(src), "r"(nbytes) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  // the same copy as bulk (TMA) transfers: the tile's z-rows are LEAF nodes of one leaf block followed by 2 nodes of its
  // +z neighbour, i.e. one 64-byte and one 32-byte cp.async.bulk per row, completing on the buffer's mbarrier
  constexpr int TROWS = G::TN / G::T;
  __shared__ __align__(8) unsigned long long s_bar[2];
  constexpr bool bulk = BULK;   // MPM_G2P_TILE=1: measured 4 % slower than the per-node cp.async
  auto stage_tile_bulk = [&](int blk, float4* dstt, unsigned long long* bar) {
    if (tid < 2 * TROWS) {
      const int row = tid >> 1, part = tid & 1;
      int oct = part << (D - 1), cell = 0, r = row, c[D > 1 ? D - 1 : 1];
#pragma unroll
      for (int d = D - 2; d >= 0; --d) { c[d] = r % G::T; r /= G::T; }
#pragma unroll
      for (int d = 0; d < D - 1; ++d) {
        const int hi = c[d] >= G::LEAF;
        oct |= hi << d;
        cell = (cell << G::LOG_LEAF) | (c[d] - hi * G::LEAF);
      }
      cell <<= G::LOG_LEAF;
      const int slot = a.pb_nbr[blk * G::NO + oct];
      const unsigned baddr = (unsigned)__cvta_generic_to_shared(bar);
      if (slot >= 0) {
        const unsigned bytes = part ? (G::T - G::LEAF) * 16u : G::LEAF * 16u;
        const unsigned daddr = (unsigned)__cvta_generic_to_shared(dstt + row * G::T + part * G::LEAF);
        const float4* src = a.grid + (size_t)slot * G::CELLS + cell;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(baddr), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(daddr), "l"(src), "r"(bytes), "r"(baddr) : "memory");
      } else {
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(baddr) : "memory");
      }
    }
  };
  if (bulk) {
    if (tid == 0) {
#pragma unroll
      for (int i = 0; i < 2; ++i)
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(&s_bar[i])), "r"(2 * TROWS) : "memory");
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
  }
  unsigned bar_phase = 0u;   // bit u: parity the next wait on buffer u expects
  if (b < npb) { if (bulk) stage_tile_bulk(b, tile_buf[0], &s_bar[0]); else stage_tile(b, tile_buf[0]); }
  int u = 0;
  uint32_t prev_lin = 0u;
  bool have_prev = false;
  // flags of a finished block from the spill patterns its staying particles showed: octant o of
  // the stencil union is touched iff some particle spills along every axis of o
  auto write_flags = [&](uint32_t lin, unsigned* seen_slot) {
    if (tid < G::NO) {
      const unsigned sn = *seen_slot;
      unsigned sup = 0u;
#pragma unroll
      for (unsigned q = 0; q < (unsigned)G::NO; ++q)
        if (((unsigned)tid & ~q) == 0u) sup |= 1u << q;
      if (sn & sup) a.next_flags[a.next_nlin + (int)lin + oct_delta_l<D>(a.L, tid)] = 1;
      __syncwarp((1u << G::NO) - 1u);
      if (tid == 0 && sn) { a.next_flags[lin] = 1; *seen_slot = 0u; }
    }
  };
  while (b < npb) {
    if (tid == 0) s_next[u] = atomicAdd(&a.st->work_g2p, 1);   // one block ahead
    const int start = a.pb_start[b], end = a.pb_start[b + 1];
    const uint32_t cur_lin = a.pb_key[b];
    unsigned seen = 0u;
    int org[D];
    {
      int rel[D];
      key_to_rel<D>(a.L, cur_lin, rel);
#pragma unroll
      for (int d = 0; d < D; ++d) org[d] = (rel[d] + a.L.ob[d]) << G::LOG_LEAF;
    }
    // software pipeline over this thread's particles: the position / tag loads run one iteration ahead of the arithmetic
    int s = start + tid;
    float xn[D];
    uint32_t matn = 0;
#pragma unroll
    for (int d = 0; d < D; ++d) xn[d] = 0.0f;
    uint32_t p1 = 0u, p2 = 0u;      // (!XS) `perm` runs two iterations ahead, the loads one
    if constexpr (XS) {
      // P2G left x and the tag of every particle at its SORTED slot of the other set: streaming reads, no perm -> x chain
      if (s < end) {
        P::load_x_rw(a.dst, s, xn);
        matn = a.dst[P::w(P::TAG, s)];
      }
    } else {
      p1 = s < end ? a.perm[s] : 0u;
      p2 = s + G2P_THREADS < end ? a.perm[s + G2P_THREADS] : 0u;
      if (s < end) {
        P::load_x(a.src, p1, xn);
        matn = __ldg(a.src + P::w(P::TAG, p1));
      }
    }
    float4* tile = tile_buf[u];
    if (bulk) {
      const unsigned baddr = (unsigned)__cvta_generic_to_shared(&s_bar[u]), par = (bar_phase >> u) & 1u;
      unsigned done = 0u;
      while (!done)
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(done) : "r"(baddr), "r"(par) : "memory");
      bar_phase ^= 1u << u;
    } else {
      asm volatile("cp.async.wait_all;" ::: "memory");
    }
    if constexpr (D == 3) {
      // (vx, vy | vz, vz): the gather below works on packed pairs; every thread patches the nodes it copied
      for (int n = tid; n < G::TN; n += G2P_THREADS) tile[n].w = tile[n].z;
    }
    // bulk copies write through the async proxy: order this CTA's generic accesses to both buffers before them
    if (bulk) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();   // tile complete; s_next visible; every warp is done with the previous block
    if (a.next_keys && have_prev) write_flags(prev_lin, &s_seen[u ^ 1]);
    const int nb_claim = s_next[u];
    if (nb_claim < npb) {
      if (bulk) stage_tile_bulk(nb_claim, tile_buf[u ^ 1], &s_bar[u ^ 1]); else stage_tile(nb_claim, tile_buf[u ^ 1]);
    }
    {   // next block's particle rows towards L2 while this one computes
      const int nb = nb_claim;
      if (nb < npb && a.pf_mode) {
        // the rows G2P reads: x and the tag of the block's own sorted slots (two runs of 128-byte rows per tile)
        const int ns = a.pb_start[nb], ne = a.pb_start[nb + 1];
        const int t0 = ns >> TILE_LOG, nt = ((ne - 1) >> TILE_LOG) - t0 + 1;
        for (int i = tid; i < 2 * nt; i += G2P_THREADS) {
          const uint32_t* tile0 = (XS ? a.dst : a.src) + (size_t)(t0 + (i >> 1)) * P::N * TILE;
          if (i & 1) prefetch_l2_range(tile0 + P::TAG * TILE, TILE * 4u);
          else prefetch_l2_range(tile0 + P::X * TILE, (uint32_t)P::XW * TILE * 4u);
        }
      }
    }
    for (; s < end; s += G2P_THREADS) {
      float x[D], fx[D], w[3][D];
      int l[D];
#pragma unroll
      for (int d = 0; d < D; ++d) x[d] = xn[d];
      const uint32_t tag = matn, mat = tag_mat(tag);      // material | static row: the only immutable word that travels
      const uint32_t pcur = p1;
      if constexpr (XS) {
        if (s + G2P_THREADS < end) {
          P::load_x_rw(a.dst, s + G2P_THREADS, xn);
          matn = a.dst[P::w(P::TAG, s + G2P_THREADS)];
        }
      } else {
        p1 = p2;
        if (s + G2P_THREADS < end) {
          P::load_x(a.src, p1, xn);
          matn = __ldg(a.src + P::w(P::TAG, p1));
        }
        p2 = s + 2 * G2P_THREADS < end ? a.perm[s + 2 * G2P_THREADS] : 0u;
      }
#pragma unroll
      for (int d = 0; d < D; ++d) {
        int base = base_index(x[d], a.K.inv_dx);
        fx[d] = x[d] * a.K.inv_dx - (float)base;
        l[d] = min(max(base + a.L.half - org[d], 0), G::LEAF - 1);
        w[0][d] = 0.5f * (1.5f - fx[d]) * (1.5f - fx[d]);
        w[1][d] = 0.75f - (fx[d] - 1.0f) * (fx[d] - 1.0f);
        w[2][d] = 0.5f * (fx[d] - 0.5f) * (fx[d] - 0.5f);
      }
      // Tensor-product evaluation of sum_ijk w_i w_j w_k g_ijk and of its first
      // moments (C = 4 inv_dx sum w g (x) (o - fx), :715-721): partial sums along
      // z, then y, then x -- 279 FMAs instead of 27 x 23, issued as 126 FFMA2 + 21 FFMA in 3D.
      float nv[D], nC[D * D], mw[3][D];
#pragma unroll
      for (int d = 0; d < D; ++d)
#pragma unroll
        for (int i = 0; i < 3; ++i) mw[i][d] = w[i][d] * ((float)i - fx[d]);
      if constexpr (D == 3) {
        // packed pairs (fma.rn.f32x2 -> FFMA2): .xy of every partial sum in one register pair, the
        // z components of the plain and the z-moment sums in another
        float2 wz2[3], mz2[3], wmz[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          wz2[k] = make_float2(w[k][2], w[k][2]);
          mz2[k] = make_float2(mw[k][2], mw[k][2]);
          wmz[k] = make_float2(w[k][2], mw[k][2]);
        }
        float2 nvxy = make_float2(0.f, 0.f), cxxy = nvxy, cyxy = nvxy, czxy = nvxy, nvcz = nvxy;   // nvcz = (nv.z, cz.z)
        float cxz = 0.f, cyz = 0.f;
#pragma unroll
        for (int i = 0; i < 3; ++i) {
          float2 B00 = make_float2(0.f, 0.f), B10 = B00, B01 = B00, Bz = B00;   // Bz = (B00.z, B01.z)
          float B10z = 0.f;
#pragma unroll
          for (int j = 0; j < 3; ++j) {
            float2 A0 = make_float2(0.f, 0.f), A1 = A0, Az = A0;              // Az = (A0.z, A1.z)
#pragma unroll
            for (int k = 0; k < 3; ++k) {
              const float4 g = tile[((l[0] + i) * G::T + (l[1] + j)) * G::T + (l[2] + k)];
              A0 = __ffma2_rn(wz2[k], make_float2(g.x, g.y), A0);
              A1 = __ffma2_rn(mz2[k], make_float2(g.x, g.y), A1);
              Az = __ffma2_rn(wmz[k], make_float2(g.z, g.w), Az);
            }
            const float2 wj = make_float2(w[j][1], w[j][1]), mj = make_float2(mw[j][1], mw[j][1]);
            B00 = __ffma2_rn(wj, A0, B00);
            B10 = __ffma2_rn(mj, A0, B10);
            B01 = __ffma2_rn(wj, A1, B01);
            Bz = __ffma2_rn(wj, Az, Bz);
            B10z = fmaf(mw[j][1], Az.x, B10z);
          }
          const float2 wi = make_float2(w[i][0], w[i][0]), mi = make_float2(mw[i][0], mw[i][0]);
          nvxy = __ffma2_rn(wi, B00, nvxy);
          cxxy = __ffma2_rn(mi, B00, cxxy);
          cyxy = __ffma2_rn(wi, B10, cyxy);
          czxy = __ffma2_rn(wi, B01, czxy);
          nvcz = __ffma2_rn(wi, Bz, nvcz);
          cxz = fmaf(mw[i][0], Bz.x, cxz);
          cyz = fmaf(w[i][0], B10z, cyz);
        }
        nv[0] = nvxy.x; nv[1] = nvxy.y; nv[2] = nvcz.x;
        const float f4 = a.K.four_inv_dx;
        nC[0] = f4 * cxxy.x; nC[1] = f4 * cyxy.x; nC[2] = f4 * czxy.x;
        nC[3] = f4 * cxxy.y; nC[4] = f4 * cyxy.y; nC[5] = f4 * czxy.y;
        nC[6] = f4 * cxz;    nC[7] = f4 * cyz;    nC[8] = f4 * nvcz.y;
      } else {
        float cx[2] = {0.f, 0.f}, cy[2] = {0.f, 0.f};
        nv[0] = nv[1] = 0.0f;
#pragma unroll
        for (int i = 0; i < 3; ++i) {
          float A0[2] = {0.f, 0.f}, A1[2] = {0.f, 0.f};
#pragma unroll
          for (int j = 0; j < 3; ++j) {
            const float4 g = tile[(l[0] + i) * G::T + (l[1] + j)];
            A0[0] += w[j][1] * g.x; A0[1] += w[j][1] * g.y;
            A1[0] += mw[j][1] * g.x; A1[1] += mw[j][1] * g.y;
          }
#pragma unroll
          for (int r = 0; r < 2; ++r) {
            nv[r] += w[i][0] * A0[r];
            cx[r] += mw[i][0] * A0[r];
            cy[r] += w[i][0] * A1[r];
          }
        }
#pragma unroll
        for (int r = 0; r < 2; ++r) {
          nC[r * 2 + 0] = a.K.four_inv_dx * cx[r];
          nC[r * 2 + 1] = a.K.four_inv_dx * cy[r];
        }
      }
      if (mat == (uint32_t)STATIONARY) {                       // :722
        // The x + (-0) identity consumes each load INSIDE this rare branch: otherwise the stores after
        // the join wait on a scoreboard shared with the next particle's prefetched loads.
        const uint32_t p = XS ? a.perm[s] : pcur;
        P::load_v(a.src, p, nv);
#pragma unroll
        for (int d = 0; d < D; ++d) nv[d] = __fadd_rn(nv[d], -0.0f);
        if (!a.K.g2p2g) {      // [g2p2g] C is a register value there: the gathered C is used (:385, 414)
#pragma unroll
          for (int i = 0; i < D * D; ++i) nC[i] = __fadd_rn(__uint_as_float(__ldg(a.src + P::w(P::C + i, p))), -0.0f);
        }
      }
      uint32_t qx[2], qv[2];                                                          // packed storage: the stored words
      P::round_v(nv, qv);                                                             // self.v[p] = new_v rounds (:723)
      if (mat != (uint32_t)STATIONARY) {
#pragma unroll
        for (int d = 0; d < D; ++d) x[d] = __fadd_rn(x[d], __fmul_rn(a.dt, nv[d]));   // :724, reads v[p] back
      }
      P::round_x(x, qx);
      uint32_t nlin_key = 0, ncell = 0, nsp = 0;
      bool nbad = false;
      P::put_x(a.dst, s, x, qx);
      P::put_v(a.dst, s, nv, qv);
#pragma unroll
      for (int d = 0; d < D; ++d) {
        vmax = fmaxf(vmax, fabsf(nv[d]));
        int nb = base_index(x[d], a.K.inv_dx);
        lo[d] = min(lo[d], nb); hi[d] = max(hi[d], nb);
        // next substep's bin of this particle (same arithmetic as k_bin_keys)
        const int g = nb + a.L.half;
        int rel = (g >> G::LOG_LEAF) - a.L.ob[d];
        if (rel < 0 || rel > a.L.eb[d] - 2) { nbad = true; rel = min(max(rel, 0), a.L.eb[d] - 2); }
        nlin_key = nlin_key * (uint32_t)a.L.eb[d] + (uint32_t)rel;
        const uint32_t lc = (uint32_t)(g & (G::LEAF - 1));
        ncell = (ncell << G::LOG_LEAF) | lc;
        nsp |= (lc >= (uint32_t)(G::LEAF - 2)) ? (1u << d) : 0u;
      }
      // slab decomposition: a particle whose new base block left this rank's columns is handed to the
      // neighbour below and dropped from the local sort (same rule as k_bin_keys)
      bool leaver = false;
      if (a.slab.enabled) {
        const int nbx = (base_index(x[0], a.K.inv_dx) + a.L.half) >> G::LOG_LEAF;
        leaver = nbx < a.slab.lo || nbx >= a.slab.hi;
      }
      if (a.next_keys && leaver) {
        a.next_keys[s] = INVALID_KEY;
      } else if (a.next_keys) {
        a.next_keys[s] = (nlin_key << G::CB) | ncell;
        if (nbad) {
          atomicOr(&a.st->next_err, ERR_BBOX);
        } else if (nlin_key == cur_lin) {
          seen |= 1u << nsp;         // the common case: flags are written once per block, below
        } else {
          // the particle changed leaf block (rare under the CFL bound): mark directly.  Plain
          // stores, no read-check: a load would put an L2 round trip on the loop's critical path
          a.next_flags[nlin_key] = 1;
          int* gf = a.next_flags + a.next_nlin;
#pragma unroll
          for (uint32_t o = 0; o < (uint32_t)G::NO; ++o)
            if ((o & ~nsp) == 0) gf[(int)nlin_key + oct_delta_l<D>(a.L, (int)o)] = 1;
        }
      }
#pragma unroll
      for (int i = 0; i < D * D; ++i) a.dst[P::w(P::C + i, s)] = __float_as_uint(nC[i]);
      if constexpr (QM == 0) if (a.slab.enabled) {
        // slab decomposition: the particle now belongs to a neighbour rank -> hand it over
        const int nbx = (base_index(x[0], a.K.inv_dx) + a.L.half) >> G::LOG_LEAF;
        const int dir = nbx < a.slab.lo ? 0 : (nbx >= a.slab.hi ? 1 : -1);
        if (dir >= 0 && a.cb.mig[dir]) {
          const int idx = atomicAdd(&a.st->mig_cnt[dir], 1);
          if (idx < a.cb.mig_cap) {
            uint32_t* m = a.cb.mig[dir] + COMM_HEADER + idx;
            const size_t mc = (size_t)a.cb.mig_cap;
#pragma unroll
            for (int d = 0; d < D; ++d) {
              m[(FL::X + d) * mc] = __float_as_uint(x[d]);
              m[(FL::V + d) * mc] = __float_as_uint(nv[d]);
            }
#pragma unroll
            for (int i = 0; i < D * D; ++i) {
              m[(FL::F + i) * mc] = a.dst[word<D>(FL::F + i, s)];   // written by P2G
              m[(FL::C + i) * mc] = __float_as_uint(nC[i]);
            }
            m[FL::JP * mc] = a.dst[word<D>(FL::JP, s)];
            const uint32_t sid = tag_sid(tag);            // the message carries the static attributes along
            m[FL::MAT * mc] = mat;
            m[FL::COLOR * mc] = a.stat.color[sid];
            m[FL::ID * mc] = a.stat.gid[sid];
            m[FL::EMIT * mc] = a.stat.emit[sid];
            if (!a.cb.fused) __threadfence_system();   // peer path: the row may live in the neighbour's memory
          }
        }
      }
      if constexpr (!XS) a.dst[P::w(P::TAG, s)] = tag;
    }
    if (a.next_keys) {
      seen = __reduce_or_sync(0xffffffffu, seen);
      if ((tid & 31) == 0 && seen) atomicOr(&s_seen[u], seen);
    }
    prev_lin = cur_lin;
    have_prev = true;
    b = nb_claim;
    u ^= 1;
  }
  asm volatile("cp.async.wait_all;" ::: "memory");
  if (a.next_keys && have_prev) {
    __syncthreads();
    write_flags(prev_lin, &s_seen[u ^ 1]);
  }
  // CTA-wide reductions, once per CTA lifetime
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
    // NaN velocities are reported as +inf so the host sees the blow-up
    if (vmax != vmax) vmax = __int_as_float(0x7f800000);
    atomicMax(&a.st->maxv_bits, __float_as_uint(vmax));
#pragma unroll
    for (int d = 0; d < D; ++d)
      if (lo[d] <= hi[d]) { atomicMin(&a.st->bb_min[d], lo[d]); atomicMax(&a.st->bb_max[d], hi[d]); }
  }
  if (a.cb.fused) {
    // fused exchange: the leavers of this CTA are in the neighbours' buffers (system-scope fence by every thread);
    // the LAST CTA to finish writes the message headers and publishes the epoch -- no separate header kernel
    __threadfence_system();
    __syncthreads();
    if (tid == 0 && atomicAdd(&a.st->g2p_done, 1) == (int)gridDim.x - 1) {
      a.st->g2p_done = 0;
      __threadfence();
      publish_migration(a.cb, a.cb.epoch + 1u, a.st);
    }
  }
}


// ------------------------------------------------------------------ seeding
__device__ __forceinline__ uint64_t splitmix64(uint64_t z) {
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
// uniform in [0,1) with 24 random bits, counter-based: (seed, particle id, draw)
__device__ __forceinline__ float rand01(uint64_t seed, uint64_t id, uint32_t draw) {
  uint64_t z = splitmix64(seed ^ splitmix64((id << 16) | draw));
  return (float)(z >> 40) * (1.0f / 16777216.0f);
}

struct SeedArgs {
  uint32_t* state;
  size_t cap;
  int64_t n0, n;
  int material, color, emitter;
  float vel[3];
  float a[3], b[3];     // cube: lower,size ; ellipsoid: center,radius
  uint64_t seed;
  const float* x;       // external positions [n][D] (may be null)
  const float* v;       // restart velocities [n][D]
  const int* mats;      // restart per-particle material / color
  const int* colors;
  int mode;             // 0 external, 1 cube, 2 ellipsoid, 3 restart
  const uint32_t* order; // non-null (mode 0): row n0 + i takes input order[i] (inputs pre-sorted by leaf block)
  int64_t id_base;       // id of input 0 (the insertion index n0 on a single-device solver; global ids with slabs)
  float* x_out;          // non-null (modes 1, 2): positions only, [n][D] -- nothing is appended (mpm_seed_generate)
  Statics stat;          // static side arrays; row n0 + i gets the static row sid0 + i
  int64_t sid0;
  int quant;             // packed storage (3D)
};

// Sort key of an external position for block-sorted seeding: absolute leaf-block coordinates, 10 bits per
// axis (the 4096^3 virtual grid of 4^3 leaves), x slowest like the substep's key.  Only the storage order
// of the new rows depends on it (ids keep the insertion order); it spares the first substep after a large
// add_particles the random gather through `perm` that an arbitrary input order causes.
// With a slab (multi-GPU) rows whose base block lies outside this rank's columns get the key 1 << 30, which sorts
// behind every owned row, and the owned rows are counted: selection and block-sorting in one sort.
template <int D>
__global__ void k_seed_keys(const float* __restrict__ x, int64_t n, float inv_dx, int half, uint32_t* __restrict__ keys,
                            uint32_t* __restrict__ vals, Slab slab, unsigned long long* __restrict__ kept) {
  using G = Geo<D>;
  unsigned long long mine_cnt = 0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    uint32_t k = 0;
    bool mine = true;
#pragma unroll
    for (int d = 0; d < D; ++d) {
      const int b = (base_index(x[i * D + d], inv_dx) + half) >> G::LOG_LEAF;
      if (d == 0 && slab.enabled) mine = b >= slab.lo && b < slab.hi;
      k = (k << 10) | (uint32_t)min(max(b, 0), 1023);
    }
    keys[i] = mine ? k : (1u << 30);
    vals[i] = (uint32_t)i;
    mine_cnt += mine ? 1u : 0u;
  }
  if (kept) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mine_cnt += __shfl_xor_sync(0xffffffffu, mine_cnt, o);
    if ((threadIdx.x & 31) == 0 && mine_cnt) atomicAdd(kept, mine_cnt);
  }
}

// seed_particle (engine/mpm_solver.py:823-838) behind seed / seed_ellipsoid /
// seed_from_external_array / recover_from_external_array.
template <int D>
__global__ void k_seed(SeedArgs a) {
  using FL = Fld<D>;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += (int64_t)gridDim.x * blockDim.x) {
    const uint32_t p = (uint32_t)(a.n0 + i);
    const int64_t src = a.order ? (int64_t)a.order[i] : i;     // input this row takes (block-sorted seeding)
    const uint64_t id = (uint64_t)(a.id_base + src);
    float x[D], v[D];
    int material = a.material, color = a.color;
#pragma unroll
    for (int d = 0; d < D; ++d) v[d] = a.vel[d];
    if (a.mode == 0 || a.mode == 3) {
#pragma unroll
      for (int d = 0; d < D; ++d) x[d] = a.x[src * D + d];
      if (a.mode == 3) {
#pragma unroll
        for (int d = 0; d < D; ++d) v[d] = a.v[i * D + d];
        material = a.mats[i]; color = a.colors[i];
      }
    } else if (a.mode == 1) {                                  // :845-848
#pragma unroll
      for (int d = 0; d < D; ++d) x[d] = __fadd_rn(a.a[d], __fmul_rn(rand01(a.seed, id, d), a.b[d]));
    } else {                                                   // :959-976
      float r[D];
      for (uint32_t t = 0; t < 1000; ++t) {
        float nsq = 0.0f;
#pragma unroll
        for (int d = 0; d < D; ++d) {
          r[d] = __fsub_rn(__fmul_rn(rand01(a.seed, id, t * D + d), 2.0f), 1.0f);
          nsq = __fadd_rn(nsq, __fmul_rn(r[d], r[d]));
        }
        if (nsq <= 1.0f) break;
      }
#pragma unroll
      for (int d = 0; d < D; ++d) x[d] = __fadd_rn(a.a[d], __fmul_rn(r[d], a.b[d]));
    }
    if (a.x_out) {                                             // positions only (distributed add_cube / add_ellipsoid)
#pragma unroll
      for (int d = 0; d < D; ++d) a.x_out[i * D + d] = x[d];
      continue;
    }
    const uint32_t sidq = (uint32_t)(a.sid0 + i);
    if constexpr (D == 3) {
      if (a.quant) {                                           // packed storage: x, v, F = I, Jp, tag [, C = 0]
        const float Fi[9] = {1.f, 0.f, 0.f, 0.f, 1.f, 0.f, 0.f, 0.f, 1.f};
        const float jp0 = material == SAND ? 0.0f : 1.0f;
        if (a.quant == 1) {
          using P = PStore<3, 1>;
          P::store_x(a.state, p, x); P::store_v(a.state, p, v); P::store_F(a.state, p, Fi);
          a.state[P::w(P::JP, p)] = __float_as_uint(jp0);
          a.state[P::w(P::TAG, p)] = make_tag((uint32_t)material, sidq);
        } else {
          using P = PStore<3, 2>;
          P::store_x(a.state, p, x); P::store_v(a.state, p, v); P::store_F(a.state, p, Fi);
          a.state[P::w(P::JP, p)] = __float_as_uint(jp0);
          a.state[P::w(P::TAG, p)] = make_tag((uint32_t)material, sidq);
#pragma unroll
          for (int i = 0; i < 9; ++i) a.state[P::w(P::C + i, p)] = 0u;
        }
        a.stat.color[sidq] = (uint32_t)color; a.stat.gid[sidq] = (uint32_t)id; a.stat.emit[sidq] = (uint32_t)a.emitter;
        continue;
      }
    }
#pragma unroll
    for (int d = 0; d < D; ++d) { stf<D>(a.state, FL::X + d, p, x[d]); stf<D>(a.state, FL::V + d, p, v[d]); }
#pragma unroll
    for (int r = 0; r < D; ++r)
#pragma unroll
      for (int c = 0; c < D; ++c) {
        stf<D>(a.state, FL::F + r * D + c, p, r == c ? 1.0f : 0.0f);
        stf<D>(a.state, FL::C + r * D + c, p, 0.0f);
      }
    stf<D>(a.state, FL::JP, p, material == SAND ? 0.0f : 1.0f);   // :831-835
    const uint32_t sid = (uint32_t)(a.sid0 + i);
    stu<D>(a.state, FL::TAG, p, make_tag((uint32_t)material, sid));
    a.stat.color[sid] = (uint32_t)color;
    a.stat.gid[sid] = (uint32_t)id;             // insertion index (block-sorted seeding: != row)
    a.stat.emit[sid] = (uint32_t)a.emitter;
  }
}


// ------------------------------------------------------------------ voxelizer
// Voxelizer.voxelize_triangles (engine/voxelizer.py:46-109), f64 as in the
// reference (precision=ti.f64, :18).  One warp per triangle; lanes walk the
// triangle's xy pixel box and add +-1 to the z-column below the surface.
// `vox` is a dense int32 box [lo, hi) of the super-sampled voxel grid.
struct VoxArgs {
  const double* tris;   // [n][9]
  int64_t ntri;
  int res[3];           // super-sampled resolution
  double dx, inv_dx;    // voxel size
  int padding;
  int lo[3], hi[3];     // allocated box
  int* vox;
};

__device__ __forceinline__ double cross2(double ax, double ay, double bx, double by) { return ax * by - ay * bx; }
__device__ __forceinline__ bool inside_ccw(double px, double py, double ax, double ay, double bx, double by,
                                           double cx, double cy) {
  return cross2(ax - px, ay - py, bx - px, by - py) >= 0 && cross2(bx - px, by - py, cx - px, cy - py) >= 0 &&
         cross2(cx - px, cy - py, ax - px, ay - py) >= 0;
}

__global__ void k_voxelize(VoxArgs a) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarp = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t i = warp; i < a.ntri; i += nwarp) {
    const double js = 1e-8;   // jitter_scale for f64 (:49-53)
    const double jit[3] = {-0.057616723909439505 * js, -0.25608986292614977 * js, 0.06716309129743714 * js};
    double A[3], B[3], C[3];
    for (int k = 0; k < 3; ++k) {
      A[k] = a.tris[i * 9 + k] + jit[k];
      B[k] = a.tris[i * 9 + 3 + k] + jit[k];
      C[k] = a.tris[i * 9 + 6 + k] + jit[k];
    }
    double bmin[2], bmax[2];
    for (int k = 0; k < 2; ++k) {
      bmin[k] = fmin(A[k], fmin(B[k], C[k]));
      bmax[k] = fmax(A[k], fmax(B[k], C[k]));
    }
    int p_min = max(a.padding, (int)floor(bmin[0] * a.inv_dx));
    int p_max = min(a.res[0] - a.padding, (int)floor(bmax[0] * a.inv_dx) + 1);
    int q_min = max(a.padding, (int)floor(bmin[1] * a.inv_dx));
    int q_max = min(a.res[1] - a.padding, (int)floor(bmax[1] * a.inv_dx) + 1);
    double e1[3] = {B[0] - A[0], B[1] - A[1], B[2] - A[2]}, e2[3] = {C[0] - A[0], C[1] - A[1], C[2] - A[2]};
    double nx = e1[1] * e2[2] - e1[2] * e2[1], ny = e1[2] * e2[0] - e1[0] * e2[2], nz = e1[0] * e2[1] - e1[1] * e2[0];
    double inv = 1.0 / sqrt(nx * nx + ny * ny + nz * nz);
    nx *= inv; ny *= inv; nz *= inv;
    if (!(fabs(nz) >= 1e-10)) continue;                         // :83-84 (also skips degenerate NaN normals)
    const int np = p_max - p_min, nq = q_max - q_min;
    if (np <= 0 || nq <= 0) continue;
    const int inc = nz > 0 ? 1 : -1;
    for (int t = lane; t < np * nq; t += 32) {
      int p = p_min + t / nq, q = q_min + t % nq;
      double px = (p + 0.5) * a.dx, py = (q + 0.5) * a.dx;
      if (inside_ccw(px, py, A[0], A[1], B[0], B[1], C[0], C[1]) ||
          inside_ccw(px, py, A[0], A[1], C[0], C[1], B[0], B[1])) {
        double dot = nx * (px - A[0]) + ny * (py - A[1]) + nz * (0.0 - A[2]);
        int height = (int)(-dot / nz * a.inv_dx + 0.5);         // truncation, as Taichi's int()
        height = min(height, a.res[1] - a.padding);             // res[1]: reference quirk (:103)
        height = min(height, a.hi[2]);
        if (p < a.lo[0] || p >= a.hi[0] || q < a.lo[1] || q >= a.hi[1]) continue;
        const size_t ny_ = a.hi[1] - a.lo[1], nz_ = a.hi[2] - a.lo[2];
        int* col = a.vox + ((size_t)(p - a.lo[0]) * ny_ + (q - a.lo[1])) * nz_;
        for (int z = max(a.padding, a.lo[2]); z < height; ++z) atomicAdd(col + (z - a.lo[2]), inc);
      }
    }
  }
}

// seed_from_voxels (engine/mpm_solver.py:1017-1047): per voxel with a positive
// winding count emit floor(s) or ceil(s) particles, s = sample_density/ss^3.
// pass 0: counts[v] = particles of voxel v; pass 1: write positions at offsets[v].
struct VoxSampleArgs {
  const int* vox;
  int lo[3], hi[3];
  int res[3];
  int sample_density, super_sample;
  float s;               // sample_density / super_sample**3
  float cell;            // dx / super_sample
  float trans[3];
  int grid_size, padding;
  uint64_t seed;
  int* counts;           // pass 0 out
  const int64_t* offsets;  // pass 1 in (exclusive prefix of counts)
  float* x_out;          // pass 1 out [n][3]
  int pass;
};

__global__ void k_voxel_sample(VoxSampleArgs a) {
  const size_t ny = a.hi[1] - a.lo[1], nz = a.hi[2] - a.lo[2];
  const size_t total = (size_t)(a.hi[0] - a.lo[0]) * ny * nz;
  for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
    int k = a.lo[2] + (int)(t % nz), j = a.lo[1] + (int)((t / nz) % ny), i = a.lo[0] + (int)(t / (nz * ny));
    int cnt = 0;
    // the reference tests `i` three times (:1027-1028); reproduced
    bool inside = (-a.grid_size / 2 + a.padding <= i) && (i < a.grid_size / 2 - a.padding);
    if (inside && a.vox[t] > 0) {
      const uint64_t vid = ((uint64_t)i * a.res[1] + j) * a.res[2] + k;
      int64_t o = a.pass ? a.offsets[t] : 0;
      for (int l = 0; l < a.sample_density + 1; ++l) {
        if (__fadd_rn(rand01(a.seed, vid, 4 * l), (float)l) < a.s) {
          if (a.pass) {
            const int ijk[3] = {i, j, k};
            for (int d = 0; d < 3; ++d)
              a.x_out[(o + cnt) * 3 + d] =
                  __fadd_rn(__fmul_rn(__fadd_rn(rand01(a.seed, vid, 4 * l + 1 + d), (float)ijk[d]), a.cell), a.trans[d]);
          }
          ++cnt;
        }
      }
    }
    if (!a.pass) a.counts[t] = cnt;
  }
}

// read-back in insertion order: out[id - begin] = field[slot]
// several consecutive state words of particles [begin, end) as rows out[id - begin][nwords]
template <int D>
__global__ void k_gather_rows(const uint32_t* __restrict__ state, Statics stat, int first, int nwords,
                              int n, int64_t begin, int64_t end, uint32_t* __restrict__ out, int quant) {
  for (uint32_t s = blockIdx.x * blockDim.x + threadIdx.x; s < (uint32_t)n; s += gridDim.x * blockDim.x) {
    const int64_t id = vword<D>(state, stat, Fld<D>::ID, s, quant);
    if (id >= begin && id < end)
      for (int w = 0; w < nwords; ++w) out[(size_t)(id - begin) * nwords + w] = vword<D>(state, stat, first + w, s, quant);
  }
}
// rows [0, n) of one (virtual) state word in storage order
template <int D>
__global__ void k_gather_raw(const uint32_t* __restrict__ state, Statics stat, int field, int n, uint32_t* __restrict__ out,
                             int quant) {
  for (uint32_t s = blockIdx.x * blockDim.x + threadIdx.x; s < (uint32_t)n; s += gridDim.x * blockDim.x)
    out[s] = vword<D>(state, stat, field, s, quant);
}
// Static rows renumbered by storage slot (distributed runs: leavers leave holes, arrivals append): pass 0 copies the
// row of every live particle to tmp[3][n], pass 1 copies back and rewrites the tags (sid = slot).
template <int D>
__global__ void k_compact_statics(uint32_t* __restrict__ state, Statics stat, int n, uint32_t* __restrict__ tmp, int pass,
                                  int quant) {
  for (uint32_t s = blockIdx.x * blockDim.x + threadIdx.x; s < (uint32_t)n; s += gridDim.x * blockDim.x) {
    if (pass == 0) {
      const uint32_t sid = tag_sid(load_tag_rt<D>(state, quant, s));
      tmp[s] = stat.color[sid]; tmp[(size_t)n + s] = stat.gid[sid]; tmp[2 * (size_t)n + s] = stat.emit[sid];
    } else {
      stat.color[s] = tmp[s]; stat.gid[s] = tmp[(size_t)n + s]; stat.emit[s] = tmp[2 * (size_t)n + s];
      const size_t w = tag_word_rt<D>(quant, s);
      state[w] = make_tag(tag_mat(state[w]), s);
    }
  }
}

// particle_info() of one slab rank: rows [x[D] v[D] material color id] of the particles whose base block lies in
// this rank's columns (rows already handed to a neighbour are skipped), compacted in storage order groups
template <int D>
__global__ void k_export_local(const uint32_t* __restrict__ state, Statics stat, int n, float inv_dx, int half, Slab slab,
                               uint32_t* __restrict__ out, unsigned long long* __restrict__ count, int quant) {
  using G = Geo<D>;
  using FL = Fld<D>;
  const int lane = threadIdx.x & 31;
  const uint32_t nround = ((uint32_t)n + 31u) & ~31u;
  for (uint32_t s = blockIdx.x * blockDim.x + threadIdx.x; s < nround; s += gridDim.x * blockDim.x) {
    bool mine = s < (uint32_t)n;
    if (mine && slab.enabled) {
      const int bx = (base_index(__uint_as_float(vword<D>(state, stat, FL::X, s, quant)), inv_dx) + half) >> G::LOG_LEAF;
      mine = bx >= slab.lo && bx < slab.hi && tag_mat(load_tag_rt<D>(state, quant, s)) != MAT_DEAD;
    }
    const unsigned m = __ballot_sync(0xffffffffu, mine);
    unsigned long long base = 0;
    if (lane == 0 && m) base = atomicAdd(count, (unsigned long long)__popc(m));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (!mine) continue;
    // field blocks sized for n rows each: [x n*D | v n*D | material n | colour n | id n]; the first *count rows are valid
    const size_t r = (size_t)(base + __popc(m & ((1u << lane) - 1u))), nn = (size_t)n;
    float ex[D], ev[D];
    load_x_rt<D>(state, quant, s, ex);
    load_v_rt<D>(state, quant, s, ev);
#pragma unroll
    for (int d = 0; d < D; ++d) { out[r * D + d] = __float_as_uint(ex[d]); out[nn * D + r * D + d] = __float_as_uint(ev[d]); }
    const uint32_t tag = load_tag_rt<D>(state, quant, s);
    out[2 * nn * D + r] = tag_mat(tag);
    out[2 * nn * D + nn + r] = stat.color[tag_sid(tag)];
    out[2 * nn * D + 2 * nn + r] = stat.gid[tag_sid(tag)];
  }
}

// ---- ParticleIO.write_particles on the device (engine/particle_io.py:42-76)
// order-preserving map float -> uint32 so integer atomics give float min / max
__device__ __forceinline__ uint32_t f2ord(float f) {
  const uint32_t b = __float_as_uint(f);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float ord2f(uint32_t u) {
  return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}
__global__ void k_ranges_init(uint32_t* r, int nwords) {
  const int i = threadIdx.x;
  if (i < nwords) r[i] = (i & 1) ? 0u : 0xffffffffu;     // [.. min, max ..]
}
// ranges[c][d][0|1] (c = 0: x, 1: v) as ordered uints; fields X and V are the first 2*D state words
template <int D>
__global__ void k_ranges(const uint32_t* __restrict__ state, size_t cap, int n, uint32_t* __restrict__ r, int quant) {
  float lo[2 * D], hi[2 * D];
#pragma unroll
  for (int f = 0; f < 2 * D; ++f) { lo[f] = __int_as_float(0x7f800000); hi[f] = __int_as_float(0xff800000); }
  for (uint32_t s = blockIdx.x * blockDim.x + threadIdx.x; s < (uint32_t)n; s += gridDim.x * blockDim.x) {
    float xv[2 * D];
    load_x_rt<D>(state, quant, s, xv);
    load_v_rt<D>(state, quant, s, xv + D);
#pragma unroll
    for (int f = 0; f < 2 * D; ++f) {
      const float a = xv[f];
      lo[f] = fminf(lo[f], a); hi[f] = fmaxf(hi[f], a);
    }
  }
#pragma unroll
  for (int f = 0; f < 2 * D; ++f) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      lo[f] = fminf(lo[f], __shfl_xor_sync(0xffffffffu, lo[f], o));
      hi[f] = fmaxf(hi[f], __shfl_xor_sync(0xffffffffu, hi[f], o));
    }
    if ((threadIdx.x & 31) == 0 && lo[f] <= hi[f]) {
      atomicMin(&r[2 * f], f2ord(lo[f]));
      atomicMax(&r[2 * f + 1], f2ord(hi[f]));
    }
  }
}
__global__ void k_ranges_decode(uint32_t* r, int nwords) {
  const int i = threadIdx.x;
  if (i < nwords) r[i] = __float_as_uint(ord2f(r[i]));
}
struct PackArgs { float lo[2][3], inv[2][3]; };
template <int D>
__global__ void k_pack_particles(const uint32_t* __restrict__ state, Statics stat, int n, PackArgs pa,
                                 uint32_t* __restrict__ xv, uint8_t* __restrict__ color, int quant) {
  using FL = Fld<D>;
  for (uint32_t s = blockIdx.x * blockDim.x + threadIdx.x; s < (uint32_t)n; s += gridDim.x * blockDim.x) {
    const uint32_t sid = tag_sid(load_tag_rt<D>(state, quant, s)), id = stat.gid[sid];
    if (id >= (uint32_t)n) continue;
    float px[D], pv[D];
    load_x_rt<D>(state, quant, s, px);
    load_v_rt<D>(state, quant, s, pv);   // ids are a permutation of [0, n) on a single-device solver; never write outside
#pragma unroll
    for (int d = 0; d < D; ++d) {
      // ((a - lo) * (1 / (hi - lo)) * (2^bits - 1) + 0.499).astype(uint32), every step rounded to f32 (:50-55)
      const float x = px[d], v = pv[d];
      const uint32_t xq = __float2uint_rz(__fadd_rn(__fmul_rn(__fmul_rn(__fsub_rn(x, pa.lo[0][d]), pa.inv[0][d]), 16777215.0f), 0.499f));
      const uint32_t vq = __float2uint_rz(__fadd_rn(__fmul_rn(__fmul_rn(__fsub_rn(v, pa.lo[1][d]), pa.inv[1][d]), 255.0f), 0.499f));
      xv[(size_t)id * D + d] = (xq << 8) + vq;
    }
    const uint32_t c = stat.color[sid];
    color[(size_t)id * 3 + 0] = (uint8_t)((c >> 16) & 255u);
    color[(size_t)id * 3 + 1] = (uint8_t)((c >> 8) & 255u);
    color[(size_t)id * 3 + 2] = (uint8_t)(c & 255u);
  }
}

template <int D>
__global__ void k_debug_binning(const uint32_t* __restrict__ state, Statics stat, int n, float inv_dx, int half,
                                int* __restrict__ out, int quant) {
  using G = Geo<D>;
  for (uint32_t p = blockIdx.x * blockDim.x + threadIdx.x; p < (uint32_t)n; p += gridDim.x * blockDim.x) {
    uint32_t id = stat.gid[tag_sid(load_tag_rt<D>(state, quant, p))];
    if (id >= (uint32_t)n) continue;
    float bx[D];
    load_x_rt<D>(state, quant, p, bx);
#pragma unroll
    for (int d = 0; d < D; ++d)
      out[(size_t)id * D + d] = (base_index(bx[d], inv_dx) + half) >> G::LOG_LEAF;
  }
}

template <int D>
__global__ void k_debug_update(Consts K, float dt, int n, const int* mat, float* F, const float* C, float* Jp,
                               float* aff, float* mass) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    float f[D * D], c[D * D], a[D * D], jp = Jp[i], m;
    for (int k = 0; k < D * D; ++k) { f[k] = F[i * D * D + k]; c[k] = C[i * D * D + k]; }
    particle_update<D>(K, dt, mat[i], f, c, jp, a, m);
    for (int k = 0; k < D * D; ++k) { F[i * D * D + k] = f[k]; aff[i * D * D + k] = a[k]; }
    Jp[i] = jp; mass[i] = m;
  }
}

}  // namespace mpm
