// libmpm_b200.so -- host side of the C ABI declared in include/mpm_b200.h.
// Orchestrates the per-substep kernel sequence on one CUDA stream; owns no
// device memory (state and workspace are bound by the caller).
#include <cub/cub.cuh>
#include <thrust/iterator/counting_iterator.h>

#include <algorithm>
#include <atomic>
#include <climits>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/mpm_b200.h"
#include "mpm_bin.cuh"
#include "mpm_comm.cuh"
#include "mpm_p2g.cuh"
#include "mpm_p2g3.cuh"
#include "mpm_g2p2g.cuh"

using namespace mpm;

struct mpm_ctx {
  mpm_params P;
  Consts K;
  int dim = 0, nf = 0, cells = 0, no = 0, cb = 0, log_leaf = 0;
  uint32_t* state[2] = {nullptr, nullptr};
  Statics stat{nullptr, nullptr, nullptr};   // static side arrays [3][cap]: colour, id, emitter by sid
  int64_t n_static = 0;                       // rows of them in use
  int nv = 0;                                 // virtual words of the read-back numbering (Fld::NV)
  int quant = 0;                              // packed x / v / F storage (quant=True, 3D; mpm_quant.cuh): 1 with use_g2p2g, 2 split
  size_t cap = 0;
  void* ws = nullptr;
  size_t ws_bytes = 0;
  int max_blocks = 0;
  // carved workspace
  Status* d_status = nullptr;
  ColliderTable* d_ct = nullptr;
  uint32_t *keys_a = nullptr, *keys_b = nullptr, *vals_a = nullptr, *vals_b = nullptr, *stage = nullptr;
  void* cub_temp = nullptr;
  unsigned long long* scan_desc = nullptr;   // tile descriptors of k_scan_excl
  size_t scan_tiles = 0;
  uint32_t scan_epoch = 0;
  int own_scan = 1, scan_grid = 0, scan2_force = 0;   // MPM_SCAN=two: always the two-level bucket scan (tests)
  size_t cub_bytes = 0;
  int* pb_start = nullptr;
  uint32_t* pb_mask = nullptr;
  int* pb_nbr = nullptr;
  uint32_t *cand_a = nullptr, *cand_b = nullptr, *gb_key = nullptr, *pb_key = nullptr;
  int *flags = nullptr, *fscan = nullptr, *cellcount = nullptr, *cellstart = nullptr;
  int *blocksum = nullptr, *blockstart = nullptr;   // two-level bucket scan (large block counts)
  int64_t table_cap = 0;
  bool dense = false;   // counting-sort path usable for the current layout
  Slab slab{0, INT_MIN, INT_MAX};
  CommBufs comm{};
  // peer path (NVLink writes into the neighbour's buffers)
  uint32_t* peer_region = nullptr;     // cudaMalloc'ed by mpm_peer_alloc: flags + the four receive buffers
  size_t peer_mig_words = 0, peer_halo_words = 0, peer_plane_words = 0;   // plane: one of the 3 rotating halo planes of a side
  int fused_fast = 1;                  // use_g2p2g, 3D: cell-owner fused kernel (MPM_G2P2G=simple: general kernel)
  int fused_halo = 1;                  // MPM_FUSED_HALO=0: legacy pack / wait / add kernels on the peer path
  void* peer_open[2] = {nullptr, nullptr};
  uint32_t epoch = 0;                  // substeps completed since mpm_peer_alloc (same on every rank)
  bool ext_box = false;          // layout box supplied by the host (global box of all ranks)
  int box_min[3] = {0, 0, 0}, box_max[3] = {0, 0, 0};
  // state of the batch being enqueued (phase API)
  // Block structure and grid exist twice (use_g2p2g: the fused kernel reads last substep's OUTPUT grid through its
  // block table while it scatters into the other one; the split path only uses set 0).  `sel` = the set that the
  // pointers flags / fscan / grid / pb_start / pb_key / pb_nbr currently refer to.
  int* flags2[2] = {nullptr, nullptr};
  int* fscan2[2] = {nullptr, nullptr};
  float4* grid2[2] = {nullptr, nullptr};
  int* pb_start2[2] = {nullptr, nullptr};
  uint32_t* pb_key2[2] = {nullptr, nullptr};
  int* pb_nbr2[2] = {nullptr, nullptr};
  int sel = 0;
  // use_g2p2g: the input grid of the next fused substep = set `sel` as left by the last completed substep
  bool fused_prev = false;       // such a grid exists (false before the first substep and after clear_particles)
  int64_t fused_n_old = 0;       // rows that existed then: rows beyond them skip the gather (:396-399)
  KeyLayout fused_L{};           // key layout of its block table
  int fused_nlin = 0, fused_npb = 0, fused_ngb = 0;
  bool keys_ready = false;     // the previous G2P of this batch already produced keys + flags
  int fuse_keys = 1;
  int sort_seed = 1;            // (MPM_SORT_SEED) block-sorted storage of large add_particles arrays
  bool in_batch = false;
  int batch_cur0 = 0, batch_enq = 0;
  const uint32_t* cur_keys = nullptr;
  const uint32_t* cur_perm = nullptr;
  const int* cur_cellstart = nullptr;
  int use_dense = 1;
  float4* grid = nullptr;
  Status* h_status = nullptr;   // pinned
  int cur = 0;
  int64_t n = 0;
  bool bbox_valid = false;
  int bb_min[3] = {0, 0, 0}, bb_max[3] = {0, 0, 0};
  KeyLayout L{};
  bool layout_valid = false;
  // structure of the last completed substep (debug getters)
  bool last_valid = false;
  const uint32_t* last_keys = nullptr;
  KeyLayout lastL{};
  Status last{};
  ColliderTable h_ct{};
  bool ct_dirty = true;
  Grav grav{};
  GridCfg gcfg{};
  int sm_count = 148;
  int grid_p2g = 148, grid_g2p = 148, grid_p2g_cell = 148;
  int g2p_cfg = 6;
  int p2g_cfg = 0;
  int pdl = 1;                  // programmatic dependent launch along the substep chain (MPM_PDL)
  bool cell_zeroed = false, flags_zeroed = false;   // tables already cleared by k_clear_grid
  int g2p_tile = 0;             // G2P tile staging: 0 cp.async per node, 1 cp.async.bulk per row (MPM_G2P_TILE)
  int pf_mode = 2;              // next-block L2 prefetch: cp.async.bulk.prefetch ranges (MPM_PREFETCH)
  int p2g_ver = 3;              // 3: mpm_p2g3.cuh (3D, dense binning); 2: mpm_p2g.cuh
  int p2g_variant = 1;   // 0 = shared-atomic scatter (first version), 1 = cell-owner
  int launches = 0;
  int done_last = 0;
  bool profiling = false;
  cudaEvent_t* cur_ev = nullptr;   // phase API: the five events of the substep being enqueued (profiling), or null
  int prof_enq = 0;                // substeps of the open batch that carry events
  std::vector<cudaEvent_t> ev;   // 5 per profiled substep
  float ms[4] = {0, 0, 0, 0};
  std::string err;
};

#define CK(call)                                                                       \
  do {                                                                                 \
    cudaError_t e_ = (call);                                                           \
    if (e_ != cudaSuccess) {                                                           \
      ctx->err = std::string(#call) + ": " + cudaGetErrorString(e_);                   \
      return MPM_E_CUDA;                                                               \
    }                                                                                  \
  } while (0)

static int fail(mpm_ctx* ctx, int code, const std::string& msg) {
  if (ctx) ctx->err = msg;
  return code;
}

static inline int gs_blocks(int64_t n, int threads, int sm) {
  int64_t b = (n + threads - 1) / threads;
  return (int)std::max<int64_t>(1, std::min<int64_t>(b, (int64_t)sm * 16));
}

// Launch with the programmatic-stream-serialization attribute (PDL): the kernel may be scheduled
// while the previous kernel of the stream drains; it parks in pdl_enter() until that one is complete.
template <typename... KArgs, typename... Args>
static cudaError_t launch_chain(bool pdl, void (*kernel)(KArgs...), int grid, int block, size_t smem, cudaStream_t s,
                                Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3((unsigned)block);
  cfg.dynamicSmemBytes = smem; cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(std::forward<Args>(args))...);
}

// ------------------------------------------------------------------ sizes
struct Carve {
  size_t off_status, off_ct, off_scratch, off_cub, off_pb_start, off_pb_mask, off_pb_nbr, off_cand_a,
      off_cand_b, off_gb_key, off_grid, off_pb_key, off_flags, off_fscan, off_cellcount, off_cellstart, off_scan_desc, total,
      off_pb_start2, off_pb_nbr2, off_grid2, off_pb_key2, off_flags2, off_fscan2, off_blocksum, off_blockstart,
      cub_bytes, scan_tiles;
  int64_t table_cap;
};

static int64_t table_capacity(int32_t max_blocks) {
  return std::min<int64_t>(std::max<int64_t>((int64_t)128 * max_blocks, (int64_t)1 << 18), (int64_t)1 << 24);
}

static size_t cub_temp_bytes(int64_t cap, int64_t ncand, int64_t nscan) {
  size_t best = 0, b = 0;
  cub::DoubleBuffer<uint32_t> k(nullptr, nullptr), v(nullptr, nullptr);
  cub::DeviceRadixSort::SortPairs(nullptr, b, k, v, (int)std::max<int64_t>(cap, 1));
  best = std::max(best, b);
  cub::DeviceRadixSort::SortKeys(nullptr, b, k, (int)std::max<int64_t>(ncand, 1));
  best = std::max(best, b);
  thrust::counting_iterator<int> it(0);
  cub::DeviceSelect::If(nullptr, b, it, BoundedOut{nullptr, 0}, (int*)nullptr, (int)std::max<int64_t>(cap, 1),
                        HeadOp{nullptr, 0});
  best = std::max(best, b);
  cub::DeviceSelect::Unique(nullptr, b, (uint32_t*)nullptr, (uint32_t*)nullptr, (int*)nullptr,
                            (int)std::max<int64_t>(ncand, 1));
  best = std::max(best, b);
  cub::DeviceScan::ExclusiveSum(nullptr, b, (int*)nullptr, (int*)nullptr, (int)std::max<int64_t>(nscan, 1));
  best = std::max(best, b);
  return best + 256;
}

static Carve carve(int dim, int64_t cap, int32_t max_blocks) {
  const int no = dim == 3 ? 8 : 4, cells = dim == 3 ? 64 : 256;
  Carve c{};
  size_t o = 0;
  auto take = [&](size_t bytes) { size_t r = o; o = align_up(o + bytes, 256); return r; };
  c.off_status = take(sizeof(Status));
  c.off_ct = take(sizeof(ColliderTable));
  c.off_scratch = take((size_t)5 * cap * 4);
  c.table_cap = table_capacity(max_blocks);
  c.cub_bytes = cub_temp_bytes(cap, (int64_t)no * max_blocks, std::max<int64_t>(c.table_cap, (int64_t)max_blocks * cells + 1));
  c.off_cub = take(c.cub_bytes);
  c.scan_tiles = (size_t)(std::max<int64_t>(c.table_cap, (int64_t)max_blocks * cells + 1) / SCAN_TILE + 2);
  c.off_scan_desc = take(c.scan_tiles * 8);
  c.off_pb_start = take((size_t)(max_blocks + 2) * 4);
  c.off_pb_mask = take((size_t)max_blocks * 4);
  c.off_pb_nbr = take((size_t)max_blocks * no * 4);
  c.off_cand_a = take((size_t)max_blocks * no * 4);
  c.off_cand_b = take((size_t)max_blocks * no * 4);
  c.off_gb_key = take(((size_t)max_blocks * no + 1) * 4);
  c.off_grid = take((size_t)max_blocks * cells * sizeof(float4));
  c.off_pb_key = take((size_t)max_blocks * 4);
  c.off_flags = take((size_t)c.table_cap * 4);
  c.off_fscan = take((size_t)c.table_cap * 4);
  // second set of block structure + grid (use_g2p2g ping-pong)
  c.off_pb_start2 = take((size_t)(max_blocks + 2) * 4);
  c.off_pb_nbr2 = take((size_t)max_blocks * no * 4);
  c.off_grid2 = take((size_t)max_blocks * cells * sizeof(float4));
  c.off_pb_key2 = take((size_t)max_blocks * 4);
  c.off_flags2 = take((size_t)c.table_cap * 4);
  c.off_fscan2 = take((size_t)c.table_cap * 4);
  c.off_blocksum = take((size_t)(max_blocks + 2) * 4);
  c.off_blockstart = take((size_t)(max_blocks + 2) * 4);
  c.off_cellcount = take(((size_t)max_blocks * cells + 1) * 4);
  c.off_cellstart = take(((size_t)max_blocks * cells + 1) * 4);
  c.total = o;
  return c;
}

extern "C" int mpm_abi_version(void) { return MPM_ABI_VERSION; }
extern "C" int mpm_state_fields(int dim) { return dim == 3 ? Geo<3>::NF : (dim == 2 ? Geo<2>::NF : -1); }
extern "C" int mpm_ctx_state_fields(mpm_ctx* ctx) { return ctx ? ctx->nf : -1; }
extern "C" int mpm_virtual_fields(int dim) { return dim == 3 ? Fld<3>::NV : (dim == 2 ? Fld<2>::NV : -1); }
extern "C" size_t mpm_workspace_bytes(int dim, int64_t capacity, int32_t max_blocks) {
  if ((dim != 2 && dim != 3) || capacity < 0 || max_blocks < 1) return 0;
  return carve(dim, capacity, max_blocks).total;
}

// ------------------------------------------------------------------ lifecycle
extern "C" int mpm_create(const mpm_params* p, mpm_ctx** out) {
  if (!p || !out) return MPM_E_INVALID;
  if (p->dim != 2 && p->dim != 3) return MPM_E_INVALID;          // engine/mpm_solver.py:76-77
  if (p->leaf != (p->dim == 3 ? 4 : 16)) return MPM_E_INVALID;
  mpm_ctx* ctx = new mpm_ctx();
  ctx->P = *p;
  ctx->dim = p->dim;
  // quant=True (flags bit 1) in 3D: bit-packed x / v / F; 1 = with use_g2p2g (no C), 2 = split substep (f32 C kept)
  ctx->quant = ((p->flags & 2) && p->dim == 3) ? ((p->flags & 1) ? 1 : 2) : 0;
  ctx->nf = ctx->quant ? quant_words(ctx->quant) : mpm_state_fields(p->dim);
  ctx->nv = p->dim == 3 ? Fld<3>::NV : Fld<2>::NV;
  ctx->cells = p->dim == 3 ? 64 : 256;
  ctx->no = p->dim == 3 ? 8 : 4;
  ctx->cb = p->dim == 3 ? 6 : 8;
  ctx->log_leaf = p->dim == 3 ? 2 : 4;
  Consts& K = ctx->K;
  K.dx = (float)p->dx; K.inv_dx = (float)p->inv_dx;
  K.p_vol = (float)p->p_vol; K.p_mass = (float)p->p_mass;
  K.mu_0 = (float)p->mu_0; K.lambda_0 = (float)p->lambda_0;
  K.alpha = (float)p->alpha;
  K.sand_coef = (float)((p->dim * p->lambda_0 + 2 * p->mu_0) / (2 * p->mu_0));
  K.water_density = (float)p->water_density;
  K.inv_dx2 = (float)(p->inv_dx * p->inv_dx);
  K.four_inv_dx = (float)(4 * p->inv_dx);
  K.support_plasticity = p->support_plasticity;
  K.g2p2g = (p->flags & 1) ? 1 : 0;
  K.clamp_F = (p->flags & 2) ? 1 : 0;
  K.v_allowed_cfl = (float)(p->dx * p->g2p2g_cfl);
  for (int d = 0; d < 3; ++d) ctx->gcfg.res[d] = p->res[d];
  ctx->gcfg.padding = p->padding;
  ctx->gcfg.grid_size = p->grid_size;
  ctx->grav.g[0] = 0.f; ctx->grav.g[1] = -9.8f; ctx->grav.g[2] = 0.f;   // :283, 296
  ctx->h_ct.n = 0;
  cudaError_t e = cudaSetDevice(p->device);
  if (e == cudaSuccess) e = cudaMallocHost((void**)&ctx->h_status, sizeof(Status));
  cudaDeviceProp prop;
  if (e == cudaSuccess) e = cudaGetDeviceProperties(&prop, p->device);
  if (e != cudaSuccess) {
    fprintf(stderr, "mpm_create: %s\n", cudaGetErrorString(e));
    delete ctx;
    return MPM_E_CUDA;
  }
  ctx->sm_count = prop.multiProcessorCount;
  int occ = 1;
  if (p->dim == 3) {
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_p2g<3>, P2G_THREADS, 0);
    ctx->grid_p2g = ctx->sm_count * std::max(occ, 1);

  } else {
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_p2g<2>, P2G_THREADS, 0);
    ctx->grid_p2g = ctx->sm_count * std::max(occ, 1);

  }
  if (const char* v = getenv("MPM_FUSE_KEYS")) ctx->fuse_keys = atoi(v);
  if (const char* v = getenv("MPM_SORT_SEED")) ctx->sort_seed = atoi(v);
  if (const char* v = getenv("MPM_SORT")) ctx->use_dense = (strcmp(v, "radix") == 0) ? 0 : 1;
  if (const char* v = getenv("MPM_P2G_CFG")) ctx->p2g_cfg = atoi(v);
  if (const char* v = getenv("MPM_P2G_VER")) ctx->p2g_ver = atoi(v);
  if (const char* v = getenv("MPM_PREFETCH")) ctx->pf_mode = atoi(v);
  if (const char* v = getenv("MPM_G2P_TILE")) ctx->g2p_tile = atoi(v);
  if (const char* v = getenv("MPM_PDL")) ctx->pdl = atoi(v);
  if (const char* v = getenv("MPM_FUSED_HALO")) ctx->fused_halo = atoi(v);
  if (const char* v = getenv("MPM_G2P2G")) ctx->fused_fast = strcmp(v, "simple") != 0;
  if (const char* v = getenv("MPM_SCAN")) { ctx->own_scan = strcmp(v, "cub") != 0; ctx->scan2_force = strcmp(v, "two") == 0; }
  {
    int occ = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_scan_excl<true>, SCAN_T, 0);
    ctx->scan_grid = ctx->sm_count * std::max(1, std::min(occ, 4));
  }
  if (const char* v = getenv("MPM_G2P_CFG")) ctx->g2p_cfg = atoi(v);
  if (const char* v = getenv("MPM_P2G")) ctx->p2g_variant = (strcmp(v, "atomic") == 0) ? 0 : 1;
  *out = ctx;
  return MPM_OK;
}

static void peer_close(mpm_ctx* ctx);
extern "C" int mpm_destroy(mpm_ctx* ctx) {
  if (!ctx) return MPM_E_INVALID;
  cudaSetDevice(ctx->P.device);
  peer_close(ctx);
  if (ctx->h_status) cudaFreeHost(ctx->h_status);
  for (cudaEvent_t e : ctx->ev) cudaEventDestroy(e);
  delete ctx;
  return MPM_OK;
}

extern "C" const char* mpm_last_error(mpm_ctx* ctx) { return ctx ? ctx->err.c_str() : "null ctx"; }

static void select_set(mpm_ctx* ctx, int q) {
  ctx->sel = q;
  ctx->flags = ctx->flags2[q]; ctx->fscan = ctx->fscan2[q]; ctx->grid = ctx->grid2[q];
  ctx->pb_start = ctx->pb_start2[q]; ctx->pb_key = ctx->pb_key2[q]; ctx->pb_nbr = ctx->pb_nbr2[q];
}

extern "C" int mpm_bind(mpm_ctx* ctx, void* s0, void* s1, void* statics, int64_t capacity, void* ws, size_t ws_bytes,
                        int32_t max_blocks) {
  if (!ctx || !s0 || !s1 || !statics || !ws || capacity < 1 || max_blocks < 1) return fail(ctx, MPM_E_INVALID, "mpm_bind: bad argument");
  if (capacity > (int64_t)TAG_SID + 1) return fail(ctx, MPM_E_INVALID, "mpm_bind: capacity must be <= 2^29 (static row index of the tag word)");
  if (capacity % 64 != 0) return fail(ctx, MPM_E_INVALID, "mpm_bind: capacity must be a multiple of 64");
  if (capacity >= (int64_t)1 << 31) return fail(ctx, MPM_E_INVALID, "mpm_bind: capacity must be < 2^31");
  Carve c = carve(ctx->dim, capacity, max_blocks);
  if (ws_bytes < c.total) return fail(ctx, MPM_E_INVALID, "mpm_bind: workspace too small");
  CK(cudaSetDevice(ctx->P.device));
  // use_g2p2g: the input grid of the next substep and its block structure live in the OLD workspace (which the
  // caller keeps alive until this call returns): remember where, they are copied over below
  const bool migrate = ctx->K.g2p2g && ctx->fused_prev && ctx->ws && ctx->ws != ws;
  int* o_flags = ctx->flags2[ctx->sel]; int* o_fscan = ctx->fscan2[ctx->sel]; float4* o_grid = ctx->grid2[ctx->sel];
  int* o_pbs = ctx->pb_start2[ctx->sel]; uint32_t* o_pbk = ctx->pb_key2[ctx->sel]; int* o_pbn = ctx->pb_nbr2[ctx->sel];
  if (migrate && (ctx->fused_npb > max_blocks || ctx->fused_ngb > max_blocks || 2 * (int64_t)ctx->fused_nlin + 1 > c.table_cap))
    return fail(ctx, MPM_E_INVALID, "mpm_bind: the new workspace cannot hold the pending g2p2g grid");
  ctx->state[0] = (uint32_t*)s0;
  ctx->state[1] = (uint32_t*)s1;
  ctx->stat.color = (uint32_t*)statics;
  ctx->stat.gid = (uint32_t*)statics + capacity;
  ctx->stat.emit = (uint32_t*)statics + 2 * capacity;
  ctx->cap = (size_t)capacity;
  ctx->cell_zeroed = false; ctx->flags_zeroed = false;   // new workspace: nothing is cleared yet
  ctx->ws = ws;
  ctx->ws_bytes = ws_bytes;
  ctx->max_blocks = max_blocks;
  char* b = (char*)ws;
  ctx->d_status = (Status*)(b + c.off_status);
  ctx->d_ct = (ColliderTable*)(b + c.off_ct);
  uint32_t* sc = (uint32_t*)(b + c.off_scratch);
  ctx->keys_a = sc; ctx->keys_b = sc + capacity; ctx->vals_a = sc + 2 * capacity;
  ctx->vals_b = sc + 3 * capacity; ctx->stage = sc + 4 * capacity;
  ctx->cub_temp = b + c.off_cub;
  ctx->scan_desc = (unsigned long long*)(b + c.off_scan_desc);
  ctx->scan_tiles = c.scan_tiles;
  cudaMemsetAsync(ctx->scan_desc, 0, c.scan_tiles * 8, 0);   // epoch 0 = never valid
  cudaStreamSynchronize(0);
  ctx->cub_bytes = c.cub_bytes;
  ctx->pb_start = (int*)(b + c.off_pb_start);
  ctx->pb_mask = (uint32_t*)(b + c.off_pb_mask);
  ctx->pb_nbr = (int*)(b + c.off_pb_nbr);
  ctx->cand_a = (uint32_t*)(b + c.off_cand_a);
  ctx->cand_b = (uint32_t*)(b + c.off_cand_b);
  ctx->gb_key = (uint32_t*)(b + c.off_gb_key);
  ctx->grid = (float4*)(b + c.off_grid);
  ctx->pb_key = (uint32_t*)(b + c.off_pb_key);
  ctx->flags = (int*)(b + c.off_flags);
  ctx->fscan = (int*)(b + c.off_fscan);
  ctx->blocksum = (int*)(b + c.off_blocksum);
  ctx->blockstart = (int*)(b + c.off_blockstart);
  ctx->cellcount = (int*)(b + c.off_cellcount);
  ctx->cellstart = (int*)(b + c.off_cellstart);
  ctx->table_cap = c.table_cap;
  ctx->ct_dirty = true;
  ctx->last_valid = false;
  ctx->pb_start2[0] = ctx->pb_start; ctx->pb_nbr2[0] = ctx->pb_nbr; ctx->grid2[0] = ctx->grid;
  ctx->pb_key2[0] = ctx->pb_key; ctx->flags2[0] = ctx->flags; ctx->fscan2[0] = ctx->fscan;
  ctx->pb_start2[1] = (int*)(b + c.off_pb_start2); ctx->pb_nbr2[1] = (int*)(b + c.off_pb_nbr2);
  ctx->grid2[1] = (float4*)(b + c.off_grid2); ctx->pb_key2[1] = (uint32_t*)(b + c.off_pb_key2);
  ctx->flags2[1] = (int*)(b + c.off_flags2); ctx->fscan2[1] = (int*)(b + c.off_fscan2);
  if (migrate) {
    const int q = ctx->sel, no = ctx->no;
    const size_t nt = 2 * (size_t)ctx->fused_nlin + 1;
    CK(cudaMemcpy(ctx->flags2[q], o_flags, nt * 4, cudaMemcpyDeviceToDevice));
    CK(cudaMemcpy(ctx->fscan2[q], o_fscan, nt * 4, cudaMemcpyDeviceToDevice));
    CK(cudaMemcpy(ctx->grid2[q], o_grid, (size_t)ctx->fused_ngb * ctx->cells * sizeof(float4), cudaMemcpyDeviceToDevice));
    CK(cudaMemcpy(ctx->pb_start2[q], o_pbs, ((size_t)ctx->fused_npb + 1) * 4, cudaMemcpyDeviceToDevice));
    CK(cudaMemcpy(ctx->pb_key2[q], o_pbk, (size_t)ctx->fused_npb * 4, cudaMemcpyDeviceToDevice));
    CK(cudaMemcpy(ctx->pb_nbr2[q], o_pbn, (size_t)ctx->fused_npb * no * 4, cudaMemcpyDeviceToDevice));
  } else if (ctx->ws != nullptr && ctx->fused_prev && ctx->K.g2p2g && ctx->ws == ws) {
    // same workspace re-bound: nothing moved
  } else {
    ctx->fused_prev = false;
  }
  select_set(ctx, ctx->K.g2p2g ? ctx->sel : 0);
  return MPM_OK;
}

extern "C" int mpm_get_state(mpm_ctx* ctx, int32_t* cur, int64_t* n) {
  if (!ctx) return MPM_E_INVALID;
  if (cur) *cur = ctx->cur;
  if (n) *n = ctx->n;
  return MPM_OK;
}
extern "C" int mpm_set_state(mpm_ctx* ctx, int32_t cur, int64_t n) {
  if (!ctx || (cur != 0 && cur != 1) || n < 0 || (size_t)n > ctx->cap) return fail(ctx, MPM_E_INVALID, "mpm_set_state: bad argument");
  if (n < ctx->n || n < ctx->fused_n_old) ctx->fused_prev = false;   // cleared or replaced: the next fused substep starts over
  ctx->cur = cur;
  ctx->n = n;
  if (n == 0) ctx->n_static = 0;
  ctx->bbox_valid = false;
  ctx->last_valid = false;
  return MPM_OK;
}
extern "C" int mpm_get_static_rows(mpm_ctx* ctx, int64_t* n_static) {
  if (!ctx || !n_static) return MPM_E_INVALID;
  *n_static = ctx->n_static;
  return MPM_OK;
}
extern "C" int mpm_set_static_rows(mpm_ctx* ctx, int64_t n_static) {
  if (!ctx || n_static < 0 || (size_t)n_static > ctx->cap) return MPM_E_INVALID;
  ctx->n_static = n_static;
  return MPM_OK;
}
// Renumber the static rows by storage slot (sid = slot, n_static = n): the distributed solver calls it between
// batches when departures have left too many holes.  Borrows the idle state set.
extern "C" int mpm_compact_statics(mpm_ctx* ctx, void* stream) {
  if (!ctx) return MPM_E_INVALID;
  if (!ctx->state[0]) return fail(ctx, MPM_E_UNBOUND, "no buffers bound");
  if (ctx->in_batch) return fail(ctx, MPM_E_INVALID, "mpm_compact_statics: inside a batch");
  CK(cudaSetDevice(ctx->P.device));
  cudaStream_t s = (cudaStream_t)stream;
  const int n = (int)ctx->n;
  if (n > 0) {
    uint32_t* tmp = ctx->state[ctx->cur ^ 1];
    const int blocks = gs_blocks(n, 256, ctx->sm_count);
    for (int pass = 0; pass < 2; ++pass) {
      if (ctx->dim == 3) k_compact_statics<3><<<blocks, 256, 0, s>>>(ctx->state[ctx->cur], ctx->stat, n, tmp, pass, ctx->quant);
      else k_compact_statics<2><<<blocks, 256, 0, s>>>(ctx->state[ctx->cur], ctx->stat, n, tmp, pass, 0);
    }
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(s));
  }
  ctx->n_static = n;
  return MPM_OK;
}

extern "C" int mpm_set_gravity(mpm_ctx* ctx, const double* g) {
  if (!ctx || !g) return MPM_E_INVALID;
  for (int d = 0; d < 3; ++d) ctx->grav.g[d] = d < ctx->dim ? (float)g[d] : 0.f;
  return MPM_OK;
}

extern "C" int mpm_set_colliders(mpm_ctx* ctx, const mpm_collider* t, int32_t n) {
  if (!ctx || n < 0 || (n > 0 && !t)) return MPM_E_INVALID;
  if (n > MAX_COLLIDERS) return fail(ctx, MPM_E_INVALID, "too many colliders (max 64)");
  ctx->h_ct.n = n;
  for (int i = 0; i < n; ++i) {
    ColliderDev& c = ctx->h_ct.c[i];
    c.kind = t[i].kind;
    c.surface = t[i].surface;
    for (int d = 0; d < 3; ++d) { c.a[d] = (float)t[i].a[d]; c.b[d] = (float)t[i].b[d]; }
    c.r2 = (float)(t[i].b[0] * t[i].b[0]);   // radius * radius evaluated in double, rounded once (:625)
    c.friction = (float)t[i].friction;
    c.unbounded = t[i].kind == MPM_COLLIDER_BBOX ? (t[i].a[0] != 0.0) : 0;
  }
  ctx->ct_dirty = true;
  return MPM_OK;
}

// ------------------------------------------------------------------ seeding
static int seed_common(mpm_ctx* ctx, SeedArgs& a, void* stream) {
  if (!ctx->state[0]) return fail(ctx, MPM_E_UNBOUND, "no buffers bound");
  if (a.n < 0 || (size_t)(ctx->n + a.n) > ctx->cap) return fail(ctx, MPM_E_INVALID, "seed: capacity exceeded");
  if (a.n == 0) return MPM_OK;
  CK(cudaSetDevice(ctx->P.device));
  cudaStream_t s = (cudaStream_t)stream;
  a.state = ctx->state[ctx->cur];
  a.cap = ctx->cap;
  a.n0 = ctx->n;
  if (a.id_base < 0) a.id_base = ctx->n;
  a.stat = ctx->stat; a.sid0 = ctx->n_static; a.quant = ctx->quant;
  if ((size_t)(ctx->n_static + a.n) > ctx->cap) return fail(ctx, MPM_E_INVALID, "seed: static rows exceed the capacity (compact first)");
  int blocks = gs_blocks(a.n, 256, ctx->sm_count);
  // A large array of external positions is stored sorted by leaf block (ids keep the insertion order): the
  // first substep then reads block-local rows instead of gathering 26 words per particle at random.  The sort
  // borrows the binning scratch, which is free between batches; the
  // distributed solver writes its global ids by row and keeps the input order.
  if (a.mode == 0 && ctx->sort_seed && a.n >= (1 << 15) && !ctx->K.g2p2g && !ctx->slab.enabled && !ctx->in_batch &&
      ctx->P.grid_size == 4096 && (size_t)a.n <= ctx->cap) {
    const int half = ctx->P.grid_size / 2;
    const Slab none{0, INT_MIN, INT_MAX};
    if (ctx->dim == 3) k_seed_keys<3><<<blocks, 256, 0, s>>>(a.x, a.n, ctx->K.inv_dx, half, ctx->keys_a, ctx->vals_a, none, nullptr);
    else k_seed_keys<2><<<blocks, 256, 0, s>>>(a.x, a.n, ctx->K.inv_dx, half, ctx->keys_a, ctx->vals_a, none, nullptr);
    cub::DoubleBuffer<uint32_t> dk(ctx->keys_a, ctx->keys_b), dv(ctx->vals_a, ctx->vals_b);
    size_t tb = ctx->cub_bytes;
    CK(cub::DeviceRadixSort::SortPairs(ctx->cub_temp, tb, dk, dv, (int)a.n, 0, 10 * ctx->dim, s));
    a.order = dv.Current();
  }
  if (ctx->dim == 3) k_seed<3><<<blocks, 256, 0, s>>>(a); else k_seed<2><<<blocks, 256, 0, s>>>(a);
  CK(cudaGetLastError());
  ctx->n += a.n;
  ctx->n_static += a.n;
  ctx->bbox_valid = false;
  ctx->last_valid = false;
  return MPM_OK;
}
static void fill3(float* dst, const double* src, int dim, float dflt) {
  for (int d = 0; d < 3; ++d) dst[d] = (src && d < dim) ? (float)src[d] : dflt;
}

extern "C" int mpm_seed_positions(mpm_ctx* ctx, const float* x_dev, int64_t n, int32_t material, int32_t color,
                                  const double* velocity, int32_t emitter, void* stream) {
  if (!ctx || (n > 0 && !x_dev)) return MPM_E_INVALID;
  SeedArgs a{};
  a.n = n; a.material = material; a.color = color; a.emitter = emitter; a.x = x_dev; a.mode = 0; a.id_base = -1;
  fill3(a.vel, velocity, ctx->dim, 0.f);
  return seed_common(ctx, a, stream);
}

// Multi-GPU seeding: of the n positions x_dev[n][dim] keep the rows whose base block lies in this rank's slab
// (all of them without a slab), stored sorted by leaf block; row i gets the id id_base + i.  Selection, ordering
// and the count come from one radix sort of (owned ? block key : 1 << 30).  n must not exceed the bound capacity
// (the sort borrows the binning scratch); the kept rows must fit behind the live ones.
extern "C" int mpm_seed_positions_slab(mpm_ctx* ctx, const float* x_dev, int64_t n, int64_t id_base, int32_t material,
                                       int32_t color, const double* velocity, int32_t emitter, int64_t* kept_out,
                                       void* stream) {
  if (!ctx || n < 0 || (n > 0 && !x_dev) || id_base < 0) return MPM_E_INVALID;
  if (kept_out) *kept_out = 0;
  if (n == 0) return MPM_OK;
  if (!ctx->state[0]) return fail(ctx, MPM_E_UNBOUND, "no buffers bound");
  if ((size_t)n > ctx->cap) return fail(ctx, MPM_E_INVALID, "mpm_seed_positions_slab: n exceeds the bound capacity");
  if (ctx->in_batch || ctx->K.g2p2g) return fail(ctx, MPM_E_INVALID, "mpm_seed_positions_slab: not inside a batch / g2p2g");
  CK(cudaSetDevice(ctx->P.device));
  cudaStream_t s = (cudaStream_t)stream;
  unsigned long long* cnt = reinterpret_cast<unsigned long long*>(ctx->stage);
  CK(cudaMemsetAsync(cnt, 0, 8, s));
  const int blocks = gs_blocks(n, 256, ctx->sm_count), half = ctx->P.grid_size / 2;
  if (ctx->dim == 3) k_seed_keys<3><<<blocks, 256, 0, s>>>(x_dev, n, ctx->K.inv_dx, half, ctx->keys_a, ctx->vals_a, ctx->slab, cnt);
  else k_seed_keys<2><<<blocks, 256, 0, s>>>(x_dev, n, ctx->K.inv_dx, half, ctx->keys_a, ctx->vals_a, ctx->slab, cnt);
  cub::DoubleBuffer<uint32_t> dk(ctx->keys_a, ctx->keys_b), dv(ctx->vals_a, ctx->vals_b);
  size_t tb = ctx->cub_bytes;
  CK(cub::DeviceRadixSort::SortPairs(ctx->cub_temp, tb, dk, dv, (int)n, 0, 31, s));
  unsigned long long kept = 0;
  CK(cudaMemcpyAsync(&kept, cnt, 8, cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  if (kept_out) *kept_out = (int64_t)kept;
  if (kept == 0) return MPM_OK;
  if ((size_t)(ctx->n + (int64_t)kept) > ctx->cap || (size_t)(ctx->n_static + (int64_t)kept) > ctx->cap)
    return fail(ctx, MPM_E_INVALID, "mpm_seed_positions_slab: capacity exceeded");
  SeedArgs a{};
  a.n = (int64_t)kept; a.material = material; a.color = color; a.emitter = emitter; a.x = x_dev; a.mode = 0;
  a.id_base = id_base; a.order = dv.Current();
  fill3(a.vel, velocity, ctx->dim, 0.f);
  a.state = ctx->state[ctx->cur]; a.cap = ctx->cap; a.n0 = ctx->n;
  a.stat = ctx->stat; a.sid0 = ctx->n_static;
  const int sb = gs_blocks(a.n, 256, ctx->sm_count);
  if (ctx->dim == 3) k_seed<3><<<sb, 256, 0, s>>>(a); else k_seed<2><<<sb, 256, 0, s>>>(a);
  CK(cudaGetLastError());
  ctx->n += a.n;
  ctx->n_static += a.n;
  ctx->bbox_valid = false;
  ctx->last_valid = false;
  return MPM_OK;
}

// Positions of seed (:840-850, mode 1) / seed_ellipsoid (:959-978, mode 2) for the particle ids [id0, id0 + n),
// written to x_out_dev[n][dim] without appending anything: the distributed add_cube / add_ellipsoid generate the
// same points as the single-device solver on every rank and keep their slab (mpm_seed_positions_slab).
extern "C" int mpm_seed_generate(mpm_ctx* ctx, int32_t mode, int64_t n, int64_t id0, const double* a3, const double* b3,
                                 uint64_t seed, float* x_out_dev, void* stream) {
  if (!ctx || (mode != 1 && mode != 2) || n < 0 || id0 < 0 || !a3 || !b3 || (n > 0 && !x_out_dev)) return MPM_E_INVALID;
  if (n == 0) return MPM_OK;
  CK(cudaSetDevice(ctx->P.device));
  SeedArgs a{};
  a.n = n; a.mode = mode; a.seed = seed; a.id_base = id0; a.x_out = x_out_dev;
  fill3(a.a, a3, ctx->dim, 0.f);
  fill3(a.b, b3, ctx->dim, 0.f);
  const int blocks = gs_blocks(n, 256, ctx->sm_count);
  cudaStream_t s = (cudaStream_t)stream;
  if (ctx->dim == 3) k_seed<3><<<blocks, 256, 0, s>>>(a); else k_seed<2><<<blocks, 256, 0, s>>>(a);
  CK(cudaGetLastError());
  return MPM_OK;
}

// particle_info() of this rank (multi-GPU): rows [x[dim] v[dim] material color id] (2 dim + 3 words) of the
// particles this rank owns, to out_dev (room for n_particles rows); the number of rows to *count_out.
extern "C" int mpm_export_local(mpm_ctx* ctx, void* out_dev, int64_t* count_out, void* stream) {
  if (!ctx || !count_out) return MPM_E_INVALID;
  *count_out = 0;
  if (ctx->n == 0) return MPM_OK;
  if (!out_dev) return MPM_E_INVALID;
  CK(cudaSetDevice(ctx->P.device));
  cudaStream_t s = (cudaStream_t)stream;
  unsigned long long* cnt = reinterpret_cast<unsigned long long*>(ctx->stage);
  CK(cudaMemsetAsync(cnt, 0, 8, s));
  const int blocks = gs_blocks(ctx->n, 256, ctx->sm_count), half = ctx->P.grid_size / 2;
  if (ctx->dim == 3) k_export_local<3><<<blocks, 256, 0, s>>>(ctx->state[ctx->cur], ctx->stat, (int)ctx->n, ctx->K.inv_dx, half, ctx->slab, (uint32_t*)out_dev, cnt, ctx->quant);
  else k_export_local<2><<<blocks, 256, 0, s>>>(ctx->state[ctx->cur], ctx->stat, (int)ctx->n, ctx->K.inv_dx, half, ctx->slab, (uint32_t*)out_dev, cnt, 0);
  CK(cudaGetLastError());
  unsigned long long c = 0;
  CK(cudaMemcpyAsync(&c, cnt, 8, cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  *count_out = (int64_t)c;
  return MPM_OK;
}
extern "C" int mpm_seed_cube(mpm_ctx* ctx, int64_t n, const double* lower, const double* size, int32_t material,
                             int32_t color, const double* velocity, int32_t emitter, uint64_t seed, void* stream) {
  if (!ctx || !lower || !size) return MPM_E_INVALID;
  SeedArgs a{};
  a.n = n; a.material = material; a.color = color; a.emitter = emitter; a.mode = 1; a.seed = seed; a.id_base = -1;
  fill3(a.vel, velocity, ctx->dim, 0.f);
  fill3(a.a, lower, ctx->dim, 0.f);
  fill3(a.b, size, ctx->dim, 0.f);
  return seed_common(ctx, a, stream);
}
extern "C" int mpm_seed_ellipsoid(mpm_ctx* ctx, int64_t n, const double* center, const double* radius,
                                  int32_t material, int32_t color, const double* velocity, int32_t emitter,
                                  uint64_t seed, void* stream) {
  if (!ctx || !center || !radius) return MPM_E_INVALID;
  SeedArgs a{};
  a.n = n; a.material = material; a.color = color; a.emitter = emitter; a.mode = 2; a.seed = seed; a.id_base = -1;
  fill3(a.vel, velocity, ctx->dim, 0.f);
  fill3(a.a, center, ctx->dim, 0.f);
  fill3(a.b, radius, ctx->dim, 0.f);
  return seed_common(ctx, a, stream);
}
extern "C" int mpm_seed_restart(mpm_ctx* ctx, const float* x_dev, const float* v_dev, const int32_t* m_dev,
                                const int32_t* c_dev, int64_t n, void* stream) {
  if (!ctx || (n > 0 && (!x_dev || !v_dev || !m_dev || !c_dev))) return MPM_E_INVALID;
  SeedArgs a{};
  a.n = n; a.x = x_dev; a.v = v_dev; a.mats = m_dev; a.colors = c_dev; a.mode = 3; a.id_base = -1;
  return seed_common(ctx, a, stream);
}

// ------------------------------------------------------------------ substep
static int upload_colliders(mpm_ctx* ctx, cudaStream_t s) {
  if (!ctx->ct_dirty) return MPM_OK;
  CK(cudaMemcpyAsync(ctx->d_ct, &ctx->h_ct, sizeof(ColliderTable), cudaMemcpyHostToDevice, s));
  CK(cudaStreamSynchronize(s));   // h_ct is pageable host memory owned by ctx
  ctx->ct_dirty = false;
  return MPM_OK;
}

static int refresh_bbox(mpm_ctx* ctx, cudaStream_t s) {
  if (ctx->bbox_valid) return MPM_OK;
  CK(cudaMemsetAsync(ctx->d_status, 0, sizeof(Status), s));
  k_reset<<<1, 1, 0, s>>>(ctx->d_status);
  int blocks = gs_blocks(ctx->n, 256, ctx->sm_count);
  if (ctx->dim == 3) k_bbox<3><<<blocks, 256, 0, s>>>(ctx->state[ctx->cur], ctx->cap, (int)ctx->n, ctx->K.inv_dx, ctx->d_status, ctx->quant);
  else k_bbox<2><<<blocks, 256, 0, s>>>(ctx->state[ctx->cur], ctx->cap, (int)ctx->n, ctx->K.inv_dx, ctx->d_status, 0);
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(ctx->h_status, ctx->d_status, sizeof(Status), cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  for (int d = 0; d < 3; ++d) { ctx->bb_min[d] = ctx->h_status->bb_min[d]; ctx->bb_max[d] = ctx->h_status->bb_max[d]; }
  ctx->bbox_valid = true;
  return MPM_OK;
}

// Keep a sticky block-aligned box around the particles so the key layout (and
// the number of radix passes) only changes when particles approach its faces.
static int update_layout(mpm_ctx* ctx) {
  const int half = ctx->P.grid_size / 2, ll = ctx->log_leaf;
  const int MARGIN = ctx->slab.enabled ? 3 : 2;
  bool keep = ctx->layout_valid && !ctx->slab.enabled && !ctx->ext_box;
  int bmin[3], bmax[3];
  for (int d = 0; d < ctx->dim; ++d) {
    const int lo = ctx->ext_box ? ctx->box_min[d] : ctx->bb_min[d];
    const int hi = ctx->ext_box ? ctx->box_max[d] : ctx->bb_max[d];
    // NaN/inf positions (a diverged simulation) show up as an absurd box: refuse before any index math
    if (lo > hi || lo < -(1 << 28) || hi > (1 << 28) || (int64_t)hi - lo > ((int64_t)1 << 24)) {
      ctx->layout_valid = false;
      return fail(ctx, MPM_E_KEY_BITS, "particle bounding box is not finite/representable (diverged simulation?)");
    }
    bmin[d] = (lo + half) >> ll;
    bmax[d] = (hi + half) >> ll;
    if (keep && !(bmin[d] >= ctx->L.ob[d] + 1 && bmax[d] <= ctx->L.ob[d] + ctx->L.eb[d] - 3)) keep = false;
  }
  if (!keep) {
    KeyLayout L{};
    L.half = half;
    double prod = 1.0;
    for (int d = 0; d < 3; ++d) { L.ob[d] = 0; L.eb[d] = 1; }
    for (int d = 0; d < ctx->dim; ++d) {
      int first = bmin[d] - MARGIN, last = bmax[d] + MARGIN;   // particle blocks covered
      if (d == 0 && ctx->slab.enabled) {
        // this rank only bins blocks of its slab; +1 ring below covers the shared column `hi`
        if (ctx->slab.lo > INT_MIN / 2) first = std::max(first, ctx->slab.lo);
        if (ctx->slab.hi < INT_MAX / 2) last = std::min(last, ctx->slab.hi - 1);
        if (last < first) last = first;
      }
      L.ob[d] = first;
      L.eb[d] = (last + 1) - first + 1;
      prod *= (double)L.eb[d];
    }
    int bits = 0;
    while (bits < 40 && (double)(1ull << bits) < prod) ++bits;
    L.key_bits = bits + ctx->cb;
    if (L.key_bits > 32) {
      ctx->layout_valid = false;
      return fail(ctx, MPM_E_KEY_BITS, "particle bounding box needs more than 32 key bits (diverged simulation?)");
    }
    ctx->L = L;
    ctx->layout_valid = true;
  }
  double nlin = 1.0;
  for (int d = 0; d < ctx->dim; ++d) nlin *= (double)ctx->L.eb[d];
  ctx->dense = ctx->use_dense && (2.0 * nlin + 1.0 <= (double)ctx->table_cap);
  return MPM_OK;
}

// Modes that only exist on the counting-sort path (use_g2p2g, quant=True): when the flag table of the current particle box
// does not fit the table this workspace was sized for, ask the host for a larger one (the table grows with max_blocks).
static int need_dense_table(mpm_ctx* ctx, const char* what) {
  double nlin = 1.0;
  for (int d = 0; d < ctx->dim; ++d) nlin *= (double)ctx->L.eb[d];
  if (ctx->use_dense && 2.0 * nlin + 1.0 <= (double)((int64_t)1 << 24) && ctx->table_cap < ((int64_t)1 << 24)) {
    char buf[200];
    snprintf(buf, sizeof buf, "%s: the block table (%lld entries) is too small for the particle box (%.0f blocks); grow max_blocks (%d)",
             what, (long long)ctx->table_cap, nlin, ctx->max_blocks);
    ctx->err = buf;
    ctx->layout_valid = false;
    return MPM_E_BLOCK_CAPACITY;
  }
  return fail(ctx, MPM_E_KEY_BITS, std::string(what) + " needs the counting-sort path (particle box too large for the flag table)");
}

// Launch configuration of one kernel instance, per device: the dynamic shared-memory opt-in
// (cudaFuncSetAttribute) and the occupancy-derived persistent grid apply to the CURRENT device only, and a
// process may drive several GPUs from several host threads (one ctx per device).
static constexpr int MAX_DEVICES = 64;
struct LaunchCache {
  std::atomic<int> grid[MAX_DEVICES];
  LaunchCache() { for (auto& g : grid) g.store(0); }
};
template <typename K>
static int cached_grid(LaunchCache& lc, mpm_ctx* ctx, K kernel, int threads, size_t smem) {
  const int dev = std::min(std::max(ctx->P.device, 0), MAX_DEVICES - 1);
  int g = lc.grid[dev].load(std::memory_order_acquire);
  if (g) return g;
  int occ = 1;
  if (smem) cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, threads, smem);
  g = ctx->sm_count * std::max(occ, 1);
  lc.grid[dev].store(g, std::memory_order_release);   // racing threads compute the same value
  return g;
}

// P2G (cell-owner) launch configurations: particles staged per pass, CTAs per SM; MPM_P2G_CFG picks one.
template <int D, int CH, int MB>
static void launch_p2g_cfg(mpm_ctx* ctx, const SubstepArgs<D>& a, cudaStream_t s) {
  static LaunchCache lc;
  constexpr size_t smem = p2g_smem_bytes<D, CH>();
  const int grid = cached_grid(lc, ctx, k_p2g_cell<D, CH, MB>, P2GCfg<D>::THREADS, smem);
  k_p2g_cell<D, CH, MB><<<grid, P2GCfg<D>::THREADS, smem, s>>>(a);
}
// third revision (mpm_p2g3.cuh): 3D, needs the counting sort's per-cell bucket starts
template <int CH, int MB>
static void launch_p2g3_cfg(mpm_ctx* ctx, const SubstepArgs<3>& a, cudaStream_t s) {
  constexpr size_t smem = p2g3_smem_bytes<CH>();
  if (a.cb.fused) {          // multi-GPU: halo inside the kernel (mpm_comm.cuh)
    static LaunchCache lcf;
    const int grid = cached_grid(lcf, ctx, k_p2g3<CH, MB, true>, P2G3::T, smem);
    launch_chain(ctx->pdl, k_p2g3<CH, MB, true>, grid, P2G3::T, smem, s, a);
    return;
  }
  static LaunchCache lc;
  const int grid = cached_grid(lc, ctx, k_p2g3<CH, MB, false>, P2G3::T, smem);
  launch_chain(ctx->pdl, k_p2g3<CH, MB, false>, grid, P2G3::T, smem, s, a);
}
template <int D>
static void launch_p2g(mpm_ctx* ctx, const SubstepArgs<D>& a, cudaStream_t s) {
  if constexpr (D == 3) {
    if (ctx->quant == 2) {     // quant=True, split substep: the cell-owner kernel on the packed accessors
      constexpr size_t smem = p2g3_smem_bytes<640>();
      static LaunchCache lcq;
      const int grid = cached_grid(lcq, ctx, k_p2g3<640, 4, false, false, true>, P2G3::T, smem);
      launch_chain(ctx->pdl, k_p2g3<640, 4, false, false, true>, grid, P2G3::T, smem, s, a);
      return;
    }
    if (ctx->p2g_ver == 3 && a.cellstart) {
      switch (ctx->p2g_cfg) {
        case 1: launch_p2g3_cfg<768, 4>(ctx, a, s); break;
        case 2: launch_p2g3_cfg<576, 4>(ctx, a, s); break;
        case 3: launch_p2g3_cfg<544, 4>(ctx, a, s); break;
        case 4: launch_p2g3_cfg<1024, 3>(ctx, a, s); break;
        default: launch_p2g3_cfg<640, 4>(ctx, a, s); break;
      }
      return;
    }
    switch (ctx->p2g_cfg) {
      case 1: launch_p2g_cfg<3, 512, 5>(ctx, a, s); break;
      case 2: launch_p2g_cfg<3, 576, 4>(ctx, a, s); break;
      case 3: launch_p2g_cfg<3, 384, 6>(ctx, a, s); break;
      case 4: launch_p2g_cfg<3, 768, 3>(ctx, a, s); break;
      default: launch_p2g_cfg<3, 640, 4>(ctx, a, s); break;
    }
  } else {
    launch_p2g_cfg<2, 1280, 1>(ctx, a, s);
  }
}

// G2P launch configurations (threads per CTA, min CTAs per SM); MPM_G2P_CFG picks one.
template <int D, int T, int MB, bool BULK = false, int QM = 0, bool XS = true>
static void launch_g2p_cfg(mpm_ctx* ctx, const SubstepArgs<D>& a, cudaStream_t s) {
  static LaunchCache lc;
  const int grid = cached_grid(lc, ctx, k_g2p<D, T, MB, BULK, QM, XS>, T, 0);
  launch_chain(ctx->pdl, k_g2p<D, T, MB, BULK, QM, XS>, grid, T, 0, s, a);
}
template <int D>
static void launch_g2p(mpm_ctx* ctx, const SubstepArgs<D>& a, cudaStream_t s) {
  if constexpr (D == 3) {
    if (ctx->quant == 2) { launch_g2p_cfg<3, 128, 5, false, 2>(ctx, a, s); return; }   // quant=True, split substep
    // after the halo variant of k_p2g3 (which does not copy x / tag to the sorted slot): follow perm
    if (a.cb.fused) { launch_g2p_cfg<3, 128, 5, false, 0, false>(ctx, a, s); return; }
  }
  if (ctx->g2p_tile == 1) { launch_g2p_cfg<D, 128, 5, true>(ctx, a, s); return; }   // cp.async.bulk rows + mbarrier
  switch (ctx->g2p_cfg) {
    case 1: launch_g2p_cfg<D, 256, 4>(ctx, a, s); break;
    case 2: launch_g2p_cfg<D, 128, 6>(ctx, a, s); break;
    case 3: launch_g2p_cfg<D, 128, 8>(ctx, a, s); break;
    case 4: launch_g2p_cfg<D, 128, 10>(ctx, a, s); break;
    case 5: launch_g2p_cfg<D, 64, 16>(ctx, a, s); break;
    case 6: launch_g2p_cfg<D, 128, 5>(ctx, a, s); break;
    case 7: launch_g2p_cfg<D, 128, 4>(ctx, a, s); break;
    default: launch_g2p_cfg<D, 256, 3>(ctx, a, s); break;
  }
}

template <int D>
static SubstepArgs<D> make_args(mpm_ctx* ctx, float dt, int cur) {
  SubstepArgs<D> a{};
  a.src = ctx->state[cur]; a.dst = ctx->state[cur ^ 1]; a.cap = ctx->cap;
  a.keys = ctx->cur_keys; a.perm = ctx->cur_perm;
  a.pb_start = ctx->pb_start; a.pb_nbr = ctx->pb_nbr; a.grid = ctx->grid; a.st = ctx->d_status;
  a.pb_key = ctx->pb_key; a.cellstart = ctx->cur_cellstart;
  a.L = ctx->L; a.K = ctx->K; a.dt = dt;
  a.slab = ctx->slab; a.cb = ctx->comm; a.stat = ctx->stat;
  a.n_rows = (int)ctx->n;
  a.pf_mode = ctx->pf_mode;
  return a;
}

// exclusive scan of n ints: one k_scan_excl launch (optionally committing the previous substep), or CUB
static int enqueue_scan(mpm_ctx* ctx, const int* in, int* out, int n, bool commit, cudaStream_t s,
                        const int* n_dev = nullptr, int n_mul = 0) {
  // the prefix travels one 32-tile window per hop: beyond ~8 hops the chain costs more than CUB's second launch
  if (ctx->own_scan && n <= 256 * SCAN_TILE) {
    ctx->scan_epoch = (ctx->scan_epoch + 1) & 0x3fffffffu;
    if (ctx->scan_epoch == 0) ctx->scan_epoch = 1;
    const int grid = std::max(1, std::min(ctx->scan_grid, (n + SCAN_TILE - 1) / SCAN_TILE));
    if (commit) CK(launch_chain(ctx->pdl, k_scan_excl<true>, grid, SCAN_T, 0, s, in, out, n, ctx->scan_desc, ctx->scan_epoch, ctx->d_status, n_dev, n_mul));
    else CK(launch_chain(ctx->pdl, k_scan_excl<false>, grid, SCAN_T, 0, s, in, out, n, ctx->scan_desc, ctx->scan_epoch, ctx->d_status, n_dev, n_mul));
    return MPM_OK;
  }
  if (commit) CK(launch_chain(ctx->pdl, k_substep_begin, 1, 1, 0, s, ctx->d_status));
  size_t tb = ctx->cub_bytes;
  CK(cub::DeviceScan::ExclusiveSum(ctx->cub_temp, tb, in, out, n, s));
  return MPM_OK;
}

// bucket starts from the per-cell counts of the npb = fscan[nlin] existing particle blocks
template <int D>
static int enqueue_cell_scan(mpm_ctx* ctx, int nlin, cudaStream_t s) {
  using G = Geo<D>;
  const int ncell = ctx->max_blocks * G::CELLS + 1;
  if (!ctx->own_scan || (ncell <= 256 * SCAN_TILE && !ctx->scan2_force))   // one look-back chain (or MPM_SCAN=cub)
    return enqueue_scan(ctx, ctx->cellcount, ctx->cellstart, ncell, false, s, ctx->fscan + nlin, G::CELLS);
  const int blocks = gs_blocks((int64_t)ctx->max_blocks * 32, 256, ctx->sm_count);
  CK(launch_chain(ctx->pdl, k_cell_sums<D>, blocks, 256, 0, s, (const int*)ctx->cellcount, (const int*)(ctx->fscan + nlin),
                  ctx->blocksum, (const Status*)ctx->d_status));
  int rc = enqueue_scan(ctx, ctx->blocksum, ctx->blockstart, ctx->max_blocks + 1, false, s, ctx->fscan + nlin, 1);
  if (rc) return rc;
  CK(launch_chain(ctx->pdl, k_cell_starts<D>, blocks, 256, 0, s, (const int*)ctx->cellcount, (const int*)(ctx->fscan + nlin),
                  (const int*)ctx->blockstart, ctx->cellstart, (const Status*)ctx->d_status));
  ctx->launches += 2;
  return MPM_OK;
}

template <int D>
static int enqueue_bin_p2g(mpm_ctx* ctx, float dt, int cur, int commit_prev, cudaStream_t s, cudaEvent_t* ev,
                           bool fuse_next = false) {
  const bool prof = ev != nullptr;
  using G = Geo<D>;
  const int n = (int)ctx->n;
  const int sm = ctx->sm_count;
  Status* st = ctx->d_status;
  const uint32_t* src = ctx->state[cur];
  if (prof && !ctx->in_batch) cudaEventRecord(ev[0], s);   // (phase API: recorded before the unpack)
  const uint32_t* keys = nullptr;
  const uint32_t* perm = nullptr;
  const int* cellstart = nullptr;
  size_t tb;
  if (ctx->dense) {
    // ---- counting sort on the (block, cell) key over a dense flag table (mpm_bin.cuh)
    int nlin = 1;
    for (int d = 0; d < D; ++d) nlin *= ctx->L.eb[d];
    const int ncell = ctx->max_blocks * G::CELLS + 1;
    if (!ctx->cell_zeroed) CK(cudaMemsetAsync(ctx->cellcount, 0, (size_t)ncell * 4, s));
    ctx->cell_zeroed = false;
    const bool fused_keys = ctx->keys_ready;   // keys and flags were written by the previous substep's G2P
    if (!fused_keys) {
      CK(cudaMemsetAsync(ctx->flags, 0, (size_t)(2 * nlin + 1) * 4, s));
      bool done = false;
      if constexpr (D == 3) {
        if (ctx->quant == 2) {
          k_bin_keys<3, 2><<<gs_blocks((n + 3) / 4, 256, sm), 256, 0, s>>>(src, ctx->cap, ctx->K.inv_dx, ctx->L, ctx->slab, ctx->keys_a,
                                                                       ctx->flags, nlin, commit_prev, st);
          done = true;
        }
      }
      if (!done)
        k_bin_keys<D><<<gs_blocks((n + 3) / 4, 256, sm), 256, 0, s>>>(src, ctx->cap, ctx->K.inv_dx, ctx->L, ctx->slab, ctx->keys_a, ctx->flags,
                                                           nlin, commit_prev, st);
    }
    ctx->keys_ready = false;
    // (a substep whose keys came from G2P is committed by the scan's first thread)
    { int rc = enqueue_scan(ctx, ctx->flags, ctx->fscan, 2 * nlin + 1, fused_keys, s); if (rc) return rc; }
    CK(launch_chain(ctx->pdl, k_bin_rank<D>, gs_blocks((n + 3) / 4, 256, sm), 256, 0, s, ctx->keys_a, ctx->fscan,
                    ctx->cellcount, ctx->vals_a, ctx->pb_key, ctx->max_blocks, st));
    // (only the cells of the npb = fscan[nlin] existing blocks: the table is sized by capacity)
    { int rc = enqueue_cell_scan<D>(ctx, nlin, s); if (rc) return rc; }
    CK(launch_chain(ctx->pdl, k_bin_scatter<D>, gs_blocks((n + 3) / 4, 256, sm), 256, 0, s, ctx->keys_a, ctx->vals_a,
                    ctx->fscan, ctx->cellstart, ctx->vals_b, st));
    CK(launch_chain(ctx->pdl, k_bin_finish<D>, gs_blocks((int64_t)ctx->max_blocks * G::NO, 256, sm), 256, 0, s,
                    ctx->flags, ctx->fscan, nlin, ctx->L, ctx->pb_key, ctx->cellstart, ctx->pb_start, ctx->pb_nbr,
                    ctx->gb_key, ctx->max_blocks, st, ctx->slab));
    keys = ctx->keys_a;
    perm = ctx->vals_b;
    cellstart = ctx->cellstart;
    // own kernels: [keys | commit] rank scatter finish, plus the two scans when they are ours (CUB's are not counted)
    ctx->launches += ctx->own_scan ? (fused_keys ? 5 : 6) : 4;   // (large tables: a scan may still go to CUB)
  } else {
    // ---- fallback: multi-pass LSD radix sort + sorted-candidate block list
    if (ctx->quant) return fail(ctx, MPM_E_KEY_BITS, "quant=True needs the counting-sort path (particle box too large for the flag table)");
    if (commit_prev) k_end<<<1, 1, 0, s>>>(st);
    k_reset<<<1, 1, 0, s>>>(st);
    k_keys<D><<<gs_blocks(n, 256, sm), 256, 0, s>>>(src, ctx->cap, n, ctx->K.inv_dx, ctx->L, ctx->keys_a, ctx->vals_a, st);
    cub::DoubleBuffer<uint32_t> dk(ctx->keys_a, ctx->keys_b), dv(ctx->vals_a, ctx->vals_b);
    tb = ctx->cub_bytes;
    CK(cub::DeviceRadixSort::SortPairs(ctx->cub_temp, tb, dk, dv, n, 0, ctx->L.key_bits, s));
    keys = dk.Current();
    perm = dv.Current();
    tb = ctx->cub_bytes;
    thrust::counting_iterator<int> it(0);
    CK(cub::DeviceSelect::If(ctx->cub_temp, tb, it, BoundedOut{ctx->pb_start, ctx->max_blocks + 1}, &st->npb, n,
                             HeadOp{keys, G::CB}, s));
    k_pb_finalize<<<1, 1, 0, s>>>(st, ctx->pb_start, n, ctx->max_blocks);
    k_pb_masks<D><<<gs_blocks((int64_t)ctx->max_blocks * 32, 256, sm), 256, 0, s>>>(keys, ctx->pb_start, ctx->L, ctx->max_blocks,
                                                                                ctx->pb_mask, ctx->cand_a, ctx->pb_key, st);
    const int ncand = ctx->max_blocks * G::NO;
    cub::DoubleBuffer<uint32_t> dc(ctx->cand_a, ctx->cand_b);
    tb = ctx->cub_bytes;
    CK(cub::DeviceRadixSort::SortKeys(ctx->cub_temp, tb, dc, ncand, 0, 32, s));
    tb = ctx->cub_bytes;
    CK(cub::DeviceSelect::Unique(ctx->cub_temp, tb, dc.Current(), ctx->gb_key, &st->ngb_raw, ncand, s));
    k_gb_finalize<<<1, 1, 0, s>>>(st, ctx->gb_key, ctx->max_blocks);
    k_nbr<D><<<gs_blocks((int64_t)ctx->max_blocks * G::NO, 256, sm), 256, 0, s>>>(keys, ctx->pb_start, ctx->pb_mask, ctx->gb_key,
                                                                              ctx->L, ctx->pb_nbr, st);
    ctx->launches += 7 + (commit_prev ? 1 : 0);
  }
  {
    // the same pass clears what the next substep of the batch needs zeroed (no memset nodes in the chain)
    int *z1 = nullptr, *z2 = nullptr, n1 = 0, n2 = 0;
    if (ctx->dense) {
      z1 = ctx->cellcount; n1 = ctx->max_blocks * G::CELLS + 1;
      ctx->cell_zeroed = true;
      if (fuse_next) {
        int nlin = 1;
        for (int d = 0; d < D; ++d) nlin *= ctx->L.eb[d];
        z2 = ctx->flags; n2 = 2 * nlin + 1;
        ctx->flags_zeroed = true;
      }
    }
    float4 *zp0 = nullptr, *zp1 = nullptr;
    int nzp = 0;
    if (ctx->comm.fused) {
      const size_t off = (size_t)((ctx->comm.epoch + 1u) % 3u) * ctx->comm.plane_blocks * HALO_SLAB;
      if (ctx->comm.plane_in[0]) zp0 = ctx->comm.plane_in[0] + off;
      if (ctx->comm.plane_in[1]) zp1 = ctx->comm.plane_in[1] + off;
      nzp = D == 3 ? ctx->L.eb[1] * ctx->L.eb[2] * HALO_SLAB : 0;
    }
    CK(launch_chain(ctx->pdl, k_clear_grid<D>, gs_blocks((int64_t)ctx->max_blocks * G::CELLS, 256, sm), 256, 0, s,
                    ctx->grid, (const Status*)st, z1, n1, z2, n2, (int)G::CELLS, zp0, zp1, nzp));
  }
  if (prof) cudaEventRecord(ev[1], s);
  ctx->cur_keys = keys; ctx->cur_perm = perm; ctx->cur_cellstart = cellstart;
  SubstepArgs<D> a = make_args<D>(ctx, dt, cur);
  if (ctx->p2g_variant == 0) k_p2g<D><<<ctx->grid_p2g, P2G_THREADS, 0, s>>>(a);
  else launch_p2g<D>(ctx, a, s);
  if (prof) cudaEventRecord(ev[2], s);
  CK(cudaGetLastError());
  ctx->launches += 2;   // clear, p2g; CUB's internal launches are not counted
  ctx->cur_keys = keys; ctx->cur_perm = perm; ctx->cur_cellstart = cellstart;
  ctx->last_keys = keys;
  return MPM_OK;
}

template <int D>
static int enqueue_grid_op(mpm_ctx* ctx, float dt, cudaStream_t s, int* zero = nullptr, int nzero = 0) {
  using G = Geo<D>;
  CK(launch_chain(ctx->pdl, k_grid_op<D>, gs_blocks((int64_t)ctx->max_blocks * G::CELLS, 256, ctx->sm_count), 256, 0, s,
                  ctx->grid, (const uint32_t*)ctx->gb_key, ctx->L, (const ColliderTable*)ctx->d_ct, ctx->grav, ctx->gcfg,
                  ctx->K.dx, dt, (ctx->K.g2p2g && ctx->K.v_allowed_cfl > 0.f) ? ctx->K.v_allowed_cfl / dt : 0.f,
                  ctx->d_status, zero, nzero, ctx->slab, ctx->comm));
  CK(cudaGetLastError());
  ctx->launches += 1;
  return MPM_OK;
}

template <int D>
static int enqueue_grid_g2p(mpm_ctx* ctx, float dt, int cur, cudaStream_t s, cudaEvent_t* ev, bool fuse_next = false) {
  const bool prof = ev != nullptr;
  Status* st = ctx->d_status;
  SubstepArgs<D> a = make_args<D>(ctx, dt, cur);
  int nlin = 1;
  for (int d = 0; d < D; ++d) nlin *= ctx->L.eb[d];
  // flag table for the fused key pass: cleared by k_clear_grid (single GPU) or, with slabs, by the grid op
  // (k_halo_add still reads this substep's flags after P2G) -- no memset node between the kernels
  const bool zero_here = fuse_next && !ctx->flags_zeroed;
  int rc = enqueue_grid_op<D>(ctx, dt, s, zero_here ? ctx->flags : nullptr, zero_here ? 2 * nlin + 1 : 0);
  if (rc) return rc;
  if (prof) cudaEventRecord(ev[3], s);
  if (fuse_next) {
    // another substep of this batch follows with the same key layout: let G2P emit its keys/flags
    ctx->flags_zeroed = false;
    a.next_keys = ctx->keys_a; a.next_flags = ctx->flags; a.next_nlin = nlin;
    ctx->keys_ready = true;
  }
  launch_g2p<D>(ctx, a, s);
  if (ctx->slab.enabled && !ctx->comm.fused) {   // (fused exchange: G2P's last CTA writes the headers)
    CK(launch_chain(ctx->pdl, k_mig_headers, 1, 1, 0, s, ctx->comm, (uint32_t)(ctx->epoch + 1), st));
    ctx->launches += 1;
  }
  if (prof) cudaEventRecord(ev[4], s);
  CK(cudaGetLastError());
  ctx->launches += 1;   // g2p
  return MPM_OK;
}

// ---- use_g2p2g (engine/mpm_solver.py:773-787): key pass, binning of the advected positions, ONE fused kernel,
// grid op -- see mpm_g2p2g.cuh.  Every substep switches to the other set of block structure + grid; the set it
// leaves behind is the next substep's input.
template <int D>
static int enqueue_fused_substep(mpm_ctx* ctx, float dt, int cur, cudaStream_t s, bool have_prev, int64_t n_old,
                                 int npb_host) {
  using G = Geo<D>;
  const int n = (int)ctx->n, sm = ctx->sm_count;
  Status* st = ctx->d_status;
  const int q_in = ctx->sel, q_out = ctx->sel ^ 1;
  int nlin = 1;
  for (int d = 0; d < D; ++d) nlin *= ctx->L.eb[d];
  FusedArgs<D> fa{};
  fa.s = make_args<D>(ctx, dt, cur);                       // (set q_in: last substep's particle blocks)
  fa.grid_in = have_prev ? ctx->grid2[q_in] : nullptr;
  fa.tin = GridTable{ctx->flags2[q_in], ctx->fscan2[q_in], ctx->fused_nlin, ctx->fused_L};
  fa.n_old = have_prev ? (int)n_old : 0;
  fa.keys = ctx->keys_a; fa.flags = ctx->flags2[q_out]; fa.nlin = nlin;
  // ---- key pass
  CK(cudaMemsetAsync(ctx->flags2[q_out], 0, (size_t)(2 * nlin + 1) * 4, s));
  const bool packed = D == 3 && ctx->quant;
  if (have_prev && ctx->fused_npb > 0) {   // (npb_host < 0: the block count of the substep enqueued just before is on the device)
    if constexpr (D == 3) {
      if (packed) k_g2p2g_keys<3, true><<<std::min(ctx->fused_npb, sm * 16), 128, 0, s>>>(fa, npb_host);
      else k_g2p2g_keys<3, false><<<std::min(ctx->fused_npb, sm * 16), 128, 0, s>>>(fa, npb_host);
    } else {
      k_g2p2g_keys<D, false><<<std::min(ctx->fused_npb, sm * 16), 128, 0, s>>>(fa, npb_host);
    }
  }
  if (fa.n_old < n) {
    if constexpr (D == 3) {
      if (packed) k_g2p2g_keys_tail<3, true><<<gs_blocks(n - fa.n_old, 256, sm), 256, 0, s>>>(fa, fa.n_old, n);
      else k_g2p2g_keys_tail<3, false><<<gs_blocks(n - fa.n_old, 256, sm), 256, 0, s>>>(fa, fa.n_old, n);
    } else {
      k_g2p2g_keys_tail<D, false><<<gs_blocks(n - fa.n_old, 256, sm), 256, 0, s>>>(fa, fa.n_old, n);
    }
  }
  ctx->launches += 2;
  // ---- binning of the advected positions into set q_out (counting sort, mpm_bin.cuh)
  select_set(ctx, q_out);
  const int ncell = ctx->max_blocks * G::CELLS + 1;
  CK(cudaMemsetAsync(ctx->cellcount, 0, (size_t)ncell * 4, s));
  { int rc = enqueue_scan(ctx, ctx->flags, ctx->fscan, 2 * nlin + 1, false, s); if (rc) return rc; }
  CK(launch_chain(ctx->pdl, k_bin_rank<D>, gs_blocks((n + 3) / 4, 256, sm), 256, 0, s, ctx->keys_a, ctx->fscan,
                  ctx->cellcount, ctx->vals_a, ctx->pb_key, ctx->max_blocks, st));
  { int rc = enqueue_cell_scan<D>(ctx, nlin, s); if (rc) return rc; }
  CK(launch_chain(ctx->pdl, k_bin_scatter<D>, gs_blocks((n + 3) / 4, 256, sm), 256, 0, s, ctx->keys_a, ctx->vals_a,
                  ctx->fscan, ctx->cellstart, ctx->vals_b, st));
  CK(launch_chain(ctx->pdl, k_bin_finish<D>, gs_blocks((int64_t)ctx->max_blocks * G::NO, 256, sm), 256, 0, s,
                  ctx->flags, ctx->fscan, nlin, ctx->L, ctx->pb_key, ctx->cellstart, ctx->pb_start, ctx->pb_nbr,
                  ctx->gb_key, ctx->max_blocks, st, ctx->slab));
  CK(launch_chain(ctx->pdl, k_clear_grid<D>, gs_blocks((int64_t)ctx->max_blocks * G::CELLS, 256, sm), 256, 0, s,
                  ctx->grid, (const Status*)st, (int*)nullptr, 0, (int*)nullptr, 0, 0, (float4*)nullptr, (float4*)nullptr, 0));
  ctx->launches += 6;
  // ---- the fused kernel over the NEW blocks, then the grid op on the output grid
  ctx->cur_keys = ctx->keys_a; ctx->cur_perm = ctx->vals_b; ctx->cur_cellstart = ctx->cellstart;
  ctx->last_keys = ctx->keys_a;
  fa.s = make_args<D>(ctx, dt, cur);                       // (set q_out)
  bool fast = false;
  if constexpr (D == 3) {
    constexpr size_t smem = p2g3_smem_bytes<640>();
    if (packed) {              // packed storage: the cell-owner kernel on the quantised accessors
      static LaunchCache lcq;
      const int grid = cached_grid(lcq, ctx, k_p2g3<640, 3, false, true, true>, P2G3::T, smem);
      CK(launch_chain(ctx->pdl, k_p2g3<640, 3, false, true, true>, grid, P2G3::T, smem, s, fa));
      fast = true;
    } else if (ctx->fused_fast) {     // cell-owner scatter (mpm_p2g3.cuh); MPM_G2P2G=simple selects the general kernel
      static LaunchCache lc;
      const int grid = cached_grid(lc, ctx, k_p2g3<640, 3, false, true>, P2G3::T, smem);
      CK(launch_chain(ctx->pdl, k_p2g3<640, 3, false, true>, grid, P2G3::T, smem, s, fa));
      fast = true;
    }
  }
  if (!fast) {
    int occ = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_g2p2g<D>, 256, 0);
    CK(launch_chain(ctx->pdl, k_g2p2g<D>, sm * std::max(occ, 1), 256, 0, s, fa));
  }
  int rc = enqueue_grid_op<D>(ctx, dt, s);
  if (rc) return rc;
  CK(launch_chain(ctx->pdl, k_fused_commit, 1, 1, 0, s, st));
  ctx->launches += 2;
  CK(cudaGetLastError());
  return MPM_OK;
}

static int substeps_g2p2g(mpm_ctx* ctx, float dt, int count, cudaStream_t s) {
  const bool d3 = ctx->dim == 3;
  for (int attempt = 0; attempt < 4; ++attempt) {
    int rc = refresh_bbox(ctx, s);
    if (rc) return rc;
    rc = update_layout(ctx);
    if (rc) return rc;
    if (!ctx->dense) return need_dense_table(ctx, "use_g2p2g");
    CK(cudaMemsetAsync(ctx->d_status, 0, sizeof(Status), s));
    k_batch_begin<<<1, 1, 0, s>>>(ctx->d_status, (int)ctx->n, (int)ctx->n_static);
    const int cur0 = ctx->cur, sel0 = ctx->sel;
    const bool prev0 = ctx->fused_prev;
    const int64_t nold0 = ctx->fused_n_old;
    const KeyLayout L0 = ctx->fused_L;
    const int nlin0 = ctx->fused_nlin, npb0 = ctx->fused_npb;
    int nlin = 1;
    for (int d = 0; d < ctx->dim; ++d) nlin *= ctx->L.eb[d];
    // enqueue optimistically: substep i reads what substep i - 1 wrote; a substep that refuses to run (capacity,
    // box) makes every later kernel a no-op and the host bookkeeping is rewound to the last completed one
    for (int i = 0; i < count; ++i) {
      const bool have_prev = i > 0 || prev0;
      if (i > 0) {
        // structure of the substep just enqueued: its block count only exists on the device (Status::npb)
        ctx->fused_L = ctx->L; ctx->fused_nlin = nlin; ctx->fused_npb = ctx->max_blocks; ctx->fused_n_old = ctx->n;
      }
      rc = d3 ? enqueue_fused_substep<3>(ctx, dt, cur0 ^ (i & 1), s, have_prev, i > 0 ? ctx->n : nold0, i > 0 ? -1 : npb0)
              : enqueue_fused_substep<2>(ctx, dt, cur0 ^ (i & 1), s, have_prev, i > 0 ? ctx->n : nold0, i > 0 ? -1 : npb0);
      if (rc) return rc;
    }
    CK(cudaMemcpyAsync(ctx->h_status, ctx->d_status, sizeof(Status), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    const Status& h = *ctx->h_status;
    ctx->cur = cur0 ^ (h.done & 1);
    select_set(ctx, sel0 ^ (h.done & 1));
    count -= h.done;
    ctx->done_last += h.done;
    ctx->bbox_valid = false;
    if (h.done > 0) {
      ctx->last = h; ctx->lastL = ctx->L; ctx->last_valid = (h.err == 0);
      ctx->fused_prev = true; ctx->fused_L = ctx->L; ctx->fused_nlin = nlin; ctx->fused_n_old = ctx->n;
      // (after an error the status block still describes the last COMPLETED substep only if it failed before
      // k_bin_finish rewrote npb / ngb, which is where capacity and box errors are raised)
      ctx->fused_npb = h.npb; ctx->fused_ngb = h.ngb;
    } else {
      ctx->fused_prev = prev0; ctx->fused_L = L0; ctx->fused_nlin = nlin0; ctx->fused_npb = npb0; ctx->fused_n_old = nold0;
    }
    if (!h.err) {
      for (int d = 0; d < 3; ++d) { ctx->bb_min[d] = h.bb_min[d]; ctx->bb_max[d] = h.bb_max[d]; }
      ctx->bbox_valid = true;
      return MPM_OK;
    }
    ctx->layout_valid = false;
    if (h.err & ERR_BLOCK_CAPACITY) {
      ctx->last.need_blocks = h.need_blocks;
      char buf[160];
      snprintf(buf, sizeof buf, "active leaf blocks (%d) exceed the bound capacity (%d)", h.need_blocks, ctx->max_blocks);
      ctx->err = buf;
      return MPM_E_BLOCK_CAPACITY;
    }
    // ERR_BBOX: the advected positions left the sticky box; rebuild it around them and go on
  }
  return fail(ctx, MPM_E_INVALID, "g2p2g substep could not establish a key layout");
}

template <int D>
static int enqueue_substep(mpm_ctx* ctx, float dt, int cur, int commit_prev, cudaStream_t s, cudaEvent_t* ev,
                           bool more_follow = false) {
  const bool fuse = more_follow && ctx->fuse_keys && ctx->dense && !ctx->slab.enabled && !ctx->K.g2p2g;
  int rc = enqueue_bin_p2g<D>(ctx, dt, cur, commit_prev, s, ev, fuse);
  if (rc) return rc;
  return enqueue_grid_g2p<D>(ctx, dt, cur, s, ev, fuse);
}

extern "C" int mpm_substeps(mpm_ctx* ctx, double dt, double t, int32_t count, void* stream) {
  (void)t;   // built-in colliders ignore t (engine/mpm_solver.py:621-640, 660-685)
  if (!ctx || count < 0) return MPM_E_INVALID;
  if (!ctx->state[0]) return fail(ctx, MPM_E_UNBOUND, "no buffers bound");
  ctx->launches = 0;
  ctx->done_last = 0;
  if (count == 0 || ctx->n == 0) return MPM_OK;
  CK(cudaSetDevice(ctx->P.device));
  cudaStream_t s = (cudaStream_t)stream;
  int rc = upload_colliders(ctx, s);
  if (rc) return rc;
  if (ctx->K.g2p2g) {
    if (ctx->slab.enabled) return fail(ctx, MPM_E_INVALID, "g2p2g mode is single-GPU");
    return substeps_g2p2g(ctx, (float)dt, count, s);
  }
  for (int attempt = 0; attempt < 3; ++attempt) {
    rc = refresh_bbox(ctx, s);
    if (rc) return rc;
    rc = update_layout(ctx);
    if (rc) return rc;
    if (ctx->quant && !ctx->dense) return need_dense_table(ctx, "quant=True");
    CK(cudaMemsetAsync(ctx->d_status, 0, sizeof(Status), s));
    k_batch_begin<<<1, 1, 0, s>>>(ctx->d_status, (int)ctx->n, (int)ctx->n_static);
    const int cur0 = ctx->cur;
    ctx->keys_ready = false;
    ctx->cell_zeroed = false; ctx->flags_zeroed = false;
    const bool prof = ctx->profiling && count <= 4096;
    if (prof)
      while ((int)ctx->ev.size() < 5 * count) {
        cudaEvent_t e;
        CK(cudaEventCreate(&e));
        ctx->ev.push_back(e);
      }
    const int enq = count;
    for (int i = 0; i < count; ++i) {
      cudaEvent_t* ev = prof ? ctx->ev.data() + 5 * i : nullptr;
      rc = ctx->dim == 3 ? enqueue_substep<3>(ctx, (float)dt, cur0 ^ (i & 1), i > 0, s, ev, i + 1 < count)
                         : enqueue_substep<2>(ctx, (float)dt, cur0 ^ (i & 1), i > 0, s, ev, i + 1 < count);
      if (rc) return rc;
    }
    k_end<<<1, 1, 0, s>>>(ctx->d_status);   // commit of the last substep
    ctx->launches += 1;
    CK(cudaMemcpyAsync(ctx->h_status, ctx->d_status, sizeof(Status), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    const Status& h = *ctx->h_status;
    ctx->cur = cur0 ^ (h.done & 1);
    count -= h.done;
    ctx->done_last += h.done;
    if (h.done > 0) {
      ctx->last = h;
      ctx->lastL = ctx->L;
      ctx->last_valid = (h.err == 0);
    }
    if (prof && !h.err) {   // per-phase device time, averaged over the batch
      for (int i = 0; i < 4; ++i) ctx->ms[i] = 0.f;
      for (int k = 0; k < enq; ++k)
        for (int i = 0; i < 4; ++i) {
          float t = 0.f;
          cudaEventElapsedTime(&t, ctx->ev[5 * k + i], ctx->ev[5 * k + i + 1]);
          ctx->ms[i] += t / enq;
        }
    }
    if (!h.err) {
      for (int d = 0; d < 3; ++d) { ctx->bb_min[d] = h.bb_min[d]; ctx->bb_max[d] = h.bb_max[d]; }
      ctx->bbox_valid = true;
      return MPM_OK;
    }
    // a substep refused to run: nothing was modified past `done`
    ctx->bbox_valid = false;
    ctx->layout_valid = false;
    if (h.err & ERR_BLOCK_CAPACITY) {
      ctx->last.need_blocks = h.need_blocks;
      char buf[160];
      snprintf(buf, sizeof buf, "active leaf blocks (%d) exceed the bound capacity (%d)", h.need_blocks, ctx->max_blocks);
      ctx->err = buf;
      return MPM_E_BLOCK_CAPACITY;
    }
    // ERR_BBOX: particles left the sticky box inside the batch; rebuild it and go on
  }
  return fail(ctx, MPM_E_INVALID, "substep could not establish a key layout");
}

extern "C" int mpm_substep(mpm_ctx* ctx, double dt, double t, void* stream) {
  return mpm_substeps(ctx, dt, t, 1, stream);
}


// ------------------------------------------------------------------ phase API (multi-GPU driver)
extern "C" int mpm_set_slab(mpm_ctx* ctx, int32_t enabled, int32_t lo_block, int32_t hi_block) {
  if (!ctx || (enabled && lo_block >= hi_block)) return fail(ctx, MPM_E_INVALID, "mpm_set_slab: empty slab");
  if (enabled && ctx->quant) return fail(ctx, MPM_E_INVALID, "mpm_set_slab: the slab decomposition runs on f32 storage (quant=False)");
  ctx->slab.enabled = enabled ? 1 : 0;
  ctx->slab.lo = enabled ? lo_block : INT_MIN;
  ctx->slab.hi = enabled ? hi_block : INT_MAX;
  ctx->layout_valid = false;
  return MPM_OK;
}

extern "C" size_t mpm_comm_bytes(int32_t dim, int32_t kind, int32_t capacity) {
  if ((dim != 2 && dim != 3) || capacity < 0) return 0;
  const size_t nf = (size_t)mpm_virtual_fields(dim), cells = dim == 3 ? 64 : 256;   // a migrating row carries its static attributes
  if (kind == 0) return 4 * (COMM_HEADER + nf * (size_t)capacity);
  return 4 * (COMM_HEADER + (size_t)capacity + (size_t)capacity * cells * 4);
}

extern "C" int mpm_bind_comm(mpm_ctx* ctx, void* mig_lo, void* mig_hi, int32_t mig_cap, void* halo_lo, void* halo_hi,
                             int32_t halo_cap) {
  if (!ctx || mig_cap < 0 || halo_cap < 0) return MPM_E_INVALID;
  ctx->comm.mig[0] = (uint32_t*)mig_lo; ctx->comm.mig[1] = (uint32_t*)mig_hi; ctx->comm.mig_cap = mig_cap;
  ctx->comm.halo[0] = (uint32_t*)halo_lo; ctx->comm.halo[1] = (uint32_t*)halo_hi; ctx->comm.halo_cap = halo_cap;
  return MPM_OK;
}

extern "C" int mpm_get_bbox(mpm_ctx* ctx, int32_t* bb_min, int32_t* bb_max, void* stream) {
  if (!ctx || !bb_min || !bb_max) return MPM_E_INVALID;
  if (!ctx->state[0]) return fail(ctx, MPM_E_UNBOUND, "no buffers bound");
  CK(cudaSetDevice(ctx->P.device));
  if (ctx->n == 0) {
    for (int d = 0; d < 3; ++d) { bb_min[d] = INT_MAX; bb_max[d] = INT_MIN; }
    return MPM_OK;
  }
  int rc = refresh_bbox(ctx, (cudaStream_t)stream);
  if (rc) return rc;
  for (int d = 0; d < 3; ++d) { bb_min[d] = ctx->bb_min[d]; bb_max[d] = ctx->bb_max[d]; }
  return MPM_OK;
}

extern "C" int mpm_set_layout_box(mpm_ctx* ctx, int32_t enabled, const int32_t* bb_min, const int32_t* bb_max) {
  if (!ctx || (enabled && (!bb_min || !bb_max))) return MPM_E_INVALID;
  ctx->ext_box = enabled != 0;
  for (int d = 0; d < 3 && enabled; ++d) { ctx->box_min[d] = bb_min[d]; ctx->box_max[d] = bb_max[d]; }
  ctx->layout_valid = false;
  return MPM_OK;
}

extern "C" int mpm_batch_begin(mpm_ctx* ctx, void* stream) {
  if (!ctx) return MPM_E_INVALID;
  if (!ctx->state[0]) return fail(ctx, MPM_E_UNBOUND, "no buffers bound");
  CK(cudaSetDevice(ctx->P.device));
  cudaStream_t s = (cudaStream_t)stream;
  ctx->launches = 0;
  ctx->done_last = 0;
  int rc = upload_colliders(ctx, s);
  if (rc) return rc;
  if (!ctx->ext_box) {
    if (ctx->n == 0) return fail(ctx, MPM_E_INVALID, "mpm_batch_begin: no particles and no layout box");
    rc = refresh_bbox(ctx, s);
    if (rc) return rc;
  }
  rc = update_layout(ctx);
  if (rc) return rc;
  if (!ctx->dense) return fail(ctx, MPM_E_INVALID, "phase API needs the counting-sort path (box too large for the flag table)");
  CK(cudaMemsetAsync(ctx->d_status, 0, sizeof(Status), s));
  k_batch_begin<<<1, 1, 0, s>>>(ctx->d_status, (int)ctx->n, (int)ctx->n_static);
  ctx->in_batch = true;
  ctx->prof_enq = 0;
  ctx->batch_cur0 = ctx->cur;
  ctx->batch_enq = 0;
  ctx->keys_ready = false; ctx->cell_zeroed = false; ctx->flags_zeroed = false;
  return MPM_OK;
}

#define REQUIRE_BATCH()                                                                   \
  if (!ctx) return MPM_E_INVALID;                                                         \
  if (!ctx->in_batch) return fail(ctx, MPM_E_INVALID, "phase call outside mpm_batch_begin/end"); \
  CK(cudaSetDevice(ctx->P.device));                                                       \
  cudaStream_t s = (cudaStream_t)stream;

extern "C" int mpm_phase_unpack(mpm_ctx* ctx, const void* from_lo, const void* from_hi, void* stream) {
  REQUIRE_BATCH();
  if (!from_lo && !from_hi) return MPM_OK;
  const int cur = ctx->batch_cur0 ^ (ctx->batch_enq & 1);
  const int blocks = gs_blocks((int64_t)ctx->comm.mig_cap * ctx->nv, 256, ctx->sm_count);
  // the last G2P emitted the coming substep's keys and flags: the appended rows need theirs too
  uint32_t* keys = ctx->keys_ready ? ctx->keys_a : nullptr;
  int nlin = 1;
  for (int d = 0; d < ctx->dim; ++d) nlin *= ctx->L.eb[d];
  if (ctx->dim == 3)
    CK(launch_chain(ctx->pdl, k_mig_unpack<3>, blocks, 256, 0, s, ctx->state[cur], ctx->cap, (uint32_t*)from_lo,
                    (uint32_t*)from_hi, ctx->comm.mig_cap, ctx->d_status, keys, ctx->flags, nlin, ctx->L, ctx->slab,
                    ctx->K.inv_dx, ctx->comm, ctx->stat));
  else
    CK(launch_chain(ctx->pdl, k_mig_unpack<2>, blocks, 256, 0, s, ctx->state[cur], ctx->cap, (uint32_t*)from_lo,
                    (uint32_t*)from_hi, ctx->comm.mig_cap, ctx->d_status, keys, ctx->flags, nlin, ctx->L, ctx->slab,
                    ctx->K.inv_dx, ctx->comm, ctx->stat));
  ctx->launches += 1;
  if (!ctx->comm.fused) {   // (fused exchange: the last CTA of the unpack kernel commits)
    CK(launch_chain(ctx->pdl, k_mig_commit, 1, 1, 0, s, (uint32_t*)from_lo, (uint32_t*)from_hi, ctx->comm.mig_cap, ctx->d_status));
    ctx->launches += 1;
  }
  CK(cudaGetLastError());
  return MPM_OK;
}

extern "C" int mpm_phase_p2g(mpm_ctx* ctx, double dt, void* stream) {
  REQUIRE_BATCH();
  const int cur = ctx->batch_cur0 ^ (ctx->batch_enq & 1);
  return ctx->dim == 3 ? enqueue_bin_p2g<3>(ctx, (float)dt, cur, ctx->batch_enq > 0, s, ctx->cur_ev)
                       : enqueue_bin_p2g<2>(ctx, (float)dt, cur, ctx->batch_enq > 0, s, ctx->cur_ev);
}

extern "C" int mpm_phase_halo_pack(mpm_ctx* ctx, void* stream) {
  REQUIRE_BATCH();
  if (!ctx->slab.enabled) return MPM_OK;
  const int blocks = gs_blocks((int64_t)ctx->max_blocks * 32, 256, ctx->sm_count);
  for (int side = 0; side < 2; ++side) {
    if (!ctx->comm.halo[side]) continue;
    const int bx = side == 0 ? ctx->slab.lo : ctx->slab.hi;
    if (ctx->dim == 3)
      CK(launch_chain(ctx->pdl, k_halo_pack<3>, blocks, 256, 0, s, (const float4*)ctx->grid, (const uint32_t*)ctx->gb_key, ctx->L, bx,
                      ctx->comm.halo[side], ctx->comm.halo_cap, side, ctx->d_status));
    else
      CK(launch_chain(ctx->pdl, k_halo_pack<2>, blocks, 256, 0, s, (const float4*)ctx->grid, (const uint32_t*)ctx->gb_key, ctx->L, bx,
                      ctx->comm.halo[side], ctx->comm.halo_cap, side, ctx->d_status));
    ctx->launches += 1;
  }
  CK(launch_chain(ctx->pdl, k_halo_headers, 1, 1, 0, s, ctx->comm, (uint32_t)(ctx->epoch + 1), ctx->d_status));
  ctx->launches += 1;
  CK(cudaGetLastError());
  return MPM_OK;
}

extern "C" int mpm_phase_halo_add(mpm_ctx* ctx, const void* from_lo, const void* from_hi, void* stream) {
  REQUIRE_BATCH();
  if (!ctx->slab.enabled) return MPM_OK;
  int nlin = 1;
  for (int d = 0; d < ctx->dim; ++d) nlin *= ctx->L.eb[d];
  const int blocks = gs_blocks((int64_t)ctx->comm.halo_cap * 32, 256, ctx->sm_count);
  const void* from[2] = {from_lo, from_hi};
  for (int side = 0; side < 2; ++side) {
    if (!from[side]) continue;
    const int bx = side == 0 ? ctx->slab.lo : ctx->slab.hi;
    if (ctx->dim == 3)
      CK(launch_chain(ctx->pdl, k_halo_add<3>, blocks, 256, 0, s, ctx->grid, (const int*)ctx->flags, (const int*)ctx->fscan, nlin, ctx->L, bx,
                      (const uint32_t*)from[side], ctx->comm.halo_cap, (const Status*)ctx->d_status));
    else
      CK(launch_chain(ctx->pdl, k_halo_add<2>, blocks, 256, 0, s, ctx->grid, (const int*)ctx->flags, (const int*)ctx->fscan, nlin, ctx->L, bx,
                      (const uint32_t*)from[side], ctx->comm.halo_cap, (const Status*)ctx->d_status));
    ctx->launches += 1;
  }
  CK(cudaGetLastError());
  return MPM_OK;
}

extern "C" int mpm_phase_g2p(mpm_ctx* ctx, double dt, void* stream) {
  REQUIRE_BATCH();
  const int cur = ctx->batch_cur0 ^ (ctx->batch_enq & 1);
  // the key pass of the next substep of this batch rides on G2P (its layout box is fixed for the batch);
  // if no substep follows the keys are simply not used
  const bool fuse = ctx->fuse_keys && ctx->dense && !ctx->K.g2p2g;
  int rc = ctx->dim == 3 ? enqueue_grid_g2p<3>(ctx, (float)dt, cur, s, ctx->cur_ev, fuse)
                         : enqueue_grid_g2p<2>(ctx, (float)dt, cur, s, ctx->cur_ev, fuse);
  if (rc) return rc;
  ctx->batch_enq += 1;
  return MPM_OK;
}

// DistributedMPMSolver.rebalance: one round of the bulk move after mpm_set_slab changed this rank's columns (see
// k_rebalance_pack).  The buffers are LOCAL send buffers of the migration format (mpm_comm_bytes(dim, 0, cap)); the
// host exchanges them and hands the received ones to mpm_phase_unpack.  out3 = rows sent to -x, to +x, rows still outside.
// Dry run of the first substep's block discovery (keys, flags, scan): how many leaf blocks the current particles need.
template <int D>
static int probe_blocks(mpm_ctx* ctx, int32_t* need, cudaStream_t s) {
  int nlin = 1;
  for (int d = 0; d < D; ++d) nlin *= ctx->L.eb[d];
  const int n = (int)ctx->n;
  CK(cudaMemsetAsync(ctx->flags, 0, (size_t)(2 * nlin + 1) * 4, s));
  k_bin_keys<D><<<gs_blocks((n + 3) / 4, 256, ctx->sm_count), 256, 0, s>>>(ctx->state[ctx->cur], ctx->cap, ctx->K.inv_dx, ctx->L,
                                                                        ctx->slab, ctx->keys_a, ctx->flags, nlin, 0, ctx->d_status);
  CK(cudaGetLastError());
  { int rc = enqueue_scan(ctx, ctx->flags, ctx->fscan, 2 * nlin + 1, false, s); if (rc) return rc; }
  int h[2] = {0, 0};
  CK(cudaMemcpyAsync(&h[0], ctx->fscan + nlin, 4, cudaMemcpyDeviceToHost, s));
  CK(cudaMemcpyAsync(&h[1], ctx->fscan + 2 * nlin, 4, cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  *need = std::max(h[0], h[1] - h[0]);
  ctx->flags_zeroed = false;
  ctx->keys_ready = false;
  return MPM_OK;
}
extern "C" int mpm_batch_probe(mpm_ctx* ctx, int32_t* need_blocks, void* stream) {
  REQUIRE_BATCH();
  if (ctx->quant) return fail(ctx, MPM_E_INVALID, "mpm_batch_probe: f32 storage only");
  if (!need_blocks) return MPM_E_INVALID;
  if (ctx->batch_enq > 0) return fail(ctx, MPM_E_INVALID, "mpm_batch_probe: only before the first substep of a batch");
  *need_blocks = 0;
  if (ctx->n == 0) return MPM_OK;
  return ctx->dim == 3 ? probe_blocks<3>(ctx, need_blocks, s) : probe_blocks<2>(ctx, need_blocks, s);
}

extern "C" int mpm_rebalance_pack(mpm_ctx* ctx, void* send_lo_dev, void* send_hi_dev, int32_t cap, int64_t* out3,
                                  void* stream) {
  if (!ctx || !out3 || cap < 0) return MPM_E_INVALID;
  if (!ctx->state[0]) return fail(ctx, MPM_E_UNBOUND, "no buffers bound");
  if (ctx->quant || ctx->K.g2p2g) return fail(ctx, MPM_E_INVALID, "mpm_rebalance_pack: split-mode f32 storage only");
  CK(cudaSetDevice(ctx->P.device));
  cudaStream_t s = (cudaStream_t)stream;
  unsigned long long* cnt = reinterpret_cast<unsigned long long*>(ctx->stage);
  CK(cudaMemsetAsync(cnt, 0, 24, s));
  out3[0] = out3[1] = out3[2] = 0;
  if (ctx->n > 0) {
    const int blocks = gs_blocks(ctx->n, 256, ctx->sm_count), half = ctx->P.grid_size / 2;
    if (ctx->dim == 3)
      k_rebalance_pack<3><<<blocks, 256, 0, s>>>(ctx->state[ctx->cur], ctx->stat, (int)ctx->n, ctx->K.inv_dx, half, ctx->slab,
                                                 (uint32_t*)send_lo_dev, (uint32_t*)send_hi_dev, cap, cnt);
    else
      k_rebalance_pack<2><<<blocks, 256, 0, s>>>(ctx->state[ctx->cur], ctx->stat, (int)ctx->n, ctx->K.inv_dx, half, ctx->slab,
                                                 (uint32_t*)send_lo_dev, (uint32_t*)send_hi_dev, cap, cnt);
  }
  if (cap > 0) k_rebalance_headers<<<1, 1, 0, s>>>((uint32_t*)send_lo_dev, (uint32_t*)send_hi_dev, cap, cnt);
  CK(cudaGetLastError());
  unsigned long long h[3] = {0, 0, 0};
  CK(cudaMemcpyAsync(h, cnt, 24, cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  out3[0] = (int64_t)std::min<unsigned long long>(h[0], (unsigned long long)cap);
  out3[1] = (int64_t)std::min<unsigned long long>(h[1], (unsigned long long)cap);
  out3[2] = (int64_t)h[2];
  ctx->bbox_valid = false;
  return MPM_OK;
}

extern "C" int mpm_batch_end(mpm_ctx* ctx, void* stream) {
  REQUIRE_BATCH();
  ctx->in_batch = false;
  if (ctx->batch_enq > 0) {   // commit of the last substep (none in a pure delivery batch)
    k_end<<<1, 1, 0, s>>>(ctx->d_status);
    ctx->launches += 1;
  }
  CK(cudaMemcpyAsync(ctx->h_status, ctx->d_status, sizeof(Status), cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  const Status& h = *ctx->h_status;
  if (ctx->prof_enq > 0 && !h.err) {   // per-phase device time, averaged over the batch
    for (int i = 0; i < 4; ++i) ctx->ms[i] = 0.f;
    for (int k = 0; k < ctx->prof_enq; ++k)
      for (int i = 0; i < 4; ++i) {
        float t = 0.f;
        cudaEventElapsedTime(&t, ctx->ev[5 * k + i], ctx->ev[5 * k + i + 1]);
        ctx->ms[i] += t / ctx->prof_enq;
      }
  }
  ctx->cur = ctx->batch_cur0 ^ (h.done & 1);
  ctx->done_last = h.done;
  if (h.done > 0) {
    ctx->last = h;
    ctx->lastL = ctx->L;
    ctx->last_valid = (h.err == 0);
  }
  ctx->bbox_valid = false;
  if (!h.err) {
    ctx->n = h.n_cur;
    ctx->n_static = h.n_static;
    if (h.n_cur > 0 && h.done > 0) {
      for (int d = 0; d < 3; ++d) { ctx->bb_min[d] = h.bb_min[d]; ctx->bb_max[d] = h.bb_max[d]; }
      ctx->bbox_valid = true;
    }
    return MPM_OK;
  }
  ctx->layout_valid = false;
  char buf[200];
  snprintf(buf, sizeof buf, "batch stopped after %d of %d substeps: device error bits 0x%x (1 block capacity %d/%d, 2 box, "
           "4 message capacity, 8 particle capacity)", h.done, ctx->batch_enq, h.err, h.need_blocks, ctx->max_blocks);
  ctx->err = buf;
  if (h.err == ERR_BLOCK_CAPACITY) { ctx->last.need_blocks = h.need_blocks; return MPM_E_BLOCK_CAPACITY; }
  return MPM_E_INVALID;
}


// ------------------------------------------------------------------ peer path (NVLink P2P, no NCCL per substep)
// Region layout (32-bit words): [16 flag words][mig from lo][mig from hi][halo from lo][halo from hi]
//   flag 0/1: migration epoch published by the lo/hi neighbour, flag 2/3: halo epoch
static uint32_t* region_mig(mpm_ctx* c, uint32_t* base, int from_side) { return base + 16 + (size_t)from_side * c->peer_mig_words; }
static uint32_t* region_halo(mpm_ctx* c, uint32_t* base, int from_side) { return base + 16 + 2 * c->peer_mig_words + (size_t)from_side * c->peer_halo_words; }
//   ... then [3 planes from lo][3 planes from hi] (fused halo, 3D): dense (y, z) block planes of float4 node records
static float4* region_plane(mpm_ctx* c, uint32_t* base, int from_side) {
  return reinterpret_cast<float4*>(base + 16 + 2 * c->peer_mig_words + 2 * c->peer_halo_words + (size_t)from_side * 3 * c->peer_plane_words);
}

extern "C" int mpm_peer_alloc(mpm_ctx* ctx, int32_t mig_cap, int32_t halo_cap) {
  if (!ctx || mig_cap < 1 || halo_cap < 1) return MPM_E_INVALID;
  CK(cudaSetDevice(ctx->P.device));
  if (ctx->peer_region) return fail(ctx, MPM_E_INVALID, "mpm_peer_alloc: already allocated");
  ctx->peer_mig_words = mpm_comm_bytes(ctx->dim, 0, mig_cap) / 4;
  ctx->peer_halo_words = mpm_comm_bytes(ctx->dim, 1, halo_cap) / 4;
  ctx->peer_plane_words = ctx->dim == 3 ? (size_t)halo_cap * HALO_SLAB * 4 : 0;
  const size_t words = 16 + 2 * ctx->peer_mig_words + 2 * ctx->peer_halo_words + 6 * ctx->peer_plane_words;
  // the one allocation this library owns: it has to be a whole cudaMalloc block to be exported over CUDA IPC
  CK(cudaMalloc((void**)&ctx->peer_region, words * 4));
  CK(cudaMemset(ctx->peer_region, 0, words * 4));
  ctx->comm.mig_cap = mig_cap;
  ctx->comm.halo_cap = halo_cap;
  ctx->comm.plane_blocks = halo_cap;
  ctx->epoch = 0;
  return MPM_OK;
}
extern "C" int mpm_peer_handle(mpm_ctx* ctx, void* out64) {
  if (!ctx || !out64 || !ctx->peer_region) return MPM_E_INVALID;
  CK(cudaSetDevice(ctx->P.device));
  cudaIpcMemHandle_t h;
  CK(cudaIpcGetMemHandle(&h, ctx->peer_region));
  static_assert(sizeof(h) == 64, "CUDA IPC handle size");
  memcpy(out64, &h, 64);
  return MPM_OK;
}
// side 0: handle of the -x neighbour, side 1: of the +x neighbour
extern "C" int mpm_peer_open(mpm_ctx* ctx, int32_t side, const void* handle64) {
  if (!ctx || (side != 0 && side != 1) || !handle64 || !ctx->peer_region) return MPM_E_INVALID;
  CK(cudaSetDevice(ctx->P.device));
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  void* base = nullptr;
  CK(cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess));
  ctx->peer_open[side] = base;
  uint32_t* pb = (uint32_t*)base;
  // what I send to my -x neighbour arrives there as "from hi" (index 1), and vice versa
  const int there = 1 - side;
  ctx->comm.mig[side] = region_mig(ctx, pb, there);
  ctx->comm.halo[side] = region_halo(ctx, pb, there);
  ctx->comm.flag_mig[side] = pb + there;
  ctx->comm.flag_halo[side] = pb + 2 + there;
  if (ctx->peer_plane_words) {
    ctx->comm.plane_out[side] = region_plane(ctx, pb, there);            // the neighbour's planes "from" my side
    ctx->comm.plane_in[side] = region_plane(ctx, ctx->peer_region, side);  // mine, filled by that neighbour
  }
  ctx->comm.wait_mig[side] = ctx->peer_region + side;
  ctx->comm.wait_halo[side] = ctx->peer_region + 2 + side;
  return MPM_OK;
}
static void peer_close(mpm_ctx* ctx) {
  for (int s = 0; s < 2; ++s)
    if (ctx->peer_open[s]) { cudaIpcCloseMemHandle(ctx->peer_open[s]); ctx->peer_open[s] = nullptr; }
  if (ctx->peer_region) { cudaFree(ctx->peer_region); ctx->peer_region = nullptr; }
}

// `count` substeps inside an open batch; every rank calls it with the same count
extern "C" int mpm_peer_substeps(mpm_ctx* ctx, double dt, int32_t count, int32_t deliver_only, void* stream) {
  REQUIRE_BATCH();
  if (!ctx->peer_region || !ctx->slab.enabled) return fail(ctx, MPM_E_INVALID, "mpm_peer_substeps: peer path not set up");
  uint32_t* me = ctx->peer_region;
  const bool has[2] = {ctx->peer_open[0] != nullptr, ctx->peer_open[1] != nullptr};
  uint32_t* mig_from[2] = {has[0] ? region_mig(ctx, me, 0) : nullptr, has[1] ? region_mig(ctx, me, 1) : nullptr};
  uint32_t* halo_from[2] = {has[0] ? region_halo(ctx, me, 0) : nullptr, has[1] ? region_halo(ctx, me, 1) : nullptr};
  const int n_iter = deliver_only ? 1 : count;
  // Fused exchange (3D, cell-owner P2G rev. 3): the shared grid column travels inside P2G as vector reductions
  // into the neighbour's halo plane, the grid op waits for the neighbours' epoch and adds the planes, G2P's last
  // CTA publishes the migration message, the unpack kernel waits for it and commits: 10 launches per substep, no
  // one-thread kernels, and the boundary blocks go first so the transfer overlaps the interior scatter.
  const bool fused = ctx->fused_halo && ctx->dim == 3 && ctx->p2g_ver == 3 && ctx->p2g_variant == 1 && ctx->dense &&
                     ctx->peer_plane_words && (int64_t)ctx->L.eb[1] * ctx->L.eb[2] <= ctx->comm.plane_blocks;
  if (fused) {
    // per-phase CUDA events (mpm_set_profiling): "sort" then includes the wait for the neighbours' migration
    // message, "grid" the wait for their halo
    const bool prof = ctx->profiling && !deliver_only && ctx->prof_enq + n_iter <= 4096;
    if (prof)
      while ((int)ctx->ev.size() < 5 * (ctx->prof_enq + n_iter)) {
        cudaEvent_t e;
        CK(cudaEventCreate(&e));
        ctx->ev.push_back(e);
      }
    for (int i = 0; i < n_iter; ++i) {
      ctx->comm.fused = 1;
      ctx->comm.epoch = ctx->epoch;
      ctx->cur_ev = prof ? ctx->ev.data() + 5 * (ctx->prof_enq++) : nullptr;
      if (ctx->cur_ev) cudaEventRecord(ctx->cur_ev[0], s);      // (re-recorded by the binning: the unpack wait counts as "sort")
      int rc = mpm_phase_unpack(ctx, mig_from[0], mig_from[1], stream);
      if (!rc && !deliver_only) rc = mpm_phase_p2g(ctx, dt, stream);
      if (!rc && !deliver_only) rc = mpm_phase_g2p(ctx, dt, stream);
      ctx->comm.fused = 0;
      ctx->cur_ev = nullptr;
      if (rc) return rc;
      if (!deliver_only) ctx->epoch += 1;
    }
    CK(cudaGetLastError());
    return MPM_OK;
  }
  for (int i = 0; i < n_iter; ++i) {
    // leavers of the neighbours' last G2P (epoch = substeps completed so far)
    CK(launch_chain(ctx->pdl, k_wait_flags, 1, 1, 0, s, (const uint32_t*)(has[0] ? me + 0 : nullptr),
                    (const uint32_t*)(has[1] ? me + 1 : nullptr), (uint32_t)ctx->epoch, ctx->d_status));
    ctx->launches += 1;
    int rc = mpm_phase_unpack(ctx, mig_from[0], mig_from[1], stream);
    if (rc) return rc;
    if (deliver_only) break;
    rc = mpm_phase_p2g(ctx, dt, stream);
    if (rc) return rc;
    rc = mpm_phase_halo_pack(ctx, stream);          // records + epoch go straight to the neighbours
    if (rc) return rc;
    CK(launch_chain(ctx->pdl, k_wait_flags, 1, 1, 0, s, (const uint32_t*)(has[0] ? me + 2 : nullptr),
                    (const uint32_t*)(has[1] ? me + 3 : nullptr), (uint32_t)(ctx->epoch + 1), ctx->d_status));
    ctx->launches += 1;
    rc = mpm_phase_halo_add(ctx, halo_from[0], halo_from[1], stream);
    if (rc) return rc;
    rc = mpm_phase_g2p(ctx, dt, stream);            // leavers + epoch go straight to the neighbours
    if (rc) return rc;
    ctx->epoch += 1;
  }
  CK(cudaGetLastError());
  return MPM_OK;
}

// local rows [0, n) of one state word in storage order (no id un-permutation)
extern "C" int mpm_download_raw(mpm_ctx* ctx, int32_t field, void* dst_host, void* stream) {
  if (!ctx || field < 0 || field >= ctx->nv) return MPM_E_INVALID;
  if (ctx->n == 0) return MPM_OK;
  if (!dst_host) return MPM_E_INVALID;
  CK(cudaSetDevice(ctx->P.device));
  cudaStream_t s = (cudaStream_t)stream;
  if (ctx->dim == 3) k_gather_raw<3><<<gs_blocks(ctx->n, 256, ctx->sm_count), 256, 0, s>>>(ctx->state[ctx->cur], ctx->stat, field, (int)ctx->n, ctx->stage, ctx->quant);
  else k_gather_raw<2><<<gs_blocks(ctx->n, 256, ctx->sm_count), 256, 0, s>>>(ctx->state[ctx->cur], ctx->stat, field, (int)ctx->n, ctx->stage, 0);
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(dst_host, ctx->stage, (size_t)ctx->n * 4, cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  return MPM_OK;
}

extern "C" int mpm_get_stats(mpm_ctx* ctx, mpm_stats* o) {
  if (!ctx || !o) return MPM_E_INVALID;
  memset(o, 0, sizeof(*o));
  o->n_particles = ctx->n;
  o->n_particle_blocks = ctx->last.npb;
  o->n_grid_blocks = std::max(ctx->last.ngb, ctx->last.need_blocks);
  o->max_blocks = ctx->max_blocks;
  o->key_bits = ctx->L.key_bits;
  for (int d = 0; d < 3; ++d) { o->bbox_min[d] = ctx->bb_min[d]; o->bbox_max[d] = ctx->bb_max[d]; }
  uint32_t bits = ctx->last.maxv_all;
  memcpy(&o->max_velocity, &bits, 4);
  o->launches = ctx->launches;
  o->substeps_done = ctx->done_last;
  bits = ctx->last.maxgv_bits;
  memcpy(&o->max_grid_velocity, &bits, 4);
  o->ms_sort = ctx->ms[0]; o->ms_p2g = ctx->ms[1]; o->ms_grid = ctx->ms[2]; o->ms_g2p = ctx->ms[3];
  return MPM_OK;
}
extern "C" int mpm_set_profiling(mpm_ctx* ctx, int32_t enabled) {
  if (!ctx) return MPM_E_INVALID;
  ctx->profiling = enabled != 0;
  return MPM_OK;
}

// ------------------------------------------------------------------ read-back
extern "C" int mpm_gather_rows(mpm_ctx* ctx, int32_t first_field, int32_t nwords, int64_t begin, int64_t end,
                               void* dst_dev, void* stream);
extern "C" int mpm_gather(mpm_ctx* ctx, int32_t field, int64_t begin, int64_t end, void* dst_dev, void* stream) {
  return mpm_gather_rows(ctx, field, 1, begin, end, dst_dev, stream);
}

extern "C" int mpm_gather_rows(mpm_ctx* ctx, int32_t first_field, int32_t nwords, int64_t begin, int64_t end,
                               void* dst_dev, void* stream) {
  if (!ctx || first_field < 0 || nwords < 1 || first_field + nwords > ctx->nv || begin < 0 || end < begin || end > ctx->n)
    return fail(ctx, MPM_E_INVALID, "mpm_gather_rows: bad range/field");
  if (end == begin) return MPM_OK;
  if (!dst_dev) return MPM_E_INVALID;
  CK(cudaSetDevice(ctx->P.device));
  // rows whose id is not present (ids that are not a permutation of [0, n): the distributed solver's
  // global ids) read as zero instead of uninitialised memory
  CK(cudaMemsetAsync(dst_dev, 0, (size_t)(end - begin) * nwords * 4, (cudaStream_t)stream));
  if (ctx->dim == 3)
    k_gather_rows<3><<<gs_blocks(ctx->n, 256, ctx->sm_count), 256, 0, (cudaStream_t)stream>>>(
        ctx->state[ctx->cur], ctx->stat, first_field, nwords, (int)ctx->n, begin, end, (uint32_t*)dst_dev, ctx->quant);
  else
    k_gather_rows<2><<<gs_blocks(ctx->n, 256, ctx->sm_count), 256, 0, (cudaStream_t)stream>>>(
        ctx->state[ctx->cur], ctx->stat, first_field, nwords, (int)ctx->n, begin, end, (uint32_t*)dst_dev, 0);
  CK(cudaGetLastError());
  return MPM_OK;
}

// ParticleIO.write_particles on the device (engine/particle_io.py:42-76)
extern "C" int mpm_particle_ranges(mpm_ctx* ctx, float* ranges_dev, void* stream) {
  if (!ctx || !ranges_dev) return MPM_E_INVALID;
  if (ctx->n <= 0) return fail(ctx, MPM_E_INVALID, "mpm_particle_ranges: no particles");
  CK(cudaSetDevice(ctx->P.device));
  cudaStream_t s = (cudaStream_t)stream;
  uint32_t* r = reinterpret_cast<uint32_t*>(ranges_dev);
  const int nw = 4 * ctx->dim;
  k_ranges_init<<<1, 32, 0, s>>>(r, nw);
  const int blocks = gs_blocks(ctx->n, 256, ctx->sm_count);
  if (ctx->dim == 3) k_ranges<3><<<blocks, 256, 0, s>>>(ctx->state[ctx->cur], ctx->cap, (int)ctx->n, r, ctx->quant);
  else k_ranges<2><<<blocks, 256, 0, s>>>(ctx->state[ctx->cur], ctx->cap, (int)ctx->n, r, 0);
  k_ranges_decode<<<1, 32, 0, s>>>(r, nw);
  CK(cudaGetLastError());
  return MPM_OK;
}
extern "C" int mpm_pack_particles(mpm_ctx* ctx, const float* lo_inv_host, uint32_t* x_and_v_dev, uint8_t* color_dev,
                                  void* stream) {
  if (!ctx || !lo_inv_host) return MPM_E_INVALID;
  if (ctx->n <= 0) return MPM_OK;
  if (!x_and_v_dev || !color_dev) return MPM_E_INVALID;
  CK(cudaSetDevice(ctx->P.device));
  PackArgs pa{};
  for (int c = 0; c < 2; ++c)
    for (int d = 0; d < ctx->dim; ++d) {
      pa.lo[c][d] = lo_inv_host[(c * ctx->dim + d) * 2];
      pa.inv[c][d] = lo_inv_host[(c * ctx->dim + d) * 2 + 1];
    }
  cudaStream_t s = (cudaStream_t)stream;
  const int blocks = gs_blocks(ctx->n, 256, ctx->sm_count);
  if (ctx->dim == 3)
    k_pack_particles<3><<<blocks, 256, 0, s>>>(ctx->state[ctx->cur], ctx->stat, (int)ctx->n, pa, x_and_v_dev, color_dev, ctx->quant);
  else
    k_pack_particles<2><<<blocks, 256, 0, s>>>(ctx->state[ctx->cur], ctx->stat, (int)ctx->n, pa, x_and_v_dev, color_dev, 0);
  CK(cudaGetLastError());
  return MPM_OK;
}

extern "C" int mpm_download(mpm_ctx* ctx, int32_t field, int64_t begin, int64_t end, void* dst_host, void* stream) {
  if (!ctx) return MPM_E_INVALID;
  int rc = mpm_gather(ctx, field, begin, end, ctx->stage, stream);
  if (rc || end == begin) return rc;
  if (!dst_host) return MPM_E_INVALID;
  cudaStream_t s = (cudaStream_t)stream;
  CK(cudaMemcpyAsync(dst_host, ctx->stage, (size_t)(end - begin) * 4, cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  return MPM_OK;
}


// ------------------------------------------------------------------ mesh seeding
extern "C" int mpm_voxelize(int32_t device, const double* tris_dev, int64_t ntri, const int32_t* res, double dx,
                            int32_t padding, const int32_t* lo, const int32_t* hi, int32_t* vox, void* stream) {
  if (!res || !lo || !hi || ntri < 0 || (ntri > 0 && (!tris_dev || !vox))) return MPM_E_INVALID;
  if (ntri == 0) return MPM_OK;
  if (cudaSetDevice(device) != cudaSuccess) return MPM_E_CUDA;
  VoxArgs a{};
  a.tris = tris_dev; a.ntri = ntri; a.dx = dx; a.inv_dx = 1.0 / dx; a.padding = padding; a.vox = vox;
  for (int d = 0; d < 3; ++d) { a.res[d] = res[d]; a.lo[d] = lo[d]; a.hi[d] = hi[d]; }
  int blocks = (int)std::min<int64_t>((ntri * 32 + 255) / 256, 148 * 16);
  k_voxelize<<<std::max(blocks, 1), 256, 0, (cudaStream_t)stream>>>(a);
  return cudaGetLastError() == cudaSuccess ? MPM_OK : MPM_E_CUDA;
}

extern "C" int mpm_voxel_sample(int32_t device, const int32_t* vox, const int32_t* res, const int32_t* lo,
                                const int32_t* hi, int32_t sample_density, int32_t super_sample, double cell,
                                const double* translation, int32_t grid_size, int32_t padding, uint64_t seed,
                                int32_t pass, int32_t* counts, const int64_t* offsets, float* x_out, void* stream) {
  if (!vox || !res || !lo || !hi || sample_density < 0 || super_sample < 1) return MPM_E_INVALID;
  if (pass == 0 ? !counts : (!offsets || !x_out)) return MPM_E_INVALID;
  if (cudaSetDevice(device) != cudaSuccess) return MPM_E_CUDA;
  VoxSampleArgs a{};
  a.vox = vox; a.sample_density = sample_density; a.super_sample = super_sample;
  a.s = (float)((double)sample_density / ((double)super_sample * super_sample * super_sample));
  a.cell = (float)cell;
  a.grid_size = grid_size; a.padding = padding; a.seed = seed;
  a.counts = counts; a.offsets = offsets; a.x_out = x_out; a.pass = pass;
  size_t total = 1;
  for (int d = 0; d < 3; ++d) {
    a.res[d] = res[d]; a.lo[d] = lo[d]; a.hi[d] = hi[d];
    a.trans[d] = translation ? (float)translation[d] : 0.f;
    if (hi[d] <= lo[d]) return MPM_OK;
    total *= (size_t)(hi[d] - lo[d]);
  }
  int blocks = (int)std::min<size_t>((total + 255) / 256, 148 * 16);
  k_voxel_sample<<<std::max(blocks, 1), 256, 0, (cudaStream_t)stream>>>(a);
  return cudaGetLastError() == cudaSuccess ? MPM_OK : MPM_E_CUDA;
}

// ------------------------------------------------------------------ debug
// the binning's single-launch exclusive scan on caller data (parity tests): out[i] = sum(in[0..i))
extern "C" int mpm_debug_scan(mpm_ctx* ctx, const int32_t* in_dev, int32_t* out_dev, int64_t n, void* stream) {
  if (!ctx || !in_dev || !out_dev || n < 1) return MPM_E_INVALID;
  if (!ctx->scan_desc) return fail(ctx, MPM_E_UNBOUND, "no buffers bound");
  if ((n + SCAN_TILE - 1) / SCAN_TILE + 1 > (int64_t)ctx->scan_tiles) return fail(ctx, MPM_E_INVALID, "mpm_debug_scan: n exceeds the bound workspace");
  CK(cudaSetDevice(ctx->P.device));
  cudaStream_t s = (cudaStream_t)stream;
  ctx->scan_epoch = (ctx->scan_epoch + 1) & 0x3fffffffu;
  if (ctx->scan_epoch == 0) ctx->scan_epoch = 1;
  const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(ctx->scan_grid, (n + SCAN_TILE - 1) / SCAN_TILE));
  CK(launch_chain(ctx->pdl, k_scan_excl<false>, grid, SCAN_T, 0, s, (const int*)in_dev, (int*)out_dev, (int)n,
                  ctx->scan_desc, ctx->scan_epoch, ctx->d_status, (const int*)nullptr, 0));
  return MPM_OK;
}

extern "C" int mpm_debug_binning(mpm_ctx* ctx, int32_t* block_host, void* stream) {
  if (!ctx || !block_host) return MPM_E_INVALID;
  if (ctx->n == 0) return MPM_OK;
  CK(cudaSetDevice(ctx->P.device));
  cudaStream_t s = (cudaStream_t)stream;
  int* out = (int*)ctx->state[ctx->cur ^ 1];   // the idle set is free between substeps
  const int half = ctx->P.grid_size / 2;
  int blocks = gs_blocks(ctx->n, 256, ctx->sm_count);
  if (ctx->dim == 3) k_debug_binning<3><<<blocks, 256, 0, s>>>(ctx->state[ctx->cur], ctx->stat, (int)ctx->n, ctx->K.inv_dx, half, out, ctx->quant);
  else k_debug_binning<2><<<blocks, 256, 0, s>>>(ctx->state[ctx->cur], ctx->stat, (int)ctx->n, ctx->K.inv_dx, half, out, 0);
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(block_host, out, (size_t)ctx->n * ctx->dim * 4, cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  return MPM_OK;
}

static void decode_block(const mpm_ctx* ctx, const KeyLayout& L, uint32_t lin, int* out) {
  for (int d = ctx->dim - 1; d >= 0; --d) {
    out[d] = (int)(lin % (uint32_t)L.eb[d]) + L.ob[d];
    lin /= (uint32_t)L.eb[d];
  }
}

extern "C" int mpm_debug_blocks(mpm_ctx* ctx, int32_t* pb_coords, int32_t* pb_counts, int32_t* npb_out,
                                int32_t* gb_coords, int32_t* ngb_out) {
  if (!ctx) return MPM_E_INVALID;
  if (!ctx->last_valid) return fail(ctx, MPM_E_INVALID, "no completed substep to inspect");
  CK(cudaSetDevice(ctx->P.device));
  const int npb = ctx->last.npb, ngb = ctx->last.ngb;
  if (npb_out) *npb_out = npb;
  if (ngb_out) *ngb_out = ngb;
  CK(cudaDeviceSynchronize());
  if (pb_coords || pb_counts) {
    std::vector<int> start(npb + 1);
    CK(cudaMemcpy(start.data(), ctx->pb_start, (size_t)(npb + 1) * 4, cudaMemcpyDeviceToHost));
    for (int b = 0; b < npb; ++b)
      if (pb_counts) pb_counts[b] = start[b + 1] - start[b];
    if (pb_coords) {
      std::vector<uint32_t> pk(npb);
      CK(cudaMemcpy(pk.data(), ctx->pb_key, (size_t)npb * 4, cudaMemcpyDeviceToHost));
      for (int b = 0; b < npb; ++b) decode_block(ctx, ctx->lastL, pk[b], pb_coords + (size_t)b * ctx->dim);
    }
  }
  if (gb_coords) {
    std::vector<uint32_t> k(ngb);
    CK(cudaMemcpy(k.data(), ctx->gb_key, (size_t)ngb * 4, cudaMemcpyDeviceToHost));
    for (int g = 0; g < ngb; ++g) decode_block(ctx, ctx->lastL, k[g], gb_coords + (size_t)g * ctx->dim);
  }
  return MPM_OK;
}

extern "C" int mpm_debug_grid(mpm_ctx* ctx, int32_t* cell_host, float* vm_host, int64_t max_cells, int64_t* ncell) {
  if (!ctx) return MPM_E_INVALID;
  if (!ctx->last_valid) return fail(ctx, MPM_E_INVALID, "no completed substep to inspect");
  CK(cudaSetDevice(ctx->P.device));
  const int ngb = ctx->last.ngb;
  const int64_t total = (int64_t)ngb * ctx->cells;
  if (ncell) *ncell = total;
  if (!cell_host || !vm_host) return MPM_OK;
  if (max_cells < total) return fail(ctx, MPM_E_INVALID, "mpm_debug_grid: buffer too small");
  CK(cudaDeviceSynchronize());
  std::vector<uint32_t> k(ngb);
  CK(cudaMemcpy(k.data(), ctx->gb_key, (size_t)ngb * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(vm_host, ctx->grid, (size_t)total * 16, cudaMemcpyDeviceToHost));
  const int half = ctx->P.grid_size / 2, ll = ctx->log_leaf, leaf = 1 << ll;
  for (int g = 0; g < ngb; ++g) {
    int blk[3] = {0, 0, 0};
    decode_block(ctx, ctx->lastL, k[g], blk);
    for (int c = 0; c < ctx->cells; ++c) {
      int32_t* o = cell_host + ((size_t)g * ctx->cells + c) * ctx->dim;
      for (int d = 0; d < ctx->dim; ++d) {
        int lc = (c >> (ll * (ctx->dim - 1 - d))) & (leaf - 1);
        o[d] = (blk[d] << ll) + lc - half;
      }
    }
  }
  return MPM_OK;
}

extern "C" int mpm_debug_particle_update(mpm_ctx* ctx, double dt, int64_t n, const int32_t* mat_h, float* F_h,
                                         const float* C_h, float* Jp_h, float* aff_h, float* mass_h) {
  if (!ctx || n < 0) return MPM_E_INVALID;
  if (n == 0) return MPM_OK;
  CK(cudaSetDevice(ctx->P.device));
  const size_t dd = (size_t)ctx->dim * ctx->dim;
  // debug-only path: uses the idle state set as scratch
  const size_t need = (size_t)n * (3 * dd + 3);
  if (need > (size_t)ctx->nf * ctx->cap) return fail(ctx, MPM_E_INVALID, "mpm_debug_particle_update: n too large for scratch");
  float* base = (float*)ctx->state[ctx->cur ^ 1];
  float *F = base, *C = F + n * dd, *A = C + n * dd, *Jp = A + n * dd, *M = Jp + n;
  int* mat = (int*)(M + n);
  CK(cudaMemcpy(F, F_h, n * dd * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(C, C_h, n * dd * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(Jp, Jp_h, n * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(mat, mat_h, n * 4, cudaMemcpyHostToDevice));
  int blocks = gs_blocks(n, 128, ctx->sm_count);
  if (ctx->dim == 3) k_debug_update<3><<<blocks, 128>>>(ctx->K, (float)dt, (int)n, mat, F, C, Jp, A, M);
  else k_debug_update<2><<<blocks, 128>>>(ctx->K, (float)dt, (int)n, mat, F, C, Jp, A, M);
  CK(cudaGetLastError());
  CK(cudaMemcpy(F_h, F, n * dd * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(aff_h, A, n * dd * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(Jp_h, Jp, n * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(mass_h, M, n * 4, cudaMemcpyDeviceToHost));
  return MPM_OK;
}
