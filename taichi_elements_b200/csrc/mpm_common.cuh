// Shared declarations of libmpm_b200: geometry traits, key layout, the device
// status block and the workspace carve-up.  See include/mpm_b200.h for the ABI
// and DESIGN.md for the data layout.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "mpm_math.cuh"

namespace mpm {

// Leaf-block geometry of the reference's sparse grid
// (engine/mpm_solver.py:154-186): 4^3 leaves in 3D, 16^2 in 2D.
template <int D> struct Geo;
template <> struct Geo<3> {
  static constexpr int LEAF = 4, LOG_LEAF = 2, CELLS = 64, CB = 6;   // CB = cell bits in a key
  static constexpr int T = LEAF + 2, TN = T * T * T, NO = 8;         // staged tile, octants
  static constexpr int NF = 2 * 3 + 2 * 9 + 2;                       // physical words per particle (see Fld)
};
template <> struct Geo<2> {
  static constexpr int LEAF = 16, LOG_LEAF = 4, CELLS = 256, CB = 8;
  static constexpr int T = LEAF + 2, TN = T * T, NO = 4;
  static constexpr int NF = 2 * 2 + 2 * 4 + 2;
};

// Word indices of one particle inside its tile (mpm_kernels.cuh: word()).  The substep moves every particle to its
// sorted slot each time, so only what the physics needs travels: x v F C Jp and ONE tag word,
//   tag = material << 29 | sid,
// where sid indexes the static side arrays (colour, id, emitter), which never move.  MAT .. EMIT are the VIRTUAL
// word numbers of the read-back ABI (include/mpm_b200.h: the reference's field order), resolved by vword().
template <int D> struct Fld {
  static constexpr int X = 0, V = D, F = 2 * D, C = 2 * D + D * D, JP = 2 * D + 2 * D * D, TAG = JP + 1, N = JP + 2;
  static constexpr int MAT = JP + 1, COLOR = JP + 2, ID = JP + 3, EMIT = JP + 4, NV = JP + 5;
};
static constexpr int TAG_SHIFT = 29;
static constexpr uint32_t TAG_SID = (1u << TAG_SHIFT) - 1u;
__host__ __device__ inline uint32_t make_tag(uint32_t material, uint32_t sid) { return (material << TAG_SHIFT) | (sid & TAG_SID); }
__host__ __device__ inline uint32_t tag_mat(uint32_t tag) { return tag >> TAG_SHIFT; }
__host__ __device__ inline uint32_t tag_sid(uint32_t tag) { return tag & TAG_SID; }
// Material code of a row that was handed to another rank while the cut planes were moved (mpm_rebalance_pack): the
// binning and the local exports skip it whatever its position says; the next sort drops the row.
static constexpr uint32_t MAT_DEAD = 7u;
// Static per-particle attributes, indexed by sid (rows are appended, never moved by the substep; the distributed
// solver compacts them between batches): packed colour, id (insertion index; global id with slabs), emitter id.
struct Statics {
  uint32_t* color;
  uint32_t* gid;
  uint32_t* emit;
};

static constexpr uint32_t INVALID_KEY = 0xFFFFFFFFu;

// Sort-key layout for one substep: leaf-block coordinates relative to a
// host-maintained ("sticky") bounding box, row-major with x slowest, then the
// cell index inside the leaf.
struct KeyLayout {
  int ob[3];   // origin of the box, absolute leaf-block coords (0 .. grid_size/leaf)
  int eb[3];   // extent in blocks (includes the +1 upper neighbour ring)
  int half;    // grid_size/2: global signed cell index = absolute cell - half
  int key_bits;
};

enum ErrBits { ERR_BLOCK_CAPACITY = 1, ERR_BBOX = 2, ERR_COMM_CAPACITY = 4, ERR_PARTICLE_CAPACITY = 8, ERR_COMM_TIMEOUT = 16, ERR_SCAN_TIMEOUT = 32 };

// Device-resident status block, copied to pinned host memory after each batch.
struct Status {
  int npb;          // particle blocks of this substep
  int ngb_raw;      // unique candidates incl. possibly INVALID_KEY
  int ngb;          // grid (active) blocks of this substep
  int err;          // ErrBits, sticky within a batch
  int done;         // substeps completed in this batch
  int work_p2g, work_g2p;
  unsigned maxv_bits;
  int bb_min[3], bb_max[3];
  int need_blocks;  // max(npb, ngb) seen, for capacity growth
  unsigned maxv_all; // max of maxv_bits over the batch
  int n_cur;        // rows of the live set in use (includes particles that left this rank's slab)
  int n_live;       // particles binned this substep (= rows of the other set after G2P)
  int mig_cnt[2];   // particles leaving to the -x / +x neighbour rank (filled by G2P)
  int halo_cnt[2];  // packed boundary-column grid blocks (-x / +x side)
  unsigned maxgv_bits;  // max |grid v|_inf after the grid op (compute_max_grid_velocity)
  int half;         // g2p2g: gather halves completed in this batch
  int next_err;     // error bits raised for the NEXT substep by G2P's fused key pass
  // multi-GPU, fused exchange (mpm_comm.cuh): particle blocks of the first / last block column of the slab
  // (their tiles touch a column shared with a neighbour; P2G takes them first), and "last CTA" counters
  int bnd_lo, bnd_hi;
  int halo_done;    // boundary blocks whose shared-column nodes have been sent (P2G)
  int g2p_done;     // CTAs of G2P that have finished (the last one publishes the migration message)
  int unpack_done;  // CTAs of k_mig_unpack that have finished (the last one commits the appended rows)
  int n_static;     // rows of the static side arrays in use (arrivals append; see Statics)
};

// Slab decomposition along x (multi-GPU): this rank owns leaf-block columns
// [lo, hi) in absolute block coordinates.  Disabled = owns everything.
struct Slab {
  int enabled, lo, hi;
};
// Fixed-capacity message buffers (32-bit words, 16-word header, word 0 = count).
//   migration: header | field-major rows  [field][mig_cap]
//   halo:      header | keys [halo_cap] | node records [halo_cap][CELLS] float4
static constexpr int COMM_HEADER = 16;
static constexpr int HALO_SLAB = 2 * 6 * 6;   // node records one boundary block sends (3D)
struct CommBufs {
  uint32_t* mig[2];        // where leavers to the -x / +x rank are written: a local send buffer (NCCL
  uint32_t* halo[2];       //   path) or the neighbour's receive buffer mapped over NVLink (peer path)
  int mig_cap, halo_cap;
  uint32_t* flag_mig[2];   // peer path: the neighbour's "data ready" epoch words to release-store
  uint32_t* flag_halo[2];
  // fused halo (peer path, 3D): the nodes of the grid column at a cut that both ranks scatter into are its first TWO
  // cell layers.  Every boundary particle block stores its partial sums of them -- a slab of 2 x 6 x 6 node records
  // of its tile -- with plain vector stores into a slot of a dense plane in the NEIGHBOUR's memory, indexed by the
  // block's (y, z) coordinates in the common key layout (no remote atomics); the neighbour's grid op adds the <= 4
  // slabs that overlap a node.  Three planes per side rotate with the substep epoch (filled / read / cleared).
  float4* plane_out[2];    // neighbour's planes for what I send to the -x / +x side (3 * plane_blocks * HALO_SLAB)
  float4* plane_in[2];     // my planes, filled by the -x / +x neighbour
  const uint32_t* wait_mig[2];    // my epoch words, written by the neighbours
  const uint32_t* wait_halo[2];
  int plane_blocks;        // capacity of one plane in leaf blocks (>= eb[1] * eb[2] of the key layout)
  int fused;               // 1: the fused exchange is active for this substep
  uint32_t epoch;          // substeps completed before this one
};

struct Grav { float g[3]; };

struct ColliderDev {
  int kind, surface;
  float a[3], b[3];
  float r2;        // sphere: radius*radius rounded once to f32
  float friction;
  int unbounded;
};
static constexpr int MAX_COLLIDERS = 64;
struct ColliderTable {
  int n;
  ColliderDev c[MAX_COLLIDERS];
};

struct GridCfg {
  int res[3];
  int padding;
  int grid_size;
};

__host__ __device__ inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

}  // namespace mpm
