/* mpm_b200.h -- C ABI of libmpm_b200.so, the sm_100a MLS-MPM substep engine
 * that sits behind taichi_elements' Python `MPMSolver` class.
 *
 * The reference (taichi-dev/taichi_elements) has no FFI of its own: its hot
 * path is a set of @ti.kernel methods JIT-compiled by Taichi.  Each entry point
 * below names the reference kernel/method (file:line under /root/reference) it
 * replaces; INTEGRATION.md shows the ctypes binding a maintainer would add to
 * engine/mpm_solver.py.
 *
 * Conventions
 *   - plain C: pointers, sizes, ints, doubles; no torch/C++ types cross the ABI.
 *   - every function returns int: 0 = MPM_OK, >0 = a recoverable condition the
 *     host must act on (grow a buffer and call again), <0 = failure;
 *     mpm_last_error(ctx) gives the text.  No exception crosses the ABI.
 *   - the library never allocates or frees device memory: the host (PyTorch)
 *     owns the particle state buffers and one workspace buffer and binds them
 *     with mpm_bind().  Pointers named *_dev are device pointers, *_host host.
 *   - one CUDA stream per call (cudaStream_t passed as void*, NULL = default);
 *     a ctx is not re-entrant; any host thread may call (the device is set per
 *     call), matching Blender's use of MPMSolver from a worker thread
 *     (blender/operators.py:403-405).
 *
 * Particle state layout (device, 32-bit words): two sets (ping-pong), each an array of TILES of 32 particles;
 * inside a tile one 128-byte row of 32 lanes per state word:
 *     state[set][capacity / 32][nf][32],   word(f, p) = ((p / 32) * nf + f) * 32 + p % 32
 *   nf = mpm_state_fields(dim) words in order  x[dim] v[dim] F[dim*dim] C[dim*dim] Jp tag
 *   (engine/mpm_solver.py:101-136, 263-273).  A warp reading one word of 32 consecutive particles touches one
 *   128-byte line (as in a structure of arrays), and every word of a particle sits at a constant offset from one
 *   address.  Particles are physically kept in grid-block order (every substep writes them to their sorted slot in
 *   the other set), so only what the physics needs travels: tag = material << 29 | sid, where sid is the row of the
 *   particle in the STATIC side arrays  statics[3][capacity] = colour, id, emitter  (:117, 129-136), which never
 *   move.  `id` is the insertion index and restores the reference's append order on read-back.
 * Read-back calls number the words as the reference orders its fields (mpm_virtual_fields(dim) of them):
 *     x[dim] v[dim] F[dim*dim] C[dim*dim] Jp material color id emitter.
 */
#ifndef MPM_B200_H
#define MPM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MPM_OK 0
#define MPM_E_BLOCK_CAPACITY 1 /* workspace too small for the active blocks: query mpm_get_stats, re-bind, retry */
#define MPM_E_KEY_BITS 2       /* particle bounding box needs > 32 key bits */
#define MPM_E_INVALID (-1)
#define MPM_E_CUDA (-2)
#define MPM_E_UNBOUND (-3)

#define MPM_ABI_VERSION 4

typedef struct mpm_ctx mpm_ctx;

/* Constants of MPMSolver.__init__ (engine/mpm_solver.py:67-99, 143-159, 198-210).
 * Doubles are the Python floats; the library rounds them to f32 exactly where
 * Taichi would (kernel-captured constants). */
typedef struct mpm_params {
  int32_t dim;             /* 2 or 3 (:76-77) */
  int32_t res[3];
  int32_t grid_size;       /* 4096, doubled while <= 2*max(res) when unbounded (:74,143-148) */
  int32_t leaf;            /* leaf_block_size: 16 (2D) / 4 (3D) (:155-159) */
  int32_t padding;         /* :52 */
  int32_t support_plasticity;
  int32_t device;          /* CUDA device ordinal */
  int32_t flags;           /* bit 0: use_g2p2g (:57), bit 1: quant (F clamp in g2p2g, :99, 415-416; with bit 0 in 3D also
                            * the bit-packed storage, see mpm_ctx_state_fields) */
  double dx, inv_dx;       /* :82-83 */
  double p_vol, p_mass;    /* :85-87 */
  double mu_0, lambda_0;   /* :203-205 */
  double alpha;            /* :208-210 */
  double water_density;    /* :61 */
  double g2p2g_cfl;        /* g2p2g_allowed_cfl if v_clamp_g2p2g else 0: grid-velocity clamp (:589, 596-598) */
} mpm_params;

/* One entry of MPMSolver.grid_postprocess, applied in order after
 * normalisation+gravity (:794-795). */
#define MPM_COLLIDER_BBOX 0   /* grid_bounding_box (:600-616); a[0] = unbounded flag */
#define MPM_COLLIDER_SPHERE 1 /* add_sphere_collider (:618-642): a = center, b[0] = radius */
#define MPM_COLLIDER_PLANE 2  /* add_surface_collider (:647-687): a = point, b = unit normal */
typedef struct mpm_collider {
  int32_t kind;
  int32_t surface; /* 0 sticky, 1 slip, 2 separate (:29-42) */
  double a[3];
  double b[3];
  double friction;
} mpm_collider;

/* Read-back of the per-substep device status block. */
typedef struct mpm_stats {
  int64_t n_particles;
  int32_t n_particle_blocks; /* leaf blocks holding >= 1 particle base (build_pid lists, :344-361) */
  int32_t n_grid_blocks;     /* leaf blocks activated by P2G writes (:582-584) */
  int32_t max_blocks;        /* bound capacity */
  int32_t key_bits;
  int32_t bbox_min[3], bbox_max[3]; /* particle base-cell bounding box (global signed cell index) */
  float max_velocity;        /* compute_max_velocity (:726-735): max over the substeps of the last call */
  int32_t launches;          /* kernels launched by the last mpm_substep(s) call */
  int32_t substeps_done;     /* substeps completed by the last mpm_substep(s) call */
  float max_grid_velocity;   /* compute_max_grid_velocity (:737-746) of the last substep's grid */
  float ms_sort, ms_p2g, ms_grid, ms_g2p; /* filled when profiling is enabled */
} mpm_stats;

int mpm_abi_version(void);

/* physical words per particle in a state set (2*dim + 2*dim*dim + 2) */
int mpm_state_fields(int dim);
/* words of the read-back numbering (2*dim + 2*dim*dim + 5) */
int mpm_virtual_fields(int dim);
/* physical words per particle of THIS context: mpm_state_fields(dim), or with the bit-packed storage of quant=True
 * (flags bit 1, dim 3; engine/mpm_solver.py:106-114, 216-247):
 *   with use_g2p2g (flags bit 0): 11 words  xq[2] vq[2] Fq[5] Jp tag          (C does not exist in that mode)
 *   split substep:                20 words  xq[2] vq[2] Fq[5] Jp tag C[9]     (C stays f32, as in the reference)
 *   x  3 x 21-bit signed fixed point, range +-2.0; v  3 x 19-bit fractions sharing a 7-bit exponent;
 *   F  9 x 16-bit signed fixed point, range +-4.1 (csrc/mpm_quant.cuh). */
int mpm_ctx_state_fields(mpm_ctx* ctx);
/* bytes the workspace must have for `capacity` particles and `max_blocks` leaf blocks */
size_t mpm_workspace_bytes(int dim, int64_t capacity, int32_t max_blocks);

/* MPMSolver.__init__ (engine/mpm_solver.py:44-310) */
int mpm_create(const mpm_params* params, mpm_ctx** out);
int mpm_destroy(mpm_ctx* ctx);
const char* mpm_last_error(mpm_ctx* ctx);

/* Bind host-owned device memory.  state0/state1: [capacity / 32][nf][32] words each (capacity % 64 == 0,
 * capacity <= 2^29); statics_dev: [3][capacity] words. */
int mpm_bind(mpm_ctx* ctx, void* state0_dev, void* state1_dev, void* statics_dev, int64_t capacity,
             void* workspace_dev, size_t workspace_bytes, int32_t max_blocks);
/* rows of the static side arrays in use (== n_particles on a single device; with slabs departures leave holes) */
int mpm_get_static_rows(mpm_ctx* ctx, int64_t* n_static);
int mpm_set_static_rows(mpm_ctx* ctx, int64_t n_static);
/* renumber the static rows by storage slot (n_static = n_particles); between batches only */
int mpm_compact_statics(mpm_ctx* ctx, void* stream);
/* Which set holds the live state, and how many particles (n_particles[None], :81). */
int mpm_get_state(mpm_ctx* ctx, int32_t* current_set, int64_t* n_particles);
int mpm_set_state(mpm_ctx* ctx, int32_t current_set, int64_t n_particles);

/* set_gravity (:316-319) */
int mpm_set_gravity(mpm_ctx* ctx, const double* g);
/* add_sphere_collider / add_surface_collider / add_bounding_box / clear_grid_postprocess
 * (:618-692): the whole ordered table is replaced. */
int mpm_set_colliders(mpm_ctx* ctx, const mpm_collider* table, int32_t n);

/* seed_from_external_array + seed_particle (:823-838, 1081-1095): append n
 * particles whose positions are x_dev[n][dim] (f32, device).  velocity may be
 * NULL (zero). */
int mpm_seed_positions(mpm_ctx* ctx, const float* x_dev, int64_t n, int32_t material, int32_t color,
                       const double* velocity, int32_t emitter, void* stream);
/* seed (:840-850): n particles uniform in lower + u*size, u from the
 * counter-based generator documented in DESIGN.md (seed, particle id, axis). */
int mpm_seed_cube(mpm_ctx* ctx, int64_t n, const double* lower, const double* size, int32_t material,
                  int32_t color, const double* velocity, int32_t emitter, uint64_t seed, void* stream);
/* seed_ellipsoid + random_point_in_unit_sphere (:959-978): rejection sampling. */
int mpm_seed_ellipsoid(mpm_ctx* ctx, int64_t n, const double* center, const double* radius,
                       int32_t material, int32_t color, const double* velocity, int32_t emitter,
                       uint64_t seed, void* stream);
/* recover_from_external_array (:1106-1126): positions, velocities, per-particle material/color */
int mpm_seed_restart(mpm_ctx* ctx, const float* x_dev, const float* v_dev, const int32_t* material_dev,
                     const int32_t* color_dev, int64_t n, void* stream);

/* With flags bit 0 (use_g2p2g) a substep is the fused order of :773-787 -- gather from the
 * previous substep's grid, advect, then bin/scatter/grid-op at the new positions -- with the
 * semantic differences of the g2p2g kernel (:363-485).
 * One substep of step() (:789-799): deactivate_all, build_pid, p2g,
 * grid_normalization_and_gravity, grid_postprocess[*], g2p, compute_max_velocity.
 * On MPM_OK the live set has flipped; on a recoverable code nothing changed. */
int mpm_substep(mpm_ctx* ctx, double dt, double t, void* stream);
/* n substeps back to back with one host synchronisation at the end */
int mpm_substeps(mpm_ctx* ctx, double dt, double t, int32_t n, void* stream);
int mpm_get_stats(mpm_ctx* ctx, mpm_stats* out);
int mpm_set_profiling(mpm_ctx* ctx, int32_t enabled);

/* copy_dynamic(_nd) / copy_ranged(_nd) (:1146-1170): particles [begin,end) of
 * one state word (field index per the layout above) in insertion order, to
 * host memory (4 bytes per particle). */
int mpm_download(mpm_ctx* ctx, int32_t field, int64_t begin, int64_t end, void* dst_host, void* stream);
/* same, device destination */
int mpm_gather_rows(mpm_ctx* ctx, int32_t first_field, int32_t nwords, int64_t begin, int64_t end, void* dst_dev,
                    void* stream); /* rows dst[i][nwords] of consecutive words (copy_dynamic_nd, :1146-1150) */
int mpm_gather(mpm_ctx* ctx, int32_t field, int64_t begin, int64_t end, void* dst_dev, void* stream);

/* ParticleIO.write_particles on the device (engine/particle_io.py:12-76): the slice-by-slice
 * read-back of x, v and colour plus the NumPy quantisation become two calls.
 * mpm_particle_ranges: min and max of every x and v component over all particles, as f32
 *   ranges_dev[2][dim][2] = [x|v][axis][min|max] (device memory, 4*dim floats) (:42-45).
 * mpm_pack_particles: x_and_v[id][axis] = (xq << 8) + vq with
 *   q = uint32(((a - lo) * inv) * (2^bits - 1) + 0.499) in f32, bits = 24 / 8 (:50-56), and
 *   color[id][0..2] = (c >> 16, c >> 8, c) & 255 (:71-72), both in insertion order.
 *   lo_inv_host[2][dim][2] = [x|v][axis][lo | 1/(hi-lo)] is host memory (the caller applies the
 *   reference's degenerate-range fix, :46-47, before inverting). */
int mpm_particle_ranges(mpm_ctx* ctx, float* ranges_dev, void* stream);
int mpm_pack_particles(mpm_ctx* ctx, const float* lo_inv_host, uint32_t* x_and_v_dev, uint8_t* color_dev,
                       void* stream);



/* ---- multi-GPU slab decomposition along x (no reference counterpart: the
 * reference is single-device; SURVEY.md 8(e), DESIGN.md "Multi-GPU") ----
 * One process per GPU.  A rank owns leaf-block columns [lo, hi) (absolute block
 * x = (cell + grid_size/2) / leaf).  The host interleaves these phase calls with
 * the two neighbour exchanges (NCCL send/recv of the fixed-capacity buffers):
 *   mpm_batch_begin
 *   per substep: [exchange migration buffers] mpm_phase_unpack, mpm_phase_p2g,
 *                mpm_phase_halo_pack, [exchange halo buffers], mpm_phase_halo_add,
 *                mpm_phase_g2p
 *   mpm_batch_end        (the only host synchronisation)
 * Message buffers (device, 32-bit words): 16-word header (word 0 = count), then
 *   migration: [word][capacity], the mpm_virtual_fields(dim) words of each particle (its static attributes travel)
 *   halo:      keys[capacity] ((by << 16) | bz), then records[capacity][cells] float4 */
int mpm_set_slab(mpm_ctx* ctx, int32_t enabled, int32_t lo_block, int32_t hi_block);
/* bytes of a message buffer: kind 0 = migration (capacity in particles), 1 = halo (capacity in leaf blocks) */
size_t mpm_comm_bytes(int32_t dim, int32_t kind, int32_t capacity);
/* send buffers of the -x (lo) and +x (hi) side; NULL where there is no neighbour */
int mpm_bind_comm(mpm_ctx* ctx, void* mig_lo_dev, void* mig_hi_dev, int32_t mig_capacity, void* halo_lo_dev,
                  void* halo_hi_dev, int32_t halo_capacity);
/* particle base-cell bounding box of this rank (INT_MAX/INT_MIN when it holds none); synchronises */
int mpm_get_bbox(mpm_ctx* ctx, int32_t* bb_min, int32_t* bb_max, void* stream);
/* key-layout box common to all ranks (the all-reduced bounding box), global signed cell indices */
int mpm_set_layout_box(mpm_ctx* ctx, int32_t enabled, const int32_t* bb_min, const int32_t* bb_max);
int mpm_batch_begin(mpm_ctx* ctx, void* stream);
/* Right after mpm_batch_begin: dry run of the block discovery of the first substep; *need_blocks = leaf blocks the current
 * particles occupy / reach.  A block-capacity miss cannot be retried inside a distributed batch (the neighbours have run
 * ahead), so the host sizes `max_blocks` from this after seeding and keeps a margin afterwards.  Synchronises. */
int mpm_batch_probe(mpm_ctx* ctx, int32_t* need_blocks, void* stream);
int mpm_phase_unpack(mpm_ctx* ctx, const void* from_lo_dev, const void* from_hi_dev, void* stream);
int mpm_phase_p2g(mpm_ctx* ctx, double dt, void* stream);
int mpm_phase_halo_pack(mpm_ctx* ctx, void* stream);
int mpm_phase_halo_add(mpm_ctx* ctx, const void* from_lo_dev, const void* from_hi_dev, void* stream);
int mpm_phase_g2p(mpm_ctx* ctx, double dt, void* stream);
int mpm_batch_end(mpm_ctx* ctx, void* stream);
/* Peer path: the producing kernels (halo pack, G2P) write their records directly into the
 * neighbour's receive buffers over NVLink (CUDA IPC mapping) and publish a per-substep epoch
 * with a system-scope release store; the consumer spins on its local epoch word (4 s
 * time-out).  No NCCL call and no host work per substep.  mpm_peer_alloc makes the one
 * cudaMalloc block this library owns (IPC needs a whole allocation); the 64-byte handle is
 * exchanged by the host (all_gather) and opened with mpm_peer_open (side 0 = -x neighbour). */
int mpm_peer_alloc(mpm_ctx* ctx, int32_t mig_capacity, int32_t halo_capacity);
int mpm_peer_handle(mpm_ctx* ctx, void* out64);
int mpm_peer_open(mpm_ctx* ctx, int32_t side, const void* handle64);
/* `count` substeps (or, deliver_only != 0, just the pending particle delivery) inside
 * mpm_batch_begin/mpm_batch_end; every rank must call it with the same arguments */
int mpm_peer_substeps(mpm_ctx* ctx, double dt, int32_t count, int32_t deliver_only, void* stream);
/* Seeding and read-back with slabs (every rank makes the same add_* calls, or hands in pre-partitioned rows):
 * mpm_seed_positions_slab: seed_from_external_array (:1081-1095) for the rows of x_dev[n][dim] whose base block
 *   lies in this rank's slab (all rows without a slab); row i carries the id id_base + i; *kept = rows appended.
 *   n <= bound capacity.  Synchronises.
 * mpm_seed_generate: the positions seed (:840-850, mode 1: lower, size) / seed_ellipsoid (:959-978, mode 2: center,
 *   radius) would give the particles with ids [id0, id0 + n), to x_out_dev[n][dim]; appends nothing.
 * mpm_export_local: particle_info (:1172-1180) of the rows this rank owns, compacted into field blocks sized for
 *   n = n_particles rows each: out_dev = [x n*dim | v n*dim | material n | color n | id n] 32-bit words; the first
 *   *count rows of every block are valid (each block is ready to be copied into its own host array).  Synchronises. */
int mpm_seed_positions_slab(mpm_ctx* ctx, const float* x_dev, int64_t n, int64_t id_base, int32_t material,
                            int32_t color, const double* velocity, int32_t emitter, int64_t* kept, void* stream);
int mpm_seed_generate(mpm_ctx* ctx, int32_t mode, int64_t n, int64_t id0, const double* a, const double* b,
                      uint64_t seed, float* x_out_dev, void* stream);
int mpm_export_local(mpm_ctx* ctx, void* out_dev, int64_t* count, void* stream);
/* Re-cutting: after mpm_set_slab moved this rank's columns, one round of the bulk move of the rows that now lie outside
 * (at most `capacity` per side) into LOCAL send buffers of the migration format; the host exchanges the buffers with the
 * neighbours and gives the received ones to mpm_phase_unpack inside mpm_batch_begin/end; repeat until no rank reports
 * rows left outside.  out3 = {rows sent to -x, rows sent to +x, rows still outside}.  Synchronises.
 * BEFORE mpm_set_slab, call it once with capacity 0 (buffers NULL): rows outside the OLD columns are leavers of the last
 * substep that were already delivered; they are marked dead so that a cut moving over them cannot revive them. */
int mpm_rebalance_pack(mpm_ctx* ctx, void* send_lo_dev, void* send_hi_dev, int32_t capacity, int64_t* out3, void* stream);
/* rows [0, n) of one state word in storage order (pair with the `id` word) */
int mpm_download_raw(mpm_ctx* ctx, int32_t field, void* dst_host, void* stream);

/* ---- mesh seeding (setup path of add_mesh, engine/mpm_solver.py:1049-1079) ---- */
/* Voxelizer.voxelize_triangles (engine/voxelizer.py:46-109): signed winding
 * count per voxel of the super-sampled grid, rasterised in f64.  `res` is the
 * super-sampled resolution, `dx` the voxel size; `voxels_dev` is a dense,
 * zero-initialised int32 box [box_lo, box_hi) (row-major x,y,z) that must
 * cover every pixel column the triangles touch. */
int mpm_voxelize(int32_t device, const double* tris_dev, int64_t ntri, const int32_t* res, double dx,
                 int32_t padding, const int32_t* box_lo, const int32_t* box_hi, int32_t* voxels_dev, void* stream);
/* seed_from_voxels (engine/mpm_solver.py:1017-1047) in two passes over the same
 * box: pass 0 writes the particle count of every voxel to counts_dev; pass 1
 * writes positions [n][3] (f32) at offsets_dev (exclusive prefix sum of the
 * counts, int64).  cell = dx / super_sample. */
int mpm_voxel_sample(int32_t device, const int32_t* voxels_dev, const int32_t* res, const int32_t* box_lo,
                     const int32_t* box_hi, int32_t sample_density, int32_t super_sample, double cell,
                     const double* translation, int32_t grid_size, int32_t padding, uint64_t seed, int32_t pass,
                     int32_t* counts_dev, const int64_t* offsets_dev, float* x_out_dev, void* stream);

/* ---- parity/debug getters (tests only; not on the hot path) ---- */
/* build_pid result: per particle (insertion order) the leaf-block coordinate
 * (base - offset) // leaf, int32 [n][dim] to host. */
int mpm_debug_binning(mpm_ctx* ctx, int32_t* block_host, void* stream);
/* the binning's single-launch exclusive prefix sum (k_scan_excl) on caller data, device pointers:
 * out[i] = in[0] + ... + in[i-1]; n is bounded by the bound workspace (tests) */
int mpm_debug_scan(mpm_ctx* ctx, const int32_t* in_dev, int32_t* out_dev, int64_t n, void* stream);
/* Structure of the LAST substep: particle blocks (coords [npb][dim], counts[npb])
 * and grid (active) blocks (coords [ngb][dim]); arrays may be NULL to query sizes. */
int mpm_debug_blocks(mpm_ctx* ctx, int32_t* pb_coords_host, int32_t* pb_counts_host, int32_t* npb,
                     int32_t* gb_coords_host, int32_t* ngb);
/* Grid of the LAST substep after the grid op: for every cell of every grid
 * block, global signed cell index [ncell][dim] and (v[dim], m) as [ncell][4]
 * floats (2D: vx, vy, m, 0). */
int mpm_debug_grid(mpm_ctx* ctx, int32_t* cell_host, float* vm_host, int64_t max_cells, int64_t* ncell);
/* Device SVD / constitutive update on host arrays (round trip through the GPU). */
int mpm_debug_particle_update(mpm_ctx* ctx, double dt, int64_t n, const int32_t* material_host,
                              float* F_host, const float* C_host, float* Jp_host, float* affine_host,
                              float* mass_host);

#ifdef __cplusplus
}
#endif
#endif /* MPM_B200_H */
