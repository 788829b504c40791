/*
 * raystrack_b200.h -- C ABI of librsk_b200.so: the B200 (sm_100a) implementation of Raystrack's
 * Monte-Carlo view-factor hot path.
 *
 * This is the drop-in boundary.  The reference (philip-ba/raystrack v1.0.2) is pure Python: its host
 * orchestrator (src/raystrack/main.py) calls Numba kernels with flat C-contiguous NumPy arrays.  The entry
 * points below are what a ctypes binding for that kernel boundary binds; each one names the reference
 * interface it replaces (paths relative to /root/reference/src/raystrack/).  INTEGRATION.md shows the
 * reference-side stub.
 *
 * Conventions
 *   - plain C: pointers + sizes, no C++/torch types.  Host pointers unless a parameter says "device".
 *   - every function returns 0 on success, a negative rsk_status otherwise; rsk_last_error() returns the
 *     message of the last failure on the calling thread.  No exceptions cross the boundary.
 *   - objects are opaque handles owned by the library; destroy in reverse order of creation.
 *   - all work of a context is ordered on one CUDA stream (its own, or one supplied by the caller, e.g.
 *     torch.cuda.current_stream().cuda_stream).  Calls that return host data synchronise that stream.
 *   - float arrays of shape [n,3] are row-major float32 exactly as the reference passes them.
 */
#ifndef RAYSTRACK_B200_H
#define RAYSTRACK_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RSK_ABI_VERSION 2

typedef enum rsk_status {
    RSK_OK = 0,
    RSK_ERR_INVALID = -1,   /* bad argument */
    RSK_ERR_CUDA = -2,      /* CUDA runtime failure (message in rsk_last_error) */
    RSK_ERR_NO_DEVICE = -3, /* no CUDA device / wrong architecture */
    RSK_ERR_OOM = -4
} rsk_status;

typedef struct rsk_ctx rsk_ctx;
typedef struct rsk_scene rsk_scene;
typedef struct rsk_emitters rsk_emitters;
typedef struct rsk_solve rsk_solve;
typedef struct rsk_geometry rsk_geometry;
typedef struct rsk_tally_block rsk_tally_block;

/* ------------------------------------------------------------------------------------------- library */

const char *rsk_last_error(void);
int rsk_abi_version(void);
/* Hash of the sources this binary was built from (the Python loader rebuilds a library that does not match its tree). */
const char *rsk_source_hash(void);
/* Number of visible CUDA devices (0 without a driver); replaces numba.cuda.is_available() (main.py:136-147). */
int rsk_device_count(int *count);

/* ------------------------------------------------------------------------------------------- context */

/* One context per (process, GPU).  `stream` is a cudaStream_t to order all work on (pass cudaStreamLegacy, i.e.
 * (void*)1, to name the legacy default stream), or NULL for a private non-blocking stream.  Replaces cuda.get_current_device()/cuda.stream() (main.py:109, 419-495). */
int rsk_ctx_create(int device_ordinal, void *stream, rsk_ctx **out);
int rsk_ctx_destroy(rsk_ctx *ctx);
int rsk_ctx_synchronize(rsk_ctx *ctx);
/* Benchmark aid: with bytes > 0 every trace launch of the context is preceded, on the same stream, by a write of that
 * many bytes of scratch memory (larger than the L2: the scene is then re-read from HBM in every iteration).  0 = off. */
int rsk_ctx_set_l2_flush(rsk_ctx *ctx, int64_t bytes);
/* CUDA-event stopwatch on the context stream (bench.py times the kernels with it); covers pipelined solves. */
int rsk_ctx_timer_start(rsk_ctx *ctx);
int rsk_ctx_timer_stop(rsk_ctx *ctx, float *elapsed_ms);
/* Number of kernels this context has launched so far (bench.py's gpu_launches). */
int rsk_ctx_launch_count(rsk_ctx *ctx, int64_t *count);
/* name[256]; info: [0]=SM count, [1]=cc major, [2]=cc minor, [3]=total global memory bytes. */
int rsk_ctx_device_info(rsk_ctx *ctx, char *name, int64_t *info);

/* ------------------------------------------------------------------------------------------- scene
 * Replaces prepare_scene + build_bvh + PreparedSolver.get_device_scene
 * (utils/prepared.py:170-243, 381-403; utils/bvh.py:14-72).
 * Input is the reference's flattened scene in mesh order: v0,e1,e2,normals float32[n_tri,3], sid int32[n_tri]
 * (mesh index of each triangle, 0 <= sid < n_surf).  With use_bvh != 0 the library builds, on the GPU, a
 * Morton-code LBVH and collapses it into 96-byte 8-wide quantised nodes; otherwise rays test every triangle
 * in input order (the reference's bvh="off" path, same tie order). */
int rsk_scene_create(rsk_ctx *ctx, const float *v0, const float *e1, const float *e2, const float *normals,
                     const int32_t *sid, int64_t n_tri, int32_t n_surf, int32_t use_bvh, rsk_scene **out);
int rsk_scene_destroy(rsk_scene *scene);
/* info: [0]=n_tri, [1]=n_surf, [2]=use_bvh, [3]=wide nodes, [4]=node bytes, [5]=triangle bytes,
 *       [6]=wide-tree depth, [7]=build time in microseconds (device). */
int rsk_scene_info(rsk_scene *scene, int64_t *info);
/* Test hook: copy the wide BVH back.  nodes: n_nodes*96 bytes; tri_index: int32[n_tri] = input index of the
 * triangle stored at each slot of the traversal-order triangle array.  Either pointer may be NULL. */
int rsk_scene_download_bvh(rsk_scene *scene, void *nodes, int32_t *tri_index);

/* ------------------------------------------------------------------------------------------- emitters
 * Replaces prepare_emitters' per-mesh arrays + PreparedSolver.get_device_emitter + the Halton tables
 * (utils/prepared.py:246-321, 405-431; utils/halton.py:9-58).  Triangles of all emitters are concatenated;
 * emitter i owns rows [tri_offset[i], tri_offset[i+1]).  g[i] is the Halton grid side (helpers.py:8-11),
 * rays_per_cell the `rays` parameter: emitter i shoots g[i]^2 * rays_per_cell rays per iteration.  A negative g[i]
 * marks a zero-area emitter: grid side -g[i], jitter and Halton values all zero (utils/prepared.py:278-287).
 * The five per-ray Halton dimensions (bases 5,2,3,7,11) and the per-cell jitter grids are generated on the
 * device with the reference's exact float64 operation order and cached in the context. */
int rsk_emitters_create(rsk_ctx *ctx, int32_t n_emit, const int64_t *tri_offset,
                        const float *tri_a, const float *tri_e1, const float *tri_e2,
                        const float *tri_u, const float *tri_v, const float *tri_n,
                        const float *tri_eps, const float *cdf,
                        const int32_t *g, int32_t rays_per_cell, rsk_emitters **out);
int rsk_emitters_destroy(rsk_emitters *em);
/* Test hook: the cached tables.  dims: float32[5][n] (rows: bases 5,2,3,7,11), n <= max rays/iteration;
 * grid_u/grid_v: float32[g*g] for a grid side used by one of the emitters. */
int rsk_emitters_download_tables(rsk_emitters *em, int64_t n, float *dims, int32_t g, float *grid_u, float *grid_v);

/* ------------------------------------------------------------------------------------------- device-side preparation
 * Replaces prepare_scene / prepare_emitters (utils/prepared.py:93-321: `_unit_rows`, `_triangle_frames`,
 * `_triangle_origin_eps`, the area CDF, `grid_from_density`, the statistics of `_emitter_plane`) for callers that
 * hold raw meshes: only vertices and faces are uploaded, the per-triangle records are computed on the GPU with
 * NumPy's float32 operation order (bit-identical to the host arrays rsk_scene_create / rsk_emitters_create take).
 * verts: float32[n_vert][3] of all meshes back to back (the reference casts vertices to float32 first,
 * prepared.py:182-188); faces: int32[n_tri][3] with indices local to their mesh; mesh i owns vertices
 * [vert_offset[i], vert_offset[i+1]) and triangles [tri_offset[i], tri_offset[i+1]).  Mesh i becomes surface i. */
int rsk_geometry_create(rsk_ctx *ctx, int32_t n_mesh, const float *verts, const int64_t *vert_offset,
                        const int32_t *faces, const int64_t *tri_offset, rsk_geometry **out);
int rsk_geometry_destroy(rsk_geometry *geometry);
/* Same result as rsk_scene_create on prepare_scene's arrays. */
int rsk_scene_from_geometry(rsk_geometry *geometry, int32_t use_bvh, rsk_scene **out);
/* Per-mesh by-products of the emitter preparation.  total_area = float(areas.sum()) in NumPy's pairwise float32
 * order (it fixes the grid side g = max(ceil(sqrt(total_area * density)), 4)); origin / normal0 = first corner and
 * unit normal of the first triangle; eps_max = largest ray-origin offset.  min_dot, worst, worst_mag are float64
 * statistics for the planarity test of `_emitter_plane` (prepared.py:133-167): min over triangles of n.normal0; max
 * over corners p of |(p - origin).normal0|; and the size of the terms of that dot product (its rounding scale). */
typedef struct rsk_mesh_summary {
    double total_area;
    float origin[3];
    float normal0[3];
    float eps_max;
    float reserved;
    double min_dot;
    double worst;
    double worst_mag;
} rsk_mesh_summary;
/* Same device object as rsk_emitters_create on prepare_emitters' arrays (faces flipped to [0,2,1] if flip_faces);
 * density is the `samples` parameter.  summary: rsk_mesh_summary[n_mesh], filled on return. */
int rsk_emitters_from_geometry(rsk_geometry *geometry, double density, int32_t rays_per_cell, int32_t flip_faces,
                               rsk_emitters **out, rsk_mesh_summary *summary);
/* g: int32[n_emit] grid sides, n_rays_once: int64[n_emit]; either may be NULL. */
int rsk_emitters_info(rsk_emitters *em, int32_t *g, int64_t *n_rays_once);
/* Test hooks: the packed device records.  records: float32[n_tri][20] = (a,eps)(e1,n.x)(e2,n.y)(u,n.z)(v,0) per
 * emitter triangle, cdf: float32[n_tri];  tri: float32[n_tri][12] = (v0,sid)(e1,0)(e2,0) and normals:
 * float32[n_tri][4] = (n,sid) in traversal order (input order for a scene without BVH). */
int rsk_emitters_download_records(rsk_emitters *em, float *records, float *cdf);
int rsk_scene_download_triangles(rsk_scene *scene, float *tri, float *normals);

/* surf_active of every emitter at once (replaces `_build_emitter_surface_mask`, main.py:167-204): active_out is
 * uint8[n_emit][n_surf]; emitter e switches off its own mesh and, if planar[e], every mesh whose bounding box
 * (centers/extents float32[n_surf][3], utils/prepared.py:359-375) lies wholly behind the plane
 * (plane_origin/plane_normal float32[n_emit][3], plane_tol float32[n_emit]).  Same float32 operation order. */
int rsk_surface_masks(rsk_ctx *ctx, int32_t n_emit, int32_t n_surf, const uint8_t *planar, const float *plane_origin,
                      const float *plane_normal, const float *plane_tol, const float *centers, const float *extents,
                      uint8_t *active_out);

/* ------------------------------------------------------------------------------------------- per-ray hook
 * Replaces build_rays + trace_cpu_[bvh_]firsthit / trace_cpu_[bvh_]hitmask called back to back
 * (utils/ray_builder.py:25-94; utils/cpu_trace.py:54-277, 540-732) for ONE emitter and ONE iteration, writing
 * per-ray outputs.  cp = [cp_grid[0..1], cp_dims[0..4]] (main.py:1810-1812).  mode 0: closest hit
 * (hit_sid = mesh index or -1, hit_front = 1 if front side); mode 1: any hit (hit_sid = 0/1 hit mask,
 * hit_front = Tregenza patch id of a miss with dz>0, else 255).  Any output pointer may be NULL.
 * first_ray/n_rays select a sub-range of the iteration's rays. */
int rsk_trace_rays(rsk_ctx *ctx, rsk_scene *scene, rsk_emitters *em, int32_t emitter,
                   const uint8_t *surf_active, int32_t emit_sid, int32_t min_sid, const float *cp,
                   int32_t mode, int64_t first_ray, int64_t n_rays,
                   float *orig, float *dirs, int32_t *hit_sid, uint8_t *hit_front);

/* Relative cost per ray of a set of emitters (no reference counterpart; used to balance multi-GPU plans):
 * one launch traces the first sample_rays rays of each emitter with the masks and skip rules of a matrix solve
 * (surf_active / emit_sid / min_sid as in rsk_matrix_begin, cp = one rotation row) and every CTA adds the SM clock
 * ticks it was resident to its emitter.  ticks: int64[n_local]; rays_out (may be NULL): the rays each entry covers. */
int rsk_emitter_costs(rsk_ctx *ctx, rsk_scene *scene, rsk_emitters *em, const int32_t *emit_ids, int32_t n_local,
                      const uint8_t *surf_active, const int32_t *emit_sid, const int32_t *min_sid, const float *cp,
                      int64_t sample_rays, int64_t *ticks, int64_t *rays_out);

/* ------------------------------------------------------------------------------------------- matrix solve
 * Replaces the emitter/iteration loops of view_factor_matrix (main.py:1757-1945): per iteration and emitter
 * build_rays -> trace -> reduce_first_hits -> Welford update -> convergence test, for a SET of emitters at once.
 *
 *   emit_ids[n_local]             emitters handled by this solve (this GPU's shard)
 *   surf_active[n_local][n_surf]  main.py:167-204 mask per emitter; emit_sid/min_sid[n_local]: main.py:1181-1182
 *   cp_table[n_rot][7]            rotations; iteration `it` of local emitter k uses row rot_base[k] + it
 *   ray_range[n_local][2]         NULL, or per job the slice [begin,end) of the emitter's rays traced by THIS
 *                                 context (multi-GPU ray-range sharding of very large emitters; see below)
 *   tol_mode: 0 = "stderr", 1 = "delta";  interval: convergence_interval (1 = the reference's CPU behaviour)
 *
 * rsk_matrix_step enqueues `n_iters` further iterations for every unconverged emitter: one fused
 * raygen+closest-hit+tally kernel and one statistics/convergence kernel per iteration, no host round trip in
 * between; emitters that converge are skipped on the device.  It then returns how many emitters are still
 * running.  rsk_matrix_read copies the integer tallies back. */
typedef struct rsk_solve_params {
    int32_t max_iters;
    int32_t min_iters;
    int32_t interval;
    int32_t tol_mode;
    double tol;
} rsk_solve_params;

int rsk_matrix_begin(rsk_ctx *ctx, rsk_scene *scene, rsk_emitters *em,
                     const int32_t *emit_ids, int32_t n_local,
                     const uint8_t *surf_active, const int32_t *emit_sid, const int32_t *min_sid,
                     const float *cp_table, int32_t n_rot, const int32_t *rot_base, const int64_t *ray_range,
                     const rsk_solve_params *params, rsk_solve **out);
/* Iterations are pipelined: iteration i is traced and folded on stream (i & 1) of the context with its own tally
 * buffer, so the trace of iteration i + 1 occupies the SMs while iteration i drains and its statistics run (statistics
 * are still folded strictly in iteration order).  A job that stops at iteration i may be traced once more before its
 * stop decision is known; those tallies are discarded, so results do not depend on the overlap.  RSK_PIPELINE=0 disables. */
int rsk_matrix_step(rsk_solve *solve, int32_t n_iters, int32_t *n_active);
/* hits_front/hits_back: int64[n_local][n_surf]; iters: int32[n_local]; total_rays: int64[n_local];
 * stderr_front/back: float64[n_local][n_surf] replicate standard errors (may be NULL). */
int rsk_matrix_read(rsk_solve *solve, int64_t *hits_front, int64_t *hits_back, int32_t *iters,
                    int64_t *total_rays, double *stderr_front, double *stderr_back);
/* The whole tally block in one contiguous copy: matrix solves int64[n_local][n_surf][2] ((front, back) pair per
 * receiver), sky solves int64[n_local][145 or 1]; what the host driver uses. */
int rsk_solve_read_block(rsk_solve *solve, int64_t *tallies, int32_t *iters, int64_t *total_rays);
/* rsk_solve_read_block without the second host copy: *tallies_view points into the context's pinned staging area
 * (valid until the next staged download of this context). */
int rsk_solve_read_block_view(rsk_solve *solve, int64_t **tallies_view, int32_t *iters, int64_t *total_rays);
/* Result rows in compressed form.  Replaces the dense read-back plus the host loop of main.py:1918-1934 (`F = hits /
 * total_rays`, keys only for F > 0) for large matrices: per job (row) the column indices (ascending: the reference's
 * key order "<r0>_front, <r0>_back, <r1>_front, ...") and float64 values F of the non-zero bins.  rsk_solve_csr /
 * rsk_tally_block_csr build the rows on the device and return row_ptr (int64[n_rows + 1]; nnz = row_ptr[n_rows]);
 * rsk_csr_fetch then copies cols int32[nnz] and vals float64[nnz] and releases the device copy.  total_rays:
 * int64[n_rows], the denominator of each row (a solve uses its own counters). */
int rsk_solve_csr(rsk_solve *solve, int64_t *row_ptr);
int rsk_tally_block_csr(rsk_tally_block *block, const int64_t *total_rays, int64_t *row_ptr);
int rsk_csr_fetch(rsk_ctx *ctx, int32_t *cols, double *vals);
/* Device pointer + element count of the int64 tally block [n_local][n_surf][2] ((front, back) per receiver) for collectives issued by the caller (torch.distributed / NCCL). */
int rsk_matrix_device_tallies(rsk_solve *solve, void **device_ptr, int64_t *n_elements);

/* ------------------------------------------------------------------------------------------- sky solve
 * Replaces the loops of view_factor_to_tregenza_sky (main.py:1997-2185): any-hit trace against every active
 * non-emitter mesh (emit_sid = emitter index, min_sid = 0), misses with dz>0 binned into the 145 Tregenza
 * patches (utils/cpu_trace.py:735-798) when discrete != 0, else counted. */
int rsk_sky_begin(rsk_ctx *ctx, rsk_scene *scene, rsk_emitters *em,
                  const int32_t *emit_ids, int32_t n_local, const uint8_t *surf_active,
                  const float *cp_table, int32_t n_rot, const int32_t *rot_base, const int64_t *ray_range,
                  const rsk_solve_params *params, int32_t discrete, rsk_solve **out);
int rsk_sky_step(rsk_solve *solve, int32_t n_iters, int32_t *n_active);
/* counts: int64[n_local][145] (discrete) or int64[n_local][1]. */
int rsk_sky_read(rsk_solve *solve, int64_t *counts, int32_t *iters, int64_t *total_rays);

/* Split-phase stepping for multi-GPU solves in which an emitter's rays are divided over several GPUs: every
 * participating context enqueues the trace of its slice, the caller all-reduces (SUM) the per-iteration tally
 * block of the shared jobs on the context's stream (device pointer from rsk_solve_device_iter_tallies:
 * uint64[n_local][n_per_job], jobs in emit_ids order), then every context enqueues the fold, so all of them
 * take identical convergence decisions.  rsk_solve_poll synchronises and returns the number of running jobs.
 * rsk_matrix_step / rsk_sky_step are exactly n_iters x (enqueue_trace, enqueue_fold) followed by a poll. */
int rsk_solve_enqueue_trace(rsk_solve *solve);
int rsk_solve_enqueue_fold(rsk_solve *solve);
int rsk_solve_poll(rsk_solve *solve, int32_t *n_active);
int rsk_solve_device_iter_tallies(rsk_solve *solve, void **device_ptr, int64_t *n_per_job);
/* Make the solve accumulate its per-iteration tallies in a caller-owned device buffer (uint64[n_local][n_per_job],
 * e.g. a torch tensor that NCCL can all-reduce in place).  Call before the first iteration; the caller keeps the
 * buffer alive until rsk_solve_destroy. */
int rsk_solve_set_iter_tally_buffer(rsk_solve *solve, void *device_ptr, int64_t n_elements);

/* ------------------------------------------------------------------------------------------- shared-ray solve
 * Replaces the loops of view_factor_matrix_and_sky (main.py:1277-1660) and trace_cpu_[bvh_]combined
 * (utils/cpu_trace.py:280-522): per iteration ONE traversal per ray yields the closest receiver hit (matrix) and
 * the any-hit flag (sky).  Both sides keep their own statistics and stop independently; when one side has
 * converged the jobs degrade to the other side's plain walk.  rsk_dual_step = n_iters x (trace, matrix fold, sky
 * fold); n_active counts running (job, side) pairs.  Results: rsk_solve_read_block on the returned handle (matrix
 * side) and on the handle from rsk_dual_sky_part (sky side); rsk_solve_destroy on the returned handle frees both. */
int rsk_dual_begin(rsk_ctx *ctx, rsk_scene *scene, rsk_emitters *em, const int32_t *emit_ids, int32_t n_local,
                   const uint8_t *surf_active, const int32_t *emit_sid, const int32_t *min_sid,
                   const float *cp_table, int32_t n_rot, const int32_t *rot_base,
                   const rsk_solve_params *matrix_params, const rsk_solve_params *sky_params, int32_t discrete,
                   rsk_solve **out);
/* The same with ray_range[n_local][2] as in rsk_matrix_begin (multi-GPU ray slices of oversized emitters).  A sliced
 * dual solve is stepped split-phase: rsk_solve_enqueue_trace(solve) traces both sides at once, the caller all-reduces
 * the iteration tallies of BOTH handles, then rsk_solve_enqueue_fold on each handle; rsk_solve_poll on each. */
int rsk_dual_begin_sliced(rsk_ctx *ctx, rsk_scene *scene, rsk_emitters *em, const int32_t *emit_ids, int32_t n_local,
                          const uint8_t *surf_active, const int32_t *emit_sid, const int32_t *min_sid,
                          const float *cp_table, int32_t n_rot, const int32_t *rot_base, const int64_t *ray_range,
                          const rsk_solve_params *matrix_params, const rsk_solve_params *sky_params, int32_t discrete,
                          rsk_solve **out);
int rsk_dual_step(rsk_solve *solve, int32_t n_iters, int32_t *n_active);
int rsk_dual_sky_part(rsk_solve *solve, rsk_solve **sky);

int rsk_solve_destroy(rsk_solve *solve);
/* Rays traced so far by this solve (all emitters, all iterations). */
int rsk_solve_rays_traced(rsk_solve *solve, int64_t *rays);
/* Diagnostic work counters of the trace kernels since the last reset: out[0] node visits, [1] triangle tests,
 * [2] triangles skipped by the surface mask, [3] rays, [4] triangle-flush trips, [5] stack pushes, [6..7] 0.  Counted
 * only by builds with -DRSK_COUNTERS=1 (scripts/kernel_variants.py); the product build returns zeros.  The reference
 * has no counterpart; SURVEY.md 8(d) measured the same quantities by instrumenting cpu_trace.py:120-277. */
int rsk_trace_counters(rsk_ctx *ctx, int64_t *out, int32_t reset);

/* ------------------------------------------------------------------------------------------- multi-GPU
 * The reference is single-device; this is the exchange step of the B200 design:
 * one process per GPU, each with its own context; emitters (matrix rows, independent units -- main.py:1758-1939)
 * are sharded over the ranks, the BVH is replicated, and the int64 tallies are summed over NVLink/NVSwitch.  NCCL is
 * bound at run time (libnccl.so.2 already in the process, else $RSK_NCCL_LIBRARY, else the system library).
 *
 * rsk_comm_unique_id: rank 0 creates the 128-byte NCCL id and hands it to the other ranks out of band (file, socket,
 * MPI, torch.distributed store ...).  rsk_comm_init joins the communicator (collective: every rank calls it); all
 * later collectives run on the context's stream, ordered with its kernels -- no host synchronisation in between. */
#define RSK_COMM_ID_BYTES 128
int rsk_comm_unique_id(uint8_t *id);
int rsk_comm_init(rsk_ctx *ctx, const uint8_t *id, int32_t rank, int32_t nranks);
int rsk_comm_destroy(rsk_ctx *ctx);
/* rank / nranks of the context's communicator (0 / 1 without one); nccl_version e.g. 22809 (0 if NCCL is not loaded). */
int rsk_comm_info(rsk_ctx *ctx, int32_t *rank, int32_t *nranks, int32_t *nccl_version);
/* In-place all-reduce of n int64 values in DEVICE memory on the context's stream; op 0 = sum, 1 = max.  A context
 * without communicator (or with one rank) returns at once. */
int rsk_allreduce_i64(rsk_ctx *ctx, void *device_ptr, int64_t n, int32_t op);
/* The same for n <= 4096 HOST values (iteration counters, "is any rank still running"); synchronises the stream. */
int rsk_allreduce_host_i64(rsk_ctx *ctx, int64_t *values, int64_t n, int32_t op);
/* Split-phase stepping (see rsk_solve_enqueue_trace): sum the per-iteration tallies of the first n_jobs jobs of
 * `solve` (the ray-split emitters, first in emit_ids on every rank) over the communicator, on the context's stream. */
int rsk_solve_allreduce_iter_tallies(rsk_solve *solve, int32_t n_jobs);
/* Device-resident int64 [n_rows][n_cols] block, zero-initialised, in which the ranks assemble the result of a
 * sharded solve: add_solve copies the totals of the solve's jobs into the rows named by their emitter ids
 * (keep: uint8[n_local] or NULL; jobs with keep == 0 are skipped -- a ray-split job is replicated on every rank and
 * must be counted once), allreduce sums the blocks of all ranks, download brings the result to the host through
 * pinned memory.  download: dst (caller memory, may be NULL) and/or *view (pointer into the context's pinned staging
 * area, valid until the next staged download of this context; may be NULL). */
int rsk_tally_block_create(rsk_ctx *ctx, int64_t n_rows, int64_t n_cols, rsk_tally_block **out);
int rsk_tally_block_add_solve(rsk_tally_block *block, rsk_solve *solve, const uint8_t *keep);
int rsk_tally_block_allreduce(rsk_tally_block *block);
int rsk_tally_block_device(rsk_tally_block *block, void **device_ptr, int64_t *n_elements);
int rsk_tally_block_download(rsk_tally_block *block, int64_t *dst, int64_t **view);
int rsk_tally_block_destroy(rsk_tally_block *block);

/* ------------------------------------------------------------------------------------------- reciprocity
 * Replaces the dense core of enforce_reciprocity_and_rowsum (utils/helpers.py:70-96): G = 0.5*(A F + (A F)^T),
 * symmetric diagonal scaling d <- d*sqrt(target/(d .* G d)) until max|d_new - d| < tol (<= max_iter sweeps),
 * F' = D G D / A.  F is float64[n][n] row-major, overwritten with F'.  target may be NULL (= area). */
int rsk_reciprocity_rowsum(rsk_ctx *ctx, int32_t n, const double *area, const double *target,
                           double *F, double tol, int32_t max_iter, int32_t *sweeps);

#ifdef __cplusplus
}
#endif
#endif /* RAYSTRACK_B200_H */
