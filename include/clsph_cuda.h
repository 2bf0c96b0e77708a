/*
 * clsph_cuda.h -- C ABI of the B200 (sm_100a) implementation of libclsph's per-step SPH path.
 *
 * This is the whole device boundary. It replaces, in the reference tree,
 *   util/cl_boilerplate.h:41-43      init_cl_single_device / make_program / readKernelFile
 *   libclsph/sph_simulation.cpp      every cl::Buffer / cl::Kernel / cl::CommandQueue use:
 *       :366-380  kernel + buffer creation        -> clsph_create
 *       :296-319  per-step scene upload           -> clsph_set_scene        (once)
 *       :283-290, :322-323, :330-333 setArg(params, terms) -> clsph_set_parameters (once)
 *       :195      enqueueWriteBuffer(particles)   -> clsph_upload_particles
 *       :199-337  bounds, locate_in_grid, sort_particles, cell table, density_pressure,
 *                 forces, advection_collision     -> clsph_step             (device resident)
 *       :339      enqueueReadBuffer(particles)    -> clsph_download_particles
 *       :173-344  simulate_single_frame(in, out)  -> clsph_simulate_single_frame (host in/out)
 * and the six OpenCL kernels of libclsph/kernels/ (grid.cl, sort.cl, sph.cl, forces.cl,
 * smoothing.cl, advection.cl, collisions.cl), which are hand-written CUDA inside the library.
 *
 * Conventions
 *   - plain C: opaque handle, PODs by pointer with the reference's exact layouts
 *     (include/clsph/clsph_types.h), caller owns every host buffer, the library owns all
 *     device memory, streams, events and communicators;
 *   - every function returns 0 on success or a CLSPH_E* code; clsph_last_error() gives the
 *     message. (The reference prints file:line and exit(-1)s, util/cl_boilerplate.h:28-34; the
 *     host wrapper in libclsph_b200/host/ keeps that behaviour on top of these codes.)
 *   - a context is bound to one GPU and one caller thread; clsph_step() only enqueues work,
 *     the calls documented as "synchronises" wait for it;
 *   - there is no CPU fallback: without a usable CUDA device clsph_create() fails.
 */
#ifndef CLSPH_CUDA_H_
#define CLSPH_CUDA_H_

#include <stddef.h>
#include <stdint.h>

#include "clsph/clsph_types.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct clsph_context clsph_context;

enum {
  CLSPH_OK = 0,
  CLSPH_EINVAL = 1,       /* bad argument (null pointer, N < 128, N > capacity, ...)            */
  CLSPH_ECUDA = 2,        /* a CUDA runtime call failed                                         */
  CLSPH_EGRID = 3,        /* a grid axis reached 1024 cells (reference: assert, cpp:247-249)    */
  CLSPH_ESTATE = 4,       /* call order (no particles uploaded, no parameters set, ...)         */
  CLSPH_ECOMM = 5,        /* NCCL / multi-GPU failure                                           */
  CLSPH_ENOMEM = 6        /* device or host allocation failed                                   */
};

/* What clsph_debug_fetch can read back. "sorted" = order of the particle array after the
 * last sub-step (the reference's output order). Taps other than the first three need
 * clsph_set_debug(ctx, 1) before the step. */
enum {
  CLSPH_TAP_SORTED_KEYS = 0,   /* uint32[N]  Morton cell key per sorted particle               */
  CLSPH_TAP_PERMUTATION = 1,   /* uint32[N]  sorted r came from pre-step index perm[r]         */
  CLSPH_TAP_CELL_TABLE = 2,    /* uint32[grid_cell_count] reference form: first idx, key >= c  */
  CLSPH_TAP_KEYS_INPUT = 3,    /* uint32[N]  key per particle in pre-step order      (debug)   */
  CLSPH_TAP_CANDIDATE_COUNT = 4, /* uint32[N] sorted: sum over 27 cells of end-start (debug)   */
  CLSPH_TAP_SUPPORT_COUNT = 5, /* uint32[N]  sorted: candidates with r/h < 1         (debug)   */
  CLSPH_TAP_DENSITY = 6,       /* float[N]   sorted                                            */
  CLSPH_TAP_PRESSURE = 7,      /* float[N]   sorted                                            */
  CLSPH_TAP_ACCELERATION = 8,  /* float[3N]  sorted, after the force pass            (debug)   */
  CLSPH_TAP_COLLISION_ITERS = 9 /* uint32[N] sorted, advect/collide loop trips       (debug)   */
};

/* Device time per stage of the step, accumulated while profiling is on. */
typedef struct clsph_stage_times {
  double ms_bounds_grid;   /* AABB reduce (first step only) + grid setup                       */
  double ms_keys;          /* cell keys + radix digit histograms                               */
  double ms_sort;          /* histogram scan + onesweep passes                                 */
  double ms_reorder;       /* gather into sorted order + cell start/end table                  */
  double ms_density;       /* density / pressure pass                                          */
  double ms_forces;        /* force pass                                                       */
  double ms_integrate;     /* leapfrog + collisions + next step's AABB                         */
  double ms_exchange;      /* multi-GPU migration + halo traffic (0 on one GPU)                */
  uint64_t substeps;       /* sub-steps covered by the sums above                              */
  uint64_t kernel_launches;/* kernels launched by the library since profiling was enabled      */
} clsph_stage_times;

/* ---- lifecycle ------------------------------------------------------------------------- */

/* Number of CUDA devices visible to the process (0 if none / no driver). Never fails. */
int clsph_device_count(void);

/* Creates a context on CUDA device `device` able to hold `max_particles` particles.
 * `cell_table_capacity` = entries of the dense Morton-indexed cell table (0 = choose from
 * max_particles); steps whose grid_cell_count exceeds it fall back to binary search over the
 * sorted keys, with identical results. Replaces sph_simulation.cpp:354-380. */
int clsph_create(clsph_context** out, int device, uint32_t max_particles, uint32_t cell_table_capacity);
void clsph_destroy(clsph_context* ctx);

/* Message for the last non-zero return on `ctx` (or of clsph_create when ctx is NULL). */
const char* clsph_last_error(const clsph_context* ctx);

/* ---- inputs ---------------------------------------------------------------------------- */

/* The three scene arrays of libclsph/scene.h:11-14 (normals 3F floats, vertices, indices 3F).
 * Uploaded once; the reference re-uploads them every sub-step (sph_simulation.cpp:296-319). */
int clsph_set_scene(clsph_context* ctx, const float* face_normals, const float* vertices,
                    size_t n_vertex_floats, const uint32_t* indices, uint32_t face_count);

/* Fluid / time-step constants and smoothing terms, as load_settings derives them
 * (sph_simulation.cpp:405-506). The grid block of `params` is ignored on input. */
int clsph_set_parameters(clsph_context* ctx, const simulation_parameters* params,
                         const precomputed_kernel_values* terms);

/* Tuning knobs; results are the same (to rounding) whatever they are set to. The environment variable
 * CLSPH_OPTIONS="name=value,name=value" applies such pairs to every context at creation. The defaults are the
 * organisation measured fastest on a B200 (profiles/r02_*); the others are kept as fallbacks and for A/B runs.
 *   "sub_cell_order"   1 (default): keep the arrays in HBM sorted by (cell key << 3 | octant of the cell), so
 *                      that each particle searches the ~27 sub-cells of side h around it (~130
 *                      candidates) instead of the 27 cells of side 2h (~1000). The reference's array
 *                      order is carried as a per-particle rank; downloads and taps are in the
 *                      reference's order exactly as with 0.
 *                      Needs a grid whose Morton cell count stays below 2^29 (z axis < 512 cells; larger grids
 *                      report CLSPH_EGRID and want 0); set it before particles are uploaded.
 *                      0: the established organisation on whole cells (k_density_lists / k_forces_lists).
 *   "count_sort"       (with sub_cell_order) 1 (default): while the grid fits the dense sub-cell table and that table
 *                      has at most 6 words per particle, particles are sorted by counting on it (count, scan the
 *                      table in place, scatter) instead of by 8-bit radix passes; the arrays come out bit for bit
 *                      the same. 0: radix passes always. 2: counting whenever the grid fits the table (A/B runs).
 *   "pair_density"     (with sub_cell_order) 1: the density pass handles two particles of a sub-cell per
 *                      thread with packed fp32 arithmetic (FADD2 / FFMA2), bitwise the same results as 0
 *                      (one particle per thread). -1 (default): 1 from 160 000 particles, 0 below (where the GPU is
 *                      not full and the shorter walk of a one-particle thread wins). "pair_variant" 0..5 selects how a thread walks its candidates
 *                      (one by one / next load ahead / four loads ahead) and whether list entries are stored
 *                      one or two at a time; 5 (default) = four ahead, in twos.
 *   "merged_rows"      (pair_density = 0) 1 (default): the density pass walks the two index ranges of each row of
 *                      sub-cells in one loop, which evens out the loop lengths between the lanes of a warp.
 *   "factored_forces"  1 (default): the list force kernel evaluates the pair terms with the per-run constants factored
 *                      out of the sums and one MUFU.RSQ per pair (needs fast_pairs); 0: add_pair_fast.
 *   "neighbour_lists"  (sub_cell_order = 0) 1 (default): the density pass stores per-particle neighbour lists in HBM
 *                      and the force pass reads them; 0: both passes search on their own.
 *   "list_rows"        list entries kept per particle (0 = derive from the rest density; rounded up to even);
 *                      particles with more neighbours fall back to the searching force kernel.
 *   "fast_pairs"       1 (default): the list force kernel evaluates each pair with one MUFU.RSQ in place of the IEEE
 *                      square root and divide (~2 ulp, far inside the 1e-4 bar; the |r| < 1e-7 decision
 *                      stays exact). 0: the reference's operation sequence.
 *   "forces_blocks"    4 (default) or 3: resident CTAs per SM the list force kernel is compiled for
 *                      (4 = 64 registers per thread, a third more warps to hide gather latency).
 *   "face_grid"        1 (default): the collision pass tests only the scene triangles registered in the grid
 *                      cells a particle's sub-step segment touches (conservative registration:
 *                      results are bit-identical to testing every triangle, 0). */
int clsph_set_option(clsph_context* ctx, const char* name, long long value);

/* Host AoS (80-byte records) -> device SoA. n must be >= 128 (sort.cl:9-20, erratum E8) and
 * <= max_particles. Replaces sph_simulation.cpp:195. */
int clsph_upload_particles(clsph_context* ctx, const particle* aos, uint32_t n);

/* ---- the step -------------------------------------------------------------------------- */

/* Enqueues `n_substeps` sub-steps, each the equivalent of one simulate_single_frame
 * (sph_simulation.cpp:173-344): padded AABB + grid sizing, Morton cell keys, stable radix
 * sort by key, cell start/end table, density + Tait pressure, pressure/viscosity/surface
 * tension forces over the 27-cell neighbourhood, leapfrog + triangle collisions. State stays
 * on the device; returns without waiting. */
int clsph_step(clsph_context* ctx, uint32_t n_substeps);

/* Waits for all enqueued work; reports CLSPH_EGRID if any sub-step overflowed the grid. */
int clsph_synchronize(clsph_context* ctx);

/* Parameters with the grid block (grid_size_*, grid_cell_count, min_point, max_point) of the
 * most recent sub-step, as sph_simulation.cpp:229-252 leaves them. Synchronises. */
int clsph_get_parameters(clsph_context* ctx, simulation_parameters* out);

/* Device SoA -> host AoS in the reference's output order (sorted by cell key, stable):
 * position, velocity, intermediate_velocity, density, pressure, grid_index as the reference
 * writes them, acceleration = 0 (sph.cl:97-99). Synchronises. Replaces sph_simulation.cpp:339. */
int clsph_download_particles(clsph_context* ctx, particle* aos_out);

/* Frame export off the critical path. A frame file needs seven floats of a particle -- position, velocity, density
 * (libclsph/file_save_delegates/houdini_file_saver.cpp:39-62 reads exactly these; colour and mass follow from the
 * density and the parameters) -- not its 80 bytes. clsph_frame_begin packs them on the device, in the reference's
 * output order, and starts the copy into `points` (n x 7 floats; page-locked memory, e.g. from clsph_host_alloc,
 * lets the copy run while further sub-steps are enqueued and computed) on a stream of its own; it does not wait.
 * clsph_frame_end waits until `points` is complete. One frame may be pending at a time. */
int clsph_frame_begin(clsph_context* ctx, float* points, uint32_t capacity);
int clsph_frame_end(clsph_context* ctx);
/* Page-locked host memory for the above (cudaHostAlloc / cudaFreeHost without a CUDA header on the caller's side). */
int clsph_host_alloc(void** out, size_t bytes);
void clsph_host_free(void* p);

/* One sub-step with host buffers on both sides, the exact shape of
 * sph_simulation::simulate_single_frame(in, out): upload, step, download, and the grid
 * block of *params rewritten. `in` may equal `out`. `terms` may be NULL to keep the ones set. */
int clsph_simulate_single_frame(clsph_context* ctx, const particle* in, particle* out,
                                simulation_parameters* params, const precomputed_kernel_values* terms);

/* ---- multi-GPU: slab decomposition, one process per GPU ------------------------------------
 * The reference is single-device; this is new functionality with the same per-particle results.
 * The fluid is cut along x by fixed planes; every sub-step all-reduces the AABB (so all ranks
 * derive the same grid and keys), migrates particles whose cell changed owner and refreshes two
 * ghost cell layers per side, in one NCCL group on the context's stream. */

/* NCCL unique id (128 bytes) to hand to every rank, e.g. through torch.distributed. */
int clsph_comm_unique_id(void* out, size_t bytes);

/* Joins the communicator. plane_lo / plane_hi = world-space x bounds of this rank's slab
 * (-INFINITY for rank 0, +INFINITY for the last rank); neighbouring ranks must pass the same
 * plane. Slabs must stay at least four grid cells (8 h) thick (two, 4 h, with sub_cell_order, where
 * ownership follows the planes themselves instead of cells snapped to them and nothing migrates in bursts). Capacities are records per
 * message per sub-step. The planes are snapped to the cell boundaries of a grid whose origin
 * follows the fluid, so occasionally a boundary jumps by one cell and a whole cell layer
 * migrates at once: emigrant_capacity must hold one cell layer of the slab's cross-section and
 * ghost_capacity two (0 = max_particles/6 + 1024 and max_particles/3 + 1024). Exceeding them, or
 * max_particles (owned + ghost copies), is reported as CLSPH_ECOMM by the next synchronising
 * call. */
int clsph_dist_init(clsph_context* ctx, int rank, int world, const void* unique_id, float plane_lo, float plane_hi,
                    uint32_t emigrant_capacity, uint32_t ghost_capacity);

/* How the ranks exchange particles and the AABB every sub-step: "peer stores ..." (each rank stores its records
 * straight into the neighbours' memory over NVLink, CUDA IPC between the processes; no collective per sub-step) or
 * "nccl ..." (send/recv + all-reduce: chosen when peer access is unavailable or CLSPH_DIST_TRANSPORT=nccl).
 * "none" before clsph_dist_init. Results are bitwise the same either way. */
const char* clsph_dist_transport(const clsph_context* ctx);

/* This rank's particles with their global ids (any unique 32-bit labels below 2^32-1). */
int clsph_dist_upload(clsph_context* ctx, const particle* aos, const uint32_t* ids, uint32_t n);

/* Owned particles (no ghost copies) in local cell-sorted order with their ids. Pass NULL arrays
 * to query the count only. Synchronises.
 * With the option sub_cell_order the padding word of each record (byte offset 76) holds the
 * particle's rank inside its cell in the REFERENCE's order: concatenating all ranks' downloads and
 * sorting by (grid_index, that word) gives exactly the array a single device -- and the reference --
 * would hold (ids passed to clsph_dist_upload must then be the indices of the initial global
 * array), and the values themselves are bitwise those of a single-device run (every rank keeps the
 * particles of a sub-cell in the reference's order, so all sums run in the same order). Without the
 * option the word is 0 and per-particle values agree to rounding only. */
int clsph_dist_download(clsph_context* ctx, particle* aos_out, uint32_t* ids_out, uint32_t capacity, uint32_t* n_out);

/* ---- observation ----------------------------------------------------------------------- */

/* The reference's `advection_collision` kernel on its own (kernels/sph.cl:64-112, argument
 * order of sph_simulation.cpp:330-333): n records in, each with its own acceleration; out gets
 * position / velocity / intermediate_velocity advanced by one sub-step with collisions against
 * the scene set by clsph_set_scene, in the same order (no sort). Exists so the one stage whose
 * input (the post-force acceleration) the step never exports can be checked in isolation.
 * Replaces the particles held by the context. Synchronises. */
int clsph_kernel_advection_collision(clsph_context* ctx, const particle* in, particle* out, uint32_t n);

/* Enables (1) / disables (0) recording of the per-stage taps marked "(debug)" above. */
int clsph_set_debug(clsph_context* ctx, int enable);

/* Copies tap `what` of the most recent sub-step into dst (bytes = exact size). Synchronises. */
int clsph_debug_fetch(clsph_context* ctx, int what, void* dst, size_t bytes);

/* Per-stage CUDA-event timing. Enabling resets the sums. Reading synchronises. */
int clsph_profile_enable(clsph_context* ctx, int enable);
int clsph_profile_read(clsph_context* ctx, clsph_stage_times* out);

/* Number of particles currently held (after migration on multi-GPU runs this changes). */
int clsph_particle_count(clsph_context* ctx, uint32_t* n);

/* How the last sub-step sorted the particles: the number of 8-bit radix passes it ran, 0 for a counting sort on the
 * dense sub-cell table (option "count_sort"). Synchronises. */
int clsph_sort_passes(clsph_context* ctx, uint32_t* passes);

/* Raw cudaStream_t the context enqueues on (for callers that time with their own events). */
void* clsph_stream(clsph_context* ctx);

#ifdef __cplusplus
}
#endif
#endif /* CLSPH_CUDA_H_ */
