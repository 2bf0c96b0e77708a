// context.cu -- the C ABI of include/clsph_cuda.h: context, device memory, the per-sub-step
// launch sequence, taps and profiling. No arithmetic of the SPH step lives here; it only
// orders the kernels of sort.cu / grid.cu / neighbors.cu / integrate.cu on one stream.
//
// Sub-step launch sequence (all device resident, no host round trip):
//   k_grid_setup (+ zeroing of the sub-cell table and the sort's scratch) -> k_keys_hist (keys + counts) -> k_scan_table
//   -> k_onesweep x4 (the first one scatters, the others return: counting sort; or the radix passes when the grid does not
//   fit the sub-cell table) -> k_reorder_sub -> [k_rank on the side stream] k_density_pairs -> k_forces_lists_direct
//   -> k_forces_sub -> k_integrate
// (default organisation; sub_cell_order = 0: k_clear_cells -> k_reorder -> k_density_lists -> k_forces_lists -> k_integrate)
// The reference's equivalent is libclsph/sph_simulation.cpp:173-344 with 17 blocking transfers.
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <string>
#include <vector>

#include "clsph_cuda.h"
#include "dist.cuh"
#include "kernels.cuh"

using namespace clsph;

namespace {

thread_local std::string g_create_error;

enum Stage { kStBounds = 0, kStExchange, kStKeys, kStSort, kStReorder, kStDensity, kStForces, kStIntegrate, kStEnd, kStages = kStEnd };

}  // namespace

struct clsph_context {
  int device = 0;
  int sm_count = 148;
  cudaStream_t stream = nullptr;
  // side stream: work of a sub-step that nothing later in the same sub-step reads (k_rank) runs next to the
  // neighbour passes instead of in front of them; forked and joined with events
  cudaStream_t side = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  // frame export (clsph_frame_begin / clsph_frame_end): packed records leave on a copy stream while the next
  // sub-steps run
  cudaStream_t copy = nullptr;
  cudaEvent_t ev_packed = nullptr, ev_copied = nullptr;
  bool frame_pending = false;
  uint32_t capacity = 0;       // max particles
  uint32_t cell_capacity = 0;  // dense table entries
  uint32_t n = 0;              // particles held
  bool have_particles = false, have_params = false, bounds_valid = false, debug = false;

  simulation_parameters params{};
  precomputed_kernel_values terms{};
  SphConst konst{};

  StateArrays state[2]{};
  int cur = 0;  // which side holds the current particles
  float4* aux = nullptr;
  float4* accel = nullptr;
  uint32_t* skey = nullptr;
  uint32_t* perm = nullptr;
  SortBuffers sort{};
  uint32_t* cell_start = nullptr;
  uint32_t* cell_end = nullptr;
  GridState* grid = nullptr;
  BoundsAcc* bounds = nullptr;
  void* aos_stage = nullptr;

  Face* faces = nullptr;
  uint32_t face_count = 0;
  // face grid (option "face_grid"): cell lists of the scene's triangles for the collision pass
  bool use_face_grid = true;   // option "face_grid"
  FaceGrid face_grid{};
  uint32_t* fg_cell_start = nullptr;
  uint32_t* fg_ids = nullptr;
  uint32_t* fg_global = nullptr;
  std::vector<float> scene_vertices;     // host copies, to (re)build the grid when the option changes
  std::vector<uint32_t> scene_indices;

  // sub-cell order (subgrid.cu): arrays sorted by (cell key << 3 | octant); rrank = index of each
  // particle in the reference's array, rr_tmp = the gathered ranks of the previous sub-step
  bool sub_order = true;        // option "sub_cell_order"
  bool merged_rows = true;      // k_density_sub<.., kMerged>: the two index ranges of a sub-cell row in one loop
  int pair_density = -1;        // k_density_pairs: two particles of a sub-cell per thread, packed fp32 (option "pair_density"); -1: by size
  int factored_forces = 1;          // option "factored_forces": pair terms with the constants factored out of the sums (default) or add_pair_fast
  int pair_variant = 5;             // option "pair_variant" (tuning): walk 0/1/2 + 3 x (entries stored two at a time)
  int count_sort = 1;               // option "count_sort": counting sort on the dense sub-cell table instead of radix passes (sort.cu); 0 never, 1 by k_grid_setup's rule, 2 whenever the grid fits the table
  uint32_t* scan_state = nullptr;   // kScanStateWords words: chunk ticket and look-back words of the table scan
  uint32_t* pair_items = nullptr;   // [capacity] items written by k_reorder_sub
  uint32_t* pair_count = nullptr;
  bool forces_dense = true;     // k_forces_lists<.., 4>: four resident CTAs per SM (option forces_blocks = 4)
  bool fast_pairs = true;       // k_forces_lists<true, ..>: add_pair_fast (option fast_pairs)
  uint32_t sub_capacity = 0;   // cells the dense sub-cell table holds (9 words each)
  uint32_t* sub_lb = nullptr;
  uint32_t* rrank = nullptr;
  uint32_t* rr_tmp = nullptr;
  // list mode (default): the density pass stores neighbour lists, the force pass reads them
  bool use_lists = true;
  uint32_t list_rows_override = 0;  // 0 = derive from the rest density
  NeighbourLists lists{};
  size_t list_words = 0;            // allocated words of lists.entries

  // multi-GPU slab decomposition (dist.cu); pid = persistent particle ids, ping-pong like the state
  DistState dist{};
  uint32_t* pid[2] = {nullptr, nullptr};
  uint32_t* export_ids = nullptr;
  // order keys across ranks (sub-cell order only): (cell key, rank in cell) of the previous sub-step,
  // ping-pong like pid, and the rank in the cell after the current one
  uint32_t* ordk[2] = {nullptr, nullptr};
  uint32_t* ordr[2] = {nullptr, nullptr};
  uint32_t* wrank = nullptr;
  uint32_t* live_idx = nullptr;  // exchange in place (sub-cell order): indices of the sort's input, see k_dist_select

  DebugTaps taps{};
  uint32_t* ref_table = nullptr;
  size_t ref_table_words = 0;

  bool profiling = false;
  std::vector<cudaEvent_t> event_pool;
  size_t events_used = 0;
  clsph_stage_times times{};
  uint64_t launches = 0;

  std::string error;
};

namespace {

int fail(clsph_context* ctx, int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  if (ctx) ctx->error = buf; else g_create_error = buf;
  return code;
}

#define CLSPH_CUDA_TRY(ctx, expr)                                                                    \
  do {                                                                                               \
    cudaError_t e__ = (expr);                                                                        \
    if (e__ != cudaSuccess)                                                                          \
      return fail(ctx, e__ == cudaErrorMemoryAllocation ? CLSPH_ENOMEM : CLSPH_ECUDA, "%s:%d: %s -> %s", \
                  __FILE__, __LINE__, #expr, cudaGetErrorString(e__));                               \
  } while (0)

template <typename T>
cudaError_t dev_alloc(T** p, size_t count) {
  return cudaMalloc(reinterpret_cast<void**>(p), std::max<size_t>(count, 1) * sizeof(T));
}

// Smallest s with sqrtf(s) / h >= 1: the reference's window 1 - clamp(floor(r/h), 0, 1) is 1
// exactly when s = |d|^2 is below it (sqrt and divide are monotone), smoothing.cl:2.
float support_threshold(float h) {
  uint32_t lo = 0u, hi = 0x7f800000u;
  while (hi - lo > 1u) {
    const uint32_t mid = lo + (hi - lo) / 2u;
    float s;
    std::memcpy(&s, &mid, 4);
    volatile float r = sqrtf(s);
    volatile float q = r / h;
    if (q >= 1.0f) hi = mid; else lo = mid;
  }
  float s;
  std::memcpy(&s, &hi, 4);
  return s;
}

// Smallest s with sqrtf(s) >= 1e-7f: smoothing.cl:23's `length(r) < 0.0000001f` as a test on s.
float degenerate_threshold() {
  uint32_t lo = 0u, hi = 0x7f800000u;
  while (hi - lo > 1u) {
    const uint32_t mid = lo + (hi - lo) / 2u;
    float s;
    std::memcpy(&s, &mid, 4);
    volatile float r = sqrtf(s);
    if (r >= 0.0000001f) hi = mid; else lo = mid;
  }
  float s;
  std::memcpy(&s, &hi, 4);
  return s;
}

void derive_constants(clsph_context* ctx) {
  const simulation_parameters& p = ctx->params;
  const precomputed_kernel_values& t = ctx->terms;
  SphConst& c = ctx->konst;
  c.h = p.h;
  volatile float h2 = p.h * p.h;
  c.h2 = h2;
  c.support_s = support_threshold(p.h);
  volatile float hm = p.h * 1.0009765625f;  // h (1 + 2^-10), see sub_bounds in subgrid.cu
  c.h_margin = hm;
  c.mass = p.particle_mass;
  c.rho0 = p.fluid_density;
  c.K = p.K;
  c.c_poly6 = t.poly_6;
  c.c_spiky = t.spiky;
  c.c_visc = t.viscosity;
  c.c_poly6_grad = t.poly_6_gradient;
  c.c_poly6_lap = t.poly_6_laplacian;
  c.mu = p.dynamic_viscosity;
  c.sigma = p.surface_tension;
  c.tension_threshold = p.surface_tension_threshold;
  c.gx = p.constant_acceleration.s[0];
  c.gy = p.constant_acceleration.s[1];
  c.gz = p.constant_acceleration.s[2];
  volatile float dt = p.time_delta * p.simulation_scale;  // sph.cl:76
  c.dt = dt;
  c.vmax = p.max_velocity;
  c.restitution = p.restitution;
  volatile float h6 = p.h;  // pown(h, 6) by sequential fp32 multiplies, smoothing.cl:24
  for (int k = 1; k < 6; ++k) h6 = h6 * p.h;
  c.spiky_degenerate = -45.f / (float)(3.14159265358979323846 * (double)h6);
  c.degenerate_s = degenerate_threshold();
}

cudaEvent_t next_event(clsph_context* ctx) {
  if (ctx->events_used == ctx->event_pool.size()) {
    cudaEvent_t e;
    cudaEventCreate(&e);
    ctx->event_pool.push_back(e);
  }
  cudaEvent_t e = ctx->event_pool[ctx->events_used++];
  cudaEventRecord(e, ctx->stream);
  return e;
}

// Folds the recorded stage boundaries into ctx->times. Caller has synchronised the stream.
void drain_events(clsph_context* ctx) {
  const size_t per_step = kStages + 1;
  for (size_t base = 0; base + per_step <= ctx->events_used; base += per_step) {
    double ms[kStages];
    for (int s = 0; s < kStages; ++s) {
      float f = 0.f;
      cudaEventElapsedTime(&f, ctx->event_pool[base + s], ctx->event_pool[base + s + 1]);
      ms[s] = f;
    }
    ctx->times.ms_bounds_grid += ms[kStBounds];
    ctx->times.ms_exchange += ms[kStExchange];
    ctx->times.ms_keys += ms[kStKeys];
    ctx->times.ms_sort += ms[kStSort];
    ctx->times.ms_reorder += ms[kStReorder];
    ctx->times.ms_density += ms[kStDensity];
    ctx->times.ms_forces += ms[kStForces];
    ctx->times.ms_integrate += ms[kStIntegrate];
    ctx->times.substeps += 1;
  }
  ctx->events_used = 0;
}

int check_device_flags(clsph_context* ctx) {
  GridState g;
  CLSPH_CUDA_TRY(ctx, cudaMemcpyAsync(&g, ctx->grid, sizeof(g), cudaMemcpyDeviceToHost, ctx->stream));
  CLSPH_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  if (g.error & 8u) {
    CLSPH_CUDA_TRY(ctx, cudaMemsetAsync(&ctx->grid->error, 0, sizeof(uint32_t), ctx->stream));
    return fail(ctx, CLSPH_ECOMM, "multi-GPU exchange: a neighbouring rank did not deliver its AABB or its particles in time "
                "(peer transport; did every rank call clsph_step the same number of times?)");
  }
  if (g.error & 2u) {
    CLSPH_CUDA_TRY(ctx, cudaMemsetAsync(&ctx->grid->error, 0, sizeof(uint32_t), ctx->stream));
    return fail(ctx, CLSPH_ECOMM, "multi-GPU buffer overflow: more local, migrating or ghost particles than the capacities given "
                "to clsph_create / clsph_dist_init (local %u, emigrants %u, ghosts %u per message)", ctx->capacity,
                ctx->dist.emax, ctx->dist.gmax);
  }
  if (g.error & 1u) {
    CLSPH_CUDA_TRY(ctx, cudaMemsetAsync(&ctx->grid->error, 0, sizeof(uint32_t), ctx->stream));
    return fail(ctx, CLSPH_EGRID, "grid overflow: %d x %d x %d cells, each axis must stay below 1024 "
                "(10-bit Morton code; the reference asserts at sph_simulation.cpp:247-249)", g.gx, g.gy, g.gz);
  }
  if (g.error & 4u) {
    CLSPH_CUDA_TRY(ctx, cudaMemsetAsync(&ctx->grid->error, 0, sizeof(uint32_t), ctx->stream));
    return fail(ctx, CLSPH_EGRID, "grid of %d x %d x %d cells is too large for the sub-cell order (its keys are the Morton cell key "
                "plus three bits, in 32: the z axis must stay below 512 cells); set the option sub_cell_order to 0 for this domain",
                g.gx, g.gy, g.gz);
  }
  return CLSPH_OK;
}

int ensure_debug_buffers(clsph_context* ctx) {
  if (ctx->taps.keys_input) return CLSPH_OK;
  CLSPH_CUDA_TRY(ctx, dev_alloc(&ctx->taps.keys_input, ctx->capacity));
  CLSPH_CUDA_TRY(ctx, dev_alloc(&ctx->taps.candidate_count, ctx->capacity));
  CLSPH_CUDA_TRY(ctx, dev_alloc(&ctx->taps.support_count, ctx->capacity));
  CLSPH_CUDA_TRY(ctx, dev_alloc(&ctx->taps.collision_iters, ctx->capacity));
  return CLSPH_OK;
}

// Rows of the neighbour-list array: about 2.5x the neighbour count at rest density
// (rho0 / m particles per unit volume times the support sphere), rounded up to 8, in [32, 256].
uint32_t list_rows_for(const clsph_context* ctx) {
  if (ctx->list_rows_override) return ctx->list_rows_override;
  const simulation_parameters& p = ctx->params;
  const double n_rest = (double)p.fluid_density / (double)p.particle_mass * 4.18879020478639 * (double)p.h * p.h * p.h;
  double rows = 2.5 * n_rest + 8.0;
  if (!(rows >= 32.0)) rows = 32.0;
  if (rows > 256.0) rows = 256.0;
  return ((uint32_t)rows + 7u) & ~7u;
}

// (Re)builds the face grid of the current scene, or drops it when the option is off.
int refresh_face_grid(clsph_context* ctx) {
  CLSPH_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  cudaFree(ctx->fg_cell_start);
  cudaFree(ctx->fg_ids);
  cudaFree(ctx->fg_global);
  ctx->fg_cell_start = ctx->fg_ids = ctx->fg_global = nullptr;
  ctx->face_grid = FaceGrid{};
  if (!ctx->use_face_grid || ctx->face_count == 0) return CLSPH_OK;
  std::vector<uint32_t> cell_start, ids, global_ids;
  FaceGrid g;
  build_face_grid(ctx->scene_vertices.data(), ctx->scene_indices.data(), ctx->face_count, &g, &cell_start, &ids, &global_ids);
  if (g.nx == 0) return CLSPH_OK;
  CLSPH_CUDA_TRY(ctx, dev_alloc(&ctx->fg_cell_start, cell_start.size()));
  CLSPH_CUDA_TRY(ctx, dev_alloc(&ctx->fg_ids, ids.size()));
  CLSPH_CUDA_TRY(ctx, dev_alloc(&ctx->fg_global, global_ids.size()));
  CLSPH_CUDA_TRY(ctx, cudaMemcpy(ctx->fg_cell_start, cell_start.data(), sizeof(uint32_t) * cell_start.size(), cudaMemcpyHostToDevice));
  if (!ids.empty())
    CLSPH_CUDA_TRY(ctx, cudaMemcpy(ctx->fg_ids, ids.data(), sizeof(uint32_t) * ids.size(), cudaMemcpyHostToDevice));
  if (!global_ids.empty())
    CLSPH_CUDA_TRY(ctx, cudaMemcpy(ctx->fg_global, global_ids.data(), sizeof(uint32_t) * global_ids.size(), cudaMemcpyHostToDevice));
  g.cell_start = ctx->fg_cell_start;
  g.ids = ctx->fg_ids;
  g.global_ids = ctx->fg_global;
  ctx->face_grid = g;
  return CLSPH_OK;
}

// Arrays of the sub-cell order, allocated the first time it is selected.
int ensure_sub(clsph_context* ctx) {
  if (!ctx->sub_order || ctx->sub_lb) return CLSPH_OK;
  CLSPH_CUDA_TRY(ctx, dev_alloc(&ctx->sub_lb, (size_t)ctx->sub_capacity * 9u));
  CLSPH_CUDA_TRY(ctx, dev_alloc(&ctx->scan_state, scan_state_words(ctx->sub_capacity)));
  CLSPH_CUDA_TRY(ctx, cudaMemsetAsync(ctx->scan_state, 0, sizeof(uint32_t) * scan_state_words(ctx->sub_capacity), ctx->stream));
  CLSPH_CUDA_TRY(ctx, dev_alloc(&ctx->rrank, ctx->capacity));
  CLSPH_CUDA_TRY(ctx, dev_alloc(&ctx->rr_tmp, ctx->capacity));
  CLSPH_CUDA_TRY(ctx, dev_alloc(&ctx->pair_items, ctx->capacity));
  CLSPH_CUDA_TRY(ctx, dev_alloc(&ctx->pair_count, 2));  // [0] items, [1] lists that overflowed in the density pass
  CLSPH_CUDA_TRY(ctx, cudaMemsetAsync(ctx->pair_count, 0, 2 * sizeof(uint32_t), ctx->stream));
  CLSPH_CUDA_TRY(ctx, cudaMemsetAsync(ctx->sub_lb, 0, sizeof(uint32_t) * 9u * (size_t)ctx->sub_capacity, ctx->stream));
  return CLSPH_OK;
}

int ensure_lists(clsph_context* ctx) {
  if (!ctx->use_lists && !ctx->sub_order) {
    ctx->lists.rows = 0;
    return CLSPH_OK;
  }
  const uint32_t rows = list_rows_for(ctx);
  const size_t words = (size_t)rows * ctx->capacity;
  if (words > ctx->list_words) {
    CLSPH_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    cudaFree(ctx->lists.entries);
    ctx->lists.entries = nullptr;
    ctx->list_words = 0;
    CLSPH_CUDA_TRY(ctx, dev_alloc(&ctx->lists.entries, words));
    ctx->list_words = words;
  }
  if (!ctx->lists.count) {
    CLSPH_CUDA_TRY(ctx, dev_alloc(&ctx->lists.window_counter, 1));
    CLSPH_CUDA_TRY(ctx, dev_alloc(&ctx->lists.count, ctx->capacity));
    CLSPH_CUDA_TRY(ctx, cudaMemsetAsync(ctx->lists.count, 0, sizeof(uint32_t) * ctx->capacity, ctx->stream));
  }
  ctx->lists.rows = rows;
  return CLSPH_OK;
}

// One sub-step, enqueued on ctx->stream. Multi-GPU differences: the AABB is all-reduced, the
// particle count lives on the device (launches are sized by the capacity), and a migration +
// ghost exchange (dist.cu) assembles the unsorted local array in the other ping-pong side before
// the sort brings it back, so `cur` does not flip.
int enqueue_substep(clsph_context* ctx) {
  cudaStream_t st = ctx->stream;
  uint64_t* lc = &ctx->launches;
  const bool multi = ctx->dist.active;
  const uint32_t n = multi ? ctx->capacity : ctx->n;  // upper bound used to size launches
  const bool prof = ctx->profiling;
  const float inf = std::numeric_limits<float>::infinity();

  if (prof) next_event(ctx);
  if (!ctx->bounds_valid) {  // first step after an upload; later steps get the AABB from the integrator
    launch_bounds_reset(ctx->bounds, st, lc);
    launch_bounds(ctx->state[ctx->cur].pos, ctx->n, ctx->bounds, ctx->sm_count, st, lc);
    ctx->bounds_valid = true;
  }
  if (multi && dist_reduce_bounds(&ctx->dist, ctx->bounds, ctx->grid, st, lc)) return fail(ctx, CLSPH_ECOMM, "%s", dist_last_error());
  const bool sub = ctx->sub_order;
  StepZero zero{};
  if (sub) {
    zero.sub_lb = ctx->sub_lb;
    zero.keys_a = ctx->sort.keys_a;
    zero.keys_b = ctx->sort.keys_b;
    zero.scan_state = ctx->scan_state;
    zero.scan_words = scan_state_words(ctx->sub_capacity);
    zero.sort_scratch = ctx->sort.scratch;
    zero.sort_scratch_words = sort_scratch_zero_words(n);
    zero.pair_count = ctx->pair_count;
  }
  launch_grid_setup(ctx->bounds, ctx->grid, ctx->params.h, ctx->n, ctx->cell_capacity, multi ? ctx->dist.plane_lo : -inf,
                    multi ? ctx->dist.plane_hi : inf, multi, sub ? 1u : 0u, ctx->sub_capacity, (uint32_t)ctx->count_sort, zero, ctx->sm_count, st, lc);
  if (prof) next_event(ctx);

  const bool in_place = multi && sub && ctx->live_idx != nullptr;  // the owned particles stay where they are
  if (multi) {
    if (in_place) {
      if (dist_exchange(&ctx->dist, ctx->state[ctx->cur], ctx->pid[ctx->cur], ctx->skey, ctx->wrank, ctx->grid, ctx->state[ctx->cur],
                        ctx->pid[ctx->cur], ctx->ordk[ctx->cur], ctx->ordr[ctx->cur], ctx->capacity, ctx->live_idx, st, lc))
        return fail(ctx, CLSPH_ECOMM, "%s", dist_last_error());
    } else {
      if (dist_exchange(&ctx->dist, ctx->state[ctx->cur], ctx->pid[ctx->cur], ctx->skey, sub ? ctx->wrank : nullptr, ctx->grid,
                        ctx->state[ctx->cur ^ 1], ctx->pid[ctx->cur ^ 1], sub ? ctx->ordk[ctx->cur ^ 1] : nullptr,
                        sub ? ctx->ordr[ctx->cur ^ 1] : nullptr, ctx->capacity, nullptr, st, lc))
        return fail(ctx, CLSPH_ECOMM, "%s", dist_last_error());
      ctx->cur ^= 1;  // the unsorted local array is the source of this step's sort
    }
  }
  StateArrays& src = ctx->state[ctx->cur];
  StateArrays& dst = ctx->state[ctx->cur ^ 1];
  if (prof) next_event(ctx);

  // sub-cell order: the pre-step keys tap is written by k_rank, in the reference's order
  launch_sort_keys(ctx->sort, src.pos, ctx->grid, n, ctx->sm_count, (ctx->debug && !sub) ? ctx->taps.keys_input : nullptr, sub,
                   sub ? ctx->sub_lb : nullptr, in_place ? ctx->live_idx : nullptr, st, lc);  // (+ table clear, histogram scan)
  if (prof) next_event(ctx);
  if (sub) launch_scan_table(ctx->sub_lb, ctx->grid, ctx->scan_state, ctx->sub_capacity, ctx->sm_count, st, lc);  // (counting sort only)
  launch_sort_passes(ctx->sort, ctx->grid, n, in_place ? ctx->live_idx : nullptr, sub ? ctx->sub_lb : nullptr, st, lc);
  if (prof) next_event(ctx);

  bool join_side = false;  // the side stream has work of this sub-step
  // the list force kernel with factored pair terms on the lists of the per-particle / pair density kernels, which
  // contain the particle itself (option "factored_forces")
  // (with the shared-memory tile kernels the factored terms only paid for ~45 neighbours per particle, profiles/r02_n_*;
  // in the direct kernel, the default, they win for both fluids, profiles/r02_s_*)
  const bool want_factored = ctx->factored_forces != 0;
  const bool factored = sub && want_factored && ctx->fast_pairs;
  // Two particles per thread halve the threads: below ~150 k particles the GPU is not full either way and the pass lasts
  // as long as one thread's walk, which is shorter with one particle per thread (100 k: 0.113 against 0.117 ms per
  // sub-step; 256 k: 0.175 against 0.171, profiles/r02_as_*). Same results, bit for bit.
  const bool pairs = sub && (ctx->pair_density < 0 ? n >= 160000u : ctx->pair_density != 0);
  if (sub) {
    launch_reorder_sub(src, dst, ctx->sort, ctx->skey, multi ? nullptr : ctx->rrank, ctx->rr_tmp, ctx->sub_lb, ctx->grid,
                       multi ? ctx->pid[ctx->cur] : nullptr, multi ? ctx->pid[ctx->cur ^ 1] : nullptr,
                       multi ? ctx->ordk[ctx->cur] : nullptr, multi ? ctx->ordr[ctx->cur] : nullptr,
                       multi ? ctx->ordk[ctx->cur ^ 1] : nullptr, multi ? ctx->ordr[ctx->cur ^ 1] : nullptr,
                       pairs ? ctx->pair_items : nullptr, ctx->pair_count, n, st, lc);
    ctx->cur ^= 1;
    if (!multi) {  // one GPU: absolute index in the reference's array
      // Nothing in this sub-step reads the new ranks, and the pass (a count over the ~40 particles of each cell,
      // latency bound) fits into the issue slots the density pass leaves free: side stream, joined at the end.
      const bool rank_on_side = !ctx->debug && !prof;
      join_side = rank_on_side;
      cudaStream_t rs = st;
      if (rank_on_side) {
        CLSPH_CUDA_TRY(ctx, cudaEventRecord(ctx->ev_fork, st));
        CLSPH_CUDA_TRY(ctx, cudaStreamWaitEvent(ctx->side, ctx->ev_fork, 0));
        rs = ctx->side;
      }
      launch_rank(ctx->skey, ctx->rr_tmp, ctx->rrank, ctx->sub_lb, ctx->sort, ctx->grid, ctx->perm,
                  ctx->debug ? ctx->taps.keys_input : nullptr, n, rs, lc);
      if (rank_on_side) CLSPH_CUDA_TRY(ctx, cudaEventRecord(ctx->ev_join, ctx->side));
    } else         // across ranks: (cell key, rank in cell), merged at export
      launch_rank_pair(dst.pos, ctx->skey, ctx->ordk[ctx->cur], ctx->ordr[ctx->cur], ctx->wrank, ctx->sub_lb, ctx->sort, ctx->grid, n,
                       st, lc);
    if (prof) next_event(ctx);
    {
      if (pairs)
        launch_density_pairs(dst.pos, dst.vel, ctx->skey, ctx->sub_lb, ctx->sort, ctx->grid, ctx->konst, ctx->aux, ctx->lists,
                             ctx->taps, ctx->debug, ctx->pair_variant, ctx->pair_items, ctx->pair_count, n, st, lc);
      else
        launch_density_sub(dst.pos, dst.vel, ctx->skey, ctx->sub_lb, ctx->sort, ctx->grid, ctx->konst, ctx->aux, ctx->lists,
                           ctx->taps, ctx->debug, ctx->merged_rows, n, st, lc);
      if (prof) next_event(ctx);
      launch_forces(dst.pos, dst.vel, ctx->aux, ctx->skey, ctx->cell_start, ctx->cell_end, ctx->grid, ctx->konst, ctx->lists,
                    false, ctx->fast_pairs, ctx->forces_dense, ctx->accel, n, st, lc, factored);
      launch_forces_sub_overflow(dst.pos, dst.vel, ctx->aux, ctx->skey, ctx->sub_lb, ctx->sort, ctx->grid, ctx->konst,
                                 ctx->lists, ctx->accel, pairs ? ctx->pair_count + 1 : nullptr, n, st, lc);
    }
    if (prof) next_event(ctx);
  } else {
    launch_clear_cells(ctx->cell_start, ctx->cell_end, ctx->grid, ctx->cell_capacity, ctx->sm_count, st, lc);
    launch_reorder(src, dst, ctx->sort, ctx->skey, ctx->perm, ctx->cell_start, ctx->cell_end, ctx->grid,
                   multi ? ctx->pid[ctx->cur] : nullptr, multi ? ctx->pid[ctx->cur ^ 1] : nullptr, n, st, lc);
    ctx->cur ^= 1;
    if (prof) next_event(ctx);

    launch_density(dst.pos, dst.vel, ctx->skey, ctx->cell_start, ctx->cell_end, ctx->grid, ctx->konst, ctx->aux, ctx->lists,
                   ctx->taps, ctx->debug, n, ctx->sm_count, st, lc);
    if (prof) next_event(ctx);
    launch_forces(dst.pos, dst.vel, ctx->aux, ctx->skey, ctx->cell_start, ctx->cell_end, ctx->grid, ctx->konst,
                  ctx->lists, true, ctx->fast_pairs, ctx->forces_dense, ctx->accel, n, st, lc);
    if (prof) next_event(ctx);
  }
  if (ctx->debug)  // the integrator consumes the acceleration; keep a copy for the tap
    CLSPH_CUDA_TRY(ctx, cudaMemcpyAsync(ctx->taps.acceleration, ctx->accel, sizeof(float4) * n,
                                        cudaMemcpyDeviceToDevice, st));
  launch_integrate(dst, ctx->accel, ctx->skey, ctx->faces, ctx->face_count, ctx->face_grid, ctx->grid, ctx->konst, ctx->bounds,
                   ctx->debug ? ctx->taps.collision_iters : nullptr, n, ctx->sm_count, st, lc);
  if (multi) dist_publish_bounds(&ctx->dist, ctx->bounds, st, lc);
  if (join_side) CLSPH_CUDA_TRY(ctx, cudaStreamWaitEvent(st, ctx->ev_join, 0));  // whatever follows sees the new ranks
  if (prof) next_event(ctx);
  CLSPH_CUDA_TRY(ctx, cudaGetLastError());
  return CLSPH_OK;
}

}  // namespace

// ================================================================================================
extern "C" {

int clsph_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

const char* clsph_last_error(const clsph_context* ctx) { return ctx ? ctx->error.c_str() : g_create_error.c_str(); }

int clsph_create(clsph_context** out, int device, uint32_t max_particles, uint32_t cell_table_capacity) {
  if (!out) return fail(nullptr, CLSPH_EINVAL, "clsph_create: out is null");
  *out = nullptr;
  if (max_particles < 128u || max_particles >= (1u << 30))
    return fail(nullptr, CLSPH_EINVAL, "clsph_create: max_particles must be in [128, 2^30), got %u", max_particles);
  int count = clsph_device_count();
  if (count <= 0) return fail(nullptr, CLSPH_ECUDA, "clsph_create: no CUDA device available (there is no CPU fallback)");
  if (device < 0 || device >= count) return fail(nullptr, CLSPH_EINVAL, "clsph_create: device %d out of range [0, %d)", device, count);

  clsph_context* ctx = new clsph_context();
  ctx->device = device;
  ctx->capacity = max_particles;
  // Morton-indexed tables are 1.1x-6.4x the cell count for compact fluids; 8 entries
  // per particle with a 4 Mi floor covers every BASELINE config; larger grids use the fallback.
  if (const char* env = std::getenv("CLSPH_NEIGHBOUR_LISTS")) ctx->use_lists = std::atoi(env) != 0;
  if (const char* env = std::getenv("CLSPH_SUB_CELL_ORDER")) ctx->sub_order = std::atoi(env) != 0;
  if (const char* env = std::getenv("CLSPH_FACE_GRID")) ctx->use_face_grid = std::atoi(env) != 0;
  // dense sub-cell table: 9 words per cell; two cells per particle covers every BASELINE config
  // (a spread-out 16 Mi river fills ~0.7 cells per particle), larger grids use binary search
  ctx->sub_capacity = cell_table_capacity ? cell_table_capacity
                                          : (uint32_t)std::min<uint64_t>(std::max<uint64_t>((uint64_t)max_particles * 2u, 1u << 20), 1u << 28);
  ctx->cell_capacity = cell_table_capacity ? cell_table_capacity
                                           : (uint32_t)std::min<uint64_t>(std::max<uint64_t>((uint64_t)max_particles * 8u, 1u << 22), 1u << 28);
#define CREATE_TRY(expr)                                                                                          \
  do {                                                                                                            \
    cudaError_t e__ = (expr);                                                                                     \
    if (e__ != cudaSuccess) {                                                                                     \
      int rc__ = fail(nullptr, e__ == cudaErrorMemoryAllocation ? CLSPH_ENOMEM : CLSPH_ECUDA, "clsph_create: %s -> %s", #expr, \
                      cudaGetErrorString(e__));                                                                   \
      clsph_destroy(ctx);                                                                                         \
      return rc__;                                                                                                \
    }                                                                                                             \
  } while (0)
  CREATE_TRY(cudaSetDevice(device));
  cudaDeviceProp prop;
  CREATE_TRY(cudaGetDeviceProperties(&prop, device));
  ctx->sm_count = prop.multiProcessorCount;
  CREATE_TRY(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
  CREATE_TRY(cudaStreamCreateWithFlags(&ctx->side, cudaStreamNonBlocking));
  CREATE_TRY(cudaStreamCreateWithFlags(&ctx->copy, cudaStreamNonBlocking));
  CREATE_TRY(cudaEventCreateWithFlags(&ctx->ev_packed, cudaEventDisableTiming));
  CREATE_TRY(cudaEventCreateWithFlags(&ctx->ev_copied, cudaEventDisableTiming));
  CREATE_TRY(cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming));
  CREATE_TRY(cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming));
  const size_t cap = max_particles;
  for (int s = 0; s < 2; ++s) {
    CREATE_TRY(dev_alloc(&ctx->state[s].pos, cap));
    CREATE_TRY(dev_alloc(&ctx->state[s].vel, cap));
    CREATE_TRY(dev_alloc(&ctx->state[s].ivel, cap));
  }
  CREATE_TRY(dev_alloc(&ctx->aux, cap));
  CREATE_TRY(dev_alloc(&ctx->accel, cap));
  CREATE_TRY(dev_alloc(&ctx->taps.acceleration, cap));
  CREATE_TRY(dev_alloc(&ctx->skey, cap));
  CREATE_TRY(dev_alloc(&ctx->perm, cap));
  CREATE_TRY(dev_alloc(&ctx->sort.keys_a, cap));
  CREATE_TRY(dev_alloc(&ctx->sort.keys_b, cap));
  CREATE_TRY(dev_alloc(&ctx->sort.vals_a, cap));
  CREATE_TRY(dev_alloc(&ctx->sort.vals_b, cap));
  CREATE_TRY(dev_alloc(&ctx->sort.scratch, sort_scratch_words(max_particles)));
  CREATE_TRY(dev_alloc(&ctx->cell_start, (size_t)ctx->cell_capacity));
  CREATE_TRY(dev_alloc(&ctx->cell_end, (size_t)ctx->cell_capacity));
  CREATE_TRY(dev_alloc(&ctx->grid, 1));
  CREATE_TRY(dev_alloc(&ctx->bounds, 1));
  CREATE_TRY(cudaMalloc(&ctx->aos_stage, cap * sizeof(particle)));
  {
    GridState g0;
    std::memset(&g0, 0, sizeof(g0));
    g0.own_hi = g0.prev_hi = 0x7fffffff;  // one GPU: every cell is owned
    CREATE_TRY(cudaMemcpy(ctx->grid, &g0, sizeof(g0), cudaMemcpyHostToDevice));
  }
  CREATE_TRY(cudaMemset(ctx->cell_start, 0, sizeof(uint32_t) * ctx->cell_capacity));
  CREATE_TRY(cudaMemset(ctx->cell_end, 0, sizeof(uint32_t) * ctx->cell_capacity));
  neighbors_init();
  CREATE_TRY(cudaGetLastError());
  if (ctx->sub_order && ensure_sub(ctx) != CLSPH_OK) {
    g_create_error = ctx->error;
    clsph_destroy(ctx);
    return CLSPH_ENOMEM;
  }
#undef CREATE_TRY
  // CLSPH_OPTIONS="name=value,name=value": clsph_set_option pairs applied to every new context, so that
  // an unchanged application -- or the whole GPU test-suite -- can be run on another kernel organisation
  if (const char* env = std::getenv("CLSPH_OPTIONS")) {
    std::string spec(env);
    size_t at = 0;
    while (at < spec.size()) {
      size_t end = spec.find(',', at);
      if (end == std::string::npos) end = spec.size();
      const std::string item = spec.substr(at, end - at);
      at = end + 1;
      if (item.empty()) continue;
      const size_t eq = item.find('=');
      int rc = eq == std::string::npos ? CLSPH_EINVAL : clsph_set_option(ctx, item.substr(0, eq).c_str(), std::atoll(item.c_str() + eq + 1));
      if (rc != CLSPH_OK) {
        const std::string why = eq == std::string::npos ? std::string("expected name=value") : ctx->error;
        clsph_destroy(ctx);
        return fail(nullptr, rc, "clsph_create: CLSPH_OPTIONS item \"%s\": %s", item.c_str(), why.c_str());
      }
    }
  }
  *out = ctx;
  return CLSPH_OK;
}

void clsph_destroy(clsph_context* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  if (ctx->stream) cudaStreamSynchronize(ctx->stream);
  for (int s = 0; s < 2; ++s) {
    cudaFree(ctx->state[s].pos);
    cudaFree(ctx->state[s].vel);
    cudaFree(ctx->state[s].ivel);
  }
  cudaFree(ctx->aux);
  cudaFree(ctx->accel);
  cudaFree(ctx->skey);
  cudaFree(ctx->perm);
  cudaFree(ctx->sort.keys_a);
  cudaFree(ctx->sort.keys_b);
  cudaFree(ctx->sort.vals_a);
  cudaFree(ctx->sort.vals_b);
  cudaFree(ctx->sort.scratch);
  cudaFree(ctx->cell_start);
  cudaFree(ctx->cell_end);
  cudaFree(ctx->grid);
  cudaFree(ctx->bounds);
  cudaFree(ctx->aos_stage);
  cudaFree(ctx->faces);
  cudaFree(ctx->taps.keys_input);
  cudaFree(ctx->taps.candidate_count);
  cudaFree(ctx->taps.support_count);
  cudaFree(ctx->taps.acceleration);
  cudaFree(ctx->taps.collision_iters);
  cudaFree(ctx->ref_table);
  dist_destroy(&ctx->dist);
  for (int s = 0; s < 2; ++s) {
    cudaFree(ctx->pid[s]);
    cudaFree(ctx->ordk[s]);
    cudaFree(ctx->ordr[s]);
  }
  cudaFree(ctx->wrank);
  cudaFree(ctx->live_idx);
  cudaFree(ctx->export_ids);
  cudaFree(ctx->lists.entries);
  cudaFree(ctx->lists.count);
  cudaFree(ctx->lists.window_counter);
  cudaFree(ctx->fg_cell_start);
  cudaFree(ctx->fg_ids);
  cudaFree(ctx->fg_global);
  cudaFree(ctx->sub_lb);
  cudaFree(ctx->scan_state);
  cudaFree(ctx->rrank);
  cudaFree(ctx->rr_tmp);
  cudaFree(ctx->pair_items);
  cudaFree(ctx->pair_count);
  for (cudaEvent_t e : ctx->event_pool) cudaEventDestroy(e);
  if (ctx->copy) cudaStreamSynchronize(ctx->copy);
  if (ctx->ev_packed) cudaEventDestroy(ctx->ev_packed);
  if (ctx->ev_copied) cudaEventDestroy(ctx->ev_copied);
  if (ctx->copy) cudaStreamDestroy(ctx->copy);
  if (ctx->side) cudaStreamSynchronize(ctx->side);
  if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
  if (ctx->ev_join) cudaEventDestroy(ctx->ev_join);
  if (ctx->side) cudaStreamDestroy(ctx->side);
  if (ctx->stream) cudaStreamDestroy(ctx->stream);
  cudaGetLastError();
  delete ctx;
}

int clsph_set_scene(clsph_context* ctx, const float* face_normals, const float* vertices, size_t n_vertex_floats,
                    const uint32_t* indices, uint32_t face_count) {
  if (!ctx) return CLSPH_EINVAL;
  if (face_count && (!face_normals || !vertices || !indices))
    return fail(ctx, CLSPH_EINVAL, "clsph_set_scene: null array with face_count = %u", face_count);
  for (size_t k = 0; k < (size_t)face_count * 3; ++k)
    if ((size_t)indices[k] * 3 + 2 >= n_vertex_floats)
      return fail(ctx, CLSPH_EINVAL, "clsph_set_scene: index %u out of range (%zu vertex floats)", indices[k], n_vertex_floats);
  CLSPH_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  CLSPH_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  cudaFree(ctx->faces);
  ctx->faces = nullptr;
  ctx->face_count = face_count;
  CLSPH_CUDA_TRY(ctx, dev_alloc(&ctx->faces, face_count));
  ctx->scene_vertices.assign(vertices, vertices + (face_count ? n_vertex_floats : 0));
  ctx->scene_indices.assign(indices, indices + (size_t)face_count * 3);
  if (int rc = refresh_face_grid(ctx)) return rc;
  if (face_count == 0) return CLSPH_OK;
  float *d_n = nullptr, *d_v = nullptr;
  uint32_t* d_i = nullptr;
  CLSPH_CUDA_TRY(ctx, dev_alloc(&d_n, (size_t)face_count * 3));
  CLSPH_CUDA_TRY(ctx, dev_alloc(&d_v, n_vertex_floats));
  CLSPH_CUDA_TRY(ctx, dev_alloc(&d_i, (size_t)face_count * 3));
  cudaMemcpyAsync(d_n, face_normals, sizeof(float) * face_count * 3, cudaMemcpyHostToDevice, ctx->stream);
  cudaMemcpyAsync(d_v, vertices, sizeof(float) * n_vertex_floats, cudaMemcpyHostToDevice, ctx->stream);
  cudaMemcpyAsync(d_i, indices, sizeof(uint32_t) * face_count * 3, cudaMemcpyHostToDevice, ctx->stream);
  launch_prepare_faces(d_n, d_v, d_i, face_count, ctx->faces, ctx->stream, &ctx->launches);
  cudaError_t e = cudaStreamSynchronize(ctx->stream);
  cudaFree(d_n);
  cudaFree(d_v);
  cudaFree(d_i);
  CLSPH_CUDA_TRY(ctx, e);
  CLSPH_CUDA_TRY(ctx, cudaGetLastError());
  return CLSPH_OK;
}

int clsph_set_parameters(clsph_context* ctx, const simulation_parameters* params, const precomputed_kernel_values* terms) {
  if (!ctx) return CLSPH_EINVAL;
  if (!params) return fail(ctx, CLSPH_EINVAL, "clsph_set_parameters: params is null");
  if (!terms && !ctx->have_params) return fail(ctx, CLSPH_EINVAL, "clsph_set_parameters: terms is null and none were set before");
  if (!(params->h > 0.f)) return fail(ctx, CLSPH_EINVAL, "clsph_set_parameters: h must be positive, got %g", (double)params->h);
  ctx->params = *params;
  if (terms) ctx->terms = *terms;
  ctx->have_params = true;
  derive_constants(ctx);
  CLSPH_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  return ensure_lists(ctx);
}

int clsph_set_option(clsph_context* ctx, const char* name, long long value) {
  if (!ctx || !name) return CLSPH_EINVAL;
  CLSPH_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  if (!std::strcmp(name, "neighbour_lists")) {
    ctx->use_lists = value != 0;
  } else if (!std::strcmp(name, "sub_cell_order")) {
    // the two organisations keep the arrays in different orders: switch only while no particles are held
    if (ctx->have_particles && ctx->sub_order != (value != 0))
      return fail(ctx, CLSPH_ESTATE, "clsph_set_option: sub_cell_order must be set before particles are uploaded");
    ctx->sub_order = value != 0;
    int rc = ensure_sub(ctx);
    if (rc) return rc;
  } else if (!std::strcmp(name, "merged_rows")) {
    ctx->merged_rows = value != 0;
  } else if (!std::strcmp(name, "pair_density")) {
    ctx->pair_density = value < 0 ? -1 : (value != 0 ? 1 : 0);
  } else if (!std::strcmp(name, "factored_forces")) {
    ctx->factored_forces = value < 0 ? -1 : (value != 0 ? 1 : 0);
  } else if (!std::strcmp(name, "count_sort")) {
    if (value < 0 || value > 2) return fail(ctx, CLSPH_EINVAL, "clsph_set_option: count_sort must be 0, 1 or 2");
    ctx->count_sort = (int)value;
  } else if (!std::strcmp(name, "pair_variant")) {
    if (value < 0 || value > 5) return fail(ctx, CLSPH_EINVAL, "clsph_set_option: pair_variant must be in [0, 5]");
    ctx->pair_variant = (int)value;
  } else if (!std::strcmp(name, "fast_pairs")) {
    ctx->fast_pairs = value != 0;
  } else if (!std::strcmp(name, "forces_blocks")) {
    if (value != 3 && value != 4) return fail(ctx, CLSPH_EINVAL, "clsph_set_option: forces_blocks must be 3 or 4");
    ctx->forces_dense = value == 4;
  } else if (!std::strcmp(name, "face_grid")) {
    ctx->use_face_grid = value != 0;
    if (int rc = refresh_face_grid(ctx)) return rc;
  } else if (!std::strcmp(name, "list_rows")) {
    if (value < 0 || value > 1024) return fail(ctx, CLSPH_EINVAL, "clsph_set_option: list_rows must be in [0, 1024]");
    ctx->list_rows_override = ((uint32_t)value + 1u) & ~1u;  // even: k_density_pairs stores entries two at a time
  } else {
    return fail(ctx, CLSPH_EINVAL, "clsph_set_option: unknown option \"%s\"", name);
  }
  return ctx->have_params ? ensure_lists(ctx) : CLSPH_OK;
}

// step_follows: the caller enqueues a sub-step before anything reads density / pressure / key of the upload.
static int upload_particles(clsph_context* ctx, const particle* aos, uint32_t n, bool step_follows) {
  if (!ctx) return CLSPH_EINVAL;
  if (!aos) return fail(ctx, CLSPH_EINVAL, "clsph_upload_particles: aos is null");
  if (n < 128u) return fail(ctx, CLSPH_EINVAL, "clsph_upload_particles: need at least 128 particles, got %u (sort.cl:9-20)", n);
  if (n > ctx->capacity) return fail(ctx, CLSPH_EINVAL, "clsph_upload_particles: %u particles exceed the capacity %u", n, ctx->capacity);
  CLSPH_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  if (ctx->frame_pending) CLSPH_CUDA_TRY(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_copied, 0));  // the staging area is in use
  // Page-locked host memory and a sub-step right behind: with CLSPH_ZERO_COPY_UPLOAD=1 the kernel reads the 48 bytes per
  // record it needs straight over the host link instead of the copy engine moving all 80. OFF by default: measured on
  // a B200 (profiles/r02_o_*) it is SLOWER end to end -- 3.79 vs 3.68 ms per step at 1 Mi particles, 15.4 vs 15.0 at
  // 4 Mi: the copy engine streams at the link's rate, SM-issued reads of 48-byte runs do not.
  const void* mapped = nullptr;
  static const bool allow_zero_copy = [] { const char* e = std::getenv("CLSPH_ZERO_COPY_UPLOAD"); return e && std::atoi(e) != 0; }();
  if (step_follows && allow_zero_copy && !ctx->dist.active) {
    cudaPointerAttributes attr;
    if (cudaPointerGetAttributes(&attr, aos) == cudaSuccess && attr.type == cudaMemoryTypeHost && attr.devicePointer)
      mapped = attr.devicePointer;
    cudaGetLastError();  // (pageable memory makes older drivers report an error here)
  }
  if (mapped) {
    launch_aos_to_soa_host(mapped, ctx->state[ctx->cur], ctx->aux, ctx->skey, ctx->sub_order ? ctx->rrank : nullptr, n, ctx->stream,
                           &ctx->launches);
  } else {
    CLSPH_CUDA_TRY(ctx, cudaMemcpyAsync(ctx->aos_stage, aos, sizeof(particle) * (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
    launch_aos_to_soa(ctx->aos_stage, ctx->state[ctx->cur], ctx->aux, ctx->skey, nullptr, ctx->sub_order ? ctx->rrank : nullptr, n,
                      ctx->stream, &ctx->launches);
  }
  const uint32_t count_and_fresh[2] = {n, 1u};
  CLSPH_CUDA_TRY(ctx, cudaMemcpyAsync(&ctx->grid->n, &count_and_fresh[0], sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream));
  CLSPH_CUDA_TRY(ctx, cudaMemcpyAsync(&ctx->grid->fresh, &count_and_fresh[1], sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream));
  if (ctx->dist.active) launch_fill_ids(ctx->pid[ctx->cur], nullptr, 0u, n, ctx->stream, &ctx->launches);
  CLSPH_CUDA_TRY(ctx, cudaGetLastError());
  ctx->n = n;
  ctx->have_particles = true;
  ctx->bounds_valid = false;
  if (ctx->dist.active) dist_invalidate_bounds(&ctx->dist);
  return CLSPH_OK;
}

int clsph_upload_particles(clsph_context* ctx, const particle* aos, uint32_t n) { return upload_particles(ctx, aos, n, false); }

/* ---- multi-GPU ---------------------------------------------------------------------------- */

int clsph_comm_unique_id(void* out, size_t bytes) {
  if (!out) return CLSPH_EINVAL;
  if (dist_unique_id(out, bytes)) return fail(nullptr, CLSPH_ECOMM, "clsph_comm_unique_id: %s", dist_last_error());
  return CLSPH_OK;
}

int clsph_dist_init(clsph_context* ctx, int rank, int world, const void* unique_id, float plane_lo, float plane_hi,
                    uint32_t emigrant_capacity, uint32_t ghost_capacity) {
  if (!ctx || !unique_id) return CLSPH_EINVAL;
  if (world < 1 || rank < 0 || rank >= world) return fail(ctx, CLSPH_EINVAL, "clsph_dist_init: rank %d of %d", rank, world);
  if (ctx->dist.active) return fail(ctx, CLSPH_ESTATE, "clsph_dist_init: already initialised");
  if (!(plane_lo < plane_hi)) return fail(ctx, CLSPH_EINVAL, "clsph_dist_init: plane_lo must be below plane_hi");
  CLSPH_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  // A slab boundary is a fixed plane snapped to the nearest cell boundary of a grid whose origin follows
  // the fluid, so now and then it jumps by one cell and a whole cell layer changes owner in one
  // sub-step: the emigrant capacity must hold a layer, the ghost capacity two.
  if (dist_init(&ctx->dist, rank, world, unique_id, plane_lo, plane_hi, emigrant_capacity ? emigrant_capacity : ctx->capacity / 6 + 1024,
                ghost_capacity ? ghost_capacity : ctx->capacity / 3 + 1024))
    return fail(ctx, CLSPH_ECOMM, "clsph_dist_init: %s", dist_last_error());
  for (int s = 0; s < 2; ++s) {
    CLSPH_CUDA_TRY(ctx, dev_alloc(&ctx->pid[s], ctx->capacity));
    CLSPH_CUDA_TRY(ctx, dev_alloc(&ctx->ordk[s], ctx->capacity));
    CLSPH_CUDA_TRY(ctx, dev_alloc(&ctx->ordr[s], ctx->capacity));
  }
  CLSPH_CUDA_TRY(ctx, dev_alloc(&ctx->wrank, ctx->capacity));
  CLSPH_CUDA_TRY(ctx, dev_alloc(&ctx->export_ids, ctx->capacity));
  {  // CLSPH_DIST_IN_PLACE=0: the copying exchange (k_dist_classify) also in the sub-cell order
    const char* e = std::getenv("CLSPH_DIST_IN_PLACE");
    if (!(e && std::atoi(e) == 0)) CLSPH_CUDA_TRY(ctx, dev_alloc(&ctx->live_idx, ctx->capacity));
  }
  return CLSPH_OK;
}

const char* clsph_dist_transport(const clsph_context* ctx) { return ctx ? dist_transport(&ctx->dist) : "none"; }

int clsph_dist_upload(clsph_context* ctx, const particle* aos, const uint32_t* ids, uint32_t n) {
  if (!ctx) return CLSPH_EINVAL;
  if (!ctx->dist.active) return fail(ctx, CLSPH_ESTATE, "clsph_dist_upload: call clsph_dist_init first");
  if (!aos || !ids || n == 0 || n > ctx->capacity) return fail(ctx, CLSPH_EINVAL, "clsph_dist_upload: bad arguments (n = %u, capacity %u)", n, ctx->capacity);
  CLSPH_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  CLSPH_CUDA_TRY(ctx, cudaMemcpyAsync(ctx->aos_stage, aos, sizeof(particle) * (size_t)n, cudaMemcpyHostToDevice, st));
  launch_aos_to_soa(ctx->aos_stage, ctx->state[ctx->cur], ctx->aux, ctx->skey, nullptr, nullptr, n, st, &ctx->launches);
  CLSPH_CUDA_TRY(ctx, cudaMemcpyAsync(ctx->export_ids, ids, sizeof(uint32_t) * (size_t)n, cudaMemcpyHostToDevice, st));
  launch_fill_ids(ctx->pid[ctx->cur], ctx->export_ids, 0u, n, st, &ctx->launches);
  const uint32_t count_and_fresh[2] = {n, 1u};
  CLSPH_CUDA_TRY(ctx, cudaMemcpyAsync(&ctx->grid->n, &count_and_fresh[0], sizeof(uint32_t), cudaMemcpyHostToDevice, st));
  CLSPH_CUDA_TRY(ctx, cudaMemcpyAsync(&ctx->grid->fresh, &count_and_fresh[1], sizeof(uint32_t), cudaMemcpyHostToDevice, st));
  CLSPH_CUDA_TRY(ctx, cudaStreamSynchronize(st));
  ctx->n = n;
  ctx->have_particles = true;
  ctx->bounds_valid = false;
  if (ctx->dist.active) dist_invalidate_bounds(&ctx->dist);
  return CLSPH_OK;
}

int clsph_dist_download(clsph_context* ctx, particle* aos_out, uint32_t* ids_out, uint32_t capacity, uint32_t* n_out) {
  if (!ctx || !n_out) return CLSPH_EINVAL;
  if (!ctx->dist.active || !ctx->have_particles) return fail(ctx, CLSPH_ESTATE, "clsph_dist_download: nothing to download");
  CLSPH_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  uint32_t* count = ctx->dist.counters + 1;
  launch_dist_export(ctx->state[ctx->cur], ctx->aux, ctx->skey, ctx->pid[ctx->cur], ctx->sub_order ? ctx->wrank : nullptr, ctx->grid,
                     ctx->aos_stage, ctx->export_ids,
                     count, ctx->capacity, st, &ctx->launches);
  uint32_t n = 0;
  CLSPH_CUDA_TRY(ctx, cudaMemcpyAsync(&n, count, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
  CLSPH_CUDA_TRY(ctx, cudaStreamSynchronize(st));
  *n_out = n;
  if (aos_out && ids_out) {
    if (n > capacity) return fail(ctx, CLSPH_EINVAL, "clsph_dist_download: %u owned particles, room for %u", n, capacity);
    CLSPH_CUDA_TRY(ctx, cudaMemcpyAsync(aos_out, ctx->aos_stage, sizeof(particle) * (size_t)n, cudaMemcpyDeviceToHost, st));
    CLSPH_CUDA_TRY(ctx, cudaMemcpyAsync(ids_out, ctx->export_ids, sizeof(uint32_t) * (size_t)n, cudaMemcpyDeviceToHost, st));
  }
  return clsph_synchronize(ctx);
}

int clsph_step(clsph_context* ctx, uint32_t n_substeps) {
  if (!ctx) return CLSPH_EINVAL;
  if (!ctx->have_particles) return fail(ctx, CLSPH_ESTATE, "clsph_step: no particles uploaded");
  if (!ctx->have_params) return fail(ctx, CLSPH_ESTATE, "clsph_step: parameters not set");
  CLSPH_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  if (ctx->debug) {
    int rc = ensure_debug_buffers(ctx);
    if (rc) return rc;
  }
  for (uint32_t k = 0; k < n_substeps; ++k) {
    if (ctx->profiling && ctx->events_used + kStages + 1 > 4096) {  // bound the event pool
      CLSPH_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
      drain_events(ctx);
    }
    int rc = enqueue_substep(ctx);
    if (rc) return rc;
  }
  return CLSPH_OK;
}

int clsph_synchronize(clsph_context* ctx) {
  if (!ctx) return CLSPH_EINVAL;
  CLSPH_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  CLSPH_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  if (ctx->profiling) drain_events(ctx);
  return check_device_flags(ctx);
}

int clsph_get_parameters(clsph_context* ctx, simulation_parameters* out) {
  if (!ctx || !out) return CLSPH_EINVAL;
  CLSPH_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  GridState g;
  CLSPH_CUDA_TRY(ctx, cudaMemcpyAsync(&g, ctx->grid, sizeof(g), cudaMemcpyDeviceToHost, ctx->stream));
  CLSPH_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  *out = ctx->params;
  out->particles_count = ctx->n;
  out->grid_size_x = g.gx;
  out->grid_size_y = g.gy;
  out->grid_size_z = g.gz;
  out->grid_cell_count = g.cell_count;
  out->min_point.s[0] = g.min_x; out->min_point.s[1] = g.min_y; out->min_point.s[2] = g.min_z; out->min_point.s[3] = 0.f;
  out->max_point.s[0] = g.max_x; out->max_point.s[1] = g.max_y; out->max_point.s[2] = g.max_z; out->max_point.s[3] = 0.f;
  return CLSPH_OK;
}

int clsph_download_particles(clsph_context* ctx, particle* aos_out) {
  if (!ctx) return CLSPH_EINVAL;
  if (!aos_out) return fail(ctx, CLSPH_EINVAL, "clsph_download_particles: aos_out is null");
  if (!ctx->have_particles) return fail(ctx, CLSPH_ESTATE, "clsph_download_particles: no particles uploaded");
  CLSPH_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  if (ctx->frame_pending) CLSPH_CUDA_TRY(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_copied, 0));  // the staging area is in use
  launch_soa_to_aos(ctx->state[ctx->cur], ctx->aux, ctx->skey, (ctx->sub_order && !ctx->dist.active) ? ctx->rrank : nullptr,
                    ctx->aos_stage, ctx->n, ctx->stream, &ctx->launches);
  CLSPH_CUDA_TRY(ctx, cudaMemcpyAsync(aos_out, ctx->aos_stage, sizeof(particle) * (size_t)ctx->n, cudaMemcpyDeviceToHost, ctx->stream));
  return clsph_synchronize(ctx);
}

/* ---- frame export ------------------------------------------------------------------------- */

int clsph_frame_begin(clsph_context* ctx, float* points, uint32_t capacity) {
  if (!ctx || !points) return CLSPH_EINVAL;
  if (!ctx->have_particles) return fail(ctx, CLSPH_ESTATE, "clsph_frame_begin: no particles uploaded");
  if (ctx->dist.active) return fail(ctx, CLSPH_ESTATE, "clsph_frame_begin: single-GPU contexts only (use clsph_dist_download)");
  if (ctx->frame_pending) return fail(ctx, CLSPH_ESTATE, "clsph_frame_begin: the previous frame was not collected with clsph_frame_end");
  if (capacity < ctx->n) return fail(ctx, CLSPH_EINVAL, "clsph_frame_begin: %u particles, room for %u", ctx->n, capacity);
  CLSPH_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  // the 28-byte records are packed on the compute stream (the state must not move on before they are taken: 20 us
  // at a million particles) into the staging area; the copy to the host then runs on its own stream
  float* staged = static_cast<float*>(ctx->aos_stage);
  launch_pack_frame(ctx->state[ctx->cur], ctx->aux, ctx->sub_order ? ctx->rrank : nullptr, staged, ctx->n, ctx->stream, &ctx->launches);
  CLSPH_CUDA_TRY(ctx, cudaEventRecord(ctx->ev_packed, ctx->stream));
  CLSPH_CUDA_TRY(ctx, cudaStreamWaitEvent(ctx->copy, ctx->ev_packed, 0));
  CLSPH_CUDA_TRY(ctx, cudaMemcpyAsync(points, staged, sizeof(float) * 7u * (size_t)ctx->n, cudaMemcpyDeviceToHost, ctx->copy));
  CLSPH_CUDA_TRY(ctx, cudaEventRecord(ctx->ev_copied, ctx->copy));
  ctx->frame_pending = true;
  return CLSPH_OK;
}

int clsph_frame_end(clsph_context* ctx) {
  if (!ctx) return CLSPH_EINVAL;
  if (!ctx->frame_pending) return CLSPH_OK;
  CLSPH_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  CLSPH_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->copy));
  ctx->frame_pending = false;
  return CLSPH_OK;
}

int clsph_host_alloc(void** out, size_t bytes) {
  if (!out) return CLSPH_EINVAL;
  *out = nullptr;
  if (cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocDefault) != cudaSuccess) {
    cudaGetLastError();
    return fail(nullptr, CLSPH_ENOMEM, "clsph_host_alloc: %zu bytes of page-locked memory", bytes);
  }
  return CLSPH_OK;
}

void clsph_host_free(void* p) {
  if (p) cudaFreeHost(p);
  cudaGetLastError();
}

int clsph_simulate_single_frame(clsph_context* ctx, const particle* in, particle* out, simulation_parameters* params,
                                const precomputed_kernel_values* terms) {
  if (!ctx) return CLSPH_EINVAL;
  if (!in || !out || !params) return fail(ctx, CLSPH_EINVAL, "clsph_simulate_single_frame: null argument");
  int rc = clsph_set_parameters(ctx, params, terms);
  if (rc) return rc;
  if ((rc = upload_particles(ctx, in, params->particles_count, true))) return rc;
  if ((rc = clsph_step(ctx, 1))) return rc;
  if ((rc = clsph_download_particles(ctx, out))) return rc;
  return clsph_get_parameters(ctx, params);
}

int clsph_kernel_advection_collision(clsph_context* ctx, const particle* in, particle* out, uint32_t n) {
  if (!ctx) return CLSPH_EINVAL;
  if (!in || !out) return fail(ctx, CLSPH_EINVAL, "clsph_kernel_advection_collision: null argument");
  if (!ctx->have_params) return fail(ctx, CLSPH_ESTATE, "clsph_kernel_advection_collision: parameters not set");
  if (n == 0 || n > ctx->capacity) return fail(ctx, CLSPH_EINVAL, "clsph_kernel_advection_collision: n = %u out of range", n);
  CLSPH_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  CLSPH_CUDA_TRY(ctx, cudaMemcpyAsync(ctx->aos_stage, in, sizeof(particle) * (size_t)n, cudaMemcpyHostToDevice, st));
  launch_aos_to_soa(ctx->aos_stage, ctx->state[ctx->cur], ctx->aux, ctx->skey, ctx->accel, ctx->sub_order ? ctx->rrank : nullptr, n,
                    st, &ctx->launches);
  CLSPH_CUDA_TRY(ctx, cudaMemcpyAsync(&ctx->grid->n, &n, sizeof(uint32_t), cudaMemcpyHostToDevice, st));
  launch_bounds_reset(ctx->bounds, st, &ctx->launches);
  if (ctx->debug) {
    int rc = ensure_debug_buffers(ctx);
    if (rc) return rc;
  }
  launch_integrate(ctx->state[ctx->cur], ctx->accel, ctx->skey, ctx->faces, ctx->face_count, ctx->face_grid, ctx->grid, ctx->konst, ctx->bounds,
                   ctx->debug ? ctx->taps.collision_iters : nullptr, n, ctx->sm_count, st, &ctx->launches);
  ctx->n = n;
  ctx->have_particles = true;
  ctx->bounds_valid = false;
  return clsph_download_particles(ctx, out);
}

int clsph_set_debug(clsph_context* ctx, int enable) {
  if (!ctx) return CLSPH_EINVAL;
  ctx->debug = enable != 0;
  if (ctx->debug) {
    CLSPH_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    return ensure_debug_buffers(ctx);
  }
  return CLSPH_OK;
}

int clsph_debug_fetch(clsph_context* ctx, int what, void* dst, size_t bytes) {
  if (!ctx || !dst) return CLSPH_EINVAL;
  if (!ctx->have_particles) return fail(ctx, CLSPH_ESTATE, "clsph_debug_fetch: no particles uploaded");
  CLSPH_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  const size_t n = ctx->n;
  const void* src = nullptr;
  size_t need = 0;
  bool needs_debug = false;
  const bool scatter = ctx->sub_order && !ctx->dist.active;  // per-particle taps go through the reference rank
  switch (what) {
    case CLSPH_TAP_SORTED_KEYS: src = ctx->skey; need = 4 * n; break;
    case CLSPH_TAP_PERMUTATION: src = ctx->perm; need = 4 * n; break;
    case CLSPH_TAP_KEYS_INPUT: src = ctx->taps.keys_input; need = 4 * n; needs_debug = true; break;
    case CLSPH_TAP_CANDIDATE_COUNT: src = ctx->taps.candidate_count; need = 4 * n; needs_debug = true; break;
    case CLSPH_TAP_SUPPORT_COUNT: src = ctx->taps.support_count; need = 4 * n; needs_debug = true; break;
    case CLSPH_TAP_COLLISION_ITERS: src = ctx->taps.collision_iters; need = 4 * n; needs_debug = true; break;
    case CLSPH_TAP_CELL_TABLE: {
      simulation_parameters p;
      int rc = clsph_get_parameters(ctx, &p);
      if (rc) return rc;
      need = 4 * (size_t)p.grid_cell_count;
      if (bytes != need) return fail(ctx, CLSPH_EINVAL, "clsph_debug_fetch: cell table is %zu bytes, got %zu", need, bytes);
      if (ctx->ref_table_words < p.grid_cell_count) {
        cudaFree(ctx->ref_table);
        ctx->ref_table = nullptr;
        ctx->ref_table_words = 0;
        CLSPH_CUDA_TRY(ctx, dev_alloc(&ctx->ref_table, (size_t)p.grid_cell_count));
        ctx->ref_table_words = p.grid_cell_count;
      }
      launch_reference_cell_table(ctx->skey, ctx->grid, ctx->ref_table, ctx->n, ctx->stream, &ctx->launches);
      src = ctx->ref_table;
      break;
    }
    case CLSPH_TAP_DENSITY:
    case CLSPH_TAP_PRESSURE:
    case CLSPH_TAP_ACCELERATION: {
      const bool acc = what == CLSPH_TAP_ACCELERATION;
      if (acc && !ctx->debug) return fail(ctx, CLSPH_ESTATE, "clsph_debug_fetch: enable clsph_set_debug before the step");
      need = acc ? 12 * n : 4 * n;
      if (bytes != need) return fail(ctx, CLSPH_EINVAL, "clsph_debug_fetch: tap %d is %zu bytes, got %zu", what, need, bytes);
      std::vector<float> host(4 * n);
      const void* from = acc ? ctx->taps.acceleration : ctx->aux;
      if (scatter) {  // sub-cell order: bring the records into the reference's order first
        launch_scatter_words(from, ctx->rrank, ctx->aos_stage, (uint32_t)n, 4u, ctx->stream, &ctx->launches);
        from = ctx->aos_stage;
      }
      CLSPH_CUDA_TRY(ctx, cudaMemcpyAsync(host.data(), from, 16 * n, cudaMemcpyDeviceToHost, ctx->stream));
      CLSPH_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
      float* o = static_cast<float*>(dst);
      for (size_t i = 0; i < n; ++i) {
        if (acc) { o[3 * i] = host[4 * i]; o[3 * i + 1] = host[4 * i + 1]; o[3 * i + 2] = host[4 * i + 2]; }
        else o[i] = host[4 * i + (what == CLSPH_TAP_PRESSURE ? 1 : 0)];
      }
      return CLSPH_OK;
    }
    default:
      return fail(ctx, CLSPH_EINVAL, "clsph_debug_fetch: unknown tap %d", what);
  }
  if (needs_debug && (!ctx->debug || !src)) return fail(ctx, CLSPH_ESTATE, "clsph_debug_fetch: enable clsph_set_debug before the step");
  if (bytes != need) return fail(ctx, CLSPH_EINVAL, "clsph_debug_fetch: tap %d is %zu bytes, got %zu", what, need, bytes);
  if (scatter && (what == CLSPH_TAP_CANDIDATE_COUNT || what == CLSPH_TAP_SUPPORT_COUNT || what == CLSPH_TAP_COLLISION_ITERS)) {
    // recorded in the internal (sub-cell) order; the other integer taps are already in the reference's
    launch_scatter_words(src, ctx->rrank, ctx->aos_stage, (uint32_t)n, 1u, ctx->stream, &ctx->launches);
    src = ctx->aos_stage;
  }
  CLSPH_CUDA_TRY(ctx, cudaMemcpyAsync(dst, src, need, cudaMemcpyDeviceToHost, ctx->stream));
  CLSPH_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  CLSPH_CUDA_TRY(ctx, cudaGetLastError());
  return CLSPH_OK;
}

int clsph_profile_enable(clsph_context* ctx, int enable) {
  if (!ctx) return CLSPH_EINVAL;
  CLSPH_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  CLSPH_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  ctx->events_used = 0;
  ctx->times = clsph_stage_times{};
  ctx->launches = 0;
  ctx->profiling = enable != 0;
  return CLSPH_OK;
}

int clsph_profile_read(clsph_context* ctx, clsph_stage_times* out) {
  if (!ctx || !out) return CLSPH_EINVAL;
  CLSPH_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  CLSPH_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  drain_events(ctx);
  *out = ctx->times;
  out->kernel_launches = ctx->launches;
  return CLSPH_OK;
}

int clsph_particle_count(clsph_context* ctx, uint32_t* n) {
  if (!ctx || !n) return CLSPH_EINVAL;
  *n = ctx->n;
  return CLSPH_OK;
}

int clsph_sort_passes(clsph_context* ctx, uint32_t* passes) {
  if (!ctx || !passes) return CLSPH_EINVAL;
  CLSPH_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  CLSPH_CUDA_TRY(ctx, cudaMemcpyAsync(passes, &ctx->grid->sort_passes, sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
  CLSPH_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  return CLSPH_OK;
}

void* clsph_stream(clsph_context* ctx) { return ctx ? (void*)ctx->stream : nullptr; }

}  // extern "C"
