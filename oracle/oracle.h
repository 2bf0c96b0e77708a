/*
 * oracle.h -- CPU restatement of libclsph's per-step SPH hot path.
 *
 * TEST INFRASTRUCTURE ONLY. Nothing under oracle/ is part of the product: only tests/,
 * __graft_entry__.smoke() and bench.py's CPU-baseline / --impl reference legs may load it,
 * and only as the checker (or as the thing timed on the CPU side). The CUDA library never
 * links, loads or calls it.
 *
 * Parity status: PINNED against the reference's own kernel sources. oracle/_ref/ is built
 * from /root/reference/libclsph/kernels/ (every .cl file, compiled as C++ behind oracle/ref_shim/) and
 * tests/test_oracle_vs_ref.py + tests/golden/ hold its outputs; see DESIGN.md "Oracle".
 *
 * Arithmetic contract (the part of OpenCL the reference leaves to the runtime):
 *   - IEEE binary32 everywhere, round-to-nearest-even; `/` and sqrtf correctly rounded;
 *   - NO contraction of a*b+c written as separate operators (build with -ffp-contract=off);
 *   - the geometric built-ins are defined with fused multiply-adds, GPU style:
 *       dot(a,b)      = fmaf(a.z,b.z, fmaf(a.y,b.y, a.x*b.x))
 *       length(v)     = sqrtf(dot(v,v))        distance(a,b) = length(a-b)
 *       normalize(v)  = v / length(v)          (component-wise IEEE divide)
 *       pown(x,n)     = ((x*x)*x)...           (n-1 sequential multiplies)
 *   - loops in the reference's order: neighbour cells z,y,x (x innermost), candidates by
 *     ascending sorted index, faces 0..F-1.
 */
#ifndef CLSPH_ORACLE_H_
#define CLSPH_ORACLE_H_

#include "clsph/clsph_types.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Optional per-stage observation points. Any pointer may be NULL. Arrays indexed by
 * "sorted" use the post-sort particle order (= the order of the output array). */
typedef struct oracle_taps {
  uint32_t* keys_input_order; /* [N]   Morton cell key of in[i]                          */
  uint32_t* permutation;      /* [N]   out[r] came from in[permutation[r]]               */
  uint32_t* cell_table;       /* [grid_cell_count] reference-form table (first idx >= c) */
  uint32_t cell_table_capacity;
  uint32_t* candidate_count;  /* [N] sorted: sum over the 27 cells of (end - start)      */
  uint32_t* support_count;    /* [N] sorted: candidates with r/h < 1 (self included)     */
  float* density;             /* [N] sorted                                              */
  float* pressure;            /* [N] sorted                                              */
  float* acceleration;        /* [3N] sorted, after the force pass (the step zeroes it)  */
  uint32_t* collision_iters;  /* [N] sorted: iterations of the advect/collide do-while   */
} oracle_taps;

/* libclsph/common/util.h:41-62 and :21-38 */
uint32_t oracle_morton_encode(uint32_t x, uint32_t y, uint32_t z);
void oracle_morton_decode(uint32_t key, uint32_t xyz[3]);

/* libclsph/sph_simulation.cpp:490-505 -- h, time_delta, max_velocity, total_mass and the
 * five smoothing constants from the JSON-level inputs. Returns initial_volume. */
float oracle_derive_constants(simulation_parameters* p, precomputed_kernel_values* t,
                              int particles_inside_influence_radius);

/* libclsph/sph_simulation.cpp:48-94 (lattice branch); acceleration/grid_index zeroed (E13). */
void oracle_init_particles(particle* buffer, const simulation_parameters* p, float initial_volume);

/* libclsph/scene.cpp:36-64 */
void oracle_face_normals(const float* vertices, const uint32_t* indices, uint32_t face_count,
                         float* normals_out);

/* libclsph/sph_simulation.cpp:201-252. Returns 0, or 1 if a grid axis reaches 1024 cells. */
int oracle_bounds_and_grid(const particle* in, simulation_parameters* p);

/* libclsph/kernels/grid.cl:43-67 */
void oracle_locate_in_grid(const particle* in, particle* out, const simulation_parameters* p);

/* libclsph/kernels/sort.cl:1-59 + libclsph/sph_simulation.cpp:107-160 (E1: intended
 * semantics). `a` holds the input and receives the result (4 passes => back in `a`).
 * perm_out (nullable): a[r] after == a[perm_out[r]] before. Returns 0, or 1 if N < 128. */
int oracle_sort_particles(particle* a, particle* scratch, uint32_t n, uint32_t* perm_out);

/* libclsph/sph_simulation.cpp:163-170 */
void oracle_cell_table(const particle* sorted, uint32_t n, uint32_t grid_cell_count,
                       uint32_t* cell_table);

/* libclsph/kernels/sph.cl:9-40 + forces.cl:15-43 + smoothing.cl:1-4 */
void oracle_density_pressure(const particle* in, particle* out, const simulation_parameters* p,
                             const precomputed_kernel_values* t, const uint32_t* cell_table,
                             uint32_t* candidate_count, uint32_t* support_count);

/* libclsph/kernels/sph.cl:42-62 (E2: intended semantics) + forces.cl:45-112 + smoothing.cl:6-34 */
void oracle_forces(const particle* in, particle* out, const simulation_parameters* p,
                   const precomputed_kernel_values* t, const uint32_t* cell_table);

/* libclsph/kernels/sph.cl:64-112 + advection.cl:6-23 + collisions.cl:15-129.
 * max_iters caps the do-while (E9; 0 = uncapped like the reference). */
void oracle_advection_collision(const particle* in, particle* out, const simulation_parameters* p,
                                const float* face_normals, const float* vertices,
                                const uint32_t* indices, uint32_t face_count, uint32_t max_iters,
                                uint32_t* collision_iters);

/* libclsph/sph_simulation.cpp:173-344 -- one sub-step, host AoS in, host AoS out (in may
 * equal out). Rewrites the grid block of *p. Returns 0 on success, 1 on grid overflow
 * (axis >= 1024 cells), 2 if N < 128, 3 on allocation failure. */
int oracle_step(const particle* in, particle* out, simulation_parameters* p,
                const precomputed_kernel_values* t, const float* face_normals,
                const float* vertices, const uint32_t* indices, uint32_t face_count,
                oracle_taps* taps);

/* Threads the per-particle loops will use (1 when built without OpenMP). */
int oracle_num_threads(void);
void oracle_set_num_threads(int n);

#define ORACLE_MAX_COLLISION_ITERS 64u

#ifdef __cplusplus
}
#endif
#endif
