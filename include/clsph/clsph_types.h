/*
 * clsph_types.h -- plain-old-data records shared by the host API, the C ABI and the
 * CUDA kernels.
 *
 * These three records ARE the binary contract with libclsph user code: callbacks receive a
 * `particle*` and a `const simulation_parameters&`, the `last_frame.bin` checkpoint is a raw
 * dump of `particle[N]`, and the frame savers index `particles[i].position.s[0]`.
 * Field names, order, sizes and offsets therefore follow the reference records
 * (libclsph/common/structures.h:16-54 in the reference tree), where `cl_float3` is the
 * OpenCL host type: four floats, 16-byte aligned, addressable as `.s[0..3]`.
 *
 * Layout (checked by the static asserts at the bottom):
 *   particle                  80 B   position@0 velocity@16 intermediate_velocity@32
 *                                    acceleration@48 density@64 pressure@68 grid_index@72
 *   simulation_parameters    128 B   15 scalars @0..56, constant_acceleration@64,
 *                                    grid_size_x/y/z@80/84/88, grid_cell_count@92,
 *                                    min_point@96, max_point@112
 *   precomputed_kernel_values 20 B   five floats
 */
#ifndef CLSPH_TYPES_H_
#define CLSPH_TYPES_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
#define CLSPH_STATIC_ASSERT(c, m) static_assert(c, m)
#else
#define CLSPH_STATIC_ASSERT(c, m) _Static_assert(c, m)
#endif
/* gcc, g++, clang and nvcc all accept the attribute spelling in both C and C++. */
#define CLSPH_ALIGN16 __attribute__((aligned(16)))

/* Host-side OpenCL scalar names that libclsph user code spells out. */
typedef float cl_float;
typedef uint32_t cl_uint;
typedef int32_t cl_int;

/* Four lanes, the fourth is padding; `.s[k]` is how the reference host code reads it
 * (e.g. libclsph/file_save_delegates/houdini_file_saver.cpp:40-45). */
typedef union cl_float3 {
  float s[4];
#if defined(__GNUC__) || defined(__clang__)
  __extension__ struct {
    float x, y, z, w;
  };
#endif
} CLSPH_ALIGN16 cl_float3;

typedef union cl_uint3 {
  uint32_t s[4];
#if defined(__GNUC__) || defined(__clang__)
  __extension__ struct {
    uint32_t x, y, z, w;
  };
#endif
} CLSPH_ALIGN16 cl_uint3;

/* Per-run constants plus the per-step grid description. The grid block (grid_size_*,
 * grid_cell_count, min_point, max_point) is rewritten on every sub-step, exactly as
 * libclsph/sph_simulation.cpp:229-252 does, because callbacks observe it. */
typedef struct simulation_parameters {
  cl_uint particles_count;             /*   0 */
  cl_float max_velocity;               /*   4  0.8 * h / time_delta                  */
  cl_float fluid_density;              /*   8  rest density rho0                     */
  cl_float total_mass;                 /*  12 */
  cl_float particle_mass;              /*  16 */
  cl_float dynamic_viscosity;          /*  20  mu                                    */
  cl_float simulation_time;            /*  24  seconds                               */
  cl_float target_fps;                 /*  28 */
  cl_float h;                          /*  32  support radius                        */
  cl_float simulation_scale;           /*  36  sub-step = time_delta * scale         */
  cl_float time_delta;                 /*  40  1 / target_fps                        */
  cl_float surface_tension_threshold;  /*  44 */
  cl_float surface_tension;            /*  48  sigma                                 */
  cl_float restitution;                /*  52  in [0, 1]                             */
  cl_float K;                          /*  56  Tait stiffness                        */

  cl_float3 constant_acceleration;     /*  64  gravity                               */

  cl_int grid_size_x;                  /*  80  cells of side 2h along x              */
  cl_int grid_size_y;                  /*  84 */
  cl_int grid_size_z;                  /*  88 */
  cl_uint grid_cell_count;             /*  92  morton(grid_size_x, _y, _z)           */
  cl_float3 min_point, max_point;      /*  96, 112  padded particle AABB             */
} simulation_parameters;

/* One fluid particle. There is no id field: identity is the array position, which is why
 * the sort permutation is part of the contract. */
typedef struct particle {
  cl_float3 position, velocity, intermediate_velocity, acceleration; /* 0,16,32,48 */
  cl_float density, pressure;                                        /* 64, 68     */
  cl_uint grid_index;                                                /* 72         */
} particle;

/* Smoothing-kernel normalisation constants (libclsph/sph_simulation.cpp:499-505). */
typedef struct precomputed_kernel_values {
  float poly_6;            /*  315 / (64 pi h^9) */
  float poly_6_gradient;   /* -945 / (32 pi h^9) */
  float poly_6_laplacian;  /* -945 / (32 pi h^9) */
  float spiky;             /*  -45 / (pi h^6)    */
  float viscosity;         /*   45 / (pi h^6)    */
} precomputed_kernel_values;

CLSPH_STATIC_ASSERT(sizeof(cl_float3) == 16, "cl_float3 must be 16 bytes");
CLSPH_STATIC_ASSERT(sizeof(particle) == 80, "particle must be 80 bytes");
CLSPH_STATIC_ASSERT(offsetof(particle, velocity) == 16, "particle.velocity");
CLSPH_STATIC_ASSERT(offsetof(particle, intermediate_velocity) == 32, "particle.intermediate_velocity");
CLSPH_STATIC_ASSERT(offsetof(particle, acceleration) == 48, "particle.acceleration");
CLSPH_STATIC_ASSERT(offsetof(particle, density) == 64, "particle.density");
CLSPH_STATIC_ASSERT(offsetof(particle, pressure) == 68, "particle.pressure");
CLSPH_STATIC_ASSERT(offsetof(particle, grid_index) == 72, "particle.grid_index");
CLSPH_STATIC_ASSERT(sizeof(simulation_parameters) == 128, "simulation_parameters must be 128 bytes");
CLSPH_STATIC_ASSERT(offsetof(simulation_parameters, K) == 56, "simulation_parameters.K");
CLSPH_STATIC_ASSERT(offsetof(simulation_parameters, constant_acceleration) == 64, "params.constant_acceleration");
CLSPH_STATIC_ASSERT(offsetof(simulation_parameters, grid_size_x) == 80, "params.grid_size_x");
CLSPH_STATIC_ASSERT(offsetof(simulation_parameters, grid_cell_count) == 92, "params.grid_cell_count");
CLSPH_STATIC_ASSERT(offsetof(simulation_parameters, min_point) == 96, "params.min_point");
CLSPH_STATIC_ASSERT(offsetof(simulation_parameters, max_point) == 112, "params.max_point");
CLSPH_STATIC_ASSERT(sizeof(precomputed_kernel_values) == 20, "precomputed_kernel_values must be 20 bytes");

#endif /* CLSPH_TYPES_H_ */
