/*
 * oracle.c -- CPU restatement of libclsph's per-step SPH hot path (see oracle.h for the
 * scope, the "test infrastructure only" rule and the arithmetic contract).
 *
 * Build: gcc -O2 -std=c11 -ffp-contract=off -mfma [-fopenmp] -fPIC -shared (oracle/Makefile).
 * -ffp-contract=off is part of the contract: every a*b+c below that is meant to be fused is
 * written as fmaf(); everything else must round after each operator.
 *
 * Citations are paths inside the reference tree (/root/reference).
 */
#include "oracle.h"

#include <limits.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

/* ------------------------------------------------------------------------------------ */
/* OpenCL built-ins the reference kernels call, under the contract stated in oracle.h.   */
/* ------------------------------------------------------------------------------------ */

typedef struct {
  float x, y, z;
} v3;

static inline v3 v3_make(float x, float y, float z) {
  v3 r = {x, y, z};
  return r;
}
static inline v3 v3_load(const cl_float3* p) { return v3_make(p->s[0], p->s[1], p->s[2]); }
static inline void v3_store(cl_float3* p, v3 a) {
  p->s[0] = a.x;
  p->s[1] = a.y;
  p->s[2] = a.z;
}
static inline v3 v3_add(v3 a, v3 b) { return v3_make(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline v3 v3_sub(v3 a, v3 b) { return v3_make(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline v3 v3_scale(v3 a, float s) { return v3_make(a.x * s, a.y * s, a.z * s); }
static inline v3 v3_div(v3 a, float s) { return v3_make(a.x / s, a.y / s, a.z / s); }
static inline v3 v3_neg(v3 a) { return v3_make(-a.x, -a.y, -a.z); }

static inline float cl_dot(v3 a, v3 b) { return fmaf(a.z, b.z, fmaf(a.y, b.y, a.x * b.x)); }
static inline float cl_length(v3 a) { return sqrtf(cl_dot(a, a)); }
static inline float cl_distance(v3 a, v3 b) { return cl_length(v3_sub(a, b)); }
static inline v3 cl_normalize(v3 a) { return v3_div(a, cl_length(a)); }
static inline float cl_pown(float x, int n) {
  float r = x;
  for (int k = 1; k < n; ++k) r = r * x;
  return r;
}
static inline float cl_clamp(float x, float lo, float hi) { return fminf(fmaxf(x, lo), hi); }

/* Support window 1 - clamp(floor(r/h), 0, 1), shared by every smoothing function
 * (libclsph/kernels/smoothing.cl:2, 8, 14, 26, 33). */
static inline float window(float r, float h) { return 1.f - cl_clamp(floorf(r / h), 0.f, 1.f); }

/* ------------------------------------------------------------------------------------ */
/* Morton code, libclsph/common/util.h                                                    */
/* ------------------------------------------------------------------------------------ */

/* util.h:41-62: spread the low 10 bits of each coordinate to every third bit. */
static inline uint32_t spread10(uint32_t v) {
  v = (v | (v << 16)) & 0x030000FFu;
  v = (v | (v << 8)) & 0x0300F00Fu;
  v = (v | (v << 4)) & 0x030C30C3u;
  v = (v | (v << 2)) & 0x09249249u;
  return v;
}

uint32_t oracle_morton_encode(uint32_t x, uint32_t y, uint32_t z) {
  return spread10(x) | (spread10(y) << 1) | (spread10(z) << 2);
}

/* util.h:4-19: pick bits 0,3,6,...,27 back into bits 0..9. */
static inline uint32_t gather10(uint32_t v) {
  uint32_t r = 0;
  for (int k = 0; k < 10; ++k) r |= (v & (1u << (3 * k))) >> (2 * k);
  return r;
}

/* util.h:21-38 */
void oracle_morton_decode(uint32_t key, uint32_t xyz[3]) {
  const uint32_t mask = 0x9249249u;
  xyz[0] = gather10(key & mask);
  xyz[1] = gather10((key >> 1) & mask);
  xyz[2] = gather10((key >> 2) & mask);
}

/* ------------------------------------------------------------------------------------ */
/* Host-side constants, lattice, normals                                                  */
/* ------------------------------------------------------------------------------------ */

/* libclsph/sph_simulation.cpp:490-505. The mixed float/double evaluation order is the
 * reference's: 4.f*M_PI and every pow() are double, the rest is float. */
float oracle_derive_constants(simulation_parameters* p, precomputed_kernel_values* t,
                              int particles_inside_influence_radius) {
  p->total_mass = (float)p->particles_count * p->particle_mass;
  float initial_volume = p->total_mass / p->fluid_density;
  float per_particle = initial_volume / (float)p->particles_count;
  float numer = 3.f * ((float)particles_inside_influence_radius * per_particle);
  p->h = cbrtf((float)((double)numer / ((double)4.f * M_PI)));
  p->time_delta = 1.f / p->target_fps;
  p->max_velocity = 0.8f * p->h / p->time_delta;

  double h9 = pow((double)p->h, 9.0), h6 = pow((double)p->h, 6.0);
  t->poly_6 = (float)(315.0 / (64.0 * M_PI * h9));
  t->poly_6_gradient = (float)(-945.0 / (32.0 * M_PI * h9));
  t->poly_6_laplacian = (float)(-945.0 / (32.0 * M_PI * h9));
  t->spiky = (float)(-45.0 / (M_PI * h6));
  t->viscosity = (float)(45.0 / (M_PI * h6));
  return initial_volume;
}

/* libclsph/sph_simulation.cpp:48-56, 71-92 */
void oracle_init_particles(particle* buffer, const simulation_parameters* p, float initial_volume) {
  int per_side = (int)ceil((double)cbrtf((float)p->particles_count));
  float side_length = cbrtf(initial_volume);
  float spacing = side_length / (float)per_side;
  uint32_t ups = (uint32_t)per_side;
  for (uint32_t i = 0; i < p->particles_count; ++i) {
    memset(&buffer[i], 0, sizeof(particle));
    buffer[i].position.s[0] = (float)(i % ups) * spacing - side_length / 2.f;
    buffer[i].position.s[1] = (float)((i / ups) % ups) * spacing;
    buffer[i].position.s[2] = (float)(i / (ups * ups)) * spacing - side_length / 2.f;
  }
}

/* libclsph/scene.cpp:36-64. `sqrt` there is the double overload applied to a float sum. */
void oracle_face_normals(const float* vertices, const uint32_t* indices, uint32_t face_count,
                         float* normals_out) {
  for (uint32_t j = 0; j < face_count; ++j) {
    const float* a = vertices + 3 * indices[3 * j + 0];
    const float* b = vertices + 3 * indices[3 * j + 1];
    const float* c = vertices + 3 * indices[3 * j + 2];
    float ux = b[0] - a[0], uy = b[1] - a[1], uz = b[2] - a[2];
    float vx = c[0] - a[0], vy = c[1] - a[1], vz = c[2] - a[2];
    float nx = uy * vz - uz * vy;
    float ny = uz * vx - ux * vz;
    float nz = ux * vy - uy * vx;
    float length = (float)sqrt((double)(nx * nx + ny * ny + nz * nz));
    normals_out[3 * j + 0] = nx / length;
    normals_out[3 * j + 1] = ny / length;
    normals_out[3 * j + 2] = nz / length;
  }
}

/* ------------------------------------------------------------------------------------ */
/* Grid: bounds, keys, sort, table                                                        */
/* ------------------------------------------------------------------------------------ */

/* libclsph/sph_simulation.cpp:201-252 */
int oracle_bounds_and_grid(const particle* in, simulation_parameters* p) {
  float cell = p->h * 2;
  float lo[3], hi[3];
  for (int a = 0; a < 3; ++a) {
    lo[a] = (float)INT_MAX; /* :204 */
    hi[a] = (float)INT_MIN; /* :205 */
  }
  for (uint32_t i = 0; i < p->particles_count; ++i)
    for (int a = 0; a < 3; ++a) {
      float v = in[i].position.s[a];
      if (v < lo[a]) lo[a] = v;
      if (v > hi[a]) hi[a] = v;
    }
  int32_t gs[3];
  for (int a = 0; a < 3; ++a) {
    lo[a] -= cell * 2; /* :221-223 */
    hi[a] += cell * 2; /* :225-227 */
    p->min_point.s[a] = lo[a];
    p->max_point.s[a] = hi[a];
    gs[a] = (int32_t)(uint32_t)((hi[a] - lo[a]) / cell); /* :237-242 */
  }
  p->grid_size_x = gs[0];
  p->grid_size_y = gs[1];
  p->grid_size_z = gs[2];
  p->grid_cell_count = oracle_morton_encode((uint32_t)gs[0], (uint32_t)gs[1], (uint32_t)gs[2]);
  return (gs[0] < 1024 && gs[1] < 1024 && gs[2] < 1024) ? 0 : 1; /* :247-249 */
}

/* libclsph/kernels/grid.cl:43-67 */
void oracle_locate_in_grid(const particle* in, particle* out, const simulation_parameters* p) {
  const float cell = p->h * 2;
#pragma omp parallel for schedule(static)
  for (uint32_t i = 0; i < p->particles_count; ++i) {
    particle q = in[i];
    uint32_t c[3];
    for (int a = 0; a < 3; ++a) c[a] = (uint32_t)((q.position.s[a] - p->min_point.s[a]) / cell);
    q.grid_index = oracle_morton_encode(c[0], c[1], c[2]);
    out[i] = q;
  }
}

enum { SORT_CHUNKS = 128, SORT_BUCKETS = 256, SORT_PASSES = 4 };

/* Chunk bounds of work-item t, libclsph/kernels/sort.cl:9-20 (inclusive end). */
static inline void chunk_bounds(uint32_t n, int t, uint32_t* first, uint32_t* last) {
  uint32_t per = n / SORT_CHUNKS;
  *first = per * (uint32_t)t;
  *last = (t == SORT_CHUNKS - 1) ? n - 1 : *first + per - 1;
}

/* sort.cl:23-59 driven by sph_simulation.cpp:114-156. Counters are laid out
 * [bucket][chunk] and scanned in that linear order, which is what makes each pass stable. */
int oracle_sort_particles(particle* a, particle* scratch, uint32_t n, uint32_t* perm_out) {
  if (n < SORT_CHUNKS) return 1; /* E8 */
  uint32_t* counts = (uint32_t*)malloc(sizeof(uint32_t) * SORT_CHUNKS * SORT_BUCKETS);
  uint32_t* ida = perm_out ? (uint32_t*)malloc(sizeof(uint32_t) * n) : NULL;
  uint32_t* idb = perm_out ? (uint32_t*)malloc(sizeof(uint32_t) * n) : NULL;
  if (!counts || (perm_out && (!ida || !idb))) {
    free(counts);
    free(ida);
    free(idb);
    return 3;
  }
  if (perm_out)
    for (uint32_t i = 0; i < n; ++i) ida[i] = i;

  particle* src = a;
  particle* dst = scratch;
  for (int pass = 0; pass < SORT_PASSES; ++pass) {
    const int shift = 8 * pass;
    memset(counts, 0, sizeof(uint32_t) * SORT_CHUNKS * SORT_BUCKETS);
#pragma omp parallel for schedule(static)
    for (int t = 0; t < SORT_CHUNKS; ++t) { /* sort_count */
      uint32_t first, last;
      chunk_bounds(n, t, &first, &last);
      for (uint32_t i = first; i <= last; ++i)
        ++counts[((src[i].grid_index >> shift) & 0xFFu) * SORT_CHUNKS + (uint32_t)t];
    }
    uint32_t running = 0; /* sph_simulation.cpp:133-138 */
    for (int k = 0; k < SORT_CHUNKS * SORT_BUCKETS; ++k) {
      uint32_t c = counts[k];
      counts[k] = running;
      running += c;
    }
#pragma omp parallel for schedule(static)
    for (int t = 0; t < SORT_CHUNKS; ++t) { /* sort */
      uint32_t first, last;
      chunk_bounds(n, t, &first, &last);
      for (uint32_t i = first; i <= last; ++i) {
        uint32_t slot = ((src[i].grid_index >> shift) & 0xFFu) * SORT_CHUNKS + (uint32_t)t;
        uint32_t at = counts[slot]++;
        dst[at] = src[i];
        if (perm_out) idb[at] = ida[i];
      }
    }
    particle* tp = src;
    src = dst;
    dst = tp;
    uint32_t* ti = ida;
    ida = idb;
    idb = ti;
  }
  /* SORT_PASSES is even, so the result is back in `a`. */
  if (perm_out) memcpy(perm_out, ida, sizeof(uint32_t) * n);
  free(counts);
  free(ida);
  free(idb);
  return 0;
}

/* libclsph/sph_simulation.cpp:163-170 */
void oracle_cell_table(const particle* sorted, uint32_t n, uint32_t grid_cell_count,
                       uint32_t* cell_table) {
  uint32_t at = 0;
  for (uint32_t c = 0; c < grid_cell_count; ++c) {
    cell_table[c] = at;
    while (at != n && sorted[at].grid_index == c) ++at;
  }
}

/* libclsph/kernels/grid.cl:22-32. A key at or beyond grid_cell_count cannot occur (cell
 * coordinates of real particles sit two cells inside the padded AABB); guarded here so a
 * malformed input cannot read outside the table. */
static inline void cell_range(uint32_t c, const uint32_t* table, const simulation_parameters* p,
                              uint32_t* first, uint32_t* end) {
  if (c >= p->grid_cell_count) {
    *first = *end = 0;
    return;
  }
  *first = table[c];
  *end = (p->grid_cell_count > c + 1) ? table[c + 1] : p->particles_count;
}

/* ------------------------------------------------------------------------------------ */
/* Density / pressure                                                                     */
/* ------------------------------------------------------------------------------------ */

/* libclsph/kernels/smoothing.cl:1-4 */
static inline float poly_6(float r, float h, const precomputed_kernel_values* t) {
  return window(r, h) * t->poly_6 * cl_pown(cl_pown(h, 2) - cl_pown(r, 2), 3);
}

/* sph.cl:9-40 with compute_density_with_grid (forces.cl:15-43) inlined. */
void oracle_density_pressure(const particle* in, particle* out, const simulation_parameters* p,
                             const precomputed_kernel_values* t, const uint32_t* cell_table,
                             uint32_t* candidate_count, uint32_t* support_count) {
#pragma omp parallel for schedule(dynamic, 256)
  for (uint32_t i = 0; i < p->particles_count; ++i) {
    uint32_t cc[3];
    oracle_morton_decode(in[i].grid_index, cc);
    const v3 xi = v3_load(&in[i].position);
    float density = 0.f;
    uint32_t n_cand = 0, n_supp = 0;
    /* unsigned wrap-around of cc-1 is the reference's behaviour too (forces.cl:25-27) */
    for (uint32_t z = cc[2] - 1; z <= cc[2] + 1; ++z)
      for (uint32_t y = cc[1] - 1; y <= cc[1] + 1; ++y)
        for (uint32_t x = cc[0] - 1; x <= cc[0] + 1; ++x) {
          uint32_t first, end;
          cell_range(oracle_morton_encode(x, y, z), cell_table, p, &first, &end);
          n_cand += end - first;
          for (uint32_t j = first; j < end; ++j) {
            float r = cl_distance(xi, v3_load(&in[j].position));
            density += p->particle_mass * poly_6(r, p->h, t);
            n_supp += (r / p->h < 1.f) ? 1u : 0u;
          }
        }
    particle q = in[i];
    q.density = density;
    /* Tait equation, sph.cl:37-39 */
    q.pressure = p->K * (cl_pown(density / p->fluid_density, 7) - 1.f);
    out[i] = q;
    if (candidate_count) candidate_count[i] = n_cand;
    if (support_count) support_count[i] = n_supp;
  }
}

/* ------------------------------------------------------------------------------------ */
/* Forces                                                                                 */
/* ------------------------------------------------------------------------------------ */

#define SPIKY_EPSILON 0.0000001f /* smoothing.cl:19 */

/* smoothing.cl:21-29. The degenerate branch returns a scalar broadcast to x,y,z (E3). */
static inline v3 spiky_gradient(v3 d, float len, float h, const precomputed_kernel_values* t) {
  if (len - SPIKY_EPSILON < 0.f && 0.f < len + SPIKY_EPSILON) {
    float c = -45.f / (float)(M_PI * (double)cl_pown(h, 6));
    return v3_make(c, c, c);
  }
  return v3_scale(v3_scale(v3_div(d, len), window(len, h) * t->spiky), cl_pown(h - len, 2));
}

/* smoothing.cl:31-34 */
static inline float viscosity_laplacian(float r, float h, const precomputed_kernel_values* t) {
  return window(r, h) * t->viscosity * (h - r);
}

/* smoothing.cl:6-10 */
static inline v3 poly_6_gradient(v3 d, float len, float h, const precomputed_kernel_values* t) {
  return v3_scale(v3_scale(d, window(len, h) * t->poly_6_gradient),
                  cl_pown(cl_pown(h, 2) - cl_pown(len, 2), 2));
}

/* smoothing.cl:12-17 */
static inline float poly_6_laplacian(float r, float h, const precomputed_kernel_values* t) {
  float w = window(fabsf(r), h);
  return w * t->poly_6_laplacian * (cl_pown(h, 2) - cl_pown(r, 2)) *
         (3.f * cl_pown(h, 2) - 7.f * cl_pown(r, 2));
}

/* sph.cl:42-62 with compute_internal_forces_with_grid (forces.cl:45-112) inlined.
 * E2: the output record is the input record with `acceleration` replaced. */
void oracle_forces(const particle* in, particle* out, const simulation_parameters* p,
                   const precomputed_kernel_values* t, const uint32_t* cell_table) {
  const float h = p->h, m = p->particle_mass;
#pragma omp parallel for schedule(dynamic, 256)
  for (uint32_t i = 0; i < p->particles_count; ++i) {
    uint32_t cc[3];
    oracle_morton_decode(in[i].grid_index, cc);
    const v3 xi = v3_load(&in[i].position);
    const v3 vi = v3_load(&in[i].velocity);
    const float rho_i = in[i].density, p_i = in[i].pressure;
    v3 pressure_term = {0.f, 0.f, 0.f}, viscosity_term = {0.f, 0.f, 0.f}, normal = {0.f, 0.f, 0.f};
    float color_field_laplacian = 0.f;

    for (uint32_t z = cc[2] - 1; z <= cc[2] + 1; ++z)
      for (uint32_t y = cc[1] - 1; y <= cc[1] + 1; ++y)
        for (uint32_t x = cc[0] - 1; x <= cc[0] + 1; ++x) {
          uint32_t first, end;
          cell_range(oracle_morton_encode(x, y, z), cell_table, p, &first, &end);
          for (uint32_t j = first; j < end; ++j) {
            const v3 d = v3_sub(xi, v3_load(&in[j].position));
            const float len = cl_length(d);
            const float rho_j = in[j].density;
            if (j != i) {
              /* forces.cl:70-77 */
              float coeff = (in[j].pressure / cl_pown(rho_j, 2) + p_i / cl_pown(rho_i, 2)) * m;
              pressure_term = v3_add(pressure_term, v3_scale(spiky_gradient(d, len, h, t), coeff));
              /* forces.cl:79-85 */
              v3 dv = v3_sub(v3_load(&in[j].velocity), vi);
              viscosity_term = v3_add(
                  viscosity_term, v3_scale(v3_scale(dv, m / rho_j), viscosity_laplacian(len, h, t)));
            }
            /* forces.cl:88-97 */
            normal = v3_add(normal, v3_scale(poly_6_gradient(d, len, h, t), m / rho_j));
            color_field_laplacian += m / rho_j * poly_6_laplacian(len, h, t);
          }
        }

    /* forces.cl:103-109 */
    v3 sum = v3_add(v3_scale(pressure_term, -rho_i), v3_scale(viscosity_term, p->dynamic_viscosity));
    float nlen = cl_length(normal);
    if (nlen > p->surface_tension_threshold)
      sum = v3_add(sum, v3_div(v3_scale(normal, -p->surface_tension * color_field_laplacian), nlen));

    /* sph.cl:53-58 */
    v3 acc = v3_add(v3_div(sum, rho_i), v3_load(&p->constant_acceleration));
    particle q = in[i];
    v3_store(&q.acceleration, acc);
    out[i] = q;
  }
}

/* ------------------------------------------------------------------------------------ */
/* Advection + collision                                                                  */
/* ------------------------------------------------------------------------------------ */

typedef struct {
  v3 point, normal;
  float depth;
  int happened;
} hit_t;

/* collisions.cl:15-89: nearest hit of segment p0->p1 over all faces, later face wins ties. */
static int detect_collision(hit_t* c, v3 p0, v3 p1, const float* face_normals,
                            const float* vertices, const uint32_t* indices, uint32_t face_count) {
  c->happened = 0;
  const v3 travel = v3_sub(p1, p0);
  for (uint32_t f = 0; f < face_count; ++f) {
    v3 n = v3_make(face_normals[3 * f + 0], face_normals[3 * f + 1], face_normals[3 * f + 2]);
    /* :27-29 orient the normal along the direction of travel */
    if (cl_dot(n, travel) / (cl_length(n) * cl_length(travel)) <= 0) n = v3_neg(n);

    const float* a = vertices + 3 * indices[3 * f + 0];
    const float* b = vertices + 3 * indices[3 * f + 1];
    const float* cc = vertices + 3 * indices[3 * f + 2];
    const v3 v0 = v3_make(a[0], a[1], a[2]);
    const v3 u = v3_sub(v3_make(b[0], b[1], b[2]), v0);
    const v3 v = v3_sub(v3_make(cc[0], cc[1], cc[2]), v0);

    float denom = cl_dot(n, travel); /* :52-56 */
    if (denom == 0.f) continue;
    float r = cl_dot(n, v3_sub(v0, p0)) / denom; /* :58 */
    if (0 <= r && r <= 1) {
      v3 hit = v3_add(p0, v3_scale(travel, r));
      v3 w = v3_sub(hit, v0);
      float uv = cl_dot(u, v), wv = cl_dot(w, v), vv = cl_dot(v, v), wu = cl_dot(w, u),
            uu = cl_dot(u, u);
      float d2 = uv * uv - uu * vv; /* :71-73 */
      float s = (uv * wv - vv * wu) / d2;
      float t = (uv * wu - uu * wv) / d2;
      if (s >= 0 && t >= 0 && s + t <= 1) {
        if (c->happened && cl_length(v3_sub(p0, hit)) > cl_length(v3_sub(p0, c->point))) continue;
        c->normal = n;
        c->point = hit;
        c->depth = cl_length(v3_sub(p1, hit));
        c->happened = 1;
      }
    }
  }
  return c->happened;
}

/* sph.cl:64-112, advection.cl:6-23, collisions.cl:91-129 */
void oracle_advection_collision(const particle* in, particle* out, const simulation_parameters* p,
                                const float* face_normals, const float* vertices,
                                const uint32_t* indices, uint32_t face_count, uint32_t max_iters,
                                uint32_t* collision_iters) {
#pragma omp parallel for schedule(dynamic, 256)
  for (uint32_t i = 0; i < p->particles_count; ++i) {
    float time_to_go = p->time_delta * p->simulation_scale;
    v3 pos = v3_load(&in[i].position);
    v3 vel = v3_load(&in[i].intermediate_velocity);
    v3 acc = v3_load(&in[i].acceleration);
    uint32_t iters = 0;
    int collided;
    do {
      /* advect: leapfrog with a speed clamp */
      v3 next_v = v3_add(vel, v3_scale(acc, time_to_go));
      if (cl_length(next_v) > p->max_velocity) next_v = v3_scale(cl_normalize(next_v), p->max_velocity);
      v3 new_pos = v3_add(pos, v3_scale(next_v, time_to_go));

      /* handle_collisions */
      v3 res_pos = new_pos;
      float used = time_to_go;
      hit_t c = {{0.f, 0.f, 0.f}, {0.f, 0.f, 0.f}, 0.f, 0};
      collided = detect_collision(&c, pos, new_pos, face_normals, vertices, indices, face_count);
      if (collided) {
        res_pos = v3_sub(c.point, v3_scale(c.normal, 0.001f)); /* :95 */
        float k = 1.f + p->restitution * c.depth / (time_to_go * cl_length(next_v));
        next_v = v3_sub(next_v, v3_scale(c.normal, k * cl_dot(next_v, c.normal))); /* :97-101 */
        used = time_to_go * (cl_length(v3_sub(res_pos, pos)) / cl_length(v3_sub(new_pos, pos)));
      }
      pos = res_pos;
      vel = next_v;
      time_to_go -= used;
      acc = v3_make(0.f, 0.f, 0.f); /* sph.cl:97-99 */
      ++iters;
    } while (collided && (max_iters == 0 || iters < max_iters));

    particle q = in[i];
    v3_store(&q.velocity, v3_div(v3_add(v3_load(&in[i].intermediate_velocity), vel), 2.f));
    v3_store(&q.intermediate_velocity, vel);
    v3_store(&q.position, pos);
    v3_store(&q.acceleration, v3_make(0.f, 0.f, 0.f)); /* E6 */
    out[i] = q;
    if (collision_iters) collision_iters[i] = iters;
  }
}

/* ------------------------------------------------------------------------------------ */
/* One sub-step                                                                           */
/* ------------------------------------------------------------------------------------ */

/* libclsph/sph_simulation.cpp:173-344 without the host<->device staging. */
int oracle_step(const particle* in, particle* out, simulation_parameters* p,
                const precomputed_kernel_values* t, const float* face_normals,
                const float* vertices, const uint32_t* indices, uint32_t face_count,
                oracle_taps* taps) {
  const uint32_t n = p->particles_count;
  if (n < SORT_CHUNKS) return 2;
  if (oracle_bounds_and_grid(in, p)) return 1;

  particle* a = (particle*)malloc(sizeof(particle) * n);
  particle* b = (particle*)malloc(sizeof(particle) * n);
  uint32_t* table = (uint32_t*)malloc(sizeof(uint32_t) * (p->grid_cell_count ? p->grid_cell_count : 1));
  uint32_t* perm = (taps && taps->permutation) ? taps->permutation : NULL;
  if (!a || !b || !table) {
    free(a);
    free(b);
    free(table);
    return 3;
  }

  oracle_locate_in_grid(in, a, p);
  if (taps && taps->keys_input_order)
    for (uint32_t i = 0; i < n; ++i) taps->keys_input_order[i] = a[i].grid_index;
  int rc = oracle_sort_particles(a, b, n, perm);
  if (rc) {
    free(a);
    free(b);
    free(table);
    return rc == 1 ? 2 : rc;
  }
  oracle_cell_table(a, n, p->grid_cell_count, table);
  if (taps && taps->cell_table) {
    uint32_t m = p->grid_cell_count < taps->cell_table_capacity ? p->grid_cell_count
                                                                 : taps->cell_table_capacity;
    memcpy(taps->cell_table, table, sizeof(uint32_t) * m);
  }

  oracle_density_pressure(a, b, p, t, table, taps ? taps->candidate_count : NULL,
                          taps ? taps->support_count : NULL);
  if (taps && taps->density)
    for (uint32_t i = 0; i < n; ++i) taps->density[i] = b[i].density;
  if (taps && taps->pressure)
    for (uint32_t i = 0; i < n; ++i) taps->pressure[i] = b[i].pressure;

  oracle_forces(b, a, p, t, table);
  if (taps && taps->acceleration)
    for (uint32_t i = 0; i < n; ++i)
      for (int k = 0; k < 3; ++k) taps->acceleration[3 * i + k] = a[i].acceleration.s[k];

  oracle_advection_collision(a, out, p, face_normals, vertices, indices, face_count,
                             ORACLE_MAX_COLLISION_ITERS, taps ? taps->collision_iters : NULL);
  free(a);
  free(b);
  free(table);
  return 0;
}

int oracle_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

void oracle_set_num_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}
