// common.cuh -- shared device-side definitions of the sm_100a SPH step.
//
// Data layout in HBM (one set per ping-pong side, all arrays of length N in cell-sorted order):
//   pos   float4  x y z .          vel   float4  vx vy vz .        ivel  float4  (leapfrog half-step velocity)
//   aux   float4  rho, p, p/rho^2, m/rho   (written by the density pass, read by the force pass)
//   skey  uint32  Morton cell key
// plus the dense cell table cell_start[c], cell_end[c] indexed by Morton key.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace clsph {

constexpr int kWarp = 32;
constexpr unsigned kFullMask = 0xffffffffu;

// Written by k_grid_setup on the device each sub-step (sph_simulation.cpp:221-252 of the
// reference), read by every later kernel of the step. Lives in device memory so that a whole
// run of sub-steps needs no host round trip.
struct GridState {
  float min_x, min_y, min_z, cell;   // padded AABB minimum, cell side 2h
  float max_x, max_y, max_z;
  float plane_lo;                    // this rank's slab in world space: plane_lo <= x < plane_hi (-inf / +inf at the ends)
  int gx, gy, gz;                    // grid_size_*
  uint32_t cell_count;               // morton(gx, gy, gz)
  uint32_t n;                        // particles on this device
  uint32_t sort_passes;              // ceil(bits(cell_count - 1) / 8), 1..4; 0: counting sort on the dense sub-cell table (sort.cu)
  uint32_t dense;                    // 1: cell_count fits the dense table; 0: binary-search fallback
  uint32_t error;                    // bit 0: a grid axis reached 1024 cells; bit 1: a multi-GPU buffer overflowed;
                                     // bit 2: sub-cell keys need more than 32 bits (Morton cell count beyond 2^29: z axis of 512+ cells)
  // Slab decomposition (multi-GPU): this rank owns the cells with own_lo <= cx < own_hi. One GPU:
  // [0, INT_MAX). Particles of other cells held locally are ghost copies of a neighbour's.
  int own_lo, own_hi;
  int prev_lo, prev_hi;              // the range of the previous sub-step (whose keys the arrays still carry)
  uint32_t fresh;                    // 1 after an upload: every local particle counts as owned
  // Sub-cell order (subgrid.cu): the arrays are sorted by (cell key << 3 | octant of the cell), so the
  // particles of a cell are grouped by the h-sized sub-cell they are in.
  uint32_t sub;                      // 1: this sub-step sorts by sub-cell keys
  uint32_t sub_dense;                // 1: cell_count fits the dense sub-cell table; 0: binary search
  float plane_hi;
  // What the last sub-step left in the dense sub-cell table, so that the launch of the next k_grid_setup can zero it
  // again (the table is all zero whenever a sub-step starts). Set by k_keys_hist, never by k_grid_setup itself.
  // table_words: 9 per cell of that step's grid (0: no table was written). table_full: a counting sort scanned the whole
  // range; otherwise k_reorder_sub wrote only rows of cells that hold particles -- those of the table_n sorted keys the
  // sort left in its "a" (table_in_b: "b") buffer -- and only these are zeroed: on a grid much larger than the fluid
  // (a slab of a long multi-GPU domain: 600 MB of table for 1 Mi particles) that is the difference between 100 and 3 us.
  uint32_t table_words, table_full, table_n, table_in_b;
};

// AABB accumulators: floats mapped to order-preserving unsigned so atomicMin/Max apply.
struct BoundsAcc {
  uint32_t lo[3];
  uint32_t hi[3];
};

// Per-run constants, passed to kernels by value.
struct SphConst {
  float h, h2, support_s;      // support_s: smallest s = |d|^2 with sqrt(s)/h >= 1 (exact window test)
  float h_margin;              // h (1 + 2^-10): search half-width of the sub-cell traversal (subgrid.cu)
  float mass, rho0, K;
  float c_poly6, c_spiky, c_visc, c_poly6_grad, c_poly6_lap;
  float mu, sigma, tension_threshold;
  float gx, gy, gz;
  float dt;                    // time_delta * simulation_scale
  float vmax, restitution;
  float spiky_degenerate;      // -45 / (pi h^6) as smoothing.cl:24 evaluates it (erratum E3)
  float degenerate_s;          // smallest s = |d|^2 with sqrt(s) >= 1e-7: smoothing.cl:23's test on s (add_pair_fast)
};

__host__ __device__ inline uint32_t float_to_ordered(float f) {
#ifdef __CUDA_ARCH__
  uint32_t b = __float_as_uint(f);
#else
  union { float f; uint32_t u; } c; c.f = f; uint32_t b = c.u;
#endif
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__host__ __device__ inline float ordered_to_float(uint32_t o) {
  uint32_t b = (o & 0x80000000u) ? (o & 0x7fffffffu) : ~o;
#ifdef __CUDA_ARCH__
  return __uint_as_float(b);
#else
  union { float f; uint32_t u; } c; c.u = b; return c.f;
#endif
}

// 10-bit-per-axis Morton code (libclsph/common/util.h:41-62 of the reference).
__host__ __device__ inline uint32_t spread10(uint32_t v) {
  v = (v | (v << 16)) & 0x030000FFu;
  v = (v | (v << 8)) & 0x0300F00Fu;
  v = (v | (v << 4)) & 0x030C30C3u;
  v = (v | (v << 2)) & 0x09249249u;
  return v;
}
__host__ __device__ inline uint32_t morton3(uint32_t x, uint32_t y, uint32_t z) {
  return spread10(x) | (spread10(y) << 1) | (spread10(z) << 2);
}
// Inverse of spread10 (util.h:4-19 does it bit by bit; this is the log-step form).
__host__ __device__ inline uint32_t compact10(uint32_t v) {
  v &= 0x09249249u;
  v = (v | (v >> 2)) & 0x030C30C3u;
  v = (v | (v >> 4)) & 0x0300F00Fu;
  v = (v | (v >> 8)) & 0x030000FFu;
  v = (v | (v >> 16)) & 0x000003FFu;
  return v;
}

// Role of a cell in the slab decomposition, from the x coordinate of its Morton key.
__device__ __forceinline__ bool cell_is_owned(uint32_t key, const GridState& g) {
  const int cx = (int)compact10(key);
  return cx >= g.own_lo && cx < g.own_hi;
}
// Ownership of a particle. Established organisation: by the cell it is in (slab boundaries snapped to cell
// boundaries). Sub-cell order: by the world-space planes themselves -- the kernels there work per particle,
// so nothing forces whole cells to change owner at once, slabs stay balanced and migration is only what
// really crosses a plane. On one GPU both say "owned".
__device__ __forceinline__ bool owned_here(float x, uint32_t key, const GridState& g) {
  if (g.sub) return x >= g.plane_lo && x < g.plane_hi;
  return cell_is_owned(key, g);
}
__device__ __forceinline__ bool slab_is_cut(const GridState& g) {
  if (g.sub) return g.plane_lo > -__int_as_float(0x7f800000) || g.plane_hi < __int_as_float(0x7f800000);
  return g.own_lo > 0 || g.own_hi != 0x7fffffff;
}
// Owned, or in the first ghost layer on either side (whose density the force pass needs).
__device__ __forceinline__ bool cell_needs_density(uint32_t key, const GridState& g) {
  const int cx = (int)compact10(key);
  return cx >= g.own_lo - 1 && cx <= g.own_hi;
}

// Cell coordinate of a position along one axis, grid.cl:56-61: (uint)((p - min) / (2h)) with a
// correctly rounded subtract and divide (no reciprocal, no contraction).
__device__ __forceinline__ uint32_t cell_coord(float p, float mn, float cell) {
  return __float2uint_rz(__fdiv_rn(__fsub_rn(p, mn), cell));
}

// Sub-cell coordinate (cells of side h): floor(2q) for the same q = (p - min) / (2h); the doubling is
// exact, so sub_coord >> 1 == cell_coord always.
__device__ __forceinline__ uint32_t sub_coord(float p, float mn, float cell) {
  const float q = __fdiv_rn(__fsub_rn(p, mn), cell);
  return __float2uint_rz(__fadd_rn(q, q));
}

// Squared distance exactly as the oracle's distance(): d = a - b, s = fma(dz,dz,fma(dy,dy,dx*dx)).
__device__ __forceinline__ float dist2_contract(float ax, float ay, float az, float bx, float by, float bz) {
  float dx = __fsub_rn(ax, bx), dy = __fsub_rn(ay, by), dz = __fsub_rn(az, bz);
  return __fmaf_rn(dz, dz, __fmaf_rn(dy, dy, __fmul_rn(dx, dx)));
}

// ---- packed fp32 (sm_100: FADD2 / FMUL2 / FFMA2, two IEEE fp32 lanes per instruction) -------------------
// Each lane rounds exactly like the scalar _rn operation, so results are bit-identical to scalar code; what
// changes is the issue slots: one instruction per two lanes. A scalar operand broadcast to both lanes costs
// nothing (the SASS operand form R.F32 next to R.F32x2.HI_LO).
#ifdef CLSPH_EMU
struct f32x2 { float lo, hi; };
__device__ __forceinline__ f32x2 f2_make(float lo, float hi) { return f32x2{lo, hi}; }
__device__ __forceinline__ float f2_lo(f32x2 a) { return a.lo; }
__device__ __forceinline__ float f2_hi(f32x2 a) { return a.hi; }
__device__ __forceinline__ f32x2 f2_sub(f32x2 a, f32x2 b) { return f32x2{__fsub_rn(a.lo, b.lo), __fsub_rn(a.hi, b.hi)}; }
__device__ __forceinline__ f32x2 f2_mul(f32x2 a, f32x2 b) { return f32x2{__fmul_rn(a.lo, b.lo), __fmul_rn(a.hi, b.hi)}; }
__device__ __forceinline__ f32x2 f2_fma(f32x2 a, f32x2 b, f32x2 c) { return f32x2{__fmaf_rn(a.lo, b.lo, c.lo), __fmaf_rn(a.hi, b.hi, c.hi)}; }
#else
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 f2_make(float lo, float hi) {
  f32x2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ float f2_lo(f32x2 a) {
  float lo, hi;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(a));
  return lo;
}
__device__ __forceinline__ float f2_hi(f32x2 a) {
  float lo, hi;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(a));
  return hi;
}
__device__ __forceinline__ f32x2 f2_sub(f32x2 a, f32x2 b) {
  f32x2 r;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ f32x2 f2_mul(f32x2 a, f32x2 b) {
  f32x2 r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ f32x2 f2_fma(f32x2 a, f32x2 b, f32x2 c) {
  f32x2 r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
#endif
__device__ __forceinline__ f32x2 f2_bcast(float v) { return f2_make(v, v); }

__device__ __forceinline__ unsigned lane_id() { return threadIdx.x & 31u; }
__device__ __forceinline__ unsigned lanemask_lt() {
#ifdef CLSPH_EMU  // tests/emu: CPU build of the kernels for logic tests, no PTX
  return (1u << lane_id()) - 1u;
#else
  unsigned m;
  asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
  return m;
#endif
}

// Cell range lookup. Dense: two loads. Fallback (grid larger than the table): lower_bound over
// the sorted key array, same answer.
__device__ __forceinline__ uint32_t lower_bound_key(const uint32_t* __restrict__ skey, uint32_t n, uint32_t key) {
  uint32_t lo = 0, hi = n;
  while (lo < hi) {
    uint32_t mid = (lo + hi) >> 1;
    if (__ldg(skey + mid) < key) lo = mid + 1; else hi = mid;
  }
  return lo;
}
__device__ __forceinline__ uint2 cell_range(uint32_t key, const GridState& g, const uint32_t* __restrict__ cell_start,
                                            const uint32_t* __restrict__ cell_end, const uint32_t* __restrict__ skey) {
  if (key >= g.cell_count) return make_uint2(0u, 0u);
  if (g.dense) return make_uint2(__ldg(cell_start + key), __ldg(cell_end + key));
  uint32_t a = lower_bound_key(skey, g.n, key);
  uint32_t b = lower_bound_key(skey, g.n, key + 1u);
  return make_uint2(a, b);
}

}  // namespace clsph
