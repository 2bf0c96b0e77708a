// subview.cuh -- index ranges of the sub-cell order: from (cell, octant range) to a range of the sorted
// arrays, and the per-particle traversal of the sub-cells around a particle (global memory). Shared by
// subgrid.cu (sort-order bookkeeping, density kernels) and dist.cu (order keys of the multi-GPU exchange).
// See the header comment of subgrid.cu for the organisation itself.
#pragma once

#include "kernels.cuh"

namespace clsph {

// What a kernel needs to turn (cell, octant range) into an index range of the sorted arrays.
struct SubView {
  const uint32_t* lb;     // dense table: lb[cell * 9 + o] = first index of the cell's octant >= o; [8] = end
  const uint32_t* fkeys;  // sorted sub-cell keys (binary-search fallback)
  uint32_t n, cell_count;
  bool dense;
};

__device__ __forceinline__ SubView make_view(const GridState& g, const uint32_t* sub_lb, const uint32_t* keys_a,
                                             const uint32_t* keys_b) {
  SubView v;
  v.lb = sub_lb;
  v.fkeys = (g.sort_passes & 1u) ? keys_b : keys_a;  // an odd number of passes ends in the "b" buffers
  v.n = g.n;
  v.cell_count = g.cell_count;
  v.dense = g.sub_dense != 0u;
  return v;
}

// Particles of cell `key` whose octant is in [o_lo, o_hi]: one contiguous range.
__device__ __forceinline__ uint2 sub_range(const SubView& v, uint32_t key, uint32_t o_lo, uint32_t o_hi) {
  if (key >= v.cell_count) return make_uint2(0u, 0u);
  if (v.dense) {
    const uint32_t* row = v.lb + (size_t)key * 9u;
    return make_uint2(__ldg(row + o_lo), __ldg(row + o_hi + 1u));
  }
  const uint32_t base = key << 3;  // key < 2^29 in sub-cell mode
  const uint32_t a = lower_bound_key(v.fkeys, v.n, base + o_lo);
  const uint32_t b = (base + o_hi == 0xFFFFFFFFu) ? v.n : lower_bound_key(v.fkeys, v.n, base + o_hi + 1u);
  return make_uint2(a, b);
}

// Sub-cell coordinates [lo, hi] along one axis that can hold a particle within the support of a
// particle at offset u = p - min. A pair inside the support has |dx| < h (1 + 2^-21); u itself
// carries up to 2^-13 h of rounding (u < 2048 h). Both are covered by searching
// [u - hm, u + hm] with hm = h (1 + 2^-10), mapped to sub-cells by the SAME monotone rounding
// sequence as sub_coord: every neighbour's sub-cell lies in [lo, hi]. Usually hi - lo = 2; 3 when
// the particle is within 2^-10 h of a sub-cell boundary.
__device__ __forceinline__ void sub_bounds(float p, float mn, float cell, float hm, uint32_t& lo, uint32_t& hi) {
  const float u = __fsub_rn(p, mn);
  const float ql = __fdiv_rn(__fsub_rn(u, hm), cell), qh = __fdiv_rn(__fadd_rn(u, hm), cell);
  lo = __float2uint_rz(__fadd_rn(ql, ql));  // negative -> 0
  // 2047 = last sub-cell a 10-bit cell coordinate can have: keeps the loops bounded for a particle that
  // has blown up (infinite or huge position; the step then reports CLSPH_EGRID anyway)
  hi = min(__float2uint_rz(__fadd_rn(qh, qh)), 2047u);
}

// *addr = v when ok, as ONE predicated store: the compiler would otherwise branch around the store
// and its address arithmetic, and in a warp that branch is nearly always taken by some lane.
__device__ __forceinline__ void store_if(bool ok, uint32_t* addr, uint32_t v) {
#ifdef CLSPH_EMU  // tests/emu: CPU build of the kernels for logic tests, no PTX
  if (ok) *addr = v;
#else
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %0, 0;\n\t@p st.global.u32 [%1], %2;\n\t}" ::"r"((uint32_t)ok),
               "l"(__cvta_generic_to_global(addr)), "r"(v)
               : "memory");
#endif
}

// Two consecutive words at an 8-byte aligned address, as ONE predicated store.
__device__ __forceinline__ void store2_if(bool ok, uint32_t* addr, uint32_t a, uint32_t b) {
#ifdef CLSPH_EMU
  if (ok) { addr[0] = a; addr[1] = b; }
#else
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %0, 0;\n\t@p st.global.v2.u32 [%1], {%2, %3};\n\t}" ::"r"((uint32_t)ok),
               "l"(__cvta_generic_to_global(addr)), "r"(a), "r"(b)
               : "memory");
#endif
}

// The same on an address that is already in the global window (cvta once per row pointer, not once per store: the
// conversion showed up as 6 % of the pair kernel's instructions). `words` = offset in 32-bit words.
#ifdef CLSPH_EMU
typedef uint32_t* global_row_t;
__device__ __forceinline__ global_row_t global_row(uint32_t* row) { return row; }
__device__ __forceinline__ void row_store_if(bool ok, global_row_t row, uint32_t words, uint32_t v) { if (ok) row[words] = v; }
__device__ __forceinline__ void row_store2_if(bool ok, global_row_t row, uint32_t words, uint32_t a, uint32_t b) {
  if (ok) { row[words] = a; row[words + 1] = b; }
}
#else
typedef unsigned long long global_row_t;
__device__ __forceinline__ global_row_t global_row(uint32_t* row) { return (global_row_t)__cvta_generic_to_global(row); }
__device__ __forceinline__ void row_store_if(bool ok, global_row_t row, uint32_t words, uint32_t v) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %0, 0;\n\t@p st.global.u32 [%1], %2;\n\t}" ::"r"((uint32_t)ok),
               "l"(row + 4ull * words), "r"(v)
               : "memory");
}
__device__ __forceinline__ void row_store2_if(bool ok, global_row_t row, uint32_t words, uint32_t a, uint32_t b) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %0, 0;\n\t@p st.global.v2.u32 [%1], {%2, %3};\n\t}" ::"r"((uint32_t)ok),
               "l"(row + 4ull * words), "r"(a), "r"(b)
               : "memory");
}
#endif

// for_each_range: calls range(begin, end) for every index range of candidates of the sub-cells around pi.
// for_each_neighbour, on top of it:
// Calls visit(j, pos[j], s, inside) for every candidate j of the sub-cells around pi, z outermost /
// x innermost; inside = (s = |pi - pj|^2) < support_s is the reference's window test (self included).
// The visitor gets every candidate so that it can stay branch-free: about one candidate in seven is
// inside, so in a warp some lane nearly always is, and a divergent "inside" branch would run for all.
template <class Range>
__device__ __forceinline__ void for_each_range(const SubView& v, const GridState& g, const SphConst& c, const float4& pi,
                                               Range&& range) {
  uint32_t xlo, xhi, ylo, yhi, zlo, zhi;
  sub_bounds(pi.x, g.min_x, g.cell, c.h_margin, xlo, xhi);
  sub_bounds(pi.y, g.min_y, g.cell, c.h_margin, ylo, yhi);
  sub_bounds(pi.z, g.min_z, g.cell, c.h_margin, zlo, zhi);
  const uint32_t cx_lo = xlo >> 1, cx_hi = xhi >> 1;
  for (uint32_t fz = zlo; fz <= zhi; ++fz) {
    const uint32_t kz = spread10(fz >> 1) << 2, oz = (fz & 1u) << 2;
    for (uint32_t fy = ylo; fy <= yhi; ++fy) {
      const uint32_t kzy = kz | (spread10(fy >> 1) << 1), ozy = oz | ((fy & 1u) << 1);
      // the x extent of the row covers two (rarely three) cells; inside a cell the octants with
      // x bit 0 and 1 are adjacent, so each cell contributes one range
      for (uint32_t cx = cx_lo; cx <= cx_hi; ++cx) {
        const uint32_t o_lo = ozy | (cx == cx_lo ? (xlo & 1u) : 0u);
        const uint32_t o_hi = ozy | (cx == cx_hi ? (xhi & 1u) : 1u);
        const uint2 r = sub_range(v, kzy | spread10(cx), o_lo, o_hi);
        range(r.x, r.y);
      }
    }
  }
}

// for_each_row: calls row(a0, a1, b0, b1) once per (z, y) row of sub-cells with the row's two index ranges
// (the x extent of a row covers two cells; a third, which only happens within 2^-10 h of a sub-cell
// boundary, is delivered as a row of its own). Walking both ranges in ONE loop matters in a warp: lanes
// sit in different sub-cells, a loop runs as long as its longest lane, and the sum of two ranges varies
// less between lanes than each of them (measured on the bench states: 164 instead of 216 candidate
// slots per lane for 123 candidates).
template <class Row>
__device__ __forceinline__ void for_each_row(const SubView& v, const GridState& g, const SphConst& c, const float4& pi,
                                             Row&& row) {
  uint32_t xlo, xhi, ylo, yhi, zlo, zhi;
  sub_bounds(pi.x, g.min_x, g.cell, c.h_margin, xlo, xhi);
  sub_bounds(pi.y, g.min_y, g.cell, c.h_margin, ylo, yhi);
  sub_bounds(pi.z, g.min_z, g.cell, c.h_margin, zlo, zhi);
  const uint32_t cx_lo = xlo >> 1, cx_hi = xhi >> 1;
  for (uint32_t fz = zlo; fz <= zhi; ++fz) {
    const uint32_t kz = spread10(fz >> 1) << 2, oz = (fz & 1u) << 2;
    for (uint32_t fy = ylo; fy <= yhi; ++fy) {
      const uint32_t kzy = kz | (spread10(fy >> 1) << 1), ozy = oz | ((fy & 1u) << 1);
      const uint2 a = sub_range(v, kzy | spread10(cx_lo), ozy | (xlo & 1u), ozy | (cx_hi == cx_lo ? (xhi & 1u) : 1u));
      uint2 b = make_uint2(0u, 0u);
      if (cx_hi > cx_lo) b = sub_range(v, kzy | spread10(cx_lo + 1u), ozy, ozy | (cx_hi == cx_lo + 1u ? (xhi & 1u) : 1u));
      row(a.x, a.y, b.x, b.y);
      if (cx_hi > cx_lo + 1u) {  // rare third cell of the row
        const uint2 e = sub_range(v, kzy | spread10(cx_hi), ozy, ozy | (xhi & 1u));
        row(e.x, e.y, 0u, 0u);
      }
    }
  }
}

template <class Visit>
__device__ __forceinline__ void for_each_neighbour(const SubView& v, const GridState& g, const SphConst& c,
                                                   const float4* pos, const float4& pi, Visit&& visit) {
  for_each_range(v, g, c, pi, [&](uint32_t begin, uint32_t end) {
    for (uint32_t j = begin; j < end; ++j) {
      const float4 pj = pos[j];
      const float s = dist2_contract(pi.x, pi.y, pi.z, pj.x, pj.y, pj.z);
      visit(j, pj, s, s < c.support_s);
    }
  });
}


// Slot for one element per true lane of a converged warp: one atomicAdd per warp, lanes get consecutive slots.
__device__ __forceinline__ uint32_t warp_append(bool want, uint32_t* counter) {
  const unsigned m = __ballot_sync(kFullMask, want);
  if (m == 0u) return 0u;
  uint32_t base = 0;
  const int leader = __ffs(m) - 1;
  if ((int)lane_id() == leader) base = atomicAdd(counter, (uint32_t)__popc(m));
  base = __shfl_sync(kFullMask, base, leader);
  return base + __popc(m & lanemask_lt());
}

// The same for a whole CTA of 256 threads (all of them must call it): ONE atomicAdd per CTA. With a million
// appending threads the per-warp form sends tens of thousands of atomics to one address, and that queue, not
// the copy, is what the kernel then waits for. Slots follow thread order inside the CTA.
__device__ __forceinline__ uint32_t block256_append(bool want, uint32_t* counter) {
  __shared__ uint32_t s_warp[8], s_base;
  const unsigned m = __ballot_sync(kFullMask, want);
  const unsigned warp = threadIdx.x >> 5;
  if (lane_id() == 0u) s_warp[warp] = (uint32_t)__popc(m);
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t total = 0;
    for (int w = 0; w < 8; ++w) { const uint32_t c = s_warp[w]; s_warp[w] = total; total += c; }
    s_base = total ? atomicAdd(counter, total) : 0u;
  }
  __syncthreads();
  const uint32_t at = s_base + s_warp[warp] + (uint32_t)__popc(m & lanemask_lt());
  __syncthreads();  // the shared words are reused by the next call
  return at;
}

}  // namespace clsph
