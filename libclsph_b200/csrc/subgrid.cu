// subgrid.cu -- "sub-cell order": the neighbour search on cells of side h inside the reference's
// grid of side 2h.
//
// The reference scans the 27 cells (side 2h) around a particle: ~1000 candidates for ~20 real
// neighbours (SURVEY hard part 1). Its observable contract -- cell keys, the stable sort
// permutation, the cell table, candidate counts, the support test -- is defined on that grid, but
// nothing forces the ARRAYS IN HBM to be kept in the reference's order. Here they are sorted by
//     sub-cell key = (Morton cell key << 3) | octant of the cell (which half along x, y, z),
// i.e. by the Morton code of a grid of side h. Consequences:
//   * the particles of a reference cell are still contiguous, in the same segment [start, end) as
//     in the reference's order -- only permuted inside it -- so cell table, sorted keys and
//     candidate counts are unchanged;
//   * the particles of each sub-cell are contiguous too, so a particle only visits the 3 x 3 x 3
//     (occasionally 4 along an axis, see sub_bounds) sub-cells around it: ~130 candidates instead
//     of ~1000, with no culling work at all, one thread per particle, no shared memory needed;
//   * the reference's array order is carried as one uint32 per particle, `rrank` = the index the
//     particle has in the reference's array. The reference's next order is the stable sort of its
//     current order by cell key, so  rrank' = cell_start + #{j in the same cell : rrank_j < rrank_i},
//     a count over the ~40 particles of the cell (k_rank). Downloads and taps scatter through it,
//     so everything observable is in the reference's order, bit for bit;
//   * inside a sub-cell the particles are kept in the reference's order too (k_reorder_sub), so the
//     arrays -- and with them the order of every floating-point sum -- depend on the state alone:
//     resident sub-steps equal host round trips bitwise, and across GPUs (order keys instead of the
//     absolute rank, ownership by the slab planes, see dist.cu) the decomposition is bitwise transparent.
//
// Replaces, like neighbors.cu, kernels/sph.cl:9-62 with forces.cl:15-112; the force pass proper is
// k_forces_lists of neighbors.cu (it only consumes the lists written here).
#include "kernels.cuh"
#include "pair_terms.cuh"
#include "subview.cuh"

namespace clsph {

namespace {

constexpr int kSubThreads = 128;

}  // namespace

// =============================================================================================
// Sub-cell table: zero at the start of a sub-step (k_grid_setup), then either the counting sort's own table
// (sort.cu) or written by the gather kernel from the sorted keys.
// =============================================================================================
// Sorted slot r takes the particle that sat at vals[r]. skey gets the CELL key (what every other
// kernel and the exported grid_index expect); rr_dst the particle's reference rank of the previous
// sub-step. Slots where the sub-cell key changes fill the octant boundaries of the dense table:
// lb[cell][o] = first index whose octant is >= o, lb[cell][8] = end of the cell; untouched (empty)
// cells stay [0, 0).
// (the body of k_reorder_sub for one sorted slot r < n; *left_out = slots of the same sub-cell before r)
__device__ __forceinline__ void
reorder_sub_slot(uint32_t r, uint32_t n, const float4* __restrict__ src_pos, const float4* __restrict__ src_vel,
                 const float4* __restrict__ src_ivel, float4* __restrict__ dst_pos, float4* __restrict__ dst_vel,
                 float4* __restrict__ dst_ivel, const uint32_t* __restrict__ keys_a, const uint32_t* __restrict__ keys_b,
                 const uint32_t* __restrict__ vals_a, const uint32_t* __restrict__ vals_b, uint32_t* __restrict__ skey,
                 const uint32_t* __restrict__ rr_src, uint32_t* __restrict__ rr_dst, uint32_t* __restrict__ sub_lb,
                 const GridState* __restrict__ grid, const uint32_t* __restrict__ src_pid, uint32_t* __restrict__ dst_pid,
                 const uint32_t* __restrict__ src_ordk, const uint32_t* __restrict__ src_ordr, uint32_t* __restrict__ dst_ordk,
                 uint32_t* __restrict__ dst_ordr, uint32_t* left_out) {
  const bool in_b = (grid->sort_passes & 1u) != 0u;
  const uint32_t* __restrict__ keys = in_b ? keys_b : keys_a;
  const uint32_t* __restrict__ vals = in_b ? vals_b : vals_a;
  const uint32_t fkey = keys[r];
  const uint32_t from = vals[r];
  // Slot inside the sub-cell. The radix sort leaves the particles of a sub-cell in the order of the
  // previous internal arrays, which depends on the history (e.g. on whether the state was re-uploaded).
  // With the reference rank at hand (one GPU) they are put in the reference's order instead: the slot is
  // the number of particles of the same sub-cell with a smaller rank, counted over the handful of equal
  // keys around r. The arrays, hence every floating-point sum, then depend on the state alone: k resident
  // sub-steps are bitwise equal to k host round trips, as in the established organisation.
  uint32_t dest = r;
  if (rr_src) {
    const uint32_t mine = rr_src[from];
    uint32_t left = 0, smaller = 0;
    for (uint32_t q = r; q > 0 && keys[q - 1] == fkey; --q) {
      ++left;
      smaller += rr_src[vals[q - 1]] < mine ? 1u : 0u;
    }
    for (uint32_t q = r + 1; q < n && keys[q] == fkey; ++q) smaller += rr_src[vals[q]] < mine ? 1u : 0u;
    dest = r - left + smaller;
    rr_dst[dest] = mine;
    *left_out = left;
  } else if (src_ordk) {
    // multi-GPU: the same with the order keys (cell key, rank in cell of the previous sub-step), which order
    // particles like the global reference rank does and which ghosts carry too. Every rank then holds the
    // particles of a sub-cell in the order a single GPU would, all sums run in the same order, and the
    // decomposition is bitwise transparent.
    const uint32_t mk = src_ordk[from], mr = src_ordr[from];
    uint32_t left = 0, smaller = 0;
    for (uint32_t q = r; q > 0 && keys[q - 1] == fkey; --q) {
      ++left;
      const uint32_t o = vals[q - 1], jk = src_ordk[o], jr = src_ordr[o];
      smaller += (jk < mk || (jk == mk && jr < mr)) ? 1u : 0u;
    }
    for (uint32_t q = r + 1; q < n && keys[q] == fkey; ++q) {
      const uint32_t o = vals[q], jk = src_ordk[o], jr = src_ordr[o];
      smaller += (jk < mk || (jk == mk && jr < mr)) ? 1u : 0u;
    }
    dest = r - left + smaller;
    *left_out = left;
  } else {
    uint32_t left = 0;
    for (uint32_t q = r; q > 0 && keys[q - 1] == fkey; --q) ++left;
    *left_out = left;
  }
  dst_pos[dest] = src_pos[from];
  dst_vel[dest] = src_vel[from];
  float4 iv = src_ivel[from];
  if (src_ordk) iv.w = 0.f;  // multi-GPU: the mark "advanced here" is k_integrate's to set again in this sub-step
  dst_ivel[dest] = iv;
  const uint32_t key = fkey >> 3, oct = fkey & 7u;
  skey[r] = key;  // the same for the whole sub-cell, whichever slot
  if (src_pid) dst_pid[dest] = src_pid[from];
  if (src_ordk) {  // multi-GPU: order keys (cell key and rank inside the cell of the previous sub-step)
    dst_ordk[dest] = src_ordk[from];
    dst_ordr[dest] = src_ordr[from];
  }
  if (!grid->sub_dense || grid->sort_passes == 0u) return;  // (a counting sort has left the finished table, sort.cu)
  const uint32_t count = grid->cell_count;  // keys are < count whenever the grid fits (see k_reorder)
  uint32_t* row = sub_lb + (size_t)key * 9u;
  if (r == 0) {
    if (key < count)
      for (uint32_t o = 0; o <= oct; ++o) row[o] = 0u;
  } else {
    const uint32_t pf = keys[r - 1];
    const uint32_t pkey = pf >> 3, poct = pf & 7u;
    if (pkey != key) {
      if (key < count)
        for (uint32_t o = 0; o <= oct; ++o) row[o] = r;
      if (pkey < count) {
        uint32_t* prow = sub_lb + (size_t)pkey * 9u;
        for (uint32_t o = poct + 1u; o <= 8u; ++o) prow[o] = r;
      }
    } else if (poct != oct && key < count) {
      for (uint32_t o = poct + 1u; o <= oct; ++o) row[o] = r;
    }
  }
  if (r == n - 1 && key < count)
    for (uint32_t o = oct + 1u; o <= 8u; ++o) row[o] = n;
}

// Pair items (k_density_pairs): the particles of a sub-cell are handled two at a time by one thread, so every
// slot at an even offset inside its sub-cell announces the pair (itself, the next slot if that is in the same
// sub-cell). The list order is arbitrary -- warps append in arrival order -- and does not influence any result.
__global__ void __launch_bounds__(256)
k_reorder_sub(const float4* __restrict__ src_pos, const float4* __restrict__ src_vel, const float4* __restrict__ src_ivel,
              float4* __restrict__ dst_pos, float4* __restrict__ dst_vel, float4* __restrict__ dst_ivel,
              const uint32_t* __restrict__ keys_a, const uint32_t* __restrict__ keys_b, const uint32_t* __restrict__ vals_a,
              const uint32_t* __restrict__ vals_b, uint32_t* __restrict__ skey, const uint32_t* __restrict__ rr_src,
              uint32_t* __restrict__ rr_dst, uint32_t* __restrict__ sub_lb, const GridState* __restrict__ grid,
              const uint32_t* __restrict__ src_pid, uint32_t* __restrict__ dst_pid, const uint32_t* __restrict__ src_ordk,
              const uint32_t* __restrict__ src_ordr, uint32_t* __restrict__ dst_ordk, uint32_t* __restrict__ dst_ordr,
              uint32_t* __restrict__ pair_items, uint32_t* __restrict__ pair_count) {
  const uint32_t n = grid->n;
  const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t left = 1u;
  if (r < n)
    reorder_sub_slot(r, n, src_pos, src_vel, src_ivel, dst_pos, dst_vel, dst_ivel, keys_a, keys_b, vals_a, vals_b, skey, rr_src,
                     rr_dst, sub_lb, grid, src_pid, dst_pid, src_ordk, src_ordr, dst_ordk, dst_ordr, &left);
  if (!pair_items) return;
  // one append per CTA: the items of 256 consecutive slots stay together and in order, which keeps the threads
  // of a density CTA on neighbouring sub-cells (L1 reuse of the candidate rows)
  __shared__ uint32_t s_warp[8], s_base;
  const bool leader = r < n && (left & 1u) == 0u;
  const unsigned m = __ballot_sync(kFullMask, leader);
  const unsigned warp = threadIdx.x >> 5;
  if (lane_id() == 0u) s_warp[warp] = (uint32_t)__popc(m);
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t total = 0;
    for (int w = 0; w < 8; ++w) { const uint32_t c = s_warp[w]; s_warp[w] = total; total += c; }
    s_base = total ? atomicAdd(pair_count, total) : 0u;
  }
  __syncthreads();
  if (leader) {
    const uint32_t* __restrict__ keys = (grid->sort_passes & 1u) ? keys_b : keys_a;
    const bool second = r + 1u < n && keys[r + 1u] == keys[r];
    pair_items[s_base + s_warp[warp] + (uint32_t)__popc(m & lanemask_lt())] = r | (second ? 0x80000000u : 0u);
  }
}

// Reference rank after this sub-step's sort (see the header comment). Also writes the permutation
// tap in the reference's terms (sorted slot -> pre-step index) and, when asked, the pre-step keys.
__global__ void __launch_bounds__(kSubThreads)
k_rank(const uint32_t* __restrict__ skey, const uint32_t* __restrict__ rr_old, uint32_t* __restrict__ rr_new,
       const uint32_t* __restrict__ sub_lb, const uint32_t* __restrict__ keys_a, const uint32_t* __restrict__ keys_b,
       const GridState* __restrict__ grid, uint32_t* __restrict__ perm_out, uint32_t* __restrict__ keys_input_tap) {
  const GridState g = *grid;
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= g.n) return;
  const SubView v = make_view(g, sub_lb, keys_a, keys_b);
  const uint32_t key = skey[i];
  const uint2 cell = sub_range(v, key, 0u, 7u);
  const uint32_t mine = rr_old[i];
  uint32_t before = 0;
  for (uint32_t j = cell.x; j < cell.y; ++j) before += (__ldg(rr_old + j) < mine) ? 1u : 0u;
  const uint32_t rank = cell.x + before;
  rr_new[i] = rank;
  perm_out[rank] = mine;
  if (keys_input_tap) keys_input_tap[mine] = key;
}

// Multi-GPU form of k_rank. Across ranks an absolute index in the reference's global array would need
// a global scan every sub-step; the position is kept instead as the pair (cell key, rank inside the
// cell), which orders particles exactly like the global index does (cell starts grow with the key). A
// cell is owned by one rank and all its particles are there, so the rank inside the cell is a local
// count over the previous pairs, compared lexicographically. The global array, when wanted, is the
// merge of the ranks' downloads by (grid_index, rank in cell): clsph_dist_download.
__global__ void __launch_bounds__(kSubThreads)
k_rank_pair(const float4* __restrict__ pos, const uint32_t* __restrict__ skey, const uint32_t* __restrict__ ordk,
            const uint32_t* __restrict__ ordr, uint32_t* __restrict__ wrank, const uint32_t* __restrict__ sub_lb,
            const uint32_t* __restrict__ keys_a, const uint32_t* __restrict__ keys_b, const GridState* __restrict__ grid) {
  const GridState g = *grid;
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= g.n) return;
  const uint32_t key = skey[i];
  // The count runs over every particle of the cell held here, owned or ghost: a cell cut by a slab plane
  // has its other half among the ghosts (it is 2h wide, the ghosts reach 2h beyond the plane), with keys.
  if (!owned_here(pos[i].x, key, g)) {  // ghost copies get their rank from their owner
    wrank[i] = 0u;
    return;
  }
  const SubView v = make_view(g, sub_lb, keys_a, keys_b);
  const uint2 cell = sub_range(v, key, 0u, 7u);
  const uint32_t mk = ordk[i], mr = ordr[i];
  uint32_t before = 0;
  for (uint32_t j = cell.x; j < cell.y; ++j) {
    const uint32_t jk = __ldg(ordk + j), jr = __ldg(ordr + j);
    before += (jk < mk || (jk == mk && jr < mr)) ? 1u : 0u;
  }
  wrank[i] = before;
}

// =============================================================================================
// Density + Tait pressure + neighbour lists, one thread per particle.
// nlist[i * list_rows + e] = e-th neighbour of particle i (indices into the sorted arrays);
// ncount[i] = neighbours found, more than list_rows = list incomplete (k_forces_sub redoes it).
// =============================================================================================
// kMerged: both index ranges of a row of sub-cells are walked in one loop (for_each_row).
template <bool kTaps, bool kMerged>
__global__ void __launch_bounds__(kSubThreads)
k_density_sub(float4* pos, float4* vel, const uint32_t* __restrict__ skey, const uint32_t* __restrict__ sub_lb,
              const uint32_t* __restrict__ keys_a, const uint32_t* __restrict__ keys_b, const GridState* __restrict__ grid,
              const SphConst c, float4* __restrict__ aux, uint32_t* __restrict__ nlist, uint32_t* __restrict__ ncount,
              uint32_t list_rows, uint32_t* __restrict__ cand_count, uint32_t* __restrict__ supp_count) {
  const GridState g = *grid;
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= g.n) return;
  const SubView v = make_view(g, sub_lb, keys_a, keys_b);
  const uint32_t key = skey[i];
  const float4 pi = pos[i];
  // multi-GPU: the particles of the slab and the ghosts within h of it get a density (they are the
  // neighbours of owned particles); ghosts farther out only supply candidates (ghost depth: k_dist_classify)
  if (!(pi.x >= g.plane_lo - c.h_margin && pi.x < g.plane_hi + c.h_margin)) return;
  uint32_t* row = nlist + (size_t)i * list_rows;
#ifndef CLSPH_EMU
  // keep the row address in registers: rebuilt from nlist + i * list_rows + cnt it costs four
  // integer instructions per candidate instead of one
  asm volatile("" : "+l"(row));
#endif
  float acc = 0.f;   // sum of (h^2 - s)^3 over the support
  uint32_t cnt = 0;
  if (kMerged) {
    for_each_row(v, g, c, pi, [&](uint32_t a0, uint32_t a1, uint32_t b0, uint32_t b1) {
      const uint32_t total = (a1 - a0) + (b1 - b0);
      uint32_t j = a0 < a1 ? a0 : b0;
      for (uint32_t k = 0; k < total; ++k) {
        const float4 pj = pos[j];
        const float s = dist2_contract(pi.x, pi.y, pi.z, pj.x, pj.y, pj.z);
        const bool inside = s < c.support_s;
        const float t = inside ? c.h2 - s : 0.f;
        acc = fmaf(t * t, t, acc);
        store_if(inside && cnt < list_rows, row + cnt, j);
        cnt += inside ? 1u : 0u;
        ++j;
        if (j == a1) j = b0;  // end of the first range: continue in the second
      }
    });
  } else {
    for_each_neighbour(v, g, c, pos, pi, [&](uint32_t j, const float4&, float s, bool inside) {
      const float t = inside ? c.h2 - s : 0.f;
      acc = fmaf(t * t, t, acc);
      store_if(inside && cnt < list_rows, row + cnt, j);
      cnt += inside ? 1u : 0u;
    });
  }
  finish_density(c, acc, i, aux, pos, vel);
  ncount[i] = cnt;
  if (kTaps) {
    // the reference's candidate count: every particle of the 27 cells around this one (forces.cl:25-40)
    const uint32_t cx = compact10(key), cy = compact10(key >> 1), cz = compact10(key >> 2);
    uint32_t total = 0;
    if (cx != 0u && cy != 0u && cz != 0u) {  // with a 0 coordinate the reference's unsigned loop does not run
      for (uint32_t z = cz - 1u; z <= cz + 1u; ++z)
        for (uint32_t y = cy - 1u; y <= cy + 1u; ++y)
          for (uint32_t x = cx - 1u; x <= cx + 1u; ++x) {
            const uint2 r = sub_range(v, morton3(x, y, z), 0u, 7u);
            total += r.y - r.x;
          }
    }
    cand_count[i] = total;
    supp_count[i] = cnt;
  }
}

// =============================================================================================
// The same pass with TWO particles of a sub-cell per thread (items of k_reorder_sub). Both see the same rows of
// sub-cells, so every candidate is loaded once and tested against both with packed fp32 instructions (FADD2,
// FMUL2, FFMA2: one issue slot per two tests, the candidate a broadcast scalar operand). Each lane rounds like
// the scalar code and every particle still meets its candidates in the same order, so densities, lists and
// counts are BITWISE those of k_density_sub<.., kMerged>. The search window is the union of the two particles'
// windows (they differ only within 2^-10 h of a sub-cell boundary); a candidate outside a particle's own window
// is outside its support and adds an exact zero.
// =============================================================================================
// kWalk: how a thread walks a row's candidates -- 0: one after the other; 1: the same with the next candidate's
// load issued before the current one is tested; 2: four loads issued, then four tests. kStore2: list entries are
// stored two at a time (one 8-byte store per two hits) instead of one by one.
template <bool kTaps, int kWalk, bool kStore2>
__global__ void __launch_bounds__(kSubThreads, 8)  // 64 registers: eight CTAs of four warps per SM
k_density_pairs(float4* pos, float4* vel, const uint32_t* __restrict__ skey, const uint32_t* __restrict__ sub_lb,
                const uint32_t* __restrict__ keys_a, const uint32_t* __restrict__ keys_b, const GridState* __restrict__ grid,
                const SphConst c, float4* __restrict__ aux, uint32_t* __restrict__ nlist, uint32_t* __restrict__ ncount,
                uint32_t list_rows, uint32_t* __restrict__ cand_count, uint32_t* __restrict__ supp_count,
                const uint32_t* __restrict__ pair_items, const uint32_t* __restrict__ pair_count) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= *pair_count) return;
  const GridState g = *grid;
  const SubView v = make_view(g, sub_lb, keys_a, keys_b);
  const uint32_t item = pair_items[t];
  const uint32_t i0 = item & 0x7FFFFFFFu;
  const bool two = (item >> 31) != 0u;
  const uint32_t i1 = two ? i0 + 1u : i0;
  const float4 p0 = pos[i0], p1 = pos[i1];
  // multi-GPU: the particles of the slab and the ghosts within h of it get a density (k_density_sub)
  const bool need0 = p0.x >= g.plane_lo - c.h_margin && p0.x < g.plane_hi + c.h_margin;
  const bool need1 = two && p1.x >= g.plane_lo - c.h_margin && p1.x < g.plane_hi + c.h_margin;
  if (!need0 && !need1) return;
  uint32_t* row0 = nlist + (size_t)i0 * list_rows;
#ifndef CLSPH_EMU
  asm volatile("" : "+l"(row0));  // keep the row address in registers (see k_density_sub)
#endif
  uint32_t* row1 = row0 + (two ? list_rows : 0u);
#ifndef CLSPH_EMU
  asm volatile("" : "+l"(row1));
#endif
  const global_row_t grow0 = global_row(row0), grow1 = global_row(row1);
  // union of the two search windows
  uint32_t xlo, xhi, ylo, yhi, zlo, zhi;
  sub_bounds(p0.x, g.min_x, g.cell, c.h_margin, xlo, xhi);
  sub_bounds(p0.y, g.min_y, g.cell, c.h_margin, ylo, yhi);
  sub_bounds(p0.z, g.min_z, g.cell, c.h_margin, zlo, zhi);
  if (two) {
    uint32_t lo, hi;
    sub_bounds(p1.x, g.min_x, g.cell, c.h_margin, lo, hi); xlo = min(xlo, lo); xhi = max(xhi, hi);
    sub_bounds(p1.y, g.min_y, g.cell, c.h_margin, lo, hi); ylo = min(ylo, lo); yhi = max(yhi, hi);
    sub_bounds(p1.z, g.min_z, g.cell, c.h_margin, lo, hi); zlo = min(zlo, lo); zhi = max(zhi, hi);
  }
  const f32x2 X = f2_make(p0.x, p1.x), Y = f2_make(p0.y, p1.y), Z = f2_make(p0.z, p1.z);
  const f32x2 H2 = f2_bcast(c.h2);
  f32x2 acc = f2_make(0.f, 0.f);  // sums of (h^2 - s)^3 over the two supports
  uint32_t cnt0 = 0, cnt1 = 0, held0 = 0, held1 = 0;
  auto test = [&](const float4& pj, uint32_t j) {
    const f32x2 dx = f2_sub(X, f2_bcast(pj.x)), dy = f2_sub(Y, f2_bcast(pj.y)), dz = f2_sub(Z, f2_bcast(pj.z));
    const f32x2 s = f2_fma(dz, dz, f2_fma(dy, dy, f2_mul(dx, dx)));
    const f32x2 d = f2_sub(H2, s);
    const bool hit0 = f2_lo(s) < c.support_s, hit1 = f2_hi(s) < c.support_s;
    const f32x2 w = f2_make(hit0 ? f2_lo(d) : 0.f, hit1 ? f2_hi(d) : 0.f);
    acc = f2_fma(f2_mul(w, w), w, acc);
    const bool in0 = hit0, in1 = hit1;
    if (kStore2) {
      // List entries leave two at a time: every lane's store is an L1 wavefront and a 32-byte L2 sector write of its
      // own (each lane writes its own row), so one 8-byte store per two hits halves that. The first hit of a pair
      // waits in a register (list_rows is even, rows are 8-byte aligned).
      const bool odd0 = (cnt0 & 1u) != 0u, odd1 = (cnt1 & 1u) != 0u;
      row_store2_if(in0 && odd0 && cnt0 < list_rows, grow0, cnt0 - 1u, held0, j);
      held0 = in0 ? j : held0;  // (after an odd hit the held value is not used again before the next even one replaces it)
      row_store2_if(in1 && odd1 && cnt1 < list_rows, grow1, cnt1 - 1u, held1, j);
      held1 = in1 ? j : held1;
    } else {
      row_store_if(in0 && cnt0 < list_rows, grow0, cnt0, j);
      row_store_if(in1 && cnt1 < list_rows, grow1, cnt1, j);
    }
    cnt0 += in0 ? 1u : 0u;
    cnt1 += in1 ? 1u : 0u;
  };
  // The row's two index ranges laid end to end. A thread walks ~160 candidates one after the other; with one load
  // in flight the pass waits on L1 / L2 latency, hence the variants.
  auto walk = [&](uint32_t a0, uint32_t a1, uint32_t b0, uint32_t b1) {
    const uint32_t la = a1 - a0, total = la + (b1 - b0);
    if (kWalk == 2) {
      const uint32_t shift = b0 - la;  // index = k + shift in the second range
      uint32_t k = 0;
      for (; k + 4u <= total; k += 4u) {
        const uint32_t j0 = k < la ? a0 + k : k + shift, j1 = k + 1u < la ? a0 + k + 1u : k + 1u + shift,
                       j2 = k + 2u < la ? a0 + k + 2u : k + 2u + shift, j3 = k + 3u < la ? a0 + k + 3u : k + 3u + shift;
        const float4 q0 = pos[j0], q1 = pos[j1], q2 = pos[j2], q3 = pos[j3];
        test(q0, j0);
        test(q1, j1);
        test(q2, j2);
        test(q3, j3);
      }
      for (; k < total; ++k) {
        const uint32_t j = k < la ? a0 + k : k + shift;
        test(pos[j], j);
      }
    } else if (kWalk == 1) {
      if (total == 0u) return;
      uint32_t j = a0 < a1 ? a0 : b0;
      float4 cur = pos[j];
      for (uint32_t k = 1; k < total; ++k) {
        uint32_t jn = j + 1u;
        if (jn == a1) jn = b0;  // end of the first range: continue in the second
        const float4 nxt = pos[jn];
        test(cur, j);
        cur = nxt;
        j = jn;
      }
      test(cur, j);
    } else {
      uint32_t j = a0 < a1 ? a0 : b0;
      for (uint32_t k = 0; k < total; ++k) {
        test(pos[j], j);
        ++j;
        if (j == a1) j = b0;
      }
    }
  };
  // Rows of sub-cells, z outermost. The index ranges of the NEXT row are looked up (four loads from the sub-cell
  // table) before the current row is walked, so that a thread does not start every row waiting for them.
  const uint32_t cx_lo = xlo >> 1, cx_hi = xhi >> 1;
  const uint32_t ny = yhi - ylo + 1u, n_rows = (zhi - zlo + 1u) * ny;
  auto row_ranges = [&](uint32_t r, uint2& a, uint2& b) {
    const uint32_t fz = zlo + r / ny, fy = ylo + r % ny;
    const uint32_t kzy = (spread10(fz >> 1) << 2) | (spread10(fy >> 1) << 1), ozy = ((fz & 1u) << 2) | ((fy & 1u) << 1);
    a = sub_range(v, kzy | spread10(cx_lo), ozy | (xlo & 1u), ozy | (cx_hi == cx_lo ? (xhi & 1u) : 1u));
    b = make_uint2(0u, 0u);
    if (cx_hi > cx_lo) b = sub_range(v, kzy | spread10(cx_lo + 1u), ozy, ozy | (cx_hi == cx_lo + 1u ? (xhi & 1u) : 1u));
  };
  uint2 na, nb;
  row_ranges(0u, na, nb);
  for (uint32_t r = 0; r < n_rows; ++r) {
    const uint2 a = na, b = nb;
    if (r + 1u < n_rows) row_ranges(r + 1u, na, nb);
    walk(a.x, a.y, b.x, b.y);
    if (cx_hi > cx_lo + 1u) {  // rare third cell of the row
      const uint32_t fz = zlo + r / ny, fy = ylo + r % ny;
      const uint32_t kzy = (spread10(fz >> 1) << 2) | (spread10(fy >> 1) << 1), ozy = ((fz & 1u) << 2) | ((fy & 1u) << 1);
      const uint2 e = sub_range(v, kzy | spread10(cx_hi), ozy, ozy | (xhi & 1u));
      walk(e.x, e.y, 0u, 0u);
    }
  }
  if (kStore2) {  // the last hit of an odd count is still held
    if ((cnt0 & 1u) && cnt0 - 1u < list_rows) row0[cnt0 - 1u] = held0;
    if (two && (cnt1 & 1u) && cnt1 - 1u < list_rows) row1[cnt1 - 1u] = held1;
  }
  if (need0) {
    finish_density(c, f2_lo(acc), i0, aux, pos, vel);
    ncount[i0] = cnt0;
  }
  if (need1) {
    finish_density(c, f2_hi(acc), i1, aux, pos, vel);
    ncount[i1] = cnt1;
  }
  // lists that did not fit (rare): k_forces_sub only looks for them when there are any
  if ((need0 && cnt0 > list_rows) || (need1 && cnt1 > list_rows)) atomicAdd(const_cast<uint32_t*>(pair_count) + 1, 1u);
  if (kTaps) {
    // the reference's candidate count: every particle of the 27 cells around this one (forces.cl:25-40);
    // both particles are in the same cell
    const uint32_t key = skey[i0];
    const uint32_t cx = compact10(key), cy = compact10(key >> 1), cz = compact10(key >> 2);
    uint32_t total = 0;
    if (cx != 0u && cy != 0u && cz != 0u) {  // with a 0 coordinate the reference's unsigned loop does not run
      for (uint32_t z = cz - 1u; z <= cz + 1u; ++z)
        for (uint32_t y = cy - 1u; y <= cy + 1u; ++y)
          for (uint32_t x = cx - 1u; x <= cx + 1u; ++x) {
            const uint2 r = sub_range(v, morton3(x, y, z), 0u, 7u);
            total += r.y - r.x;
          }
    }
    if (need0) { cand_count[i0] = total; supp_count[i0] = cnt0; }
    if (need1) { cand_count[i1] = total; supp_count[i1] = cnt1; }
  }
}

// Forces for the particles whose list overflowed: same traversal, pair terms evaluated in place.
__global__ void __launch_bounds__(kSubThreads)
k_forces_sub(const float4* __restrict__ pos, const float4* __restrict__ vel, const float4* __restrict__ aux,
             const uint32_t* __restrict__ skey, const uint32_t* __restrict__ sub_lb, const uint32_t* __restrict__ keys_a,
             const uint32_t* __restrict__ keys_b, const GridState* __restrict__ grid, const SphConst c,
             const uint32_t* __restrict__ ncount, uint32_t list_rows, float4* __restrict__ accel,
             const uint32_t* __restrict__ overflowed) {
  if (overflowed && *overflowed == 0u) return;  // the density pass found no list that overflowed
  const GridState g = *grid;
  const SubView v = make_view(g, sub_lb, keys_a, keys_b);
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < g.n; i += gridDim.x * blockDim.x) {
    if (!(ncount[i] > list_rows)) continue;
    const float4 pi = pos[i];
    if (!owned_here(pi.x, skey[i], g)) continue;  // multi-GPU: ghosts get no force
    const float4 vi = vel[i];
    ForceSums sums;
    for_each_neighbour(v, g, c, pos, pi, [&](uint32_t j, const float4& pj, float, bool inside) {
      if (inside) add_pair(sums, c, j == i, pi, vi, pi.w, pj, vel[j]);  // rare path: the exact pair terms
    });
    accel[i] = finish_force(sums, c, aux[i].x);
  }
}

// dst[rrank[i]] = src[i], `words` 32-bit words per item: internal order -> the reference's order.
__global__ void __launch_bounds__(256) k_scatter_words(const uint32_t* __restrict__ src, const uint32_t* __restrict__ rrank,
                                                       uint32_t* __restrict__ dst, uint32_t n, uint32_t words) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t at = rrank[i];
  for (uint32_t w = 0; w < words; ++w) dst[(size_t)at * words + w] = src[(size_t)i * words + w];
}

// ---------------------------------------------------------------------------------------------
void launch_reorder_sub(const StateArrays& src, const StateArrays& dst, const SortBuffers& sort, uint32_t* skey,
                        const uint32_t* rr_src, uint32_t* rr_dst, uint32_t* sub_lb, const GridState* grid,
                        const uint32_t* src_pid, uint32_t* dst_pid, const uint32_t* src_ordk, const uint32_t* src_ordr,
                        uint32_t* dst_ordk, uint32_t* dst_ordr, uint32_t* pair_items, uint32_t* pair_count, uint32_t n_launch,
                        cudaStream_t stream, uint64_t* launches) {
  // (pair_count -- items, overflowing lists -- was zeroed by the launch of k_grid_setup)
  k_reorder_sub<<<(n_launch + 255) / 256, 256, 0, stream>>>(src.pos, src.vel, src.ivel, dst.pos, dst.vel, dst.ivel, sort.keys_a,
                                                            sort.keys_b, sort.vals_a, sort.vals_b, skey, rr_src, rr_dst, sub_lb,
                                                            grid, src_pid, dst_pid, src_ordk, src_ordr, dst_ordk, dst_ordr,
                                                            pair_items, pair_count);
  if (launches) ++*launches;
}

void launch_rank(const uint32_t* skey, const uint32_t* rr_old, uint32_t* rr_new, const uint32_t* sub_lb,
                 const SortBuffers& sort, const GridState* grid, uint32_t* perm_out, uint32_t* keys_input_tap,
                 uint32_t n_launch, cudaStream_t stream, uint64_t* launches) {
  k_rank<<<(n_launch + kSubThreads - 1) / kSubThreads, kSubThreads, 0, stream>>>(skey, rr_old, rr_new, sub_lb, sort.keys_a,
                                                                                 sort.keys_b, grid, perm_out, keys_input_tap);
  if (launches) ++*launches;
}

void launch_rank_pair(const float4* pos, const uint32_t* skey, const uint32_t* ordk, const uint32_t* ordr, uint32_t* wrank,
                      const uint32_t* sub_lb, const SortBuffers& sort, const GridState* grid, uint32_t n_launch,
                      cudaStream_t stream, uint64_t* launches) {
  k_rank_pair<<<(n_launch + kSubThreads - 1) / kSubThreads, kSubThreads, 0, stream>>>(pos, skey, ordk, ordr, wrank, sub_lb,
                                                                                      sort.keys_a, sort.keys_b, grid);
  if (launches) ++*launches;
}

void launch_density_sub(float4* pos, float4* vel, const uint32_t* skey, const uint32_t* sub_lb, const SortBuffers& sort,
                        const GridState* grid, const SphConst& c, float4* aux, const NeighbourLists& lists,
                        const DebugTaps& taps, bool debug, bool merged, uint32_t n_launch, cudaStream_t stream,
                        uint64_t* launches) {
  const unsigned blocks = (n_launch + kSubThreads - 1) / kSubThreads;
  uint32_t* cand = debug ? taps.candidate_count : nullptr;
  uint32_t* supp = debug ? taps.support_count : nullptr;
  if (debug && merged)
    k_density_sub<true, true><<<blocks, kSubThreads, 0, stream>>>(pos, vel, skey, sub_lb, sort.keys_a, sort.keys_b, grid, c, aux,
                                                                  lists.entries, lists.count, lists.rows, cand, supp);
  else if (debug)
    k_density_sub<true, false><<<blocks, kSubThreads, 0, stream>>>(pos, vel, skey, sub_lb, sort.keys_a, sort.keys_b, grid, c, aux,
                                                                   lists.entries, lists.count, lists.rows, cand, supp);
  else if (merged)
    k_density_sub<false, true><<<blocks, kSubThreads, 0, stream>>>(pos, vel, skey, sub_lb, sort.keys_a, sort.keys_b, grid, c, aux,
                                                                   lists.entries, lists.count, lists.rows, cand, supp);
  else
    k_density_sub<false, false><<<blocks, kSubThreads, 0, stream>>>(pos, vel, skey, sub_lb, sort.keys_a, sort.keys_b, grid, c, aux,
                                                                    lists.entries, lists.count, lists.rows, cand, supp);
  if (launches) ++*launches;
}

namespace {
struct PairArgs {
  float4 *pos, *vel;
  const uint32_t *skey, *sub_lb, *keys_a, *keys_b;
  const GridState* grid;
  SphConst c;
  float4* aux;
  uint32_t *nlist, *ncount;
  uint32_t list_rows;
  uint32_t *cand, *supp;
  const uint32_t *pair_items, *pair_count;
  unsigned blocks;
  cudaStream_t stream;
};
template <bool kTaps, int kWalk, bool kStore2>
void launch_pairs_variant(const PairArgs& a) {
  k_density_pairs<kTaps, kWalk, kStore2><<<a.blocks, kSubThreads, 0, a.stream>>>(a.pos, a.vel, a.skey, a.sub_lb, a.keys_a, a.keys_b, a.grid, a.c, a.aux,
                                                                                 a.nlist, a.ncount, a.list_rows, a.cand, a.supp, a.pair_items,
                                                                                 a.pair_count);
}
}  // namespace

void launch_density_pairs(float4* pos, float4* vel, const uint32_t* skey, const uint32_t* sub_lb, const SortBuffers& sort,
                          const GridState* grid, const SphConst& c, float4* aux, const NeighbourLists& lists,
                          const DebugTaps& taps, bool debug, int variant, const uint32_t* pair_items, const uint32_t* pair_count,
                          uint32_t n_launch, cudaStream_t stream, uint64_t* launches) {
  // the item count lives on the device (between n / 2 and n): sized for the worst case, surplus blocks leave at once
  const unsigned blocks = (n_launch + kSubThreads - 1) / kSubThreads;
  uint32_t* cand = debug ? taps.candidate_count : nullptr;
  uint32_t* supp = debug ? taps.support_count : nullptr;
  const PairArgs a{pos, vel, skey, sub_lb, sort.keys_a, sort.keys_b, grid, c, aux, lists.entries, lists.count, lists.rows, cand, supp,
                   pair_items, pair_count, blocks, stream};
  if (debug) {
    launch_pairs_variant<true, 2, true>(a);
  } else {
    switch (variant) {
      case 0: launch_pairs_variant<false, 0, false>(a); break;
      case 1: launch_pairs_variant<false, 1, false>(a); break;
      case 2: launch_pairs_variant<false, 2, false>(a); break;
      case 3: launch_pairs_variant<false, 0, true>(a); break;
      case 4: launch_pairs_variant<false, 1, true>(a); break;
      default: launch_pairs_variant<false, 2, true>(a); break;
    }
  }
  if (launches) ++*launches;
}

void launch_forces_sub_overflow(const float4* pos, const float4* vel, const float4* aux, const uint32_t* skey,
                                const uint32_t* sub_lb, const SortBuffers& sort, const GridState* grid, const SphConst& c,
                                const NeighbourLists& lists, float4* accel, const uint32_t* overflowed, uint32_t n_launch,
                                cudaStream_t stream, uint64_t* launches) {
  // a resident grid striding over the particles: usually every CTA returns at its first instruction
  const unsigned blocks = std::max(1u, std::min((n_launch + kSubThreads - 1) / kSubThreads, 148u * 16u));
  k_forces_sub<<<blocks, kSubThreads, 0, stream>>>(pos, vel, aux, skey, sub_lb, sort.keys_a, sort.keys_b, grid, c, lists.count, lists.rows,
                                                   accel, overflowed);
  if (launches) ++*launches;
}

void launch_scatter_words(const void* src, const uint32_t* rrank, void* dst, uint32_t n, uint32_t words,
                          cudaStream_t stream, uint64_t* launches) {
  k_scatter_words<<<(n + 255) / 256, 256, 0, stream>>>((const uint32_t*)src, rrank, (uint32_t*)dst, n, words);
  if (launches) ++*launches;
}

}  // namespace clsph
