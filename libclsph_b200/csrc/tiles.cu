// tiles.cu -- the two neighbour passes of the sub-cell order as TILE kernels.
//
// Replaces, like neighbors.cu / subgrid.cu, kernels/sph.cl:9-62 with forces.cl:15-112 and smoothing.cl
// of the reference; same candidate sets (a subset of the reference's 27 cells that provably holds every
// particle inside the support), same support test s < support_s, same pair formulas.
//
// ncu of the per-particle kernels (profiles/r02_a_summary.md) showed where their time went: 23 SASS
// instructions per candidate, a third of them address arithmetic, range switching and list bookkeeping;
// one 16-byte global load per candidate per lane at lane-divergent addresses (5 sectors per request); and one
// scattered 4-byte store per neighbour found (23 M sectors written to L2 for 20 M list entries). Here:
//
//   * the work unit is a BLOCK of 2 x 2 x 2 grid cells = 4 x 4 x 4 sub-cells of side h. In the Morton order
//     of the sorted arrays a block is 8 consecutive cells, so its particles are one contiguous range;
//   * a CTA stages the positions of the 6 x 6 x 6 sub-cells around the block (the block plus one sub-cell of
//     halo) in shared memory with bulk asynchronous copies (cp.async.bulk, one per sub-cell, completion on an
//     mbarrier), laid out ROW-MAJOR (z, y, x): the three sub-cells a particle visits along x in one (z, y)
//     row are one contiguous range of slots -- 9 ranges per particle instead of 18 -- and every row ends in
//     three far-away sentinel slots, so rows are walked four candidates at a time without a tail;
//   * a thread takes TWO particles of the same sub-cell (same candidate ranges): one 16-byte shared-memory
//     load per candidate serves two distance tests;
//   * hits are recorded as one bit per candidate in a 64-bit mask per row (9 masks per particle, written with
//     coalesced stores) instead of index lists: the force pass rebuilds the same rows from the same table;
//   * the force pass (k_forces_tiles) stages positions and velocities of the same region, expands each
//     particle's masks into a list of 16-bit slot numbers in shared memory and runs one thread per particle
//     over its list: both gathers of a pair come from shared memory.
//
// Particles the tiles cannot serve exactly are handed to the per-particle kernels of subgrid.cu through a
// list (TileCtl::n_slow): particles within sub_delta of a sub-cell boundary (their neighbours may sit two
// sub-cells away after rounding, see wide_target), rows of more than 64 candidates, particles with more
// neighbours than the force list holds, and whole blocks whose region exceeds the staging capacity.
//
// Summation order: rows z-major, y, then x, candidates in array order -- the order of k_density_sub's walk,
// so densities are bit-identical to the per-particle kernel's, on one GPU and across a slab decomposition.
#include "kernels.cuh"
#include "pair_terms.cuh"
#include "subview.cuh"

namespace clsph {

namespace {

constexpr int kTileThreads = 128;
constexpr int kRegionSide = 6;                       // sub-cells per axis of the staged region
constexpr int kRowEntries = kRegionSide + 1;         // per (z, y) row: 6 sub-cells + 1 pad entry of sentinels
constexpr int kRows = kRegionSide * kRegionSide;     // 36
constexpr int kEntries = kRows * kRowEntries;        // 252
constexpr int kPadSlots = 3;                         // sentinel slots after each row
constexpr int kTargetCells = 64;                     // sub-cells of the block proper
constexpr float kSentinel = 1e18f;                   // |d|^2 ~ 3e36: finite, never inside a support

// Header of the dynamic shared memory (all tile kernels): tables, then the staged arrays.
struct TileTables {
  unsigned long long mbar;        // mbarrier of the bulk copies
  uint32_t block;                 // block id being processed (Morton code of the block coordinates), ~0u = none left
  uint32_t total;                 // staged slots, pads included
  uint32_t n_work;                // work items of the block: particle pairs (density) or particles (forces)
  uint32_t pad;
  uint32_t warp_sum[4];
  uint32_t soff[kEntries + 4];    // first slot of every region entry; [kEntries] = total
  uint32_t gstart[kEntries + 4];  // index of the entry's first particle in the sorted arrays
  uint32_t tcum[kTargetCells + 4];// exclusive running sum of work items over the 64 target sub-cells (memory order)
};

__device__ __forceinline__ uint32_t warp_inclusive_scan_u32(uint32_t v) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t t = __shfl_up_sync(kFullMask, v, o);
    if (lane_id() >= (unsigned)o) v += t;
  }
  return v;
}

// Exclusive scan of one value per thread over the 128 threads of the CTA; *total gets the sum.
__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t* warp_sum, uint32_t* total) {
  const unsigned warp = threadIdx.x >> 5;
  const uint32_t inc = warp_inclusive_scan_u32(v);
  if (lane_id() == 31u) warp_sum[warp] = inc;
  __syncthreads();
  uint32_t before = 0, all = 0;
#pragma unroll
  for (unsigned w = 0; w < kTileThreads / 32; ++w) {
    const uint32_t s = warp_sum[w];
    if (w < warp) before += s;
    all += s;
  }
  __syncthreads();
  *total = all;
  return before + inc - v;
}

// ---- bulk asynchronous copy global -> shared with mbarrier completion (UBLKCP in SASS) ------------------
#ifndef CLSPH_EMU
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
#endif
__device__ __forceinline__ void tile_barrier_init(unsigned long long* bar) {
#ifndef CLSPH_EMU
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"((uint32_t)kTileThreads) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
#else
  (void)bar;
#endif
  __syncthreads();
}
// Every thread announces the bytes its own copies will deliver and arrives; the phase completes when all
// 128 threads have arrived and all announced bytes have landed.
__device__ __forceinline__ void tile_barrier_arrive(unsigned long long* bar, uint32_t bytes) {
#ifndef CLSPH_EMU
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
#else
  (void)bar; (void)bytes;
#endif
}
__device__ __forceinline__ void tile_barrier_wait(unsigned long long* bar, uint32_t parity) {
#ifndef CLSPH_EMU
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "TILE_WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra TILE_DONE_%=;\n\t"
      "bra TILE_WAIT_%=;\n\t"
      "TILE_DONE_%=:\n\t}" ::"r"(smem_addr(bar)), "r"(parity)
      : "memory");
#else
  (void)bar; (void)parity;
  __syncthreads();  // the copies of the emulator build are synchronous: a block barrier orders them
#endif
}
// `count` float4 from global src to shared dst (both 16-byte aligned).
__device__ __forceinline__ void tile_copy(float4* dst, const float4* src, uint32_t count, unsigned long long* bar) {
#ifndef CLSPH_EMU
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_addr(dst)),
               "l"(__cvta_generic_to_global(src)), "r"(count * 16u), "r"(smem_addr(bar))
               : "memory");
#else
  (void)bar;
  for (uint32_t k = 0; k < count; ++k) dst[k] = src[k];
#endif
}
// Shared memory last touched by ordinary loads / stores is about to be written by the copy engine.
__device__ __forceinline__ void tile_fence_before_copies() {
#ifndef CLSPH_EMU
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
#endif
}

// Region entry (row-major, 7 entries per row) of target sub-cell t, t in the order of the sorted arrays:
// t = cell of the block (z y x bits) << 3 | octant of the cell (z y x bits).
__device__ __forceinline__ uint32_t target_entry(uint32_t t) {
  const uint32_t lx = 1u + ((t >> 3) & 1u) * 2u + (t & 1u);
  const uint32_t ly = 1u + ((t >> 4) & 1u) * 2u + ((t >> 1) & 1u);
  const uint32_t lz = 1u + ((t >> 5) & 1u) * 2u + ((t >> 2) & 1u);
  return (lz * kRegionSide + ly) * kRowEntries + lx;
}

// Fetches the next block (persistent CTAs pull from a counter) and fills the tables of its region:
// soff / gstart for the 252 entries, T.total. Returns false when no block is left. Ends with a barrier.
__device__ __forceinline__ bool tile_begin(TileTables& T, const SubView& v, const uint32_t* __restrict__ blist,
                                           const uint32_t* n_blocks, uint32_t* next_block, uint32_t (&my_count)[2],
                                           uint32_t (&my_entry)[2]) {
  if (threadIdx.x == 0) {
    const uint32_t b = atomicAdd(next_block, 1u);
    T.block = b < *n_blocks ? blist[b] : 0xFFFFFFFFu;
  }
  __syncthreads();
  const uint32_t block = T.block;
  if (block == 0xFFFFFFFFu) return false;
  // block coordinates: the block id is the cell key >> 3, itself a Morton code (of the halved cell coordinates)
  const int bx = (int)compact10(block), by = (int)compact10(block >> 1), bz = (int)compact10(block >> 2);
  uint32_t sum = 0;
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const uint32_t e = threadIdx.x * 2u + (uint32_t)k;
    my_entry[k] = e;
    uint32_t count = 0, first = 0;
    if (e < (uint32_t)kEntries) {
      const uint32_t row = e / kRowEntries, lx = e - row * kRowEntries;
      if (lx == (uint32_t)kRegionSide) {
        count = kPadSlots;
      } else {
        const uint32_t lz = row / kRegionSide, ly = row - lz * kRegionSide;
        const int sx = 4 * bx - 1 + (int)lx, sy = 4 * by - 1 + (int)ly, sz = 4 * bz - 1 + (int)lz;  // sub-cell coordinates
        if (sx >= 0 && sy >= 0 && sz >= 0 && sx < 2048 && sy < 2048 && sz < 2048) {
          const uint32_t key = morton3((uint32_t)sx >> 1, (uint32_t)sy >> 1, (uint32_t)sz >> 1);
          const uint32_t o = ((uint32_t)sx & 1u) | (((uint32_t)sy & 1u) << 1) | (((uint32_t)sz & 1u) << 2);
          const uint2 r = sub_range(v, key, o, o);
          first = r.x;
          count = r.y - r.x;
        }
      }
      T.gstart[e] = first;
    }
    my_count[k] = count;
    sum += count;
  }
  uint32_t total;
  const uint32_t before = block_exclusive_scan(sum, T.warp_sum, &total);
  if (my_entry[0] < (uint32_t)kEntries) T.soff[my_entry[0]] = before;
  if (my_entry[1] < (uint32_t)kEntries) T.soff[my_entry[1]] = before + my_count[0];
  if (threadIdx.x == 0) {
    T.soff[kEntries] = total;
    T.total = total;
  }
  __syncthreads();
  return true;
}

// Copies the region into shared memory: arrays[a][slot] <- src[a][particle]. Each thread issues the copies of
// its own two entries (sentinels for the pad entries) and waits for everybody's. `parity` flips per block.
template <int kArrays>
__device__ __forceinline__ void tile_stage(TileTables& T, float4* const (&dst)[kArrays], const float4* const (&src)[kArrays],
                                           const uint32_t (&my_count)[2], const uint32_t (&my_entry)[2], uint32_t parity) {
  tile_fence_before_copies();
  uint32_t bytes = 0;
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const uint32_t e = my_entry[k];
    if (e >= (uint32_t)kEntries || my_count[k] == 0u) continue;
    const uint32_t at = T.soff[e];
    if (e % kRowEntries == (uint32_t)kRegionSide) {
      const float4 far = make_float4(kSentinel, kSentinel, kSentinel, 0.f);
#pragma unroll
      for (int a = 0; a < kArrays; ++a)
        for (int q = 0; q < kPadSlots; ++q) dst[a][at + q] = far;
    } else {
      bytes += my_count[k] * 16u * kArrays;
    }
  }
  tile_barrier_arrive(&T.mbar, bytes);
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const uint32_t e = my_entry[k];
    if (e >= (uint32_t)kEntries || my_count[k] == 0u || e % kRowEntries == (uint32_t)kRegionSide) continue;
#pragma unroll
    for (int a = 0; a < kArrays; ++a) tile_copy(dst[a] + T.soff[e], src[a] + T.gstart[e], my_count[k], &T.mbar);
  }
  tile_barrier_wait(&T.mbar, parity);
  __syncthreads();  // the sentinel stores of other threads
}

// Largest t in [0, 64) with tcum[t] <= item (tcum ascending, tcum[0] = 0, item < tcum[64]).
__device__ __forceinline__ uint32_t find_target(const uint32_t* tcum, uint32_t item) {
  uint32_t t = 0;
#pragma unroll
  for (uint32_t step = 32; step > 0; step >>= 1)
    if (tcum[t + step] <= item) t += step;
  return t;
}

// A particle is "wide" when it lies within sub_delta (in units of h) of a boundary of its sub-cell along some
// axis. For every other particle i, a particle j inside its support has a sub-cell coordinate within +-1 of
// i's on every axis: with f = 2 fl(fl(p - min) / cell) as k_keys_hist computes it, |f_i - f_j| < 1 + sub_delta
// (a pair inside the support has |dx| < h (1 + 2^-21), and each f carries a relative error below 2^-23, see
// k_grid_setup), so floor(f_j) is in [floor(f_i) - 1, floor(f_i) + 1] whenever the fraction of f_i is in
// [sub_delta, 1 - sub_delta). Wide particles (a few in ten thousand) go to the per-particle kernel, whose
// search window is the proven one of sub_bounds.
__device__ __forceinline__ bool wide_axis(float p, float mn, float cell, float delta) {
  const float q = __fdiv_rn(__fsub_rn(p, mn), cell);
  const float f = __fadd_rn(q, q);
  const float frac = f - floorf(f);
  return !(frac >= delta && frac < 1.f - delta);
}
__device__ __forceinline__ bool wide_target(const float4& p, const GridState& g) {
  return wide_axis(p.x, g.min_x, g.cell, g.sub_delta) || wide_axis(p.y, g.min_y, g.cell, g.sub_delta) ||
         wide_axis(p.z, g.min_z, g.cell, g.sub_delta);
}

// One candidate against the two particles of the thread: support test, density term, hit bit.
__device__ __forceinline__ void test_pair(const float4& q, uint32_t bit, const float4& p0, const float4& p1, float support_s,
                                          float h2, float& acc0, float& acc1, uint32_t& m0, uint32_t& m1) {
  const float s0 = dist2_contract(p0.x, p0.y, p0.z, q.x, q.y, q.z);
  const float s1 = dist2_contract(p1.x, p1.y, p1.z, q.x, q.y, q.z);
  const float t0 = h2 - s0, t1 = h2 - s1;
  if (s0 < support_s) {
    acc0 = fmaf(t0 * t0, t0, acc0);
    m0 |= bit;
  }
  if (s1 < support_s) {
    acc1 = fmaf(t1 * t1, t1, acc1);
    m1 |= bit;
  }
}

// Up to 32 candidates starting at p (len > 0), four at a time; reads up to 3 slots past len (sentinels or
// particles two sub-cells away, never inside the support of a particle that is not wide).
__device__ __forceinline__ void walk32(const float4* p, uint32_t len, const float4& p0, const float4& p1, float support_s,
                                       float h2, float& acc0, float& acc1, uint32_t& m0, uint32_t& m1) {
#pragma unroll
  for (int ch = 0; ch < 8; ++ch) {
    if ((uint32_t)(4 * ch) >= len) break;
    const float4 q0 = p[4 * ch], q1 = p[4 * ch + 1], q2 = p[4 * ch + 2], q3 = p[4 * ch + 3];
    test_pair(q0, 1u << (4 * ch), p0, p1, support_s, h2, acc0, acc1, m0, m1);
    test_pair(q1, 2u << (4 * ch), p0, p1, support_s, h2, acc0, acc1, m0, m1);
    test_pair(q2, 4u << (4 * ch), p0, p1, support_s, h2, acc0, acc1, m0, m1);
    test_pair(q3, 8u << (4 * ch), p0, p1, support_s, h2, acc0, acc1, m0, m1);
  }
}

}  // namespace

// =============================================================================================
// Density + Tait pressure + hit masks.
// nmask[r * mask_stride + i] = hits of particle i in row r (bit b = the b-th slot of the row), ncount[i] =
// number of hits (the support count, self included); bit 31 of ncount = no masks, see kNoMasks.
// =============================================================================================
__global__ void __launch_bounds__(kTileThreads)
k_density_tiles(float4* pos, float4* vel, const uint32_t* __restrict__ sub_lb, const uint32_t* __restrict__ keys_a,
                const uint32_t* __restrict__ keys_b, const GridState* __restrict__ grid, const SphConst c,
                float4* __restrict__ aux, unsigned long long* __restrict__ nmask, size_t mask_stride, uint32_t* __restrict__ ncount,
                uint32_t list_cap, const uint32_t* __restrict__ blist, TileCtl* ctl, uint32_t* __restrict__ slow,
                uint32_t slot_cap) {
  extern __shared__ float4 tile_smem[];
  TileTables& T = *reinterpret_cast<TileTables*>(tile_smem);
  float4* cand = tile_smem + (sizeof(TileTables) + 15) / 16;
  const GridState g = *grid;
  const SubView v = make_view(g, sub_lb, keys_a, keys_b);
  const unsigned warp = threadIdx.x >> 5, lane = lane_id();
  tile_barrier_init(&T.mbar);
  uint32_t parity = 0;
  uint32_t my_count[2], my_entry[2];
  while (tile_begin(T, v, blist, &ctl->n_blocks, &ctl->next_density, my_count, my_entry)) {
    // work items: pairs of particles of the same target sub-cell
    uint32_t pairs = 0;
    if (threadIdx.x < (unsigned)kTargetCells) {
      const uint32_t e = target_entry(threadIdx.x);
      pairs = (T.soff[e + 1] - T.soff[e] + 1u) >> 1;
    }
    uint32_t n_work;
    const uint32_t before = block_exclusive_scan(pairs, T.warp_sum, &n_work);
    if (threadIdx.x <= (unsigned)kTargetCells) T.tcum[threadIdx.x] = threadIdx.x == (unsigned)kTargetCells ? n_work : before;
    const bool fits = T.total + 4u <= slot_cap;
    if (!fits) {
      // the region does not fit the staging area: every particle of the block goes to the per-particle kernel
      if (threadIdx.x < (unsigned)kTargetCells) {
        const uint32_t e = target_entry(threadIdx.x);
        const uint32_t cnt = T.soff[e + 1] - T.soff[e], first = T.gstart[e];
        for (uint32_t k = 0; k < cnt; ++k) slow[atomicAdd(&ctl->n_slow, 1u)] = first + k;
      }
      __syncthreads();
      continue;
    }
    {
      float4* const dst[1] = {cand};
      const float4* const src[1] = {pos};
      tile_stage<1>(T, dst, src, my_count, my_entry, parity);
      parity ^= 1u;
    }
    for (uint32_t base = warp * 32u; base < n_work; base += kTileThreads) {
      const uint32_t item = base + lane;
      const bool live = item < n_work;
      uint32_t e = 0, k0 = 0, cnt = 0;
      if (live) {
        const uint32_t t = find_target(T.tcum, item);
        e = target_entry(t);
        cnt = T.soff[e + 1] - T.soff[e];
        k0 = (item - T.tcum[t]) * 2u;
      }
      const bool has0 = live, has1 = live && k0 + 1u < cnt;
      const uint32_t slot0 = T.soff[e] + k0;
      const float4 p0 = live ? cand[slot0] : make_float4(kSentinel, kSentinel, kSentinel, 0.f);
      const float4 p1 = has1 ? cand[slot0 + 1u] : p0;
      const uint32_t i0 = T.gstart[e] + k0;
      // multi-GPU: the particles of the slab and the ghosts within h of it get a density (k_density_sub)
      const bool need0 = has0 && p0.x >= g.plane_lo - c.h_margin && p0.x < g.plane_hi + c.h_margin;
      const bool need1 = has1 && p1.x >= g.plane_lo - c.h_margin && p1.x < g.plane_hi + c.h_margin;
      // rows of the sub-cell: entries (lz + dz, ly + dy, lx - 1 .. lx + 1)
      uint32_t row_at[9], row_len[9];
      bool long_row = false;
#pragma unroll
      for (int r = 0; r < 9; ++r) {
        const int dz = r / 3 - 1, dy = r % 3 - 1;
        const uint32_t er = (uint32_t)((int)e + (dz * kRegionSide + dy) * kRowEntries - 1);
        row_at[r] = T.soff[live ? er : 0u];
        row_len[r] = T.soff[live ? er + 3u : 0u] - row_at[r];
        long_row |= row_len[r] > 64u;
      }
      bool slow0 = need0 && (long_row || wide_target(p0, g));
      bool slow1 = need1 && (long_row || wide_target(p1, g));
      float acc0 = 0.f, acc1 = 0.f;
      uint32_t cnt0 = 0, cnt1 = 0;
      if ((need0 && !slow0) || (need1 && !slow1)) {
#pragma unroll
        for (int r = 0; r < 9; ++r) {
          uint32_t lo0 = 0, lo1 = 0, hi0 = 0, hi1 = 0;
          const float4* p = cand + row_at[r];
          const uint32_t len = min(row_len[r], 64u);
          if (len > 0u) walk32(p, len, p0, p1, c.support_s, c.h2, acc0, acc1, lo0, lo1);
          if (len > 32u) walk32(p + 32, len - 32u, p0, p1, c.support_s, c.h2, acc0, acc1, hi0, hi1);
          cnt0 += (uint32_t)(__popc(lo0) + __popc(hi0));
          cnt1 += (uint32_t)(__popc(lo1) + __popc(hi1));
          if (need0 && !slow0) nmask[(size_t)r * mask_stride + i0] = (unsigned long long)lo0 | ((unsigned long long)hi0 << 32);
          if (need1 && !slow1) nmask[(size_t)r * mask_stride + i0 + 1u] = (unsigned long long)lo1 | ((unsigned long long)hi1 << 32);
        }
      }
      // more neighbours than the force pass lists: that pass is the per-particle kernel's too
      slow0 |= need0 && cnt0 > list_cap;
      slow1 |= need1 && cnt1 > list_cap;
      if (need0 && !slow0) {
        finish_density(c, acc0, i0, aux, pos, vel);
        ncount[i0] = cnt0;
      }
      if (need1 && !slow1) {
        finish_density(c, acc1, i0 + 1u, aux, pos, vel);
        ncount[i0 + 1u] = cnt1;
      }
      const uint32_t at0 = warp_append(slow0, &ctl->n_slow);
      if (slow0) slow[at0] = i0;
      const uint32_t at1 = warp_append(slow1, &ctl->n_slow);
      if (slow1) slow[at1] = i0 + 1u;
    }
    __syncthreads();  // the staging area is reused by the next block
  }
}

// =============================================================================================
// Forces from the masks: one thread per particle, neighbours gathered from shared memory.
// =============================================================================================
template <bool kFast>
__global__ void __launch_bounds__(kTileThreads)
k_forces_tiles(const float4* __restrict__ pos, const float4* __restrict__ vel, const float4* __restrict__ aux,
               const uint32_t* __restrict__ sub_lb, const uint32_t* __restrict__ keys_a, const uint32_t* __restrict__ keys_b,
               const GridState* __restrict__ grid, const SphConst c, const unsigned long long* __restrict__ nmask,
               size_t mask_stride, const uint32_t* __restrict__ ncount, uint32_t list_cap, const uint32_t* __restrict__ blist,
               TileCtl* ctl, float4* __restrict__ accel, uint32_t slot_cap) {
  extern __shared__ float4 tile_smem[];
  TileTables& T = *reinterpret_cast<TileTables*>(tile_smem);
  float4* cpos = tile_smem + (sizeof(TileTables) + 15) / 16;
  float4* cvel = cpos + slot_cap;
  unsigned short* lists = reinterpret_cast<unsigned short*>(cvel + slot_cap);  // [list_cap][kTileThreads]
  const GridState g = *grid;
  const SubView v = make_view(g, sub_lb, keys_a, keys_b);
  const unsigned warp = threadIdx.x >> 5, lane = lane_id();
  tile_barrier_init(&T.mbar);
  uint32_t parity = 0;
  uint32_t my_count[2], my_entry[2];
  while (tile_begin(T, v, blist, &ctl->n_blocks, &ctl->next_forces, my_count, my_entry)) {
    uint32_t mine = 0;
    if (threadIdx.x < (unsigned)kTargetCells) {
      const uint32_t e = target_entry(threadIdx.x);
      mine = T.soff[e + 1] - T.soff[e];
    }
    uint32_t n_work;
    const uint32_t before = block_exclusive_scan(mine, T.warp_sum, &n_work);
    if (threadIdx.x <= (unsigned)kTargetCells) T.tcum[threadIdx.x] = threadIdx.x == (unsigned)kTargetCells ? n_work : before;
    if (T.total + 4u > slot_cap) {  // the density pass sent this block's particles to the per-particle kernels
      __syncthreads();
      continue;
    }
    {
      float4* const dst[2] = {cpos, cvel};
      const float4* const src[2] = {pos, vel};
      tile_stage<2>(T, dst, src, my_count, my_entry, parity);
      parity ^= 1u;
    }
    unsigned short* my_list = lists + threadIdx.x;
    for (uint32_t base = warp * 32u; base < n_work; base += kTileThreads) {
      const uint32_t item = base + lane;
      const bool live = item < n_work;
      uint32_t e = 0, k = 0;
      if (live) {
        const uint32_t t = find_target(T.tcum, item);
        e = target_entry(t);
        k = item - T.tcum[t];
      }
      const uint32_t self = T.soff[e] + k;
      const uint32_t i = T.gstart[e] + k;
      const float4 pi = live ? cpos[self] : make_float4(0.f, 0.f, 0.f, 0.f);
      const float4 vi = live ? cvel[self] : make_float4(0.f, 0.f, 0.f, 0.f);
      uint32_t count = live ? ncount[i] : 0u;
      // multi-GPU: ghosts get no force; particles without masks are the per-particle kernel's
      const bool mineok = live && owned_here(pi.x, 0u, g) && !(count & kNoMasks);
      if (!mineok) count = 0u;
      // masks -> list of slots, rows in order, bits ascending: the order of the density pass
      uint32_t filled = 0;
#pragma unroll
      for (int r = 0; r < 9; ++r) {
        const int dz = r / 3 - 1, dy = r % 3 - 1;
        const uint32_t er = (uint32_t)((int)e + (dz * kRegionSide + dy) * kRowEntries - 1);
        const uint32_t row_at = T.soff[live ? er : 0u];
        unsigned long long m = mineok ? nmask[(size_t)r * mask_stride + i] : 0ull;
        uint32_t lo = (uint32_t)m, hi = (uint32_t)(m >> 32);
        while (lo) {
          const uint32_t b = (uint32_t)__ffs((int)lo) - 1u;
          lo &= lo - 1u;
          my_list[(size_t)filled * kTileThreads] = (unsigned short)(row_at + b);
          ++filled;
        }
        while (hi) {
          const uint32_t b = (uint32_t)__ffs((int)hi) - 1u;
          hi &= hi - 1u;
          my_list[(size_t)filled * kTileThreads] = (unsigned short)(row_at + 32u + b);
          ++filled;
        }
      }
      // (filled == count by construction; count <= list_cap because the density pass checked it)
      ForceSums sums;
      uint32_t q = 0;
      for (; q + 2u <= filled; q += 2u) {  // two neighbours per trip: four independent loads in flight
        const uint32_t ja = my_list[(size_t)q * kTileThreads], jb = my_list[(size_t)(q + 1u) * kTileThreads];
        const float4 pa = cpos[ja], va = cvel[ja], pb = cpos[jb], vb = cvel[jb];
        add_pair_sel<kFast>(sums, c, ja == self, pi, vi, pi.w, pa, va);
        add_pair_sel<kFast>(sums, c, jb == self, pi, vi, pi.w, pb, vb);
      }
      if (q < filled) {
        const uint32_t j = my_list[(size_t)q * kTileThreads];
        add_pair_sel<kFast>(sums, c, j == self, pi, vi, pi.w, cpos[j], cvel[j]);
      }
      if (mineok) accel[i] = finish_force(sums, c, aux[i].x);
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------------
// Launchers
// ---------------------------------------------------------------------------------------------
namespace {
size_t density_smem(uint32_t slot_cap) { return ((sizeof(TileTables) + 15) / 16 + slot_cap) * sizeof(float4); }
size_t forces_smem(uint32_t slot_cap, uint32_t list_cap) {
  return ((sizeof(TileTables) + 15) / 16 + 2 * (size_t)slot_cap) * sizeof(float4) + (size_t)list_cap * kTileThreads * sizeof(unsigned short);
}
int g_smem_limit = 0;   // opt-in maximum of dynamic shared memory per CTA
int g_smem_per_sm = 0;
}  // namespace

void tiles_init() {
  int dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&g_smem_limit, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
  cudaDeviceGetAttribute(&g_smem_per_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, dev);
  if (g_smem_limit <= 0) g_smem_limit = 227 * 1024;
  if (g_smem_per_sm <= 0) g_smem_per_sm = 228 * 1024;
  cudaFuncSetAttribute(k_density_tiles, cudaFuncAttributeMaxDynamicSharedMemorySize, g_smem_limit);
  cudaFuncSetAttribute(k_forces_tiles<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, g_smem_limit);
  cudaFuncSetAttribute(k_forces_tiles<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, g_smem_limit);
}

// Staging capacities (slots) for a fluid with `per_sub_cell` particles per sub-cell at rest density: room for
// 1.8x the rest population of the region, reduced (never below 1.3x) until `ctas` CTAs fit on an SM.
TilePlan tiles_plan(double per_sub_cell, uint32_t list_cap) {
  if (g_smem_limit == 0) tiles_init();
  const double expect = per_sub_cell * kRegionSide * kRegionSide * kRegionSide + kRows * kPadSlots + 4;
  auto clamp_slots = [&](double want, double floor_slots, size_t (*bytes)(uint32_t, uint32_t), uint32_t lc, int ctas) {
    uint32_t slots = (uint32_t)std::max(want, 1024.0);
    const size_t budget = (size_t)g_smem_per_sm / (size_t)ctas - 1024;  // 1 KB per CTA is reserved by the system
    while (slots > (uint32_t)floor_slots && bytes(slots, lc) > budget) slots -= 64;
    while (bytes(slots, lc) > (size_t)g_smem_limit && slots > 256) slots -= 64;
    return slots & ~3u;
  };
  TilePlan p;
  p.density_slots = clamp_slots(1.8 * expect, 1.3 * expect, [](uint32_t s, uint32_t) { return density_smem(s); }, 0, 5);
  p.forces_slots = clamp_slots(1.8 * expect, 1.3 * expect, [](uint32_t s, uint32_t lc) { return forces_smem(s, lc); }, list_cap, 3);
  // one capacity for both passes: a block the density pass could not stage must be skipped by the force pass too
  p.density_slots = p.forces_slots = std::min(p.density_slots, p.forces_slots);
  p.density_smem = density_smem(p.density_slots);
  p.forces_smem = forces_smem(p.forces_slots, list_cap);
  return p;
}

void launch_density_tiles(float4* pos, float4* vel, const uint32_t* sub_lb, const SortBuffers& sort, const GridState* grid,
                          const SphConst& c, float4* aux, const TileLists& tl, const TilePlan& plan, int sm_count,
                          cudaStream_t stream, uint64_t* launches) {
  const int per_sm = std::max(1, std::min(16, (int)((size_t)g_smem_per_sm / (plan.density_smem + 1024))));
  k_density_tiles<<<sm_count * per_sm, kTileThreads, plan.density_smem, stream>>>(
      pos, vel, sub_lb, sort.keys_a, sort.keys_b, grid, c, aux, tl.masks, tl.mask_stride, tl.count, tl.list_cap, tl.blocks, tl.ctl,
      tl.slow, plan.density_slots);
  if (launches) ++*launches;
}

void launch_forces_tiles(const float4* pos, const float4* vel, const float4* aux, const uint32_t* sub_lb, const SortBuffers& sort,
                         const GridState* grid, const SphConst& c, const TileLists& tl, const TilePlan& plan, bool fast_pairs,
                         float4* accel, int sm_count, cudaStream_t stream, uint64_t* launches) {
  const int per_sm = std::max(1, std::min(16, (int)((size_t)g_smem_per_sm / (plan.forces_smem + 1024))));
  if (fast_pairs)
    k_forces_tiles<true><<<sm_count * per_sm, kTileThreads, plan.forces_smem, stream>>>(
        pos, vel, aux, sub_lb, sort.keys_a, sort.keys_b, grid, c, tl.masks, tl.mask_stride, tl.count, tl.list_cap, tl.blocks, tl.ctl,
        accel, plan.forces_slots);
  else
    k_forces_tiles<false><<<sm_count * per_sm, kTileThreads, plan.forces_smem, stream>>>(
        pos, vel, aux, sub_lb, sort.keys_a, sort.keys_b, grid, c, tl.masks, tl.mask_stride, tl.count, tl.list_cap, tl.blocks, tl.ctl,
        accel, plan.forces_slots);
  if (launches) ++*launches;
}

}  // namespace clsph
