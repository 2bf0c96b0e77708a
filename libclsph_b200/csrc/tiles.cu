// tiles.cu -- the density pass of the sub-cell order as a TILE kernel.
//
// Replaces, like neighbors.cu / subgrid.cu, kernels/sph.cl:9-40 with forces.cl:15-43 of the reference; same
// candidate sets (a subset of the reference's 27 cells that provably holds every particle inside the support),
// same support test s < support_s, same density sum. It also writes the neighbour lists the force pass
// (k_forces_lists, neighbors.cu) walks.
//
// ncu of the per-particle kernel (profiles/r02_a_summary.md) showed where its time went: ~27 SASS instructions per
// candidate slot, one 16-byte global load per candidate per lane at lane-divergent addresses, and six of seven
// candidates fail the support test after paying for all of it. Here:
//
//   * the work unit is a BLOCK of 2 x 2 x 2 grid cells = 4 x 4 x 4 sub-cells of side h. In the Morton order
//     of the sorted arrays a block is 8 consecutive cells, so its particles are one contiguous range;
//   * a CTA stages the positions of the 6 x 6 x 6 sub-cells around the block (the block plus one sub-cell of
//     halo) in shared memory with bulk asynchronous copies (cp.async.bulk, one per sub-cell, completion on an
//     mbarrier; UBLKCP in SASS), laid out ROW-MAJOR (z, y, x): the three sub-cells a particle visits along x in
//     one (z, y) row are one contiguous range of slots -- 9 ranges per particle instead of 18;
//   * next to the fp32 positions it keeps a HALF-PRECISION copy in region coordinates (units of h, two candidates
//     per 16-byte record) and the global index of every slot;
//   * a thread takes one particle. Phase 1 walks each row two candidates per packed-half instruction (HADD2, HMUL2,
//     HFMA2, HSETP2: ~6 instructions per candidate instead of ~13) and keeps the candidates with
//     |d|^2 < 1.03 h^2 in a bit mask -- a superset of the support by a margin ~10x the half-precision error
//     (see pre_pair). Phase 2 runs the exact fp32 test, the density term and the list store over the set bits
//     only: about one candidate in six.
//
// Particles the tiles cannot serve exactly are handed to the warp-per-particle kernel k_density_slow (subgrid.cu)
// through a list (TileCtl::n_slow): particles within 1.5 sub_delta of a sub-cell boundary (their neighbours may
// sit two sub-cells away after rounding, see wide_target), rows of more than 63 candidates, and whole blocks
// whose region exceeds the staging capacity.
//
// Summation order: rows z-major, y, then x, candidates in array order, then the particle's own term -- the order
// of k_density_slow's walk, so densities and lists are bit-identical whichever kernel serves a particle, on one
// GPU and across a slab decomposition.
#include "kernels.cuh"
#include "pair_terms.cuh"
#include "subview.cuh"

namespace clsph {

namespace {

constexpr int kTileThreads = 160;                    // five warps: a block holds ~9.2 warps of particles at rest density, two rounds
constexpr int kTileCtas = 5;                         // resident CTAs per SM the kernel is compiled and planned for
constexpr int kRegionSide = 6;                       // sub-cells per axis of the staged region
constexpr int kRowEntries = kRegionSide + 1;         // per (z, y) row: 6 sub-cells + 1 pad entry of sentinels
constexpr int kRows = kRegionSide * kRegionSide;     // 36
constexpr int kEntries = kRows * kRowEntries;        // 252
constexpr int kPadSlots = 3;                         // sentinel slots after each row
constexpr int kTailSlots = 4;                        // sentinel slots after the last row (phase 1 reads up to 4 slots past a row)
constexpr int kTargetCells = 64;                     // sub-cells of the block proper
constexpr float kSentinel = 1e18f;                   // |d|^2 ~ 3e36: finite, never inside a support; +inf in half precision

// Header of the dynamic shared memory: tables, then the staged arrays.
struct TileTables {
  unsigned long long mbar;        // mbarrier of the bulk copies
  uint32_t block;                 // block id being processed (Morton code of the block coordinates), ~0u = none left
  uint32_t total;                 // staged slots, pads included
  float ox, oy, oz;               // world position of the region's corner (sub-cell coordinates 4 b - 1)
  uint32_t pad;
  uint32_t warp_sum[8];
  uint32_t soff[kEntries + 4];    // first slot of every region entry; [kEntries] = total
  uint32_t gstart[kEntries + 4];  // index of the entry's first particle in the sorted arrays
  uint32_t tcum[kTargetCells + 4];// exclusive running sum of particles over the 64 target sub-cells (memory order)
};
constexpr int kTableFloat4 = (int)((sizeof(TileTables) + 15) / 16);

__device__ __forceinline__ uint32_t warp_inclusive_scan_u32(uint32_t v) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t t = __shfl_up_sync(kFullMask, v, o);
    if (lane_id() >= (unsigned)o) v += t;
  }
  return v;
}

// Exclusive scan of one value per thread over the kThreads threads of the CTA; *total gets the sum.
template <int kThreads>
__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t* warp_sum, uint32_t* total) {
  const unsigned warp = threadIdx.x >> 5;
  const uint32_t inc = warp_inclusive_scan_u32(v);
  if (lane_id() == 31u) warp_sum[warp] = inc;
  __syncthreads();
  uint32_t before = 0, all = 0;
#pragma unroll
  for (unsigned w = 0; w < kThreads / 32; ++w) {
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
template <int kThreads>
__device__ __forceinline__ void tile_barrier_init(unsigned long long* bar) {
#ifndef CLSPH_EMU
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"((uint32_t)kThreads) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
#else
  (void)bar;
#endif
  __syncthreads();
}
// Every thread announces the bytes its own copies will deliver and arrives; the phase completes when all
// threads have arrived and all announced bytes have landed.
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
// Entry where row r (0..8: dz = r / 3 - 1, dy = r % 3 - 1) of a target at entry e starts: sub-cell lx - 1 of that row.
__device__ __forceinline__ uint32_t row_entry(uint32_t e, int r) {
  const int dz = r / 3 - 1, dy = r - 3 * (r / 3) - 1;
  return (uint32_t)((int)e + (dz * kRegionSide + dy) * kRowEntries - 1);
}

// Fetches the next block (persistent CTAs pull from a counter) and fills the tables of its region:
// soff / gstart for the 252 entries, T.total, the region's corner. Returns false when no block is left. Ends with a barrier.
template <int kThreads>
__device__ __forceinline__ bool tile_begin(TileTables& T, const SubView& v, const GridState& g, float sub,
                                           const uint32_t* __restrict__ blist, const uint32_t* n_blocks, uint32_t* next_block,
                                           uint32_t (&my_count)[2], uint32_t (&my_entry)[2]) {
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
  const uint32_t before = block_exclusive_scan<kThreads>(sum, T.warp_sum, &total);
  if (my_entry[0] < (uint32_t)kEntries) T.soff[my_entry[0]] = before;
  if (my_entry[1] < (uint32_t)kEntries) T.soff[my_entry[1]] = before + my_count[0];
  if (threadIdx.x == 0) {
    T.soff[kEntries] = total;
    T.total = total;
    T.ox = fmaf((float)(4 * bx - 1), sub, g.min_x);
    T.oy = fmaf((float)(4 * by - 1), sub, g.min_y);
    T.oz = fmaf((float)(4 * bz - 1), sub, g.min_z);
  }
  __syncthreads();
  return true;
}

// Work items of the block per target sub-cell (pairs of particles for the density pass, particles for the force
// pass), as an exclusive running sum in T.tcum. Returns their total. Ends with a barrier.
template <int kThreads, bool kPairs>
__device__ __forceinline__ uint32_t tile_work(TileTables& T) {
  uint32_t items = 0;
  if (threadIdx.x < (unsigned)kTargetCells) {
    const uint32_t e = target_entry(threadIdx.x);
    const uint32_t cnt = T.soff[e + 1] - T.soff[e];
    items = kPairs ? (cnt + 1u) >> 1 : cnt;
  }
  uint32_t n_work;
  const uint32_t before = block_exclusive_scan<kThreads>(items, T.warp_sum, &n_work);
  if (threadIdx.x <= (unsigned)kTargetCells) T.tcum[threadIdx.x] = threadIdx.x == (unsigned)kTargetCells ? n_work : before;
  __syncthreads();
  return n_work;
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
// [sub_delta, 1 - sub_delta). The test here avoids the IEEE divide: f' = fl(fl(p - min) * fl(2 / cell)) differs
// from f by less than 3 * 2^-24 f <= F 2^-22 = (sub_delta - 2^-20) / 2, so a fraction of f' in
// [1.5 sub_delta, 1 - 1.5 sub_delta) puts the fraction of f in [sub_delta, 1 - sub_delta) (and the two floors
// agree). Wide particles (a few in ten thousand) go to k_density_slow, whose search window is the proven one of
// sub_bounds.
__device__ __forceinline__ bool wide_axis(float p, float mn, float inv_sub, float delta) {
  const float f = __fmul_rn(__fsub_rn(p, mn), inv_sub);
  const float frac = f - floorf(f);
  return !(frac >= delta && frac < 1.f - delta);
}

// ---- packed half precision (two candidates per instruction) -----------------------------------------------
// A record holds two consecutive slots a = 2k, b = 2k + 1 as (x_a, x_b), (y_a, y_b), (z_a, z_b) in region
// coordinates u = (p - corner) / h, 0 <= u <~ 6. Error budget of the phase-1 test, per axis, in units of h:
// conversion to half <= 2^-9 each for target and candidate (u < 8), the subtraction adds <= 2^-11 |d|, so
// |d^ - d| <= 0.0044 and, for a pair inside the support (|d| < 1 + 2^-21), |d^| <= 1.0077, |d^|^2 <= 1.0154; the
// three half-precision roundings of the sum add a factor <= 1.0015: s^ <= 1.0170 < 1.03125 = kPreThreshold. The
// fp32 rounding of u itself (<= 2048 * 2^-22 h even at the far end of the largest grid) is inside the same slack.
// Sentinels and blown-up particles become +-inf or NaN and fail the test, as they fail the exact one.
constexpr float kPreThreshold = 1.03125f;
#ifndef CLSPH_EMU
__device__ __forceinline__ uint32_t half_bits(float v) {
  unsigned short h;
  asm("cvt.rn.f16.f32 %0, %1;" : "=h"(h) : "f"(v));
  return (uint32_t)h;
}
#else
__device__ __forceinline__ uint32_t half_bits(float v) {
  const _Float16 h = (_Float16)v;
  unsigned short b;
  std::memcpy(&b, &h, 2);
  return (uint32_t)b;
}
__device__ __forceinline__ float half_value(uint32_t bits) {
  const unsigned short b = (unsigned short)bits;
  _Float16 h;
  std::memcpy(&h, &b, 2);
  return (float)h;
}
__device__ __forceinline__ bool pre_test_emu(float xi, float yi, float zi, float xj, float yj, float zj) {
  const _Float16 dx = (_Float16)((_Float16)xi - (_Float16)xj), dy = (_Float16)((_Float16)yi - (_Float16)yj),
                 dz = (_Float16)((_Float16)zi - (_Float16)zj);
  _Float16 s = (_Float16)(dx * dx);
  s = (_Float16)((float)dy * (float)dy + (float)s);  // products of halves are exact in fp32: one rounding, like HFMA2
  s = (_Float16)((float)dz * (float)dz + (float)s);
  return (float)s < kPreThreshold;
}
#endif
__device__ __forceinline__ uint32_t half_pair(float v) {
  const uint32_t b = half_bits(v);
  return b | (b << 16);
}
// Both candidates of a record against the target (xi, yi, zi: the target's coordinate in both halves): the hit of
// slot a sets `bit` in we, the hit of slot b sets it in wo.
__device__ __forceinline__ void pre_pair(const uint4& rec, uint32_t xi, uint32_t yi, uint32_t zi, uint32_t thr, uint32_t bit,
                                         uint32_t& we, uint32_t& wo) {
#ifndef CLSPH_EMU
  asm("{\n\t.reg .b32 dx, dy, dz, s;\n\t.reg .pred p, q;\n\t"
      "sub.f16x2 dx, %2, %5;\n\t"
      "sub.f16x2 dy, %3, %6;\n\t"
      "sub.f16x2 dz, %4, %7;\n\t"
      "mul.f16x2 s, dx, dx;\n\t"
      "fma.rn.f16x2 s, dy, dy, s;\n\t"
      "fma.rn.f16x2 s, dz, dz, s;\n\t"
      "setp.lt.f16x2 p|q, s, %8;\n\t"
      "@p or.b32 %0, %0, %9;\n\t"
      "@q or.b32 %1, %1, %9;\n\t}"
      : "+r"(we), "+r"(wo)
      : "r"(xi), "r"(yi), "r"(zi), "r"(rec.x), "r"(rec.y), "r"(rec.z), "r"(thr), "r"(bit));
#else
  (void)thr;
  const float x = half_value(xi), y = half_value(yi), z = half_value(zi);
  if (pre_test_emu(x, y, z, half_value(rec.x), half_value(rec.y), half_value(rec.z))) we |= bit;
  if (pre_test_emu(x, y, z, half_value(rec.x >> 16), half_value(rec.y >> 16), half_value(rec.z >> 16))) wo |= bit;
#endif
}

// Phase 1 over up to 32 slots starting at record `rec` (slot pairs; `slots` > 0 counts from the record's first
// slot): bit b of the result = slot b may be inside the support. Reads whole records, four slots per trip: up to
// three slots past the end, masked off by the caller.
__device__ __forceinline__ uint32_t pre_walk(const uint4* rec, uint32_t slots, uint32_t xi, uint32_t yi, uint32_t zi, uint32_t thr) {
  uint32_t we = 0, wo = 0, bit = 1u;
#pragma unroll 1
  for (uint32_t k = 0; k < slots; k += 4u, rec += 2, bit <<= 4) {
    const uint4 a = rec[0], b = rec[1];
    pre_pair(a, xi, yi, zi, thr, bit, we, wo);
    pre_pair(b, xi, yi, zi, thr, bit << 2, we, wo);
  }
  return we | (wo << 1);
}

// Phase 2 over the set bits of `w` (bit b = slot first + b), ascending: exact support test, density term, list entry.
__device__ __forceinline__ void exact_walk(uint32_t w, uint32_t first, const float4* __restrict__ cand, const uint32_t* __restrict__ gidx,
                                           const float4& pi, float support_s, float h2, float& acc, uint32_t& cnt, uint32_t* row,
                                           uint32_t list_rows) {
  while (w) {
    const uint32_t slot = first + (uint32_t)__ffs((int)w) - 1u;
    w &= w - 1u;
    const float4 q = cand[slot];
    const uint32_t j = gidx[slot];
    const float s = dist2_contract(pi.x, pi.y, pi.z, q.x, q.y, q.z);
    const bool inside = s < support_s;
    const float t = inside ? h2 - s : 0.f;
    acc = fmaf(t * t, t, acc);
    store_if(inside && cnt < list_rows, row + cnt, j);
    cnt += inside ? 1u : 0u;
  }
}

}  // namespace

// =============================================================================================
// Density + Tait pressure + neighbour lists.
// nlist[i * list_rows + e] = e-th neighbour of particle i (index into the sorted arrays; the particle itself is
// not listed), ncount[i] = neighbours found (more than list_rows: list incomplete, k_forces_sub redoes the
// particle).
// =============================================================================================
__global__ void __launch_bounds__(kTileThreads, kTileCtas)
k_density_tiles(float4* pos, float4* vel, const uint32_t* __restrict__ sub_lb, const uint32_t* __restrict__ keys_a,
                const uint32_t* __restrict__ keys_b, const GridState* __restrict__ grid, const SphConst c,
                float4* __restrict__ aux, uint32_t* __restrict__ nlist, uint32_t* __restrict__ ncount, uint32_t list_rows,
                const uint32_t* __restrict__ blist, TileCtl* ctl, uint32_t* __restrict__ slow, uint32_t slot_cap) {
  extern __shared__ float4 tile_smem[];
  TileTables& T = *reinterpret_cast<TileTables*>(tile_smem);
  float4* cand = tile_smem + kTableFloat4;                                  // [slot_cap] fp32 positions
  uint4* hrec = reinterpret_cast<uint4*>(cand + slot_cap);                  // [slot_cap / 2] half-precision records
  uint32_t* gidx = reinterpret_cast<uint32_t*>(hrec + slot_cap / 2);        // [slot_cap] index in the sorted arrays
  const GridState g = *grid;
  const SubView v = make_view(g, sub_lb, keys_a, keys_b);
  const unsigned warp = threadIdx.x >> 5, lane = lane_id();
  const float sub = g.cell * 0.5f, inv_sub = 2.f / g.cell, inv_h = 1.f / c.h;
  const float wide_delta = 1.5f * g.sub_delta;
  const uint32_t thr = half_pair(kPreThreshold);
  tile_barrier_init<kTileThreads>(&T.mbar);
  uint32_t parity = 0;
  uint32_t my_count[2], my_entry[2];
  while (tile_begin<kTileThreads>(T, v, g, sub, blist, &ctl->n_blocks, &ctl->next_density, my_count, my_entry)) {
    const uint32_t n_work = tile_work<kTileThreads, false>(T);  // the particles of the block
    if (T.total + (uint32_t)kTailSlots > slot_cap) {
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
    // half-precision records and global indices of the thread's own entries (pads: sentinels -> +inf)
    {
      const float ox = T.ox, oy = T.oy, oz = T.oz;
      unsigned short* hs = reinterpret_cast<unsigned short*>(hrec);
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        const uint32_t e = my_entry[k];
        if (e >= (uint32_t)kEntries) continue;
        const uint32_t at = T.soff[e], first = T.gstart[e];
        for (uint32_t q = 0; q < my_count[k]; ++q) {
          const uint32_t slot = at + q;
          const float4 p = cand[slot];
          unsigned short* r = hs + (size_t)(slot >> 1) * 8u + (slot & 1u);
          r[0] = (unsigned short)half_bits((p.x - ox) * inv_h);
          r[2] = (unsigned short)half_bits((p.y - oy) * inv_h);
          r[4] = (unsigned short)half_bits((p.z - oz) * inv_h);
          gidx[slot] = first + q;
        }
      }
      if (threadIdx.x < (unsigned)kTailSlots) {  // phase 1 may read this far past the last row
        const uint32_t slot = T.total + threadIdx.x;
        unsigned short* r = hs + (size_t)(slot >> 1) * 8u + (slot & 1u);
        r[0] = r[2] = r[4] = (unsigned short)half_bits(kSentinel);
      }
    }
    __syncthreads();
    for (uint32_t base = warp * 32u; base < n_work; base += kTileThreads) {
      const uint32_t item = base + lane;
      const bool live = item < n_work;
      uint32_t e = target_entry(0), k = 0;
      if (live) {
        const uint32_t t = find_target(T.tcum, item);
        e = target_entry(t);
        k = item - T.tcum[t];
      }
      const uint32_t self = T.soff[e] + k;
      const uint32_t i = T.gstart[e] + k;
      const float4 pi = live ? cand[self] : make_float4(kSentinel, kSentinel, kSentinel, 0.f);
      // multi-GPU: the particles of the slab and the ghosts within h of it get a density (k_density_sub)
      const bool need = live && pi.x >= g.plane_lo - c.h_margin && pi.x < g.plane_hi + c.h_margin;
      bool long_row = false;
#pragma unroll
      for (int r = 0; r < 9; ++r) {
        const uint32_t er = row_entry(e, r);
        long_row |= T.soff[er + 3u] - T.soff[er] > 63u;
      }
      const bool wide = wide_axis(pi.x, g.min_x, inv_sub, wide_delta) || wide_axis(pi.y, g.min_y, inv_sub, wide_delta) ||
                        wide_axis(pi.z, g.min_z, inv_sub, wide_delta);
      const bool is_slow = need && (long_row || wide);
      const bool go = need && !is_slow;
      float acc = 0.f;
      uint32_t cnt = 0;
      if (go) {
        const uint32_t xi = half_pair((pi.x - T.ox) * inv_h), yi = half_pair((pi.y - T.oy) * inv_h),
                       zi = half_pair((pi.z - T.oz) * inv_h);
        uint32_t* row = nlist + (size_t)i * list_rows;
#ifndef CLSPH_EMU
        asm volatile("" : "+l"(row));  // keep the row address in registers (see k_density_sub)
#endif
#pragma unroll 1
        for (int r = 0; r < 9; ++r) {
          const uint32_t er = row_entry(e, r);
          const uint32_t at = T.soff[er];
          const uint32_t len = T.soff[er + 3u] - at;
          const uint32_t first = at & ~1u, lead = at - first, span = lead + len;  // records start at even slots; span <= 64
          uint32_t w0 = 0, w1 = 0;
          const uint4* rec = hrec + (first >> 1);
          if (len > 0u) w0 = pre_walk(rec, min(span, 32u), xi, yi, zi, thr);
          if (span > 32u) w1 = pre_walk(rec + 16, span - 32u, xi, yi, zi, thr);
          // the slots of the row proper: [lead, span) of the 64 bits; without the particle itself
          unsigned long long w = (unsigned long long)w0 | ((unsigned long long)w1 << 32);
          w &= (span >= 64u ? ~0ull : (1ull << span) - 1ull) & ~((1ull << lead) - 1ull);
          if (r == 4) w &= ~(1ull << (self - first));
          exact_walk((uint32_t)w, first, cand, gidx, pi, c.support_s, c.h2, acc, cnt, row, list_rows);
          exact_walk((uint32_t)(w >> 32), first + 32u, cand, gidx, pi, c.support_s, c.h2, acc, cnt, row, list_rows);
        }
        finish_density(c, tile_self_density(acc, pi, c), i, aux, pos, vel);
        ncount[i] = cnt;
      }
      const uint32_t at_slow = warp_append(is_slow, &ctl->n_slow);
      if (is_slow) slow[at_slow] = i;
    }
    __syncthreads();  // the staging area is reused by the next block
  }
}

// ---------------------------------------------------------------------------------------------
// Launchers
// ---------------------------------------------------------------------------------------------
namespace {
size_t density_smem(uint32_t slot_cap) { return (size_t)kTableFloat4 * 16u + (size_t)slot_cap * (16u + 8u + 4u); }
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
}

// Staging capacity (slots) for a fluid with `per_sub_cell` particles per sub-cell at rest density: as much as
// lets the kernel keep its resident CTAs per SM, between 1.25x and 1.8x the rest population of a region.
TilePlan tiles_plan(double per_sub_cell) {
  if (g_smem_limit == 0) tiles_init();
  const double expect = per_sub_cell * kRegionSide * kRegionSide * kRegionSide + kRows * kPadSlots + kTailSlots;
  uint32_t slots = (uint32_t)std::max(1.8 * expect, 1024.0);
  const uint32_t floor_slots = (uint32_t)std::max(1.25 * expect, 512.0);
  const size_t budget = (size_t)g_smem_per_sm / (size_t)kTileCtas - 1024;  // 1 KB per CTA is reserved by the system
  while (slots > floor_slots && density_smem(slots) > budget) slots -= 32;
  while (density_smem(slots) > (size_t)g_smem_limit && slots > 256) slots -= 32;
  TilePlan p;
  p.density_slots = slots & ~3u;
  p.density_smem = density_smem(p.density_slots);
  return p;
}

void launch_density_tiles(float4* pos, float4* vel, const uint32_t* sub_lb, const SortBuffers& sort, const GridState* grid,
                          const SphConst& c, float4* aux, const NeighbourLists& lists, const TileLists& tl, const TilePlan& plan,
                          int sm_count, cudaStream_t stream, uint64_t* launches) {
  const int per_sm = std::max(1, std::min(kTileCtas, (int)((size_t)g_smem_per_sm / (plan.density_smem + 1024))));
  k_density_tiles<<<sm_count * per_sm, kTileThreads, plan.density_smem, stream>>>(
      pos, vel, sub_lb, sort.keys_a, sort.keys_b, grid, c, aux, lists.entries, lists.count, lists.rows, tl.blocks, tl.ctl, tl.slow,
      plan.density_slots);
  if (launches) ++*launches;
}

}  // namespace clsph
