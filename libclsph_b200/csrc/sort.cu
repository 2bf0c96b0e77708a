// sort.cu -- cell keys + stable LSD radix sort of (key, index) pairs, onesweep style.
//
// Replaces, in the reference: kernels/grid.cl:43-67 (locate_in_grid), kernels/sort.cl:23-59
// (sort_count / sort, 128 work-items) and the host scan of libclsph/sph_simulation.cpp:128-143.
// Contract: the permutation equals a stable sort of the input order by Morton cell key.
//
// Structure per sub-step:
//   k_keys_hist   one read of positions: key per particle + all digit histograms at once
//   k_scan_hist   exclusive scan of the 256-bin histogram of every pass (one CTA)
//   k_onesweep    per 8-bit digit: one read + one write of the pairs. Each CTA takes a tile in
//                 arrival order, ranks its keys with warp-aggregated match_any multisplit,
//                 publishes its digit counts and resolves its global offsets by decoupled
//                 look-back over the preceding tiles (no separate scan pass, no atomics on keys).
// Only ceil(bits(grid_cell_count-1)/8) passes run; the others return at once.
//
// Algorithmic traffic per particle: keys 16 B read + 4 B write + 4 B re-read for the histogram
// accounting of SURVEY 8(d) (here fused: the key never leaves registers), then 16 B per pass.
#include <cstdlib>

#include "kernels.cuh"

namespace clsph {

namespace {

constexpr uint32_t kFlagAggregate = 1u << 30;
constexpr uint32_t kFlagInclusive = 2u << 30;
constexpr uint32_t kValueMask = (1u << 30) - 1u;

__device__ __forceinline__ uint32_t warp_inclusive_scan(uint32_t v) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    uint32_t t = __shfl_up_sync(kFullMask, v, o);
    if (lane_id() >= (unsigned)o) v += t;
  }
  return v;
}

// Exclusive scan of one value per thread over a 256-thread CTA. `scratch` holds 8 words.
__device__ __forceinline__ uint32_t block256_exclusive_scan(uint32_t v, uint32_t* scratch) {
  const unsigned warp = threadIdx.x >> 5, lane = lane_id();
  uint32_t inc = warp_inclusive_scan(v);
  if (lane == 31) scratch[warp] = inc;
  __syncthreads();
  uint32_t warp_prefix = 0;
#pragma unroll
  for (unsigned w = 0; w < 8; ++w)
    if (w < warp) warp_prefix += scratch[w];
  __syncthreads();
  return warp_prefix + inc - v;
}

}  // namespace

// ---------------------------------------------------------------------------------------------
// Keys + histograms.  grid.cl:56-64: key = morton((uint)((p - min) / (2h)) per axis).
// ---------------------------------------------------------------------------------------------
// kSub: the key gets three more bits, the octant of the cell the particle is in (which half along
// x, y, z): floor(2q) & 1 with q = (p - min) / (2h) as above; 2q is exact, and floor(2q) >> 1 ==
// floor(q), so the cell part is unchanged and a stable sort by the longer key is a stable sort by
// cell refined by octant.
// Two chores ride along, each of which used to be a launch of its own (5 us apiece at 1 Mi particles, where the whole
// sub-step is 520): the sub-cell table of the gather pass is cleared (sub_lb, nothing else touches it here), and the
// last CTA to finish scans the histograms into digit_base (what k_scan_hist does; `done` counts the CTAs).
template <bool kSub>
__global__ void __launch_bounds__(256) k_keys_hist(const float4* __restrict__ pos, uint32_t* __restrict__ keys,
                                                   const GridState* __restrict__ grid, uint32_t* __restrict__ hist,
                                                   uint32_t* __restrict__ sub_lb, uint32_t* __restrict__ digit_base,
                                                   uint32_t* __restrict__ done, const uint32_t* __restrict__ index) {
  __shared__ uint32_t s_hist[kMaxSortPasses * kRadix];
  __shared__ uint32_t s_scratch[8];
  __shared__ bool s_last;
  for (int i = threadIdx.x; i < kMaxSortPasses * kRadix; i += blockDim.x) s_hist[i] = 0;
  __syncthreads();
  if (kSub && sub_lb && grid->sub_dense) {
    const size_t words = (size_t)grid->cell_count * 9u;
    for (size_t w = (size_t)blockIdx.x * blockDim.x + threadIdx.x; w < words; w += (size_t)gridDim.x * blockDim.x) sub_lb[w] = 0u;
  }

  const float mnx = grid->min_x, mny = grid->min_y, mnz = grid->min_z, cell = grid->cell;
  const uint32_t n = grid->n;
  const int passes = (int)grid->sort_passes;
  const uint32_t stride = gridDim.x * blockDim.x;
  // warp-uniform trip count so that the match/ballot intrinsics see whole warps
  for (uint32_t base = blockIdx.x * blockDim.x + (threadIdx.x & ~31u); base < n; base += stride) {
    const uint32_t i = base + lane_id();
    const bool valid = i < n;
    const unsigned vmask = __ballot_sync(kFullMask, valid);
    if (!valid) continue;
    // `index` (multi-GPU exchange in place): entry i of the sort's input is the particle at index[i]
    const float4 p = pos[index ? index[i] : i];
    uint32_t key;
    if (kSub) {
      const uint32_t fx = sub_coord(p.x, mnx, cell), fy = sub_coord(p.y, mny, cell), fz = sub_coord(p.z, mnz, cell);
      key = (morton3(fx >> 1, fy >> 1, fz >> 1) << 3) | (fx & 1u) | ((fy & 1u) << 1) | ((fz & 1u) << 2);
    } else {
      key = morton3(cell_coord(p.x, mnx, cell), cell_coord(p.y, mny, cell), cell_coord(p.z, mnz, cell));
    }
    keys[i] = key;
    for (int pass = 0; pass < passes; ++pass) {
      const uint32_t d = (key >> (8 * pass)) & 0xFFu;
      // neighbouring particles mostly share a cell: aggregate equal digits inside the warp
      const unsigned peers = __match_any_sync(vmask, d);
      if ((peers & lanemask_lt()) == 0) atomicAdd(&s_hist[pass * kRadix + d], (uint32_t)__popc(peers));
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < kMaxSortPasses * kRadix; i += blockDim.x) {
    const uint32_t c = s_hist[i];
    if (c) atomicAdd(&hist[i], c);
  }
  if (!done) return;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = atomicAdd(done, 1u) == gridDim.x - 1u;
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  for (int pass = 0; pass < passes; ++pass) {  // blockDim.x == kRadix
    const uint32_t v = atomicAdd(&hist[pass * kRadix + threadIdx.x], 0u);
    digit_base[pass * kRadix + threadIdx.x] = block256_exclusive_scan(v, s_scratch);
  }
}

// One CTA of 256 threads: digit_base[pass][d] = number of keys whose digit in `pass` is < d.
__global__ void __launch_bounds__(256) k_scan_hist(const uint32_t* __restrict__ hist, uint32_t* __restrict__ digit_base,
                                                   const GridState* __restrict__ grid) {
  __shared__ uint32_t scratch[8];
  const int passes = (int)grid->sort_passes;
  for (int pass = 0; pass < passes; ++pass) {
    const uint32_t v = hist[pass * kRadix + threadIdx.x];
    digit_base[pass * kRadix + threadIdx.x] = block256_exclusive_scan(v, scratch);
  }
}

// ---------------------------------------------------------------------------------------------
// One onesweep pass.
// ---------------------------------------------------------------------------------------------
// kItems keys per thread: tiles of 256 * kItems keys. 16 amortises the look-back and the per-tile histogram
// work; 8 doubles the CTAs, which only pays when 16 would leave most SMs without a tile (~100 k keys).
template <int kItems>
__global__ void __launch_bounds__(kSortThreads)
k_onesweep(const uint32_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in, uint32_t* __restrict__ keys_out,
           uint32_t* __restrict__ vals_out, const GridState* __restrict__ grid, const uint32_t* __restrict__ digit_base,
           uint32_t* __restrict__ tile_counter, volatile uint32_t* status, int pass) {
  constexpr int kWarps = kSortThreads / 32;
  constexpr int kTile = kSortThreads * kItems;
  __shared__ uint32_t s_keys[kTile];
  __shared__ uint32_t s_vals[kTile];
  __shared__ uint32_t s_warp_hist[kWarps][kRadix];
  __shared__ uint32_t s_digit_start[kRadix];
  __shared__ uint32_t s_digit_global[kRadix];
  __shared__ uint32_t s_scratch[8];
  __shared__ uint32_t s_tile;

  if (pass >= (int)grid->sort_passes) return;
  const uint32_t n = grid->n;
  const unsigned tid = threadIdx.x, warp = tid >> 5, lane = lane_id();

  // Tiles are handed out in arrival order, so every tile this CTA may wait on is already running.
  if (tid == 0) s_tile = atomicAdd(tile_counter + pass, 1u);
  for (int i = tid; i < kWarps * kRadix; i += kSortThreads) (&s_warp_hist[0][0])[i] = 0;
  __syncthreads();
  const uint32_t tile = s_tile;
  const uint64_t tile_base64 = (uint64_t)tile * kTile;
  if (tile_base64 >= n) return;
  const uint32_t tile_base = (uint32_t)tile_base64;
  const int shift = 8 * pass;
  digit_base += pass * kRadix;
  status += (size_t)pass * gridDim.x * kRadix;

  // Each warp owns a contiguous run of 32*kItems keys and walks it 32 at a time: tile order
  // = (warp, round, lane), which is what makes the ranks stable.
  uint32_t key[kItems], val[kItems], rank[kItems];
  const uint32_t warp_base = tile_base + warp * (32 * kItems);
#pragma unroll
  for (int k = 0; k < kItems; ++k) {
    const uint32_t idx = warp_base + k * 32 + lane;
    const bool ok = idx < n;
    key[k] = ok ? keys_in[idx] : 0xFFFFFFFFu;  // padding sorts to the very end of the tile
    val[k] = (vals_in != nullptr && ok) ? vals_in[idx] : idx;
  }
  // the matches of the kItems rounds do not depend on each other: issued back to back (their latency is what a
  // pass over a million keys -- one wave of tiles -- mostly consists of), then the short dependent part per round
  unsigned match[kItems];
#pragma unroll
  for (int k = 0; k < kItems; ++k) match[k] = __match_any_sync(kFullMask, (key[k] >> shift) & 0xFFu);
#pragma unroll
  for (int k = 0; k < kItems; ++k) {
    const uint32_t d = (key[k] >> shift) & 0xFFu;
    const unsigned peers = match[k];
    const int leader = __ffs(peers) - 1;
    uint32_t prev = 0;
    if ((int)lane == leader) {
      prev = s_warp_hist[warp][d];
      s_warp_hist[warp][d] = prev + __popc(peers);
    }
    prev = __shfl_sync(kFullMask, prev, leader);
    rank[k] = prev + __popc(peers & lanemask_lt());
    __syncwarp();
  }
  __syncthreads();

  // Thread d owns digit d: tile count, per-warp exclusive offsets, publish, look back.
  {
    const unsigned d = tid;
    uint32_t run = 0;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) {
      const uint32_t c = s_warp_hist[w][d];
      s_warp_hist[w][d] = run;
      run += c;
    }
    const uint32_t count = run;
    status[(size_t)tile * kRadix + d] = (tile == 0 ? kFlagInclusive : kFlagAggregate) | count;

    const uint32_t start = block256_exclusive_scan(count, s_scratch);
    s_digit_start[d] = start;

    uint32_t exclusive = 0;
    if (tile > 0) {
      uint32_t t = tile - 1;
      while (true) {
        const uint32_t v = status[(size_t)t * kRadix + d];
        const uint32_t flag = v & ~kValueMask;
        if (flag == 0) continue;  // predecessor has not published yet
        exclusive += v & kValueMask;
        if (flag == kFlagInclusive) break;
        --t;
      }
      status[(size_t)tile * kRadix + d] = kFlagInclusive | (exclusive + count);
    }
    s_digit_global[d] = digit_base[d] + exclusive - start;
  }
  __syncthreads();

  // Scatter inside shared memory to tile-sorted order, then stream out coalesced runs.
#pragma unroll
  for (int k = 0; k < kItems; ++k) {
    const uint32_t d = (key[k] >> shift) & 0xFFu;
    const uint32_t at = s_digit_start[d] + s_warp_hist[warp][d] + rank[k];
    s_keys[at] = key[k];
    s_vals[at] = val[k];
  }
  __syncthreads();
  const uint32_t tile_n = min((uint32_t)kTile, n - tile_base);
  for (uint32_t i = tid; i < tile_n; i += kSortThreads) {
    const uint32_t k = s_keys[i];
    const uint32_t at = s_digit_global[(k >> shift) & 0xFFu] + i;
    keys_out[at] = k;
    vals_out[at] = s_vals[i];
  }
}

// ---------------------------------------------------------------------------------------------
// Host-side launchers
// ---------------------------------------------------------------------------------------------
// Keys per thread of a onesweep tile for n keys (see k_onesweep). CLSPH_SORT_ITEMS=8|16 overrides (tuning).
static int sort_items_for(uint32_t n) {
  static const int forced = [] { const char* e = getenv("CLSPH_SORT_ITEMS"); return e ? atoi(e) : 0; }();
  if (forced == 8 || forced == 16) return forced;
  return n <= (1u << 18) ? 8 : 16;  // measured on a B200: 8 wins at 100 k keys (34 vs 38 us), 16 from 1 Mi up (56 vs 64 us)
}
uint32_t sort_tiles_for(uint32_t n) {
  const uint32_t tile = (uint32_t)kSortThreads * (uint32_t)sort_items_for(n);
  return (n + tile - 1) / tile;
}

size_t sort_scratch_words(uint32_t max_particles) {
  // hist[4][256] + digit_base[4][256] + tile_counter[4] (padded to 64) + status[4][tiles][256]
  const uint32_t most_tiles = (max_particles + kSortThreads * 8u - 1u) / (kSortThreads * 8u);  // the smaller tile
  return (size_t)2 * kMaxSortPasses * kRadix + 64 + (size_t)kMaxSortPasses * most_tiles * kRadix;
}

namespace {
struct ScratchLayout {
  uint32_t *hist, *digit_base, *tile_counter, *status;
};
ScratchLayout layout_of(const SortBuffers& b) {
  ScratchLayout l;
  l.hist = b.scratch;
  l.digit_base = l.hist + kMaxSortPasses * kRadix;
  l.tile_counter = l.digit_base + kMaxSortPasses * kRadix;
  l.status = l.tile_counter + 64;
  return l;
}
}  // namespace

// Zeroes the scratch, computes keys into keys_a and all digit histograms.
void launch_sort_keys(const SortBuffers& b, const float4* pos, const GridState* grid, uint32_t n_launch, int sm_count,
                      uint32_t* keys_tap, bool sub_keys, uint32_t* sub_lb, const uint32_t* index, cudaStream_t stream,
                      uint64_t* launches) {
  const ScratchLayout l = layout_of(b);
  const uint32_t tiles = sort_tiles_for(n_launch);
  // histograms, tile counters and the look-back status words of the tiles in use start at zero
  const size_t zero_words = (size_t)2 * kMaxSortPasses * kRadix + 64 + (size_t)kMaxSortPasses * tiles * kRadix;
  cudaMemsetAsync(b.scratch, 0, zero_words * sizeof(uint32_t), stream);
  const unsigned hist_blocks = (unsigned)std::min<uint64_t>(((uint64_t)n_launch + 255) / 256, (uint64_t)sm_count * 8);
  uint32_t* done = l.tile_counter + 8;  // (words 0..3 are the tile counters of the passes; all zeroed above)
  if (sub_keys) k_keys_hist<true><<<std::max(1u, hist_blocks), 256, 0, stream>>>(pos, b.keys_a, grid, l.hist, sub_lb, l.digit_base, done, index);
  else k_keys_hist<false><<<std::max(1u, hist_blocks), 256, 0, stream>>>(pos, b.keys_a, grid, l.hist, nullptr, l.digit_base, done, index);
  if (launches) ++*launches;
  if (keys_tap) launch_copy_u32(b.keys_a, keys_tap, n_launch, stream, launches);
}

// Histogram scan + the four digit passes (those beyond grid->sort_passes return immediately).
// first_vals: payload of the first pass (null: the identity) -- the index list of an exchange in place.
void launch_sort_passes(const SortBuffers& b, const GridState* grid, uint32_t n_launch, const uint32_t* first_vals,
                        cudaStream_t stream, uint64_t* launches) {
  const ScratchLayout l = layout_of(b);
  const uint32_t tiles = sort_tiles_for(n_launch);
  // (the histogram scan is done by the last CTA of k_keys_hist)
  // pass 0: a -> b (identity payload), pass 1: b -> a, pass 2: a -> b, pass 3: b -> a
  for (int pass = 0; pass < kMaxSortPasses; ++pass) {
    const bool even = (pass & 1) == 0;
    if (sort_items_for(n_launch) == 8)
      k_onesweep<8><<<tiles, kSortThreads, 0, stream>>>(even ? b.keys_a : b.keys_b, pass == 0 ? first_vals : (even ? b.vals_a : b.vals_b),
                                                        even ? b.keys_b : b.keys_a, even ? b.vals_b : b.vals_a, grid, l.digit_base,
                                                        l.tile_counter, l.status, pass);
    else
      k_onesweep<16><<<tiles, kSortThreads, 0, stream>>>(even ? b.keys_a : b.keys_b, pass == 0 ? first_vals : (even ? b.vals_a : b.vals_b),
                                                         even ? b.keys_b : b.keys_a, even ? b.vals_b : b.vals_a, grid, l.digit_base,
                                                         l.tile_counter, l.status, pass);
  }
  if (launches) *launches += kMaxSortPasses;
}

}  // namespace clsph
