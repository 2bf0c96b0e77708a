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
// Sub-cell order with the grid inside the dense sub-cell table (grid->sort_passes == 0, decided on the device by
// k_grid_setup; every BASELINE config): no radix pass at all. The table the neighbour search needs anyway -- first
// index of every (cell, octant) -- IS the exclusive scan of the per-sub-cell particle counts, so
//   k_keys_hist    counts: arrival number of each particle in its sub-cell, one warp-aggregated atomic per run of equal keys
//   k_scan_table   exclusive scan of the table in place (chunks in arrival order, look-back over the chunk sums)
//   k_onesweep(0)  scatter: slot = table[sub-cell] + arrival number
// The order inside a sub-cell is then that of the atomics, i.e. arbitrary -- which is all the radix sort's "order of the
// previous arrays" was to k_reorder_sub, which places the particles of a sub-cell by their reference rank (order keys
// across GPUs) either way. Same arrays, bit for bit, at a third of the time (profiles/r02_c_summary.md).
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
// Counting sort (grid->sort_passes == 0, kSub only): instead of the digit histograms, the particle takes its arrival
// number among those of its sub-cell from the table (zero when the kernel starts, see k_grid_setup); the unsorted key
// goes to keys_count and the arrival number to arrival (the "b" buffers; the sorted pairs will be in the "a" buffers).
template <bool kSub>
__global__ void __launch_bounds__(256) k_keys_hist(const float4* __restrict__ pos, uint32_t* __restrict__ keys,
                                                   const GridState* __restrict__ grid, uint32_t* __restrict__ hist,
                                                   uint32_t* __restrict__ sub_lb, uint32_t* __restrict__ digit_base,
                                                   uint32_t* __restrict__ done, const uint32_t* __restrict__ index,
                                                   uint32_t* __restrict__ keys_count, uint32_t* __restrict__ arrival,
                                                   GridState* __restrict__ table_note) {
  __shared__ uint32_t s_hist[kMaxSortPasses * kRadix];
  __shared__ uint32_t s_scratch[8];
  __shared__ bool s_last;
  const float mnx = grid->min_x, mny = grid->min_y, mnz = grid->min_z, cell = grid->cell;
  const uint32_t n = grid->n;
  const int passes = (int)grid->sort_passes;
  const uint32_t stride = gridDim.x * blockDim.x;
  // the table is written in this sub-step (here or by k_reorder_sub): a note for the next k_grid_setup, which zeroes it
  // (table_note is `grid` itself; these four words are read by nothing else, see common.cuh)
  if (kSub && sub_lb && blockIdx.x == 0 && threadIdx.x == 0) {
    table_note->table_words = grid->sub_dense ? grid->cell_count * 9u : 0u;
    table_note->table_full = passes == 0 ? 1u : 0u;
    table_note->table_n = n;
    table_note->table_in_b = (uint32_t)passes & 1u;
  }
  if (kSub && passes == 0) {
    const uint32_t last_key = grid->cell_count * 8u - 1u;
    for (uint32_t base = blockIdx.x * blockDim.x + (threadIdx.x & ~31u); base < n; base += stride) {
      const uint32_t i = base + lane_id();
      const bool valid = i < n;
      const unsigned vmask = __ballot_sync(kFullMask, valid);
      if (!valid) continue;
      const float4 p = pos[index ? index[i] : i];
      const uint32_t fx = sub_coord(p.x, mnx, cell), fy = sub_coord(p.y, mny, cell), fz = sub_coord(p.z, mnz, cell);
      uint32_t key = (morton3(fx >> 1, fy >> 1, fz >> 1) << 3) | (fx & 1u) | ((fy & 1u) << 1) | ((fz & 1u) << 2);
      key = min(key, last_key);  // (the bounding box holds every particle, so this never bites; it keeps the atomic in the table)
      // the arrays are in last sub-step's sub-cell order: neighbouring lanes mostly share the sub-cell, one atomic per run
      const unsigned peers = __match_any_sync(vmask, key);
      const int leader = __ffs(peers) - 1;
      uint32_t first = 0;
      if ((int)lane_id() == leader) first = atomicAdd(&sub_lb[(size_t)(key >> 3) * 9u + (key & 7u)], (uint32_t)__popc(peers));
      first = __shfl_sync(peers, first, leader);
      keys_count[i] = key;
      arrival[i] = first + (uint32_t)__popc(peers & lanemask_lt());
    }
    return;
  }
  for (int i = threadIdx.x; i < kMaxSortPasses * kRadix; i += blockDim.x) s_hist[i] = 0;
  __syncthreads();
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
// Counting sort: exclusive scan of the sub-cell table, in place.
// ---------------------------------------------------------------------------------------------
// table[cell * 9 + o] = particles in octant o of the cell (word 8 of a cell: 0)  ->  number of particles with a smaller
// sub-cell key = first index of the octant, and word 8 = end of the cell: the layout the gather kernels read (subview.cuh).
// (An empty cell comes out as [x, x) where k_reorder_sub's table leaves [0, 0): every reader takes ranges.)
// One CTA per SM; each takes chunks of 16 Ki words in arrival order (so every chunk it may wait for is running or done)
// and keeps its chunk in registers: sum, publish, look back over the preceding chunks' sums, add, write. The look-back is
// what the scan costs (an L2 round trip per step on a one-wave table), so a warp reads 128 predecessors per step.
// state[0] = chunk ticket, state[1 + chunk] = flag | value as in k_onesweep (a count of particles: < 2^30).
constexpr int kScanThreads = 1024;
constexpr uint32_t kScanChunk = kScanThreads * 16u;

__global__ void __launch_bounds__(kScanThreads) k_scan_table(uint32_t* __restrict__ table, const GridState* __restrict__ grid,
                                                             uint32_t* __restrict__ state) {
  __shared__ uint32_t s_warp[32];
  __shared__ uint32_t s_chunk, s_exclusive;
  if (grid->sort_passes != 0u) return;
  const unsigned tid = threadIdx.x, warp = tid >> 5, lane = lane_id();
  const uint32_t words = grid->cell_count * 9u;
  volatile uint32_t* status = state + 1;
  for (;;) {
    __syncthreads();  // the shared words of the previous chunk have been read
    if (tid == 0) s_chunk = atomicAdd(state, 1u);
    __syncthreads();
    const uint32_t chunk = s_chunk;
    const uint64_t chunk_lo = (uint64_t)chunk * kScanChunk;
    if (chunk_lo >= words) return;
    const uint32_t lo = (uint32_t)chunk_lo + tid * 16u;
    uint32_t v[16];
    if (lo + 16u <= words) {
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const uint4 x = *reinterpret_cast<const uint4*>(table + lo + 4 * q);
        v[4 * q] = x.x; v[4 * q + 1] = x.y; v[4 * q + 2] = x.z; v[4 * q + 3] = x.w;
      }
    } else {
#pragma unroll
      for (int k = 0; k < 16; ++k) v[k] = lo + (uint32_t)k < words ? table[lo + k] : 0u;
    }
    uint32_t mine = 0;
#pragma unroll
    for (int k = 0; k < 16; ++k) mine += v[k];
    const uint32_t inc = warp_inclusive_scan(mine);
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    if (warp == 0) {
      const uint32_t w = s_warp[lane];
      const uint32_t winc = warp_inclusive_scan(w);
      s_warp[lane] = winc - w;  // what the warps before this one hold
      const uint32_t total = __shfl_sync(kFullMask, winc, 31);
      if (lane == 0) status[chunk] = (chunk == 0 ? kFlagInclusive : kFlagAggregate) | total;
      uint32_t exclusive = 0;
      if (chunk != 0) {
        bool found = false;
        for (int64_t nearest_chunk = (int64_t)chunk - 1; !found; nearest_chunk -= 128) {
          uint32_t x[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) {  // four loads in flight per lane; before the first chunk: nothing, "inclusive"
            const int64_t t = nearest_chunk - 32 * q - (int64_t)lane;
            x[q] = kFlagInclusive + 0u;
            if (t >= 0) x[q] = status[t];
          }
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            if (found) break;
            const int64_t t = nearest_chunk - 32 * q - (int64_t)lane;
            uint32_t y = x[q];
            while ((y & ~kValueMask) == 0u) y = status[t];  // that chunk has not published yet
            const unsigned inclusive = __ballot_sync(kFullMask, (y & ~kValueMask) == kFlagInclusive);
            const int nearest = __ffs(inclusive) - 1;  // lane 0 holds the nearest chunk of the 32
            uint32_t part = (nearest < 0 || (int)lane <= nearest) ? (y & kValueMask) : 0u;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(kFullMask, part, o);
            exclusive += part;
            found = nearest >= 0;
          }
        }
        if (lane == 0) status[chunk] = kFlagInclusive | (exclusive + total);
      }
      if (lane == 0) s_exclusive = exclusive;
    }
    __syncthreads();
    uint32_t run = s_exclusive + s_warp[warp] + inc - mine;
    if (lo + 16u <= words) {
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        uint4 out;
        out.x = run; run += v[4 * q];
        out.y = run; run += v[4 * q + 1];
        out.z = run; run += v[4 * q + 2];
        out.w = run; run += v[4 * q + 3];
        *reinterpret_cast<uint4*>(table + lo + 4 * q) = out;
      }
    } else {
#pragma unroll
      for (int k = 0; k < 16; ++k) {
        if (lo + (uint32_t)k < words) table[lo + k] = run;
        run += v[k];
      }
    }
  }
}

// The scatter of the counting sort, run by the launch of radix pass 0: sorted slot = first index of the particle's
// sub-cell + its arrival number there. keys_count / arrival as written by k_keys_hist; index: see there.
__device__ __forceinline__ void count_scatter(const uint32_t* __restrict__ keys_count, const uint32_t* __restrict__ arrival,
                                              const uint32_t* __restrict__ index, const uint32_t* __restrict__ table,
                                              uint32_t* __restrict__ keys_sorted, uint32_t* __restrict__ vals_sorted, uint32_t n) {
  // (the launch has one thread per 8 or 16 keys: eight independent key -> table -> slot chains in flight per thread)
  constexpr int kBatch = 8;
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t first = blockIdx.x * blockDim.x + threadIdx.x; first < n; first += stride * kBatch) {
    uint32_t key[kBatch], at[kBatch], val[kBatch];
#pragma unroll
    for (int k = 0; k < kBatch; ++k) {
      const uint32_t i = first + (uint32_t)k * stride;
      key[k] = i < n ? keys_count[i] : 0u;
      at[k] = i < n ? arrival[i] : 0u;
      val[k] = (index && i < n) ? index[i] : i;
    }
#pragma unroll
    for (int k = 0; k < kBatch; ++k) at[k] += table[(size_t)(key[k] >> 3) * 9u + (key[k] & 7u)];
#pragma unroll
    for (int k = 0; k < kBatch; ++k) {
      if (first + (uint32_t)k * stride < n) {
        keys_sorted[at[k]] = key[k];
        vals_sorted[at[k]] = val[k];
      }
    }
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
           uint32_t* __restrict__ tile_counter, volatile uint32_t* status, int pass, const uint32_t* __restrict__ table,
           uint32_t* __restrict__ keys_sorted, uint32_t* __restrict__ vals_sorted) {
  constexpr int kWarps = kSortThreads / 32;
  constexpr int kTile = kSortThreads * kItems;
  __shared__ uint32_t s_keys[kTile];
  __shared__ uint32_t s_vals[kTile];
  __shared__ uint32_t s_warp_hist[kWarps][kRadix];
  __shared__ uint32_t s_digit_start[kRadix];
  __shared__ uint32_t s_digit_global[kRadix];
  __shared__ uint32_t s_scratch[8];
  __shared__ uint32_t s_tile;

  if (pass >= (int)grid->sort_passes) {
    // counting sort: keys_out / vals_out of pass 0 (the "b" buffers) hold the unsorted keys and the arrival numbers
    if (pass == 0 && table) count_scatter(keys_out, vals_out, vals_in, table, keys_sorted, vals_sorted, grid->n);
    return;
  }
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

// Leading words of the scratch that a sort of n keys needs zeroed.
size_t sort_scratch_zero_words(uint32_t n) {
  return (size_t)2 * kMaxSortPasses * kRadix + 64 + (size_t)kMaxSortPasses * sort_tiles_for(n) * kRadix;
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
  // histograms, tile counters and the look-back status words of the tiles in use start at zero (sub-cell order: the
  // launch of k_grid_setup has done that, see sort_scratch_zero_words)
  if (!sub_keys) cudaMemsetAsync(b.scratch, 0, sort_scratch_zero_words(n_launch) * sizeof(uint32_t), stream);
  const unsigned hist_blocks = (unsigned)std::min<uint64_t>(((uint64_t)n_launch + 255) / 256, (uint64_t)sm_count * 8);
  uint32_t* done = l.tile_counter + 8;  // (words 0..3 are the tile counters of the passes; all zeroed above)
  if (sub_keys)
    k_keys_hist<true><<<std::max(1u, hist_blocks), 256, 0, stream>>>(pos, b.keys_a, grid, l.hist, sub_lb, l.digit_base, done, index, b.keys_b,
                                                                    b.vals_b, const_cast<GridState*>(grid));
  else
    k_keys_hist<false><<<std::max(1u, hist_blocks), 256, 0, stream>>>(pos, b.keys_a, grid, l.hist, nullptr, l.digit_base, done, index, nullptr,
                                                                     nullptr, nullptr);
  if (launches) ++*launches;
  if (keys_tap) launch_copy_u32(b.keys_a, keys_tap, n_launch, stream, launches);
}

// Words of the scan's state for a table of sub_capacity cells: the ticket and one word per chunk.
uint32_t scan_state_words(uint32_t sub_capacity) {
  return (uint32_t)(((uint64_t)sub_capacity * 9u + kScanChunk - 1u) / kScanChunk) + 2u;
}

// Counting sort only (returns at once otherwise): the scan of the sub-cell table. scan_state: scan_state_words()
// words, zeroed by k_grid_setup.
void launch_scan_table(uint32_t* sub_lb, const GridState* grid, uint32_t* scan_state, uint32_t sub_capacity, int sm_count,
                       cudaStream_t stream, uint64_t* launches) {
  const uint32_t most_chunks = scan_state_words(sub_capacity) - 2u;
  k_scan_table<<<std::max(1u, std::min<uint32_t>((uint32_t)sm_count, most_chunks)), kScanThreads, 0, stream>>>(sub_lb, grid, scan_state);
  if (launches) ++*launches;
}

// Histogram scan + the four digit passes (those beyond grid->sort_passes return immediately).
// first_vals: payload of the first pass (null: the identity) -- the index list of an exchange in place.
// table: the scanned sub-cell table of a counting sort (null: radix passes only), whose scatter pass 0's launch runs.
void launch_sort_passes(const SortBuffers& b, const GridState* grid, uint32_t n_launch, const uint32_t* first_vals,
                        const uint32_t* table, cudaStream_t stream, uint64_t* launches) {
  const ScratchLayout l = layout_of(b);
  const uint32_t tiles = sort_tiles_for(n_launch);
  // (the histogram scan is done by the last CTA of k_keys_hist)
  // pass 0: a -> b (identity payload), pass 1: b -> a, pass 2: a -> b, pass 3: b -> a
  for (int pass = 0; pass < kMaxSortPasses; ++pass) {
    const bool even = (pass & 1) == 0;
    if (sort_items_for(n_launch) == 8)
      k_onesweep<8><<<tiles, kSortThreads, 0, stream>>>(even ? b.keys_a : b.keys_b, pass == 0 ? first_vals : (even ? b.vals_a : b.vals_b),
                                                        even ? b.keys_b : b.keys_a, even ? b.vals_b : b.vals_a, grid, l.digit_base,
                                                        l.tile_counter, l.status, pass, table, b.keys_a, b.vals_a);
    else
      k_onesweep<16><<<tiles, kSortThreads, 0, stream>>>(even ? b.keys_a : b.keys_b, pass == 0 ? first_vals : (even ? b.vals_a : b.vals_b),
                                                         even ? b.keys_b : b.keys_a, even ? b.vals_b : b.vals_a, grid, l.digit_base,
                                                         l.tile_counter, l.status, pass, table, b.keys_a, b.vals_a);
  }
  if (launches) *launches += kMaxSortPasses;
}

}  // namespace clsph
