// grid.cu -- per-step grid bookkeeping on the device: AABB, grid sizing, gather into sorted
// order, cell start/end table, AoS <-> SoA conversion at the API boundary.
//
// Replaces the host-resident fragments of the reference step:
//   libclsph/sph_simulation.cpp:201-217  serial min/max over positions   -> k_bounds (+ fused in the integrator)
//   :221-252                             padding, grid_size, cell count  -> k_grid_setup (one thread, same fp32 ops)
//   :158-170                             D2H of the sorted array + serial cell table -> k_reorder
//   :195, :264, :278, :339               whole-array AoS transfers       -> k_aos_to_soa / k_soa_to_aos (API edge only)
#include "kernels.cuh"

namespace clsph {

namespace {

__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(kFullMask, v, o));
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(kFullMask, v, o));
  return v;
}

}  // namespace

// sentinels of sph_simulation.cpp:204-205: min starts at (float)INT_MAX, max at (float)INT_MIN
__global__ void k_bounds_reset(BoundsAcc* acc) {
  if (threadIdx.x < 3) {
    acc->lo[threadIdx.x] = float_to_ordered(2147483648.f);
    acc->hi[threadIdx.x] = float_to_ordered(-2147483648.f);
  }
}

__global__ void __launch_bounds__(256) k_bounds(const float4* __restrict__ pos, uint32_t n, BoundsAcc* acc) {
  float lo[3] = {2147483648.f, 2147483648.f, 2147483648.f};
  float hi[3] = {-2147483648.f, -2147483648.f, -2147483648.f};
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const float4 p = pos[i];
    lo[0] = fminf(lo[0], p.x); hi[0] = fmaxf(hi[0], p.x);
    lo[1] = fminf(lo[1], p.y); hi[1] = fmaxf(hi[1], p.y);
    lo[2] = fminf(lo[2], p.z); hi[2] = fmaxf(hi[2], p.z);
  }
  __shared__ float s_lo[8][3], s_hi[8][3];
  const unsigned warp = threadIdx.x >> 5;
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    lo[a] = warp_min(lo[a]);
    hi[a] = warp_max(hi[a]);
    if (lane_id() == 0) { s_lo[warp][a] = lo[a]; s_hi[warp][a] = hi[a]; }
  }
  __syncthreads();
  if (threadIdx.x < 3) {
    float l = s_lo[0][threadIdx.x], h = s_hi[0][threadIdx.x];
    for (unsigned w = 1; w < blockDim.x / 32; ++w) { l = fminf(l, s_lo[w][threadIdx.x]); h = fmaxf(h, s_hi[w][threadIdx.x]); }
    atomicMin(&acc->lo[threadIdx.x], float_to_ordered(l));
    atomicMax(&acc->hi[threadIdx.x], float_to_ordered(h));
  }
}

// sph_simulation.cpp:221-252, operation for operation in fp32:
//   cell = h*2;  min -= cell*2;  max += cell*2;  grid_size = (uint)((max - min) / cell);
//   grid_cell_count = morton(grid_size).  Also re-arms the AABB accumulator for the next step.
// Multi-GPU: plane_lo / plane_hi are the fixed world-space x planes bounding this rank's slab
// (-inf / +inf at the ends); they are snapped to the nearest cell boundary of the current grid, the
// same on every rank because the AABB was all-reduced. keep_n: the particle count lives on the
// device (it changes with migration) and must not be overwritten.
// sub_mode: sort by sub-cell keys (cell key << 3 | octant), which need 3 more key bits; sub_capacity =
// cells the dense sub-cell table can hold. count_sort (0 / 1 / 2, the option): when the grid fits that table -- and,
// with 1, the table is small beside the particles -- the sub-step sorts by counting on it (sort_passes = 0, see sort.cu)
// instead of by radix passes.
// The other threads of the launch zero what the previous sub-step left in the sub-cell table (grid->table_words, which
// thread 0 does not touch), the state words of the table scan and the sort's histograms and look-back words: every later
// kernel of the step is ordered behind this launch, and nothing before it in the step reads any of them.
__global__ void __launch_bounds__(256) k_grid_setup(BoundsAcc* acc, GridState* grid, float h, uint32_t n, uint32_t cell_capacity,
                                                    float plane_lo, float plane_hi, int keep_n, uint32_t sub_mode, uint32_t sub_capacity,
                                                    uint32_t count_sort, const StepZero zero) {
  if (zero.sub_lb) {
    const size_t words = grid->table_words;
    const size_t stride = (size_t)gridDim.x * blockDim.x, first = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (grid->table_full) {
      for (size_t w = first; w < words; w += stride) zero.sub_lb[w] = 0u;
    } else if (words) {  // the rows of the cells that held particles: where the sorted keys change cell
      const uint32_t* __restrict__ keys = grid->table_in_b ? zero.keys_b : zero.keys_a;
      const uint32_t n_old = grid->table_n;
      for (size_t r = first; r < n_old; r += stride) {
        const uint32_t key = keys[r] >> 3;
        if ((r == 0 || (keys[r - 1] >> 3) != key) && ((size_t)key + 1u) * 9u <= words) {
          uint32_t* row = zero.sub_lb + (size_t)key * 9u;
#pragma unroll
          for (int o = 0; o < 9; ++o) row[o] = 0u;
        }
      }
    }
    for (size_t w = first; w < zero.scan_words; w += stride) zero.scan_state[w] = 0u;
    for (size_t w = first; w < zero.sort_scratch_words; w += stride) zero.sort_scratch[w] = 0u;
    if (first < 2 && zero.pair_count) zero.pair_count[first] = 0u;  // pair items, overflowing lists (subgrid.cu)
  }
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const float cell = __fmul_rn(h, 2.f);
  const float pad = __fmul_rn(cell, 2.f);
  float mn[3], mx[3];
  int gs[3];
  uint32_t err = 0;
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    mn[a] = __fsub_rn(ordered_to_float(acc->lo[a]), pad);
    mx[a] = __fadd_rn(ordered_to_float(acc->hi[a]), pad);
    // saturating conversion, kept unsigned for the test: an infinite extent must not wrap to -1
    const uint32_t cells = __float2uint_rz(__fdiv_rn(__fsub_rn(mx[a], mn[a]), cell));
    gs[a] = (int)min(cells, 0x7fffffffu);
    if (cells >= 1024u) err |= 1u;  // the reference asserts here (:247-249)
    acc->lo[a] = float_to_ordered(2147483648.f);
    acc->hi[a] = float_to_ordered(-2147483648.f);
  }
  const uint32_t count = morton3((uint32_t)gs[0], (uint32_t)gs[1], (uint32_t)gs[2]);
  grid->min_x = mn[0]; grid->min_y = mn[1]; grid->min_z = mn[2]; grid->cell = cell;
  grid->max_x = mx[0]; grid->max_y = mx[1]; grid->max_z = mx[2];
  grid->plane_lo = plane_lo;
  grid->plane_hi = plane_hi;
  grid->gx = gs[0]; grid->gy = gs[1]; grid->gz = gs[2];
  grid->cell_count = count;
  if (!keep_n) grid->n = n;
  grid->prev_lo = grid->own_lo;
  grid->prev_hi = grid->own_hi;
  const float inf = __int_as_float(0x7f800000);
  // nearest cell boundary: floor((plane - min) / cell + 0.5), clamped to the grid
  grid->own_lo = (plane_lo == -inf) ? 0 : max(0, min(gs[0], (int)floorf((plane_lo - mn[0]) / cell + 0.5f)));
  grid->own_hi = (plane_hi == inf) ? 0x7fffffff : max(0, min(gs[0], (int)floorf((plane_hi - mn[0]) / cell + 0.5f)));
  if (sub_mode && count > (1u << 29)) err |= 4u;  // (key << 3 | octant) would not fit 32 bits
  // largest sort key: count - 1, or 8 * count - 1 with the octant bits appended
  const uint32_t top = sub_mode ? (count >= (1u << 29) ? 0xFFFFFFFFu : max(count * 8u, 2u) - 1u) : (count > 1u ? count - 1u : 1u);
  const uint32_t bits = 32u - (uint32_t)__clz((int)top);
  grid->dense = (count <= cell_capacity) ? 1u : 0u;
  grid->sub = sub_mode ? 1u : 0u;
  grid->sub_dense = (sub_mode && count <= sub_capacity) ? 1u : 0u;
  // A counting sort scans (and the next sub-step zeroes) the whole table, so it only pays while the table is small beside
  // the particles. Measured on B200s (profiles/r02_ak_*, r02_am_*): 48 us against the radix passes' 77 at 2 table words
  // per particle (1 Mi particles, one GPU), 119 against 90 at 13 (one slab of eight, whose table spans the whole domain): break-even near 7.
  const uint32_t n_now = keep_n ? grid->n : n;
  const bool counting = count_sort && sub_mode && !err && count <= sub_capacity &&
                        (count_sort == 2u || (uint64_t)count * 9u <= (uint64_t)n_now * 6u);  // (option count_sort = 2: whenever it fits)
  grid->sort_passes = err ? 4u : counting ? 0u : max(1u, (bits + 7u) / 8u);
  grid->error |= err;  // sticky until the host reads and clears it
}

__global__ void __launch_bounds__(256) k_clear_cells(uint32_t* __restrict__ cell_start, uint32_t* __restrict__ cell_end,
                                                     const GridState* __restrict__ grid) {
  if (!grid->dense) return;
  const uint32_t count = grid->cell_count;
  for (uint32_t c = blockIdx.x * blockDim.x + threadIdx.x; c < count; c += gridDim.x * blockDim.x) {
    cell_start[c] = 0u;
    cell_end[c] = 0u;
  }
}

// Sorted slot r takes the particle that sat at perm[r]; neighbouring slots with different keys
// mark where a cell starts / ends (the table of sph_simulation.cpp:163-170 without the serial walk).
__global__ void __launch_bounds__(256)
k_reorder(const float4* __restrict__ src_pos, const float4* __restrict__ src_vel, const float4* __restrict__ src_ivel,
          float4* __restrict__ dst_pos, float4* __restrict__ dst_vel, float4* __restrict__ dst_ivel,
          const uint32_t* __restrict__ keys_a, const uint32_t* __restrict__ keys_b, const uint32_t* __restrict__ vals_a,
          const uint32_t* __restrict__ vals_b, uint32_t* __restrict__ skey, uint32_t* __restrict__ perm_out,
          uint32_t* __restrict__ cell_start, uint32_t* __restrict__ cell_end, const GridState* __restrict__ grid,
          const uint32_t* __restrict__ src_pid, uint32_t* __restrict__ dst_pid) {
  const uint32_t n = grid->n;
  const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  // an odd number of passes leaves the result in the "b" buffers
  const bool in_b = (grid->sort_passes & 1u) != 0u;
  const uint32_t* __restrict__ keys = in_b ? keys_b : keys_a;
  const uint32_t* __restrict__ vals = in_b ? vals_b : vals_a;
  const uint32_t key = keys[r];
  const uint32_t from = vals[r];
  dst_pos[r] = src_pos[from];
  dst_vel[r] = src_vel[from];
  dst_ivel[r] = src_ivel[from];
  skey[r] = key;
  perm_out[r] = from;
  if (src_pid) dst_pid[r] = src_pid[from];
  // keys are < cell_count whenever the grid fits (Morton is monotone in each coordinate); the
  // bound check only matters after a grid overflow, which is reported as CLSPH_EGRID
  const uint32_t count = grid->cell_count;
  if (grid->dense) {
    if (r == 0) {
      if (key < count) cell_start[key] = 0u;
    } else {
      const uint32_t prev = keys[r - 1];
      if (prev != key) {
        if (key < count) cell_start[key] = r;
        if (prev < count) cell_end[prev] = r;
      }
    }
    if (r == n - 1 && key < count) cell_end[key] = n;
  }
}

// 80-byte AoS record = 5 x 16 B: position, velocity, intermediate_velocity, acceleration,
// {density, pressure, grid_index, pad}.
__global__ void __launch_bounds__(256) k_aos_to_soa(const float4* __restrict__ aos, float4* __restrict__ pos,
                                                    float4* __restrict__ vel, float4* __restrict__ ivel,
                                                    float4* __restrict__ aux, uint32_t* __restrict__ skey,
                                                    float4* __restrict__ accel, uint32_t* __restrict__ rrank, uint32_t n) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float4* rec = aos + (size_t)i * 5;
  pos[i] = rec[0];
  vel[i] = rec[1];
  ivel[i] = rec[2];
  if (accel) accel[i] = rec[3];
  const float4 tail = rec[4];
  aux[i] = make_float4(tail.x, tail.y, 0.f, 0.f);
  skey[i] = __float_as_uint(tail.z);
  if (rrank) rrank[i] = i;  // the uploaded order is the reference's order
}

__global__ void __launch_bounds__(256) k_soa_to_aos(const float4* __restrict__ pos, const float4* __restrict__ vel,
                                                    const float4* __restrict__ ivel, const float4* __restrict__ aux,
                                                    const uint32_t* __restrict__ skey, const uint32_t* __restrict__ rrank,
                                                    float4* __restrict__ aos, uint32_t n) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float4* rec = aos + (size_t)(rrank ? rrank[i] : i) * 5;  // sub-cell order: slot in the reference's array
  float4 p = pos[i], v = vel[i], iv = ivel[i];
  p.w = 0.f; v.w = 0.f; iv.w = 0.f;
  const float4 a = aux[i];
  rec[0] = p;
  rec[1] = v;
  rec[2] = iv;
  rec[3] = make_float4(0.f, 0.f, 0.f, 0.f);  // the step leaves acceleration at zero (sph.cl:97-99)
  rec[4] = make_float4(a.x, a.y, __uint_as_float(skey[i]), 0.f);
}

// Reference-form table for the debug tap: table[c] = first sorted index whose key is >= c.
__global__ void __launch_bounds__(256) k_reference_cell_table(const uint32_t* __restrict__ skey,
                                                              const GridState* __restrict__ grid,
                                                              uint32_t* __restrict__ table) {
  const uint32_t n = grid->n, count = grid->cell_count;
  const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  const uint32_t key = skey[r];
  if (r == 0) {
    for (uint32_t c = 0; c <= key && c < count; ++c) table[c] = 0u;
  } else {
    const uint32_t prev = skey[r - 1];
    for (uint32_t c = prev + 1; c <= key && c < count; ++c) table[c] = r;
  }
  if (r == n - 1)
    for (uint32_t c = key + 1; c < count; ++c) table[c] = n;
}

__global__ void __launch_bounds__(256) k_copy_u32(const uint32_t* __restrict__ src, uint32_t* __restrict__ dst, uint32_t n) {
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) dst[i] = src[i];
}

// ---------------------------------------------------------------------------------------------
static inline unsigned blocks_for(uint32_t n, unsigned threads) { return (n + threads - 1) / threads; }

void launch_bounds_reset(BoundsAcc* acc, cudaStream_t stream, uint64_t* launches) {
  k_bounds_reset<<<1, 32, 0, stream>>>(acc);
  if (launches) ++*launches;
}

void launch_bounds(const float4* pos, uint32_t n, BoundsAcc* acc, int sm_count, cudaStream_t stream, uint64_t* launches) {
  const unsigned blocks = std::min<unsigned>(blocks_for(n, 256), (unsigned)sm_count * 8u);
  k_bounds<<<blocks, 256, 0, stream>>>(pos, n, acc);
  if (launches) ++*launches;
}

void launch_grid_setup(BoundsAcc* acc, GridState* grid, float h, uint32_t n, uint32_t cell_capacity, float plane_lo,
                       float plane_hi, bool keep_n, uint32_t sub_mode, uint32_t sub_capacity, uint32_t count_sort, const StepZero& zero,
                       int sm_count, cudaStream_t stream, uint64_t* launches) {
  // with a sub-cell table to zero: enough threads to stream a few MB; without: the one thread that does the set-up
  if (zero.sub_lb)
    k_grid_setup<<<(unsigned)sm_count * 4u, 256, 0, stream>>>(acc, grid, h, n, cell_capacity, plane_lo, plane_hi, keep_n ? 1 : 0, sub_mode,
                                                             sub_capacity, count_sort, zero);
  else
    k_grid_setup<<<1, 32, 0, stream>>>(acc, grid, h, n, cell_capacity, plane_lo, plane_hi, keep_n ? 1 : 0, sub_mode, sub_capacity, 0u,
                                       StepZero{});
  if (launches) ++*launches;
}

void launch_clear_cells(uint32_t* cell_start, uint32_t* cell_end, const GridState* grid, uint32_t cell_capacity,
                        int sm_count, cudaStream_t stream, uint64_t* launches) {
  const unsigned blocks = std::min<unsigned>(std::max(1u, blocks_for(cell_capacity, 256)), (unsigned)sm_count * 8u);
  k_clear_cells<<<blocks, 256, 0, stream>>>(cell_start, cell_end, grid);
  if (launches) ++*launches;
}

void launch_reorder(const StateArrays& src, const StateArrays& dst, const SortBuffers& sort, uint32_t* skey,
                    uint32_t* perm_out, uint32_t* cell_start, uint32_t* cell_end, const GridState* grid,
                    const uint32_t* src_pid, uint32_t* dst_pid, uint32_t n_launch, cudaStream_t stream,
                    uint64_t* launches) {
  k_reorder<<<blocks_for(n_launch, 256), 256, 0, stream>>>(src.pos, src.vel, src.ivel, dst.pos, dst.vel, dst.ivel,
                                                           sort.keys_a, sort.keys_b, sort.vals_a, sort.vals_b, skey,
                                                           perm_out, cell_start, cell_end, grid, src_pid, dst_pid);
  if (launches) ++*launches;
}

// Upload straight from page-locked HOST memory (zero copy): the three vectors a sub-step needs of each 80-byte record
// -- position, velocity, half-step velocity, bytes [0, 48) -- read over the host link by the kernel itself, three
// lanes per record so that a warp's loads are runs of 48 consecutive bytes. Density, pressure and key of the record
// are not read (the sub-step that follows recomputes all three), which leaves one 32-byte sector in five untouched.
__global__ void __launch_bounds__(256) k_aos_to_soa_host(const float4* __restrict__ aos, float4* __restrict__ pos,
                                                         float4* __restrict__ vel, float4* __restrict__ ivel,
                                                         float4* __restrict__ aux, uint32_t* __restrict__ skey,
                                                         uint32_t* __restrict__ rrank, uint32_t n) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t i = t / 3u, part = t - i * 3u;
  if (i >= n) return;
  const float4 v = aos[(size_t)i * 5u + part];
  if (part == 0u) {
    pos[i] = v;
    aux[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    skey[i] = 0u;
    if (rrank) rrank[i] = i;  // the uploaded order is the reference's order
  } else if (part == 1u) {
    vel[i] = v;
  } else {
    ivel[i] = v;
  }
}

void launch_aos_to_soa_host(const void* aos_host_mapped, const StateArrays& dst, float4* aux, uint32_t* skey, uint32_t* rrank,
                            uint32_t n, cudaStream_t stream, uint64_t* launches) {
  const uint64_t threads = (uint64_t)n * 3u;
  k_aos_to_soa_host<<<(unsigned)((threads + 255) / 256), 256, 0, stream>>>((const float4*)aos_host_mapped, dst.pos, dst.vel, dst.ivel, aux,
                                                                         skey, rrank, n);
  if (launches) ++*launches;
}

void launch_aos_to_soa(const void* aos, const StateArrays& dst, float4* aux, uint32_t* skey, float4* accel,
                       uint32_t* rrank, uint32_t n, cudaStream_t stream, uint64_t* launches) {
  k_aos_to_soa<<<blocks_for(n, 256), 256, 0, stream>>>((const float4*)aos, dst.pos, dst.vel, dst.ivel, aux, skey, accel, rrank, n);
  if (launches) ++*launches;
}

// What a frame file needs of a particle (file_save_delegates/houdini_file_saver.cpp:39-62 of the reference reads
// position, velocity and density; colour and mass follow from the density and the parameters): seven floats, in
// the reference's output order.
__global__ void __launch_bounds__(256) k_pack_frame(const float4* __restrict__ pos, const float4* __restrict__ vel,
                                                    const float4* __restrict__ aux, const uint32_t* __restrict__ rrank,
                                                    float* __restrict__ out, uint32_t n) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float4 p = pos[i], v = vel[i];
  float* rec = out + (size_t)(rrank ? rrank[i] : i) * 7;
  rec[0] = p.x; rec[1] = p.y; rec[2] = p.z;
  rec[3] = v.x; rec[4] = v.y; rec[5] = v.z;
  rec[6] = aux[i].x;
}

void launch_pack_frame(const StateArrays& src, const float4* aux, const uint32_t* rrank, float* out, uint32_t n, cudaStream_t stream,
                       uint64_t* launches) {
  k_pack_frame<<<blocks_for(n, 256), 256, 0, stream>>>(src.pos, src.vel, aux, rrank, out, n);
  if (launches) ++*launches;
}

void launch_soa_to_aos(const StateArrays& src, const float4* aux, const uint32_t* skey, const uint32_t* rrank, void* aos,
                       uint32_t n, cudaStream_t stream, uint64_t* launches) {
  k_soa_to_aos<<<blocks_for(n, 256), 256, 0, stream>>>(src.pos, src.vel, src.ivel, aux, skey, rrank, (float4*)aos, n);
  if (launches) ++*launches;
}

void launch_reference_cell_table(const uint32_t* skey, const GridState* grid, uint32_t* table, uint32_t n_launch,
                                 cudaStream_t stream, uint64_t* launches) {
  k_reference_cell_table<<<blocks_for(n_launch, 256), 256, 0, stream>>>(skey, grid, table);
  if (launches) ++*launches;
}

void launch_copy_u32(const uint32_t* src, uint32_t* dst, uint32_t n, cudaStream_t stream, uint64_t* launches) {
  const unsigned blocks = std::min<unsigned>(std::max(1u, blocks_for(n, 256)), 148u * 8u);
  k_copy_u32<<<blocks, 256, 0, stream>>>(src, dst, n);
  if (launches) ++*launches;
}

}  // namespace clsph
