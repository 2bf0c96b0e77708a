// neighbors.cu -- the two neighbour passes: density/pressure and forces.
//
// Replaces kernels/sph.cl:9-62 with forces.cl:15-112 and smoothing.cl:1-34 of the reference.
// The reference lets every particle walk all candidates of its 27 cells (cell side 2h, so about
// 1000 candidates for about 20 real neighbours). Results here are identical in meaning (same
// candidate set, same support test s < support_s <=> sqrt(s)/h < 1, same per-pair formulas), but
// the work is organised for the SM:
//
//   * candidates are staged in shared memory as float4 (x, y, z, index) and culled while staging
//     against the bounding box of the particles that will use them: a candidate farther than h
//     from the box cannot be inside any member's support, exactly (the box test uses the same
//     fused distance formula on component-wise smaller offsets, and fp32 ops are monotone);
//   * the support test runs with lanes across candidates (one coalesced LDS.128 each, every
//     lane busy whatever the cell occupancy), one particle at a time;
//   * the expensive per-pair force terms run one thread per particle over a compacted neighbour
//     list, with register accumulators, so the ten force sums need no cross-lane reduction.
//
// Two organisations of the pair work are built (clsph_context picks one, default = lists):
//   lists   k_density_lists does the search ONCE per whole grid cell, with a second, octant-level
//           cull in shared memory, accumulates the density and writes every particle's neighbour
//           indices to HBM; k_forces_lists then runs one thread per particle over its own list.
//           Particles whose list would exceed the row budget are flagged and redone by the
//           searching force kernel.
//   twice   k_density + k_forces each stage and test on their own (no list memory).
//
// Between the passes the density kernel leaves, per particle j, p_j/rho_j^2 in pos[j].w and
// m/rho_j in vel[j].w (the w lanes are otherwise unused), so a pair costs two 16-byte gathers.
//
// All of these are FP32-issue / shared-memory bound, not HBM bound (SURVEY hard part 1); their
// HBM traffic is the 16-32 B/particle of coalesced reads, L2-resident candidate gathers and, in
// list mode, about 2 x 4 B x (neighbours per particle) of list traffic.
#include <cstdlib>

#include "kernels.cuh"
#include "pair_terms.cuh"

namespace clsph {

namespace {

// ---- two-pass ("twice") kernels and the list-mode fallback: 8 warps per CTA
constexpr int kNbWarps = 8;
constexpr int kNbThreads = kNbWarps * 32;
constexpr int kCandCap = 384;                // staged candidates per warp (float4 each)
constexpr int kListLen = 48;                 // neighbour-list entries per particle before a flush
constexpr int kListStride = kListLen + 1;    // odd word stride: lanes hit distinct banks
constexpr size_t kDensitySmem = (size_t)kNbWarps * kCandCap * sizeof(float4);
constexpr size_t kForceSmem = kDensitySmem + (size_t)kNbWarps * 32 * kListStride * sizeof(uint32_t);

// ---- list mode
constexpr int kDlWarps = 4;                  // warps per CTA of k_density_lists (small CTAs: units vary in size)
constexpr int kDlThreads = kDlWarps * 32;
constexpr int kCap1 = 352;                   // level-1 staged candidates per warp (unit box), + 32 padding
constexpr int kCap2 = 160;                   // level-2 staged candidates per warp (octant group box), + 32 padding
constexpr int kOwnedCellMax = 64;            // cells up to this size are processed whole by one warp
constexpr int kSplitMin = 12;                // targets needed before a unit is cut into octant groups
constexpr int kRing = 3;                     // candidate chunks in flight per warp (cp.async ring, 32 x float4 each)
constexpr int kDlWarpSmem = kCap1 + 32 + kCap2 + 32 + kRing * 32;  // float4 per warp
constexpr size_t kDensityListSmem = (size_t)kDlWarps * kDlWarpSmem * sizeof(float4);
constexpr int kDlCtasPerSm = 5;
constexpr int kFlWarps = 8;                  // warps per CTA of k_forces_lists
constexpr int kTileStride = 33;              // 32 entries per particle + 1: lanes hit distinct banks

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFullMask, v, o);
  return v;
}
__device__ __forceinline__ uint32_t warp_sum_u32(uint32_t v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFullMask, v, o);
  return v;
}
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
__device__ __forceinline__ uint32_t warp_max_u32(uint32_t v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(kFullMask, v, o));
  return v;
}

// Axis-aligned box, identical in every lane.
struct Box {
  float lx, ly, lz, hx, hy, hz;
};
__device__ __forceinline__ Box segment_box(bool in_seg, const float4& p) {
  const float inf = __int_as_float(0x7f800000);
  Box b;
  b.lx = warp_min(in_seg ? p.x : inf);
  b.ly = warp_min(in_seg ? p.y : inf);
  b.lz = warp_min(in_seg ? p.z : inf);
  b.hx = warp_max(in_seg ? p.x : -inf);
  b.hy = warp_max(in_seg ? p.y : -inf);
  b.hz = warp_max(in_seg ? p.z : -inf);
  return b;
}

// Lower bound of |x_i - c|^2 over every particle i inside the box, evaluated with the same
// rounding sequence as dist2_contract on offsets that are component-wise <= the true ones, so
// "bound >= support_s" implies "true s >= support_s" bit-exactly (fp32 ops are monotone).
__device__ __forceinline__ float box_dist2(const Box& b, float cx, float cy, float cz) {
  const float ex = fmaxf(fmaxf(__fsub_rn(b.lx, cx), __fsub_rn(cx, b.hx)), 0.f);
  const float ey = fmaxf(fmaxf(__fsub_rn(b.ly, cy), __fsub_rn(cy, b.hy)), 0.f);
  const float ez = fmaxf(fmaxf(__fsub_rn(b.lz, cz), __fsub_rn(cz, b.hz)), 0.f);
  return __fmaf_rn(ez, ez, __fmaf_rn(ey, ey, __fmul_rn(ex, ex)));
}

// Ranges of the 27 neighbour cells of `key`, one per lane 0..26, in the reference's visiting
// order (z outermost, x innermost: forces.cl:25-27); lane 13 is the cell itself. Other lanes get
// an empty range.
__device__ __forceinline__ uint2 neighbour_cell_range(uint32_t key, const GridState& g,
                                                      const uint32_t* __restrict__ cell_start,
                                                      const uint32_t* __restrict__ cell_end,
                                                      const uint32_t* __restrict__ skey) {
  const unsigned lane = lane_id();
  if (lane >= 27) return make_uint2(0u, 0u);
  const uint32_t cx = compact10(key), cy = compact10(key >> 1), cz = compact10(key >> 2);
  // Cell coordinates of real particles are >= 1 (two cells of padding); with a 0 coordinate the
  // reference's unsigned loop `for (x = c-1; x <= c+1; ++x)` would not run at all.
  if (cx == 0u || cy == 0u || cz == 0u) return make_uint2(0u, 0u);
  const uint32_t x = cx + (lane % 3u) - 1u, y = cy + ((lane / 3u) % 3u) - 1u, z = cz + (lane / 9u) - 1u;
  return cell_range(morton3(x, y, z), g, cell_start, cell_end, skey);
}

// 16-byte asynchronous global -> shared copy (LDGSTS); each lane later reads only what it copied
// itself, so cp.async.wait_group alone orders it.
#ifdef CLSPH_EMU  // tests/emu: CPU build of the kernels for logic tests, the copy completes at once
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) { memcpy(smem_dst, gmem_src, 16); }
__device__ __forceinline__ void cp_async_commit() {}
template <int kPending>
__device__ __forceinline__ void cp_async_wait() {}
#else
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
  const unsigned dst = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int kPending>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(kPending) : "memory"); }
#endif

// Rounds a staged list up to a multiple of 32 with candidates at +infinity (never inside a
// support), so the support-test loops need no bounds check. The lists have 32 spare slots.
__device__ __forceinline__ uint32_t pad_list(float4* list, uint32_t count) {
  const uint32_t padded = (count + 31u) & ~31u;
  const float inf = __int_as_float(0x7f800000);
  if (count + lane_id() < padded) list[count + lane_id()] = make_float4(inf, inf, inf, 0.f);
  return padded;
}

}  // namespace

// =============================================================================================
// Two-pass mode: density + Tait pressure. One warp = 32 consecutive sorted particles; lanes with
// the same cell key form a segment that shares one staged candidate list.
// =============================================================================================
template <bool kTaps>
__global__ void __launch_bounds__(kNbThreads, 4)
k_density(float4* pos, float4* vel, const uint32_t* __restrict__ skey, const uint32_t* __restrict__ cell_start,
          const uint32_t* __restrict__ cell_end, const GridState* __restrict__ grid, const SphConst c,
          float4* __restrict__ aux, uint32_t* __restrict__ cand_count, uint32_t* __restrict__ supp_count) {
  extern __shared__ float4 s_dyn[];
  const unsigned warp = threadIdx.x >> 5, lane = lane_id();
  float4* s_cand = s_dyn + warp * kCandCap;

  const GridState g = *grid;
  const uint32_t base = (blockIdx.x * kNbWarps + warp) * 32u;
  if (base >= g.n) return;
  const uint32_t i = base + lane;
  const bool valid = i < g.n;
  const float4 pi = valid ? pos[i] : make_float4(0.f, 0.f, 0.f, 0.f);
  const uint32_t key_i = valid ? skey[i] : 0xFFFFFFFFu;

  float acc = 0.f;               // sum of (h^2 - s)^3 over the support of particle i
  uint32_t n_supp = 0, n_cand = 0;

  // For every particle of the current segment: test the staged candidates, lanes across candidates.
  auto process = [&](unsigned seg_mask, uint32_t staged) {
    for (unsigned todo = seg_mask; todo; todo &= todo - 1u) {
      const int p = __ffs(todo) - 1;
      const float xi = __shfl_sync(kFullMask, pi.x, p), yi = __shfl_sync(kFullMask, pi.y, p),
                  zi = __shfl_sync(kFullMask, pi.z, p);
      float part = 0.f;
      uint32_t cnt = 0;
      for (uint32_t q = lane; q < staged; q += 32u) {
        const float4 cj = s_cand[q];
        const float s = dist2_contract(xi, yi, zi, cj.x, cj.y, cj.z);
        if (s < c.support_s) {
          const float t = c.h2 - s;
          part = fmaf(t * t, t, part);
          ++cnt;
        }
      }
      part = warp_sum(part);
      if (kTaps) cnt = warp_sum_u32(cnt);
      if ((int)lane == p) {
        acc += part;
        n_supp += cnt;
      }
    }
  };

  for (unsigned remaining = __ballot_sync(kFullMask, valid); remaining;) {
    const int leader = __ffs(remaining) - 1;
    const uint32_t seg_key = __shfl_sync(kFullMask, key_i, leader);
    const bool in_seg = valid && key_i == seg_key;
    const unsigned seg_mask = __ballot_sync(kFullMask, in_seg);
    remaining &= ~seg_mask;
    if (!cell_needs_density(seg_key, g)) continue;  // multi-GPU: outer ghost layer, candidates only

    const Box box = segment_box(in_seg, pi);
    const uint2 rng = neighbour_cell_range(seg_key, g, cell_start, cell_end, skey);
    if (kTaps) {
      const uint32_t total = warp_sum_u32(rng.y - rng.x);
      if (in_seg) n_cand = total;
    }

    uint32_t staged = 0;
    for (int cell = 0; cell < 27; ++cell) {
      const uint32_t first = __shfl_sync(kFullMask, rng.x, cell), end = __shfl_sync(kFullMask, rng.y, cell);
      for (uint32_t j0 = first; j0 < end; j0 += 32u) {
        const uint32_t j = j0 + lane;
        bool keep = j < end;
        float4 pj = make_float4(0.f, 0.f, 0.f, 0.f);
        if (keep) {
          pj = pos[j];
          keep = box_dist2(box, pj.x, pj.y, pj.z) < c.support_s;
        }
        const unsigned m = __ballot_sync(kFullMask, keep);
        if (staged + (uint32_t)__popc(m) > (uint32_t)kCandCap) {  // list full: consume it first
          __syncwarp();
          process(seg_mask, staged);
          __syncwarp();
          staged = 0;
        }
        if (keep) s_cand[staged + __popc(m & lanemask_lt())] = make_float4(pj.x, pj.y, pj.z, __uint_as_float(j));
        staged += __popc(m);
      }
    }
    __syncwarp();
    process(seg_mask, staged);
    __syncwarp();
  }

  if (valid) {
    finish_density(c, acc, i, aux, pos, vel);
    if (kTaps) {
      cand_count[i] = n_cand;
      supp_count[i] = n_supp;
    }
  }
}

// =============================================================================================
// Two-pass mode: forces (pressure, viscosity, surface tension); a = F / rho + g.
// kOnlyOverflow: list-mode fallback; only windows holding a particle whose neighbour list
// overflowed (ncount > list_rows) do any work, and they redo all of their particles.
// =============================================================================================
template <bool kOnlyOverflow>
__global__ void __launch_bounds__(kNbThreads, 2)
k_forces(const float4* __restrict__ pos, const float4* __restrict__ vel, const float4* __restrict__ aux,
         const uint32_t* __restrict__ skey, const uint32_t* __restrict__ cell_start, const uint32_t* __restrict__ cell_end,
         const GridState* __restrict__ grid, const SphConst c, float4* __restrict__ accel,
         const uint32_t* __restrict__ ncount, uint32_t list_rows) {
  extern __shared__ float4 s_dyn[];
  const unsigned warp = threadIdx.x >> 5, lane = lane_id();
  float4* s_cand = s_dyn + warp * kCandCap;
  uint32_t* s_list = reinterpret_cast<uint32_t*>(s_dyn + kNbWarps * kCandCap) + warp * 32 * kListStride;

  const GridState g = *grid;
  const uint32_t base = (blockIdx.x * kNbWarps + warp) * 32u;
  if (base >= g.n) return;
  const uint32_t i = base + lane;
  const bool valid = i < g.n;
  if (kOnlyOverflow) {
    const bool overflowed = valid && ncount[i] > list_rows;
    if (!__any_sync(kFullMask, overflowed)) return;
  }
  const float4 pi = valid ? pos[i] : make_float4(0.f, 0.f, 0.f, 0.f);
  const float4 vi = valid ? vel[i] : make_float4(0.f, 0.f, 0.f, 0.f);
  const float rho_i = valid ? aux[i].x : 1.f;
  const uint32_t key_i = valid ? skey[i] : 0xFFFFFFFFu;

  ForceSums sums;         // register accumulators of this lane's own particle
  uint32_t my_count = 0;  // entries waiting in this lane's neighbour list

  // One thread per particle: consume the lane's own neighbour list (global indices).
  auto flush = [&]() {
    __syncwarp();
    const uint32_t* mine = s_list + lane * kListStride;
    for (uint32_t e = 0; e < my_count; ++e) {
      const uint32_t j = mine[e];
      add_pair(sums, c, j == i, pi, vi, pi.w, pos[j], vel[j]);
    }
    my_count = 0;
    __syncwarp();
  };

  // Support test of the staged candidates for each particle of the segment; survivors are
  // appended to that particle's list (ballot-compacted, so list order = candidate order).
  auto cull = [&](unsigned seg_mask, uint32_t staged) {
    for (unsigned todo = seg_mask; todo; todo &= todo - 1u) {
      const int p = __ffs(todo) - 1;
      const float xi = __shfl_sync(kFullMask, pi.x, p), yi = __shfl_sync(kFullMask, pi.y, p),
                  zi = __shfl_sync(kFullMask, pi.z, p);
      uint32_t count_p = __shfl_sync(kFullMask, my_count, p);
      uint32_t* list_p = s_list + p * kListStride;
      for (uint32_t q0 = 0; q0 < staged; q0 += 32u) {
        const uint32_t q = q0 + lane;
        bool hit = q < staged;
        float4 cj = make_float4(0.f, 0.f, 0.f, 0.f);
        if (hit) {
          cj = s_cand[q];
          hit = dist2_contract(xi, yi, zi, cj.x, cj.y, cj.z) < c.support_s;
        }
        const unsigned m = __ballot_sync(kFullMask, hit);
        if (m == 0u) continue;
        if (count_p + (uint32_t)__popc(m) > (uint32_t)kListLen) {  // list full: run the pair terms now
          if ((int)lane == p) my_count = count_p;
          flush();
          count_p = 0;
        }
        if (hit) list_p[count_p + __popc(m & lanemask_lt())] = __float_as_uint(cj.w);
        count_p += __popc(m);
      }
      if ((int)lane == p) my_count = count_p;
    }
  };

  for (unsigned remaining = __ballot_sync(kFullMask, valid); remaining;) {
    const int leader = __ffs(remaining) - 1;
    const uint32_t seg_key = __shfl_sync(kFullMask, key_i, leader);
    const bool in_seg = valid && key_i == seg_key;
    const unsigned seg_mask = __ballot_sync(kFullMask, in_seg);
    remaining &= ~seg_mask;
    if (!cell_is_owned(seg_key, g)) continue;  // multi-GPU: ghosts get no force

    const Box box = segment_box(in_seg, pi);
    const uint2 rng = neighbour_cell_range(seg_key, g, cell_start, cell_end, skey);

    uint32_t staged = 0;
    for (int cell = 0; cell < 27; ++cell) {
      const uint32_t first = __shfl_sync(kFullMask, rng.x, cell), end = __shfl_sync(kFullMask, rng.y, cell);
      for (uint32_t j0 = first; j0 < end; j0 += 32u) {
        const uint32_t j = j0 + lane;
        bool keep = j < end;
        float4 pj = make_float4(0.f, 0.f, 0.f, 0.f);
        if (keep) {
          pj = pos[j];
          keep = box_dist2(box, pj.x, pj.y, pj.z) < c.support_s;
        }
        const unsigned m = __ballot_sync(kFullMask, keep);
        if (staged + (uint32_t)__popc(m) > (uint32_t)kCandCap) {
          __syncwarp();
          cull(seg_mask, staged);
          __syncwarp();
          staged = 0;
        }
        if (keep) s_cand[staged + __popc(m & lanemask_lt())] = make_float4(pj.x, pj.y, pj.z, __uint_as_float(j));
        staged += __popc(m);
      }
    }
    __syncwarp();
    cull(seg_mask, staged);
    __syncwarp();
  }
  flush();

  if (valid && cell_is_owned(key_i, g)) accel[i] = finish_force(sums, c, rho_i);
}

// =============================================================================================
// List mode, pass 1: density + pressure + per-particle neighbour lists
// =============================================================================================
// nlist is particle-major: nlist[i * list_rows + e] = e-th neighbour (sorted index) of particle
// i, so the survivors of one support-test iteration (all for the same particle) land in one
// 128-byte line. ncount[i] = neighbours found; more than list_rows means the list is incomplete.
//
// Work unit = one grid cell. A cell of up to kOwnedCellMax particles is processed whole (two
// particles per lane) by the warp whose 32-particle window holds the cell's first particle; larger
// cells are processed window by window. Per unit:
//   1. candidates of the 27 cells are culled against the unit's bounding box into shared memory
//      (level 1, ~30 % survive);
//   2. with kSplitMin or more particles the unit is cut at the box centre into 8 octant groups,
//      and level 1 is culled again against each group's box (level 2, ~1/3 survive) -- both culls
//      are exact lower bounds of the true distance, so no neighbour can be lost;
//   3. each particle tests its group's list with lanes across candidates; survivors add to the
//      density and go straight to the particle's list in HBM.
template <bool kTaps>
__global__ void __launch_bounds__(kDlThreads, kDlCtasPerSm)
k_density_lists(float4* pos, float4* vel, const uint32_t* __restrict__ skey, const uint32_t* __restrict__ cell_start,
                const uint32_t* __restrict__ cell_end, const GridState* __restrict__ grid, const SphConst c,
                float4* __restrict__ aux, uint32_t* __restrict__ nlist, uint32_t* __restrict__ ncount,
                uint32_t list_rows, uint32_t* __restrict__ window_counter, uint32_t* __restrict__ cand_count,
                uint32_t* __restrict__ supp_count) {
  extern __shared__ float4 s_dyn[];
  const unsigned warp = threadIdx.x >> 5, lane = lane_id();
  float4* s_l1 = s_dyn + warp * kDlWarpSmem;    // level 1: culled against the unit's box
  float4* s_l2 = s_l1 + kCap1 + 32;             // level 2: culled against one octant group's box
  float4* s_ring = s_l2 + kCap2 + 32;           // raw candidate chunks in flight

  const GridState g = *grid;
  // Persistent warps: every warp keeps taking the next 32-particle window until none is left, so
  // uneven units (empty windows, crowded cells) do not leave warp slots idle.
  for (;;) {
  uint32_t window = 0;
  if (lane == 0) window = atomicAdd(window_counter, 1u);
  window = __shfl_sync(kFullMask, window, 0);
  if ((uint64_t)window * 32u >= g.n) break;
  const uint32_t base = window * 32u;
  const uint32_t i = base + lane;
  const bool valid = i < g.n;
  const uint32_t key_i = valid ? skey[i] : 0xFFFFFFFFu;

  for (unsigned remaining = __ballot_sync(kFullMask, valid); remaining;) {
    const int leader = __ffs(remaining) - 1;
    const uint32_t seg_key = __shfl_sync(kFullMask, key_i, leader);
    const bool in_seg = valid && key_i == seg_key;
    remaining &= ~__ballot_sync(kFullMask, in_seg);
    if (!cell_needs_density(seg_key, g)) continue;  // multi-GPU: outer ghost layer, candidates only

    const uint2 rng = neighbour_cell_range(seg_key, g, cell_start, cell_end, skey);
    const uint32_t cs = __shfl_sync(kFullMask, rng.x, 13), ce = __shfl_sync(kFullMask, rng.y, 13);
    const bool owned = ce - cs <= (uint32_t)kOwnedCellMax;
    if (owned && cs < base) continue;  // an earlier window owns this cell

    // Targets of this unit: two per lane for an owned cell, the lane's own particle otherwise.
    const uint32_t ti0 = owned ? cs + lane : i, ti1 = cs + 32u + lane;
    const bool tv0 = owned ? ti0 < ce : in_seg, tv1 = owned && ti1 < ce;
    const float4 tp0 = tv0 ? pos[ti0] : make_float4(0.f, 0.f, 0.f, 0.f);
    const float4 tp1 = tv1 ? pos[ti1] : make_float4(0.f, 0.f, 0.f, 0.f);
    const unsigned tm0 = __ballot_sync(kFullMask, tv0), tm1 = __ballot_sync(kFullMask, tv1);
    float acc0 = 0.f, acc1 = 0.f;   // sum of (h^2 - s)^3
    uint32_t cnt0 = 0, cnt1 = 0;    // neighbours found so far

    const float inf = __int_as_float(0x7f800000);
    Box box;
    box.lx = warp_min(fminf(tv0 ? tp0.x : inf, tv1 ? tp1.x : inf));
    box.ly = warp_min(fminf(tv0 ? tp0.y : inf, tv1 ? tp1.y : inf));
    box.lz = warp_min(fminf(tv0 ? tp0.z : inf, tv1 ? tp1.z : inf));
    box.hx = warp_max(fmaxf(tv0 ? tp0.x : -inf, tv1 ? tp1.x : -inf));
    box.hy = warp_max(fmaxf(tv0 ? tp0.y : -inf, tv1 ? tp1.y : -inf));
    box.hz = warp_max(fmaxf(tv0 ? tp0.z : -inf, tv1 ? tp1.z : -inf));
    uint32_t n_cand = 0;
    if (kTaps) n_cand = warp_sum_u32(rng.y - rng.x);

    // Support test of a padded staged list for every target in `mask` of one register slot.
    auto cull_slot = [&](unsigned mask, const float4* list, uint32_t padded, const float4& tp, uint32_t ti, float& acc,
                         uint32_t& cnt) {
      for (unsigned todo = mask; todo; todo &= todo - 1u) {
        const int p = __ffs(todo) - 1;
        const float xi = __shfl_sync(kFullMask, tp.x, p), yi = __shfl_sync(kFullMask, tp.y, p),
                    zi = __shfl_sync(kFullMask, tp.z, p);
        uint32_t* row = nlist + (size_t)__shfl_sync(kFullMask, ti, p) * list_rows;
        uint32_t cnt_p = __shfl_sync(kFullMask, cnt, p);
        float part = 0.f;
        for (uint32_t q = lane; q < padded; q += 32u) {  // padded is a multiple of 32: uniform trip count
          const float4 cj = list[q];
          const float s = dist2_contract(xi, yi, zi, cj.x, cj.y, cj.z);
          const bool hit = s < c.support_s;
          if (hit) {
            const float t = c.h2 - s;
            part = fmaf(t * t, t, part);
          }
          const unsigned m = __ballot_sync(kFullMask, hit);
          if (m != 0u) {
            const uint32_t at = cnt_p + __popc(m & lanemask_lt());
            if (hit && at < list_rows) row[at] = __float_as_uint(cj.w);
            cnt_p += __popc(m);
          }
        }
        part = warp_sum(part);
        if ((int)lane == p) {
          acc += part;
          cnt = cnt_p;
        }
      }
    };
    auto cull = [&](unsigned m0, unsigned m1, const float4* list, uint32_t padded) {
      cull_slot(m0, list, padded, tp0, ti0, acc0, cnt0);
      cull_slot(m1, list, padded, tp1, ti1, acc1, cnt1);
    };

    // Level 2: one pass per occupied octant of the unit's box. `padded1` = level-1 entries, padded.
    auto process_level1 = [&](uint32_t padded1) {
      if (__popc(tm0) + __popc(tm1) < kSplitMin) {
        cull(tm0, tm1, s_l1, padded1);
        return;
      }
      const float mx = 0.5f * (box.lx + box.hx), my = 0.5f * (box.ly + box.hy), mz = 0.5f * (box.lz + box.hz);
      const int oct0 = (tp0.x > mx ? 1 : 0) | (tp0.y > my ? 2 : 0) | (tp0.z > mz ? 4 : 0);
      const int oct1 = (tp1.x > mx ? 1 : 0) | (tp1.y > my ? 2 : 0) | (tp1.z > mz ? 4 : 0);
      for (int o = 0; o < 8; ++o) {
        const unsigned m0 = __ballot_sync(kFullMask, tv0 && oct0 == o), m1 = __ballot_sync(kFullMask, tv1 && oct1 == o);
        if ((m0 | m1) == 0u) continue;
        // coordinates <= the centre are in [lo, centre], the others in [centre, hi]
        Box gb;
        gb.lx = (o & 1) ? mx : box.lx; gb.hx = (o & 1) ? box.hx : mx;
        gb.ly = (o & 2) ? my : box.ly; gb.hy = (o & 2) ? box.hy : my;
        gb.lz = (o & 4) ? mz : box.lz; gb.hz = (o & 4) ? box.hz : mz;
        uint32_t staged2 = 0;
        for (uint32_t q = lane; q < padded1; q += 32u) {
          const float4 cj = s_l1[q];
          const bool keep = box_dist2(gb, cj.x, cj.y, cj.z) < c.support_s;  // padding is at +inf: never kept
          const unsigned m = __ballot_sync(kFullMask, keep);
          if (staged2 + (uint32_t)__popc(m) > (uint32_t)kCap2) {
            const uint32_t full2 = pad_list(s_l2, staged2);
            __syncwarp();
            cull(m0, m1, s_l2, full2);
            __syncwarp();
            staged2 = 0;
          }
          if (keep) s_l2[staged2 + __popc(m & lanemask_lt())] = cj;
          staged2 += __popc(m);
        }
        const uint32_t padded2 = pad_list(s_l2, staged2);
        __syncwarp();
        cull(m0, m1, s_l2, padded2);
        __syncwarp();
      }
    };

    // Level 1: the 27 cells in the reference's visiting order, culled against the unit's box.
    // Raw chunks of 32 candidates travel global -> shared through a ring of kRing cp.async
    // groups, so each lane has kRing loads in flight while it tests an older chunk.
    int cur_cell = -1;
    uint32_t cur_j0 = 0, cur_end = 0;
    auto next_chunk = [&]() -> bool {  // advances (cur_cell, cur_j0) to the next non-empty chunk; uniform
      cur_j0 += 32u;
      while (cur_j0 >= cur_end) {
        if (++cur_cell >= 27) return false;
        cur_j0 = __shfl_sync(kFullMask, rng.x, cur_cell);
        cur_end = __shfl_sync(kFullMask, rng.y, cur_cell);
      }
      return true;
    };
    bool live[kRing];
    bool mine_valid[kRing];
    uint32_t mine_j[kRing];
    bool exhausted = false;
    auto issue = [&](int slot) {
      live[slot] = !exhausted && next_chunk();
      if (!live[slot]) exhausted = true;
      mine_j[slot] = cur_j0 + lane;
      mine_valid[slot] = live[slot] && mine_j[slot] < cur_end;
      if (mine_valid[slot]) cp_async16(s_ring + slot * 32 + lane, pos + mine_j[slot]);
      cp_async_commit();  // always, so the group count stays in step with the slots
    };
    uint32_t staged1 = 0;
    auto consume = [&](int slot) {
      cp_async_wait<kRing - 1>();  // all but the kRing-1 newest groups have landed: this slot's did
      bool keep = mine_valid[slot];
      float4 pj = make_float4(0.f, 0.f, 0.f, 0.f);
      if (keep) {
        pj = s_ring[slot * 32 + lane];
        keep = box_dist2(box, pj.x, pj.y, pj.z) < c.support_s;
      }
      const unsigned m = __ballot_sync(kFullMask, keep);
      if (staged1 + (uint32_t)__popc(m) > (uint32_t)kCap1) {
        const uint32_t full1 = pad_list(s_l1, staged1);
        __syncwarp();
        process_level1(full1);
        __syncwarp();
        staged1 = 0;
      }
      if (keep) s_l1[staged1 + __popc(m & lanemask_lt())] = make_float4(pj.x, pj.y, pj.z, __uint_as_float(mine_j[slot]));
      staged1 += __popc(m);
    };
#pragma unroll
    for (int slot = 0; slot < kRing; ++slot) issue(slot);
    for (bool more = true; more;) {
#pragma unroll
      for (int slot = 0; slot < kRing; ++slot) {
        if (!live[slot]) { more = false; break; }
        consume(slot);
        issue(slot);
      }
    }
    cp_async_wait<0>();
    const uint32_t padded1 = pad_list(s_l1, staged1);
    __syncwarp();
    process_level1(padded1);
    __syncwarp();

    if (tv0) {
      finish_density(c, acc0, ti0, aux, pos, vel);
      ncount[ti0] = cnt0;
      if (kTaps) { cand_count[ti0] = n_cand; supp_count[ti0] = cnt0; }
    }
    if (tv1) {
      finish_density(c, acc1, ti1, aux, pos, vel);
      ncount[ti1] = cnt1;
      if (kTaps) { cand_count[ti1] = n_cand; supp_count[ti1] = cnt1; }
    }
  }
  }  // persistent window loop
}

// =============================================================================================
// List mode, pass 2: forces from the stored neighbour lists, one thread per particle.
// The warp first copies its 32 rows, 32 entries at a time, into a shared-memory tile with
// coalesced 128-byte reads; each lane then walks its own row.
// =============================================================================================
// kFast: pair terms through add_pair_fast (sub-cell organisation). kBlocks: resident CTAs per SM the
// register allocation aims at (3: up to 85 registers; 4: 64 registers, a third more warps to hide the
// latency of the neighbour gathers, which is what this kernel waits for).
template <bool kFast, int kBlocks>
__global__ void __launch_bounds__(kFlWarps * 32, kBlocks)
k_forces_lists(const float4* __restrict__ pos, const float4* __restrict__ vel, const float4* __restrict__ aux,
               const uint32_t* __restrict__ nlist, const uint32_t* __restrict__ ncount, uint32_t list_rows,
               const uint32_t* __restrict__ skey, const GridState* __restrict__ grid, const SphConst c,
               float4* __restrict__ accel) {
  __shared__ uint32_t s_tile[kFlWarps][32 * kTileStride];
  const unsigned warp = threadIdx.x >> 5, lane = lane_id();
  const GridState g = *grid;
  const uint32_t n = g.n;
  const uint32_t base = (blockIdx.x * kFlWarps + warp) * 32u;
  if (base >= n) return;
  const uint32_t i = base + lane;
  const float4 pi = i < n ? pos[i] : make_float4(0.f, 0.f, 0.f, 0.f);
  const bool valid = i < n && owned_here(pi.x, skey[min(i, n - 1u)], g);  // multi-GPU: ghosts get no force
  uint32_t count = valid ? ncount[i] : 0u;
  const bool listed = count <= list_rows;  // otherwise redone by k_forces<true>
  if (!listed) count = 0u;
  const float4 vi = valid ? vel[i] : make_float4(0.f, 0.f, 0.f, 0.f);
  ForceSums sums;
  uint32_t* tile = s_tile[warp];
  const uint32_t max_count = warp_max_u32(count);
  for (uint32_t e0 = 0; e0 < max_count; e0 += 32u) {
#pragma unroll
    for (int p0 = 0; p0 < 32; p0 += 8) {  // rows p of the tile <- entries e0 .. e0+31 of particle base+p, 8 loads in flight
      uint32_t v[8];
      bool ok[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        ok[k] = e0 + lane < __shfl_sync(kFullMask, count, p0 + k);
        v[k] = ok[k] ? nlist[(size_t)(base + p0 + k) * list_rows + e0 + lane] : 0u;
      }
#pragma unroll
      for (int k = 0; k < 8; ++k)
        if (ok[k]) tile[(p0 + k) * kTileStride + lane] = v[k];
    }
    __syncwarp();
    const uint32_t mine = count > e0 ? min(count - e0, 32u) : 0u;
    const uint32_t* row = tile + lane * kTileStride;
    uint32_t e = 0;
    for (; e + 2 <= mine; e += 2) {  // two neighbours per trip: four independent gathers in flight
      const uint32_t ja = row[e], jb = row[e + 1];
      const float4 pa = pos[ja], va = vel[ja], pb = pos[jb], vb = vel[jb];
      add_pair_sel<kFast>(sums, c, ja == i, pi, vi, pi.w, pa, va);
      add_pair_sel<kFast>(sums, c, jb == i, pi, vi, pi.w, pb, vb);
    }
    if (e < mine) {
      const uint32_t j = row[e];
      add_pair_sel<kFast>(sums, c, j == i, pi, vi, pi.w, pos[j], vel[j]);
    }
    __syncwarp();
  }
  if (valid && listed) accel[i] = finish_force(sums, c, aux[i].x);
}

// The same pass with the pair terms of pair_terms.cuh's second half: per-run constants factored out of the sums, one
// MUFU.RSQ per pair, no branch -- ~46 instead of ~71 instructions per pair. The particle itself is in its list (the
// density kernels write it there); for that entry 1/r is taken as 0, which leaves exactly its own term. A particle
// with a degenerate pair (coincident particles, smoothing.cl:23; found by the exact test on s) is redone on the spot
// with the reference's formulas. Measured against k_forces_lists<fast> (profiles/r02_c_summary.md): 59 M instead of
// 88 M warp instructions at 1 Mi particles, the same 0.14 ms (both wait on the same gathers); 0.81 instead of 0.91 ms
// at 4 Mi mucus with 45 neighbours per particle -- the library picks by the fluid (context.cu).
template <int kBlocks>
__global__ void __launch_bounds__(kFlWarps * 32, kBlocks)
k_forces_lists_factored(const float4* __restrict__ pos, const float4* __restrict__ vel, const float4* __restrict__ aux,
                    const uint32_t* __restrict__ nlist, const uint32_t* __restrict__ ncount, uint32_t list_rows,
                    const uint32_t* __restrict__ skey, const GridState* __restrict__ grid, const SphConst c,
                    float4* __restrict__ accel) {
  __shared__ uint32_t s_tile[kFlWarps][32 * kTileStride];
  const unsigned warp = threadIdx.x >> 5, lane = lane_id();
  const GridState g = *grid;
  const uint32_t n = g.n;
  const uint32_t base = (blockIdx.x * kFlWarps + warp) * 32u;
  if (base >= n) return;
  const uint32_t i = base + lane;
  const float4 pi = i < n ? pos[i] : make_float4(0.f, 0.f, 0.f, 0.f);
  const bool valid = i < n && owned_here(pi.x, skey[min(i, n - 1u)], g);  // multi-GPU: ghosts get no force
  uint32_t count = valid ? ncount[i] : 0u;
  const bool listed = count <= list_rows;  // otherwise redone by k_forces_sub
  if (!listed) count = 0u;
  const float4 vi = valid ? vel[i] : make_float4(0.f, 0.f, 0.f, 0.f);
  ForceSums sums;
  bool degenerate = false;
  uint32_t* tile = s_tile[warp];
  const uint32_t max_count = warp_max_u32(count);
  for (uint32_t e0 = 0; e0 < max_count; e0 += 32u) {
#pragma unroll
    for (int p0 = 0; p0 < 32; p0 += 8) {  // rows p of the tile <- entries e0 .. e0+31 of particle base+p, 8 loads in flight
      uint32_t v[8];
      bool ok[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        ok[k] = e0 + lane < __shfl_sync(kFullMask, count, p0 + k);
        v[k] = ok[k] ? nlist[(size_t)(base + p0 + k) * list_rows + e0 + lane] : 0u;
      }
#pragma unroll
      for (int k = 0; k < 8; ++k)
        if (ok[k]) tile[(p0 + k) * kTileStride + lane] = v[k];
    }
    __syncwarp();
    const uint32_t mine = count > e0 ? min(count - e0, 32u) : 0u;
    const uint32_t* row = tile + lane * kTileStride;
    uint32_t e = 0;
    for (; e + 2 <= mine; e += 2) {  // two neighbours per trip: four independent gathers in flight
      const uint32_t ja = row[e], jb = row[e + 1];
      const float4 pa = pos[ja], va = vel[ja], pb = pos[jb], vb = vel[jb];
      float sa, sb;
      const TilePair oa = tile_pair_ops(c, pi, vi, pa, va, sa, ja == i);
      const TilePair ob = tile_pair_ops(c, pi, vi, pb, vb, sb, jb == i);
      degenerate |= (sa < c.degenerate_s) | (sb < c.degenerate_s);
      tile_pair_add(sums, oa);
      tile_pair_add(sums, ob);
    }
    if (e < mine) {
      const uint32_t j = row[e];
      float s;
      const TilePair o = tile_pair_ops(c, pi, vi, pos[j], vel[j], s, j == i);
      degenerate |= s < c.degenerate_s;
      tile_pair_add(sums, o);
    }
    __syncwarp();
  }
  if (!(valid && listed)) return;
  if (!degenerate) {
    accel[i] = tile_finish_force(sums, c, aux[i].x, vi.w, true);
    return;
  }
  ForceSums exact;
  const uint32_t* mine = nlist + (size_t)i * list_rows;
  for (uint32_t e = 0; e < count; ++e) {
    const uint32_t j = mine[e];
    add_pair(exact, c, j == i, pi, vi, pi.w, pos[j], vel[j]);
  }
  accel[i] = finish_force(exact, c, aux[i].x);
}

// =============================================================================================
// The same pass WITHOUT the shared-memory tile: every lane reads its own row four entries at a time (one 16-byte
// load that does not allocate in L1: the lists are streamed once, the L1 is for the gathered positions and
// velocities, and the 34 KB tile per CTA took a third of it) and gathers the four neighbours together -- eight
// independent loads in flight per thread. kFactored: pair terms as in k_forces_lists_factored, else add_pair_fast.
// =============================================================================================
__device__ __forceinline__ uint4 load_list4(const uint32_t* p) {
#ifdef CLSPH_EMU
  return *reinterpret_cast<const uint4*>(p);
#else
  uint4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
#endif
}

template <bool kFactored, int kBlocks>
__global__ void __launch_bounds__(128, kBlocks)
k_forces_lists_direct(const float4* __restrict__ pos, const float4* __restrict__ vel, const float4* __restrict__ aux,
                      const uint32_t* __restrict__ nlist, const uint32_t* __restrict__ ncount, uint32_t list_rows,
                      const uint32_t* __restrict__ skey, const GridState* __restrict__ grid, const SphConst c,
                      float4* __restrict__ accel) {
  const GridState g = *grid;
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= g.n) return;
  const float4 pi = pos[i];
  if (!owned_here(pi.x, skey[i], g)) return;  // multi-GPU: ghosts get no force
  const uint32_t count = ncount[i];
  if (count > list_rows) return;  // redone by k_forces_sub
  const float4 vi = vel[i];
  const uint32_t* row = nlist + (size_t)i * list_rows;  // 16-byte aligned: list_rows is a multiple of 8 (or even and >= 4 ...)
  ForceSums sums;
  bool degenerate = false;
  auto pair = [&](uint32_t j, const float4& pj, const float4& vj) {
    if (kFactored) {
      float s;
      const TilePair o = tile_pair_ops(c, pi, vi, pj, vj, s, j == i);
      degenerate |= s < c.degenerate_s;
      tile_pair_add(sums, o);
    } else {
      add_pair_sel<true>(sums, c, j == i, pi, vi, pi.w, pj, vj);
    }
  };
  // the next four entries are requested before the current four neighbours are gathered: the list load (L2, the
  // lists do not allocate in L1) is otherwise the head of every trip's dependency chain
  uint32_t e = 0;
  uint4 q = count ? load_list4(row) : make_uint4(0u, 0u, 0u, 0u);
  for (; e + 4u <= count; e += 4u) {
    const uint4 qn = e + 4u < count ? load_list4(row + e + 4u) : q;
    const float4 pa = pos[q.x], va = vel[q.x], pb = pos[q.y], vb = vel[q.y], pc = pos[q.z], vc = vel[q.z], pd = pos[q.w], vd = vel[q.w];
    pair(q.x, pa, va);
    pair(q.y, pb, vb);
    pair(q.z, pc, vc);
    pair(q.w, pd, vd);
    q = qn;
  }
  if (e < count) {  // one to three entries left in the last quad (already in q)
    const uint32_t left = count - e;
    const uint32_t jb = left > 1u ? q.y : q.x, jc = left > 2u ? q.z : q.x;
    const float4 pa = pos[q.x], va = vel[q.x], pb = pos[jb], vb = vel[jb], pc = pos[jc], vc = vel[jc];
    pair(q.x, pa, va);
    if (left > 1u) pair(jb, pb, vb);
    if (left > 2u) pair(jc, pc, vc);
  }
  if (!kFactored) {
    accel[i] = finish_force(sums, c, aux[i].x);
    return;
  }
  if (!degenerate) {
    accel[i] = tile_finish_force(sums, c, aux[i].x, vi.w, true);
    return;
  }
  ForceSums exact;
  for (uint32_t k = 0; k < count; ++k) {
    const uint32_t j = row[k];
    add_pair(exact, c, j == i, pi, vi, pi.w, pos[j], vel[j]);
  }
  accel[i] = finish_force(exact, c, aux[i].x);
}

// ---------------------------------------------------------------------------------------------
void neighbors_init() {
  cudaFuncSetAttribute(k_density<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kDensitySmem);
  cudaFuncSetAttribute(k_density<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kDensitySmem);
  cudaFuncSetAttribute(k_forces<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kForceSmem);
  cudaFuncSetAttribute(k_forces<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kForceSmem);
  cudaFuncSetAttribute(k_density_lists<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kDensityListSmem);
  cudaFuncSetAttribute(k_density_lists<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kDensityListSmem);
}

void launch_density(float4* pos, float4* vel, const uint32_t* skey, const uint32_t* cell_start, const uint32_t* cell_end,
                    const GridState* grid, const SphConst& c, float4* aux, const NeighbourLists& lists,
                    const DebugTaps& taps, bool debug, uint32_t n_launch, int sm_count, cudaStream_t stream,
                    uint64_t* launches) {
  uint32_t* cand = debug ? taps.candidate_count : nullptr;
  uint32_t* supp = debug ? taps.support_count : nullptr;
  if (lists.rows) {
    // persistent: one resident set of CTAs, warps pull 32-particle windows from a counter
    const unsigned needed = (n_launch + kDlThreads - 1) / kDlThreads;
    const unsigned blocks = std::max(1u, std::min(needed, (unsigned)sm_count * kDlCtasPerSm));
    cudaMemsetAsync(lists.window_counter, 0, sizeof(uint32_t), stream);
    if (debug)
      k_density_lists<true><<<blocks, kDlThreads, kDensityListSmem, stream>>>(pos, vel, skey, cell_start, cell_end, grid, c,
                                                                              aux, lists.entries, lists.count, lists.rows,
                                                                              lists.window_counter, cand, supp);
    else
      k_density_lists<false><<<blocks, kDlThreads, kDensityListSmem, stream>>>(pos, vel, skey, cell_start, cell_end, grid, c,
                                                                               aux, lists.entries, lists.count, lists.rows,
                                                                               lists.window_counter, cand, supp);
  } else {
    const unsigned blocks = (n_launch + kNbThreads - 1) / kNbThreads;
    if (debug)
      k_density<true><<<blocks, kNbThreads, kDensitySmem, stream>>>(pos, vel, skey, cell_start, cell_end, grid, c, aux, cand, supp);
    else
      k_density<false><<<blocks, kNbThreads, kDensitySmem, stream>>>(pos, vel, skey, cell_start, cell_end, grid, c, aux, cand, supp);
  }
  if (launches) ++*launches;
}

void launch_forces(const float4* pos, const float4* vel, const float4* aux, const uint32_t* skey,
                   const uint32_t* cell_start, const uint32_t* cell_end, const GridState* grid, const SphConst& c,
                   const NeighbourLists& lists, bool search_fallback, bool fast_pairs, bool dense_occupancy, float4* accel,
                   uint32_t n_launch, cudaStream_t stream, uint64_t* launches, bool factored) {
  const unsigned blocks = (n_launch + kNbThreads - 1) / kNbThreads;
  // Default for the sub-cell order: the variant without the shared-memory tile (measured, profiles/r02_s_*: with the
  // factored pair terms 0.125 instead of 0.137 ms at 1 Mi water, 0.753 instead of 0.793 at 4 Mi mucus).
  // CLSPH_FORCES_DIRECT=0 selects the tile kernels, =2 the 72-register build.
  static const int direct = [] { const char* e = getenv("CLSPH_FORCES_DIRECT"); return e ? atoi(e) : 1; }();
  if (lists.rows && direct && fast_pairs && !search_fallback && (lists.rows % 4u) == 0u) {
    const unsigned dblocks = (n_launch + 127) / 128;
    if (factored && direct == 2) k_forces_lists_direct<true, 6><<<dblocks, 128, 0, stream>>>(pos, vel, aux, lists.entries, lists.count, lists.rows, skey, grid, c, accel);
    else if (factored) k_forces_lists_direct<true, 8><<<dblocks, 128, 0, stream>>>(pos, vel, aux, lists.entries, lists.count, lists.rows, skey, grid, c, accel);
    else if (direct == 2) k_forces_lists_direct<false, 6><<<dblocks, 128, 0, stream>>>(pos, vel, aux, lists.entries, lists.count, lists.rows, skey, grid, c, accel);
    else k_forces_lists_direct<false, 8><<<dblocks, 128, 0, stream>>>(pos, vel, aux, lists.entries, lists.count, lists.rows, skey, grid, c, accel);
    if (launches) ++*launches;
  } else if (lists.rows && factored) {  // pair terms with the constants factored out of the sums (sub-cell order, fast pairs)
    const unsigned lblocks = (n_launch + kFlWarps * 32 - 1) / (kFlWarps * 32);
    if (dense_occupancy)
      k_forces_lists_factored<4><<<lblocks, kFlWarps * 32, 0, stream>>>(pos, vel, aux, lists.entries, lists.count, lists.rows, skey, grid, c, accel);
    else
      k_forces_lists_factored<3><<<lblocks, kFlWarps * 32, 0, stream>>>(pos, vel, aux, lists.entries, lists.count, lists.rows, skey, grid, c, accel);
    if (launches) ++*launches;
  } else if (lists.rows) {
    const unsigned lblocks = (n_launch + kFlWarps * 32 - 1) / (kFlWarps * 32);
    // <false, 3> is the instantiation the GPU parity suite has passed; the others are selected by the options
    // fast_pairs / forces_blocks (clsph_cuda.h)
    if (fast_pairs && dense_occupancy)
      k_forces_lists<true, 4><<<lblocks, kFlWarps * 32, 0, stream>>>(pos, vel, aux, lists.entries, lists.count, lists.rows, skey,
                                                                    grid, c, accel);
    else if (fast_pairs)
      k_forces_lists<true, 3><<<lblocks, kFlWarps * 32, 0, stream>>>(pos, vel, aux, lists.entries, lists.count, lists.rows, skey,
                                                                    grid, c, accel);
    else if (dense_occupancy)
      k_forces_lists<false, 4><<<lblocks, kFlWarps * 32, 0, stream>>>(pos, vel, aux, lists.entries, lists.count, lists.rows, skey,
                                                                     grid, c, accel);
    else
      k_forces_lists<false, 3><<<lblocks, kFlWarps * 32, 0, stream>>>(pos, vel, aux, lists.entries, lists.count, lists.rows, skey,
                                                                     grid, c, accel);
    if (launches) ++*launches;
    if (search_fallback) {
      // particles with more neighbours than list rows: redone with the searching kernel (exits at once elsewhere)
      k_forces<true><<<blocks, kNbThreads, kForceSmem, stream>>>(pos, vel, aux, skey, cell_start, cell_end, grid, c, accel,
                                                                 lists.count, lists.rows);
      if (launches) ++*launches;
    }
  } else {
    k_forces<false><<<blocks, kNbThreads, kForceSmem, stream>>>(pos, vel, aux, skey, cell_start, cell_end, grid, c, accel,
                                                                nullptr, 0u);
    if (launches) ++*launches;
  }
}

}  // namespace clsph
