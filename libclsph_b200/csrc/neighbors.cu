// neighbors.cu -- the two neighbour passes: density/pressure and forces.
//
// Replaces kernels/sph.cl:9-62 with forces.cl:15-112 and smoothing.cl:1-34 of the reference.
// The reference lets every particle walk all candidates of its 27 cells (cell side 2h, so about
// 1000 candidates for about 20 real neighbours). Results here are identical in meaning (same
// candidate set, same support test s < support_s <=> sqrt(s)/h < 1, same per-pair formulas), but
// the work is organised for the SM:
//
//   * one warp owns 32 consecutive cell-sorted particles; lanes with the same cell key form a
//     "segment" that shares one candidate list;
//   * the candidates of the segment's 27 cells are staged ONCE in shared memory as float4
//     (x, y, z, index), and while staging they are culled against the segment's bounding box:
//     a candidate farther than h from the box cannot be inside any member's support, exactly
//     (the box test uses the same fused distance formula on component-wise smaller offsets).
//     This keeps about 30 % of the 27-cell candidates;
//   * the support test then runs with lanes across candidates (one coalesced LDS.128 each,
//     every lane busy whatever the cell occupancy), per particle of the segment;
//   * density: survivors accumulate (h^2 - s)^3 and one shuffle reduction per particle;
//   * forces: survivors are compacted (ballot + popc) into a per-particle neighbour list in
//     shared memory; the expensive pair terms then run one thread per particle over its own
//     list with register accumulators, so no cross-lane reduction of the ten force sums.
//
// Both kernels are FP32-issue / shared-memory bound, not HBM bound (SURVEY hard part 1); their
// HBM traffic is the 16-32 B/particle of coalesced reads plus L2-resident candidate gathers.
#include "kernels.cuh"

namespace clsph {

namespace {

constexpr int kNbWarps = 8;                  // warps per CTA
constexpr int kNbThreads = kNbWarps * 32;
constexpr int kCandCap = 384;                // staged candidates per warp (float4 each)
constexpr int kListLen = 48;                 // neighbour-list entries per particle before a flush
constexpr int kListStride = kListLen + 1;    // odd word stride: lanes hit distinct banks

constexpr size_t kDensitySmem = (size_t)kNbWarps * kCandCap * sizeof(float4);
constexpr size_t kForceSmem = kDensitySmem + (size_t)kNbWarps * 32 * kListStride * sizeof(uint32_t);

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

// Bounding box of the segment's particles, identical in every lane.
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
// order (z outermost, x innermost: forces.cl:25-27). Other lanes get an empty range.
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

}  // namespace

// =============================================================================================
// Density + Tait pressure
// =============================================================================================
template <bool kTaps>
__global__ void __launch_bounds__(kNbThreads, 4)
k_density(const float4* __restrict__ pos, const uint32_t* __restrict__ skey, const uint32_t* __restrict__ cell_start,
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
    // forces.cl:33-36 / smoothing.cl:1-4: rho = sum m * C6 * (h^2 - r^2)^3 ; sph.cl:37-39 Tait.
    const float rho = c.mass * c.c_poly6 * acc;
    const float q = rho / c.rho0;
    const float q2 = q * q, q4 = q2 * q2;
    const float prs = c.K * (q4 * q2 * q - 1.f);
    aux[i] = make_float4(rho, prs, prs / (rho * rho), c.mass / rho);
    if (kTaps) {
      cand_count[i] = n_cand;
      supp_count[i] = n_supp;
    }
  }
}

// =============================================================================================
// Forces: pressure (spiky), viscosity, surface tension (poly6 colour field); a = F / rho + g
// =============================================================================================
__global__ void __launch_bounds__(kNbThreads, 2)
k_forces(const float4* __restrict__ pos, const float4* __restrict__ vel, const float4* __restrict__ aux,
         const uint32_t* __restrict__ skey, const uint32_t* __restrict__ cell_start, const uint32_t* __restrict__ cell_end,
         const GridState* __restrict__ grid, const SphConst c, float4* __restrict__ accel) {
  extern __shared__ float4 s_dyn[];
  const unsigned warp = threadIdx.x >> 5, lane = lane_id();
  float4* s_cand = s_dyn + warp * kCandCap;
  uint32_t* s_list = reinterpret_cast<uint32_t*>(s_dyn + kNbWarps * kCandCap) + warp * 32 * kListStride;

  const GridState g = *grid;
  const uint32_t base = (blockIdx.x * kNbWarps + warp) * 32u;
  if (base >= g.n) return;
  const uint32_t i = base + lane;
  const bool valid = i < g.n;
  const float4 pi = valid ? pos[i] : make_float4(0.f, 0.f, 0.f, 0.f);
  const float4 vi = valid ? vel[i] : make_float4(0.f, 0.f, 0.f, 0.f);
  const float4 ai = valid ? aux[i] : make_float4(1.f, 0.f, 0.f, 0.f);  // rho, p, p/rho^2, m/rho
  const uint32_t key_i = valid ? skey[i] : 0xFFFFFFFFu;

  // register accumulators of this lane's own particle
  float px = 0.f, py = 0.f, pz = 0.f;   // pressure term        forces.cl:70-77
  float wx = 0.f, wy = 0.f, wz = 0.f;   // viscosity term       forces.cl:79-85
  float nx = 0.f, ny = 0.f, nz = 0.f;   // colour-field normal  forces.cl:88-91
  float lap = 0.f;                      // colour-field laplacian forces.cl:93-97
  uint32_t my_count = 0;                // entries waiting in this lane's neighbour list

  // One thread per particle: consume the lane's own neighbour list (global indices).
  auto flush = [&]() {
    __syncwarp();
    const uint32_t* mine = s_list + lane * kListStride;
    for (uint32_t e = 0; e < my_count; ++e) {
      const uint32_t j = mine[e];
      const float4 pj = pos[j];
      const float4 vj = vel[j];
      const float4 aj = aux[j];
      const float dx = pi.x - pj.x, dy = pi.y - pj.y, dz = pi.z - pj.z;
      const float s = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
      const float r = sqrtf(s);
      const float mass_over_rho = aj.w;
      if (j != i) {
        float gx, gy, gz;
        if (r < 0.0000001f) {  // smoothing.cl:23-25 (erratum E3): a scalar broadcast to x, y, z
          gx = gy = gz = c.spiky_degenerate;
        } else {               // smoothing.cl:26-28 with the window == 1
          const float hr = c.h - r;
          const float k = c.c_spiky * hr * hr / r;
          gx = k * dx; gy = k * dy; gz = k * dz;
        }
        const float pc = (aj.z + ai.z) * c.mass;
        px = fmaf(pc, gx, px); py = fmaf(pc, gy, py); pz = fmaf(pc, gz, pz);
        const float vc = mass_over_rho * (c.c_visc * (c.h - r));  // smoothing.cl:31-34
        wx = fmaf(vj.x - vi.x, vc, wx); wy = fmaf(vj.y - vi.y, vc, wy); wz = fmaf(vj.z - vi.z, vc, wz);
      }
      const float t = c.h2 - r * r;
      const float gc = mass_over_rho * (c.c_poly6_grad * t * t);  // smoothing.cl:6-10
      nx = fmaf(gc, dx, nx); ny = fmaf(gc, dy, ny); nz = fmaf(gc, dz, nz);
      lap = fmaf(mass_over_rho, c.c_poly6_lap * t * (3.f * c.h2 - 7.f * r * r), lap);  // smoothing.cl:12-17
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

  if (valid) {
    const float rho = ai.x;
    // forces.cl:103-109
    float fx = -rho * px + wx * c.mu, fy = -rho * py + wy * c.mu, fz = -rho * pz + wz * c.mu;
    const float nlen = sqrtf(fmaf(nz, nz, fmaf(ny, ny, nx * nx)));
    if (nlen > c.tension_threshold) {
      const float k = -c.sigma * lap / nlen;
      fx = fmaf(k, nx, fx); fy = fmaf(k, ny, fy); fz = fmaf(k, nz, fz);
    }
    // sph.cl:53-58
    accel[i] = make_float4(fx / rho + c.gx, fy / rho + c.gy, fz / rho + c.gz, 0.f);
  }
}

// ---------------------------------------------------------------------------------------------
void neighbors_init() {
  cudaFuncSetAttribute(k_density<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kDensitySmem);
  cudaFuncSetAttribute(k_density<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kDensitySmem);
  cudaFuncSetAttribute(k_forces, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kForceSmem);
}

void launch_density(const float4* pos, const uint32_t* skey, const uint32_t* cell_start, const uint32_t* cell_end,
                    const GridState* grid, const SphConst& c, float4* aux, const DebugTaps& taps, bool debug,
                    uint32_t n_launch, cudaStream_t stream, uint64_t* launches) {
  const unsigned blocks = (n_launch + kNbThreads - 1) / kNbThreads;
  if (debug)
    k_density<true><<<blocks, kNbThreads, kDensitySmem, stream>>>(pos, skey, cell_start, cell_end, grid, c, aux,
                                                                  taps.candidate_count, taps.support_count);
  else
    k_density<false><<<blocks, kNbThreads, kDensitySmem, stream>>>(pos, skey, cell_start, cell_end, grid, c, aux,
                                                                   nullptr, nullptr);
  if (launches) ++*launches;
}

void launch_forces(const float4* pos, const float4* vel, const float4* aux, const uint32_t* skey,
                   const uint32_t* cell_start, const uint32_t* cell_end, const GridState* grid, const SphConst& c,
                   float4* accel, uint32_t n_launch, cudaStream_t stream, uint64_t* launches) {
  const unsigned blocks = (n_launch + kNbThreads - 1) / kNbThreads;
  k_forces<<<blocks, kNbThreads, kForceSmem, stream>>>(pos, vel, aux, skey, cell_start, cell_end, grid, c, accel);
  if (launches) ++*launches;
}

}  // namespace clsph
