// integrate.cu -- leapfrog advection with speed clamp, segment-vs-triangle collision and
// response, plus the AABB reduction the NEXT sub-step needs (so it never costs a pass of its own).
//
// Replaces kernels/sph.cl:64-112, kernels/advection.cl:6-23 and kernels/collisions.cl:15-129 of
// the reference. Every hit / no-hit decision is a discrete outcome, so this file evaluates the
// reference's expressions with explicit round-to-nearest intrinsics in the reference's
// operation order (no FMA contraction except inside dot()/length(), which the oracle contract
// defines as fused): given identical inputs the results are bit-identical to the oracle.
//
// Per-triangle quantities that do not depend on the particle (edges, their dot products, the
// barycentric determinant, |n|) are evaluated once in k_prepare_faces with the same operations,
// which is where the reference's per-particle-per-face redundancy goes.
#include <cmath>

#include "kernels.cuh"

namespace clsph {

namespace {

struct V3 {
  float x, y, z;
};
__device__ __forceinline__ V3 mk(float x, float y, float z) { V3 r; r.x = x; r.y = y; r.z = z; return r; }
__device__ __forceinline__ V3 add(V3 a, V3 b) { return mk(__fadd_rn(a.x, b.x), __fadd_rn(a.y, b.y), __fadd_rn(a.z, b.z)); }
__device__ __forceinline__ V3 sub(V3 a, V3 b) { return mk(__fsub_rn(a.x, b.x), __fsub_rn(a.y, b.y), __fsub_rn(a.z, b.z)); }
__device__ __forceinline__ V3 scale(V3 a, float s) { return mk(__fmul_rn(a.x, s), __fmul_rn(a.y, s), __fmul_rn(a.z, s)); }
__device__ __forceinline__ V3 divs(V3 a, float s) { return mk(__fdiv_rn(a.x, s), __fdiv_rn(a.y, s), __fdiv_rn(a.z, s)); }
__device__ __forceinline__ V3 neg(V3 a) { return mk(-a.x, -a.y, -a.z); }
__device__ __forceinline__ float dot(V3 a, V3 b) { return __fmaf_rn(a.z, b.z, __fmaf_rn(a.y, b.y, __fmul_rn(a.x, b.x))); }
__device__ __forceinline__ float length(V3 a) { return __fsqrt_rn(dot(a, a)); }

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

constexpr uint32_t kMaxCollisionIters = 64;  // erratum E9: the reference loop has no cap

}  // namespace

__global__ void k_prepare_faces(const float* __restrict__ normals, const float* __restrict__ vertices,
                                const uint32_t* __restrict__ indices, uint32_t face_count, Face* __restrict__ faces) {
  const uint32_t f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= face_count) return;
  const V3 n = mk(normals[3 * f + 0], normals[3 * f + 1], normals[3 * f + 2]);
  const uint32_t i0 = indices[3 * f + 0], i1 = indices[3 * f + 1], i2 = indices[3 * f + 2];
  const V3 a = mk(vertices[3 * i0 + 0], vertices[3 * i0 + 1], vertices[3 * i0 + 2]);
  const V3 b = mk(vertices[3 * i1 + 0], vertices[3 * i1 + 1], vertices[3 * i1 + 2]);
  const V3 cc = mk(vertices[3 * i2 + 0], vertices[3 * i2 + 1], vertices[3 * i2 + 2]);
  const V3 u = sub(b, a), v = sub(cc, a);  // collisions.cl:49-50
  Face o;
  o.nx = n.x; o.ny = n.y; o.nz = n.z; o.nlen = length(n);
  o.ax = a.x; o.ay = a.y; o.az = a.z; o.uv = dot(u, v);
  o.ux = u.x; o.uy = u.y; o.uz = u.z; o.uu = dot(u, u);
  o.vx = v.x; o.vy = v.y; o.vz = v.z; o.vv = dot(v, v);
  o.det = __fsub_rn(__fmul_rn(o.uv, o.uv), __fmul_rn(o.uu, o.vv));  // collisions.cl:71
  o.pad0 = o.pad1 = o.pad2 = 0.f;
  faces[f] = o;
}

// kGrid: only the faces registered in the face-grid cells touched by the segment are tested (plus
// the grid's "global" faces). A face that passes the hit test has its intersection point inside the
// segment's box and inside the triangle, both up to rounding that is orders of magnitude below the
// padding used at registration (build_face_grid), so it is registered in the cell holding that
// point: no hit is lost. A face may be met twice (two cells) and out of index order, so "ties go to
// the later face" (collisions.cl:77-80, where faces come in ascending order) is applied as
// "nearer wins; at equal distance the higher face index wins", which is the same thing.
template <bool kGrid>
__global__ void __launch_bounds__(256, 4)  // 64 registers (60 bytes of spills): half instead of a third of the warps resident
k_integrate(float4* __restrict__ pos, float4* __restrict__ vel, float4* __restrict__ ivel,
            const float4* __restrict__ accel, const uint32_t* __restrict__ skey, const Face* __restrict__ faces,
            uint32_t face_count, const FaceGrid fg, const GridState* __restrict__ grid, const SphConst c,
            BoundsAcc* next_bounds, uint32_t* __restrict__ iters_tap) {
  const GridState g = *grid;
  const uint32_t n = g.n;
  const bool sliced = slab_is_cut(g);  // multi-GPU: ghosts are not advanced
  // sub-cell order across GPUs: the w lane of the half-step velocity marks "advanced here", i.e. owned in this
  // sub-step, which is what the next exchange and the export go by (the position may since have left the slab)
  const float owned_mark = (sliced && g.sub) ? 1.f : 0.f;
  float lo[3] = {2147483648.f, 2147483648.f, 2147483648.f};
  float hi[3] = {-2147483648.f, -2147483648.f, -2147483648.f};

  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const float4 p4 = pos[i];
    if (sliced && !owned_here(p4.x, skey[i], g)) continue;
    const float4 iv4 = ivel[i], a4 = accel[i];
    V3 x = mk(p4.x, p4.y, p4.z);
    V3 v = mk(iv4.x, iv4.y, iv4.z);
    V3 acc = mk(a4.x, a4.y, a4.z);
    const V3 v_half_old = v;
    float time_to_go = c.dt;  // sph.cl:76
    uint32_t iters = 0;
    bool collided;
    do {
      // advection.cl:14-20
      V3 nv = add(v, scale(acc, time_to_go));
      const float speed = length(nv);
      if (speed > c.vmax) nv = scale(divs(nv, speed), c.vmax);
      const V3 np = add(x, scale(nv, time_to_go));

      // collisions.cl:15-89: nearest hit of x -> np over all faces, ties go to the later face
      const V3 travel = sub(np, x);
      const float travel_len = length(travel);
      collided = false;
      V3 hit_n = mk(0.f, 0.f, 0.f), hit_p = mk(0.f, 0.f, 0.f);
      float hit_depth = 0.f, hit_dist = 0.f;
      uint32_t hit_face = 0;
      auto test_face = [&](uint32_t f) {
        const float4* fr = reinterpret_cast<const float4*>(faces + f);
        const float4 f0 = __ldg(fr + 0), f1 = __ldg(fr + 1);
        const V3 n0 = mk(f0.x, f0.y, f0.z);
        const V3 a = mk(f1.x, f1.y, f1.z);
        const float nd = dot(n0, travel);
        if (nd == 0.f) return;  // :54-56 (the oriented denominator is +-nd)
        const float na = dot(n0, sub(a, x));
        // The plane parameter is r = (+-na) / (+-nd) = na / nd whatever the orientation, and a hit
        // needs 0 <= r <= 1 (:58-60). Two division-free rejections that can never drop such a face:
        // opposite signs with a quotient that cannot underflow to -0, and |na| beyond |nd| by more
        // than the half ulp that could still round the quotient down to 1. Nearly every face of a
        // scene leaves here: a sub-step moves a particle by millimetres.
        if ((na < 0.f) != (nd < 0.f) && fabsf(na) > 1e-30f * fabsf(nd)) return;
        if (fabsf(na) > fabsf(nd) * 1.0000002f) return;
        // :27-29 orient the normal along the travel direction
        const bool flip = __fdiv_rn(nd, __fmul_rn(f0.w, travel_len)) <= 0.f;
        const float denom = flip ? -nd : nd;  // dot(-n, d) == -dot(n, d) exactly
        const float r = __fdiv_rn(flip ? -na : na, denom);  // :58
        if (!(0.f <= r && r <= 1.f)) return;
        const float4 f2 = __ldg(fr + 2), f3 = __ldg(fr + 3), f4 = __ldg(fr + 4);
        const V3 u = mk(f2.x, f2.y, f2.z), vv3 = mk(f3.x, f3.y, f3.z);
        const float uu = f2.w, vv = f3.w, uv = f1.w, det = f4.x;
        const V3 hit = add(x, scale(travel, r));  // :61
        const V3 w = sub(hit, a);
        const float wv = dot(w, vv3), wu = dot(w, u);
        const float s = __fdiv_rn(__fsub_rn(__fmul_rn(uv, wv), __fmul_rn(vv, wu)), det);  // :72
        const float t = __fdiv_rn(__fsub_rn(__fmul_rn(uv, wu), __fmul_rn(uu, wv)), det);  // :73
        if (s >= 0.f && t >= 0.f && __fadd_rn(s, t) <= 1.f) {
          const float dist = length(sub(x, hit));
          if (collided && (dist > hit_dist || (dist == hit_dist && f < hit_face))) return;  // :77-80
          hit_n = flip ? neg(n0) : n0;
          hit_p = hit;
          hit_depth = length(sub(np, hit));
          hit_dist = dist;
          hit_face = f;
          collided = true;
        }
      };
      // Face lists to visit: every face (one list, ids = 0 .. face_count-1), or the grid's global
      // list followed by the list of each cell the segment's box overlaps.
      int cx0 = 0, cy0 = 0, cz0 = 0, wx = 0, wy = 0;
      int lists = -1;  // -1: every face
      if (kGrid) {
        // cells overlapped by the segment's box, widened by 1 % of a cell (the registration is widened too)
        const float x0 = (fminf(x.x, np.x) - fg.ox) * fg.inv - 0.01f, x1 = (fmaxf(x.x, np.x) - fg.ox) * fg.inv + 0.01f;
        const float y0 = (fminf(x.y, np.y) - fg.oy) * fg.inv - 0.01f, y1 = (fmaxf(x.y, np.y) - fg.oy) * fg.inv + 0.01f;
        const float z0 = (fminf(x.z, np.z) - fg.oz) * fg.inv - 0.01f, z1 = (fmaxf(x.z, np.z) - fg.oz) * fg.inv + 0.01f;
        // anything odd (NaN, a huge span) keeps the full scan
        const bool finite = x0 <= x1 && y0 <= y1 && z0 <= z1 && fabsf(x0) < 1e9f && fabsf(x1) < 1e9f && fabsf(y0) < 1e9f &&
                            fabsf(y1) < 1e9f && fabsf(z0) < 1e9f && fabsf(z1) < 1e9f;
        if (finite) {
          cx0 = max(0, (int)floorf(x0));
          cy0 = max(0, (int)floorf(y0));
          cz0 = max(0, (int)floorf(z0));
          wx = min(fg.nx - 1, (int)floorf(x1)) - cx0 + 1;
          wy = min(fg.ny - 1, (int)floorf(y1)) - cy0 + 1;
          const int wz = min(fg.nz - 1, (int)floorf(z1)) - cz0 + 1;
          // a box entirely outside the grid meets no registered face: only the global list remains
          const long long cells = (wx <= 0 || wy <= 0 || wz <= 0) ? 0 : (long long)wx * wy * wz;
          if (cells <= 64) lists = 1 + (int)cells;
        }
      }
      for (int li = 0; li < (lists < 0 ? 1 : lists); ++li) {
        const uint32_t* ids = nullptr;
        uint32_t e = 0, e1 = face_count;
        if (kGrid && lists >= 0) {
          if (li == 0) {
            ids = fg.global_ids;
            e1 = fg.n_global;
          } else {
            const int k = li - 1;
            const int cx = cx0 + k % wx, cy = cy0 + (k / wx) % wy, cz = cz0 + k / (wx * wy);
            const uint32_t cell = ((uint32_t)cz * (uint32_t)fg.ny + (uint32_t)cy) * (uint32_t)fg.nx + (uint32_t)cx;
            ids = fg.ids;
            e = __ldg(fg.cell_start + cell);
            e1 = __ldg(fg.cell_start + cell + 1);
          }
        }
        for (; e < e1; ++e) test_face(ids ? __ldg(ids + e) : e);
      }

      // collisions.cl:91-128
      V3 res = np;
      float used = time_to_go;
      if (collided) {
        res = sub(hit_p, scale(hit_n, 0.001f));
        const float k = __fadd_rn(1.f, __fdiv_rn(__fmul_rn(c.restitution, hit_depth), __fmul_rn(time_to_go, length(nv))));
        nv = sub(nv, scale(hit_n, __fmul_rn(k, dot(nv, hit_n))));
        used = __fmul_rn(time_to_go, __fdiv_rn(length(sub(res, x)), length(sub(np, x))));
      }
      x = res;
      v = nv;
      time_to_go = __fsub_rn(time_to_go, used);
      acc = mk(0.f, 0.f, 0.f);  // sph.cl:97-99
      ++iters;
    } while (collided && iters < kMaxCollisionIters);

    // sph.cl:103-108
    const V3 v_out = divs(add(v_half_old, v), 2.f);
    pos[i] = make_float4(x.x, x.y, x.z, 0.f);
    vel[i] = make_float4(v_out.x, v_out.y, v_out.z, 0.f);
    ivel[i] = make_float4(v.x, v.y, v.z, owned_mark);
    if (iters_tap) iters_tap[i] = iters;

    lo[0] = fminf(lo[0], x.x); hi[0] = fmaxf(hi[0], x.x);
    lo[1] = fminf(lo[1], x.y); hi[1] = fmaxf(hi[1], x.y);
    lo[2] = fminf(lo[2], x.z); hi[2] = fmaxf(hi[2], x.z);
  }

  // AABB of the new positions for the next sub-step's grid (sph_simulation.cpp:201-217)
  __shared__ float s_lo[8][3], s_hi[8][3];
  const unsigned warp = threadIdx.x >> 5;
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    const float l = warp_min(lo[a]), h = warp_max(hi[a]);
    if (lane_id() == 0) { s_lo[warp][a] = l; s_hi[warp][a] = h; }
  }
  __syncthreads();
  if (threadIdx.x < 3) {
    float l = s_lo[0][threadIdx.x], h = s_hi[0][threadIdx.x];
    for (unsigned w = 1; w < blockDim.x / 32; ++w) { l = fminf(l, s_lo[w][threadIdx.x]); h = fmaxf(h, s_hi[w][threadIdx.x]); }
    atomicMin(&next_bounds->lo[threadIdx.x], float_to_ordered(l));
    atomicMax(&next_bounds->hi[threadIdx.x], float_to_ordered(h));
  }
}

void launch_prepare_faces(const float* normals, const float* vertices, const uint32_t* indices, uint32_t face_count,
                          Face* faces, cudaStream_t stream, uint64_t* launches) {
  if (face_count == 0) return;
  k_prepare_faces<<<(face_count + 127) / 128, 128, 0, stream>>>(normals, vertices, indices, face_count, faces);
  if (launches) ++*launches;
}

void launch_integrate(const StateArrays& s, const float4* accel, const uint32_t* skey, const Face* faces,
                      uint32_t face_count, const FaceGrid& face_grid, const GridState* grid, const SphConst& c,
                      BoundsAcc* next_bounds, uint32_t* iters_tap, uint32_t n_launch, int sm_count, cudaStream_t stream,
                      uint64_t* launches) {
  const unsigned blocks = std::max(1u, std::min<unsigned>((n_launch + 255) / 256, (unsigned)sm_count * 8u));
  if (face_grid.nx > 0 && face_count > 0)
    k_integrate<true><<<blocks, 256, 0, stream>>>(s.pos, s.vel, s.ivel, accel, skey, faces, face_count, face_grid, grid, c,
                                                  next_bounds, iters_tap);
  else
    k_integrate<false><<<blocks, 256, 0, stream>>>(s.pos, s.vel, s.ivel, accel, skey, faces, face_count, face_grid, grid, c,
                                                   next_bounds, iters_tap);
  if (launches) ++*launches;
}

// ---------------------------------------------------------------------------------------------
// Host: face grid construction (double precision, conservative).
// ---------------------------------------------------------------------------------------------
void build_face_grid(const float* vertices, const uint32_t* indices, uint32_t face_count, FaceGrid* out,
                     std::vector<uint32_t>* cell_start, std::vector<uint32_t>* ids, std::vector<uint32_t>* global_ids) {
  FaceGrid g{};
  cell_start->clear();
  ids->clear();
  global_ids->clear();
  struct Tri {
    double lo[3], hi[3], v0[3], n[3], pad;
  };
  std::vector<Tri> tris;
  std::vector<uint32_t> tri_face;
  std::vector<double> extents;
  double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
  for (uint32_t f = 0; f < face_count; ++f) {
    double v[3][3];
    bool finite = true;
    for (int k = 0; k < 3; ++k)
      for (int a = 0; a < 3; ++a) {
        v[k][a] = (double)vertices[3 * (size_t)indices[3 * f + k] + a];
        finite = finite && std::isfinite(v[k][a]) && std::fabs(v[k][a]) < 1e9;
      }
    double u[3], w[3];
    for (int a = 0; a < 3; ++a) { u[a] = v[1][a] - v[0][a]; w[a] = v[2][a] - v[0][a]; }
    const double uu = u[0] * u[0] + u[1] * u[1] + u[2] * u[2], ww = w[0] * w[0] + w[1] * w[1] + w[2] * w[2];
    const double uw = u[0] * w[0] + u[1] * w[1] + u[2] * w[2];
    // conditioning of the barycentric solve (collisions.cl:71-73): sin^2 of the angle between the edges.
    // Below 1e-3 the fp32 inside test could accept points visibly outside the triangle: such faces
    // (and non-finite ones) are tested for every particle, exactly like the reference does.
    const bool regular = finite && uu > 0.0 && ww > 0.0 && (uu * ww - uw * uw) >= 1e-3 * uu * ww;
    if (!regular) {
      global_ids->push_back(f);
      continue;
    }
    Tri t;
    double extent = 0.0;
    for (int a = 0; a < 3; ++a) {
      t.lo[a] = std::min(v[0][a], std::min(v[1][a], v[2][a]));
      t.hi[a] = std::max(v[0][a], std::max(v[1][a], v[2][a]));
      extent = std::max(extent, t.hi[a] - t.lo[a]);
      t.v0[a] = v[0][a];
    }
    t.n[0] = u[1] * w[2] - u[2] * w[1];
    t.n[1] = u[2] * w[0] - u[0] * w[2];
    t.n[2] = u[0] * w[1] - u[1] * w[0];
    // 2 % of the triangle's size + 0.1 mm: >= 100x the distance by which a point accepted by the
    // fp32 inside test can lie outside a triangle this well conditioned
    t.pad = 0.02 * extent + 1e-4;
    for (int a = 0; a < 3; ++a) {
      lo[a] = std::min(lo[a], t.lo[a] - t.pad);
      hi[a] = std::max(hi[a], t.hi[a] + t.pad);
    }
    tris.push_back(t);
    tri_face.push_back(f);
    extents.push_back(extent);
  }
  if (tris.empty()) {  // nothing to grid: nx = 0 selects the full scan
    *out = g;
    return;
  }
  std::vector<double> sorted = extents;
  std::sort(sorted.begin(), sorted.end());
  const double median = sorted[sorted.size() / 2];
  const double span = std::max(hi[0] - lo[0], std::max(hi[1] - lo[1], hi[2] - lo[2]));
  // cell side: about half a typical triangle, at most 64 and at least 4 cells along the longest axis
  double cell = std::min(std::max(0.5 * median, span / 64.0), span / 4.0);
  if (!(cell > 0.0)) cell = 1.0;
  int dims[3];
  for (int a = 0; a < 3; ++a) dims[a] = std::max(1, std::min(64, (int)std::ceil((hi[a] - lo[a]) / cell)));
  const double inv = 1.0 / cell;
  const size_t cells = (size_t)dims[0] * dims[1] * dims[2];
  std::vector<std::vector<uint32_t>> lists(cells);
  const double slack = 0.02;  // in cells: covers the device's fp32 cell arithmetic (it widens by 0.01 itself)
  for (size_t k = 0; k < tris.size(); ++k) {
    const Tri& t = tris[k];
    int c0[3], c1[3];
    for (int a = 0; a < 3; ++a) {
      c0[a] = std::max(0, (int)std::floor((t.lo[a] - t.pad - lo[a]) * inv - slack));
      c1[a] = std::min(dims[a] - 1, (int)std::floor((t.hi[a] + t.pad - lo[a]) * inv + slack));
    }
    const double nabs[3] = {std::fabs(t.n[0]), std::fabs(t.n[1]), std::fabs(t.n[2])};
    const double half = 0.5 * cell + t.pad + slack * cell;
    for (int cz = c0[2]; cz <= c1[2]; ++cz)
      for (int cy = c0[1]; cy <= c1[1]; ++cy)
        for (int cx = c0[0]; cx <= c1[0]; ++cx) {
          // skip cells whose (widened) box lies entirely on one side of the triangle's plane
          const double ctr[3] = {lo[0] + (cx + 0.5) * cell, lo[1] + (cy + 0.5) * cell, lo[2] + (cz + 0.5) * cell};
          const double d = t.n[0] * (ctr[0] - t.v0[0]) + t.n[1] * (ctr[1] - t.v0[1]) + t.n[2] * (ctr[2] - t.v0[2]);
          const double r = (nabs[0] + nabs[1] + nabs[2]) * half;
          if (std::fabs(d) > r) continue;
          lists[((size_t)cz * dims[1] + cy) * dims[0] + cx].push_back(tri_face[k]);  // k ascending => ids ascending
        }
  }
  cell_start->resize(cells + 1);
  for (size_t c = 0; c < cells; ++c) {
    (*cell_start)[c] = (uint32_t)ids->size();
    ids->insert(ids->end(), lists[c].begin(), lists[c].end());
  }
  (*cell_start)[cells] = (uint32_t)ids->size();
  g.ox = (float)lo[0]; g.oy = (float)lo[1]; g.oz = (float)lo[2];
  g.inv = (float)inv;
  g.nx = dims[0]; g.ny = dims[1]; g.nz = dims[2];
  g.n_global = (uint32_t)global_ids->size();
  *out = g;
}

}  // namespace clsph
