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

__global__ void __launch_bounds__(256)
k_integrate(float4* __restrict__ pos, float4* __restrict__ vel, float4* __restrict__ ivel,
            const float4* __restrict__ accel, const uint32_t* __restrict__ skey, const Face* __restrict__ faces,
            uint32_t face_count, const GridState* __restrict__ grid, const SphConst c, BoundsAcc* next_bounds,
            uint32_t* __restrict__ iters_tap) {
  const GridState g = *grid;
  const uint32_t n = g.n;
  const bool sliced = g.own_lo > 0 || g.own_hi != 0x7fffffff;  // multi-GPU: ghosts are not advanced
  float lo[3] = {2147483648.f, 2147483648.f, 2147483648.f};
  float hi[3] = {-2147483648.f, -2147483648.f, -2147483648.f};

  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    if (sliced && !cell_is_owned(skey[i], g)) continue;
    const float4 p4 = pos[i], iv4 = ivel[i], a4 = accel[i];
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
      for (uint32_t f = 0; f < face_count; ++f) {
        const float4* fr = reinterpret_cast<const float4*>(faces + f);
        const float4 f0 = __ldg(fr + 0), f1 = __ldg(fr + 1);
        const V3 n0 = mk(f0.x, f0.y, f0.z);
        const V3 a = mk(f1.x, f1.y, f1.z);
        const float nd = dot(n0, travel);
        if (nd == 0.f) continue;  // :54-56 (the oriented denominator is +-nd)
        const float na = dot(n0, sub(a, x));
        // The plane parameter is r = (+-na) / (+-nd) = na / nd whatever the orientation, and a hit
        // needs 0 <= r <= 1 (:58-60). Two division-free rejections that can never drop such a face:
        // opposite signs with a quotient that cannot underflow to -0, and |na| beyond |nd| by more
        // than the half ulp that could still round the quotient down to 1. Nearly every face of a
        // scene leaves here: a sub-step moves a particle by millimetres.
        if ((na < 0.f) != (nd < 0.f) && fabsf(na) > 1e-30f * fabsf(nd)) continue;
        if (fabsf(na) > fabsf(nd) * 1.0000002f) continue;
        // :27-29 orient the normal along the travel direction
        const bool flip = __fdiv_rn(nd, __fmul_rn(f0.w, travel_len)) <= 0.f;
        const float denom = flip ? -nd : nd;  // dot(-n, d) == -dot(n, d) exactly
        const float r = __fdiv_rn(flip ? -na : na, denom);  // :58
        if (!(0.f <= r && r <= 1.f)) continue;
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
          if (collided && dist > hit_dist) continue;  // :77-80
          hit_n = flip ? neg(n0) : n0;
          hit_p = hit;
          hit_depth = length(sub(np, hit));
          hit_dist = dist;
          collided = true;
        }
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
    ivel[i] = make_float4(v.x, v.y, v.z, 0.f);
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
                      uint32_t face_count, const GridState* grid, const SphConst& c, BoundsAcc* next_bounds,
                      uint32_t* iters_tap, uint32_t n_launch, int sm_count, cudaStream_t stream, uint64_t* launches) {
  const unsigned blocks = std::min<unsigned>((n_launch + 255) / 256, (unsigned)sm_count * 8u);
  k_integrate<<<std::max(1u, blocks), 256, 0, stream>>>(s.pos, s.vel, s.ivel, accel, skey, faces, face_count, grid, c,
                                                        next_bounds, iters_tap);
  if (launches) ++*launches;
}

}  // namespace clsph
