// pair_terms.cuh -- per-pair and per-particle arithmetic of the two neighbour passes, shared by
// the kernels of neighbors.cu and subgrid.cu. Replaces forces.cl:33-36, :64-109, smoothing.cl:1-34
// and sph.cl:37-39, :53-58 of the reference.
#pragma once

#include "common.cuh"

namespace clsph {

// Sums of one particle's force pass (forces.cl:50-55).
struct ForceSums {
  float px = 0.f, py = 0.f, pz = 0.f;   // pressure term          forces.cl:70-77
  float wx = 0.f, wy = 0.f, wz = 0.f;   // viscosity term         forces.cl:79-85
  float nx = 0.f, ny = 0.f, nz = 0.f;   // colour-field normal    forces.cl:88-91
  float lap = 0.f;                      // colour-field laplacian forces.cl:93-97
};

// Contribution of neighbour j (known to be inside the support, window == 1) to particle i.
// pj = (x, y, z, p_j/rho_j^2), vj = (vx, vy, vz, m/rho_j); a_i = p_i/rho_i^2.
__device__ __forceinline__ void add_pair(ForceSums& f, const SphConst& c, bool is_self, const float4& pi, const float4& vi,
                                         float a_i, const float4& pj, const float4& vj) {
  const float dx = pi.x - pj.x, dy = pi.y - pj.y, dz = pi.z - pj.z;
  const float s = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
  const float r = sqrtf(s);
  const float mass_over_rho = vj.w;
  if (!is_self) {
    float gx, gy, gz;
    if (r < 0.0000001f) {  // smoothing.cl:23-25 (erratum E3): a scalar broadcast to x, y, z
      gx = gy = gz = c.spiky_degenerate;
    } else {               // smoothing.cl:26-28
      const float hr = c.h - r;
      const float k = c.c_spiky * hr * hr / r;
      gx = k * dx; gy = k * dy; gz = k * dz;
    }
    const float pc = (pj.w + a_i) * c.mass;
    f.px = fmaf(pc, gx, f.px); f.py = fmaf(pc, gy, f.py); f.pz = fmaf(pc, gz, f.pz);
    const float vc = mass_over_rho * (c.c_visc * (c.h - r));  // smoothing.cl:31-34
    f.wx = fmaf(vj.x - vi.x, vc, f.wx); f.wy = fmaf(vj.y - vi.y, vc, f.wy); f.wz = fmaf(vj.z - vi.z, vc, f.wz);
  }
  const float t = c.h2 - r * r;
  const float gc = mass_over_rho * (c.c_poly6_grad * t * t);  // smoothing.cl:6-10
  f.nx = fmaf(gc, dx, f.nx); f.ny = fmaf(gc, dy, f.ny); f.nz = fmaf(gc, dz, f.nz);
  f.lap = fmaf(mass_over_rho, c.c_poly6_lap * t * (3.f * c.h2 - 7.f * r * r), f.lap);  // smoothing.cl:12-17
}

// The same contribution with one MUFU.RSQ in place of the IEEE square root and the IEEE divide
// (about 20 of add_pair's ~80 instructions): r = s rsqrt(s) and 1/r = rsqrt(s) are good to ~2 ulp and
// h^2 - r^2 is taken as h^2 - s, all far inside the 1e-4 bar these sums are held to. The one discrete
// decision, smoothing.cl:23's |r| < 1e-7, is kept exact as s < degenerate_s (smallest s whose rounded
// square root reaches 1e-7; sqrt is monotone). Selected by the option fast_pairs; the default keeps add_pair.
__device__ __forceinline__ void add_pair_fast(ForceSums& f, const SphConst& c, bool is_self, const float4& pi,
                                              const float4& vi, float a_i, const float4& pj, const float4& vj) {
  const float dx = pi.x - pj.x, dy = pi.y - pj.y, dz = pi.z - pj.z;
  const float s = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
  const float inv_r = rsqrtf(s);             // +inf at s = 0 or denormal (self, coincident particles): not used there
  float r = s * inv_r;
  if (s < c.degenerate_s) r = sqrtf(s);      // rare: keep r finite and exact where rsqrt is not usable
  const float mass_over_rho = vj.w;
  if (!is_self) {
    float gx, gy, gz;
    if (s < c.degenerate_s) {
      gx = gy = gz = c.spiky_degenerate;
    } else {
      const float hr = c.h - r;
      const float k = c.c_spiky * hr * hr * inv_r;
      gx = k * dx; gy = k * dy; gz = k * dz;
    }
    const float pc = (pj.w + a_i) * c.mass;
    f.px = fmaf(pc, gx, f.px); f.py = fmaf(pc, gy, f.py); f.pz = fmaf(pc, gz, f.pz);
    const float vc = mass_over_rho * (c.c_visc * (c.h - r));
    f.wx = fmaf(vj.x - vi.x, vc, f.wx); f.wy = fmaf(vj.y - vi.y, vc, f.wy); f.wz = fmaf(vj.z - vi.z, vc, f.wz);
  }
  const float t = c.h2 - s;
  const float gc = mass_over_rho * (c.c_poly6_grad * t * t);
  f.nx = fmaf(gc, dx, f.nx); f.ny = fmaf(gc, dy, f.ny); f.nz = fmaf(gc, dz, f.nz);
  f.lap = fmaf(mass_over_rho, c.c_poly6_lap * t * (3.f * c.h2 - 7.f * s), f.lap);
}
template <bool kFast>
__device__ __forceinline__ void add_pair_sel(ForceSums& f, const SphConst& c, bool is_self, const float4& pi,
                                             const float4& vi, float a_i, const float4& pj, const float4& vj) {
  if (kFast) add_pair_fast(f, c, is_self, pi, vi, a_i, pj, vj);
  else add_pair(f, c, is_self, pi, vi, a_i, pj, vj);
}

// forces.cl:103-109 and sph.cl:53-58: F = -rho P + mu V (+ surface tension), a = F / rho + g.
__device__ __forceinline__ float4 finish_force(const ForceSums& f, const SphConst& c, float rho) {
  float fx = -rho * f.px + f.wx * c.mu, fy = -rho * f.py + f.wy * c.mu, fz = -rho * f.pz + f.wz * c.mu;
  const float nlen = sqrtf(fmaf(f.nz, f.nz, fmaf(f.ny, f.ny, f.nx * f.nx)));
  if (nlen > c.tension_threshold) {
    const float k = -c.sigma * f.lap / nlen;
    fx = fmaf(k, f.nx, fx); fy = fmaf(k, f.ny, fy); fz = fmaf(k, f.nz, fz);
  }
  return make_float4(fx / rho + c.gx, fy / rho + c.gy, fz / rho + c.gz, 0.f);
}

// ---- factored pair terms (k_forces_lists_factored, neighbors.cu) --------------------------------------------------
// Same formulas with the per-run constants factored out of the sums (they are applied once per particle in
// tile_finish_force) and one MUFU.RSQ for r and 1/r: 36 floating-point instructions per pair, no branch.
//   P += (p_j/rho_j^2 + p_i/rho_i^2) (h - r)^2 / r * d        pressure      x m c_spiky
//   W += (v_j - v_i) (m/rho_j) (h - r)                        viscosity     x c_visc
//   N += (m/rho_j) (h^2 - s)^2 * d                            colour normal x c_poly6_grad
//   L += (m/rho_j) (h^2 - s) (3 h^2 - 7 s)                    colour laplacian x c_poly6_lap
// Valid for j != i and s >= degenerate_s; the callers exclude the particle itself (its only contribution, to L,
// is tile_self_lap) and hand particles with a degenerate pair (coincident particles, smoothing.cl:23) to the
// exact add_pair. Both kernels evaluate operands and sums with THESE functions in the same order, so a particle
// gets the same bits whichever kernel serves it.
struct TilePair {
  float kp, dx, dy, dz, vc, ux, uy, uz, gc, lw;
};
__device__ __forceinline__ float rsqrt_fast(float s) {
#ifdef CLSPH_EMU
  return 1.0f / sqrtf(s);
#else
  float r;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(s));
  return r;
#endif
}
// is_self (lists that contain the particle itself): 1/r is taken as 0, which leaves exactly the particle's own term --
// d = 0 and v_j - v_i = 0 wipe the three vector sums, (h^2 - 0)(3 h^2 - 0) m/rho_i is tile_self_lap.
__device__ __forceinline__ TilePair tile_pair_ops(const SphConst& c, const float4& pi, const float4& vi, const float4& pj,
                                                  const float4& vj, float& s_out, bool is_self = false) {
  TilePair o;
  o.dx = pi.x - pj.x; o.dy = pi.y - pj.y; o.dz = pi.z - pj.z;
  const float s = fmaf(o.dz, o.dz, fmaf(o.dy, o.dy, o.dx * o.dx));
  s_out = is_self ? c.h2 : s;  // (never "degenerate")
  const float inv_r = is_self ? 0.f : rsqrt_fast(s);
  const float hr = c.h - s * inv_r;
  o.kp = (pj.w + pi.w) * (hr * hr * inv_r);
  o.vc = vj.w * hr;
  o.ux = vj.x - vi.x; o.uy = vj.y - vi.y; o.uz = vj.z - vi.z;
  const float t = c.h2 - s;
  o.gc = vj.w * (t * t);
  o.lw = vj.w * (t * fmaf(-7.f, s, 3.f * c.h2));
  return o;
}
__device__ __forceinline__ void tile_pair_add(ForceSums& f, const TilePair& o) {
  f.px = fmaf(o.kp, o.dx, f.px); f.py = fmaf(o.kp, o.dy, f.py); f.pz = fmaf(o.kp, o.dz, f.pz);
  f.wx = fmaf(o.ux, o.vc, f.wx); f.wy = fmaf(o.uy, o.vc, f.wy); f.wz = fmaf(o.uz, o.vc, f.wz);
  f.nx = fmaf(o.gc, o.dx, f.nx); f.ny = fmaf(o.gc, o.dy, f.ny); f.nz = fmaf(o.gc, o.dz, f.nz);
  f.lap += o.lw;
}
// The particle's own term (s = 0: only the laplacian of the colour field is non-zero), then the constants, then
// finish_force. mor_i = m / rho_i.
__device__ __forceinline__ float4 tile_finish_force(ForceSums f, const SphConst& c, float rho, float mor_i, bool self_listed = false) {
  if (!self_listed) f.lap += mor_i * (c.h2 * (3.f * c.h2));
  const float kp = c.mass * c.c_spiky;
  f.px *= kp; f.py *= kp; f.pz *= kp;
  f.wx *= c.c_visc; f.wy *= c.c_visc; f.wz *= c.c_visc;
  f.nx *= c.c_poly6_grad; f.ny *= c.c_poly6_grad; f.nz *= c.c_poly6_grad;
  f.lap *= c.c_poly6_lap;
  return finish_force(f, c, rho);
}

// forces.cl:33-36 / smoothing.cl:1-4: rho = sum m C6 (h^2 - r^2)^3; sph.cl:37-39: Tait pressure.
// Writes (rho, p) to aux[i] and the two per-neighbour factors the force pass gathers.
__device__ __forceinline__ void finish_density(const SphConst& c, float sum_cubed, uint32_t i, float4* __restrict__ aux,
                                               float4* pos, float4* vel) {
  const float rho = c.mass * c.c_poly6 * sum_cubed;
  const float q = rho / c.rho0;
  const float q2 = q * q, q4 = q2 * q2;
  const float prs = c.K * (q4 * q2 * q - 1.f);
  aux[i] = make_float4(rho, prs, 0.f, 0.f);
  reinterpret_cast<float*>(pos + i)[3] = prs / (rho * rho);  // only the w lane: other warps may be reading x, y, z
  reinterpret_cast<float*>(vel + i)[3] = c.mass / rho;
}

}  // namespace clsph
