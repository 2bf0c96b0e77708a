// dist.cu -- multi-GPU slab decomposition: migration + ghost-layer exchange over NCCL.
//
// The reference is single-device; this is new functionality with the same per-particle
// results. One process per GPU. The fluid is cut into slabs along x by fixed world-space planes,
// snapped each sub-step to the nearest cell boundary of the GLOBAL grid (the AABB is all-reduced,
// so every rank derives the same grid and the same Morton keys). Rank r owns the cells with
// own_lo <= cx < own_hi. Per sub-step, on the library's stream, with no host round trip:
//
//   1. all-reduce (min / max) of the six AABB accumulators                       [2 x 12 bytes]
//   2. k_dist_classify: every previously owned particle gets its new cell; it is copied to the
//      unsorted local array, and additionally
//        - to the left / right message as an EMIGRANT if its cell now belongs to a neighbour
//          (it stays locally as a ghost copy, which is exactly what the neighbour's first ghost
//          layer on this side needs), or
//        - to the left / right message as a GHOST if it lies in this rank's two outermost layers;
//   3. one NCCL group: fixed-size send + recv with each neighbour (counts travel in the header);
//   4. k_dist_unpack appends immigrants (owned) and ghosts to the local array; k_dist_finish
//      publishes the local count;
//   5. the ordinary sub-step runs on owned + ghost particles; roles follow from the key:
//      density for owned and first-ghost-layer cells, forces and integration for owned cells.
//
// Two ghost layers make one exchange per sub-step enough: the first layer's densities are
// recomputed locally from the second. Slabs must be at least four cells thick.
//
// In the sub-cell order (subgrid.cu) the kernels take their roles per particle, and the same protocol runs
// on the planes themselves instead of cells snapped to them: a particle belongs to the rank whose planes
// contain it (owned_here), migrates only when it really crosses one, and the ghosts are the particles within
// 2h of a plane. Which particles a rank advanced is marked by k_integrate in the w lane of the half-step
// velocity. Migrating particles and ghosts carry order keys (cell key and rank in cell of the previous
// sub-step), from which every rank keeps its arrays in the order a single GPU would: results are bitwise
// those of a single-GPU run, and the ranks' downloads merge into the reference's global array.
#include <dlfcn.h>
#include <nccl.h>

#include <cstdio>

#include "dist.cuh"
#include "subview.cuh"

namespace clsph {

// ---------------------------------------------------------------------------------------------
// NCCL through dlopen: the single-GPU library must not depend on libnccl being present.
// ---------------------------------------------------------------------------------------------
namespace {

struct NcclApi {
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  bool ok = false;
};

NcclApi& nccl() {
  static NcclApi api;
  static bool tried = false;
  if (tried) return api;
  tried = true;
  // RTLD_NOLOAD first: inside a torch process the bundled libnccl is already mapped
  api.handle = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
  if (!api.handle) api.handle = dlopen("libnccl.so.2", RTLD_NOW);
  if (!api.handle) api.handle = dlopen("libnccl.so", RTLD_NOW);
  if (!api.handle) return api;
#define LOAD(field, name) *(void**)(&api.field) = dlsym(api.handle, name)
  LOAD(GetUniqueId, "ncclGetUniqueId");
  LOAD(CommInitRank, "ncclCommInitRank");
  LOAD(CommDestroy, "ncclCommDestroy");
  LOAD(AllReduce, "ncclAllReduce");
  LOAD(Send, "ncclSend");
  LOAD(Recv, "ncclRecv");
  LOAD(GroupStart, "ncclGroupStart");
  LOAD(GroupEnd, "ncclGroupEnd");
  LOAD(GetErrorString, "ncclGetErrorString");
#undef LOAD
  api.ok = api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.AllReduce && api.Send && api.Recv &&
           api.GroupStart && api.GroupEnd && api.GetErrorString;
  return api;
}

thread_local char g_nccl_error[256];

bool nccl_check(ncclResult_t r, const char* what) {
  if (r == ncclSuccess) return true;
  snprintf(g_nccl_error, sizeof(g_nccl_error), "%s: %s", what, nccl().GetErrorString ? nccl().GetErrorString(r) : "?");
  return false;
}

// message layout: header | emigrant records (4 x float4) | ghost records (2 x float4)
struct MsgHeader {
  uint32_t n_emigrants, n_ghosts, pad0, pad1;
};
__host__ __device__ inline float4* msg_emigrants(void* msg) { return reinterpret_cast<float4*>(static_cast<char*>(msg) + 16); }
__host__ __device__ inline float4* msg_ghosts(void* msg, uint32_t emax) { return msg_emigrants(msg) + (size_t)emax * 4; }

}  // namespace

const char* dist_last_error() { return g_nccl_error; }

// ---------------------------------------------------------------------------------------------
// Kernels
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_dist_classify(const float4* __restrict__ pos, const float4* __restrict__ vel, const float4* __restrict__ ivel,
                const uint32_t* __restrict__ pid, const uint32_t* __restrict__ skey, const uint32_t* __restrict__ wrank,
                GridState* grid, float4* __restrict__ u_pos, float4* __restrict__ u_vel, float4* __restrict__ u_ivel,
                uint32_t* __restrict__ u_pid, uint32_t* __restrict__ u_ordk, uint32_t* __restrict__ u_ordr,
                uint32_t* __restrict__ u_count, uint32_t capacity, void* send_left, void* send_right, uint32_t emax,
                uint32_t gmax) {
  const GridState g = *grid;
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if ((i & ~31u) >= g.n) return;  // whole warp out of range
  bool owned = i < g.n;
  if (owned && !g.fresh) {
    if (g.sub) {
      // advanced here in the last sub-step (k_integrate's mark); the rest are ghost copies. A slab that is not
      // cut at all (world size 1) has no ghosts and k_integrate writes no marks: everything is owned.
      owned = ivel[i].w == 1.f || !slab_is_cut(g);
    } else {
      const int cx_old = (int)compact10(skey[i]);
      owned = cx_old >= g.prev_lo && cx_old < g.prev_hi;  // the rest are last step's ghost copies: dropped
    }
  }
  float4 p = make_float4(0.f, 0.f, 0.f, 0.f), v = p, iv = p;
  uint32_t id = 0;
  // Order key (sub-cell order only, wrank != null): where the particle stands in the reference's GLOBAL
  // array = (its cell key, its rank inside that cell) of the previous sub-step, compared
  // lexicographically; right after an upload (0, id): ids are the indices of the uploaded global array.
  uint32_t ok_k = 0, ok_r = 0;
  int cx = 0;
  if (owned) {
    p = pos[i]; v = vel[i]; iv = ivel[i]; id = pid[i];
    cx = (int)cell_coord(p.x, g.min_x, g.cell);
    if (wrank) {
      ok_k = g.fresh ? 0u : skey[i];
      ok_r = g.fresh ? id : wrank[i];
    }
  }
  // Established organisation: ownership by cell (planes snapped to cell boundaries). Sub-cell order: by the
  // planes themselves (owned_here in common.cuh), so only particles that really cross a plane migrate.
  const float inf = __int_as_float(0x7f800000);
  const bool has_left = g.sub ? g.plane_lo > -inf : g.own_lo > 0;
  const bool has_right = g.sub ? g.plane_hi < inf : g.own_hi != 0x7fffffff;
  const bool go_left = owned && has_left && (g.sub ? p.x < g.plane_lo : cx < g.own_lo);
  const bool go_right = owned && has_right && (g.sub ? p.x >= g.plane_hi : cx >= g.own_hi);
  const bool stay = owned && !go_left && !go_right;
  if (g.sub) iv.w = 0.f;  // the mark is k_integrate's to set again, for what this rank advances in this sub-step
  // Ghost depth. The neighbour needs, beyond its boundary, the particles within h (candidates of its own
  // particles) whose densities it recomputes, hence those within 2h. The established kernels work on whole
  // cells of side 2h, which makes that two cell layers; the sub-cell order searches by sub-cells of side
  // h and takes roles per particle, so the particles within 2h (1 + 2^-10) of the plane are enough: half
  // the ghost volume.
  const float depth = g.cell * 1.0009765625f;
  const bool ghost_left = stay && has_left && (g.sub ? p.x < g.plane_lo + depth : cx < g.own_lo + 2);
  const bool ghost_right = stay && has_right && (g.sub ? p.x >= g.plane_hi - depth : cx >= g.own_hi - 2);

  // local array: stayers as owned, emigrants as ghost copies (their new key marks them as such)
  const uint32_t at = warp_append(owned, u_count);
  if (owned) {
    if (at < capacity) {
      u_pos[at] = p; u_vel[at] = v; u_ivel[at] = iv; u_pid[at] = id;
      if (u_ordk) { u_ordk[at] = ok_k; u_ordr[at] = ok_r; }
    } else atomicOr(&grid->error, 2u);
  }
  MsgHeader* hl = static_cast<MsgHeader*>(send_left);
  MsgHeader* hr = static_cast<MsgHeader*>(send_right);
  uint32_t e;
  e = warp_append(go_left, &hl->n_emigrants);
  if (go_left) {
    if (e < emax) { float4* r = msg_emigrants(send_left) + (size_t)e * 4; r[0] = p; r[1] = v; r[2] = iv; r[3] = make_float4(__uint_as_float(id), __uint_as_float(ok_k), __uint_as_float(ok_r), 0.f); }
    else atomicOr(&grid->error, 2u);
  }
  e = warp_append(go_right, &hr->n_emigrants);
  if (go_right) {
    if (e < emax) { float4* r = msg_emigrants(send_right) + (size_t)e * 4; r[0] = p; r[1] = v; r[2] = iv; r[3] = make_float4(__uint_as_float(id), __uint_as_float(ok_k), __uint_as_float(ok_r), 0.f); }
    else atomicOr(&grid->error, 2u);
  }
  // ghosts travel as (position, velocity); with order keys (sub-cell order) these ride in the two w lanes,
  // which are free between the integrator and the density pass
  float4 gp = p, gv = v;
  if (wrank) { gp.w = __uint_as_float(ok_k); gv.w = __uint_as_float(ok_r); }
  e = warp_append(ghost_left, &hl->n_ghosts);
  if (ghost_left) {
    if (e < gmax) { float4* r = msg_ghosts(send_left, emax) + (size_t)e * 2; r[0] = gp; r[1] = gv; }
    else atomicOr(&grid->error, 2u);
  }
  e = warp_append(ghost_right, &hr->n_ghosts);
  if (ghost_right) {
    if (e < gmax) { float4* r = msg_ghosts(send_right, emax) + (size_t)e * 2; r[0] = gp; r[1] = gv; }
    else atomicOr(&grid->error, 2u);
  }
}

// Appends one received message: its emigrants become owned particles here, its ghosts are
// candidates for the neighbour passes. Thread t < emax handles emigrant t, the rest ghost t - emax.
__global__ void __launch_bounds__(256)
k_dist_unpack(void* msg, uint32_t emax, uint32_t gmax, GridState* grid, float4* __restrict__ u_pos,
              float4* __restrict__ u_vel, float4* __restrict__ u_ivel, uint32_t* __restrict__ u_pid,
              uint32_t* __restrict__ u_ordk, uint32_t* __restrict__ u_ordr, uint32_t* __restrict__ u_count,
              uint32_t capacity) {
  const MsgHeader h = *static_cast<const MsgHeader*>(msg);
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t ne = min(h.n_emigrants, emax), ng = min(h.n_ghosts, gmax);
  const bool is_e = t < ne;
  const bool is_g = t >= emax && t - emax < ng;
  float4 p = make_float4(0.f, 0.f, 0.f, 0.f), v = p, iv = p;
  uint32_t id = 0xFFFFFFFFu;  // ghosts carry no identity here
  uint32_t ok_k = 0xFFFFFFFFu, ok_r = 0u;  // ... and, without order keys, no place in the reference's order
  if (is_e) {
    const float4* r = msg_emigrants(msg) + (size_t)t * 4;
    p = r[0]; v = r[1]; iv = r[2]; id = __float_as_uint(r[3].x);
    ok_k = __float_as_uint(r[3].y); ok_r = __float_as_uint(r[3].z);
  } else if (is_g) {
    const float4* r = msg_ghosts(msg, emax) + (size_t)(t - emax) * 2;
    p = r[0]; v = r[1];
    if (u_ordk) {  // the sender's order keys, see k_dist_classify
      ok_k = __float_as_uint(p.w); ok_r = __float_as_uint(v.w);
      p.w = 0.f; v.w = 0.f;
    }
  }
  const uint32_t at = warp_append(is_e || is_g, u_count);
  if (is_e || is_g) {
    if (at < capacity) {
      u_pos[at] = p; u_vel[at] = v; u_ivel[at] = iv; u_pid[at] = id;
      if (u_ordk) { u_ordk[at] = ok_k; u_ordr[at] = ok_r; }
    } else atomicOr(&grid->error, 2u);
  }
}

__global__ void k_dist_finish(GridState* grid, const uint32_t* u_count, uint32_t capacity) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    grid->n = min(*u_count, capacity);
    grid->fresh = 0u;
  }
}

// Owned particles only, compacted in local order, as 80-byte AoS records + ids (download).
__global__ void __launch_bounds__(256)
k_dist_export(const float4* __restrict__ pos, const float4* __restrict__ vel, const float4* __restrict__ ivel,
              const float4* __restrict__ aux, const uint32_t* __restrict__ skey, const uint32_t* __restrict__ pid,
              const uint32_t* __restrict__ wrank, const GridState* __restrict__ grid, float4* __restrict__ aos,
              uint32_t* __restrict__ ids, uint32_t* __restrict__ out_count) {
  const GridState g = *grid;
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if ((i & ~31u) >= g.n) return;
  // sub-cell order: owned = advanced here in the last sub-step (its position may since have left the slab)
  const bool owned = i < g.n && (g.fresh || (g.sub ? (ivel[i].w == 1.f || !slab_is_cut(g)) : cell_is_owned(skey[i], g)));
  const uint32_t at = warp_append(owned, out_count);
  if (!owned) return;
  float4 p = pos[i], v = vel[i], iv = ivel[i];
  p.w = 0.f; v.w = 0.f; iv.w = 0.f;
  const float4 a = aux[i];
  float4* rec = aos + (size_t)at * 5;
  rec[0] = p; rec[1] = v; rec[2] = iv;
  rec[3] = make_float4(0.f, 0.f, 0.f, 0.f);
  // padding word: rank inside the cell in the reference's order (0 when not tracked), see clsph_dist_download
  rec[4] = make_float4(a.x, a.y, __uint_as_float(skey[i]), __uint_as_float((wrank && !g.fresh) ? wrank[i] : 0u));
  ids[at] = pid[i];
}

__global__ void k_fill_ids(uint32_t* pid, const uint32_t* src, uint32_t first, uint32_t n) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) pid[i] = src ? src[i] : first + i;
}

// ---------------------------------------------------------------------------------------------
// Host side
// ---------------------------------------------------------------------------------------------
int dist_unique_id(void* out, size_t bytes) {
  if (!nccl().ok) { snprintf(g_nccl_error, sizeof(g_nccl_error), "libnccl.so.2 could not be loaded"); return 1; }
  if (bytes < sizeof(ncclUniqueId)) { snprintf(g_nccl_error, sizeof(g_nccl_error), "unique id needs %zu bytes", sizeof(ncclUniqueId)); return 1; }
  ncclUniqueId id;
  if (!nccl_check(nccl().GetUniqueId(&id), "ncclGetUniqueId")) return 1;
  memcpy(out, &id, sizeof(id));
  return 0;
}

int dist_init(DistState* d, int rank, int world, const void* id_bytes, float plane_lo, float plane_hi, uint32_t emax,
              uint32_t gmax) {
  if (!nccl().ok) { snprintf(g_nccl_error, sizeof(g_nccl_error), "libnccl.so.2 could not be loaded"); return 1; }
  ncclUniqueId id;
  memcpy(&id, id_bytes, sizeof(id));
  ncclComm_t comm;
  if (!nccl_check(nccl().CommInitRank(&comm, world, id, rank), "ncclCommInitRank")) return 1;
  d->comm = comm;
  d->rank = rank;
  d->world = world;
  d->plane_lo = plane_lo;
  d->plane_hi = plane_hi;
  d->emax = emax;
  d->gmax = gmax;
  d->msg_bytes = 16 + (size_t)emax * 64 + (size_t)gmax * 32;
  for (int s = 0; s < 2; ++s) {
    if (cudaMalloc(&d->send[s], d->msg_bytes) != cudaSuccess || cudaMalloc(&d->recv[s], d->msg_bytes) != cudaSuccess) {
      snprintf(g_nccl_error, sizeof(g_nccl_error), "cudaMalloc of %zu-byte message buffers failed", d->msg_bytes);
      return 1;
    }
    cudaMemset(d->send[s], 0, 16);
    cudaMemset(d->recv[s], 0, 16);
  }
  if (cudaMalloc(&d->counters, 64) != cudaSuccess) return 1;
  cudaMemset(d->counters, 0, 64);
  d->active = true;
  return 0;
}

void dist_destroy(DistState* d) {
  if (!d->active) return;
  for (int s = 0; s < 2; ++s) {
    cudaFree(d->send[s]);
    cudaFree(d->recv[s]);
  }
  cudaFree(d->counters);
  if (d->comm && nccl().ok) nccl().CommDestroy(static_cast<ncclComm_t>(d->comm));
  d->active = false;
}

int dist_allreduce_bounds(DistState* d, BoundsAcc* acc, cudaStream_t stream) {
  ncclComm_t comm = static_cast<ncclComm_t>(d->comm);
  if (!nccl_check(nccl().GroupStart(), "ncclGroupStart")) return 1;
  bool ok = nccl_check(nccl().AllReduce(acc->lo, acc->lo, 3, ncclUint32, ncclMin, comm, stream), "ncclAllReduce(min)");
  ok = ok && nccl_check(nccl().AllReduce(acc->hi, acc->hi, 3, ncclUint32, ncclMax, comm, stream), "ncclAllReduce(max)");
  if (!nccl_check(nccl().GroupEnd(), "ncclGroupEnd")) return 1;
  return ok ? 0 : 1;
}

int dist_exchange(DistState* d, const StateArrays& prev, const uint32_t* prev_pid, const uint32_t* skey,
                  const uint32_t* wrank, GridState* grid, const StateArrays& u, uint32_t* u_pid, uint32_t* u_ordk,
                  uint32_t* u_ordr, uint32_t capacity, cudaStream_t stream, uint64_t* launches) {
  ncclComm_t comm = static_cast<ncclComm_t>(d->comm);
  uint32_t* u_count = d->counters;
  cudaMemsetAsync(u_count, 0, sizeof(uint32_t), stream);
  cudaMemsetAsync(d->send[0], 0, 16, stream);
  cudaMemsetAsync(d->send[1], 0, 16, stream);
  const unsigned blocks = (capacity + 255) / 256;
  k_dist_classify<<<blocks, 256, 0, stream>>>(prev.pos, prev.vel, prev.ivel, prev_pid, skey, wrank, grid, u.pos, u.vel, u.ivel,
                                              u_pid, u_ordk, u_ordr, u_count, capacity, d->send[0], d->send[1], d->emax,
                                              d->gmax);
  const bool has_left = d->rank > 0, has_right = d->rank + 1 < d->world;
  if (!nccl_check(nccl().GroupStart(), "ncclGroupStart")) return 1;
  bool ok = true;
  if (has_left) {
    ok = ok && nccl_check(nccl().Send(d->send[0], d->msg_bytes, ncclUint8, d->rank - 1, comm, stream), "ncclSend(left)");
    ok = ok && nccl_check(nccl().Recv(d->recv[0], d->msg_bytes, ncclUint8, d->rank - 1, comm, stream), "ncclRecv(left)");
  }
  if (has_right) {
    ok = ok && nccl_check(nccl().Send(d->send[1], d->msg_bytes, ncclUint8, d->rank + 1, comm, stream), "ncclSend(right)");
    ok = ok && nccl_check(nccl().Recv(d->recv[1], d->msg_bytes, ncclUint8, d->rank + 1, comm, stream), "ncclRecv(right)");
  }
  if (!nccl_check(nccl().GroupEnd(), "ncclGroupEnd") || !ok) return 1;
  const unsigned ublocks = (d->emax + d->gmax + 255) / 256;
  if (has_left)
    k_dist_unpack<<<ublocks, 256, 0, stream>>>(d->recv[0], d->emax, d->gmax, grid, u.pos, u.vel, u.ivel, u_pid, u_ordk, u_ordr,
                                               u_count, capacity);
  if (has_right)
    k_dist_unpack<<<ublocks, 256, 0, stream>>>(d->recv[1], d->emax, d->gmax, grid, u.pos, u.vel, u.ivel, u_pid, u_ordk, u_ordr,
                                               u_count, capacity);
  k_dist_finish<<<1, 32, 0, stream>>>(grid, u_count, capacity);
  if (launches) *launches += 2 + (has_left ? 1 : 0) + (has_right ? 1 : 0);
  return 0;
}

void launch_dist_export(const StateArrays& s, const float4* aux, const uint32_t* skey, const uint32_t* pid,
                        const uint32_t* wrank, const GridState* grid, void* aos, uint32_t* ids, uint32_t* out_count,
                        uint32_t capacity, cudaStream_t stream, uint64_t* launches) {
  cudaMemsetAsync(out_count, 0, sizeof(uint32_t), stream);
  k_dist_export<<<(capacity + 255) / 256, 256, 0, stream>>>(s.pos, s.vel, s.ivel, aux, skey, pid, wrank, grid, (float4*)aos,
                                                            ids, out_count);
  if (launches) ++*launches;
}

void launch_fill_ids(uint32_t* pid, const uint32_t* src, uint32_t first, uint32_t n, cudaStream_t stream,
                     uint64_t* launches) {
  k_fill_ids<<<(n + 255) / 256, 256, 0, stream>>>(pid, src, first, n);
  if (launches) ++*launches;
}

}  // namespace clsph
