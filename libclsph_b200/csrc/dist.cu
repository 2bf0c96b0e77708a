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
//
// Transport. On a box where the GPUs reach each other's memory (NVLink / NVSwitch, CUDA IPC between the rank
// processes) steps 1 and 3 do not call NCCL at all: every rank owns a MAILBOX in its HBM that the others map.
//   * k_dist_classify stores the emigrant and ghost records straight into the neighbour's mailbox while it
//     compacts the local array -- the transfer is the kernel's own stores, overlapping its other work, and only the
//     records that exist travel; its last CTA then publishes the counts and the sub-step's sequence number
//     (release at system scope), k_dist_unpack waits for that number (acquire) before it reads;
//   * the AABB travels the same way: k_bounds_publish stores the six accumulators of a rank into a slot of every
//     mailbox right after the integrator has produced them (the end of the PREVIOUS sub-step, so the stores are long
//     done when they are needed), k_bounds_gather waits for the slots of all ranks and reduces them.
// No rank ever waits for a rank that waits for it: a kernel that stores never waits, a kernel that waits only needs
// kernels that precede it in the other ranks' streams. Slots are double-buffered by the parity of the sequence
// number (a rank two slabs away may run ahead by part of a sub-step); the record areas need no second buffer
// because a neighbour can only store sub-step k's records after this rank has published its AABB for k, i.e. after
// it has unpacked k - 1. Waits give up after ~30 s and flag CLSPH_ECOMM instead of hanging the GPU.
// NCCL remains for the start-up (exchange of the IPC handles) and as the transport when peer access is unavailable
// (CLSPH_DIST_TRANSPORT=nccl forces it).
#include <dlfcn.h>
#include <nccl.h>

#include <cstdio>
#include <cstdlib>
#include <vector>

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

// Where k_dist_classify puts what goes to one neighbour.
struct MsgOut {
  uint32_t* counts;   // [0] emigrants, [1] ghosts appended so far (local memory)
  float4* emigrants;  // records: the local send buffer (NCCL) or the neighbour's mailbox (peer stores)
  float4* ghosts;
};

// Peer transport: the last CTA of k_dist_classify to finish publishes the counts and the sequence number.
struct PeerSignal {
  uint32_t* done;          // ticket counter (null: NCCL transport, nothing to publish)
  uint32_t* header_left;   // header of my message in the left / right neighbour's mailbox (null: no such neighbour)
  uint32_t* header_right;
  uint32_t seq;
};
// One received message for k_dist_unpack.
struct MsgIn {
  const uint32_t* header;  // emigrants, ghosts, sequence number (null: no neighbour on this side)
  const float4* emigrants;
  const float4* ghosts;
};

// Mailbox of a rank (peer transport), mapped by every other rank:
//   [0, 2048)     AABB slots [parity][source rank][8 words]: lo[3], hi[3], sequence number, pad
//   [2048, 2112)  message headers [side][8 words]: emigrants, ghosts, sequence number; side 0 = from the left neighbour
//   [4096, ...)   records of side 0, then of side 1: emigrants (emax x 64 B), ghosts (gmax x 32 B)
constexpr int kMaxPeers = 32;
constexpr size_t kMailboxRecords = 4096;
__host__ __device__ inline uint32_t* mailbox_bounds(void* box, uint32_t parity, uint32_t rank) {
  return static_cast<uint32_t*>(box) + ((size_t)parity * kMaxPeers + rank) * 8u;
}
__host__ __device__ inline const uint32_t* mailbox_bounds(const void* box, uint32_t parity, uint32_t rank) {
  return static_cast<const uint32_t*>(box) + ((size_t)parity * kMaxPeers + rank) * 8u;
}
__host__ __device__ inline uint32_t* mailbox_header(void* box, int side) { return static_cast<uint32_t*>(box) + 512 + side * 8; }
inline size_t side_bytes(uint32_t emax, uint32_t gmax) { return (size_t)emax * 64 + (size_t)gmax * 32; }
inline float4* mailbox_emigrants(void* box, int side, uint32_t emax, uint32_t gmax) {
  return reinterpret_cast<float4*>(static_cast<char*>(box) + kMailboxRecords + (size_t)side * side_bytes(emax, gmax));
}
inline float4* mailbox_ghosts(void* box, int side, uint32_t emax, uint32_t gmax) { return mailbox_emigrants(box, side, emax, gmax) + (size_t)emax * 4; }

}  // namespace

const char* dist_last_error() { return g_nccl_error; }

// ---------------------------------------------------------------------------------------------
// Kernels
// ---------------------------------------------------------------------------------------------
// ---- peer transport: flags and waits ---------------------------------------------------------------------
__device__ __forceinline__ uint32_t load_flag(const uint32_t* p) { return *reinterpret_cast<const volatile uint32_t*>(p); }
__device__ __forceinline__ void store_flag(uint32_t* p, uint32_t v) { *reinterpret_cast<volatile uint32_t*>(p) = v; }
// Waits until *flag == want. Gives up after kSpinLimit clock ticks (a peer that died must not hang this GPU).
constexpr long long kSpinLimit = 60000000000ll;  // ~30 s of GPU clock: a rank may be busy on its host between sub-steps (tests compare on rank 0)
__device__ __forceinline__ bool wait_flag(const uint32_t* flag, uint32_t want, long long limit = kSpinLimit) {
  const long long t0 = clock64();
  while (load_flag(flag) != want)
    if (clock64() - t0 > limit) return false;
  __threadfence_system();  // acquire: what the peer stored before the flag is visible from here on
  return true;
}

__global__ void __launch_bounds__(256)
k_dist_classify(const float4* __restrict__ pos, const float4* __restrict__ vel, const float4* __restrict__ ivel,
                const uint32_t* __restrict__ pid, const uint32_t* __restrict__ skey, const uint32_t* __restrict__ wrank,
                GridState* grid, float4* __restrict__ u_pos, float4* __restrict__ u_vel, float4* __restrict__ u_ivel,
                uint32_t* __restrict__ u_pid, uint32_t* __restrict__ u_ordk, uint32_t* __restrict__ u_ordr,
                uint32_t* __restrict__ u_count, uint32_t capacity, const MsgOut left, const MsgOut right, uint32_t emax,
                uint32_t gmax, const PeerSignal sig) {
  const GridState g = *grid;
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  bool stored_remotely = false;
  if (blockIdx.x * blockDim.x < g.n) {  // (a CTA wholly out of range only takes part in the signalling below)
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
  const uint32_t at = block256_append(owned, u_count);
  if (owned) {
    if (at < capacity) {
      u_pos[at] = p; u_vel[at] = v; u_ivel[at] = iv; u_pid[at] = id;
      if (u_ordk) { u_ordk[at] = ok_k; u_ordr[at] = ok_r; }
    } else atomicOr(&grid->error, 2u);
  }
  // the records go where the transport wants them: the local send buffer (NCCL) or the neighbour's mailbox (peer
  // stores over NVLink); the counters are local either way
  uint32_t e;
  e = warp_append(go_left, left.counts);
  if (go_left) {
    if (e < emax) { float4* r = left.emigrants + (size_t)e * 4; r[0] = p; r[1] = v; r[2] = iv; r[3] = make_float4(__uint_as_float(id), __uint_as_float(ok_k), __uint_as_float(ok_r), 0.f); }
    else atomicOr(&grid->error, 2u);
  }
  e = warp_append(go_right, right.counts);
  if (go_right) {
    if (e < emax) { float4* r = right.emigrants + (size_t)e * 4; r[0] = p; r[1] = v; r[2] = iv; r[3] = make_float4(__uint_as_float(id), __uint_as_float(ok_k), __uint_as_float(ok_r), 0.f); }
    else atomicOr(&grid->error, 2u);
  }
  // ghosts travel as (position, velocity); with order keys (sub-cell order) these ride in the two w lanes,
  // which are free between the integrator and the density pass
  float4 gp = p, gv = v;
  if (wrank) { gp.w = __uint_as_float(ok_k); gv.w = __uint_as_float(ok_r); }
  e = warp_append(ghost_left, left.counts + 1);
  if (ghost_left) {
    if (e < gmax) { float4* r = left.ghosts + (size_t)e * 2; r[0] = gp; r[1] = gv; }
    else atomicOr(&grid->error, 2u);
  }
  e = warp_append(ghost_right, right.counts + 1);
  if (ghost_right) {
    if (e < gmax) { float4* r = right.ghosts + (size_t)e * 2; r[0] = gp; r[1] = gv; }
    else atomicOr(&grid->error, 2u);
  }
  stored_remotely = go_left || go_right || ghost_left || ghost_right;
  }
  if (!sig.done) return;
  // Peer transport: the records are in the neighbours' mailboxes once every CTA has passed this point; the last
  // one to arrive publishes the counts, then the sequence number the receivers wait for (release at system scope).
  if (stored_remotely) __threadfence_system();  // the few threads that stored into a mailbox
  __syncthreads();
  if (threadIdx.x != 0) return;
  __threadfence();  // device scope is enough here (7500 CTAs: a system-scope fence in each cost ~25 us); the last CTA's is not
  if (atomicAdd(sig.done, 1u) != gridDim.x - 1u) return;
  __threadfence_system();
  if (sig.header_left) { store_flag(sig.header_left, atomicAdd(left.counts, 0u)); store_flag(sig.header_left + 1, atomicAdd(left.counts + 1, 0u)); }
  if (sig.header_right) { store_flag(sig.header_right, atomicAdd(right.counts, 0u)); store_flag(sig.header_right + 1, atomicAdd(right.counts + 1, 0u)); }
  __threadfence_system();
  if (sig.header_left) store_flag(sig.header_left + 2, sig.seq);
  if (sig.header_right) store_flag(sig.header_right + 2, sig.seq);
}

// ---------------------------------------------------------------------------------------------
// The same decisions WITHOUT copying the local array (sub-cell order only). The particles this rank advanced
// stay where they are; what the sort is given is a list of their indices (`live`, then the indices of the
// records k_dist_unpack appends behind the old array), so last sub-step's ghost copies simply are not in it.
// Per local particle this pass reads a position and a mark and writes three words, instead of moving 120 bytes:
// the copy was 0.06 of the 0.10 ms an exchange cost at 8 GPUs. The order keys of the previous sub-step are
// written in place (ordk / ordr at the particle's own index) for the gather pass to read.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_dist_select(const float4* __restrict__ pos, const float4* __restrict__ vel, const float4* __restrict__ ivel,
              const uint32_t* __restrict__ pid, const uint32_t* __restrict__ skey, const uint32_t* __restrict__ wrank,
              GridState* grid, uint32_t* __restrict__ ordk, uint32_t* __restrict__ ordr, uint32_t* __restrict__ live,
              uint32_t* __restrict__ live_count, uint32_t capacity, const MsgOut left, const MsgOut right, uint32_t emax,
              uint32_t gmax, const PeerSignal sig) {
  const GridState g = *grid;
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  bool stored_remotely = false;
  if (blockIdx.x * blockDim.x < g.n) {  // (a CTA wholly out of range only takes part in the signalling below)
    bool owned = i < g.n;
    // advanced here in the last sub-step (k_integrate's mark); the rest are ghost copies. A slab that is not cut at
    // all (world size 1) has no ghosts and k_integrate writes no marks: everything is owned.
    if (owned && !g.fresh) owned = ivel[i].w == 1.f || !slab_is_cut(g);
    float4 p = make_float4(0.f, 0.f, 0.f, 0.f);
    uint32_t id = 0, ok_k = 0, ok_r = 0;
    if (owned) {
      p = pos[i];
      id = pid[i];
      ok_k = g.fresh ? 0u : skey[i];
      ok_r = g.fresh ? id : wrank[i];
      ordk[i] = ok_k;
      ordr[i] = ok_r;
    }
    const float inf = __int_as_float(0x7f800000);
    const bool has_left = g.plane_lo > -inf, has_right = g.plane_hi < inf;
    const bool go_left = owned && has_left && p.x < g.plane_lo;
    const bool go_right = owned && has_right && p.x >= g.plane_hi;
    const bool stay = owned && !go_left && !go_right;
    const float depth = g.cell * 1.0009765625f;  // ghost depth 2h (1 + 2^-10), see k_dist_classify
    const bool ghost_left = stay && has_left && p.x < g.plane_lo + depth;
    const bool ghost_right = stay && has_right && p.x >= g.plane_hi - depth;
    const uint32_t at = block256_append(owned, live_count);
    if (owned) {
      if (at < capacity) live[at] = i;
      else atomicOr(&grid->error, 2u);
    }
    stored_remotely = go_left || go_right || ghost_left || ghost_right;
    if (__any_sync(kFullMask, stored_remotely)) {  // few warps touch a plane: only they read the rest of the record
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f), iv = v;
      if (stored_remotely) { v = vel[i]; iv = ivel[i]; iv.w = 0.f; }
      const float4 tag = make_float4(__uint_as_float(id), __uint_as_float(ok_k), __uint_as_float(ok_r), 0.f);
      uint32_t e = warp_append(go_left, left.counts);
      if (go_left) {
        if (e < emax) { float4* r = left.emigrants + (size_t)e * 4; r[0] = p; r[1] = v; r[2] = iv; r[3] = tag; }
        else atomicOr(&grid->error, 2u);
      }
      e = warp_append(go_right, right.counts);
      if (go_right) {
        if (e < emax) { float4* r = right.emigrants + (size_t)e * 4; r[0] = p; r[1] = v; r[2] = iv; r[3] = tag; }
        else atomicOr(&grid->error, 2u);
      }
      float4 gp = p, gv = v;  // ghosts travel as (position, velocity) with the order keys in the two w lanes
      gp.w = __uint_as_float(ok_k); gv.w = __uint_as_float(ok_r);
      e = warp_append(ghost_left, left.counts + 1);
      if (ghost_left) {
        if (e < gmax) { float4* r = left.ghosts + (size_t)e * 2; r[0] = gp; r[1] = gv; }
        else atomicOr(&grid->error, 2u);
      }
      e = warp_append(ghost_right, right.counts + 1);
      if (ghost_right) {
        if (e < gmax) { float4* r = right.ghosts + (size_t)e * 2; r[0] = gp; r[1] = gv; }
        else atomicOr(&grid->error, 2u);
      }
    }
  }
  if (!sig.done) return;
  if (stored_remotely) __threadfence_system();  // the few threads that stored into a mailbox
  __syncthreads();
  if (threadIdx.x != 0) return;
  __threadfence();  // device scope is enough here (7500 CTAs: a system-scope fence in each cost ~25 us); the last CTA's is not
  if (atomicAdd(sig.done, 1u) != gridDim.x - 1u) return;
  __threadfence_system();
  if (sig.header_left) { store_flag(sig.header_left, atomicAdd(left.counts, 0u)); store_flag(sig.header_left + 1, atomicAdd(left.counts + 1, 0u)); }
  if (sig.header_right) { store_flag(sig.header_right, atomicAdd(right.counts, 0u)); store_flag(sig.header_right + 1, atomicAdd(right.counts + 1, 0u)); }
  __threadfence_system();
  if (sig.header_left) store_flag(sig.header_left + 2, sig.seq);
  if (sig.header_right) store_flag(sig.header_right + 2, sig.seq);
}

// The six AABB accumulators of this rank into slot [parity of seq][rank] of every rank's mailbox (its own too).
__global__ void k_bounds_publish(const BoundsAcc* acc, void* const* mailboxes, int rank, int world, uint32_t seq) {
  const int r = (int)threadIdx.x;
  if (blockIdx.x != 0 || r >= world) return;
  uint32_t* slot = mailbox_bounds(mailboxes[r], seq & 1u, (uint32_t)rank);
  for (int a = 0; a < 3; ++a) { store_flag(slot + a, acc->lo[a]); store_flag(slot + 3 + a, acc->hi[a]); }
  __threadfence_system();
  store_flag(slot + 6, seq);
}

// Teardown: tells every rank that this one has finished storing (its stream is drained), then waits until all
// others have said so too: only then may a mailbox be freed. A few seconds at most, in case a peer is gone.
constexpr uint32_t kClosed = 0xC105ED00u;
__global__ void k_dist_close(void* const* mailboxes, void* mailbox, int rank, int world) {
  const int r = (int)threadIdx.x;
  if (blockIdx.x != 0 || r >= world) return;
  __threadfence_system();
  store_flag(mailbox_bounds(mailboxes[r], 0u, (uint32_t)rank) + 7, kClosed);
  wait_flag(mailbox_bounds(mailbox, 0u, (uint32_t)r) + 7, kClosed, kSpinLimit / 4);
}

// Waits for the slots of all ranks and leaves the global AABB in acc (what the all-reduce of the NCCL transport does).
__global__ void k_bounds_gather(BoundsAcc* acc, void* mailbox, int world, uint32_t seq, GridState* grid) {
  const unsigned r = threadIdx.x;  // one warp
  uint32_t lo[3] = {0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu}, hi[3] = {0u, 0u, 0u};
  if ((int)r < world) {
    const uint32_t* slot = mailbox_bounds(mailbox, seq & 1u, r);
    if (!wait_flag(slot + 6, seq)) atomicOr(&grid->error, 8u);
    for (int a = 0; a < 3; ++a) { lo[a] = load_flag(slot + a); hi[a] = load_flag(slot + 3 + a); }
  }
  for (int a = 0; a < 3; ++a)
    for (int o = 16; o > 0; o >>= 1) {
      lo[a] = min(lo[a], __shfl_xor_sync(kFullMask, lo[a], o));
      hi[a] = max(hi[a], __shfl_xor_sync(kFullMask, hi[a], o));
    }
  if (r == 0)
    for (int a = 0; a < 3; ++a) { acc->lo[a] = lo[a]; acc->hi[a] = hi[a]; }
}

// Appends the received messages: their emigrants become owned particles here, their ghosts are candidates for
// the neighbour passes. The first half of the grid serves the message from the left, the second half the one from
// the right; in a message, thread t < emax handles emigrant t, the rest ghost t - emax.
// wait_seq != 0 (peer transport): a message is complete when header[2] == wait_seq.
// The last CTA to finish publishes the local particle count (what k_dist_finish does otherwise) and re-arms the
// counters of the exchange for the next sub-step.
__global__ void __launch_bounds__(256)
k_dist_unpack(const MsgIn from_left, const MsgIn from_right, uint32_t wait_seq, uint32_t emax, uint32_t gmax,
              GridState* grid, float4* __restrict__ u_pos,
              float4* __restrict__ u_vel, float4* __restrict__ u_ivel, uint32_t* __restrict__ u_pid,
              uint32_t* __restrict__ u_ordk, uint32_t* __restrict__ u_ordr, uint32_t* counters,
              uint32_t capacity, uint32_t* __restrict__ live) {
  __shared__ uint32_t s_counts[2];
  uint32_t* u_count = counters;  // entries of the local array (copying exchange) or of the index list `live` (in place)
  const unsigned half = gridDim.x >> 1;
  const bool right = blockIdx.x >= half;
  const MsgIn msg = right ? from_right : from_left;
  if (threadIdx.x == 0) {
    bool ok = msg.header != nullptr;
    if (ok && wait_seq) {
      ok = wait_flag(msg.header + 2, wait_seq);
      if (!ok) atomicOr(&grid->error, 8u);
    }
    s_counts[0] = ok ? load_flag(msg.header) : 0u;
    s_counts[1] = ok ? load_flag(msg.header + 1) : 0u;
  }
  __syncthreads();
  const uint32_t t = (blockIdx.x - (right ? half : 0u)) * blockDim.x + threadIdx.x;
  const uint32_t ne = min(s_counts[0], emax), ng = min(s_counts[1], gmax);
  const bool is_e = t < ne;
  const bool is_g = t >= emax && t - emax < ng;
  float4 p = make_float4(0.f, 0.f, 0.f, 0.f), v = p, iv = p;
  uint32_t id = 0xFFFFFFFFu;  // ghosts carry no identity here
  uint32_t ok_k = 0xFFFFFFFFu, ok_r = 0u;  // ... and, without order keys, no place in the reference's order
  if (is_e) {
    const float4* r = msg.emigrants + (size_t)t * 4;  // (L2 loads: the lines were written from outside this SM's L1)
    p = __ldcg(r); v = __ldcg(r + 1); iv = __ldcg(r + 2);
    const float4 tag = __ldcg(r + 3);
    id = __float_as_uint(tag.x);
    ok_k = __float_as_uint(tag.y); ok_r = __float_as_uint(tag.z);
  } else if (is_g) {
    const float4* r = msg.ghosts + (size_t)(t - emax) * 2;
    p = __ldcg(r); v = __ldcg(r + 1);
    if (u_ordk) {  // the sender's order keys, see k_dist_classify
      ok_k = __float_as_uint(p.w); ok_r = __float_as_uint(v.w);
      p.w = 0.f; v.w = 0.f;
    }
  }
  const bool rec = is_e || is_g;
  uint32_t at = warp_append(rec, u_count);  // place in the local array (copying exchange) or in the index list (in place)
  bool fits = at < capacity;
  if (live) {
    // in place: the records go behind the previous array (grid->n is still its length), the list gets their indices
    const uint32_t slot = grid->n + warp_append(rec, counters + 8);
    fits = fits && slot < capacity;
    if (rec && fits) live[at] = slot;
    at = slot;
  }
  if (rec) {
    if (fits) {
      u_pos[at] = p; u_vel[at] = v; u_ivel[at] = iv; u_pid[at] = id;
      if (u_ordk) { u_ordk[at] = ok_k; u_ordr[at] = ok_r; }
    } else atomicOr(&grid->error, 2u);
  }
  // the last CTA: local particle count, counters back to zero (counters[1], the export count, is not ours)
  __threadfence();
  __syncthreads();
  if (threadIdx.x != 0) return;
  if (atomicAdd(counters + 3, 1u) != gridDim.x - 1u) return;
  __threadfence();
  grid->n = min(atomicAdd(u_count, 0u), capacity);
  grid->fresh = 0u;
  counters[0] = 0u; counters[2] = 0u; counters[3] = 0u;
  counters[4] = 0u; counters[5] = 0u; counters[6] = 0u; counters[7] = 0u; counters[8] = 0u;
}

__global__ void k_dist_finish(GridState* grid, uint32_t* counters, uint32_t capacity) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    grid->n = min(counters[0], capacity);
    grid->fresh = 0u;
    counters[0] = 0u; counters[2] = 0u; counters[3] = 0u;
    counters[4] = 0u; counters[5] = 0u; counters[6] = 0u; counters[7] = 0u; counters[8] = 0u;
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

// CLSPH_DIST_TIMING=1 (diagnostics): CUDA events around the pieces of the exchange of up to 256 sub-steps,
// averaged and printed to stderr when the context is destroyed.
namespace {
struct ExchangeTiming {
  bool on = false;
  std::vector<cudaEvent_t> ev;  // 5 per sub-step: start, after classify, after signal, after unpack, end
  size_t used = 0;
};
ExchangeTiming g_timing;
void timing_mark(cudaStream_t stream) {
  if (!g_timing.on || g_timing.used >= g_timing.ev.size()) return;
  cudaEventRecord(g_timing.ev[g_timing.used++], stream);
}
void timing_report(int rank) {
  if (!g_timing.on) return;
  cudaDeviceSynchronize();
  double sum[4] = {0, 0, 0, 0};
  size_t steps = g_timing.used / 5;
  const size_t skip = steps > 20 ? 10 : 0;  // warm-up sub-steps
  for (size_t k = skip; k < steps; ++k)
    for (int p = 0; p < 4; ++p) {
      float ms = 0.f;
      cudaEventElapsedTime(&ms, g_timing.ev[k * 5 + p], g_timing.ev[k * 5 + p + 1]);
      sum[p] += ms;
    }
  const double n = steps > skip ? (double)(steps - skip) : 1.0;
  fprintf(stderr, "clsph dist timing rank %d over %zu sub-steps (us): classify %.1f, signal %.1f, unpack+wait %.1f, finish %.1f\n", rank,
          steps - skip, 1e3 * sum[0] / n, 1e3 * sum[1] / n, 1e3 * sum[2] / n, 1e3 * sum[3] / n);
  for (cudaEvent_t e : g_timing.ev) cudaEventDestroy(e);
  g_timing = ExchangeTiming{};
}
}  // namespace

// Peer transport set-up: one mailbox per rank, its IPC handle all-gathered (as an all-reduce(max) over a zeroed
// table in which every rank fills its own row: the one collective dist.cu already uses), every mailbox mapped.
// All ranks agree on the outcome (a second all-reduce), so either all use peer stores or all use NCCL.
static bool setup_peer_transport(DistState* d) {
  const char* env = getenv("CLSPH_DIST_TRANSPORT");
  bool want = !(env && !strcmp(env, "nccl")) && d->world <= kMaxPeers;
  ncclComm_t comm = static_cast<ncclComm_t>(d->comm);
  const size_t words_per_rank = sizeof(cudaIpcMemHandle_t) / 4;
  uint32_t* table = nullptr;   // device: world rows of handle words, then one word of "this rank is fine"
  const size_t table_words = words_per_rank * d->world;
  if (cudaMalloc(&table, (table_words + 1) * 4) != cudaSuccess) return false;
  cudaMemset(table, 0, (table_words + 1) * 4);
  d->mailbox_bytes = kMailboxRecords + 2 * side_bytes(d->emax, d->gmax);
  cudaIpcMemHandle_t mine;
  memset(&mine, 0, sizeof(mine));
  if (want) {
    want = cudaMalloc(&d->mailbox, d->mailbox_bytes) == cudaSuccess && cudaMemset(d->mailbox, 0, kMailboxRecords) == cudaSuccess &&
           cudaIpcGetMemHandle(&mine, d->mailbox) == cudaSuccess;
    cudaGetLastError();
  }
  if (want) cudaMemcpy(table + words_per_rank * d->rank, &mine, sizeof(mine), cudaMemcpyHostToDevice);
  bool ok = nccl_check(nccl().AllReduce(table, table, table_words, ncclUint32, ncclMax, comm, nullptr), "ncclAllReduce(handles)");
  ok = ok && cudaStreamSynchronize(nullptr) == cudaSuccess;
  std::vector<cudaIpcMemHandle_t> handles(d->world);
  if (ok) cudaMemcpy(handles.data(), table, table_words * 4, cudaMemcpyDeviceToHost);
  bool mapped = ok && want;
  for (int r = 0; r < d->world && mapped; ++r) {
    if (r == d->rank) { d->peer_mailbox[r] = d->mailbox; continue; }
    void* p = nullptr;
    if (cudaIpcOpenMemHandle(&p, handles[r], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); mapped = false; break; }
    d->peer_mailbox[r] = p;
  }
  // 1 = this rank cannot: max over ranks tells everybody
  const uint32_t bad = mapped ? 0u : 1u;
  cudaMemcpy(table + table_words, &bad, 4, cudaMemcpyHostToDevice);
  uint32_t any_bad = 1u;
  if (ok && nccl_check(nccl().AllReduce(table + table_words, table + table_words, 1, ncclUint32, ncclMax, comm, nullptr), "ncclAllReduce(transport)") &&
      cudaStreamSynchronize(nullptr) == cudaSuccess)
    cudaMemcpy(&any_bad, table + table_words, 4, cudaMemcpyDeviceToHost);
  cudaFree(table);
  if (any_bad) {
    for (int r = 0; r < d->world; ++r) {
      if (r != d->rank && d->peer_mailbox[r]) cudaIpcCloseMemHandle(d->peer_mailbox[r]);
      d->peer_mailbox[r] = nullptr;
    }
    cudaFree(d->mailbox);
    d->mailbox = nullptr;
    cudaGetLastError();
    return false;
  }
  if (cudaMalloc(&d->peer_table, sizeof(void*) * kMaxPeers) != cudaSuccess) return false;
  cudaMemcpy(d->peer_table, d->peer_mailbox, sizeof(void*) * d->world, cudaMemcpyHostToDevice);
  return true;
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
  d->seq = 0;
  d->bounds_published = false;
  if (cudaMalloc(&d->counters, 64) != cudaSuccess) return 1;
  cudaMemset(d->counters, 0, 64);
  d->peer = setup_peer_transport(d);
  if (const char* t = getenv("CLSPH_DIST_TIMING")) {
    if (atoi(t) != 0 && !g_timing.on) {
      g_timing.on = true;
      g_timing.ev.resize(5 * 256);
      for (cudaEvent_t& e : g_timing.ev) cudaEventCreate(&e);
    }
  }
  if (!d->peer) {  // NCCL messages: local send and receive buffers
    for (int s = 0; s < 2; ++s) {
      if (cudaMalloc(&d->send[s], d->msg_bytes) != cudaSuccess || cudaMalloc(&d->recv[s], d->msg_bytes) != cudaSuccess) {
        snprintf(g_nccl_error, sizeof(g_nccl_error), "cudaMalloc of %zu-byte message buffers failed", d->msg_bytes);
        return 1;
      }
      cudaMemset(d->send[s], 0, 16);
      cudaMemset(d->recv[s], 0, 16);
    }
  }
  d->active = true;
  return 0;
}

const char* dist_transport(const DistState* d) {
  if (!d->active) return "none";
  return d->peer ? "peer stores into the neighbours' mailboxes (CUDA IPC over NVLink), no collective per sub-step" : "nccl send/recv + all-reduce";
}

void dist_destroy(DistState* d) {
  if (!d->active) return;
  timing_report(d->rank);
  for (int s = 0; s < 2; ++s) {
    cudaFree(d->send[s]);
    cudaFree(d->recv[s]);
  }
  if (d->peer) {
    // the other ranks may still be storing into this mailbox (the AABB they publish at the end of their last
    // sub-step): leave together
    k_dist_close<<<1, 32>>>(d->peer_table, d->mailbox, d->rank, d->world);
    cudaDeviceSynchronize();
    for (int r = 0; r < d->world; ++r)
      if (r != d->rank && d->peer_mailbox[r]) cudaIpcCloseMemHandle(d->peer_mailbox[r]);
    cudaFree(d->peer_table);
    cudaFree(d->mailbox);
  }
  cudaFree(d->counters);
  if (d->comm && nccl().ok) nccl().CommDestroy(static_cast<ncclComm_t>(d->comm));
  d->active = false;
}

// Start of a sub-step: the global AABB into acc. Advances the sequence number.
int dist_reduce_bounds(DistState* d, BoundsAcc* acc, GridState* grid, cudaStream_t stream, uint64_t* launches) {
  ++d->seq;
  if (d->peer) {
    if (!d->bounds_published) {  // first sub-step after an upload: nothing was published at the end of a previous one
      k_bounds_publish<<<1, 32, 0, stream>>>(acc, d->peer_table, d->rank, d->world, d->seq);
      if (launches) ++*launches;
    }
    d->bounds_published = false;
    k_bounds_gather<<<1, 32, 0, stream>>>(acc, d->mailbox, d->world, d->seq, grid);
    if (launches) ++*launches;
    return 0;
  }
  ncclComm_t comm = static_cast<ncclComm_t>(d->comm);
  if (!nccl_check(nccl().GroupStart(), "ncclGroupStart")) return 1;
  bool ok = nccl_check(nccl().AllReduce(acc->lo, acc->lo, 3, ncclUint32, ncclMin, comm, stream), "ncclAllReduce(min)");
  ok = ok && nccl_check(nccl().AllReduce(acc->hi, acc->hi, 3, ncclUint32, ncclMax, comm, stream), "ncclAllReduce(max)");
  if (!nccl_check(nccl().GroupEnd(), "ncclGroupEnd")) return 1;
  return ok ? 0 : 1;
}

// End of a sub-step (the integrator has accumulated the next AABB of this rank's particles): peer transport
// sends it on its way now, a whole neighbour-search ahead of when it is needed.
void dist_publish_bounds(DistState* d, const BoundsAcc* acc, cudaStream_t stream, uint64_t* launches) {
  if (!d->peer) return;
  k_bounds_publish<<<1, 32, 0, stream>>>(acc, d->peer_table, d->rank, d->world, d->seq + 1u);
  d->bounds_published = true;
  if (launches) ++*launches;
}

// The particles were replaced (upload): what was published for the next sub-step is stale. Skipping two
// sequence numbers keeps the slot parity and makes every rank wait for the fresh values.
void dist_invalidate_bounds(DistState* d) {
  if (d->bounds_published) d->seq += 2u;
  d->bounds_published = false;
}

// `live` != null: in place (k_dist_select): u must be prev, u_pid prev_pid; the sort then takes its input through `live`.
int dist_exchange(DistState* d, const StateArrays& prev, const uint32_t* prev_pid, const uint32_t* skey,
                  const uint32_t* wrank, GridState* grid, const StateArrays& u, uint32_t* u_pid, uint32_t* u_ordk,
                  uint32_t* u_ordr, uint32_t capacity, uint32_t* live, cudaStream_t stream, uint64_t* launches) {
  ncclComm_t comm = static_cast<ncclComm_t>(d->comm);
  // counters (zero at the start of every exchange: the last kernel of the previous one re-arms them):
  // [0] local count, [1] export count (clsph_dist_download), [2] / [3] CTAs done in classify / unpack,
  // [4..5] emigrants, ghosts to the left, [6..7] to the right
  uint32_t* u_count = d->counters;
  const bool has_left = d->rank > 0, has_right = d->rank + 1 < d->world;
  const unsigned blocks = (capacity + 255) / 256;
  const unsigned ublocks = (d->emax + d->gmax + 255) / 256;
  if (d->peer) {
    timing_mark(stream);
    // my left neighbour receives "from its right" (side 1), my right neighbour "from its left" (side 0)
    void* lbox = has_left ? d->peer_mailbox[d->rank - 1] : d->mailbox;  // (without a neighbour nothing is ever appended)
    void* rbox = has_right ? d->peer_mailbox[d->rank + 1] : d->mailbox;
    const MsgOut left{d->counters + 4, mailbox_emigrants(lbox, 1, d->emax, d->gmax), mailbox_ghosts(lbox, 1, d->emax, d->gmax)};
    const MsgOut right{d->counters + 6, mailbox_emigrants(rbox, 0, d->emax, d->gmax), mailbox_ghosts(rbox, 0, d->emax, d->gmax)};
    const PeerSignal sig{d->counters + 2, has_left ? mailbox_header(lbox, 1) : nullptr, has_right ? mailbox_header(rbox, 0) : nullptr, d->seq};
    if (live)
      k_dist_select<<<blocks, 256, 0, stream>>>(prev.pos, prev.vel, prev.ivel, prev_pid, skey, wrank, grid, u_ordk, u_ordr, live, u_count,
                                                capacity, left, right, d->emax, d->gmax, sig);
    else
      k_dist_classify<<<blocks, 256, 0, stream>>>(prev.pos, prev.vel, prev.ivel, prev_pid, skey, wrank, grid, u.pos, u.vel, u.ivel,
                                                  u_pid, u_ordk, u_ordr, u_count, capacity, left, right, d->emax, d->gmax, sig);
    timing_mark(stream);
    timing_mark(stream);
    if (has_left || has_right) {
      const MsgIn from_left{has_left ? mailbox_header(d->mailbox, 0) : nullptr, mailbox_emigrants(d->mailbox, 0, d->emax, d->gmax),
                            mailbox_ghosts(d->mailbox, 0, d->emax, d->gmax)};
      const MsgIn from_right{has_right ? mailbox_header(d->mailbox, 1) : nullptr, mailbox_emigrants(d->mailbox, 1, d->emax, d->gmax),
                             mailbox_ghosts(d->mailbox, 1, d->emax, d->gmax)};
      k_dist_unpack<<<2 * ublocks, 256, 0, stream>>>(from_left, from_right, d->seq, d->emax, d->gmax, grid, u.pos, u.vel, u.ivel, u_pid,
                                                     u_ordk, u_ordr, d->counters, capacity, live);
    } else {
      k_dist_finish<<<1, 32, 0, stream>>>(grid, d->counters, capacity);
    }
    timing_mark(stream);
    timing_mark(stream);
    if (launches) *launches += 2;
    return 0;
  }
  cudaMemsetAsync(d->send[0], 0, 16, stream);
  cudaMemsetAsync(d->send[1], 0, 16, stream);
  const MsgOut left{static_cast<uint32_t*>(d->send[0]), msg_emigrants(d->send[0]), msg_ghosts(d->send[0], d->emax)};
  const MsgOut right{static_cast<uint32_t*>(d->send[1]), msg_emigrants(d->send[1]), msg_ghosts(d->send[1], d->emax)};
  const PeerSignal none{nullptr, nullptr, nullptr, 0u};
  if (live)
    k_dist_select<<<blocks, 256, 0, stream>>>(prev.pos, prev.vel, prev.ivel, prev_pid, skey, wrank, grid, u_ordk, u_ordr, live, u_count,
                                              capacity, left, right, d->emax, d->gmax, none);
  else
    k_dist_classify<<<blocks, 256, 0, stream>>>(prev.pos, prev.vel, prev.ivel, prev_pid, skey, wrank, grid, u.pos, u.vel, u.ivel,
                                                u_pid, u_ordk, u_ordr, u_count, capacity, left, right, d->emax, d->gmax, none);
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
  if (has_left || has_right) {
    const MsgIn from_left{has_left ? static_cast<const uint32_t*>(d->recv[0]) : nullptr, msg_emigrants(d->recv[0]), msg_ghosts(d->recv[0], d->emax)};
    const MsgIn from_right{has_right ? static_cast<const uint32_t*>(d->recv[1]) : nullptr, msg_emigrants(d->recv[1]), msg_ghosts(d->recv[1], d->emax)};
    k_dist_unpack<<<2 * ublocks, 256, 0, stream>>>(from_left, from_right, 0u, d->emax, d->gmax, grid, u.pos, u.vel, u.ivel, u_pid, u_ordk,
                                                   u_ordr, d->counters, capacity, live);
  } else {
    k_dist_finish<<<1, 32, 0, stream>>>(grid, d->counters, capacity);
  }
  if (launches) *launches += 2;
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
