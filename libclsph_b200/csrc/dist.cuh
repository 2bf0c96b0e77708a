// dist.cuh -- multi-GPU slab decomposition (see dist.cu).
#pragma once

#include <cstring>

#include "kernels.cuh"

namespace clsph {

struct DistState {
  bool active = false;
  void* comm = nullptr;       // ncclComm_t
  int rank = 0, world = 1;
  float plane_lo = 0.f, plane_hi = 0.f;  // world-space x planes of this rank's slab (-inf / +inf at the ends)
  uint32_t emax = 0, gmax = 0;           // record capacities of one message: emigrants, ghosts
  size_t msg_bytes = 0;
  void* send[2] = {nullptr, nullptr};    // [0] left neighbour, [1] right neighbour
  void* recv[2] = {nullptr, nullptr};
  uint32_t* counters = nullptr;          // [0] local particle count being assembled, [1] export count, [4..7] records per neighbour
  // peer transport (see dist.cu): records and AABBs are stored straight into the other ranks' mailboxes
  bool peer = false;
  void* mailbox = nullptr;               // this rank's mailbox
  size_t mailbox_bytes = 0;
  void* peer_mailbox[32] = {};           // every rank's mailbox as mapped into this process ([rank] = mailbox)
  void** peer_table = nullptr;           // the same table on the device
  uint32_t seq = 0;                      // sequence number of the current sub-step, the same on every rank
  bool bounds_published = false;         // the AABB for sub-step seq + 1 is already on its way
};

const char* dist_last_error();
int dist_unique_id(void* out, size_t bytes);
int dist_init(DistState* d, int rank, int world, const void* id_bytes, float plane_lo, float plane_hi, uint32_t emax,
              uint32_t gmax);
void dist_destroy(DistState* d);
const char* dist_transport(const DistState* d);
int dist_reduce_bounds(DistState* d, BoundsAcc* acc, GridState* grid, cudaStream_t stream, uint64_t* launches);
void dist_publish_bounds(DistState* d, const BoundsAcc* acc, cudaStream_t stream, uint64_t* launches);
void dist_invalidate_bounds(DistState* d);
// prev (sorted by last step's keys, owned + ghosts) -> u (unsorted: owned + new ghosts); sets grid->n.
// wrank / u_ordk / u_ordr (all null or all set): order keys of the sub-cell order, see k_dist_classify.
int dist_exchange(DistState* d, const StateArrays& prev, const uint32_t* prev_pid, const uint32_t* skey,
                  const uint32_t* wrank, GridState* grid, const StateArrays& u, uint32_t* u_pid, uint32_t* u_ordk,
                  uint32_t* u_ordr, uint32_t capacity, uint32_t* live, cudaStream_t stream, uint64_t* launches);
void launch_dist_export(const StateArrays& s, const float4* aux, const uint32_t* skey, const uint32_t* pid,
                        const uint32_t* wrank, const GridState* grid, void* aos, uint32_t* ids, uint32_t* out_count,
                        uint32_t capacity, cudaStream_t stream, uint64_t* launches);
void launch_fill_ids(uint32_t* pid, const uint32_t* src, uint32_t first, uint32_t n, cudaStream_t stream,
                     uint64_t* launches);

}  // namespace clsph
