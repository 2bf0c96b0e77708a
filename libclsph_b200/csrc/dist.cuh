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
  uint32_t* counters = nullptr;          // [0] local particle count being assembled, [1] export count
};

const char* dist_last_error();
int dist_unique_id(void* out, size_t bytes);
int dist_init(DistState* d, int rank, int world, const void* id_bytes, float plane_lo, float plane_hi, uint32_t emax,
              uint32_t gmax);
void dist_destroy(DistState* d);
int dist_allreduce_bounds(DistState* d, BoundsAcc* acc, cudaStream_t stream);
// prev (sorted by last step's keys, owned + ghosts) -> u (unsorted: owned + new ghosts); sets grid->n.
// wrank / u_ordk / u_ordr (all null or all set): order keys of the sub-cell order, see k_dist_classify.
int dist_exchange(DistState* d, const StateArrays& prev, const uint32_t* prev_pid, const uint32_t* skey,
                  const uint32_t* wrank, GridState* grid, const StateArrays& u, uint32_t* u_pid, uint32_t* u_ordk,
                  uint32_t* u_ordr, uint32_t capacity, cudaStream_t stream, uint64_t* launches);
void launch_dist_export(const StateArrays& s, const float4* aux, const uint32_t* skey, const uint32_t* pid,
                        const uint32_t* wrank, const GridState* grid, void* aos, uint32_t* ids, uint32_t* out_count,
                        uint32_t capacity, cudaStream_t stream, uint64_t* launches);
void launch_fill_ids(uint32_t* pid, const uint32_t* src, uint32_t first, uint32_t n, cudaStream_t stream,
                     uint64_t* launches);

}  // namespace clsph
