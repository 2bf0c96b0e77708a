// kernels.cuh -- launcher declarations shared by the translation units of libclsph_cuda.so.
#pragma once

#include <algorithm>
#include <vector>

#include "common.cuh"

namespace clsph {

constexpr int kRadix = 256;
constexpr int kMaxSortPasses = 4;
constexpr int kSortThreads = 256;
constexpr int kSortItems = 16;
constexpr int kSortTile = kSortThreads * kSortItems;

// One ping-pong side of the particle state (device pointers, capacity = max_particles).
struct StateArrays {
  float4* pos;
  float4* vel;
  float4* ivel;
};

struct SortBuffers {
  uint32_t* keys_a;   // keys in input order (pass 0 source)
  uint32_t* keys_b;
  uint32_t* vals_a;
  uint32_t* vals_b;
  uint32_t* scratch;  // sort_scratch_words() words
};

// Device words zeroed by the launch of k_grid_setup at the start of a sub-step (sub-cell order).
struct StepZero {
  uint32_t* sub_lb;            // what the last sub-step wrote (GridState::table_words ...)
  const uint32_t* keys_a;      // the sort's key buffers: the sorted keys of the last sub-step are still in one of them
  const uint32_t* keys_b;
  uint32_t* scan_state;
  uint32_t scan_words;
  uint32_t* sort_scratch;
  size_t sort_scratch_words;
  uint32_t* pair_count;        // 2 words (nullable)
};

// Precomputed triangle record for the collision pass (5 x float4).
struct Face {
  float nx, ny, nz, nlen;   // scene normal and its length()
  float ax, ay, az, uv;     // vertex 0, dot(u, v)
  float ux, uy, uz, uu;     // edge v1 - v0, dot(u, u)
  float vx, vy, vz, vv;     // edge v2 - v0, dot(v, v)
  float det, pad0, pad1, pad2;  // uv*uv - uu*vv
};

// Uniform grid over the scene's triangles (integrate.cu, built on the host by build_face_grid):
// a particle only tests the faces registered in the cells its sub-step segment touches. The
// registration is conservative (padded boxes), so the faces that can pass the reference's hit test
// are always among them and the result is bit-identical to testing every face.
struct FaceGrid {
  float ox, oy, oz;   // origin
  float inv;          // cells per unit length
  int nx, ny, nz;     // 0 cells: no grid, every face is tested
  uint32_t n_global;  // faces tested for every particle (ill-conditioned or non-finite triangles)
  const uint32_t* cell_start;  // [nx * ny * nz + 1] offsets into ids
  const uint32_t* ids;         // face indices per cell, ascending
  const uint32_t* global_ids;  // ascending
};

struct DebugTaps {           // all nullable; sorted order unless stated
  uint32_t* keys_input;      // pre-step order
  uint32_t* candidate_count;
  uint32_t* support_count;
  float4* acceleration;      // always allocated (the force pass writes it for the integrator)
  uint32_t* collision_iters;
};

// ---- sort.cu
uint32_t sort_tiles_for(uint32_t n);
size_t sort_scratch_words(uint32_t max_particles);
// sub_keys: keys are (cell key << 3 | octant) instead of the cell key (subgrid.cu).
void launch_sort_keys(const SortBuffers& b, const float4* pos, const GridState* grid, uint32_t n_launch, int sm_count,
                      uint32_t* keys_tap, bool sub_keys, uint32_t* sub_lb, const uint32_t* index, cudaStream_t stream, uint64_t* launches);
void launch_sort_passes(const SortBuffers& b, const GridState* grid, uint32_t n_launch, const uint32_t* first_vals, const uint32_t* table,
                        cudaStream_t stream, uint64_t* launches);
// counting sort on the dense sub-cell table (grid->sort_passes == 0): scan state words, scan of the table
uint32_t scan_state_words(uint32_t sub_capacity);
size_t sort_scratch_zero_words(uint32_t n);  // leading words of SortBuffers::scratch that must be zero before a sort of n keys
void launch_scan_table(uint32_t* sub_lb, const GridState* grid, uint32_t* scan_state, uint32_t sub_capacity, int sm_count,
                       cudaStream_t stream, uint64_t* launches);

// ---- grid.cu
void launch_bounds_reset(BoundsAcc* acc, cudaStream_t stream, uint64_t* launches);
void launch_bounds(const float4* pos, uint32_t n, BoundsAcc* acc, int sm_count, cudaStream_t stream, uint64_t* launches);
void launch_grid_setup(BoundsAcc* acc, GridState* grid, float h, uint32_t n, uint32_t cell_capacity, float plane_lo,
                       float plane_hi, bool keep_n, uint32_t sub_mode, uint32_t sub_capacity, uint32_t count_sort, const StepZero& zero,
                       int sm_count, cudaStream_t stream, uint64_t* launches);
void launch_clear_cells(uint32_t* cell_start, uint32_t* cell_end, const GridState* grid, uint32_t cell_capacity,
                        int sm_count, cudaStream_t stream, uint64_t* launches);
// Gathers `src` into `dst` through the sort permutation, writes sorted keys and the cell table.
void launch_reorder(const StateArrays& src, const StateArrays& dst, const SortBuffers& sort, uint32_t* skey,
                    uint32_t* perm_out, uint32_t* cell_start, uint32_t* cell_end, const GridState* grid,
                    const uint32_t* src_pid, uint32_t* dst_pid, uint32_t n_launch, cudaStream_t stream,
                    uint64_t* launches);
// rrank (nullable): reference rank per particle (subgrid.cu); set to the identity on upload, and
// record i goes to slot rrank[i] of the AoS array on download.
void launch_aos_to_soa(const void* aos, const StateArrays& dst, float4* aux, uint32_t* skey, float4* accel,
                       uint32_t* rrank, uint32_t n, cudaStream_t stream, uint64_t* launches);
// The same from page-locked host memory mapped into the device's address space, reading only what a sub-step needs.
void launch_aos_to_soa_host(const void* aos_host_mapped, const StateArrays& dst, float4* aux, uint32_t* skey, uint32_t* rrank,
                            uint32_t n, cudaStream_t stream, uint64_t* launches);
void launch_soa_to_aos(const StateArrays& src, const float4* aux, const uint32_t* skey, const uint32_t* rrank, void* aos,
                       uint32_t n, cudaStream_t stream, uint64_t* launches);
// Seven floats per particle (position, velocity, density) in the reference's order: what a frame file needs.
void launch_pack_frame(const StateArrays& src, const float4* aux, const uint32_t* rrank, float* out, uint32_t n, cudaStream_t stream,
                       uint64_t* launches);
void launch_reference_cell_table(const uint32_t* skey, const GridState* grid, uint32_t* table, uint32_t n_launch,
                                 cudaStream_t stream, uint64_t* launches);
void launch_copy_u32(const uint32_t* src, uint32_t* dst, uint32_t n, cudaStream_t stream, uint64_t* launches);

// Per-particle neighbour lists in HBM (list mode). rows == 0 selects the two-pass kernels.
struct NeighbourLists {
  uint32_t* entries = nullptr;  // [capacity][rows]: entries[i * rows + e] = e-th neighbour of particle i
  uint32_t* count = nullptr;    // [capacity]: entries of particle i, 0xFFFFFFFF = more than `rows`
  uint32_t rows = 0;
  uint32_t* window_counter = nullptr;  // work counter of the persistent list-building kernel
};

// ---- neighbors.cu
void neighbors_init();  // opts the kernels into their dynamic shared memory size
// Also leaves p/rho^2 in pos[].w and m/rho in vel[].w for the force pass.
void launch_density(float4* pos, float4* vel, const uint32_t* skey, const uint32_t* cell_start, const uint32_t* cell_end,
                    const GridState* grid, const SphConst& c, float4* aux, const NeighbourLists& lists,
                    const DebugTaps& taps, bool debug, uint32_t n_launch, int sm_count, cudaStream_t stream,
                    uint64_t* launches);
// Only particles of owned cells get an acceleration (multi-GPU: ghosts are skipped).
// search_fallback: in list mode, also run the searching kernel for particles whose list overflowed
// (sub-cell order has its own, launch_forces_sub_overflow). fast_pairs: pair terms through add_pair_fast.
// dense_occupancy: the list kernel compiled for four resident CTAs per SM instead of three.
void launch_forces(const float4* pos, const float4* vel, const float4* aux, const uint32_t* skey,
                   const uint32_t* cell_start, const uint32_t* cell_end, const GridState* grid, const SphConst& c,
                   const NeighbourLists& lists, bool search_fallback, bool fast_pairs, bool dense_occupancy, float4* accel,
                   uint32_t n_launch, cudaStream_t stream, uint64_t* launches, bool factored = false);

// ---- subgrid.cu: sub-cell order (arrays sorted by cell key << 3 | octant)
void launch_reorder_sub(const StateArrays& src, const StateArrays& dst, const SortBuffers& sort, uint32_t* skey,
                        const uint32_t* rr_src, uint32_t* rr_dst, uint32_t* sub_lb, const GridState* grid,
                        const uint32_t* src_pid, uint32_t* dst_pid, const uint32_t* src_ordk, const uint32_t* src_ordr,
                        uint32_t* dst_ordk, uint32_t* dst_ordr, uint32_t* pair_items, uint32_t* pair_count, uint32_t n_launch,
                        cudaStream_t stream, uint64_t* launches);
// Density + pressure + neighbour lists, two particles of a sub-cell per thread (packed fp32: FADD2 / FFMA2).
void launch_density_pairs(float4* pos, float4* vel, const uint32_t* skey, const uint32_t* sub_lb, const SortBuffers& sort,
                          const GridState* grid, const SphConst& c, float4* aux, const NeighbourLists& lists,
                          const DebugTaps& taps, bool debug, int variant, const uint32_t* pair_items, const uint32_t* pair_count,
                          uint32_t n_launch, cudaStream_t stream, uint64_t* launches);
void launch_rank_pair(const float4* pos, const uint32_t* skey, const uint32_t* ordk, const uint32_t* ordr, uint32_t* wrank,
                      const uint32_t* sub_lb, const SortBuffers& sort, const GridState* grid, uint32_t n_launch,
                      cudaStream_t stream, uint64_t* launches);
void launch_rank(const uint32_t* skey, const uint32_t* rr_old, uint32_t* rr_new, const uint32_t* sub_lb,
                 const SortBuffers& sort, const GridState* grid, uint32_t* perm_out, uint32_t* keys_input_tap,
                 uint32_t n_launch, cudaStream_t stream, uint64_t* launches);
void launch_density_sub(float4* pos, float4* vel, const uint32_t* skey, const uint32_t* sub_lb, const SortBuffers& sort,
                        const GridState* grid, const SphConst& c, float4* aux, const NeighbourLists& lists,
                        const DebugTaps& taps, bool debug, bool merged, uint32_t n_launch, cudaStream_t stream,
                        uint64_t* launches);
void launch_forces_sub_overflow(const float4* pos, const float4* vel, const float4* aux, const uint32_t* skey,
                                const uint32_t* sub_lb, const SortBuffers& sort, const GridState* grid, const SphConst& c,
                                const NeighbourLists& lists, float4* accel, const uint32_t* overflowed, uint32_t n_launch,
                                cudaStream_t stream, uint64_t* launches);
void launch_scatter_words(const void* src, const uint32_t* rrank, void* dst, uint32_t n, uint32_t words,
                          cudaStream_t stream, uint64_t* launches);

// ---- integrate.cu
void launch_prepare_faces(const float* normals, const float* vertices, const uint32_t* indices, uint32_t face_count,
                          Face* faces, cudaStream_t stream, uint64_t* launches);
void launch_integrate(const StateArrays& s, const float4* accel, const uint32_t* skey, const Face* faces,
                      uint32_t face_count, const FaceGrid& face_grid, const GridState* grid, const SphConst& c,
                      BoundsAcc* next_bounds, uint32_t* iters_tap, uint32_t n_launch, int sm_count, cudaStream_t stream,
                      uint64_t* launches);
// Host side: cell lists for the triangles (vertices 3V floats, indices 3F). Fills the three vectors
// and the geometry of `out`; the caller uploads them and sets the pointers.
void build_face_grid(const float* vertices, const uint32_t* indices, uint32_t face_count, FaceGrid* out,
                     std::vector<uint32_t>* cell_start, std::vector<uint32_t>* ids, std::vector<uint32_t>* global_ids);

}  // namespace clsph
