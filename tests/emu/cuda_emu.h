// cuda_emu.h -- TEST INFRASTRUCTURE ONLY: a CPU stand-in for the slice of CUDA that
// libclsph_b200/csrc uses, so that the KERNEL LOGIC (indexing, warp collectives, barriers,
// look-back, list bookkeeping) can be exercised by `pytest -m "not gpu"` in a container that has
// no GPU. It is not a product path and not a fallback: the product (libclsph_cuda.so, capi.py)
// never loads it; only tests/test_emu_*.py build and load tests/emu/_build/libclsph_emu.so.
//
// Model: a kernel launch runs its blocks one after the other; the threads of a block are
// coroutines (ucontext) on the calling OS thread. A thread runs until it reaches a warp
// collective (__shfl_sync, __ballot_sync, __match_any_sync, __syncwarp, ...) or __syncthreads();
// the collective completes when every thread named by its mask (every live thread of the block
// for __syncthreads) waits in the same kind of collective with the same mask, as Volta+
// independent thread scheduling requires. A collective that names an exited lane, or a block in
// which every live thread waits and nothing can complete, aborts with a diagnostic.
//
// What this cannot show: performance, memory-model races, misaligned or out-of-bounds accesses
// that happen to land in mapped host memory. The GPU tests (-m gpu) remain the parity gate.
#pragma once

#include <algorithm>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <type_traits>

#define CLSPH_EMU 1
#define __host__
#define __device__
#define __global__
#define __forceinline__ inline __attribute__((always_inline))
#define __launch_bounds__(...)

// ---- vector types -------------------------------------------------------------------------
struct alignas(16) float4 {
  float x, y, z, w;
};
inline float4 make_float4(float x, float y, float z, float w) { float4 r; r.x = x; r.y = y; r.z = z; r.w = w; return r; }
struct alignas(8) uint2 {
  unsigned x, y;
};
inline uint2 make_uint2(unsigned x, unsigned y) { uint2 r; r.x = x; r.y = y; return r; }
struct uint3 {
  unsigned x, y, z;
};
struct alignas(16) uint4 {
  unsigned x, y, z, w;
};
inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { uint4 v; v.x = x; v.y = y; v.z = z; v.w = w; return v; }
struct dim3 {
  unsigned x, y, z;
  dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {}
};

// ---- execution context ----------------------------------------------------------------------
namespace emu {

struct ThreadCtx {
  uint3 tid, bid;
  dim3 bdim, gdim;
};
extern thread_local ThreadCtx* cur;

enum Op { kShflIdx, kShflXor, kShflUp, kBallot, kAny, kMatchAny, kSyncWarp };
uint64_t warp_collective(Op op, unsigned mask, uint64_t in, int aux);
void block_barrier();
void* dyn_smem();

// Runs `fn(arg)` once per thread of every block of the grid.
void launch_raw(dim3 grid, dim3 block, size_t smem_bytes, void (*fn)(void*), void* arg);

template <class F>
inline void launch(dim3 grid, dim3 block, size_t smem_bytes, F&& f) {
  launch_raw(grid, block, smem_bytes, [](void* p) { (*static_cast<typename std::remove_reference<F>::type*>(p))(); },
             const_cast<void*>(static_cast<const void*>(&f)));
}

template <class T>
inline uint64_t to_bits(T v) {
  static_assert(sizeof(T) <= 8, "collective payload too large");
  uint64_t b = 0;
  std::memcpy(&b, &v, sizeof(T));
  return b;
}
template <class T>
inline T from_bits(uint64_t b) {
  T v;
  std::memcpy(&v, &b, sizeof(T));
  return v;
}

void* fake_dlopen(const char* name, int flags);
void* fake_dlsym(void* handle, const char* name);

}  // namespace emu

#define threadIdx (emu::cur->tid)
#define blockIdx (emu::cur->bid)
#define blockDim (emu::cur->bdim)
#define gridDim (emu::cur->gdim)

// ---- warp / block collectives -------------------------------------------------------------------
template <class T>
inline T __shfl_sync(unsigned mask, T v, int src, int width = 32) {
  (void)width;
  return emu::from_bits<T>(emu::warp_collective(emu::kShflIdx, mask, emu::to_bits(v), src));
}
template <class T>
inline T __shfl_xor_sync(unsigned mask, T v, int lane_mask, int width = 32) {
  (void)width;
  return emu::from_bits<T>(emu::warp_collective(emu::kShflXor, mask, emu::to_bits(v), lane_mask));
}
template <class T>
inline T __shfl_up_sync(unsigned mask, T v, unsigned delta, int width = 32) {
  (void)width;
  return emu::from_bits<T>(emu::warp_collective(emu::kShflUp, mask, emu::to_bits(v), (int)delta));
}
inline unsigned __ballot_sync(unsigned mask, int pred) {
  return (unsigned)emu::warp_collective(emu::kBallot, mask, pred ? 1u : 0u, 0);
}
inline int __any_sync(unsigned mask, int pred) { return (int)emu::warp_collective(emu::kAny, mask, pred ? 1u : 0u, 0); }
template <class T>
inline unsigned __match_any_sync(unsigned mask, T v) {
  return (unsigned)emu::warp_collective(emu::kMatchAny, mask, emu::to_bits(v), 0);
}
inline void __syncwarp(unsigned mask = 0xffffffffu) { emu::warp_collective(emu::kSyncWarp, mask, 0, 0); }
inline void __syncthreads() { emu::block_barrier(); }

// ---- arithmetic intrinsics (build with -ffp-contract=off: one rounding per operation) --------------
inline float __fadd_rn(float a, float b) { volatile float r = a + b; return r; }
inline float __fsub_rn(float a, float b) { volatile float r = a - b; return r; }
inline float __fmul_rn(float a, float b) { volatile float r = a * b; return r; }
inline float __fdiv_rn(float a, float b) { volatile float r = a / b; return r; }
inline float __fmaf_rn(float a, float b, float c) { return std::fmaf(a, b, c); }
inline float __fsqrt_rn(float a) { return sqrtf(a); }
inline float rsqrtf(float a) { return 1.0f / sqrtf(a); }  // MUFU.RSQ on the GPU (2 ulp)
inline unsigned __float2uint_rz(float f) {  // cvt.rzi.u32.f32 saturates; NaN -> 0
  if (!(f > 0.f)) return 0u;
  if (f >= 4294967296.f) return 0xffffffffu;
  return (unsigned)f;
}
inline unsigned __float_as_uint(float f) { return emu::from_bits<unsigned>(emu::to_bits(f)); }
inline float __uint_as_float(unsigned u) { return emu::from_bits<float>(emu::to_bits(u)); }
inline float __int_as_float(int i) { return emu::from_bits<float>(emu::to_bits(i)); }
inline int __float_as_int(float f) { return emu::from_bits<int>(emu::to_bits(f)); }
inline int __popc(unsigned v) { return __builtin_popcount(v); }
inline int __ffs(int v) { return __builtin_ffs(v); }
inline int __clz(int v) { return v == 0 ? 32 : __builtin_clz((unsigned)v); }
template <class T>
inline T __ldg(const T* p) { return *p; }

template <class T>
inline T min(T a, T b) { return b < a ? b : a; }
template <class T>
inline T max(T a, T b) { return a < b ? b : a; }

// ---- atomics (one OS thread per emulated device: plain read-modify-write) ---------------------------
template <class T>
inline T atomicAdd(T* p, T v) { T o = *p; *p = o + v; return o; }
template <class T>
inline T atomicMin(T* p, T v) { T o = *p; if (v < o) *p = v; return o; }
template <class T>
inline T atomicMax(T* p, T v) { T o = *p; if (o < v) *p = v; return o; }
template <class T>
inline T atomicOr(T* p, T v) { T o = *p; *p = o | v; return o; }

// ---- runtime API ----------------------------------------------------------------------------------
typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorMemoryAllocation = 2, cudaErrorInvalidValue = 1 };
enum cudaMemcpyKind { cudaMemcpyHostToHost, cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice };
enum { cudaStreamNonBlocking = 1 };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
typedef struct emuStream* cudaStream_t;
typedef struct emuEvent* cudaEvent_t;
struct cudaDeviceProp {
  int multiProcessorCount;
  char name[64];
};

enum cudaDeviceAttr { cudaDevAttrMaxSharedMemoryPerBlockOptin = 97, cudaDevAttrMaxSharedMemoryPerMultiprocessor = 81, cudaDevAttrMultiProcessorCount = 16 };
// ---- peer memory between ranks: in the emulator every rank is a thread of one process, a handle is the pointer
struct cudaIpcMemHandle_t { char reserved[64]; };
enum { cudaIpcMemLazyEnablePeerAccess = 1 };
inline cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t* h, void* p) {
  std::memset(h, 0, sizeof(*h));
  std::memcpy(h->reserved, &p, sizeof(p));
  return cudaSuccess;
}
inline cudaError_t cudaIpcOpenMemHandle(void** p, cudaIpcMemHandle_t h, unsigned) {
  std::memcpy(p, h.reserved, sizeof(*p));
  return cudaSuccess;
}
inline cudaError_t cudaIpcCloseMemHandle(void*) { return cudaSuccess; }
inline void __threadfence_system() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
inline void __threadfence() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
template <class T>
inline T __ldcg(const T* p) { return *p; }
long long clock64();  // nanoseconds (the product code only compares differences with a generous limit)
inline cudaError_t cudaGetDevice(int* d) { *d = 0; return cudaSuccess; }
inline cudaError_t cudaDeviceGetAttribute(int* v, cudaDeviceAttr a, int) {  // the B200's figures
  *v = a == cudaDevAttrMultiProcessorCount ? 4 : (a == cudaDevAttrMaxSharedMemoryPerBlockOptin ? 227 * 1024 : 228 * 1024);  // (4 SMs: small grids reach the persistent paths)
  return cudaSuccess;
}
cudaError_t cudaGetDeviceCount(int* n);
cudaError_t cudaSetDevice(int d);
cudaError_t cudaGetDeviceProperties(cudaDeviceProp* p, int d);
cudaError_t cudaGetLastError();
const char* cudaGetErrorString(cudaError_t e);
cudaError_t cudaMalloc(void** p, size_t bytes);
template <class T>
inline cudaError_t cudaMalloc(T** p, size_t bytes) { return cudaMalloc(reinterpret_cast<void**>(p), bytes); }
cudaError_t cudaFree(void* p);
cudaError_t cudaMemcpy(void* dst, const void* src, size_t bytes, cudaMemcpyKind kind);
cudaError_t cudaMemcpyAsync(void* dst, const void* src, size_t bytes, cudaMemcpyKind kind, cudaStream_t s = nullptr);
cudaError_t cudaMemset(void* dst, int v, size_t bytes);
cudaError_t cudaMemsetAsync(void* dst, int v, size_t bytes, cudaStream_t s = nullptr);
cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned flags);
cudaError_t cudaStreamSynchronize(cudaStream_t s);
inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
enum { cudaHostAllocDefault = 0 };
// every host pointer doubles as its own device pointer here, so the zero-copy upload kernel runs in the emulator too
enum cudaMemoryType { cudaMemoryTypeUnregistered = 0, cudaMemoryTypeHost = 1, cudaMemoryTypeDevice = 2 };
struct cudaPointerAttributes { cudaMemoryType type; void* devicePointer; void* hostPointer; int device; };
inline cudaError_t cudaPointerGetAttributes(cudaPointerAttributes* a, const void* p) {
  a->type = cudaMemoryTypeHost; a->devicePointer = const_cast<void*>(p); a->hostPointer = const_cast<void*>(p); a->device = 0;
  return cudaSuccess;
}
inline cudaError_t cudaHostAlloc(void** p, size_t bytes, unsigned) { *p = std::malloc(bytes); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
inline cudaError_t cudaFreeHost(void* p) { std::free(p); return cudaSuccess; }
cudaError_t cudaStreamDestroy(cudaStream_t s);
cudaError_t cudaEventCreate(cudaEvent_t* e);
enum { cudaEventDisableTiming = 2 };
inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { return cudaEventCreate(e); }
inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned = 0) { return cudaSuccess; }  // launches are synchronous here
cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t s = nullptr);
cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t a, cudaEvent_t b);
cudaError_t cudaEventDestroy(cudaEvent_t e);
template <class F>
inline cudaError_t cudaFuncSetAttribute(F, cudaFuncAttribute, int) { return cudaSuccess; }
