// cuda_emu.cpp -- scheduler and runtime shim of the test-only CUDA emulator (see cuda_emu.h).
#include "cuda_emu.h"

#include <pthread.h>
#include <sys/mman.h>
#include <time.h>
#include <ucontext.h>
#include <unistd.h>

#include <cstdio>
#include <vector>

namespace emu {

thread_local ThreadCtx* cur = nullptr;

namespace {

constexpr size_t kStackBytes = 256 * 1024;
constexpr size_t kDynSmemBytes = 256 * 1024;

struct Thread {
  ThreadCtx ctx;
  ucontext_t uc;
  bool done = false;
  // pending warp collective
  bool waiting = false, released = false;
  Op op = kSyncWarp;
  unsigned mask = 0;
  uint64_t in = 0, out = 0;
  int aux = 0;
  // block barrier
  bool at_barrier = false, barrier_released = false;
};

struct Block {
  std::vector<Thread> threads;
  unsigned nthreads = 0, live = 0, barrier_count = 0;
  uint64_t progress = 0;
  ucontext_t sched;
  Thread* running = nullptr;
  void (*fn)(void*) = nullptr;
  void* arg = nullptr;
};

thread_local Block* blk = nullptr;
thread_local char* stack_pool = nullptr;
thread_local size_t stack_pool_threads = 0;
thread_local void* dyn_smem_buf = nullptr;

[[noreturn]] void fatal(const char* what) {
  std::fprintf(stderr, "cuda_emu: %s\n", what);
  if (blk && blk->running) {
    const Thread* t = blk->running;
    std::fprintf(stderr, "  in block %u thread %u (block of %u threads, %u live)\n", t->ctx.bid.x, t->ctx.tid.x, blk->nthreads, blk->live);
  }
  std::fflush(stderr);
  std::abort();
}

void yield() {
  Thread* t = blk->running;
  swapcontext(&t->uc, &blk->sched);
}

void trampoline() {
  Block* b = blk;
  Thread* t = b->running;
  b->fn(b->arg);
  t->done = true;
  --b->live;
  ++b->progress;
  // returning resumes uc_link = the scheduler
}

char* stacks_for(size_t threads) {
  if (threads > stack_pool_threads) {
    if (stack_pool) munmap(stack_pool, stack_pool_threads * kStackBytes);
    void* p = mmap(nullptr, threads * kStackBytes, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
    if (p == MAP_FAILED) fatal("mmap of coroutine stacks failed");
    stack_pool = static_cast<char*>(p);
    stack_pool_threads = threads;
  }
  return stack_pool;
}

}  // namespace

void* dyn_smem() {
  if (!dyn_smem_buf && posix_memalign(&dyn_smem_buf, 128, kDynSmemBytes) != 0) fatal("dynamic shared memory allocation failed");
  return dyn_smem_buf;
}

uint64_t warp_collective(Op op, unsigned mask, uint64_t in, int aux) {
  Block* b = blk;
  if (!b) fatal("warp collective outside a kernel");
  Thread* t = b->running;
  const unsigned linear = t->ctx.tid.x;
  const unsigned lane = linear & 31u, base = linear & ~31u;
  if (!((mask >> lane) & 1u)) fatal("warp collective whose mask does not name the calling lane");
  t->waiting = true;
  t->released = false;
  t->op = op;
  t->mask = mask;
  t->in = in;
  t->aux = aux;
  for (;;) {
    if (t->released) break;
    bool ready = true;
    for (unsigned l = 0; l < 32 && ready; ++l) {
      if (!((mask >> l) & 1u)) continue;
      if (base + l >= b->nthreads) fatal("warp collective names a lane beyond the block");
      const Thread& o = b->threads[base + l];
      if (o.done) fatal("warp collective names a lane that has exited");
      ready = o.waiting && !o.released && o.mask == mask && o.op == op;
    }
    if (ready) {
      unsigned ballot = 0;
      for (unsigned l = 0; l < 32; ++l)
        if (((mask >> l) & 1u) && (b->threads[base + l].in & 1u)) ballot |= 1u << l;
      for (unsigned l = 0; l < 32; ++l) {
        if (!((mask >> l) & 1u)) continue;
        Thread& o = b->threads[base + l];
        int src = (int)l;
        switch (op) {
          case kShflIdx: src = o.aux & 31; break;
          case kShflXor: src = (int)(l ^ (unsigned)o.aux); break;
          case kShflUp: src = (int)l - o.aux; break;
          default: break;
        }
        switch (op) {
          case kShflIdx:
          case kShflXor:
          case kShflUp:
            o.out = (src >= 0 && src < 32 && ((mask >> src) & 1u)) ? b->threads[base + src].in : o.in;
            break;
          case kBallot: o.out = ballot; break;
          case kAny: o.out = ballot != 0u; break;
          case kMatchAny: {
            unsigned m = 0;
            for (unsigned k = 0; k < 32; ++k)
              if (((mask >> k) & 1u) && b->threads[base + k].in == o.in) m |= 1u << k;
            o.out = m;
            break;
          }
          case kSyncWarp: o.out = 0; break;
        }
      }
      for (unsigned l = 0; l < 32; ++l)
        if ((mask >> l) & 1u) b->threads[base + l].released = true;
      ++b->progress;
      break;
    }
    yield();
  }
  t->waiting = false;
  return t->out;
}

void block_barrier() {
  Block* b = blk;
  if (!b) fatal("__syncthreads outside a kernel");
  Thread* t = b->running;
  t->at_barrier = true;
  t->barrier_released = false;
  ++b->barrier_count;
  for (;;) {
    if (t->barrier_released) break;
    if (b->barrier_count == b->live) {
      for (Thread& o : b->threads)
        if (o.at_barrier) {
          o.at_barrier = false;
          o.barrier_released = true;
        }
      b->barrier_count = 0;
      ++b->progress;
      break;
    }
    yield();
  }
  t->barrier_released = false;
}

void launch_raw(dim3 grid, dim3 block, size_t smem_bytes, void (*fn)(void*), void* arg) {
  if (blk) fatal("nested kernel launch");
  if (smem_bytes > kDynSmemBytes) fatal("dynamic shared memory request beyond the emulator's buffer");
  if (block.y != 1 || block.z != 1 || grid.y != 1 || grid.z != 1) fatal("only 1-D launches are emulated");
  const unsigned nthreads = block.x;
  if (nthreads == 0 || nthreads > 1024) fatal("bad block size");
  if (grid.x == 0) fatal("empty grid (cudaErrorInvalidConfiguration on a GPU)");
  Block b;
  b.threads.resize(nthreads);
  b.nthreads = nthreads;
  b.fn = fn;
  b.arg = arg;
  char* stacks = stacks_for(nthreads);
  ThreadCtx* saved_cur = cur;
  blk = &b;
  for (unsigned bx = 0; bx < grid.x; ++bx) {
    b.live = nthreads;
    b.barrier_count = 0;
    for (unsigned i = 0; i < nthreads; ++i) {
      Thread& t = b.threads[i];
      t.done = t.waiting = t.released = t.at_barrier = t.barrier_released = false;
      t.ctx.tid = {i, 0u, 0u};
      t.ctx.bid = {bx, 0u, 0u};
      t.ctx.bdim = block;
      t.ctx.gdim = grid;
      getcontext(&t.uc);
      t.uc.uc_stack.ss_sp = stacks + (size_t)i * kStackBytes;
      t.uc.uc_stack.ss_size = kStackBytes;
      t.uc.uc_link = &b.sched;
      makecontext(&t.uc, trampoline, 0);
    }
    // EMU_SCHEDULE=reverse | random[:seed] runs the threads of a block in another order; results that
    // depend on it point at a missing barrier (the default is ascending thread index)
    static const char* schedule = std::getenv("EMU_SCHEDULE");
    static thread_local uint64_t rng = 0;
    if (rng == 0) rng = (schedule && std::strchr(schedule, ':')) ? std::strtoull(std::strchr(schedule, ':') + 1, nullptr, 10) * 2654435761u + 88172645463325252ull : 88172645463325252ull;
    const bool reverse = schedule && !std::strncmp(schedule, "reverse", 7), shuffle = schedule && !std::strncmp(schedule, "random", 6);
    std::vector<unsigned> order(nthreads);
    for (unsigned i = 0; i < nthreads; ++i) order[i] = reverse ? nthreads - 1 - i : i;
    while (b.live > 0) {
      const uint64_t before = b.progress;
      if (shuffle)
        for (unsigned i = nthreads - 1; i > 0; --i) {
          rng ^= rng << 13; rng ^= rng >> 7; rng ^= rng << 17;
          std::swap(order[i], order[rng % (i + 1)]);
        }
      for (unsigned k = 0; k < nthreads; ++k) {
        const unsigned i = order[k];
        Thread& t = b.threads[i];
        if (t.done) continue;
        b.running = &t;
        cur = &t.ctx;
        swapcontext(&b.sched, &t.uc);
      }
      if (b.progress == before) {
        std::fprintf(stderr, "cuda_emu: deadlock in block %u: %u live threads, none can proceed\n", bx, b.live);
        for (unsigned i = 0; i < nthreads; ++i) {
          const Thread& t = b.threads[i];
          if (t.done) continue;
          std::fprintf(stderr, "  thread %u: %s op=%d mask=%08x\n", i, t.at_barrier ? "at __syncthreads" : (t.waiting ? "in warp collective" : "running"),
                       (int)t.op, t.mask);
        }
        std::abort();
      }
    }
  }
  b.running = nullptr;
  blk = nullptr;
  cur = saved_cur;
}

}  // namespace emu

// ================================================================================================
// Runtime API. Device memory = host memory, each allocation followed by an inaccessible guard page
// so that an overrun past the end faults instead of corrupting a neighbour. Fresh memory is
// filled with 0xA5 (a GPU does not zero cudaMalloc'ed memory either).
// ================================================================================================
namespace {

struct Allocation {
  void* user;
  void* map;
  size_t map_bytes;
};
std::vector<Allocation>& allocations() {
  static std::vector<Allocation> a;
  return a;
}
pthread_mutex_t g_alloc_lock = PTHREAD_MUTEX_INITIALIZER;

double now_ms() {
  timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
}

}  // namespace

struct emuStream {
  int dummy;
};
struct emuEvent {
  double ms;
};

cudaError_t cudaGetDeviceCount(int* n) {
  *n = 1;
  return cudaSuccess;
}
cudaError_t cudaSetDevice(int) { return cudaSuccess; }
cudaError_t cudaGetDeviceProperties(cudaDeviceProp* p, int) {
  std::memset(p, 0, sizeof(*p));
  p->multiProcessorCount = 4;  // small grids: persistent kernels loop more, which is what we want to test
  std::snprintf(p->name, sizeof(p->name), "cuda_emu (CPU, tests only)");
  return cudaSuccess;
}
cudaError_t cudaGetLastError() { return cudaSuccess; }
const char* cudaGetErrorString(cudaError_t e) { return e == cudaSuccess ? "no error" : (e == cudaErrorMemoryAllocation ? "out of memory" : "error"); }

#ifdef EMU_ASAN  // AddressSanitizer build: its red zones replace the guard page (both sides, and use after free)
cudaError_t cudaMalloc(void** p, size_t bytes) {
  void* m = nullptr;
  if (posix_memalign(&m, 256, bytes ? bytes : 1) != 0) return cudaErrorMemoryAllocation;
  std::memset(m, 0xA5, bytes);
  *p = m;
  return cudaSuccess;
}
cudaError_t cudaFree(void* p) {
  free(p);
  return cudaSuccess;
}
#else
cudaError_t cudaMalloc(void** p, size_t bytes) {
  const size_t page = (size_t)sysconf(_SC_PAGESIZE);
  const size_t need = (bytes + 15) & ~(size_t)15;
  const size_t data_pages = (need + page - 1) / page;
  const size_t map_bytes = (data_pages + 1) * page;
  void* m = mmap(nullptr, map_bytes, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
  if (m == MAP_FAILED) return cudaErrorMemoryAllocation;
  char* guard = static_cast<char*>(m) + data_pages * page;
  mprotect(guard, page, PROT_NONE);
  char* user = guard - need;  // the end of the buffer touches the guard page; 16-byte aligned
  std::memset(user, 0xA5, need);
  pthread_mutex_lock(&g_alloc_lock);
  allocations().push_back({user, m, map_bytes});
  pthread_mutex_unlock(&g_alloc_lock);
  *p = user;
  return cudaSuccess;
}
cudaError_t cudaFree(void* p) {
  if (!p) return cudaSuccess;
  pthread_mutex_lock(&g_alloc_lock);
  auto& a = allocations();
  for (size_t i = 0; i < a.size(); ++i)
    if (a[i].user == p) {
      munmap(a[i].map, a[i].map_bytes);
      a[i] = a.back();
      a.pop_back();
      pthread_mutex_unlock(&g_alloc_lock);
      return cudaSuccess;
    }
  pthread_mutex_unlock(&g_alloc_lock);
  std::fprintf(stderr, "cuda_emu: cudaFree of a pointer cudaMalloc did not return\n");
  std::abort();
}
#endif
cudaError_t cudaMemcpy(void* dst, const void* src, size_t bytes, cudaMemcpyKind) {
  std::memmove(dst, src, bytes);
  return cudaSuccess;
}
cudaError_t cudaMemcpyAsync(void* dst, const void* src, size_t bytes, cudaMemcpyKind, cudaStream_t) {
  std::memmove(dst, src, bytes);
  return cudaSuccess;
}
cudaError_t cudaMemset(void* dst, int v, size_t bytes) {
  std::memset(dst, v, bytes);
  return cudaSuccess;
}
cudaError_t cudaMemsetAsync(void* dst, int v, size_t bytes, cudaStream_t) {
  std::memset(dst, v, bytes);
  return cudaSuccess;
}
cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) {
  *s = new emuStream{0};
  return cudaSuccess;
}
cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
cudaError_t cudaStreamDestroy(cudaStream_t s) {
  delete s;
  return cudaSuccess;
}
cudaError_t cudaEventCreate(cudaEvent_t* e) {
  *e = new emuEvent{0.0};
  return cudaSuccess;
}
cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t) {
  e->ms = now_ms();
  return cudaSuccess;
}
cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t a, cudaEvent_t b) {
  *ms = (float)(b->ms - a->ms);
  return cudaSuccess;
}
cudaError_t cudaEventDestroy(cudaEvent_t e) {
  delete e;
  return cudaSuccess;
}

#include <chrono>
long long clock64() {
  return std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
