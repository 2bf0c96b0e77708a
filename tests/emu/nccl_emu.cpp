// nccl_emu.cpp -- in-process NCCL stand-in for the emulator build (tests only). Ranks are OS
// threads of one process; sends are buffered (eager), receives block on a condition variable,
// all-reduce is a three-phase rendezvous. dist.cu reaches these through emu::fake_dlsym.
#include "nccl_emu.h"

#include <condition_variable>
#include <cstdio>
#include <deque>
#include <map>
#include <mutex>
#include <string>
#include <vector>

namespace {

struct Message {
  std::vector<char> bytes;
};

struct Group {
  int world = 0;
  int joined = 0;
  std::mutex m;
  std::condition_variable cv;
  std::map<std::pair<int, int>, std::deque<Message>> mail;  // (src, dst) -> queue
  // all-reduce rendezvous
  int ar_arrived = 0, ar_done = 0;
  uint64_t ar_generation = 0;
  std::vector<const void*> ar_src;
  std::vector<unsigned> ar_result;
};

std::mutex g_registry_lock;
std::map<unsigned long long, Group*> g_registry;
unsigned long long g_next_id = 1;

struct PendingRecv {
  void* dst;
  size_t bytes;
  int peer;
};
thread_local int g_group_depth = 0;
thread_local std::vector<PendingRecv> g_pending;

size_t dtype_bytes(ncclDataType_t t) { return (t == ncclInt8 || t == ncclUint8) ? 1 : 4; }

}  // namespace

struct emuNcclComm {
  Group* group;
  int rank, world;
};

namespace {

void do_recv(emuNcclComm* c, const PendingRecv& r) {
  Group* g = c->group;
  std::unique_lock<std::mutex> lk(g->m);
  auto key = std::make_pair(r.peer, c->rank);
  g->cv.wait(lk, [&] { return !g->mail[key].empty(); });
  Message msg = std::move(g->mail[key].front());
  g->mail[key].pop_front();
  lk.unlock();
  if (msg.bytes.size() != r.bytes) {
    std::fprintf(stderr, "nccl_emu: recv of %zu bytes matched a send of %zu bytes\n", r.bytes, msg.bytes.size());
    std::abort();
  }
  std::memcpy(r.dst, msg.bytes.data(), r.bytes);
}

ncclResult_t emuGetUniqueId(ncclUniqueId* id) {
  std::lock_guard<std::mutex> lk(g_registry_lock);
  std::memset(id, 0, sizeof(*id));
  const unsigned long long v = g_next_id++;
  std::memcpy(id->internal, &v, sizeof(v));
  return ncclSuccess;
}

ncclResult_t emuCommInitRank(ncclComm_t* comm, int world, ncclUniqueId id, int rank) {
  unsigned long long v;
  std::memcpy(&v, id.internal, sizeof(v));
  Group* g;
  {
    std::lock_guard<std::mutex> lk(g_registry_lock);
    Group*& slot = g_registry[v];
    if (!slot) {
      slot = new Group();
      slot->world = world;
      slot->ar_src.resize(world);
    }
    g = slot;
  }
  if (g->world != world || rank < 0 || rank >= world) return ncclInvalidArgument;
  std::unique_lock<std::mutex> lk(g->m);
  ++g->joined;
  g->cv.notify_all();
  g->cv.wait(lk, [&] { return g->joined >= g->world; });
  *comm = new emuNcclComm{g, rank, world};
  return ncclSuccess;
}

ncclResult_t emuCommDestroy(ncclComm_t comm) {
  delete comm;  // the group itself stays registered: peers may still be draining it
  return ncclSuccess;
}

ncclResult_t emuAllReduce(const void* send, void* recv, size_t count, ncclDataType_t dt, ncclRedOp_t op, ncclComm_t comm, cudaStream_t) {
  if (dt != ncclUint32 || (op != ncclMin && op != ncclMax)) return ncclInvalidArgument;
  Group* g = comm->group;
  std::unique_lock<std::mutex> lk(g->m);
  const uint64_t gen = g->ar_generation;
  g->ar_src[comm->rank] = send;
  if (++g->ar_arrived == g->world) {  // last to arrive reduces for everyone
    g->ar_result.assign(count, 0u);
    for (size_t k = 0; k < count; ++k) {
      unsigned acc = static_cast<const unsigned*>(g->ar_src[0])[k];
      for (int r = 1; r < g->world; ++r) {
        const unsigned v = static_cast<const unsigned*>(g->ar_src[r])[k];
        acc = op == ncclMin ? std::min(acc, v) : std::max(acc, v);
      }
      g->ar_result[k] = acc;
    }
    g->ar_arrived = 0;
    g->ar_done = 0;
    ++g->ar_generation;
    g->cv.notify_all();
  } else {
    g->cv.wait(lk, [&] { return g->ar_generation != gen; });
  }
  std::memcpy(recv, g->ar_result.data(), count * sizeof(unsigned));
  // nobody may start the next all-reduce (and overwrite ar_result) before everyone has copied
  if (++g->ar_done == g->world) g->cv.notify_all();
  else g->cv.wait(lk, [&] { return g->ar_done == g->world || g->ar_generation != gen + 1; });
  return ncclSuccess;
}

ncclResult_t emuSend(const void* src, size_t count, ncclDataType_t dt, int peer, ncclComm_t comm, cudaStream_t) {
  Group* g = comm->group;
  Message msg;
  msg.bytes.assign(static_cast<const char*>(src), static_cast<const char*>(src) + count * dtype_bytes(dt));
  std::lock_guard<std::mutex> lk(g->m);
  g->mail[std::make_pair(comm->rank, peer)].push_back(std::move(msg));
  g->cv.notify_all();
  return ncclSuccess;
}

thread_local ncclComm_t g_group_comm = nullptr;

ncclResult_t emuRecv(void* dst, size_t count, ncclDataType_t dt, int peer, ncclComm_t comm, cudaStream_t) {
  PendingRecv r{dst, count * dtype_bytes(dt), peer};
  if (g_group_depth > 0) {
    g_pending.push_back(r);
    g_group_comm = comm;
  } else {
    do_recv(comm, r);
  }
  return ncclSuccess;
}

ncclResult_t emuGroupStart() {
  ++g_group_depth;
  return ncclSuccess;
}

ncclResult_t emuGroupEnd() {
  if (--g_group_depth == 0) {
    for (const PendingRecv& r : g_pending) do_recv(g_group_comm, r);
    g_pending.clear();
  }
  return ncclSuccess;
}

const char* emuGetErrorString(ncclResult_t r) { return r == ncclSuccess ? "no error" : "nccl_emu error"; }

int g_fake_handle;

}  // namespace

namespace emu {

void* fake_dlopen(const char*, int) { return &g_fake_handle; }

void* fake_dlsym(void*, const char* name) {
  const std::string n(name);
  if (n == "ncclGetUniqueId") return (void*)&emuGetUniqueId;
  if (n == "ncclCommInitRank") return (void*)&emuCommInitRank;
  if (n == "ncclCommDestroy") return (void*)&emuCommDestroy;
  if (n == "ncclAllReduce") return (void*)&emuAllReduce;
  if (n == "ncclSend") return (void*)&emuSend;
  if (n == "ncclRecv") return (void*)&emuRecv;
  if (n == "ncclGroupStart") return (void*)&emuGroupStart;
  if (n == "ncclGroupEnd") return (void*)&emuGroupEnd;
  if (n == "ncclGetErrorString") return (void*)&emuGetErrorString;
  return nullptr;
}

}  // namespace emu
