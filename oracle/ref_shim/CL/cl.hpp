/*
 * CL/cl.hpp -- a minimal in-process stand-in for the Khronos OpenCL C++ bindings
 * (TEST INFRASTRUCTURE, see oracle/oracle.h).
 *
 * This image has no OpenCL headers, ICD or CPU device, so the reference cannot be built
 * against a real runtime. This header provides exactly the slice of `cl::` that the reference
 * host code uses (libclsph/sph_simulation.cpp, util/cl_boilerplate.cpp) so those files compile
 * UNMODIFIED with g++; the "device" it exposes executes the reference's own kernels,
 * compiled natively from libclsph/kernels/ by oracle/build_ref.sh (see cl_runtime.cpp).
 *
 * One platform ("clsph-ref-shim"), one CPU device. Buffers are host allocations; queues are
 * in-order and synchronous; enqueueNDRangeKernel runs work-groups across OpenMP threads.
 */
#ifndef CLSPH_REF_SHIM_CL_HPP_
#define CLSPH_REF_SHIM_CL_HPP_

#include <array>
#include <cassert>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <functional>
#include <iostream>
#include <limits>
#include <memory>
#include <sstream>
#include <string>
#include <utility>
#include <vector>

/* ---- cl_platform.h slice --------------------------------------------------------------- */
typedef int32_t cl_int;
typedef uint32_t cl_uint;
typedef uint64_t cl_ulong;
typedef float cl_float;
typedef cl_uint cl_bool;
typedef cl_ulong cl_bitfield;
typedef cl_bitfield cl_mem_flags;
typedef cl_bitfield cl_device_type;
typedef cl_bitfield cl_command_queue_properties;

typedef union {
  cl_float __attribute__((aligned(16))) s[4];
  __extension__ struct { cl_float x, y, z, w; };
} __attribute__((aligned(16))) cl_float4;
typedef cl_float4 cl_float3;
typedef union {
  cl_uint __attribute__((aligned(16))) s[4];
  __extension__ struct { cl_uint x, y, z, w; };
} __attribute__((aligned(16))) cl_uint4;
typedef cl_uint4 cl_uint3;

/* ---- cl.h constants used by the reference ---------------------------------------------- */
#define CL_SUCCESS 0
#define CL_INVALID_VALUE -30
#define CL_INVALID_PLATFORM -32
#define CL_INVALID_DEVICE -33
#define CL_INVALID_PROGRAM -44
#define CL_INVALID_KERNEL_NAME -46
#define CL_INVALID_ARG_INDEX -49
#define CL_FALSE 0
#define CL_TRUE 1
#define CL_MEM_READ_WRITE (1 << 0)
#define CL_MEM_ALLOC_HOST_PTR (1 << 4)
#define CL_DEVICE_TYPE_ALL 0xFFFFFFFF
#define CL_PLATFORM_NAME 0x0902
#define CL_DEVICE_NAME 0x102B
#define CL_DEVICE_MAX_WORK_GROUP_SIZE 0x1004
#define CL_DEVICE_LOCAL_MEM_SIZE 0x1023
#define CL_PROGRAM_BUILD_LOG 0x1183

namespace cl {

namespace shim {
/* Largest work-group and local memory of the shim device. 4096 x 80 B = 320 KiB must fit
 * (libclsph/sph_simulation.cpp:189-192), as it does on PoCL's CPU device. */
const size_t kMaxWorkGroupSize = 4096;
const cl_ulong kLocalMemSize = 1u << 20;

template <int Name> struct info;
template <> struct info<CL_PLATFORM_NAME> { typedef std::string type; };
template <> struct info<CL_DEVICE_NAME> { typedef std::string type; };
template <> struct info<CL_DEVICE_MAX_WORK_GROUP_SIZE> { typedef size_t type; };
template <> struct info<CL_DEVICE_LOCAL_MEM_SIZE> { typedef cl_ulong type; };
template <> struct info<CL_PROGRAM_BUILD_LOG> { typedef std::string type; };

/* One kernel argument as the host set it. */
struct arg_slot {
  enum kind_t { UNSET, BYTES, BUFFER, LOCAL } kind = UNSET;
  std::vector<unsigned char> bytes;   /* BYTES: by-value argument                    */
  std::shared_ptr<std::vector<unsigned char>> buffer; /* BUFFER: backing store      */
  size_t local_bytes = 0;             /* LOCAL: per-work-group scratch size          */
};

/* Implemented in cl_runtime.cpp: runs kernel `name` over [0, global) in groups of `local`. */
cl_int launch(const std::string& name, const std::vector<arg_slot>& args, size_t global, size_t local);
bool kernel_exists(const std::string& name);
}  // namespace shim

class Device {
 public:
  template <int Name>
  typename shim::info<Name>::type getInfo() const { return get(std::integral_constant<int, Name>()); }

 private:
  std::string get(std::integral_constant<int, CL_DEVICE_NAME>) const { return "clsph-ref-shim CPU (g++ native kernels)"; }
  size_t get(std::integral_constant<int, CL_DEVICE_MAX_WORK_GROUP_SIZE>) const { return shim::kMaxWorkGroupSize; }
  cl_ulong get(std::integral_constant<int, CL_DEVICE_LOCAL_MEM_SIZE>) const { return shim::kLocalMemSize; }
};

class Platform {
 public:
  static cl_int get(std::vector<Platform>* out) {
    out->assign(1, Platform());
    return CL_SUCCESS;
  }
  template <int Name>
  typename shim::info<Name>::type getInfo() const { return "clsph-ref-shim"; }
  cl_int getDevices(cl_device_type, std::vector<Device>* out) const {
    out->assign(1, Device());
    return CL_SUCCESS;
  }
};

class Context {
 public:
  Context() {}
  Context(const std::vector<Device>&, void* = NULL, void* = NULL, void* = NULL, cl_int* err = NULL) {
    if (err) *err = CL_SUCCESS;
  }
};

class Buffer {
 public:
  Buffer() {}
  Buffer(const Context&, cl_mem_flags, size_t size, void* = NULL, cl_int* err = NULL)
      : store_(std::make_shared<std::vector<unsigned char>>(size)) {
    if (err) *err = CL_SUCCESS;
  }
  std::shared_ptr<std::vector<unsigned char>> store_;
};

class NDRange {
 public:
  NDRange() : size_(0), null_(true) {}
  NDRange(size_t n) : size_(n), null_(false) {}
  size_t size_;
  bool null_;
};
static const NDRange NullRange;

class Program {
 public:
  typedef std::vector<std::pair<const char*, size_t>> Sources;
  Program() {}
  Program(const Context&, const Sources&, cl_int* err = NULL) {
    if (err) *err = CL_SUCCESS;
  }
  /* Nothing to JIT: the kernels were compiled from the same sources ahead of time. */
  cl_int build(const std::vector<Device>&, const char* = NULL) const { return CL_SUCCESS; }
  template <int Name>
  typename shim::info<Name>::type getBuildInfo(const Device&) const {
    return "clsph-ref-shim: kernels precompiled by oracle/build_ref.sh";
  }
};

class Kernel {
 public:
  Kernel() {}
  Kernel(const Program&, const char* name, cl_int* err = NULL) : name_(name) {
    if (err) *err = shim::kernel_exists(name_) ? CL_SUCCESS : CL_INVALID_KERNEL_NAME;
  }
  cl_int setArg(cl_uint index, const Buffer& b) {
    shim::arg_slot& s = slot(index);
    s.kind = shim::arg_slot::BUFFER;
    s.buffer = b.store_;
    return CL_SUCCESS;
  }
  template <typename T>
  cl_int setArg(cl_uint index, const T& value) {
    shim::arg_slot& s = slot(index);
    s.kind = shim::arg_slot::BYTES;
    s.bytes.resize(sizeof(T));
    std::memcpy(s.bytes.data(), &value, sizeof(T));
    return CL_SUCCESS;
  }
  cl_int setArg(cl_uint index, size_t size, const void* ptr) {
    shim::arg_slot& s = slot(index);
    if (ptr == NULL) {
      s.kind = shim::arg_slot::LOCAL;
      s.local_bytes = size;
    } else {
      s.kind = shim::arg_slot::BYTES;
      s.bytes.assign((const unsigned char*)ptr, (const unsigned char*)ptr + size);
    }
    return CL_SUCCESS;
  }
  std::string name_;
  std::vector<shim::arg_slot> args_;

 private:
  shim::arg_slot& slot(cl_uint index) {
    if (args_.size() <= index) args_.resize(index + 1);
    return args_[index];
  }
};

class CommandQueue {
 public:
  CommandQueue() {}
  CommandQueue(const Context&, const Device&, cl_command_queue_properties = 0, cl_int* err = NULL) {
    if (err) *err = CL_SUCCESS;
  }
  template <typename P>
  cl_int enqueueFillBuffer(const Buffer& b, P pattern, size_t offset, size_t size) const {
    if (!b.store_ || offset + size > b.store_->size()) return CL_INVALID_VALUE;
    for (size_t at = 0; at + sizeof(P) <= size; at += sizeof(P))
      std::memcpy(b.store_->data() + offset + at, &pattern, sizeof(P));
    return CL_SUCCESS;
  }
  cl_int enqueueReadBuffer(const Buffer& b, cl_bool, size_t offset, size_t size, void* dst) const {
    if (!b.store_ || offset + size > b.store_->size()) return CL_INVALID_VALUE;
    std::memcpy(dst, b.store_->data() + offset, size);
    return CL_SUCCESS;
  }
  cl_int enqueueWriteBuffer(const Buffer& b, cl_bool, size_t offset, size_t size, const void* src) const {
    if (!b.store_ || offset + size > b.store_->size()) return CL_INVALID_VALUE;
    std::memcpy(b.store_->data() + offset, src, size);
    return CL_SUCCESS;
  }
  cl_int enqueueNDRangeKernel(const Kernel& k, const NDRange&, const NDRange& global,
                              const NDRange& local) const {
    return shim::launch(k.name_, k.args_, global.size_, local.null_ ? 0 : local.size_);
  }
};

}  // namespace cl

#endif
