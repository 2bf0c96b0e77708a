/*
 * cl_runtime.cpp -- the "device" of the OpenCL shim (TEST INFRASTRUCTURE, see
 * oracle/oracle.h): this translation unit textually includes the reference's kernel program
 * (libclsph/kernels/sph.cl, which pulls in every other .cl file and common/*.h) and
 * dispatches cl::CommandQueue::enqueueNDRangeKernel to those functions.
 *
 * The kernel sources are NOT part of this repository. oracle/build_ref.sh reads them from
 * the reference tree and writes errata-patched copies to oracle/_ref/gen/ (git-ignored);
 * REF_KERNEL_PROGRAM is the path of that generated sph.cl. The patches are the three
 * one-liners listed in build_ref.sh (E1, E2 and one OpenCL-only vector literal).
 *
 * Built with -ffp-contract=off so the reference's expressions are evaluated as written.
 */
#include "CL/cl.hpp"

#include <map>
#include <omp.h>

#include "cl_device.h"

#ifndef REF_KERNEL_PROGRAM
#error "REF_KERNEL_PROGRAM must point at the generated copy of libclsph/kernels/sph.cl"
#endif

namespace clsph_ref_device {
thread_local work_item_state g_wi;
#include REF_KERNEL_PROGRAM
}  // namespace clsph_ref_device

/* The .cl sources and cl_device.h leave these macros behind. */
#undef kernel
#undef global
#undef __global
#undef __local
#undef constant

namespace dev = clsph_ref_device;
using cl::shim::arg_slot;

namespace {

template <typename T>
T by_value(const arg_slot& a) {
  T v;
  assert(a.kind == arg_slot::BYTES && a.bytes.size() == sizeof(T));
  std::memcpy(&v, a.bytes.data(), sizeof(T));
  return v;
}
template <typename T>
T* in_buffer(const arg_slot& a) {
  assert(a.kind == arg_slot::BUFFER && a.buffer);
  return reinterpret_cast<T*>(a.buffer->data());
}

/* Argument order of each trampoline = the setArg order in libclsph/sph_simulation.cpp
 * (:257, :120-122, :145-147, :283-290, :322-323, :330-333). */
typedef std::function<void(void* local_mem)> work_item_body;
typedef std::function<work_item_body(const std::vector<arg_slot>&)> binder;

std::map<std::string, binder>& registry() {
  static std::map<std::string, binder> r;
  if (!r.empty()) return r;
  r["locate_in_grid"] = [](const std::vector<arg_slot>& a) -> work_item_body {
    auto in = in_buffer<const dev::particle>(a[0]);
    auto out = in_buffer<dev::particle>(a[1]);
    auto prm = by_value<dev::simulation_parameters>(a[2]);
    return [=](void*) { dev::locate_in_grid(in, out, prm); };
  };
  r["sort_count"] = [](const std::vector<arg_slot>& a) -> work_item_body {
    auto in = in_buffer<const dev::particle>(a[0]);
    auto counts = in_buffer<volatile unsigned int>(a[1]);
    auto prm = by_value<dev::simulation_parameters>(a[2]);
    int threads = by_value<int>(a[3]), pass = by_value<int>(a[4]), width = by_value<int>(a[5]);
    return [=](void*) { dev::sort_count(in, counts, prm, threads, pass, width); };
  };
  r["sort"] = [](const std::vector<arg_slot>& a) -> work_item_body {
    auto in = in_buffer<const dev::particle>(a[0]);
    auto out = in_buffer<dev::particle>(a[1]);
    auto starts = in_buffer<unsigned int>(a[2]);
    auto prm = by_value<dev::simulation_parameters>(a[3]);
    int threads = by_value<int>(a[4]), pass = by_value<int>(a[5]), width = by_value<int>(a[6]);
    return [=](void*) { dev::sort(in, out, starts, prm, threads, pass, width); };
  };
  r["density_pressure"] = [](const std::vector<arg_slot>& a) -> work_item_body {
    auto in = in_buffer<const dev::particle>(a[0]);
    auto out = in_buffer<dev::particle>(a[2]);
    auto prm = by_value<dev::simulation_parameters>(a[3]);
    auto terms = by_value<dev::precomputed_kernel_values>(a[4]);
    auto table = in_buffer<const unsigned int>(a[5]);
    return [=](void* lm) {
      dev::density_pressure(in, static_cast<dev::particle*>(lm), out, prm, terms, table);
    };
  };
  r["forces"] = [](const std::vector<arg_slot>& a) -> work_item_body {
    auto in = in_buffer<const dev::particle>(a[0]);
    auto out = in_buffer<dev::particle>(a[1]);
    auto prm = by_value<dev::simulation_parameters>(a[2]);
    auto terms = by_value<dev::precomputed_kernel_values>(a[3]);
    auto table = in_buffer<const unsigned int>(a[4]);
    return [=](void*) { dev::forces(in, out, prm, terms, table); };
  };
  r["advection_collision"] = [](const std::vector<arg_slot>& a) -> work_item_body {
    auto in = in_buffer<const dev::particle>(a[0]);
    auto out = in_buffer<dev::particle>(a[1]);
    auto prm = by_value<dev::simulation_parameters>(a[2]);
    auto terms = by_value<dev::precomputed_kernel_values>(a[3]);
    auto table = in_buffer<const unsigned int>(a[4]);
    auto normals = in_buffer<const float>(a[5]);
    auto vertices = in_buffer<const float>(a[6]);
    auto indices = in_buffer<const dev::uint>(a[7]);
    dev::uint faces = by_value<dev::uint>(a[8]);
    return [=](void*) {
      dev::advection_collision(in, out, prm, terms, table, normals, vertices, indices, faces);
    };
  };
  return r;
}

}  // namespace

namespace cl {
namespace shim {

bool kernel_exists(const std::string& name) { return registry().count(name) != 0; }

cl_int launch(const std::string& name, const std::vector<arg_slot>& args, size_t global, size_t local) {
  auto it = registry().find(name);
  if (it == registry().end()) return CL_INVALID_KERNEL_NAME;
  if (local == 0) local = 1; /* cl::NullRange: the implementation picks */
  if (local > kMaxWorkGroupSize || global % local != 0) return CL_INVALID_VALUE;
  size_t local_bytes = 0;
  for (const arg_slot& a : args) {
    if (a.kind == arg_slot::UNSET) return CL_INVALID_ARG_INDEX;
    if (a.kind == arg_slot::LOCAL) local_bytes = std::max(local_bytes, a.local_bytes);
  }
  if (local_bytes > kLocalMemSize) return CL_INVALID_VALUE;
  const work_item_body body = it->second(args);
  const long groups = static_cast<long>(global / local);
#pragma omp parallel
  {
    std::vector<unsigned char> local_mem(local_bytes ? local_bytes : 16);
#pragma omp for schedule(dynamic, 1)
    for (long g = 0; g < groups; ++g) {
      for (size_t l = 0; l < local; ++l) {
        dev::g_wi.group_id = static_cast<size_t>(g);
        dev::g_wi.local_id = l;
        dev::g_wi.local_size = local;
        dev::g_wi.global_id = static_cast<size_t>(g) * local + l;
        body(local_mem.data());
      }
    }
  }
  return CL_SUCCESS;
}

}  // namespace shim
}  // namespace cl
