/*
 * cl_device.h -- the OpenCL C "device side" that the reference's .cl files need, written as
 * plain C++ so g++ can compile those files where they lie (TEST INFRASTRUCTURE, see
 * oracle/oracle.h). It supplies what an OpenCL runtime would: vector types with their
 * operators, address-space and kernel qualifiers (erased), work-item id functions and the
 * built-in math the kernels call.
 *
 * The numeric definitions below are the arithmetic contract of oracle/oracle.h (OpenCL leaves
 * them to the runtime: libclsph picks "first platform, first device" and pins no version):
 * fused dot/length/distance, normalize = v / length(v), pown by sequential multiplies,
 * IEEE `/` and sqrt. The translation unit that includes this header is built with
 * -ffp-contract=off, so the reference's own expressions round after every operator.
 *
 * Work-items of one work-group run sequentially on one host thread in local-id order, which
 * is what makes async_work_group_copy (kernels/sph.cl:22-27) implementable as "item 0 copies".
 */
#ifndef CLSPH_REF_CL_DEVICE_H_
#define CLSPH_REF_CL_DEVICE_H_

#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstring>

namespace clsph_ref_device {

/* ---- work-item state (set by the launcher in cl_runtime.cpp) ------------------------- */
struct work_item_state {
  size_t global_id, group_id, local_id, local_size;
};
extern thread_local work_item_state g_wi;

inline size_t get_global_id(unsigned) { return g_wi.global_id; }
inline size_t get_group_id(unsigned) { return g_wi.group_id; }
inline size_t get_local_id(unsigned) { return g_wi.local_id; }
inline size_t get_local_size(unsigned) { return g_wi.local_size; }

/* ---- qualifiers ---------------------------------------------------------------------- */
#define kernel
#define global
#define __global
#define __local
#define constant

typedef unsigned int uint;

/* ---- vector types -------------------------------------------------------------------- */
/* float3 occupies 16 bytes, 16-byte aligned, like OpenCL's (and like the host's cl_float3),
 * so the structs in common/structures.h get the layout the host side writes. */
struct alignas(16) float3 {
  float x, y, z, pad_;
  float3() = default;
  float3(float a, float b, float c) : x(a), y(b), z(c), pad_(0.f) {}
  float3(float a) : x(a), y(a), z(a), pad_(0.f) {} /* scalar widening, smoothing.cl:24 */
};
struct alignas(16) uint3 {
  uint x, y, z, pad_;
  uint3() = default;
  uint3(uint a, uint b, uint c) : x(a), y(b), z(c), pad_(0u) {}
};
struct alignas(8) uint2 {
  uint x, y;
  uint2() = default;
  uint2(uint a, uint b) : x(a), y(b) {}
};

inline float3 operator+(float3 a, float3 b) { return float3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline float3 operator-(float3 a, float3 b) { return float3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline float3 operator-(float3 a) { return float3(-a.x, -a.y, -a.z); }
inline float3 operator*(float3 a, float s) { return float3(a.x * s, a.y * s, a.z * s); }
inline float3 operator*(float s, float3 a) { return float3(s * a.x, s * a.y, s * a.z); }
inline float3 operator/(float3 a, float s) { return float3(a.x / s, a.y / s, a.z / s); }
inline float3& operator+=(float3& a, float3 b) { return a = a + b; }
inline float3& operator-=(float3& a, float3 b) { return a = a - b; }

/* ---- built-ins ----------------------------------------------------------------------- */
inline float dot(float3 a, float3 b) { return std::fma(a.z, b.z, std::fma(a.y, b.y, a.x * b.x)); }
inline float length(float3 a) { return std::sqrt(dot(a, a)); }
inline float length(float a) { return std::fabs(a); }
inline float distance(float3 a, float3 b) { return length(a - b); }
inline float3 normalize(float3 a) { return a / length(a); }
inline float pown(float x, int n) {
  float r = x;
  for (int k = 1; k < n; ++k) r = r * x;
  return r;
}
inline float floor(float x) { return std::floor(x); }
inline float clamp(float x, float lo, float hi) { return std::fmin(std::fmax(x, lo), hi); }
template <typename T>
inline float convert_float(T v) { return static_cast<float>(v); }

/* ---- async copy: work-items of a group run in order on one thread, item 0 does it ---- */
typedef int event_t;
inline event_t async_work_group_copy(char* dst, const char* src, size_t n, event_t) {
  if (g_wi.local_id == 0) std::memcpy(dst, src, n);
  return 0;
}
inline void wait_group_events(int, event_t*) {}

}  // namespace clsph_ref_device

#endif
