// pipe_rates.cu -- issue rates of the arithmetic the neighbour passes are built from, on the GPU at hand:
// FFMA vs the packed FFMA2 (two fp32 lanes per instruction, new in sm_100) vs HFMA2, and shared-memory loads.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/pipe_rates tools/micro/pipe_rates.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>

constexpr int kIters = 4096;

template <int kMode>
__global__ void __launch_bounds__(256) k_rate(float* out, float seed) {
  float a0 = seed + threadIdx.x, a1 = a0 + 1.f, a2 = a0 + 2.f, a3 = a0 + 3.f, a4 = a0 + 4.f, a5 = a0 + 5.f, a6 = a0 + 6.f, a7 = a0 + 7.f;
  const float m = 0.999f, c = 0.001f;
  if (kMode == 0) {  // 8 independent FFMA chains
#pragma unroll 1
    for (int k = 0; k < kIters; ++k) {
      a0 = fmaf(a0, m, c); a1 = fmaf(a1, m, c); a2 = fmaf(a2, m, c); a3 = fmaf(a3, m, c);
      a4 = fmaf(a4, m, c); a5 = fmaf(a5, m, c); a6 = fmaf(a6, m, c); a7 = fmaf(a7, m, c);
    }
  } else if (kMode == 1) {  // 4 independent FFMA2 chains (same number of fp32 lanes as mode 0)
    unsigned long long p0, p1, p2, p3, mm, cc;
    asm("mov.b64 %0, {%1, %2};" : "=l"(p0) : "f"(a0), "f"(a1));
    asm("mov.b64 %0, {%1, %2};" : "=l"(p1) : "f"(a2), "f"(a3));
    asm("mov.b64 %0, {%1, %2};" : "=l"(p2) : "f"(a4), "f"(a5));
    asm("mov.b64 %0, {%1, %2};" : "=l"(p3) : "f"(a6), "f"(a7));
    asm("mov.b64 %0, {%1, %1};" : "=l"(mm) : "f"(m));
    asm("mov.b64 %0, {%1, %1};" : "=l"(cc) : "f"(c));
#pragma unroll 1
    for (int k = 0; k < kIters; ++k) {
      asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p0) : "l"(mm), "l"(cc));
      asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p1) : "l"(mm), "l"(cc));
      asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p2) : "l"(mm), "l"(cc));
      asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p3) : "l"(mm), "l"(cc));
      asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p0) : "l"(mm), "l"(cc));
      asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p1) : "l"(mm), "l"(cc));
      asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p2) : "l"(mm), "l"(cc));
      asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p3) : "l"(mm), "l"(cc));
    }
    asm("mov.b64 {%0, %1}, %2;" : "=f"(a0), "=f"(a1) : "l"(p0));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(a2), "=f"(a3) : "l"(p1));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(a4), "=f"(a5) : "l"(p2));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(a6), "=f"(a7) : "l"(p3));
  } else if (kMode == 2) {  // 8 independent HFMA2 chains
    uint32_t h0 = __float_as_uint(a0), h1 = __float_as_uint(a1), h2 = __float_as_uint(a2), h3 = __float_as_uint(a3),
             h4 = __float_as_uint(a4), h5 = __float_as_uint(a5), h6 = __float_as_uint(a6), h7 = __float_as_uint(a7);
    const uint32_t hm = 0x3bff3bffu, hc = 0x14001400u;
#pragma unroll 1
    for (int k = 0; k < kIters; ++k) {
      asm volatile("fma.rn.f16x2 %0, %0, %1, %2;" : "+r"(h0) : "r"(hm), "r"(hc));
      asm volatile("fma.rn.f16x2 %0, %0, %1, %2;" : "+r"(h1) : "r"(hm), "r"(hc));
      asm volatile("fma.rn.f16x2 %0, %0, %1, %2;" : "+r"(h2) : "r"(hm), "r"(hc));
      asm volatile("fma.rn.f16x2 %0, %0, %1, %2;" : "+r"(h3) : "r"(hm), "r"(hc));
      asm volatile("fma.rn.f16x2 %0, %0, %1, %2;" : "+r"(h4) : "r"(hm), "r"(hc));
      asm volatile("fma.rn.f16x2 %0, %0, %1, %2;" : "+r"(h5) : "r"(hm), "r"(hc));
      asm volatile("fma.rn.f16x2 %0, %0, %1, %2;" : "+r"(h6) : "r"(hm), "r"(hc));
      asm volatile("fma.rn.f16x2 %0, %0, %1, %2;" : "+r"(h7) : "r"(hm), "r"(hc));
    }
    a0 = __uint_as_float(h0 ^ h1 ^ h2 ^ h3 ^ h4 ^ h5 ^ h6 ^ h7);
  } else if (kMode == 3) {  // 4 FFMA + 4 integer adds (alu pipe) interleaved: do the two pipes dual-issue?
    int i0 = threadIdx.x, i1 = 1, i2 = 2, i3 = 3;
#pragma unroll 1
    for (int k = 0; k < kIters; ++k) {
      a0 = fmaf(a0, m, c); i0 += i1; a1 = fmaf(a1, m, c); i1 += i2; a2 = fmaf(a2, m, c); i2 += i3; a3 = fmaf(a3, m, c); i3 += i0;
      a4 = fmaf(a4, m, c); i0 ^= i2; a5 = fmaf(a5, m, c); i1 ^= i3; a6 = fmaf(a6, m, c); i2 ^= i0; a7 = fmaf(a7, m, c); i3 ^= i1;
    }
    a0 += (float)(i0 + i1 + i2 + i3);
  } else if (kMode == 4) {  // 8 FADD2-style packed adds: add.f32x2
    unsigned long long p0, p1, p2, p3, cc;
    asm("mov.b64 %0, {%1, %2};" : "=l"(p0) : "f"(a0), "f"(a1));
    asm("mov.b64 %0, {%1, %2};" : "=l"(p1) : "f"(a2), "f"(a3));
    asm("mov.b64 %0, {%1, %2};" : "=l"(p2) : "f"(a4), "f"(a5));
    asm("mov.b64 %0, {%1, %2};" : "=l"(p3) : "f"(a6), "f"(a7));
    asm("mov.b64 %0, {%1, %1};" : "=l"(cc) : "f"(c));
#pragma unroll 1
    for (int k = 0; k < kIters; ++k) {
      asm volatile("add.f32x2 %0, %0, %1;" : "+l"(p0) : "l"(cc));
      asm volatile("add.f32x2 %0, %0, %1;" : "+l"(p1) : "l"(cc));
      asm volatile("add.f32x2 %0, %0, %1;" : "+l"(p2) : "l"(cc));
      asm volatile("add.f32x2 %0, %0, %1;" : "+l"(p3) : "l"(cc));
      asm volatile("add.f32x2 %0, %0, %1;" : "+l"(p0) : "l"(cc));
      asm volatile("add.f32x2 %0, %0, %1;" : "+l"(p1) : "l"(cc));
      asm volatile("add.f32x2 %0, %0, %1;" : "+l"(p2) : "l"(cc));
      asm volatile("add.f32x2 %0, %0, %1;" : "+l"(p3) : "l"(cc));
    }
    asm("mov.b64 {%0, %1}, %2;" : "=f"(a0), "=f"(a1) : "l"(p0));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(a2), "=f"(a3) : "l"(p1));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(a4), "=f"(a5) : "l"(p2));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(a6), "=f"(a7) : "l"(p3));
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

template <int kMode>
double run(const char* what, int lanes_per_inst, float* out) {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int blocks = sms * 8;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  k_rate<kMode><<<blocks, 256>>>(out, 1.f);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  k_rate<kMode><<<blocks, 256>>>(out, 1.f);
  cudaEventRecord(e1);
  cudaDeviceSynchronize();
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  const double warp_inst = (double)blocks * 8 /*warps*/ * kIters * 8;
  const double rate = warp_inst / (ms * 1e-3);  // warp instructions per second, whole GPU
  int khz = 0;
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  printf("%-34s %8.3f ms  %7.1f G warp-inst/s = %5.2f per SM sub-partition per cycle at %d MHz; %6.1f T lane-ops/s\n", what, ms, rate / 1e9,
         rate / (sms * 4.0 * khz * 1e3), khz / 1000, rate * 32 * lanes_per_inst / 1e12);
  return rate;
}

int main() {
  float* out;
  cudaMalloc(&out, 148 * 8 * 256 * 4 * 2);
  run<0>("FFMA (8 chains)", 1, out);
  run<1>("FFMA2 fma.rn.f32x2 (4 chains x 2)", 2, out);
  run<4>("FADD2 add.f32x2 (4 chains x 2)", 2, out);
  run<2>("HFMA2 fma.rn.f16x2 (8 chains)", 2, out);
  run<3>("FFMA + IADD/LOP interleaved", 1, out);
  cudaError_t e = cudaDeviceSynchronize();
  printf("status: %s\n", cudaGetErrorString(e));
  return e != cudaSuccess;
}
