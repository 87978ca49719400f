// Micro-benchmark: issue rate of scalar FFMA vs packed FFMA2 (fma.rn.f32x2) on sm_100a, per SM sub-partition.
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/_bin/ubench_ffma2 tools/ubench_ffma2.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void ffma2(float2& acc, float2 a, float2 b) {
  unsigned long long d = *reinterpret_cast<unsigned long long*>(&acc);
  const unsigned long long aa = *reinterpret_cast<unsigned long long*>(&a), bb = *reinterpret_cast<unsigned long long*>(&b);
  asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(d) : "l"(aa), "l"(bb));
  acc = *reinterpret_cast<float2*>(&d);
}

template <bool PACKED, int NACC>
__global__ void k(float2* out, float2 a, float2 b, int iters) {
  float2 acc[NACC];
#pragma unroll
  for (int i = 0; i < NACC; ++i) acc[i] = make_float2(threadIdx.x * 1e-3f + i, i * 0.5f);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) {
      if (PACKED) {
        ffma2(acc[i], a, b);
      } else {
        asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(acc[i].x) : "f"(a.x), "f"(b.x));
        asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(acc[i].y) : "f"(a.y), "f"(b.y));
      }
    }
  }
  float2 s = make_float2(0, 0);
#pragma unroll
  for (int i = 0; i < NACC; ++i) { s.x += acc[i].x; s.y += acc[i].y; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <bool PACKED>
void run(int warps, float2* out) {
  const int iters = 4096, NACC = 16;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<PACKED, NACC><<<148, warps * 32>>>(out, make_float2(1.0001f, 0.9999f), make_float2(0.5f, 0.25f), 16);
  cudaEventRecord(e0);
  k<PACKED, NACC><<<148, warps * 32>>>(out, make_float2(1.0001f, 0.9999f), make_float2(0.5f, 0.25f), iters);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  const double fma_lane = (double)148 * warps * 32 * iters * NACC * 2;     // scalar FMAs
  int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  printf("%s warps/SM %2d: %.3f ms  %.1f TFLOP/s  (%.1f FMA lanes/clk/SM at %d MHz nominal)\n", PACKED ? "FFMA2" : "FFMA ", warps, ms,
         2 * fma_lane / ms / 1e9, fma_lane / 148 / (ms * 1e-3 * clk * 1e3), clk / 1000);
}

int main() {
  float2* out; cudaMalloc(&out, 148 * 1024 * sizeof(float2));
  for (int w : {4, 8, 13, 16, 32}) { run<false>(w, out); run<true>(w, out); }
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
