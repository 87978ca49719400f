// Micro-benchmark 2: what bounds small-N tcgen05.mma kind::tf32 (M128 K8, A in TMEM)?  (a) dependency through the
// accumulator -> round-robin over NACC accumulators; (b) the issuing thread -> NISS warps issue concurrently.
// build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I neural_pde_surrogates_b200/csrc -o tools/_bin/ubench_mma2 tools/ubench_mma2.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "pdes_ptx.cuh"
using namespace pdes;

template <int N, int NACC, int NISS, bool A_TMEM>
__global__ void __launch_bounds__(128, 1) k(int niter, long long* cycles) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) unsigned long long bar[4];
  __shared__ uint32_t slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < (128 * 8 + 256 * 8); i += 128) reinterpret_cast<float*>(smem)[i] = 0.001f * (i % 97);
  if (tid == 0) { for (int i = 0; i < 4; ++i) ptx::mbar_init(&bar[i], 1); ptx::fence_mbar_init(); }
  if (warp == 0) { ptx::tmem_alloc(&slot, 512); ptx::tmem_relinquish(); }
  ptx::fence_proxy_async();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tm = slot;
  if (warp < NISS && lane == 0) {
    const uint32_t idesc = ptx::idesc_tf32(128, N);
    const uint64_t da = ptx::smem_desc_noswizzle(ptx::smem_u32(smem), (128 / 8) * 128, 128);
    const uint64_t db = ptx::smem_desc_noswizzle(ptx::smem_u32(smem) + 128 * 8 * 4, (uint32_t)(N / 8) * 128, 128);
    const uint32_t d0 = tm + (uint32_t)(warp * NACC * N);
    uint32_t ph = 0;
    const long long t0 = clock64();
    for (int it = 0; it < niter; ++it) {
#pragma unroll
      for (int j = 0; j < 48; ++j) {
        if (A_TMEM) ptx::mma_tf32_ta(d0 + (uint32_t)((j % NACC) * N), tm + 448 + (uint32_t)((j % 4) * 8), db, idesc, 1u);
        else ptx::mma_tf32(d0 + (uint32_t)((j % NACC) * N), da, db, idesc, 1u);
      }
      ptx::tc_commit(&bar[warp]);
      ptx::mbar_wait(&bar[warp], ph);
      ph ^= 1;
    }
    const long long t1 = clock64();
    if (blockIdx.x == 0 && warp == 0) cycles[0] = t1 - t0;
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) { ptx::tc_fence_after(); ptx::tmem_dealloc(tm, 512); }
}

template <int N, int NACC, int NISS, bool A_TMEM>
void run(long long* d) {
  const int niter = 40;
  auto kk = k<N, NACC, NISS, A_TMEM>;
  cudaFuncSetAttribute(kk, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  kk<<<148, 128, 64 * 1024>>>(niter, d);
  long long c = 0; cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
  printf("A %s N=%3d  %d accumulators/issuer, %d issuing warps: %7.1f cycles per MMA per issuer, %7.1f cycles per MMA overall (%s)\n",
         A_TMEM ? "TMEM" : "smem", N, NACC, NISS, (double)c / (niter * 48), (double)c / (niter * 48 * NISS), cudaGetErrorString(cudaGetLastError()));
}

int main() {
  long long* d; cudaMalloc(&d, 8);
  run<32, 1, 1, true>(d); run<32, 2, 1, true>(d); run<32, 4, 1, true>(d);
  run<32, 1, 2, true>(d); run<32, 2, 2, true>(d); run<32, 1, 4, true>(d);
  run<32, 1, 1, false>(d); run<32, 2, 1, false>(d); run<32, 1, 2, false>(d); run<32, 1, 4, false>(d);
  run<64, 1, 1, true>(d); run<64, 2, 1, true>(d); run<64, 1, 2, true>(d);
  run<96, 1, 1, true>(d); run<96, 2, 1, true>(d); run<96, 1, 2, true>(d);
  run<192, 1, 1, true>(d); run<192, 2, 1, true>(d); run<192, 1, 2, true>(d); run<192, 1, 1, false>(d);
  run<256, 1, 1, true>(d); run<256, 1, 1, false>(d);
  return 0;
}
