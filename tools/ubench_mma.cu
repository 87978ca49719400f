// Micro-benchmark: sustained tcgen05.mma kind::tf32 rate per SM for the K3b tile shapes (M = 128, K = 8 per instruction),
// A from shared memory vs tensor memory.  One CTA per SM, one thread issues NITER x NPER MMAs into one accumulator.
// build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I neural_pde_surrogates_b200/csrc -o tools/_bin/ubench_mma tools/ubench_mma.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "pdes_ptx.cuh"
using namespace pdes;

__global__ void __launch_bounds__(128, 1) k(int N, int a_tmem, int nper, int niter, long long* cycles, int nacc = 1, int na = 1) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) unsigned long long bar;
  __shared__ uint32_t slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < (128 * 8 + 256 * 8) * 4 / 4; i += 128) reinterpret_cast<float*>(smem)[i] = 0.001f * (i % 97);
  if (tid == 0) { ptx::mbar_init(&bar, 1); ptx::fence_mbar_init(); }
  if (warp == 0) { ptx::tmem_alloc(&slot, 512); ptx::tmem_relinquish(); }
  ptx::fence_proxy_async();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tm = slot;
  if (tid == 0) {
    const uint32_t idesc = ptx::idesc_tf32(128, N);
    const uint64_t da = ptx::smem_desc_noswizzle(ptx::smem_u32(smem), (128 / 8) * 128, 128);
    const uint64_t db = ptx::smem_desc_noswizzle(ptx::smem_u32(smem) + 128 * 8 * 4, (uint32_t)(N / 8) * 128, 128);
    uint32_t ph = 0;
    const long long t0 = clock64();
    for (int it = 0; it < niter; ++it) {
      for (int j = 0; j < nper; ++j) {
        if (a_tmem) ptx::mma_tf32_ta(tm + (uint32_t)((j % nacc) * N), tm + 448 + (uint32_t)((j % na) * 8), db, idesc, 1u);
        else ptx::mma_tf32(tm, da, db, idesc, 1u);
      }
      ptx::tc_commit(&bar);
      ptx::mbar_wait(&bar, ph);
      ph ^= 1;
    }
    const long long t1 = clock64();
    if (blockIdx.x == 0) cycles[0] = t1 - t0;
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) { ptx::tc_fence_after(); ptx::tmem_dealloc(tm, 512); }
}

int main() {
  long long* d; cudaMalloc(&d, 8);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  for (int a_tmem = 0; a_tmem < 2; ++a_tmem)
    for (int N : {16, 32, 64, 128, 192, 208, 256})
      for (int nper : {6, 96}) {
        const int niter = 2000 / nper + 1;
        k<<<148, 128, 64 * 1024>>>(N, a_tmem, nper, niter, d);
        long long c = 0; cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
        const double per = (double)c / (niter * nper);
        printf("A from %s  N=%3d  %2d MMAs per commit+wait: %7.1f cycles/MMA  -> %6.0f MAC/clk/SM (%s)\n", a_tmem ? "TMEM" : "smem",
               N, nper, per, 128.0 * N * 8 / per, cudaGetErrorString(cudaGetLastError()));
      }
  // independent accumulators / A operands: is the ~100-cycle minimum a dependency through the accumulator?
  for (int N : {32, 64})
    for (int nacc : {1, 2, 4})
      for (int na : {1, 4}) {
        const int nper = 96, niter = 21;
        k<<<148, 128, 64 * 1024>>>(N, 1, nper, niter, d, nacc, na);
        long long c = 0; cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
        printf("A from TMEM  N=%3d  %d accumulators, %d A operands round-robin: %7.1f cycles/MMA (%s)\n", N, nacc, na,
               (double)c / (niter * nper), cudaGetErrorString(cudaGetLastError()));
      }
  return 0;
}
