// Minimal CPU emulation of the CUDA execution model (TEST INFRASTRUCTURE ONLY).
//
// The build container has nvcc but no GPU, and GPU minutes are scarce.  When the kernels in this
// directory are compiled with `g++ -x c++ -DPDES_CPU_EMU`, this header stands in for the CUDA
// runtime so that the *same kernel source* (index arithmetic, shared-memory staging, guards) can be
// run on host arrays through the same C ABI.  One std::thread per CUDA thread of a block, blocks run
// one after another, __syncthreads() is a pthread barrier.  The product library is never built this
// way: `_native.py` only loads the nvcc-built sm_100a library and raises if it is missing.
#pragma once
#ifdef PDES_CPU_EMU

#include <pthread.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __shared__ static
#define __launch_bounds__(...)
#define __align__(n) alignas(n)

struct float2 { float x, y; };
struct alignas(16) float4 { float x, y, z, w; };
static inline float2 make_float2(float x, float y) { return float2{x, y}; }
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }

struct uint3 { unsigned x, y, z; };
struct dim3 {
  unsigned x, y, z;
  dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};

typedef void* cudaStream_t;
typedef int cudaError_t;
enum { cudaSuccess = 0 };

namespace pdes_emu {
inline thread_local uint3 t_threadIdx, t_blockIdx;
inline dim3 g_blockDim, g_gridDim;
inline pthread_barrier_t g_barrier;
alignas(128) inline unsigned char g_dyn_smem[232448];

template <class F>
void launch(dim3 grid, dim3 block, size_t smem_bytes, F body) {
  if (smem_bytes > sizeof(g_dyn_smem)) { std::fprintf(stderr, "emu: smem too large\n"); std::abort(); }
  g_blockDim = block; g_gridDim = grid;
  const unsigned nt = block.x * block.y * block.z;
  pthread_barrier_init(&g_barrier, nullptr, nt);
  std::vector<std::thread> pool;
  pool.reserve(nt);
  for (unsigned t = 0; t < nt; ++t) {
    pool.emplace_back([=]() {
      t_threadIdx.x = t % block.x;
      t_threadIdx.y = (t / block.x) % block.y;
      t_threadIdx.z = t / (block.x * block.y);
      for (unsigned bz = 0; bz < grid.z; ++bz)
        for (unsigned by = 0; by < grid.y; ++by)
          for (unsigned bx = 0; bx < grid.x; ++bx) {
            t_blockIdx.x = bx; t_blockIdx.y = by; t_blockIdx.z = bz;
            body();
            pthread_barrier_wait(&g_barrier);   // next block reuses shared memory
          }
    });
  }
  for (auto& th : pool) th.join();
  pthread_barrier_destroy(&g_barrier);
}
}  // namespace pdes_emu

#define threadIdx (pdes_emu::t_threadIdx)
#define blockIdx (pdes_emu::t_blockIdx)
#define blockDim (pdes_emu::g_blockDim)
#define gridDim (pdes_emu::g_gridDim)

static inline void __syncthreads() { pthread_barrier_wait(&pdes_emu::g_barrier); }
template <class T> static inline T __ldg(const T* p) { return *p; }
static inline unsigned __float_as_uint(float f) { unsigned u; std::memcpy(&u, &f, 4); return u; }
static inline float __uint_as_float(unsigned u) { float f; std::memcpy(&f, &u, 4); return f; }
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline const char* cudaGetErrorString(cudaError_t) { return "emu"; }

#define PDES_DYN_SMEM(T, name) T* name = reinterpret_cast<T*>(pdes_emu::g_dyn_smem)
#define PDES_LAUNCH(kernel, grid, block, smem, stream, ...) \
  pdes_emu::launch((grid), (block), (smem), [&]() { kernel(__VA_ARGS__); })
#define PDES_SET_SMEM(kernel, bytes) (0)
#define PDES_MAX_CARVEOUT(kernel) do { } while (0)
#define PDES_LAUNCH_PDL PDES_LAUNCH
#define PDES_GRID_DEP_WAIT() do { } while (0)
#define PDES_GRID_DEP_LAUNCH() do { } while (0)

#endif  // PDES_CPU_EMU
