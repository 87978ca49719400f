// Shared helpers for the sm_100a kernels of the U-FNO spectral block.
#pragma once

#include "pdes_emu.h"
#ifndef PDES_CPU_EMU
#include <cuda_runtime.h>
#include <cstdlib>
#define PDES_DYN_SMEM(T, name)                                      \
  extern __shared__ __align__(16) unsigned char _pdes_dyn_smem[];   \
  T* name = reinterpret_cast<T*>(_pdes_dyn_smem)
#define PDES_LAUNCH(kernel, grid, block, smem, stream, ...) \
  kernel<<<(grid), (block), (smem), (cudaStream_t)(stream)>>>(__VA_ARGS__)
#define PDES_SET_SMEM(kernel, bytes) \
  cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(bytes))
// Programmatic dependent launch: the kernel may be scheduled while its predecessor in the stream is still draining (after
// every predecessor CTA executed griddepcontrol.launch_dependents or exited); it MUST execute PDES_GRID_DEP_WAIT() before
// its first read of anything an earlier kernel wrote.  Hides the launch latency and the prologue (barrier init, TMEM
// allocation, table loads) of the chain kernels behind the tail of the previous one.  PDES_NO_PDL=1 restores plain launches.
namespace pdes {
inline bool pdl_enabled() {
  static const bool on = getenv("PDES_NO_PDL") == nullptr;
  return on;
}
template <class K, class... A>
inline void launch_pdl(K kernel, dim3 grid, dim3 block, size_t smem, cudaStream_t stream, A... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  cudaLaunchKernelEx(&cfg, kernel, args...);
}
}  // namespace pdes
#define PDES_LAUNCH_PDL(kernel, grid, block, smem, stream, ...) \
  pdes::launch_pdl(kernel, (grid), (block), (smem), (cudaStream_t)(stream), __VA_ARGS__)
#define PDES_GRID_DEP_WAIT() asm volatile("griddepcontrol.wait;" ::: "memory")
#define PDES_GRID_DEP_LAUNCH() asm volatile("griddepcontrol.launch_dependents;" ::: "memory")
// Ask for the maximum shared-memory carve-out for a kernel of the block chain (once per call site).  The tcgen05 kernels
// need ~220 KB of shared memory; a neighbour that runs with the default (small) carve-out forces the SMs to drain and
// switch their L1 / shared-memory split at every kernel boundary of the chain.
#define PDES_MAX_CARVEOUT(kernel)                                                                      \
  do {                                                                                                 \
    static bool _pdes_done = false;                                                                    \
    if (!_pdes_done) {                                                                                 \
      cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared); \
      _pdes_done = true;                                                                               \
    }                                                                                                  \
  } while (0)
#endif

#include <cstddef>
#include <cstdint>
#include <cstdio>

#include "../../include/pdes_b200.h"

namespace pdes {

constexpr int kMaxDynSmem = 200 * 1024;   // leave headroom below the 227 KB sm_100a CTA limit

// ---- error plumbing ---------------------------------------------------------------------------------
void set_error(const char* fmt, ...);
int check_launch(const char* what);

#define PDES_REQUIRE(cond, code, ...)         \
  do {                                        \
    if (!(cond)) {                            \
      pdes::set_error(__VA_ARGS__);           \
      return (code);                          \
    }                                         \
  } while (0)

// driver entry point cuTensorMapEncodeTiled fetched through the runtime (no libcuda link); nullptr if unavailable
void* tensor_map_encoder();
// K1 on tcgen05 (spectral_dft_tc.cu): -1 = shape / build outside that kernel, otherwise a PDES_* code
int dft_fwd_tc_try(const float* x0, int C0, const float* x1, int C1, int B, int H, int W, int m1, int m2, const float* tables,
                   int herm_scale, float* X, float* X2, int CinP, void* stream);
// TMA-fed K2 (spectral_mix_tma.cu): reduction splits it wants (0 = unsupported shape) and the launcher
// (PDES_ERR_UNSUPPORTED = caller falls back to the generic kernels)
int mix_tma_splits(int B, int Cred, int Cn, int m1, int m2);
int mix_tma_launch(bool conj, const float* Xin, const float* w1, const float* w2, float* P, int nsplit, int B, int Cred,
                   int Cn, int Cw_in, int Cw_out, int m1, int m2, int H, void* stream);

int mix_dw_tma_launch(const float* X, const float* GO, float* gw1, float* gw2, int B, int Cin, int Cout, int m1, int m2,
                      int H, void* stream);

// ---- twiddle table blob layout (floats) ---------------------------------------------------------------
// twh    [H][2]        (cos, sin)(2 pi j / H)
// twa    [W][NC4]      K1 stage A: col 2l -> cos(2 pi l w / W), col 2l+1 -> -sin(2 pi l w / W), zero padded
// tinv_f [2*m2][W]     K3b forward:  row 2l -> s_l cos(2 pi l w / W), row 2l+1 -> -s_l sin(..), s_l = c_l/(HW)
// tinv_b [2*m2][W]     K3b backward: same with s_l = 1
// herm   [m2]          s_l
// twp    [m1+1][npp][2] K3a (v2): (cos, sin)(2 pi j p / H) for the folded row pairs p = 0 .. H/2 (npp = pairs rounded up to 8)
// tw2    [m2][W+1][2]  K1 fast path stage 2: (cos, sin)(2 pi l w / W), row padded by one entry (bank spread)
// (order in the blob: twh, twp, tw2, twa, tinv_f, tinv_b, herm)
struct TableLayout {
  int nc4, npp;
  size_t twh, twa, tinv_f, tinv_b, herm, twp, tw2, total;
};
__host__ __device__ inline size_t round4(size_t n) { return (n + 3) & ~size_t(3); }
__host__ __device__ inline TableLayout table_layout(int H, int W, int m1, int m2) {
  TableLayout t;
  t.nc4 = (int)round4((size_t)2 * m2);
  t.npp = (H / 2 + 1 + 7) / 8 * 8;
  t.twh = 0;
  t.twp = t.twh + round4((size_t)2 * H);                       // (offsets of twh and twp do not depend on W)
  t.tw2 = t.twp + round4((size_t)(m1 + 1) * t.npp * 2);        // [m2][W+1] (cos, sin)(2 pi l w / W): K1 fast path, one bulk copy
  t.twa = t.tw2 + round4((size_t)m2 * (W + 1) * 2);
  t.tinv_f = t.twa + (size_t)W * t.nc4;
  t.tinv_b = t.tinv_f + round4((size_t)2 * m2 * W);
  t.herm = t.tinv_b + round4((size_t)2 * m2 * W);
  t.total = t.herm + round4((size_t)m2);
  return t;
}

// retained row k -> frequency index (proc_fno.py:266-269)
__host__ __device__ inline int kx_of(int k, int m1, int H) { return k < m1 ? k : H - 2 * m1 + k; }
// first-block rows overwritten by the second block when 2*m1 > H
__host__ __device__ inline bool row_dead(int k, int m1, int H) { return k < m1 && k >= H - m1; }

// geometry of the packed spectral weights Wp[m][tile][chunk][row][16 i][re|im] (spectral_mix_tc.cu, spectral_mix_adj_tc.cu)
constexpr int kMtBK = 16;          // input channels per chunk
constexpr int kMtMaxN = 64;        // padded 2B (two split accumulators x double buffering = 4 * N <= 256 TMEM columns)
__host__ __device__ inline int mt_cinp(int Cin) { return (Cin + kMtBK - 1) / kMtBK * kMtBK; }
__host__ __device__ inline int mt_npad(int B) { int n = (2 * B + 15) & ~15; return n < 16 ? 16 : n; }
__host__ __device__ inline int mt_ntile(int Cout) { return (Cout + 127) / 128; }
__host__ __device__ inline int mt_to(int Cout) { const int nt = mt_ntile(Cout); return (((Cout + nt - 1) / nt) + 7) & ~7; }

__host__ __device__ inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
__host__ __device__ inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

// exact GELU and its derivative (nn.GELU() default, approximate='none')
__device__ __forceinline__ float gelu_f(float v) { return 0.5f * v * (1.0f + erff(v * 0.70710678118654752440f)); }
// GELU with erf from Abramowitz-Stegun 7.1.26 (|erf error| <= 6e-7 in fp32 => GELU rel. L2 error ~1e-7, two orders
// below the 1e-5 parity budget) in ~16 instructions instead of ~34: used where the epilogue is issue-bound.
__device__ __forceinline__ float gelu_fast_f(float v) {
  const float x = v * 0.70710678118654752440f;
  const float ax = fabsf(x);
#ifdef PDES_CPU_EMU
  const float t = 1.0f / fmaf(0.3275911f, ax, 1.0f);
  const float ex = expf(-ax * ax);
#else
  const float t = __fdividef(1.0f, fmaf(0.3275911f, ax, 1.0f));
  const float ex = __expf(-ax * ax);
#endif
  float poly = fmaf(1.061405429f, t, -1.453152027f);
  poly = fmaf(poly, t, 1.421413741f);
  poly = fmaf(poly, t, -0.284496736f);
  poly = fmaf(poly, t, 0.254829592f);
  poly *= t;
  const float e = 1.0f - poly * ex;
  return 0.5f * v * (1.0f + copysignf(e, x));
}
// acc += a * b on both halves with ONE instruction (fma.rn.f32x2, SASS FFMA2, new on sm_100).  A complex MAC
// acc += x * w is two of them: ffma2(acc, (x.re, x.re), w); ffma2(acc, (x.im, x.im), (-w.im, w.re)); the scalar
// broadcasts are free operand modifiers.  Same flop rate as FFMA but half the issue slots.
__device__ __forceinline__ void ffma2(float2& acc, float2 a, float2 b) {
#ifdef PDES_CPU_EMU
  acc.x = fmaf(a.x, b.x, acc.x);
  acc.y = fmaf(a.y, b.y, acc.y);
#else
  unsigned long long d = *reinterpret_cast<unsigned long long*>(&acc);
  const unsigned long long aa = *reinterpret_cast<unsigned long long*>(&a), bb = *reinterpret_cast<unsigned long long*>(&b);
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(d) : "l"(aa), "l"(bb));
  acc = *reinterpret_cast<float2*>(&d);
#endif
}
// two GELUs at once (same Abramowitz-Stegun erf as gelu_fast_f): the polynomial and the affine steps are packed
// FFMA2 / FMUL2, only the two MUFU pairs (rcp, ex2) and the sign handling stay scalar: ~10 instructions per output
// instead of ~16.
__device__ __forceinline__ float2 gelu_fast2_f(float2 v) {
#ifdef PDES_CPU_EMU
  return make_float2(gelu_fast_f(v.x), gelu_fast_f(v.y));
#else
  const float2 zero = make_float2(0.0f, 0.0f);
  float2 x = zero;
  ffma2(x, v, make_float2(0.70710678118654752440f, 0.70710678118654752440f));
  const float2 ax = make_float2(fabsf(x.x), fabsf(x.y));
  float2 den = make_float2(1.0f, 1.0f);
  ffma2(den, make_float2(0.3275911f, 0.3275911f), ax);
  float2 t, ex;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t.x) : "f"(den.x));                   // den >= 1: no range handling needed
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t.y) : "f"(den.y));
  float2 q = zero;
  ffma2(q, make_float2(-1.4426950408889634f * ax.x, -1.4426950408889634f * ax.y), ax);   // -ax^2 * log2(e)
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(ex.x) : "f"(q.x));                     // q <= 0: underflow to 0 is the right limit
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(ex.y) : "f"(q.y));
  float2 poly = make_float2(-1.453152027f, -1.453152027f);
  ffma2(poly, make_float2(1.061405429f, 1.061405429f), t);
  float2 p2 = make_float2(1.421413741f, 1.421413741f);
  ffma2(p2, poly, t);
  float2 p3 = make_float2(-0.284496736f, -0.284496736f);
  ffma2(p3, p2, t);
  float2 p4 = make_float2(0.254829592f, 0.254829592f);
  ffma2(p4, p3, t);
  float2 pe = zero;
  ffma2(pe, p4, t);                                                             // poly * t
  float2 e = make_float2(1.0f, 1.0f);
  ffma2(e, make_float2(-pe.x, -pe.y), ex);                                      // 1 - poly * exp(-x^2)
  const float2 sgn = make_float2(copysignf(e.x, x.x), copysignf(e.y, x.y));
  float2 hv = zero;
  ffma2(hv, v, make_float2(0.5f, 0.5f));
  float2 r = hv;
  ffma2(r, hv, sgn);                                                            // 0.5 v (1 + erf)
  return r;
#endif
}
__device__ __forceinline__ float gelu_grad_f(float v) {
  const float cdf = 0.5f * (1.0f + erff(v * 0.70710678118654752440f));
  const float pdf = 0.39894228040143267794f * expf(-0.5f * v * v);
  return cdf + v * pdf;
}

}  // namespace pdes
