// Fused per-pixel temporal decoder (SURVEY.md 8(f) "next #2"): the two Conv1d layers of the reference's TimeConvDense
// (dec_grid.py:97-146) applied along the 3*tw axis of every pixel,
//
//     h[o][l] = b1[o] + sum_k w1[o][k] * z[2l + k]            Conv1d(1 -> 2, k = ceil(tw/2), stride 2)
//     y[t]    = b2    + sum_{o,k} w2[o][k] * act(h[o][t + k])  Conv1d(2 -> 1, k = ceil(tw/4) + 1 (+1 if tw % 4 == 0))
//
// PyTorch runs them as cuDNN convolutions with batch = B*H*W = 98 304 and 1-2 channels after a permute copy (about
// 6 ms per training step, forward + backward).  Here thread = pixel: z is read straight from the [B][3tw][H*W] output of
// the 1x1 pre-decoder (coalesced along pixels, no permute), everything else stays in registers / shared memory, and the
// result is written in the [B][1][tw][H*W] layout the caller needs.  The backward recomputes h, produces dz and
// per-block partial sums of the 45 weight gradients, reduced in a fixed order by a second tiny kernel (deterministic).
#include "pdes_common.cuh"

namespace pdes {
namespace {

constexpr int kTvThreads = 128;

template <int TW>
struct TvGeom {
  static constexpr int NZ = 3 * TW;
  static constexpr int KA = (TW + 1) / 2;
  static constexpr int KB = (TW + 3) / 4 + 1 + (TW % 4 == 0 ? 1 : 0);
  static constexpr int L1 = (NZ - KA) / 2 + 1;
  static constexpr int NW = 2 * KA + 2 + 2 * KB + 1;      // w1, b1, w2, b2
  static_assert(L1 - KB + 1 == TW, "decoder geometry must map 3*tw inputs to tw outputs");
};

template <int TW>
__device__ __forceinline__ void tv_load_weights(float* sw, const float* w1, const float* b1, const float* w2, const float* b2) {
  using G = TvGeom<TW>;
  for (int i = threadIdx.x; i < G::NW; i += kTvThreads) {
    float v;
    if (i < 2 * G::KA) v = w1[i];
    else if (i < 2 * G::KA + 2) v = b1[i - 2 * G::KA];
    else if (i < 2 * G::KA + 2 + 2 * G::KB) v = w2[i - 2 * G::KA - 2];
    else v = b2[0];
    sw[i] = v;
  }
  __syncthreads();
}

template <int TW>
__global__ void __launch_bounds__(kTvThreads)
k_timeconv_fwd(const float* __restrict__ z, const float* __restrict__ w1, const float* __restrict__ b1,
               const float* __restrict__ w2, const float* __restrict__ b2, float* __restrict__ out, int HW, int act) {
  using G = TvGeom<TW>;
  __shared__ float sw[G::NW];
  tv_load_weights<TW>(sw, w1, b1, w2, b2);
  const float* sw1 = sw;
  const float* sb1 = sw + 2 * G::KA;
  const float* sw2 = sb1 + 2;
  const float sb2 = sw2[2 * G::KB];
  const int pix = blockIdx.x * kTvThreads + threadIdx.x;
  const int b = blockIdx.y;
  if (pix >= HW) return;
  const float* zp = z + (size_t)b * G::NZ * HW + pix;
  float zr[G::NZ];
#pragma unroll
  for (int j = 0; j < G::NZ; ++j) zr[j] = __ldg(zp + (size_t)j * HW);
  float y[TW];
#pragma unroll
  for (int t = 0; t < TW; ++t) y[t] = sb2;
#pragma unroll
  for (int o = 0; o < 2; ++o) {
#pragma unroll
    for (int l = 0; l < G::L1; ++l) {
      float h = sb1[o];
#pragma unroll
      for (int k = 0; k < G::KA; ++k) h = fmaf(sw1[o * G::KA + k], zr[2 * l + k], h);
      const float a = act == PDES_ACT_GELU ? gelu_f(h) : h;
#pragma unroll
      for (int k = 0; k < G::KB; ++k) {
        const int t = l - k;
        if (t >= 0 && t < TW) y[t] = fmaf(sw2[o * G::KB + k], a, y[t]);
      }
    }
  }
  float* op = out + (size_t)b * TW * HW + pix;
#pragma unroll
  for (int t = 0; t < TW; ++t) op[(size_t)t * HW] = y[t];
}

// dz, and per-block partial sums of (dw1, db1, dw2, db2) in part[block][NW]
template <int TW>
__global__ void __launch_bounds__(kTvThreads)
k_timeconv_bwd(const float* __restrict__ z, const float* __restrict__ gy, const float* __restrict__ w1,
               const float* __restrict__ b1, const float* __restrict__ w2, const float* __restrict__ b2,
               float* __restrict__ dz, float* __restrict__ part, int HW, int act) {
  using G = TvGeom<TW>;
  PDES_DYN_SMEM(float, dyn);                       // zs [NZ][T], dzs [NZ][T], gys [TW][T]  (column = thread: conflict-free)
  float* zs = dyn;
  float* dzs = zs + G::NZ * kTvThreads;
  float* gys = dzs + G::NZ * kTvThreads;
  __shared__ float sw[G::NW];
  tv_load_weights<TW>(sw, w1, b1, w2, b2);
  const float* sw1 = sw;
  const float* sb1 = sw + 2 * G::KA;
  const float* sw2 = sb1 + 2;
  const int tid = threadIdx.x;
  const int pix = blockIdx.x * kTvThreads + tid;
  const int b = blockIdx.y;
  const bool valid = pix < HW;
  const float* zp = z + (size_t)b * G::NZ * HW + pix;
  const float* gp = gy + (size_t)b * TW * HW + pix;
  for (int j = 0; j < G::NZ; ++j) {
    zs[j * kTvThreads + tid] = valid ? __ldg(zp + (size_t)j * HW) : 0.0f;
    dzs[j * kTvThreads + tid] = 0.0f;
  }
  float gsum = 0.0f;
  for (int t = 0; t < TW; ++t) {
    const float g = valid ? __ldg(gp + (size_t)t * HW) : 0.0f;
    gys[t * kTvThreads + tid] = g;
    gsum += g;
  }
  float dw1a[2][G::KA], db1a[2], dw2a[2][G::KB];
#pragma unroll
  for (int o = 0; o < 2; ++o) {
    db1a[o] = 0.0f;
#pragma unroll
    for (int k = 0; k < G::KA; ++k) dw1a[o][k] = 0.0f;
#pragma unroll
    for (int k = 0; k < G::KB; ++k) dw2a[o][k] = 0.0f;
  }
#pragma unroll
  for (int o = 0; o < 2; ++o) {
#pragma unroll 1
    for (int l = 0; l < G::L1; ++l) {
      float h = sb1[o];
#pragma unroll
      for (int k = 0; k < G::KA; ++k) h = fmaf(sw1[o * G::KA + k], zs[(2 * l + k) * kTvThreads + tid], h);
      const float a = act == PDES_ACT_GELU ? gelu_f(h) : h;
      float da = 0.0f;
#pragma unroll
      for (int k = 0; k < G::KB; ++k) {
        const int t = l - k;
        if (t >= 0 && t < TW) {
          const float g = gys[t * kTvThreads + tid];
          da = fmaf(sw2[o * G::KB + k], g, da);
          dw2a[o][k] = fmaf(g, a, dw2a[o][k]);
        }
      }
      const float dh = act == PDES_ACT_GELU ? da * gelu_grad_f(h) : da;
      db1a[o] += dh;
#pragma unroll
      for (int k = 0; k < G::KA; ++k) {
        const int j = (2 * l + k) * kTvThreads + tid;
        dw1a[o][k] = fmaf(dh, zs[j], dw1a[o][k]);
        dzs[j] = fmaf(sw1[o * G::KA + k], dh, dzs[j]);
      }
    }
  }
  if (valid) {
    float* dp = dz + (size_t)b * G::NZ * HW + pix;
    for (int j = 0; j < G::NZ; ++j) dp[(size_t)j * HW] = dzs[j * kTvThreads + tid];
  }
  __syncthreads();                                  // zs is dead: reuse it as the [NW][T] reduction buffer
  float* red = zs;
#pragma unroll
  for (int o = 0; o < 2; ++o) {
#pragma unroll
    for (int k = 0; k < G::KA; ++k) red[(o * G::KA + k) * kTvThreads + tid] = dw1a[o][k];
    red[(2 * G::KA + o) * kTvThreads + tid] = db1a[o];
#pragma unroll
    for (int k = 0; k < G::KB; ++k) red[(2 * G::KA + 2 + o * G::KB + k) * kTvThreads + tid] = dw2a[o][k];
  }
  red[(G::NW - 1) * kTvThreads + tid] = gsum;
  __syncthreads();
  if (tid < G::NW) {
    float s = 0.0f;
    for (int i = 0; i < kTvThreads; ++i) s += red[tid * kTvThreads + i];
    part[((size_t)blockIdx.y * gridDim.x + blockIdx.x) * G::NW + tid] = s;
  }
}

template <int TW>
__global__ void __launch_bounds__(64)
k_timeconv_reduce(const float* __restrict__ part, int nblocks, float* __restrict__ dw1, float* __restrict__ db1,
                  float* __restrict__ dw2, float* __restrict__ db2) {
  using G = TvGeom<TW>;
  const int q = threadIdx.x;
  if (q >= G::NW) return;
  double s = 0.0;
  for (int i = 0; i < nblocks; ++i) s += (double)part[(size_t)i * G::NW + q];
  const float v = (float)s;
  if (q < 2 * G::KA) dw1[q] = v;
  else if (q < 2 * G::KA + 2) db1[q - 2 * G::KA] = v;
  else if (q < 2 * G::KA + 2 + 2 * G::KB) dw2[q - 2 * G::KA - 2] = v;
  else db2[0] = v;
}

}  // namespace
}  // namespace pdes

extern "C" {

int pdes_timeconv_ok(int time_window, int num_c) { return (time_window == 25 && num_c == 1) ? 1 : 0; }

int pdes_timeconv_forward(const float* z, const float* w1, const float* b1, const float* w2, const float* b2, float* out,
                          int B, int HW, int time_window, int act, void* stream) {
  using namespace pdes;
  PDES_REQUIRE(z && w1 && b1 && w2 && b2 && out, PDES_ERR_ARG, "pdes_timeconv_forward: null pointer");
  PDES_REQUIRE(B > 0 && HW > 0, PDES_ERR_ARG, "pdes_timeconv_forward: non-positive size");
  PDES_REQUIRE(act == PDES_ACT_NONE || act == PDES_ACT_GELU, PDES_ERR_ARG, "pdes_timeconv_forward: unknown activation");
  PDES_REQUIRE(pdes_timeconv_ok(time_window, 1) && B <= 65535, PDES_ERR_UNSUPPORTED,
               "pdes_timeconv_forward: only time_window = 25, one field (the twophase decoder) is built");
  auto kfn = k_timeconv_fwd<25>;
  PDES_LAUNCH(kfn, dim3((unsigned)ceil_div(HW, kTvThreads), (unsigned)B), dim3(kTvThreads), 0, stream, z, w1, b1, w2, b2, out, HW, act);
  return check_launch("pdes_timeconv_forward");
}

size_t pdes_timeconv_bwd_workspace_floats(int B, int HW, int time_window) {
  if (B <= 0 || HW <= 0 || !pdes_timeconv_ok(time_window, 1)) return 0;
  return (size_t)B * pdes::ceil_div(HW, pdes::kTvThreads) * pdes::TvGeom<25>::NW;
}

int pdes_timeconv_backward(const float* z, const float* gy, const float* w1, const float* b1, const float* w2,
                           const float* b2, float* dz, float* dw1, float* db1, float* dw2, float* db2, float* ws, int B,
                           int HW, int time_window, int act, void* stream) {
  using namespace pdes;
  PDES_REQUIRE(z && gy && w1 && b1 && w2 && b2 && dz && dw1 && db1 && dw2 && db2 && ws, PDES_ERR_ARG,
               "pdes_timeconv_backward: null pointer");
  PDES_REQUIRE(B > 0 && HW > 0, PDES_ERR_ARG, "pdes_timeconv_backward: non-positive size");
  PDES_REQUIRE(act == PDES_ACT_NONE || act == PDES_ACT_GELU, PDES_ERR_ARG, "pdes_timeconv_backward: unknown activation");
  PDES_REQUIRE(pdes_timeconv_ok(time_window, 1) && B <= 65535, PDES_ERR_UNSUPPORTED,
               "pdes_timeconv_backward: only time_window = 25, one field (the twophase decoder) is built");
  using G = TvGeom<25>;
  const size_t smem = (size_t)(2 * G::NZ + 25) * kTvThreads * sizeof(float);
  const int nbx = ceil_div(HW, kTvThreads);
  auto kfn = k_timeconv_bwd<25>;
  PDES_SET_SMEM(kfn, smem);
  PDES_LAUNCH(kfn, dim3((unsigned)nbx, (unsigned)B), dim3(kTvThreads), smem, stream, z, gy, w1, b1, w2, b2, dz, ws, HW, act);
  if (int e = check_launch("pdes_timeconv_backward")) return e;
  auto rfn = k_timeconv_reduce<25>;
  PDES_LAUNCH(rfn, dim3(1), dim3(64), 0, stream, ws, nbx * B, dw1, db1, dw2, db2);
  return check_launch("pdes_timeconv_reduce");
}

}  // extern "C"
