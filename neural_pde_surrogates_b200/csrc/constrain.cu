// Output constraints of one model application, fused (inference / push-forward applications, no autograd):
//
//   d   = u_last + steps[t] * delta                       add_delta            dec_grid.py:8-23 (dec_delta_mode = 'per_step')
//   v   = tanh(d);  v = v - mask * v                      activation_wrapper.py:34-36, :25-31
//   new = sum_hw v,  prev = sum_hw u_last;  dif = (1 - new/prev) * 100;  dif = tanh(dif / cap[t]) / 100 * cap[t]
//   out = v / new * ((1 - dif) * prev);  out = out - mask * out        'individual_static' volume preservation :80-105
//
// The reference runs this as ~25 element-wise / reduction launches over [B,1,tw,H,W]; here ONE CTA per (sample, frame)
// makes two passes over its H*W pixels (the frame is 24 KB, L1/L2-resident between the passes): pass 1 accumulates the
// two sums, pass 2 recomputes v (a tanh is cheaper than a round trip through memory) and writes the result.
// `steps` and `cap` are the reference's own cumulative sums (computed once by torch.cumsum on the host side of the
// binding) so that their fp32 rounding is reproduced exactly.  Sums are reduced in a fixed order (deterministic).
#include "pdes_common.cuh"

namespace pdes {
namespace {

constexpr int kCoThreads = 256;

__device__ __forceinline__ float co_block_sum(float v, float* red) {
  const int tid = threadIdx.x;
  red[tid] = v;
  __syncthreads();
  for (int s = kCoThreads / 2; s > 0; s >>= 1) {
    if (tid < s) red[tid] += red[tid + s];
    __syncthreads();
  }
  const float r = red[0];
  __syncthreads();
  return r;
}

__global__ void __launch_bounds__(kCoThreads)
k_constrain_fwd(const float* __restrict__ delta, const float* __restrict__ x, const float* __restrict__ mask, int mask_bs,
                const float* __restrict__ steps, const float* __restrict__ cap, float* __restrict__ out, int tw, int HW,
                int use_tanh, int use_mask, int use_volume) {
  __shared__ float red[kCoThreads];
  const int t = blockIdx.x, b = blockIdx.y, tid = threadIdx.x;
  const float* ul = x + ((size_t)b * tw + (tw - 1)) * HW;          // last frame of the previous state
  const float* dl = delta + ((size_t)b * tw + t) * HW;
  const float* mk = use_mask ? mask + (size_t)b * mask_bs : nullptr;
  float* o = out + ((size_t)b * tw + t) * HW;
  const float st = __ldg(steps + t);
  auto value = [&](int p) {
    float v = __ldg(ul + p) + st * __ldg(dl + p);
    if (use_tanh) v = tanhf(v);
    if (use_mask) { const float m = __ldg(mk + p); v = v - m * v; }
    return v;
  };
  float scale = 1.0f;
  if (use_volume) {
    float s_new = 0.0f, s_prev = 0.0f;
    for (int p = tid; p < HW; p += kCoThreads) {
      s_new += value(p);
      s_prev += __ldg(ul + p);
    }
    const float new_tot = co_block_sum(s_new, red);
    const float prev_tot = co_block_sum(s_prev, red);
    const float c = __ldg(cap + t);
    float dif = (1.0f - new_tot / prev_tot) * 100.0f;
    dif = tanhf(dif / c) / 100.0f * c;
    scale = (1.0f - dif) * prev_tot;                             // out = (v / new_tot) * scale
    for (int p = tid; p < HW; p += kCoThreads) {
      float v = value(p) / new_tot * scale;
      if (use_mask) { const float m = __ldg(mk + p); v = v - m * v; }
      o[p] = v;
    }
  } else {
    for (int p = tid; p < HW; p += kCoThreads) o[p] = value(p);
  }
}

}  // namespace
}  // namespace pdes

extern "C" int pdes_constrain_forward(const float* delta, const float* x, const float* mask, int mask_bstride,
                                      const float* steps, const float* cap, float* out, int B, int tw, int HW,
                                      int use_tanh, int use_mask, int use_volume, void* stream) {
  using namespace pdes;
  PDES_REQUIRE(delta && x && steps && out, PDES_ERR_ARG, "pdes_constrain_forward: null pointer");
  PDES_REQUIRE(B > 0 && tw > 0 && HW > 0, PDES_ERR_ARG, "pdes_constrain_forward: non-positive size");
  PDES_REQUIRE(!use_mask || mask != nullptr, PDES_ERR_ARG, "pdes_constrain_forward: mask expected");
  PDES_REQUIRE(!use_volume || cap != nullptr, PDES_ERR_ARG, "pdes_constrain_forward: cap expected");
  PDES_REQUIRE(B <= 65535, PDES_ERR_UNSUPPORTED, "pdes_constrain_forward: batch too large");
  auto kfn = k_constrain_fwd;
  PDES_LAUNCH(kfn, dim3((unsigned)tw, (unsigned)B), dim3(kCoThreads), 0, stream, delta, x, mask, mask_bstride, steps, cap, out,
              tw, HW, use_tanh, use_mask, use_volume);
  return check_launch("pdes_constrain_forward");
}
