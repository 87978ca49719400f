// Small helpers of the block backward: activation backward, 1x1 weight/bias gradients, transpose.
#include "pdes_common.cuh"

namespace pdes {
namespace {

// g_pre = g_out * act'(pre)     (GeluBackward of reference proc_ufno.py:118 / proc_fno.py:153-154)
__global__ void __launch_bounds__(256)
k_act_bwd(const float* __restrict__ g, const float* __restrict__ pre, float* __restrict__ out, size_t n, int vec) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (vec) {
    const size_t n4 = n / 4;
    const float4* g4 = reinterpret_cast<const float4*>(g);
    const float4* p4 = reinterpret_cast<const float4*>(pre);
    float4* o4 = reinterpret_cast<float4*>(out);
    for (; i < n4; i += stride) {
      const float4 a = __ldg(g4 + i), p = __ldg(p4 + i);
      o4[i] = make_float4(a.x * gelu_grad_f(p.x), a.y * gelu_grad_f(p.y), a.z * gelu_grad_f(p.z),
                          a.w * gelu_grad_f(p.w));
    }
  } else {
    for (; i < n; i += stride) out[i] = __ldg(g + i) * gelu_grad_f(__ldg(pre + i));
  }
}

// out[k][m] = in[m][k]
__global__ void __launch_bounds__(256)
k_transpose(const float* __restrict__ in, float* __restrict__ out, int M, int K) {
  __shared__ float tile[32][33];
  const int k0 = blockIdx.x * 32, m0 = blockIdx.y * 32;
  const int tx = threadIdx.x, ty = threadIdx.y;   // block (32, 8)
  for (int r = ty; r < 32; r += 8) {
    const int m = m0 + r, k = k0 + tx;
    tile[r][tx] = (m < M && k < K) ? __ldg(in + (size_t)m * K + k) : 0.0f;
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    const int k = k0 + r, m = m0 + tx;
    if (k < K && m < M) out[(size_t)k * M + m] = tile[tx][r];
  }
}

// 1x1-conv weight gradient: part[s][o][i] = sum_{p in slab s} g[b,o,p] * xin[b,i,p]; column i == K is the
// all-ones input, i.e. the bias gradient.  (ConvolutionBackward of reference proc_fno.py:143.)
constexpr int kWgBM = 64, kWgBN = 64, kWgBK = 16, kWgLd = 68, kWgThreads = 128;

__global__ void __launch_bounds__(kWgThreads)
k_wgrad(const float* __restrict__ g, const float* __restrict__ x0, int C0, const float* __restrict__ x1, int C1,
        float* __restrict__ part, int M, int HW, int slab_px, int slabs_per_b) {
  __align__(16) __shared__ float As[2][kWgBK][kWgLd];
  __align__(16) __shared__ float Bs[2][kWgBK][kWgLd];
  const int K = C0 + C1;
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int n0 = blockIdx.x * kWgBN, m0 = blockIdx.y * kWgBM;
  const int s = blockIdx.z;
  const int b = s / slabs_per_b;
  const int pbeg = (s % slabs_per_b) * slab_px;
  const int pend = (pbeg + slab_px < HW) ? (pbeg + slab_px) : HW;
  const bool vec = (HW % 4 == 0) && (slab_px % 4 == 0) && aligned16(g) && aligned16(x0) &&
                   (x1 == nullptr || aligned16(x1));

  float acc[8][4];
#pragma unroll
  for (int a = 0; a < 8; ++a)
#pragma unroll
    for (int c = 0; c < 4; ++c) acc[a][c] = 0.0f;

  float4 ra[2], rb[2];
  auto load4 = [&](const float* p, int nv) -> float4 {
    if (nv >= 4 && vec) return __ldg(reinterpret_cast<const float4*>(p));
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (nv > 0) v.x = __ldg(p);
    if (nv > 1) v.y = __ldg(p + 1);
    if (nv > 2) v.z = __ldg(p + 2);
    if (nv > 3) v.w = __ldg(p + 3);
    return v;
  };
  auto load_tile = [&](int t) {
    const int pt = pbeg + t * kWgBK;
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const int f = tid + q * kWgThreads;
      const int row = f >> 2, pq = (f & 3) * 4;
      const int pp = pt + pq;
      int nv = pend - pp;
      nv = nv < 0 ? 0 : (nv > 4 ? 4 : nv);
      const int o = m0 + row;
      ra[q] = (o < M && nv > 0) ? load4(g + ((size_t)b * M + o) * HW + pp, nv) : make_float4(0.f, 0.f, 0.f, 0.f);
      const int i = n0 + row;
      if (nv > 0 && i < K) {
        const float* src = (i < C0) ? x0 + ((size_t)b * C0 + i) * HW + pp : x1 + ((size_t)b * C1 + (i - C0)) * HW + pp;
        rb[q] = load4(src, nv);
      } else if (nv > 0 && i == K) {
        rb[q] = make_float4(1.f, nv > 1 ? 1.f : 0.f, nv > 2 ? 1.f : 0.f, nv > 3 ? 1.f : 0.f);
      } else {
        rb[q] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
  };
  auto store_tile = [&](int buf) {
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const int f = tid + q * kWgThreads;
      const int row = f >> 2, pq = (f & 3) * 4;
      As[buf][pq + 0][row] = ra[q].x; As[buf][pq + 1][row] = ra[q].y;
      As[buf][pq + 2][row] = ra[q].z; As[buf][pq + 3][row] = ra[q].w;
      Bs[buf][pq + 0][row] = rb[q].x; Bs[buf][pq + 1][row] = rb[q].y;
      Bs[buf][pq + 2][row] = rb[q].z; Bs[buf][pq + 3][row] = rb[q].w;
    }
  };

  const int nt = ceil_div(pend - pbeg, kWgBK);
  if (nt > 0) {
    load_tile(0);
    store_tile(0);
  }
  __syncthreads();
  for (int t = 0; t < nt; ++t) {
    const int buf = t & 1;
    if (t + 1 < nt) load_tile(t + 1);
#pragma unroll
    for (int kk = 0; kk < kWgBK; ++kk) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][kk][32 + ty * 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][kk][tx * 4]);
      const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float bv[4] = {b0.x, b0.y, b0.z, b0.w};
#pragma unroll
      for (int a = 0; a < 8; ++a)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[a][c] = fmaf(av[a], bv[c], acc[a][c]);
    }
    if (t + 1 < nt) store_tile(buf ^ 1);
    __syncthreads();
  }
#pragma unroll
  for (int a = 0; a < 8; ++a) {
    const int o = m0 + (a >> 2) * 32 + ty * 4 + (a & 3);
    if (o >= M) continue;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int i = n0 + tx * 4 + c;
      if (i <= K) part[((size_t)s * M + o) * (K + 1) + i] = acc[a][c];
    }
  }
}

__global__ void __launch_bounds__(256)
k_wgrad_reduce(const float* __restrict__ part, int nslab, int M, int K, float* __restrict__ dW,
               float* __restrict__ dbias) {
  const int n = M * (K + 1);
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n) return;
  float sum = 0.0f;
  for (int s = 0; s < nslab; ++s) sum += __ldg(part + (size_t)s * n + idx);
  const int o = idx / (K + 1), i = idx % (K + 1);
  if (i < K) {
    if (dW != nullptr) dW[(size_t)o * K + i] = sum;
  } else if (dbias != nullptr) {
    dbias[o] = sum;
  }
}

struct WgradPlan { int slab_px, slabs_per_b, nslab; };
WgradPlan wgrad_plan(int B, int M, int K, int HW) {
  const int tiles = ceil_div(K + 1, kWgBN) * ceil_div(M, kWgBM);
  int spb = ceil_div(300, tiles * B);
  const int max_spb = ceil_div(HW, kWgBK);
  if (spb > max_spb) spb = max_spb;
  if (spb < 1) spb = 1;
  int slab_px = ceil_div(ceil_div(HW, spb), kWgBK) * kWgBK;
  WgradPlan p;
  p.slab_px = slab_px;
  p.slabs_per_b = ceil_div(HW, slab_px);
  p.nslab = p.slabs_per_b * B;
  return p;
}

}  // namespace
}  // namespace pdes

extern "C" {

int pdes_act_bwd(const float* g_out, const float* pre, float* g_pre, size_t n, int act, void* stream) {
  using namespace pdes;
  PDES_REQUIRE(g_out && pre && g_pre, PDES_ERR_ARG, "pdes_act_bwd: null pointer");
  PDES_REQUIRE(act == PDES_ACT_GELU, PDES_ERR_ARG, "pdes_act_bwd: only GELU has a device backward (act=%d)", act);
  if (n == 0) return PDES_OK;
  const int vec = (n % 4 == 0) && aligned16(g_out) && aligned16(pre) && aligned16(g_pre);
  const size_t work = vec ? n / 4 : n;
  size_t blocks = (work + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  auto kfn = k_act_bwd;
  PDES_LAUNCH(kfn, dim3((unsigned)blocks), dim3(256), 0, stream, g_out, pre, g_pre, n, vec);
  return check_launch("pdes_act_bwd");
}

int pdes_transpose(const float* in, float* out, int M, int K, void* stream) {
  using namespace pdes;
  PDES_REQUIRE(in && out && M > 0 && K > 0, PDES_ERR_ARG, "pdes_transpose: bad arguments");
  auto kfn = k_transpose;
  PDES_LAUNCH(kfn, dim3((unsigned)ceil_div(K, 32), (unsigned)ceil_div(M, 32)), dim3(32, 8), 0, stream, in, out, M, K);
  return check_launch("pdes_transpose");
}

size_t pdes_wgrad_workspace_floats(int B, int M, int K, int HW) {
  if (B <= 0 || M <= 0 || K <= 0 || HW <= 0) return 0;
  const pdes::WgradPlan p = pdes::wgrad_plan(B, M, K, HW);
  return (size_t)p.nslab * M * (K + 1);
}

int pdes_wgrad(const float* g, const float* x0, int C0, const float* x1, int C1, float* dW, float* dbias, float* ws,
               int B, int M, int HW, void* stream) {
  using namespace pdes;
  PDES_REQUIRE(g && x0 && ws, PDES_ERR_ARG, "pdes_wgrad: null pointer");
  PDES_REQUIRE(B > 0 && M > 0 && HW > 0 && C0 > 0 && C1 >= 0, PDES_ERR_ARG, "pdes_wgrad: non-positive size");
  PDES_REQUIRE((C1 == 0) == (x1 == nullptr), PDES_ERR_ARG, "pdes_wgrad: x1/C1 mismatch");
  const int K = C0 + C1;
  const WgradPlan p = wgrad_plan(B, M, K, HW);
  PDES_REQUIRE(p.nslab <= 65535, PDES_ERR_UNSUPPORTED, "pdes_wgrad: too many slabs");
  auto kfn = k_wgrad;
  const dim3 grid((unsigned)ceil_div(K + 1, kWgBN), (unsigned)ceil_div(M, kWgBM), (unsigned)p.nslab);
  PDES_LAUNCH(kfn, grid, dim3(kWgThreads), 0, stream, g, x0, C0, x1, C1, ws, M, HW, p.slab_px, p.slabs_per_b);
  if (int e = check_launch("pdes_wgrad")) return e;
  auto rfn = k_wgrad_reduce;
  const int n = M * (K + 1);
  PDES_LAUNCH(rfn, dim3((unsigned)ceil_div(n, 256)), dim3(256), 0, stream, ws, p.nslab, M, K, dW, dbias);
  return check_launch("pdes_wgrad_reduce");
}

}  // extern "C"
