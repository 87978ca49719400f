// K3: pruned inverse DFT fused with the 1x1 conv, bias, residual (U-Net branch) and GELU.
//
//   K3a  Z[b][h][j][c]   = sum_k e^{+2 pi i kx_k h/H} O[b][c][k][l]          (j = 2l + {re,im}, c fastest)
//   K3b  pre[b][o][h][w] = sum_j Z[b][h][j][o] T[j][w] + sum_i At[i][o] xin[b][i][h][w] + bias[o] + res[b][o][h][w]
//        out             = act(pre)
// Replaces zero-padding + torch.fft.irfft2 (reference proc_fno.py:265-269,287), self.w(x) (:143), x1 + x2 (:146),
// h_fno + h_unet and the activation (proc_ufno.py:118 / proc_fno.py:153-154): one pass over the activations,
// one vectorised write.  In K3b the W-axis inverse DFT is simply 2*m2 extra reduction steps of the channel GEMM
// whose "weight" operand is the row's Z vector, so the spectral term never exists as a tensor in HBM.
#include "pdes_common.cuh"

namespace pdes {
namespace {

// ------------------------------------------------------------------------------------------------- K3a
// One CTA = (b, 8 channels): the 8 x (2*m1*m2) spectrum tile is summed over the K2 split partials ONCE into shared
// memory (coalesced along the mode index), then every thread owns one channel and HT rows h and accumulates LCH
// columns at a time: per k it loads HT twiddles (warp-broadcast) and LCH spectrum values for 4*HT*LCH FMAs.
constexpr int kIhOT = 8;                    // channels per CTA (=> 32-byte output sectors)
constexpr int kIhHG = 32;                   // threads per channel; thread owns rows hg, hg+32, ...
constexpr int kIhThreads = kIhOT * kIhHG;
constexpr int kIhLd = kIhOT + 1;            // padded channel stride (float2) of the shared tile

template <int HT, int LCH>
__global__ void __launch_bounds__(kIhThreads)
k_inv_h(const float2* __restrict__ P, int nsplit, int B, int C, int H, int m1, int m2,
        const float* __restrict__ twh_g, float* __restrict__ Z) {
  PDES_DYN_SMEM(float2, smem2);
  const int K = 2 * m1, M2 = K * m2, J = 2 * m2;
  float2* Os = smem2;                               // [M2][kIhLd]
  float2* twh = Os + (size_t)M2 * kIhLd;            // [H]
  const int tid = threadIdx.x;
  const int ot = tid % kIhOT, hg = tid / kIhOT;
  const int c0 = blockIdx.x * kIhOT, b = blockIdx.y;

  for (int i = tid; i < H; i += kIhThreads) twh[i] = make_float2(__ldg(twh_g + 2 * i), __ldg(twh_g + 2 * i + 1));
  // sum the K2 split partials: 8 elements per thread per batch, all loads of one split issued back to back
  {
    constexpr int UN = 8;
    const size_t sstride = (size_t)B * C * M2;
    for (int base0 = tid; base0 < kIhOT * M2; base0 += kIhThreads * UN) {
      const float2* src[UN];
      int dst[UN];
      float2 a[UN];
#pragma unroll
      for (int u = 0; u < UN; ++u) {
        const int idx = base0 + u * kIhThreads;
        const int oo = idx / M2, m = idx - oo * M2;
        const bool ok = idx < kIhOT * M2 && c0 + oo < C;
        src[u] = ok ? P + ((size_t)b * C + c0 + oo) * M2 + m : nullptr;
        dst[u] = (idx < kIhOT * M2) ? m * kIhLd + oo : -1;
        a[u] = make_float2(0.f, 0.f);
      }
      for (int sp = 0; sp < nsplit; ++sp) {
#pragma unroll
        for (int u = 0; u < UN; ++u) {
          if (src[u] != nullptr) {
            const float2 pv = __ldg(src[u] + (size_t)sp * sstride);
            a[u].x += pv.x; a[u].y += pv.y;
          }
        }
      }
#pragma unroll
      for (int u = 0; u < UN; ++u)
        if (dst[u] >= 0) Os[dst[u]] = a[u];
    }
  }
  __syncthreads();

  const bool cvalid = c0 + ot < C;
  for (int hbase = 0; hbase < H; hbase += kIhHG * HT) {
    for (int l0 = 0; l0 < m2; l0 += LCH) {
      float ar[LCH][HT], ai[LCH][HT];
#pragma unroll
      for (int q = 0; q < LCH; ++q)
#pragma unroll
        for (int i = 0; i < HT; ++i) ar[q][i] = ai[q][i] = 0.0f;
      // twiddle index of (k, h0) and its increment per 32 rows, advanced incrementally: kx grows by 1 per k
      // inside each half, so no integer division in the loop (only at k = 0 and k = m1)
      const unsigned h0m = (unsigned)(hbase + hg) % (unsigned)H, s32 = (unsigned)kIhHG % (unsigned)H;
      unsigned j0 = 0, step = 0;
      const bool full = l0 + LCH <= m2;
      const float2* orow = Os + (size_t)l0 * kIhLd + ot;
      for (int k = 0; k < K; ++k) {
        if (k == m1) {
          const unsigned kx = (unsigned)kx_of(k, m1, H);
          j0 = (unsigned)(((unsigned long long)kx * h0m) % (unsigned)H);
          step = (unsigned)(((unsigned long long)kx * s32) % (unsigned)H);
        }
        float2 t[HT];                                     // e^{+i theta} = (cos, sin)
        unsigned j = j0;
#pragma unroll
        for (int i = 0; i < HT; ++i) {
          t[i] = twh[j];
          j += step;
          if (j >= (unsigned)H) j -= (unsigned)H;
        }
        j0 += h0m;
        if (j0 >= (unsigned)H) j0 -= (unsigned)H;
        step += s32;
        if (step >= (unsigned)H) step -= (unsigned)H;
        if (full) {                                        // all LCH columns valid: no per-column guards
#pragma unroll
          for (int q = 0; q < LCH; ++q) {
            const float2 o = orow[q * kIhLd];
#pragma unroll
            for (int i = 0; i < HT; ++i) {
              ar[q][i] = fmaf(o.x, t[i].x, fmaf(-o.y, t[i].y, ar[q][i]));
              ai[q][i] = fmaf(o.x, t[i].y, fmaf(o.y, t[i].x, ai[q][i]));
            }
          }
        } else {
#pragma unroll
          for (int q = 0; q < LCH; ++q) {
            if (l0 + q < m2) {
              const float2 o = orow[q * kIhLd];
#pragma unroll
              for (int i = 0; i < HT; ++i) {
                ar[q][i] = fmaf(o.x, t[i].x, fmaf(-o.y, t[i].y, ar[q][i]));
                ai[q][i] = fmaf(o.x, t[i].y, fmaf(o.y, t[i].x, ai[q][i]));
              }
            }
          }
        }
        orow += m2 * kIhLd;
      }
      if (cvalid) {
#pragma unroll
        for (int i = 0; i < HT; ++i) {
          const int h = hbase + hg + i * kIhHG;
          if (h < H) {
#pragma unroll
            for (int q = 0; q < LCH; ++q) {
              if (l0 + q < m2) {
                float* z = Z + (((size_t)b * H + h) * J + 2 * (l0 + q)) * C + c0 + ot;
                z[0] = ar[q][i];
                z[C] = ai[q][i];
              }
            }
          }
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------- K3b
constexpr int kBM = 64, kBN = 128, kBK = 16, kGemmThreads = 128;

struct InvWParams {
  const float* Z;      // [B][H][2*m2][M] or null
  const float* At;     // [K][lda] or null
  int lda;
  const float* x0; int C0;
  const float* x1; int C1;
  const float* bias;   // [M] or null
  const float* res;    // [B][M][HW] or null
  const float* T;      // [2*m2][W]
  float* out;          // [B][M][HW]
  float* pre;          // [B][M][HW] or null
  int M, H, W, m2, act;
};

__device__ __forceinline__ float4 load4g(const float* p, int nvalid, bool vec) {
  if (nvalid >= 4 && vec) return __ldg(reinterpret_cast<const float4*>(p));
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (nvalid > 0) v.x = __ldg(p);
  if (nvalid > 1) v.y = __ldg(p + 1);
  if (nvalid > 2) v.z = __ldg(p + 2);
  if (nvalid > 3) v.w = __ldg(p + 3);
  return v;
}

__device__ __forceinline__ void store4g(float* p, float4 v, int nvalid, bool vec) {
  if (nvalid >= 4 && vec) { *reinterpret_cast<float4*>(p) = v; return; }
  if (nvalid > 0) p[0] = v.x;
  if (nvalid > 1) p[1] = v.y;
  if (nvalid > 2) p[2] = v.z;
  if (nvalid > 3) p[3] = v.w;
}

__device__ __forceinline__ int clamp04(int n) { return n < 0 ? 0 : (n > 4 ? 4 : n); }

template <bool ROWVEC>
__global__ void __launch_bounds__(kGemmThreads)
k_inv_w_gemm(InvWParams p) {
  __align__(16) __shared__ float As[2][kBK][kBM];
  __align__(16) __shared__ float Bs[2][kBK][kBN];

  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int p0 = blockIdx.x * kBN, m0 = blockIdx.y * kBM, b = blockIdx.z;
  const int HW = p.H * p.W, W = p.W, M = p.M;
  const int K = (p.At != nullptr) ? (p.C0 + p.C1) : 0;
  const int oa[2] = {m0 + ty * 4, m0 + 32 + ty * 4};
  const int pg[2] = {p0 + tx * 4, p0 + 64 + tx * 4};

  float acc[8][8];
#pragma unroll
  for (int a = 0; a < 8; ++a)
#pragma unroll
    for (int c = 0; c < 8; ++c) acc[a][c] = 0.0f;

  // ---------------- spectral term: 2*m2 reduction steps whose A operand is the row's Z vector
  if (p.Z != nullptr) {
    const int J = 2 * p.m2;
    const float* Zb = p.Z + (size_t)b * p.H * J * M;
    const bool zvec = (M % 4 == 0) && aligned16(p.Z);
    if (ROWVEC) {
      const bool tvec = aligned16(p.T);
      int hrow[2], wcol[2];
      bool pv[2];
#pragma unroll
      for (int g = 0; g < 2; ++g) {
        pv[g] = pg[g] < HW;               // W % 4 == 0 -> a 4-pixel group is all-valid or all-invalid
        hrow[g] = pv[g] ? pg[g] / W : 0;
        wcol[g] = pv[g] ? pg[g] % W : 0;
      }
#pragma unroll 2
      for (int j = 0; j < J; ++j) {
        float tb[8], za[2][8];
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          const float4 t4 = load4g(p.T + (size_t)j * W + wcol[g], pv[g] ? 4 : 0, tvec);
          tb[g * 4 + 0] = t4.x; tb[g * 4 + 1] = t4.y; tb[g * 4 + 2] = t4.z; tb[g * 4 + 3] = t4.w;
#pragma unroll
          for (int og = 0; og < 2; ++og) {
            const float4 z4 = load4g(Zb + ((size_t)hrow[g] * J + j) * M + oa[og], pv[g] ? clamp04(M - oa[og]) : 0, zvec);
            za[g][og * 4 + 0] = z4.x; za[g][og * 4 + 1] = z4.y; za[g][og * 4 + 2] = z4.z; za[g][og * 4 + 3] = z4.w;
          }
        }
#pragma unroll
        for (int a = 0; a < 8; ++a)
#pragma unroll
          for (int c = 0; c < 8; ++c) acc[a][c] = fmaf(za[c >> 2][a], tb[c], acc[a][c]);
      }
    } else {
      // generic W: every pixel of the micro tile may sit in a different row
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const int pp = pg[c >> 2] + (c & 3);
        if (pp < HW) {
          const int hh = pp / W, ww = pp % W;
          for (int j = 0; j < J; ++j) {
            const float t = __ldg(p.T + (size_t)j * W + ww);
            const float* zr = Zb + ((size_t)hh * J + j) * M;
#pragma unroll
            for (int a = 0; a < 8; ++a) {
              const int o = oa[a >> 2] + (a & 3);
              if (o < M) acc[a][c] = fmaf(__ldg(zr + o), t, acc[a][c]);
            }
          }
        }
      }
    }
  }

  // ---------------- channel GEMM: acc[o][p] += sum_i At[i][o] * xin[b][i][p]
  if (K > 0) {
    const bool avec = (p.lda % 4 == 0) && aligned16(p.At);
    const bool bvec = (HW % 4 == 0) && aligned16(p.x0) && (p.x1 == nullptr || aligned16(p.x1));
    const int nt = ceil_div(K, kBK);
    float4 ra[2], rb[4];

    auto load_tile = [&](int t) {
      const int k0 = t * kBK;
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const int f = tid + q * kGemmThreads;
        const int kk = f >> 4, m4 = (f & 15) * 4;
        const int k = k0 + kk;
        const int nv = (k < K) ? clamp04(M - (m0 + m4)) : 0;
        ra[q] = (nv > 0) ? load4g(p.At + (size_t)k * p.lda + m0 + m4, nv, avec) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int f = tid + q * kGemmThreads;
        const int kk = f >> 5, n4 = (f & 31) * 4;
        const int i = k0 + kk, pp = p0 + n4;
        const int nv = (i < K) ? clamp04(HW - pp) : 0;
        if (nv > 0) {
          const float* src = (i < p.C0) ? p.x0 + ((size_t)b * p.C0 + i) * HW + pp
                                        : p.x1 + ((size_t)b * p.C1 + (i - p.C0)) * HW + pp;
          rb[q] = load4g(src, nv, bvec);
        } else {
          rb[q] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
    };
    auto store_tile = [&](int buf) {
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const int f = tid + q * kGemmThreads;
        *reinterpret_cast<float4*>(&As[buf][f >> 4][(f & 15) * 4]) = ra[q];
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int f = tid + q * kGemmThreads;
        *reinterpret_cast<float4*>(&Bs[buf][f >> 5][(f & 31) * 4]) = rb[q];
      }
    };

    load_tile(0);
    store_tile(0);
    __syncthreads();
    for (int t = 0; t < nt; ++t) {
      const int buf = t & 1;
      if (t + 1 < nt) load_tile(t + 1);
#pragma unroll
      for (int kk = 0; kk < kBK; ++kk) {
        const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 4]);
        const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][kk][32 + ty * 4]);
        const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][kk][tx * 4]);
        const float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][kk][64 + tx * 4]);
        const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
        const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
        for (int a = 0; a < 8; ++a)
#pragma unroll
          for (int c = 0; c < 8; ++c) acc[a][c] = fmaf(av[a], bv[c], acc[a][c]);
      }
      if (t + 1 < nt) store_tile(buf ^ 1);
      __syncthreads();
    }
  }

  // ---------------- epilogue: + bias + residual, optional pre-activation save, activation, store
  const bool ovec = (HW % 4 == 0) && aligned16(p.out) && (p.res == nullptr || aligned16(p.res)) &&
                    (p.pre == nullptr || aligned16(p.pre));
#pragma unroll
  for (int a = 0; a < 8; ++a) {
    const int o = oa[a >> 2] + (a & 3);
    if (o >= M) continue;
    const float bv = (p.bias != nullptr) ? __ldg(p.bias + o) : 0.0f;
#pragma unroll
    for (int g = 0; g < 2; ++g) {
      const int nv = clamp04(HW - pg[g]);
      if (nv <= 0) continue;
      const size_t off = ((size_t)b * M + o) * HW + pg[g];
      float4 v = make_float4(acc[a][g * 4 + 0] + bv, acc[a][g * 4 + 1] + bv, acc[a][g * 4 + 2] + bv,
                             acc[a][g * 4 + 3] + bv);
      if (p.res != nullptr) {
        const float4 r = load4g(p.res + off, nv, ovec);
        v.x += r.x; v.y += r.y; v.z += r.z; v.w += r.w;
      }
      if (p.pre != nullptr) store4g(p.pre + off, v, nv, ovec);
      if (p.act == PDES_ACT_GELU) {
        v.x = gelu_f(v.x); v.y = gelu_f(v.y); v.z = gelu_f(v.z); v.w = gelu_f(v.w);
      }
      store4g(p.out + off, v, nv, ovec);
    }
  }
}

}  // namespace
}  // namespace pdes

extern "C" {

int pdes_inv_h(const float* P, int nsplit, int B, int C, int H, int m1, int m2, const float* tables, float* Z,
               void* stream) {
  using namespace pdes;
  PDES_REQUIRE(P && tables && Z, PDES_ERR_ARG, "pdes_inv_h: null pointer");
  PDES_REQUIRE(nsplit >= 1 && B > 0 && C > 0 && H > 0 && m1 > 0 && m2 > 0 && m1 <= H, PDES_ERR_ARG,
               "pdes_inv_h: bad sizes");
  PDES_REQUIRE(B <= 65535, PDES_ERR_UNSUPPORTED, "pdes_inv_h: grid too large");
  const size_t smem = ((size_t)2 * m1 * m2 * kIhLd + (size_t)H) * sizeof(float2);
  PDES_REQUIRE(smem <= (size_t)kMaxDynSmem, PDES_ERR_UNSUPPORTED,
               "pdes_inv_h: modes (%d,%d) with H=%d need %zu B of shared memory", m1, m2, H, smem);
  const dim3 grid((unsigned)ceil_div(C, kIhOT), (unsigned)B);
  const TableLayout t = table_layout(H, 1, m1, m2);   // twh offset does not depend on W
  const float2* P2 = reinterpret_cast<const float2*>(P);
#define PDES_INVH_LAUNCH(HT, LCH)                                                                        \
  do {                                                                                                   \
    auto kfn = k_inv_h<HT, LCH>;                                                                         \
    if (smem > 48 * 1024) PDES_SET_SMEM(kfn, smem);                                                      \
    PDES_LAUNCH(kfn, grid, dim3(kIhThreads), smem, stream, P2, nsplit, B, C, H, m1, m2, tables + t.twh, Z); \
  } while (0)
  const int ht = H <= kIhHG * 3 ? 3 : (H <= kIhHG * 4 ? 4 : 8);
  if (m2 % 5 == 0) {
    if (ht == 3) PDES_INVH_LAUNCH(3, 5); else if (ht == 4) PDES_INVH_LAUNCH(4, 5); else PDES_INVH_LAUNCH(8, 5);
  } else {
    if (ht == 3) PDES_INVH_LAUNCH(3, 4); else if (ht == 4) PDES_INVH_LAUNCH(4, 4); else PDES_INVH_LAUNCH(8, 4);
  }
#undef PDES_INVH_LAUNCH
  return check_launch("pdes_inv_h");
}

int pdes_inv_w_gemm(const float* Z, const float* At, int lda, const float* x0, int C0, const float* x1, int C1,
                    const float* bias, const float* res, const float* tables, int backward_scale, float* out,
                    float* pre, int B, int M, int H, int W, int m1, int m2, int act, void* stream) {
  using namespace pdes;
  PDES_REQUIRE(out != nullptr, PDES_ERR_ARG, "pdes_inv_w_gemm: null output");
  PDES_REQUIRE(Z != nullptr || At != nullptr, PDES_ERR_ARG, "pdes_inv_w_gemm: neither spectral nor 1x1 term given");
  PDES_REQUIRE(B > 0 && M > 0 && H > 0 && W > 0, PDES_ERR_ARG, "pdes_inv_w_gemm: non-positive size");
  PDES_REQUIRE(B <= 65535 && ceil_div(M, kBM) <= 65535, PDES_ERR_UNSUPPORTED, "pdes_inv_w_gemm: grid too large");
  PDES_REQUIRE(act == PDES_ACT_NONE || act == PDES_ACT_GELU, PDES_ERR_ARG, "pdes_inv_w_gemm: unknown activation %d", act);
  if (Z != nullptr) {
    PDES_REQUIRE(tables != nullptr, PDES_ERR_ARG, "pdes_inv_w_gemm: spectral term needs tables");
    PDES_REQUIRE(m1 > 0 && m2 > 0 && m2 <= W / 2 + 1, PDES_ERR_ARG, "pdes_inv_w_gemm: modes out of range");
  }
  if (At != nullptr) {
    PDES_REQUIRE(x0 != nullptr && C0 > 0 && C1 >= 0 && lda >= M, PDES_ERR_ARG, "pdes_inv_w_gemm: bad 1x1 operands");
    PDES_REQUIRE((C1 == 0) == (x1 == nullptr), PDES_ERR_ARG, "pdes_inv_w_gemm: x1/C1 mismatch");
  }
  InvWParams p;
  p.Z = Z; p.At = At; p.lda = lda; p.x0 = x0; p.C0 = C0; p.x1 = x1; p.C1 = C1; p.bias = bias; p.res = res;
  p.T = nullptr;
  if (Z != nullptr) {
    const TableLayout t = table_layout(H, W, m1, m2);
    p.T = tables + (backward_scale ? t.tinv_b : t.tinv_f);
  }
  p.out = out; p.pre = pre; p.M = M; p.H = H; p.W = W; p.m2 = m2; p.act = act;
  const dim3 grid((unsigned)ceil_div(H * W, kBN), (unsigned)ceil_div(M, kBM), (unsigned)B);
  if (W % 4 == 0) {
    auto kfn = k_inv_w_gemm<true>;
    PDES_LAUNCH(kfn, grid, dim3(kGemmThreads), 0, stream, p);
  } else {
    auto kfn = k_inv_w_gemm<false>;
    PDES_LAUNCH(kfn, grid, dim3(kGemmThreads), 0, stream, p);
  }
  return check_launch("pdes_inv_w_gemm");
}

}  // extern "C"
