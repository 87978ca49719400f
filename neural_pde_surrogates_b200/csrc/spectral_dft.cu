// K1: pruned forward DFT  x[B,C,H,W] (real) -> X[B,C,2*m1,m2] (complex), only the retained modes.
//
// Replaces torch.fft.rfft2 + mode slicing (reference proc_fno.py:261,267,269).  One CTA per image (b,c):
//   stage A  Y[h,l]  = sum_w x[h,w] e^{-2 pi i l w/W}         rows staged in shared memory, 4x4 register tiles
//   stage B  X[k,l]  = sum_h Y[h,l] e^{-2 pi i kx_k h/H}      Y kept in shared memory
// The image is read from HBM exactly once (coalesced 128-bit loads); everything else stays on chip.
// Generic in (H, W, m1, m2): rows are processed in chunks of R so that large grids still fit.
#include "pdes_common.cuh"
#include "pdes_ptx.cuh"

namespace pdes {
namespace {

constexpr int kK1Threads = 128;

__global__ void __launch_bounds__(kK1Threads)
k_dft_fwd(const float* __restrict__ x0, int C0, const float* __restrict__ x1, int C1, int H, int W, int m1, int m2,
          int nc4, int R, int xs_stride, const float* __restrict__ twa_g, const float* __restrict__ twh_g,
          const float* __restrict__ lscale, float* __restrict__ X, float2* __restrict__ X2, int B, int CinP) {
  PDES_DYN_SMEM(float, smem);
  float* xs = smem;                                     // [R][xs_stride]
  float* twa = xs + round4((size_t)R * xs_stride);      // [W][nc4]
  float* ys = twa + (size_t)W * nc4;                    // [H][2*m2]
  float* twh = ys + round4((size_t)H * 2 * m2);         // [H][2]

  const int C = C0 + C1;
  const int img = blockIdx.x;
  const int b = img / C, c = img % C;
  const float* src = (c < C0) ? x0 + ((size_t)b * C0 + c) * H * W : x1 + ((size_t)b * C1 + (c - C0)) * H * W;
  const int tid = threadIdx.x, nt = blockDim.x;
  const int J = 2 * m2;

  for (int i = tid; i < W * nc4; i += nt) twa[i] = __ldg(twa_g + i);
  for (int i = tid; i < 2 * H; i += nt) twh[i] = __ldg(twh_g + i);

  const bool vec = (W % 4 == 0) && aligned16(src);
  const int ncg = nc4 / 4;

  for (int h0 = 0; h0 < H; h0 += R) {
    const int rc = (H - h0 < R) ? (H - h0) : R;
    __syncthreads();   // previous chunk fully consumed; tables visible
    if (vec) {
      const float4* s4 = reinterpret_cast<const float4*>(src + (size_t)h0 * W);
      const int nq = rc * W / 4;
      for (int q = tid; q < nq; q += nt) {
        const float4 v = __ldg(s4 + q);
        const int e = q * 4;
        float* d = xs + (e / W) * xs_stride + (e % W);
        d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
      }
    } else {
      const float* s1 = src + (size_t)h0 * W;
      for (int e = tid; e < rc * W; e += nt) xs[(e / W) * xs_stride + (e % W)] = __ldg(s1 + e);
    }
    __syncthreads();

    // ---- stage A: 4 rows x 4 table columns per item
    const int nrg = ceil_div(rc, 4);
    for (int item = tid; item < nrg * ncg; item += nt) {
      const int rg = item / ncg, cg = item % ncg;
      const int r0 = rg * 4;
      const float* xr[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int rr = (r0 + j < rc) ? (r0 + j) : (rc - 1);
        xr[j] = xs + rr * xs_stride;
      }
      float acc[4][4];
#pragma unroll
      for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[j][e] = 0.0f;
      const float* tw = twa + cg * 4;
#pragma unroll 4
      for (int w = 0; w < W; ++w) {
        const float4 t = *reinterpret_cast<const float4*>(tw + (size_t)w * nc4);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float xv = xr[j][w];
          acc[j][0] = fmaf(xv, t.x, acc[j][0]);
          acc[j][1] = fmaf(xv, t.y, acc[j][1]);
          acc[j][2] = fmaf(xv, t.z, acc[j][2]);
          acc[j][3] = fmaf(xv, t.w, acc[j][3]);
        }
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (r0 + j < rc) {
          float* y = ys + (size_t)(h0 + r0 + j) * J + cg * 4;
#pragma unroll
          for (int e = 0; e < 4; ++e)
            if (cg * 4 + e < J) y[e] = acc[j][e];
        }
      }
    }
  }
  __syncthreads();

  // ---- stage B: one retained mode (k,l) per work item
  const int nout = 2 * m1 * m2;
  for (int n = tid; n < nout; n += nt) {
    const int k = n / m2, l = n % m2;
    const int kx = kx_of(k, m1, H);
    float ar = 0.0f, ai = 0.0f;
    int j = 0;
    const float* yl = ys + 2 * l;
    for (int h = 0; h < H; ++h) {
      const float yr = yl[(size_t)h * J], yi = yl[(size_t)h * J + 1];
      const float cs = twh[2 * j], sn = twh[2 * j + 1];
      ar = fmaf(yr, cs, fmaf(yi, sn, ar));      // (yr + i yi)(cs - i sn)
      ai = fmaf(yi, cs, fmaf(-yr, sn, ai));
      j += kx;
      if (j >= H) j -= H;
    }
    if (lscale != nullptr) {
      const float sc = __ldg(lscale + l);
      ar *= sc; ai *= sc;
    }
    float* o = X + ((size_t)img * nout + n) * 2;
    o[0] = ar; o[1] = ai;
    if (X2 != nullptr) X2[((size_t)n * B + b) * CinP + c] = make_float2(ar, ai);      // mode-major copy for K2 on tcgen05
  }
}


// ------------------------------------------------------------------------------------------------------------
// Fast path for the shipped config (H=96, W=64, modes 10x10): one 64-thread CTA per image.
//   * the image arrives with ONE bulk async copy (TMA 1-D, SASS UBLKCP) signalled on an mbarrier;
//   * stage 1 (H axis): thread = column w.  Real input => rows +kx and -kx share their sums: P = sum x cos,
//     Q = sum x sin for kx = 0..m1, with h and H-h folded (cos even, sin odd).  All twiddles are compile-time
//     constants (immediates), the loop is fully unrolled: ~1000 FFMA-imm per column, no table loads;
//   * stage 2 (W axis): 110 work items (kx, l), each four real dot products over w; X[+kx] and X[-kx] come out
//     of the same four sums.
// ~4x fewer instructions per image than the generic kernel.
namespace ct {
constexpr double kPi = 3.14159265358979323846264338327950288;
constexpr double sin_taylor(double x) {
  double t = x, s = x;
  for (int n = 1; n <= 20; ++n) { t *= -x * x / ((2.0 * n) * (2.0 * n + 1.0)); s += t; }
  return s;
}
constexpr double cos_taylor(double x) {
  double t = 1.0, s = 1.0;
  for (int n = 1; n <= 20; ++n) { t *= -x * x / ((2.0 * n - 1.0) * (2.0 * n)); s += t; }
  return s;
}
template <int N> struct Tw { float c[N]; float s[N]; };
template <int N> constexpr Tw<N> make_tw() {
  Tw<N> t{};
  for (int j = 0; j < N; ++j) {
    double a = 2.0 * kPi * j / N;
    if (a > kPi) a -= 2.0 * kPi;
    t.c[j] = (float)cos_taylor(a);
    t.s[j] = (float)sin_taylor(a);
  }
  return t;
}
}  // namespace ct

__device__ constexpr ct::Tw<96> kTw96 = ct::make_tw<96>();

// stage 1 of the fast path for the |kx| range [KB, KE): column DFT with +-kx pairing and h folding, compile-time twiddles
template <int H, int W, int NK, int KB, int KE>
__device__ __forceinline__ void dft_fast_stage1(const float* __restrict__ col, float (*pq)[NK][W + 1], int w) {
  float P[KE - KB], Q[KE - KB];
  const float v0 = col[0], vh = col[(H / 2) * W];
#pragma unroll
  for (int kx = KB; kx < KE; ++kx) {
    P[kx - KB] = (kx & 1) ? (v0 - vh) : (v0 + vh);
    Q[kx - KB] = 0.0f;
  }
#pragma unroll
  for (int h = 1; h < H / 2; ++h) {
    const float a = col[h * W], bq = col[(H - h) * W];
    const float e = a + bq, o = a - bq;
#pragma unroll
    for (int kx = KB; kx < KE; ++kx) {
      P[kx - KB] = fmaf(e, kTw96.c[(kx * h) % H], P[kx - KB]);
      if (kx > 0) Q[kx - KB] = fmaf(o, kTw96.s[(kx * h) % H], Q[kx - KB]);
    }
  }
#pragma unroll
  for (int kx = KB; kx < KE; ++kx) {
    pq[0][kx][w] = P[kx - KB];
    pq[1][kx][w] = Q[kx - KB];
  }
}

// 2*W threads per image: the two halves of the CTA take the two halves of the |kx| range in stage 1 (same shared image,
// twice the warps per SM: the kernel is latency-bound at 2 warps per image) and share the 110 stage-2 items.
template <int H, int W, int M1, int M2>
__global__ void __launch_bounds__(2 * W)
k_dft_fwd_fast(const float* __restrict__ x0, int C0, const float* __restrict__ x1, int C1,
               const float* __restrict__ twa_g, int nc4, const float* __restrict__ tw2_g, const float* __restrict__ lscale,
               float* __restrict__ X, float2* __restrict__ X2, int B, int CinP) {
  static_assert(H == 96, "twiddle table instantiated for H = 96");
  static_assert(H % 2 == 0 && 2 * M1 <= H && M2 <= W / 2 + 1, "fast path preconditions");
  constexpr int NK = M1 + 1;                 // |kx| = 0 .. M1
  __align__(128) __shared__ float img[H * W];
  __shared__ float pq[2][NK][W + 1];
  __align__(16) __shared__ float2 tw2[M2][W + 1];
  __align__(8) __shared__ unsigned long long mbar;

  const int tid = threadIdx.x;
  const int C = C0 + C1;
  const int im = blockIdx.x;
  const int b = im / C, c = im % C;
  const float* src = (c < C0) ? x0 + ((size_t)b * C0 + c) * H * W : x1 + ((size_t)b * C1 + (c - C0)) * H * W;

#ifndef PDES_CPU_EMU
  if (tid == 0) {
    ptx::mbar_init(&mbar, 1);
    ptx::fence_mbar_init();
    // the stage-2 twiddle table arrives with the image: a second bulk copy on the same mbarrier (ncu: the per-thread
    // __ldg fill of this table was 21 % of the kernel's stall samples, a dependent global-load phase in every CTA)
    ptx::mbar_arrive_expect_tx(&mbar, H * W * 4 + M2 * (W + 1) * 8);
    // the image is read again by K3b two kernels later: keep it in L2 (evict-last) across K2's weight stream
    ptx::bulk_g2s_hint(img, src, H * W * 4, &mbar, ptx::l2_policy_evict_last());
    ptx::bulk_g2s(&tw2[0][0], tw2_g, M2 * (W + 1) * 8, &mbar);
  }
  (void)twa_g; (void)nc4;
#else
  (void)mbar; (void)twa_g; (void)nc4;
  for (int i = tid; i < H * W; i += 2 * W) img[i] = src[i];
  for (int i = tid; i < M2 * (W + 1) * 2; i += 2 * W) (&tw2[0][0].x)[i] = tw2_g[i];
#endif
  __syncthreads();                            // mbarrier init (and, emulated, the copies) visible to everyone
#ifndef PDES_CPU_EMU
  ptx::mbar_wait(&mbar, 0);
#endif

  // ---- stage 1 (see dft_fast_stage1)
  {
    constexpr int KH = (NK + 1) / 2;
    const int w = tid & (W - 1);
    if (tid < W) dft_fast_stage1<H, W, NK, 0, KH>(img + w, pq, w);
    else dft_fast_stage1<H, W, NK, KH, NK>(img + w, pq, w);
  }
  __syncthreads();

  PDES_GRID_DEP_LAUNCH();                     // the next kernel of the chain (K2) may be scheduled behind this CTA's tail
  // ---- stage 2: row DFT of (P -+ iQ) to the M2 kept columns
  float* Xo = X + (size_t)im * (2 * M1 * M2) * 2;
  for (int item = tid; item < NK * M2; item += 2 * W) {
    const int kxi = item / M2, l = item % M2;
    const float* pp = pq[0][kxi];
    const float* qq = pq[1][kxi];
    const float2* tt = tw2[l];
    float a = 0.f, bs = 0.f, cq = 0.f, d = 0.f;
#pragma unroll 8
    for (int w = 0; w < W; ++w) {
      const float p = pp[w], q = qq[w];
      const float2 t = tt[w];
      a = fmaf(p, t.x, a);
      bs = fmaf(p, t.y, bs);
      cq = fmaf(q, t.x, cq);
      d = fmaf(q, t.y, d);
    }
    const float sc = (lscale != nullptr) ? __ldg(lscale + l) : 1.0f;
    if (kxi < M1) {                          // +kx -> row k = kx:  sum (P - iQ)(c - is)
      float* o = Xo + ((size_t)kxi * M2 + l) * 2;
      const float vr = sc * (a - d), vi = -sc * (bs + cq);
      o[0] = vr;
      o[1] = vi;
      if (X2 != nullptr) X2[((size_t)(kxi * M2 + l) * B + b) * CinP + c] = make_float2(vr, vi);
    }
    if (kxi > 0) {                           // -kx -> row k = 2*M1 - kx:  sum (P + iQ)(c - is)
      float* o = Xo + ((size_t)(2 * M1 - kxi) * M2 + l) * 2;
      const float vr = sc * (a + d), vi = sc * (cq - bs);
      o[0] = vr;
      o[1] = vi;
      if (X2 != nullptr) X2[((size_t)((2 * M1 - kxi) * M2 + l) * B + b) * CinP + c] = make_float2(vr, vi);
    }
  }
}

}  // namespace
}  // namespace pdes

namespace pdes {
namespace {
int dft_fwd_impl(const float* x0, int C0, const float* x1, int C1, int B, int H, int W, int m1, int m2,
                 const float* tables, int herm_scale, float* X, float* X2f, int CinP, void* stream) {
  float2* X2 = reinterpret_cast<float2*>(X2f);
  PDES_REQUIRE(x0 != nullptr && tables != nullptr && X != nullptr, PDES_ERR_ARG, "pdes_dft_fwd: null pointer");
  PDES_REQUIRE(B > 0 && C0 > 0 && C1 >= 0 && H > 0 && W > 0, PDES_ERR_ARG, "pdes_dft_fwd: non-positive size");
  PDES_REQUIRE((C1 == 0) == (x1 == nullptr), PDES_ERR_ARG, "pdes_dft_fwd: x1/C1 mismatch");
  PDES_REQUIRE(m1 > 0 && m2 > 0 && m1 <= H && m2 <= W / 2 + 1, PDES_ERR_ARG,
               "modes (%d,%d) exceed the grid (%d,%d): need m1 <= H and m2 <= W/2+1", m1, m2, H, W);
  const TableLayout t = table_layout(H, W, m1, m2);
  {
    const int rc = dft_fwd_tc_try(x0, C0, x1, C1, B, H, W, m1, m2, tables, herm_scale, X, X2f, CinP, stream);
    if (rc >= 0) return rc;
  }
  if (H == 96 && W == 64 && m1 == 10 && m2 == 10 && aligned16(x0) && (x1 == nullptr || aligned16(x1)) && aligned16(tables)) {
    auto kfast = k_dft_fwd_fast<96, 64, 10, 10>;
    PDES_MAX_CARVEOUT(kfast);
    PDES_LAUNCH(kfast, dim3((unsigned)(B * (C0 + C1))), dim3(128), 0, stream, x0, C0, x1, C1, tables + t.twa, t.nc4,
                tables + t.tw2, herm_scale ? tables + t.herm : nullptr, X, X2, B, CinP);
    return check_launch("pdes_dft_fwd(fast)");
  }
  const int xs_stride = (W % 2 == 0) ? W + 1 : W;
  const size_t fixed = (size_t)W * t.nc4 + round4((size_t)H * 2 * m2) + round4((size_t)2 * H);
  auto bytes_for = [&](int R) { return (round4((size_t)R * xs_stride) + fixed) * sizeof(float); };
  int R = H;
  if (bytes_for(R) > 48 * 1024) {
    while (R > 4 && bytes_for(R) > 96 * 1024) R = (R > 8) ? ((R / 2 + 3) & ~3) : 4;
    while (R > 4 && bytes_for(R) > (size_t)kMaxDynSmem) R = (R > 8) ? ((R / 2 + 3) & ~3) : 4;
  }
  PDES_REQUIRE(bytes_for(R) <= (size_t)kMaxDynSmem, PDES_ERR_UNSUPPORTED,
               "pdes_dft_fwd: H=%d W=%d m2=%d needs %zu B of shared memory", H, W, m2, bytes_for(R));
  const size_t smem = bytes_for(R);
  auto kfn = k_dft_fwd;
  if (smem > 48 * 1024) PDES_SET_SMEM(kfn, smem);
  PDES_MAX_CARVEOUT(kfn);
  const float* lscale = herm_scale ? tables + t.herm : nullptr;
  PDES_LAUNCH(kfn, dim3((unsigned)(B * (C0 + C1))), dim3(kK1Threads), smem, stream, x0, C0, x1, C1, H, W, m1, m2,
              t.nc4, R, xs_stride, tables + t.twa, tables + t.twh, lscale, X, X2, B, CinP);
  return check_launch("pdes_dft_fwd");
}
}  // namespace
}  // namespace pdes

extern "C" int pdes_dft_fwd(const float* x0, int C0, const float* x1, int C1, int B, int H, int W, int m1, int m2,
                            const float* tables, int herm_scale, float* X, void* stream) {
  return pdes::dft_fwd_impl(x0, C0, x1, C1, B, H, W, m1, m2, tables, herm_scale, X, nullptr, 0, stream);
}

/* K1 that also writes the mode-major copy X2[m][b][i_pad] (complex, i_pad = channels rounded up to 16) read by
 * pdes_mix_tc_fwd; pad columns are left untouched (the consumer masks them). */
extern "C" int pdes_dft_fwd2(const float* x0, int C0, const float* x1, int C1, int B, int H, int W, int m1, int m2,
                             const float* tables, int herm_scale, float* X, float* X2, void* stream) {
  using namespace pdes;
  PDES_REQUIRE(X2 != nullptr, PDES_ERR_ARG, "pdes_dft_fwd2: null X2");
  return dft_fwd_impl(x0, C0, x1, C1, B, H, W, m1, m2, tables, herm_scale, X, X2, (C0 + C1 + 15) / 16 * 16, stream);
}
