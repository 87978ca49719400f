// K2 on the 5th-generation tensor cores: per-mode complex channel mixing as a real 2x2-block GEMM (tcgen05 + TMEM,
// TMA-staged tiles, 3xTF32 split => fp32-faithful), plus the packed master copy of the spectral weights it streams
// and the H-axis inverse DFT (K3a) that consumes its output.
//
//   reference: compl_mul2d = einsum("bixy,ioxy->boxy") x2, proc_fno.py:253-255,266-269
//
// Per retained mode m = (k, l) the einsum is the GEMM  O_m[o, b] = sum_i W_m[i, o] X_m[b, i]  (complex).  With the
// output channel on the M side (TMEM lanes), N = (b, re|im) and K = i it becomes, in real arithmetic,
//     D_m[o, (b,re)] = sum_i Wr[o,i] Xr[b,i] - Wi[o,i] Xi[b,i]
//     D_m[o, (b,im)] = sum_i Wr[o,i] Xi[b,i] + Wi[o,i] Xr[b,i]
// i.e.  D = Wr * XA + Wi * XB  with XA[(b,re|im)][i] = (Xr | Xi) and XB = (-Xi | Xr): the weights are read ONCE and
// never expanded to the 2x2 real block in memory.  The kernel is bound by streaming the weights
// (16 * Cin * Cout * m1 * m2 bytes); the tensor pipe is a few percent busy by construction.
//
// Layouts
//   Wp  packed master copy  [m][o][i_pad][re|im] fp32, i_pad = Cin rounded up to 16, zero padded; rows of the first
//       weight block that the reference overwrites when 2*m1 > H (proc_fno.py:266-269) are stored as zeros.  Built by
//       pdes_mix_tc_pack() once per weight version (the caller caches it); the parameters themselves, Adam and the
//       gradient all-reduce keep the reference layout [Cin][Cout][m1][m2].
//   X2  [m][b][i_pad] complex, mode-major copy of the retained spectrum written by K1 (pad columns are never read
//       unmasked);
//   O2  [2][m][b][o] complex: partial 0 and (for work items split between two CTAs) partial 1, summed by K3a.
//
// Work decomposition: an item = (mode m, tile of <= 128 output channels) = i_pad/16 chunks of 16 input channels; the
// flattened chunk stream is cut into equal contiguous ranges, one per CTA (persistent, one CTA per SM), so the load
// balance is within one chunk; an item that straddles a cut is finished by the next CTA into partial 1.
//
// Pipeline of one CTA (448 threads):
//   warp 13      TMA: per chunk one 3-D box [TO rows o][16 i x (re,im)] of Wp (128-byte swizzle) and one 2-D box
//                [B rows][16 i x (re,im)] of X2 into a raw ring
//   warps 0-7    convert, two groups taking alternate chunks: thread = output channel = TMEM lane; reads its 128-byte
//                row, de-interleaves re/im, splits hi/lo and writes the A operand straight into TENSOR MEMORY
//                (tcgen05.st); also turns the raw X rows into the four canonical K-major B blocks (XA/XB x hi/lo)
//   warp 12      one thread issues 12 tcgen05.mma (M128, N = 2B padded to 16, K8, kind::tf32) per chunk: the four
//                small cross terms go to a LOW accumulator, the two hi*hi terms to a HIGH accumulator (tcgen05's fp32
//                accumulation truncates; keeping the large terms in their own, shorter sum halves the error)
//   warps 8-11   epilogue: TMEM -> registers (hi + lo) -> O2, coalesced float2 per lane; double-buffered accumulators
#include "pdes_common.cuh"
#include "pdes_ptx.cuh"
#ifndef PDES_CPU_EMU
#include <cuda.h>
#include <cstring>
#endif

namespace pdes {

constexpr int kMtBK = 16;          // input channels per chunk
constexpr int kMtMaxN = 64;        // padded 2B (two split accumulators x double buffering = 4 * N <= 256 TMEM columns)

__host__ __device__ inline int mt_cinp(int Cin) { return (Cin + kMtBK - 1) / kMtBK * kMtBK; }
__host__ __device__ inline int mt_npad(int B) { int n = (2 * B + 15) & ~15; return n < 16 ? 16 : n; }
__host__ __device__ inline int mt_ntile(int Cout) { return (Cout + 127) / 128; }
__host__ __device__ inline int mt_to(int Cout) { const int nt = mt_ntile(Cout); return (((Cout + nt - 1) / nt) + 7) & ~7; }

namespace {

// ------------------------------------------------------------------------------------------------- weight pack
// [Cin][Cout][m1*m2] complex (x2 blocks) -> Wp[m][o][i_pad] complex.  Per output channel this is a 2-D transpose between
// the input-channel and the mode index: a 32 x 32 tile goes through shared memory so that both the loads (along the
// mode index) and the stores (along i) are coalesced.
__global__ void __launch_bounds__(256)
k_mix_tc_pack(const float2* __restrict__ w1, const float2* __restrict__ w2, float2* __restrict__ Wp, int Cin, int Cout,
              int CinP, int m1, int m2, int H) {
  __shared__ float2 tile[32][33];
  const int MM = m1 * m2;
  const int o = blockIdx.x;
  const int i0 = blockIdx.y * 32;
  const int mt_per_half = (MM + 31) / 32;
  const int half = blockIdx.z / mt_per_half, mm0 = (blockIdx.z % mt_per_half) * 32;
  const float2* w = half ? w2 : w1;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;            // 32 x 8
  for (int r = ty; r < 32; r += 8) {                                  // r = input channel, tx = mode
    const int i = i0 + r, mm = mm0 + tx;
    float2 v = make_float2(0.f, 0.f);
    if (i < Cin && mm < MM) v = __ldg(w + ((size_t)i * Cout + o) * MM + mm);
    tile[r][tx] = v;
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {                                  // r = mode, tx = input channel
    const int mm = mm0 + r, i = i0 + tx;
    if (mm < MM && i < CinP) {
      const int k = half * m1 + mm / m2;
      float2 v = tile[tx][r];
      if (row_dead(k, m1, H)) v = make_float2(0.f, 0.f);
      Wp[((size_t)(half * MM + mm) * Cout + o) * CinP + i] = v;
    }
  }
}

// ------------------------------------------------------------------------------------------------- K3a (v2)
// Z[b][h][2l + {re,im}][o] = sum_k e^{+2 pi i kx_k h / H} (O2[0][m][b][o] + [item split] O2[1][m][b][o]),  m = k*m2 + l.
// CTA = (32 output channels, one l, one sample): lanes run along o (coalesced loads of O2 and stores of Z), the four
// warps take the row pairs (h, H-h) round-robin.  A pair shares its four real sums:
//     P = sum_k cos(t_k h) O_k,  Q = sum_k sin(t_k h) O_k  =>  z(h) = P + iQ,  z(H-h) = P - iQ,
// which halves the multiplies for any (H, m1).  Twiddles come from a per-CTA shared table (warp-broadcast loads).
struct SplitRule { int nck, per, ntile, to; };       // how K2 cut its chunk stream (to find the items with two partials)
__device__ __forceinline__ bool item_is_split(const SplitRule& r, int item) {
  return (item * r.nck) / r.per != ((item + 1) * r.nck - 1) / r.per;
}

constexpr int kIh2Threads = 128, kIh2HP = 4;           // row pairs per thread per pass
__global__ void __launch_bounds__(kIh2Threads)
k_inv_h2(const float2* __restrict__ O2, SplitRule rule, int B, int C, int H, int m1, int m2,
         const float* __restrict__ twh_g, float* __restrict__ Z) {
  PDES_DYN_SMEM(float2, sm2);
  const int K = 2 * m1, M2 = K * m2, J = 2 * m2;
  const int npair = H / 2 + 1;
  float2* Os = sm2;                                    // [K][32]
  float2* tw = Os + (size_t)K * 32;                    // [npair][K]
  const int tid = threadIdx.x, lane = tid & 31, wq = tid >> 5;
  const int o0 = blockIdx.x * 32, l = blockIdx.y, b = blockIdx.z;
  const int o = o0 + lane;
  for (int idx = tid; idx < npair * K; idx += kIh2Threads) {
    const int p = idx / K, k = idx - p * K;
    const int j = (int)(((long long)kx_of(k, m1, H) * p) % H);
    tw[idx] = make_float2(__ldg(twh_g + 2 * j), __ldg(twh_g + 2 * j + 1));
  }
  const size_t pstride = (size_t)M2 * B * C;
  for (int idx = tid; idx < K * 32; idx += kIh2Threads) {
    const int k = idx >> 5, oo = o0 + (idx & 31);
    float2 v = make_float2(0.f, 0.f);
    if (oo < C) {
      const int m = k * m2 + l;
      const size_t at = ((size_t)m * B + b) * C + oo;
      v = __ldg(O2 + at);
      if (item_is_split(rule, m * rule.ntile + oo / rule.to)) {
        const float2 u = __ldg(O2 + pstride + at);
        v.x += u.x; v.y += u.y;
      }
    }
    Os[idx] = v;
  }
  __syncthreads();
  for (int p0 = wq * kIh2HP; p0 < npair; p0 += 4 * kIh2HP) {
    float pr[kIh2HP], pi[kIh2HP], qr[kIh2HP], qi[kIh2HP];
#pragma unroll
    for (int e = 0; e < kIh2HP; ++e) pr[e] = pi[e] = qr[e] = qi[e] = 0.0f;
    for (int k = 0; k < K; ++k) {
      const float2 ov = Os[k * 32 + lane];
#pragma unroll
      for (int e = 0; e < kIh2HP; ++e) {
        const int p = (p0 + e < npair) ? (p0 + e) : (npair - 1);
        const float2 t = tw[p * K + k];                                   // warp-uniform address: broadcast
        pr[e] = fmaf(t.x, ov.x, pr[e]);
        pi[e] = fmaf(t.x, ov.y, pi[e]);
        qr[e] = fmaf(t.y, ov.x, qr[e]);
        qi[e] = fmaf(t.y, ov.y, qi[e]);
      }
    }
    if (o < C) {
#pragma unroll
      for (int e = 0; e < kIh2HP; ++e) {
        const int h = p0 + e;
        if (h < npair) {
          float* z = Z + (((size_t)b * H + h) * J + 2 * l) * C + o;
          z[0] = pr[e] - qi[e];                                           // Re (P + iQ)
          z[C] = pi[e] + qr[e];
          const int h2 = H - h;
          if (h != 0 && h2 != h && h2 < H) {
            float* z2 = Z + (((size_t)b * H + h2) * J + 2 * l) * C + o;
            z2[0] = pr[e] + qi[e];                                        // Re (P - iQ)
            z2[C] = pi[e] - qr[e];
          }
        }
      }
    }
  }
}

#ifndef PDES_CPU_EMU
// ------------------------------------------------------------------------------------------------- K2 on tcgen05
constexpr int kMtThreads = 448, kMtMmaWarp = 12, kMtTmaWarp = 13;
constexpr int kMtNST = 4;                              // TMEM A stages / shared-memory B stages
constexpr int kMtMaxRaw = 8;

struct MtBars {
  unsigned long long full[kMtNST], empty[kMtNST], raw_full[kMtMaxRaw], raw_empty[kMtMaxRaw], acc_full[2], acc_empty[2];
};

struct MtParams {
  float2* O2;
  int B, Cin, Cout, CinP, nmodes, npad, bp, ntile, to, nck, per, nch_total, nraw;
};

__device__ __forceinline__ float mt_hi(float x) { return __uint_as_float(__float_as_uint(x) & 0xffffe000u); }

__global__ void __launch_bounds__(kMtThreads, 1)
k_mix_tc(MtParams p, const __grid_constant__ CUtensorMap tmap_w, const __grid_constant__ CUtensorMap tmap_x) {
  PDES_DYN_SMEM(unsigned char, smem_raw);
  unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int npad = p.npad;
  const uint32_t w_slot = 128u * 128u;                                     // 16 KB: [128 rows][128 B], rows >= TO unused
  const uint32_t x_slot = (uint32_t)p.bp * 128u;                           // [bp rows][16 complex]
  const uint32_t raw_slot = w_slot + ((x_slot + 1023u) & ~1023u);
  const uint32_t lbo = (uint32_t)(npad / 8) * 128u + 16u;                  // +16: the 4 k-quads of a row fall in different banks
  const uint32_t blk = 4u * lbo;                                           // one canonical [npad x 16] block
  const uint32_t b_stage = 4u * blk;                                       // XA_hi, XA_lo, XB_hi, XB_lo
  unsigned char* sRaw = base;
  unsigned char* sB = sRaw + (size_t)p.nraw * raw_slot;
  __shared__ __align__(8) MtBars bars;
  __shared__ uint32_t tmem_slot;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int NRAW = p.nraw;
  const int c_beg = (int)blockIdx.x * p.per;
  const int c_end = (c_beg + p.per < p.nch_total) ? (c_beg + p.per) : p.nch_total;

  if (tid == 0) {
    for (int i = 0; i < kMtNST; ++i) { ptx::mbar_init(&bars.full[i], 4); ptx::mbar_init(&bars.empty[i], 1); }
    for (int i = 0; i < NRAW; ++i) { ptx::mbar_init(&bars.raw_full[i], 1); ptx::mbar_init(&bars.raw_empty[i], 4); }
    for (int i = 0; i < 2; ++i) { ptx::mbar_init(&bars.acc_full[i], 1); ptx::mbar_init(&bars.acc_empty[i], 4); }
    ptx::fence_mbar_init();
  }
  if (warp == kMtMmaWarp) {
    ptx::tmem_alloc(&tmem_slot, 512);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  // TMEM columns: accumulator a (0/1): high sum at a*2*npad, low sum at a*2*npad + npad; A stage s at 256 + 64*s:
  // [Wr_hi 16][Wr_lo 16][Wi_hi 16][Wi_lo 16]
  const uint32_t ta0 = 256u;

  if (warp < 8) {
    // ================================================================== convert
    const int grp = warp >> 2, cw = warp & 3;
    const int row = tid & 127;                                             // output channel within the tile = TMEM lane
    for (int c = c_beg, g = 0; c < c_end; ++c, ++g) {
      if ((g & 1) != grp) continue;
      const int s = g % kMtNST, r = g % NRAW;
      const int kc = c % p.nck;
      if (g >= kMtNST) ptx::mbar_wait(&bars.empty[s], (uint32_t)(((g / kMtNST) - 1) & 1));
      ptx::tc_fence_after();
      ptx::mbar_wait(&bars.raw_full[r], (uint32_t)((g / NRAW) & 1));
      const unsigned char* slot = sRaw + (size_t)r * raw_slot;
      // ---- weights: this thread's 128-byte row (16 complex), 128-byte swizzle: chunk j sits at j ^ (row & 7)
      uint32_t rh[16], rl[16], ih[16], il[16];
      {
        const unsigned char* wrow = slot + (uint32_t)row * 128u;
        const uint32_t sw = (uint32_t)(row & 7);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 v = *reinterpret_cast<const float4*>(wrow + (((uint32_t)j ^ sw) << 4));   // (re, im) of i = 2j, 2j+1
          float h;
          h = mt_hi(v.x); rh[2 * j] = __float_as_uint(h); rl[2 * j] = __float_as_uint(v.x - h);
          h = mt_hi(v.y); ih[2 * j] = __float_as_uint(h); il[2 * j] = __float_as_uint(v.y - h);
          h = mt_hi(v.z); rh[2 * j + 1] = __float_as_uint(h); rl[2 * j + 1] = __float_as_uint(v.z - h);
          h = mt_hi(v.w); ih[2 * j + 1] = __float_as_uint(h); il[2 * j + 1] = __float_as_uint(v.w - h);
        }
      }
      const uint32_t trow = tmem_base + ((uint32_t)(cw * 32) << 16) + ta0 + (uint32_t)s * 64u;
      ptx::tmem_st16(trow, rh);
      ptx::tmem_st16(trow + 16u, rl);
      ptx::tmem_st16(trow + 32u, ih);
      ptx::tmem_st16(trow + 48u, il);
      // ---- spectrum rows -> canonical K-major B blocks: element (n, kk) at (kk/4)*lbo + (n/8)*128 + (n%8)*16 + (kk%4)*4
      {
        const unsigned char* xraw = slot + w_slot;
        unsigned char* sb = sB + (size_t)s * b_stage;
        const int i0 = kc * kMtBK;
        for (int e = row; e < p.bp * kMtBK; e += 128) {
          const int bb = e >> 4, kk = e & 15;
          float2 x = *reinterpret_cast<const float2*>(xraw + (uint32_t)bb * 128u + (uint32_t)kk * 8u);
          if (bb >= p.B || i0 + kk >= p.Cin) x = make_float2(0.f, 0.f);
          const float xr_h = mt_hi(x.x), xi_h = mt_hi(x.y);
          const float xr_l = x.x - xr_h, xi_l = x.y - xi_h;
          const int n0 = 2 * bb, n1 = n0 + 1;
          const uint32_t kof = (uint32_t)(kk >> 2) * lbo + (uint32_t)(kk & 3) * 4u;
          const uint32_t o0f = kof + (uint32_t)(n0 >> 3) * 128u + (uint32_t)(n0 & 7) * 16u;
          const uint32_t o1f = kof + (uint32_t)(n1 >> 3) * 128u + (uint32_t)(n1 & 7) * 16u;
          *reinterpret_cast<float*>(sb + o0f) = xr_h;                        // XA: (re | im)
          *reinterpret_cast<float*>(sb + o1f) = xi_h;
          *reinterpret_cast<float*>(sb + blk + o0f) = xr_l;
          *reinterpret_cast<float*>(sb + blk + o1f) = xi_l;
          *reinterpret_cast<float*>(sb + 2 * blk + o0f) = -xi_h;             // XB: (-im | re)
          *reinterpret_cast<float*>(sb + 2 * blk + o1f) = xr_h;
          *reinterpret_cast<float*>(sb + 3 * blk + o0f) = -xi_l;
          *reinterpret_cast<float*>(sb + 3 * blk + o1f) = xr_l;
        }
      }
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&bars.raw_empty[r]);                   // raw slot fully consumed
      ptx::tmem_st_wait();
      ptx::tc_fence_before();
      ptx::fence_proxy_async();                                              // the B blocks went through st.shared
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&bars.full[s]);
    }
  } else if (warp == kMtMmaWarp) {
    if (lane == 0) {
      // ================================================================ MMA issue
      const uint32_t idesc = ptx::idesc_tf32(128, npad);
      const uint64_t bd0 = ptx::smem_desc_noswizzle(ptx::smem_u32(sB), lbo, 128);
      uint32_t part = 0;
      bool fresh = true;
      for (int c = c_beg, g = 0; c < c_end; ++c, ++g) {
        const int s = g % kMtNST;
        const int kc = c % p.nck;
        const uint32_t a = part & 1u;
        if (fresh) {
          if (part >= 2) ptx::mbar_wait(&bars.acc_empty[a], ((part >> 1) - 1) & 1u);
          ptx::tc_fence_after();
        }
        ptx::mbar_wait(&bars.full[s], (uint32_t)((g / kMtNST) & 1));
        ptx::tc_fence_after();
        const uint32_t d_hi = tmem_base + a * 2u * (uint32_t)npad, d_lo = d_hi + (uint32_t)npad;
        const uint32_t ta = tmem_base + ta0 + (uint32_t)s * 64u;
        const uint64_t sd = (uint64_t)(((uint32_t)s * b_stage) >> 4);
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
          const uint64_t kb = sd + (uint64_t)((ks * 2 * lbo) >> 4);
          const uint64_t xa_hi = bd0 + kb, xa_lo = bd0 + kb + (blk >> 4), xb_hi = bd0 + kb + ((2 * blk) >> 4),
                         xb_lo = bd0 + kb + ((3 * blk) >> 4);
          const uint32_t wr_hi = ta + ks * 8, wr_lo = ta + 16 + ks * 8, wi_hi = ta + 32 + ks * 8, wi_lo = ta + 48 + ks * 8;
          const uint32_t acc0 = (fresh && ks == 0) ? 0u : 1u;
          ptx::mma_tf32_ta(d_lo, wr_lo, xa_hi, idesc, acc0);                 // cross terms -> low accumulator
          ptx::mma_tf32_ta(d_lo, wr_hi, xa_lo, idesc, 1u);
          ptx::mma_tf32_ta(d_lo, wi_lo, xb_hi, idesc, 1u);
          ptx::mma_tf32_ta(d_lo, wi_hi, xb_lo, idesc, 1u);
          ptx::mma_tf32_ta(d_hi, wr_hi, xa_hi, idesc, acc0);                 // hi*hi terms -> high accumulator
          ptx::mma_tf32_ta(d_hi, wi_hi, xb_hi, idesc, 1u);
        }
        ptx::tc_commit(&bars.empty[s]);
        fresh = false;
        if (kc == p.nck - 1 || c + 1 == c_end) {                             // item (or this CTA's part of it) complete
          ptx::tc_commit(&bars.acc_full[a]);
          ++part;
          fresh = true;
        }
      }
    }
  } else if (warp == kMtTmaWarp) {
    if (lane == 0) {
      // ================================================================ TMA producer
      const uint32_t wbytes = (uint32_t)p.to * 128u, xbytes = (uint32_t)p.bp * 128u;
      for (int c = c_beg, g = 0; c < c_end; ++c, ++g) {
        const int r = g % NRAW;
        if (g >= NRAW) ptx::mbar_wait(&bars.raw_empty[r], (uint32_t)(((g / NRAW) - 1) & 1));
        const int item = c / p.nck, kc = c - item * p.nck;
        const int m = item / p.ntile, t = item - m * p.ntile;
        unsigned char* slot = sRaw + (size_t)r * raw_slot;
        ptx::mbar_arrive_expect_tx(&bars.raw_full[r], wbytes + xbytes);
        ptx::tma_load_3d(slot, &tmap_w, kc * 2 * kMtBK, t * p.to, m, &bars.raw_full[r]);
        ptx::tma_load_2d(slot + w_slot, &tmap_x, kc * 2 * kMtBK, m * p.B, &bars.raw_full[r]);
      }
    }
  } else if (warp >= 8 && warp < 12) {
    // ==================================================================== epilogue: lane = output channel
    const int quad = warp & 3;
    const int rowl = quad * 32 + lane;
    const size_t pstride = (size_t)p.nmodes * p.B * p.Cout;
    uint32_t part = 0;
    for (int c = c_beg; c < c_end;) {
      const int item = c / p.nck, kc0 = c - item * p.nck;
      int cl = (item + 1) * p.nck;                                           // end of this part
      if (cl > c_end) cl = c_end;
      const int m = item / p.ntile, t = item - m * p.ntile;
      const int o = t * p.to + rowl;
      const bool valid = rowl < p.to && o < p.Cout;
      const uint32_t a = part & 1u;
      ptx::mbar_wait(&bars.acc_full[a], (part >> 1) & 1u);
      ptx::tc_fence_after();
      const uint32_t tb = tmem_base + ((uint32_t)(quad * 32) << 16) + a * 2u * (uint32_t)npad;
      float2* dst = p.O2 + (kc0 != 0 ? pstride : 0) + ((size_t)m * p.B) * p.Cout + o;   // partial 1 = continuation of a split item
      for (int n0 = 0; n0 < npad; n0 += 8) {
        uint32_t vh[8], vl[8];
        ptx::tmem_ld8(tb + (uint32_t)n0, vh);
        ptx::tmem_ld8(tb + (uint32_t)(npad + n0), vl);
        ptx::tmem_ld_wait();
        if (n0 + 8 >= npad) {                                                // accumulators fully read: hand them back
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive(&bars.acc_empty[a]);
        }
        if (valid) {
#pragma unroll
          for (int e = 0; e < 8; e += 2) {
            const int bb = (n0 + e) >> 1;
            if (bb < p.B)
              dst[(size_t)bb * p.Cout] = make_float2(__uint_as_float(vh[e]) + __uint_as_float(vl[e]),
                                                     __uint_as_float(vh[e + 1]) + __uint_as_float(vl[e + 1]));
          }
        }
      }
      ++part;
      c = cl;
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == kMtMmaWarp) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, 512);
  }
}

typedef CUresult (*MtEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int g_mt_sms = 0;
#endif  // !PDES_CPU_EMU

struct MtGeom { int CinP, npad, bp, ntile, to, nck, nitems, nch, G, per; };
inline MtGeom mt_geom(int B, int Cin, int Cout, int m1, int m2, int sms) {
  MtGeom g;
  g.CinP = mt_cinp(Cin);
  g.npad = mt_npad(B);
  g.bp = g.npad / 2;
  g.ntile = mt_ntile(Cout);
  g.to = mt_to(Cout);
  g.nck = g.CinP / kMtBK;
  g.nitems = 2 * m1 * m2 * g.ntile;
  g.nch = g.nitems * g.nck;
  g.G = sms < g.nitems ? sms : g.nitems;             // a CTA's share is >= one item => an item has at most two parts
  if (g.G < 1) g.G = 1;
  g.per = (g.nch + g.G - 1) / g.G;
  if (g.per < g.nck) g.per = g.nck;
  g.G = (g.nch + g.per - 1) / g.per;
  return g;
}
inline int mt_sms() {
#ifdef PDES_CPU_EMU
  return 148;
#else
  if (g_mt_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_mt_sms, cudaDevAttrMultiProcessorCount, dev);
    if (g_mt_sms <= 0) g_mt_sms = 148;
  }
  return g_mt_sms;
#endif
}

}  // namespace
}  // namespace pdes

extern "C" {

/* 1 when the tensor-core K2 covers the shape: 2B (padded to 16) <= 64 and the tensor-core mode is on. */
int pdes_mix_tc_ok(int B, int Cin, int Cout, int m1, int m2) {
#ifdef PDES_CPU_EMU
  (void)B; (void)Cin; (void)Cout; (void)m1; (void)m2;
  return 0;
#else
  if (B <= 0 || Cin <= 0 || Cout <= 0 || m1 <= 0 || m2 <= 0) return 0;
  if (pdes_get_tensor_core_mode() < 2 || pdes::tensor_map_encoder() == nullptr) return 0;
  if (pdes::mt_npad(B) > pdes::kMtMaxN) return 0;
  if ((long)2 * m1 * m2 * B >= (1L << 30) || (long)pdes::mt_cinp(Cin) * 2 > (1L << 30)) return 0;
  return 1;
#endif
}

size_t pdes_mix_tc_pack_floats(int Cin, int Cout, int m1, int m2) {
  if (Cin <= 0 || Cout <= 0 || m1 <= 0 || m2 <= 0) return 0;
  return (size_t)2 * m1 * m2 * Cout * pdes::mt_cinp(Cin) * 2;
}

size_t pdes_mix_tc_x2_floats(int B, int Cin, int m1, int m2) {
  if (B <= 0 || Cin <= 0 || m1 <= 0 || m2 <= 0) return 0;
  return (size_t)2 * m1 * m2 * B * pdes::mt_cinp(Cin) * 2;
}

size_t pdes_mix_tc_o2_floats(int B, int Cout, int m1, int m2) {
  if (B <= 0 || Cout <= 0 || m1 <= 0 || m2 <= 0) return 0;
  return (size_t)2 * 2 * m1 * m2 * B * Cout * 2;
}

/* Packed master copy of the spectral weights (see the header comment).  Works in every build (plain CUDA kernel). */
int pdes_mix_tc_pack(const float* w1, const float* w2, float* Wp, int Cin, int Cout, int H, int m1, int m2, void* stream) {
  using namespace pdes;
  PDES_REQUIRE(w1 && w2 && Wp, PDES_ERR_ARG, "pdes_mix_tc_pack: null pointer");
  PDES_REQUIRE(Cin > 0 && Cout > 0 && m1 > 0 && m2 > 0 && H > 0 && m1 <= H, PDES_ERR_ARG, "pdes_mix_tc_pack: bad sizes");
  const int CinP = mt_cinp(Cin), MM = m1 * m2;
  PDES_REQUIRE(Cout <= 65535, PDES_ERR_UNSUPPORTED, "pdes_mix_tc_pack: too many output channels");
  auto kfn = k_mix_tc_pack;
  const dim3 grid((unsigned)Cout, (unsigned)ceil_div(CinP, 32), (unsigned)(2 * ceil_div(MM, 32)));
  PDES_LAUNCH(kfn, grid, dim3(256), 0, stream, reinterpret_cast<const float2*>(w1), reinterpret_cast<const float2*>(w2),
              reinterpret_cast<float2*>(Wp), Cin, Cout, CinP, m1, m2, H);
  return check_launch("pdes_mix_tc_pack");
}

/* K2 on tcgen05: O2 = mix(X2, Wp).  X2 [2MM][B][CinP] complex (written by pdes_dft_fwd2), O2 [2][2MM][B][Cout] complex. */
int pdes_mix_tc_fwd(const float* X2, const float* Wp, float* O2, int B, int Cin, int Cout, int m1, int m2, void* stream) {
  using namespace pdes;
#ifdef PDES_CPU_EMU
  (void)X2; (void)Wp; (void)O2; (void)B; (void)Cin; (void)Cout; (void)m1; (void)m2; (void)stream;
  set_error("pdes_mix_tc_fwd: tcgen05 path is not available in the CPU emulation build");
  return PDES_ERR_UNSUPPORTED;
#else
  PDES_REQUIRE(X2 && Wp && O2, PDES_ERR_ARG, "pdes_mix_tc_fwd: null pointer");
  PDES_REQUIRE(pdes_mix_tc_ok(B, Cin, Cout, m1, m2), PDES_ERR_UNSUPPORTED, "pdes_mix_tc_fwd: shape not supported (B=%d)", B);
  PDES_REQUIRE(aligned16(X2) && aligned16(Wp) && aligned16(O2), PDES_ERR_ARG, "pdes_mix_tc_fwd: pointers must be 16-byte aligned");
  const MtGeom g = mt_geom(B, Cin, Cout, m1, m2, mt_sms());
  MtEncodeFn enc = reinterpret_cast<MtEncodeFn>(tensor_map_encoder());
  alignas(64) CUtensorMap tw, tx;
  memset(&tw, 0, sizeof(tw));
  memset(&tx, 0, sizeof(tx));
  const int nmodes = 2 * m1 * m2;
  {
    const cuuint64_t gdim[3] = {(cuuint64_t)g.CinP * 2, (cuuint64_t)Cout, (cuuint64_t)nmodes};
    const cuuint64_t gstr[2] = {(cuuint64_t)g.CinP * 8, (cuuint64_t)g.CinP * 8 * (cuuint64_t)Cout};
    const cuuint32_t box[3] = {32, (cuuint32_t)g.to, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    const CUresult r = enc(&tw, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(Wp), gdim, gstr, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    PDES_REQUIRE(r == CUDA_SUCCESS, PDES_ERR_UNSUPPORTED, "pdes_mix_tc_fwd: cuTensorMapEncodeTiled(Wp) failed (%d)", (int)r);
  }
  {
    const cuuint64_t gdim[2] = {(cuuint64_t)g.CinP * 2, (cuuint64_t)nmodes * (cuuint64_t)B};
    const cuuint64_t gstr[1] = {(cuuint64_t)g.CinP * 8};
    const cuuint32_t box[2] = {32, (cuuint32_t)g.bp};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = enc(&tx, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(X2), gdim, gstr, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    PDES_REQUIRE(r == CUDA_SUCCESS, PDES_ERR_UNSUPPORTED, "pdes_mix_tc_fwd: cuTensorMapEncodeTiled(X2) failed (%d)", (int)r);
  }
  MtParams p;
  p.O2 = reinterpret_cast<float2*>(O2);
  p.B = B; p.Cin = Cin; p.Cout = Cout; p.CinP = g.CinP; p.nmodes = nmodes; p.npad = g.npad; p.bp = g.bp;
  p.ntile = g.ntile; p.to = g.to; p.nck = g.nck; p.per = g.per; p.nch_total = g.nch;
  const size_t raw_slot = 128 * 128 + (((size_t)g.bp * 128 + 1023) & ~size_t(1023));
  const size_t lbo = (size_t)(g.npad / 8) * 128 + 16;
  const size_t fixed = (size_t)kMtNST * 16 * lbo + 2048;
  int nraw = (int)((226 * 1024 - fixed) / raw_slot);
  if (nraw > kMtMaxRaw) nraw = kMtMaxRaw;
  PDES_REQUIRE(nraw >= 2, PDES_ERR_UNSUPPORTED, "pdes_mix_tc_fwd: not enough shared memory");
  p.nraw = nraw;
  const size_t smem = (size_t)nraw * raw_slot + (size_t)kMtNST * 16 * lbo + 1024;
  auto kfn = k_mix_tc;
  PDES_SET_SMEM(kfn, smem);
  PDES_LAUNCH(kfn, dim3((unsigned)g.G), dim3(kMtThreads), smem, stream, p, tw, tx);
  return check_launch("pdes_mix_tc_fwd");
#endif
}

/* K3a for the layout K2-on-tcgen05 writes: Z[b][h][2l+ri][c] from O2 (two partials, see pdes_mix_tc_fwd).  `Cin` is
 * the reduction width K2 ran with (it fixes which items carry a second partial). */
int pdes_inv_h_modes(const float* O2, int B, int Cin, int C, int H, int m1, int m2, const float* tables, float* Z,
                     void* stream) {
  using namespace pdes;
  PDES_REQUIRE(O2 && tables && Z, PDES_ERR_ARG, "pdes_inv_h_modes: null pointer");
  PDES_REQUIRE(B > 0 && C > 0 && H > 0 && m1 > 0 && m2 > 0 && m1 <= H && Cin > 0, PDES_ERR_ARG, "pdes_inv_h_modes: bad sizes");
  PDES_REQUIRE(B <= 65535 && m2 <= 65535, PDES_ERR_UNSUPPORTED, "pdes_inv_h_modes: grid too large");
  const MtGeom g = mt_geom(B, Cin, C, m1, m2, mt_sms());
  SplitRule rule;
  rule.nck = g.nck; rule.per = g.per; rule.ntile = g.ntile; rule.to = g.to;
  const size_t smem = ((size_t)2 * m1 * 32 + (size_t)(H / 2 + 1) * 2 * m1) * sizeof(float2);
  PDES_REQUIRE(smem <= (size_t)kMaxDynSmem, PDES_ERR_UNSUPPORTED, "pdes_inv_h_modes: needs %zu B of shared memory", smem);
  auto kfn = k_inv_h2;
  if (smem > 48 * 1024) PDES_SET_SMEM(kfn, smem);
  PDES_LAUNCH(kfn, dim3((unsigned)ceil_div(C, 32), (unsigned)m2, (unsigned)B), dim3(kIh2Threads), smem, stream,
              reinterpret_cast<const float2*>(O2), rule, B, C, H, m1, m2, tables /* twh [H][2] sits at offset 0 of the blob */, Z);
  return check_launch("pdes_inv_h_modes");
}

}  // extern "C"
