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
//   Wp  packed master copy  [m][tile][chunk][row][16 i][re|im] fp32 (tile = <= 128 output channels, chunk = 16 input
//       channels, i padded to a multiple of 16 with zeros): one contiguous 128-byte-row block per chunk, in the order
//       the kernel streams them; rows of the first weight block that the reference overwrites when 2*m1 > H
//       (proc_fno.py:266-269) are stored as zeros.  Built by
//       pdes_mix_tc_pack() once per weight version (the caller caches it); the parameters themselves, Adam and the
//       gradient all-reduce keep the reference layout [Cin][Cout][m1][m2].
//   X2  [m][b][i_pad] complex, mode-major copy of the retained spectrum written by K1 (pad columns are never read
//       unmasked);
//   O2  [2][m][b][o] complex: partial 0 and partial 1 (the part of a work item that a second CTA finished; zero for the
//       items that one CTA computed completely), summed by K3a.
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
#include <cstdlib>
#include <cstring>
#endif

namespace pdes {


namespace {

// ------------------------------------------------------------------------------------------------- weight pack
// [Cin][Cout][m1*m2] complex (x2 blocks) -> Wp[m][tile][chunk][row][16 i][re|im]: one (mode, output-channel tile, 16-channel
// chunk) = `to` rows x 128 bytes = ONE CONTIGUOUS block, and the blocks sit in exactly the order the kernel streams them,
// so every CTA reads one contiguous region of HBM (strided 128-byte rows reached only 2.4 TB/s: DRAM page misses).
// Per output channel this is a 2-D transpose between the input-channel and the mode index: a 32 x 32 tile goes through
// shared memory so that the loads (along the mode index) and the stores (128-byte segments along i) are coalesced.
__global__ void __launch_bounds__(256)
k_mix_tc_pack(const float2* __restrict__ w1, const float2* __restrict__ w2, float2* __restrict__ Wp, int Cin, int Cout,
              int CinP, int m1, int m2, int H, int ntile, int to) {
  __shared__ float2 tile[32][33];
  const int MM = m1 * m2;
  const int o = blockIdx.x;                                             // 0 .. ntile*to - 1 (rows >= Cout are zero padding)
  const int i0 = blockIdx.y * 32;
  const int mt_per_half = (MM + 31) / 32;
  const int half = blockIdx.z / mt_per_half, mm0 = (blockIdx.z % mt_per_half) * 32;
  const float2* w = half ? w2 : w1;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;            // 32 x 8
  for (int r = ty; r < 32; r += 8) {                                  // r = input channel, tx = mode
    const int i = i0 + r, mm = mm0 + tx;
    float2 v = make_float2(0.f, 0.f);
    if (i < Cin && mm < MM && o < Cout) v = __ldg(w + ((size_t)i * Cout + o) * MM + mm);
    tile[r][tx] = v;
  }
  __syncthreads();
  const int nck = CinP / kMtBK;
  const int t = o / to, row = o - t * to;
  for (int r = ty; r < 32; r += 8) {                                  // r = mode, tx = input channel
    const int mm = mm0 + r, i = i0 + tx;
    if (mm < MM && i < CinP) {
      const int k = half * m1 + mm / m2;
      float2 v = tile[tx][r];
      if (row_dead(k, m1, H)) v = make_float2(0.f, 0.f);
      const int m = half * MM + mm, kc = i / kMtBK, ii = i - kc * kMtBK;
      Wp[((((size_t)m * ntile + t) * nck + kc) * to + row) * kMtBK + ii] = v;
    }
  }
}

// ------------------------------------------------------------------------------------------------- K3a (v2)
// Z[b][h][2l + {re,im}][o] = sum_k e^{+2 pi i kx_k h / H} (O2[0][m][b][o] + O2[1][m][b][o]),  m = k*m2 + l.
// CTA = (32 output channels, one l, one sample): lanes run along o (coalesced loads of O2 and stores of Z), the four
// warps take the row pairs (h, H-h) round-robin.  A pair shares its four real sums:
//     P = sum_k cos(t_k h) O_k,  Q = sum_k sin(t_k h) O_k  =>  z(h) = P + iQ,  z(H-h) = P - iQ,
// which halves the multiplies for any (H, m1).  Twiddles come from a per-CTA shared table (warp-broadcast loads).
// Rows +kx and -kx are folded as well: with S_j = O[kx=j] + O[kx=-j], D_j = O[kx=j] - O[kx=-j] (j = 0..m1; j = 0 has no
// partner, j = m1 only the negative one) the sums run over m1 + 1 terms instead of 2*m1:
//     P = sum_j cos(2 pi j h/H) S_j,  Q = sum_j sin(2 pi j h/H) D_j.
// Each thread keeps HP = 8 row pairs in registers (32 accumulators), so one spectrum load and four 128-bit twiddle loads
// feed 32 FMAs.
constexpr int kIh2HP = 8;                                // row pairs per thread
constexpr int kIh2Warps = 3;
// (A persistent, cp.async double-buffered variant of this kernel was measured on B200 and was not faster -- 23.0 vs
// 20.5 us in the cold-cache ncu pass -- so the simple one-CTA-per-item form stays.)
__global__ void __launch_bounds__(32 * kIh2Warps)
k_inv_h2(const float2* __restrict__ O2, int B, int C, int H, int m1, int m2,
         const float2* __restrict__ twp_g, float* __restrict__ Z) {
  PDES_DYN_SMEM(float2, sm2);
  const int K = 2 * m1, M2 = K * m2, J = 2 * m2, NJ = m1 + 1;
  const int npair = H / 2 + 1;
  const int npp = (npair + kIh2HP - 1) / kIh2HP * kIh2HP;        // pairs padded to the register block
  float2* Ss = sm2;                                    // [NJ][32]
  float2* Ds = Ss + (size_t)NJ * 32;                   // [NJ][32]
  float2* tw = Ds + (size_t)NJ * 32;                   // [NJ][npp]  (cos, sin)(2 pi j p / H)
  const int tid = threadIdx.x, nthr = blockDim.x, lane = tid & 31, wq = tid >> 5, nwarp = nthr >> 5;
  const int o0 = blockIdx.x * 32, l = blockIdx.y, b = blockIdx.z;
  const int o = o0 + lane;
  // every global load of the block is issued before its one barrier
  for (int idx = tid; idx < NJ * npp; idx += nthr) tw[idx] = __ldg(twp_g + idx);     // host-built table, coalesced
  PDES_GRID_DEP_WAIT();                                // O2 is written by the previous kernel of the chain (K2)
  const size_t pstride = (size_t)M2 * B * C;
  auto load_o = [&](int k, int oo) -> float2 {         // O[k][l] of channel oo = partial 0 + partial 1 (zero unless split)
    const size_t at = ((size_t)(k * m2 + l) * B + b) * C + oo;
    const float2 v = __ldg(O2 + at), u = __ldg(O2 + pstride + at);
    return make_float2(v.x + u.x, v.y + u.y);
  };
  for (int idx = tid; idx < NJ * 32; idx += nthr) {
    const int j = idx >> 5, oo = o0 + (idx & 31);
    float2 sv = make_float2(0.f, 0.f), dv = make_float2(0.f, 0.f);
    if (oo < C) {
      float2 a = make_float2(0.f, 0.f), c = make_float2(0.f, 0.f);
      if (j < m1) a = load_o(j, oo);                   // +kx = j lives in row k = j
      if (j > 0) c = load_o(K - j, oo);                // -kx = -j lives in row k = 2*m1 - j
      sv = make_float2(a.x + c.x, a.y + c.y);
      dv = make_float2(a.x - c.x, a.y - c.y);
    }
    Ss[idx] = sv;
    Ds[idx] = dv;
  }
  __syncthreads();
  PDES_GRID_DEP_LAUNCH();                              // K3b may be scheduled on an SM as soon as its K3a CTAs are gone
  for (int p0 = wq * kIh2HP; p0 < npair; p0 += nwarp * kIh2HP) {
    float pr[kIh2HP], pi[kIh2HP], qr[kIh2HP], qi[kIh2HP];
#pragma unroll
    for (int e = 0; e < kIh2HP; ++e) pr[e] = pi[e] = qr[e] = qi[e] = 0.0f;
    for (int j = 0; j < NJ; ++j) {
      const float2 sv = Ss[j * 32 + lane], dv = Ds[j * 32 + lane];
      const float4* t4 = reinterpret_cast<const float4*>(tw + (size_t)j * npp + p0);       // warp-uniform: broadcast
#pragma unroll
      for (int e = 0; e < kIh2HP; e += 2) {
        const float4 t = t4[e >> 1];                                       // (cos, sin) of pairs p0+e, p0+e+1
        pr[e] = fmaf(t.x, sv.x, pr[e]);         pi[e] = fmaf(t.x, sv.y, pi[e]);
        qr[e] = fmaf(t.y, dv.x, qr[e]);         qi[e] = fmaf(t.y, dv.y, qi[e]);
        pr[e + 1] = fmaf(t.z, sv.x, pr[e + 1]); pi[e + 1] = fmaf(t.z, sv.y, pi[e + 1]);
        qr[e + 1] = fmaf(t.w, dv.x, qr[e + 1]); qi[e + 1] = fmaf(t.w, dv.y, qi[e + 1]);
      }
    }
    if (o < C) {
      float* const zb = Z + ((size_t)b * H * J + 2 * l) * C + o;          // row h of this (sample, l, channel): zb + h * J * C
      const size_t rs = (size_t)J * C;
#pragma unroll
      for (int e = 0; e < kIh2HP; ++e) {
        const int h = p0 + e;
        if (h < npair) {
          float* z = zb + (size_t)h * rs;
          z[0] = pr[e] - qi[e];                                           // Re (P + iQ)
          z[C] = pi[e] + qr[e];
          const int h2 = H - h;
          if (h != 0 && h2 != h) {
            float* z2 = zb + (size_t)h2 * rs;
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
// Facts measured on B200 that shape this kernel (tools/ubench_mma2.cu, tools/ablate_mix.py, profiles/r02_k2_*):
//   * kind::tf32 TRUNCATES its operands (feeding unmasked fp32 bits gives bit-identical results), so the "hi" half of
//     the 3xTF32 split needs no conversion at all: the MMA reads the raw fp32 tile the TMA delivered;
//   * one thread cannot issue a small-N tcgen05.mma faster than every ~49 cycles (N = 32: tensor-pipe floor 16), four
//     issuing warps reach ~18 cycles per MMA;
//   * every scalar instruction of a single-thread role is latency-exposed (a runtime integer division ~150 cycles), and
//     one loop iteration of any role costs several hundred cycles of mbarrier round trips.
// Hence: the K index of the GEMM is the interleaved (input channel, re|im) pair exactly as it sits in the packed
// weights, A_hi = the raw 128-byte-swizzled TMA tile in shared memory, only A_lo = W - trunc(W) goes through registers
// into tensor memory; 16 convert warps / 4 issuing warps each own every fourth chunk (and its two ring stages), and all
// per-chunk index arithmetic is strength-reduced to counters.
//
//   D[o, n] = sum_{k = (i, re|im)} Wp[o][k] * Bx[n][k],   Bx[(b,re)][(i,re|im)] = (Xr, -Xi),  Bx[(b,im)][(i,re|im)] = (Xi, Xr)
constexpr int kMtCvtWarps = 16, kMtEpi0 = 16, kMtIssue0 = 20, kMtNIssue = 4, kMtTma0 = 24;
constexpr int kMtThreads = 32 * 28;
constexpr int kMtMaxStages = 8;

struct MtBars {
  unsigned long long raw_full[kMtMaxStages], full[kMtMaxStages], empty[kMtMaxStages], acc_full[2], acc_empty[2];
};

struct MtParams {
  float2* O2;
  int B, Cin, Cout, CinP, nmodes, npad, bp, ntile, to, nck, per, nch_total, nst, nbuf;
  uint32_t stage_bytes, x_off, b_off;
  int dbg;       // diagnostic builds (-DPDES_MT_ABLATE) only: 1 no MMA, 2 no TMA, 4 no convert work, 8 no epilogue stores
};
#ifdef PDES_MT_ABLATE
#define MT_DBG(bit) ((p.dbg & (bit)) != 0)
__device__ long long g_mt_trace[3 * 64 * 4];          // [role: 0 producer, 1 convert, 2 MMA][chunk < 64][4 stamps], CTA 0
#define MT_TRACE(role, g, k) do { if (blockIdx.x == 0 && (g) < 64) g_mt_trace[((role) * 64 + (g)) * 4 + (k)] = clock64(); } while (0)
#else
#define MT_DBG(bit) false
#define MT_TRACE(role, g, k) do { } while (0)
#endif

__device__ __forceinline__ float mt_hi(float x) { return __uint_as_float(__float_as_uint(x) & 0xffffe000u); }
// shared-memory matrix descriptor, K-major, 128-byte swizzle (8-row x 128-byte atoms 1024 bytes apart), sm_100 version
__device__ __forceinline__ uint64_t mt_desc_sw128(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;                      // leading byte offset: unused for a K extent within one swizzle span
  d |= (uint64_t)(1024 >> 4) << 32;            // stride byte offset: next 8-row group
  d |= (uint64_t)1 << 46;                      // descriptor version 1 (Blackwell)
  d |= (uint64_t)2 << 61;                      // SWIZZLE_128B
  return d;
}

__global__ void __launch_bounds__(kMtThreads, 1)
k_mix_tc(MtParams p, const __grid_constant__ CUtensorMap tmap_w, const __grid_constant__ CUtensorMap tmap_x) {
  PDES_DYN_SMEM(unsigned char, smem_raw);
  unsigned char* base = ptx::align_smem_1024(smem_raw);
  const int npad = p.npad;
  const uint32_t lbo = (uint32_t)(npad / 8) * 128u + 16u;                  // +16: the k-quads of a row fall in different banks
  const uint32_t blk = 8u * lbo;                                           // one canonical [npad x 32] block (8 k-quads)
  const uint32_t stage_bytes = p.stage_bytes, x_off = p.x_off, b_off = p.b_off;   // stage: [W raw 16 KB][X raw][B_hi][B_lo]
  __shared__ __align__(8) MtBars bars;
  __shared__ uint32_t tmem_slot;

  const int tid = threadIdx.x, warp = ptx::uniform_warp_idx(), lane = tid & 31;
  const int NST = p.nst;                                                    // 8 or 4 (a multiple of the 4 chunk owners)
  const int c_beg = (int)blockIdx.x * p.per;
  const int c_end = (c_beg + p.per < p.nch_total) ? (c_beg + p.per) : p.nch_total;
  const int nloc = c_end - c_beg;
  const int item0 = c_beg / p.nck, kc_first = c_beg - item0 * p.nck;       // the only runtime divisions: once per thread
  const int nbuf = p.nbuf;

  if (tid == 0) {
    for (int i = 0; i < NST; ++i) { ptx::mbar_init(&bars.raw_full[i], 1); ptx::mbar_init(&bars.full[i], 4); ptx::mbar_init(&bars.empty[i], 1); }
    for (int i = 0; i < 2; ++i) { ptx::mbar_init(&bars.acc_full[i], kMtNIssue); ptx::mbar_init(&bars.acc_empty[i], 4); }
    ptx::fence_mbar_init();
  }
  if (warp == kMtIssue0) {
    ptx::tmem_alloc(&tmem_slot, 512);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  // TMEM columns: accumulator of issuer q in buffer a at (a * 4 + q) * npad (<= 256 in total); A_lo of stage s at 256 + 32*s
  const uint32_t ta0 = 256u;

  if (warp < kMtCvtWarps) {
    // ================================================================== convert: group q = chunks g = q (mod 4)
    const int q = warp >> 2, cw = warp & 3;
    const int row = tid & 127;                                             // output channel within the tile = TMEM lane
    const uint32_t sw = (uint32_t)(row & 7);
    const uint32_t trow0 = tmem_base + ((uint32_t)(cw * 32) << 16) + ta0;
    const int ii = row & 15, bb0 = row >> 4;                               // this thread's spectrum elements: (bb0 + 8 t, ii)
    // Bx element (n, k = 2 ii + ri) at (k/4)*lbo + (n/8)*128 + (n%8)*16 + (k%4)*4; rows n = 2 bb, 2 bb + 1 share an 8-row group
    const uint32_t kof = (uint32_t)(ii >> 1) * lbo + (uint32_t)(ii & 1) * 8u;
    int s = q;                                                             // ring stage of chunk g = q, q + 4, ...
    uint32_t ph = 0;                                                       // parity of raw_full[s]
    int kc = kc_first + q;
    while (kc >= p.nck) kc -= p.nck;
    for (int g = q; g < nloc; g += 4) {
      if (cw == 0 && lane == 0) MT_TRACE(1, g, 0);
      ptx::mbar_wait(&bars.raw_full[s], ph);                                // (the producer already waited for empty[s])
      if (cw == 0 && lane == 0) MT_TRACE(1, g, 1);
      ptx::tc_fence_after();
      unsigned char* st = base + (uint32_t)s * stage_bytes;
      if (!MT_DBG(4)) {
        // ---- A_lo = W - trunc(W): this thread's 128-byte row (k = 0..31), 128-byte swizzle: chunk j sits at j ^ (row & 7)
        const unsigned char* wrow = st + (uint32_t)row * 128u;
        const uint32_t trow = trow0 + (uint32_t)s * 32u;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          uint32_t lo[16];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float4 v = *reinterpret_cast<const float4*>(wrow + ((((uint32_t)(half * 4 + j)) ^ sw) << 4));
            lo[4 * j + 0] = __float_as_uint(v.x - mt_hi(v.x));
            lo[4 * j + 1] = __float_as_uint(v.y - mt_hi(v.y));
            lo[4 * j + 2] = __float_as_uint(v.z - mt_hi(v.z));
            lo[4 * j + 3] = __float_as_uint(v.w - mt_hi(v.w));
          }
          ptx::tmem_st16(trow + (uint32_t)half * 16u, lo);
        }
        // ---- spectrum rows -> Bx_hi (raw values: the tensor core truncates) and Bx_lo, canonical K-major blocks
        const unsigned char* xraw = st + x_off;
        unsigned char* sb = st + b_off;
        const bool kvalid = kc * kMtBK + ii < p.Cin;
        for (int bb = bb0; bb < p.bp; bb += 8) {
          float2 x = *reinterpret_cast<const float2*>(xraw + (uint32_t)bb * 128u + (uint32_t)ii * 8u);
          if (bb >= p.B || !kvalid) x = make_float2(0.f, 0.f);
          const float xr_l = x.x - mt_hi(x.x), xi_l = x.y - mt_hi(x.y);
          const int n0 = 2 * bb;
          const uint32_t o0f = kof + (uint32_t)(n0 >> 3) * 128u + (uint32_t)(n0 & 7) * 16u;
          *reinterpret_cast<float2*>(sb + o0f) = make_float2(x.x, -x.y);            // row (b, re): (Xr, -Xi)
          *reinterpret_cast<float2*>(sb + o0f + 16u) = make_float2(x.y, x.x);       // row (b, im): (Xi,  Xr)
          *reinterpret_cast<float2*>(sb + blk + o0f) = make_float2(xr_l, -xi_l);
          *reinterpret_cast<float2*>(sb + blk + o0f + 16u) = make_float2(xi_l, xr_l);
        }
      }
      ptx::tmem_st_wait();
      ptx::tc_fence_before();
      ptx::fence_proxy_async();                                              // the Bx blocks went through st.shared
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&bars.full[s]);
      if (cw == 0 && lane == 0) MT_TRACE(1, g, 3);
      s += 4; if (s >= NST) { s -= NST; ph ^= 1u; }
      kc += 4; while (kc >= p.nck) kc -= p.nck;
    }
  } else if (warp >= kMtIssue0 && warp < kMtIssue0 + kMtNIssue) {
    {
      // ================================================================ MMA issue: issuer q = chunks g = q (mod 4), accumulator q
      // (the whole warp runs this loop converged with warp-uniform values; one elected lane issues, see pdes_ptx.cuh)
      const int q = warp - kMtIssue0;
      const uint32_t idesc = ptx::idesc_tf32(128, npad);
      const uint32_t sbase = ptx::smem_u32(base);
      const uint64_t kstepb = (uint64_t)((2u * lbo) >> 4);                   // B: 8 k = two k-quads
      // Every issuer walks EVERY part of the CTA's range with the same protocol -- wait for the accumulator buffer to be
      // drained, issue its chunks of the part (possibly none), commit acc_full -- so an arrival for part P can never land
      // on acc_full[a] before the phase of part P - nbuf has completed (an issuer without chunks in a short part would
      // otherwise run ahead and complete that phase early: seen as a rare wrong result with nck < 4).
      int s = q;
      uint32_t fph = 0;
      int g = q;                                                             // next local chunk of this issuer
      int pb = 0, pe = p.nck - kc_first;                                     // local chunk range [pb, pe) of the current part
      for (int part = 0; pb < nloc; ++part) {
        if (pe > nloc) pe = nloc;
        const uint32_t a = (nbuf == 2) ? (uint32_t)(part & 1) : 0u;
        if (part >= nbuf) {
          ptx::mbar_wait(&bars.acc_empty[a], (uint32_t)((((nbuf == 2) ? (part >> 1) : part) - 1) & 1));
          ptx::tc_fence_after();
        }
        const uint32_t dacc = tmem_base + (a * 4u + (uint32_t)q) * (uint32_t)npad;
        bool fresh = true;
        for (; g < pe; g += kMtNIssue) {
          if (q == 0 && lane == 0) MT_TRACE(2, g, 0);
          ptx::mbar_wait(&bars.full[s], fph);
          if (q == 0 && lane == 0) MT_TRACE(2, g, 1);
          ptx::tc_fence_after();
          const uint32_t sst = sbase + (uint32_t)s * stage_bytes;
          const uint64_t a_hi = mt_desc_sw128(sst);                          // + 2 (= 32 bytes) per k-step of 8
          const uint64_t b_hi = ptx::smem_desc_noswizzle(sst + b_off, lbo, 128), b_lo = b_hi + (uint64_t)(blk >> 4);
          const uint32_t a_lo = tmem_base + ta0 + (uint32_t)s * 32u;
          if (!MT_DBG(1)) {
            ptx::mma_tf32_ws(dacc, a_hi, b_hi, idesc, fresh ? 0u : 1u);                         // hi * hi
            ptx::mma_tf32_ws(dacc, a_hi + 2, b_hi + kstepb, idesc, 1u);
            ptx::mma_tf32_ws(dacc, a_hi + 4, b_hi + 2 * kstepb, idesc, 1u);
            ptx::mma_tf32_ws(dacc, a_hi + 6, b_hi + 3 * kstepb, idesc, 1u);
            ptx::mma_tf32_ws(dacc, a_hi, b_lo, idesc, 1u);                                      // hi * lo
            ptx::mma_tf32_ws(dacc, a_hi + 2, b_lo + kstepb, idesc, 1u);
            ptx::mma_tf32_ws(dacc, a_hi + 4, b_lo + 2 * kstepb, idesc, 1u);
            ptx::mma_tf32_ws(dacc, a_hi + 6, b_lo + 3 * kstepb, idesc, 1u);
            ptx::mma_tf32_ta_ws(dacc, a_lo, b_hi, idesc, 1u);                                   // lo * hi
            ptx::mma_tf32_ta_ws(dacc, a_lo + 8, b_hi + kstepb, idesc, 1u);
            ptx::mma_tf32_ta_ws(dacc, a_lo + 16, b_hi + 2 * kstepb, idesc, 1u);
            ptx::mma_tf32_ta_ws(dacc, a_lo + 24, b_hi + 3 * kstepb, idesc, 1u);
          }
          ptx::tc_commit_ws(&bars.empty[s]);
          if (q == 0 && lane == 0) MT_TRACE(2, g, 2);
          fresh = false;
          s += 4; if (s >= NST) { s -= NST; fph ^= 1u; }
        }
        ptx::tc_commit_ws(&bars.acc_full[a]);
        pb = pe;
        pe += p.nck;
      }
    }
  } else if (warp >= kMtTma0 && warp < kMtTma0 + 4) {
    {
      // ================================================================ TMA producers: producer q = chunks g = q (mod 4)
      // (a single producer thread needs ~600 cycles per chunk -- wait, expect_tx, two TMA issues, counters -- and paced
      // the whole kernel; four of them make the pipeline four independent lanes producer -> convert group -> issuer)
      const int q = warp - kMtTma0;
      const uint32_t wbytes = (uint32_t)p.to * 128u, xbytes = (uint32_t)p.bp * 128u;
      int m = item0 / p.ntile, t = item0 - m * p.ntile, kc = kc_first + q;
      while (kc >= p.nck) { kc -= p.nck; if (++t == p.ntile) { t = 0; ++m; } }
      int s = q;
      const uint64_t pol_stream = ptx::l2_policy_evict_first();
      uint32_t eph = 1;                                                      // parity of empty[s] (first pass: no wait)
      // The weights do not depend on the previous kernel of the chain (K1 writes only the spectrum X2): the first pass over
      // the ring streams its weight blocks BEFORE the grid-dependency wait (programmatic dependent launch), so the ~3 us
      // until the first bytes arrive overlap K1's tail; the spectrum boxes of those stages follow after the wait.
      if (!MT_DBG(2)) {
        int sa = q;
        for (int g = q; g < nloc && g < NST; g += 4, sa += 4) {
          ptx::mbar_arrive_expect_tx_ws(&bars.raw_full[sa], wbytes + xbytes);
          ptx::tma_load_2d_ws_hint(base + (uint32_t)sa * stage_bytes, &tmap_w, 0, (c_beg + g) * p.to, &bars.raw_full[sa], pol_stream);
        }
      }
      PDES_GRID_DEP_WAIT();
      for (int g = q; g < nloc; g += 4) {
        if (q == 0 && lane == 0) MT_TRACE(0, g, 0);
        if (g >= NST) ptx::mbar_wait(&bars.empty[s], eph);
        if (q == 0 && lane == 0) MT_TRACE(0, g, 1);
        unsigned char* st = base + (uint32_t)s * stage_bytes;
        if (MT_DBG(2)) {
          ptx::mbar_arrive_ws(&bars.raw_full[s]);
        } else {
          if (g >= NST) {
            ptx::mbar_arrive_expect_tx_ws(&bars.raw_full[s], wbytes + xbytes);
            // the packed weights are read exactly once per forward (59 MB): evict-first, so that they do not push the
            // block input out of L2 between K1 and K3b
            ptx::tma_load_2d_ws_hint(st, &tmap_w, 0, (c_beg + g) * p.to, &bars.raw_full[s], pol_stream);   // one contiguous block
          }
          ptx::tma_load_2d_ws(st + x_off, &tmap_x, kc * 2 * kMtBK, m * p.B, &bars.raw_full[s]);
        }
        if (q == 0 && lane == 0) MT_TRACE(0, g, 2);
        s += 4; if (s >= NST) { s -= NST; eph ^= 1u; }
        kc += 4;
        while (kc >= p.nck) { kc -= p.nck; if (++t == p.ntile) { t = 0; ++m; } }
      }
    }
  } else if (warp >= kMtEpi0 && warp < kMtEpi0 + 4) {
    // ==================================================================== epilogue: lane = output channel
    const int quad = warp & 3;
    const int rowl = quad * 32 + lane;
    const size_t pstride = (size_t)p.nmodes * p.B * p.Cout;
    uint32_t part = 0;
    int m = item0 / p.ntile, t = item0 - m * p.ntile, kc0 = kc_first;
    for (int c = c_beg; c < c_end;) {
      int cl = c + (p.nck - kc0);                                            // end of this part
      if (cl > c_end) cl = c_end;
      const int o = t * p.to + rowl;
      const bool valid = rowl < p.to && o < p.Cout;
      const uint32_t a = (nbuf == 2) ? (part & 1u) : 0u;
      ptx::mbar_wait(&bars.acc_full[a], ((nbuf == 2) ? (part >> 1) : part) & 1u);
      ptx::tc_fence_after();
      const uint32_t tb = tmem_base + ((uint32_t)(quad * 32) << 16) + a * 4u * (uint32_t)npad;
      float2* dst = p.O2 + (kc0 != 0 ? pstride : 0) + ((size_t)m * p.B) * p.Cout + o;   // partial 1 = continuation of a split item
      // a part that covers its whole item also zeroes partial 1, so K3a adds the two partials unconditionally
      float2* zero1 = (kc0 == 0 && cl - c == p.nck) ? dst + pstride : nullptr;
      // accumulator q received the chunks g = q (mod 4) of this part; a part shorter than 4 chunks leaves some untouched
      const int g0 = c - c_beg, nch_part = cl - c;
      bool used[4];
#pragma unroll
      for (int qq = 0; qq < 4; ++qq) used[qq] = nch_part >= 4 || (((qq - g0) & 3) < nch_part);
      for (int n0 = 0; n0 < npad; n0 += 8) {
        uint32_t v0[8], v1[8], v2[8], v3[8];
        ptx::tmem_ld8(tb + (uint32_t)n0, v0);
        ptx::tmem_ld8(tb + (uint32_t)(npad + n0), v1);
        ptx::tmem_ld8(tb + (uint32_t)(2 * npad + n0), v2);
        ptx::tmem_ld8(tb + (uint32_t)(3 * npad + n0), v3);
        ptx::tmem_ld_wait();
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          if (!used[0]) v0[e] = 0u;
          if (!used[1]) v1[e] = 0u;
          if (!used[2]) v2[e] = 0u;
          if (!used[3]) v3[e] = 0u;
        }
        if (n0 + 8 >= npad) {                                                // accumulators fully read: hand them back
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive(&bars.acc_empty[a]);
        }
        if (valid && !MT_DBG(8)) {
#pragma unroll
          for (int e = 0; e < 8; e += 2) {
            const int bb = (n0 + e) >> 1;
            if (bb < p.B) {
              const float re = (__uint_as_float(v2[e]) + __uint_as_float(v3[e])) + (__uint_as_float(v1[e]) + __uint_as_float(v0[e]));
              const float im = (__uint_as_float(v2[e + 1]) + __uint_as_float(v3[e + 1])) + (__uint_as_float(v1[e + 1]) + __uint_as_float(v0[e + 1]));
              dst[(size_t)bb * p.Cout] = make_float2(re, im);
              if (zero1 != nullptr) zero1[(size_t)bb * p.Cout] = make_float2(0.f, 0.f);
            }
          }
        }
      }
      ++part;
      c = cl;
      kc0 = 0;
      if (++t == p.ntile) { t = 0; ++m; }
    }
  }
  PDES_GRID_DEP_LAUNCH();                      // K3a may be scheduled as this CTA drains
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == kMtIssue0) {
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

#if defined(PDES_MT_ABLATE) && !defined(PDES_CPU_EMU)
int pdes_mt_trace_read(long long* host) { return (int)cudaMemcpyFromSymbol(host, pdes::g_mt_trace, sizeof(long long) * 3 * 64 * 4); }
#endif

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
  return (size_t)2 * m1 * m2 * pdes::mt_ntile(Cout) * pdes::mt_to(Cout) * pdes::mt_cinp(Cin) * 2;
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
  const int CinP = mt_cinp(Cin), MM = m1 * m2, ntile = mt_ntile(Cout), to = mt_to(Cout);
  PDES_REQUIRE(Cout <= 65535, PDES_ERR_UNSUPPORTED, "pdes_mix_tc_pack: too many output channels");
  auto kfn = k_mix_tc_pack;
  const dim3 grid((unsigned)(ntile * to), (unsigned)ceil_div(CinP, 32), (unsigned)(2 * ceil_div(MM, 32)));
  PDES_LAUNCH(kfn, grid, dim3(256), 0, stream, reinterpret_cast<const float2*>(w1), reinterpret_cast<const float2*>(w2),
              reinterpret_cast<float2*>(Wp), Cin, Cout, CinP, m1, m2, H, ntile, to);
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
    const cuuint64_t gdim[2] = {32, (cuuint64_t)g.nch * (cuuint64_t)g.to};      // [chunk][row][16 i x (re,im)]: rows of 128 bytes
    const cuuint64_t gstr[1] = {128};
    const cuuint32_t box[2] = {32, (cuuint32_t)g.to};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = enc(&tw, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(Wp), gdim, gstr, box, estr,
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
  p.nbuf = (8 * g.npad <= 256) ? 2 : 1;                                    // double-buffered accumulators when they fit
  p.dbg = 0;
#ifdef PDES_MT_ABLATE
  if (const char* e = getenv("PDES_MT_DBG")) p.dbg = atoi(e);
#endif
  const size_t lbo = (size_t)(g.npad / 8) * 128 + 16;
  const size_t x_off = 128 * 128, b_off = x_off + (((size_t)g.bp * 128 + 127) & ~size_t(127));
  const size_t stage_bytes = (b_off + 2 * 8 * lbo + 1023) & ~size_t(1023);
  p.stage_bytes = (uint32_t)stage_bytes; p.x_off = (uint32_t)x_off; p.b_off = (uint32_t)b_off;
  p.nst = (kMtMaxStages * stage_bytes + 2048 <= 227 * 1024) ? kMtMaxStages : 4;
  PDES_REQUIRE((size_t)p.nst * stage_bytes + 2048 <= 227 * 1024, PDES_ERR_UNSUPPORTED, "pdes_mix_tc_fwd: not enough shared memory");
  const size_t smem = (size_t)p.nst * stage_bytes + 1024;
  auto kfn = k_mix_tc;
  PDES_SET_SMEM(kfn, smem);
  PDES_LAUNCH_PDL(kfn, dim3((unsigned)g.G), dim3(kMtThreads), smem, stream, p, tw, tx);
  return check_launch("pdes_mix_tc_fwd");
#endif
}

/* K3a for the layout K2-on-tcgen05 writes: Z[b][h][2l+ri][c] from O2 (two partials, see pdes_mix_tc_fwd). */
int pdes_inv_h_modes(const float* O2, int B, int C, int H, int m1, int m2, const float* tables, float* Z, void* stream) {
  using namespace pdes;
  PDES_REQUIRE(O2 && tables && Z, PDES_ERR_ARG, "pdes_inv_h_modes: null pointer");
  PDES_REQUIRE(B > 0 && C > 0 && H > 0 && m1 > 0 && m2 > 0 && m1 <= H, PDES_ERR_ARG, "pdes_inv_h_modes: bad sizes");
  const int npair = H / 2 + 1, npp = ceil_div(npair, kIh2HP) * kIh2HP;
  const size_t smem = ((size_t)2 * (m1 + 1) * 32 + (size_t)(m1 + 1) * npp) * sizeof(float2);
  PDES_REQUIRE(smem <= (size_t)kMaxDynSmem && (long)H * m1 < (1L << 31), PDES_ERR_UNSUPPORTED,
               "pdes_inv_h_modes: needs %zu B of shared memory", smem);
  PDES_REQUIRE(B <= 65535 && m2 <= 65535, PDES_ERR_UNSUPPORTED, "pdes_inv_h_modes: grid too large");
  int nwarp = ceil_div(npair, kIh2HP);
  // Three warps per CTA (each takes every third block of 8 row pairs): at the shipped shape the 960 CTAs are then all
  // resident at once (7 per SM by registers).  With one warp per block (7 warps) only 592 fit and the grid ran as 1.6
  // waves: 16.6 -> 14.5 us back to back, 4 us off the chain in situ (tools/sweep_k3a.py).
  if (nwarp > kIh2Warps) nwarp = kIh2Warps;
  auto kfn = k_inv_h2;
  if (smem > 48 * 1024) PDES_SET_SMEM(kfn, smem);
  PDES_MAX_CARVEOUT(kfn);
  PDES_LAUNCH_PDL(kfn, dim3((unsigned)ceil_div(C, 32), (unsigned)m2, (unsigned)B), dim3((unsigned)(32 * nwarp)), smem, stream,
              reinterpret_cast<const float2*>(O2), B, C, H, m1, m2,
              reinterpret_cast<const float2*>(tables + table_layout(H, 2, m1, m2).twp), Z);
  return check_launch("pdes_inv_h_modes");
}

}  // extern "C"
