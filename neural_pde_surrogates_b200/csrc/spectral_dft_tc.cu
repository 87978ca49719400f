// K1 with BOTH DFT stages on the 5th-generation tensor cores (tcgen05 + TMEM + TMA), for grids of width 64.
// OPT-IN (PDES_K1_TC=1): correct (3xTF32, ~1.5e-6 rel. L2) but, measured on B200, SLOWER than the FFMA kernel at the
// shipped shape (35 us vs 27 us at B = 16) -- see "Result" below.  Kept because it pins down how MN-major tf32 operands
// work on sm_100a, which the next kernels can use.
//
//   reference: torch.fft.rfft2(x) followed by the two row slices and the column slice, proc_fno.py:261,267,269
//
// Stage 1, the column DFT  P[kx][w] = sum_h x[h][w] cos(2 pi kx h / H),  Q = same with sin,  kx = 0 .. m1, is the GEMM
//     D1[(image, w)][(kx, cos|sin)] = sum_h  x^T[(image, w)][h] * E[h][(kx, cos|sin)]
// with M = 2 images x 64 columns = 128 TMEM lanes, N = 32 (P in columns 0..15, Q in 16..31), K = H:
//   * A = x^T is MN-major (the 64 pixels of a row are contiguous).  For tf32 an MN-major shared-memory operand must use
//     the SWIZZLE_128B_BASE32B layout type (descriptor layout 1; with the ordinary SWIZZLE_128B it multiplies as zeros):
//     128-byte rows of 32 M-elements per K index whose four 32-byte chunks are XOR-ed with (row % 4), which is what a TMA
//     box {32 floats, H rows} with CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B writes.  So the fp32 image is multiplied where it
//     lands: kind::tf32 truncates its operands, the raw tile is the "hi" half of the 3xTF32 split, and only the residual
//     lo = x - trunc(x) is computed by threads and written to TENSOR MEMORY (tcgen05.st) as a TMEM A operand;
//   * B = E is a constant K-major table (hi / lo halves) built once per CTA from the host's float64 twiddles.
// Stage 2, the row DFT to the m2 kept columns, is the GEMM
//     D2[(image, kx, P|Q)][(l, cos|sin)] = sum_w  D1^T[(kx, P|Q)][w] * F[w][(l, cos|sin)]:
// lane (image, w) of D1 holds exactly row k = w of the MN-major operand D1^T, so the accumulator goes TMEM -> registers
// -> one swizzled 128-byte row in shared memory (raw = hi, residual = lo) and is multiplied again.  The last warps combine
// the P and Q rows (lanes n and n + 16) into X[kx][l] / X[-kx][l] and write X and the mode-major copy X2.
//
// One persistent CTA per SM, 640 threads: warp 0 TMA (2-stage ring, L2 evict-last), warps 1-2 stage-1 MMA issue (alternate
// pairs), warp 3 stage-2 MMA issue, warps 4-7 lo -> TMEM, warps 8-11 D1 -> stage-2 operand, warps 12/13/16/17 output.
//
// Result (B200, B = 16, in-kernel clock trace tools/trace_k1_tc.py): every MN-major tf32 MMA (M 128, N 32, K 8) occupies
// the tensor pipe for ~95 cycles whatever N is, so the 60 MMAs of an image pair take ~5.7 k cycles against the ~2.1 k
// cycles in which HBM delivers the pair: the kernel is tensor-pipe bound at 35 us.  N = 32 (22 useful columns) is inherent
// to a pruned DFT that keeps 11 row frequencies; with single-pass TF32 (one MMA instead of three) it would be ~14 us, but
// that is outside the fp32 parity bar.  The FFMA kernel (27 us, issue-bound) therefore stays the default.
#include "pdes_common.cuh"
#include "pdes_ptx.cuh"
#ifndef PDES_CPU_EMU
#include <cuda.h>
#include <cstdlib>
#include <cstring>
#endif

namespace pdes {

#ifndef PDES_CPU_EMU
namespace {

constexpr int kD1W = 64, kD1N = 32, kD1Stages = 2, kD1Threads = 640, kD1MaxH = 96, kD1MaxNK = 16, kD1MaxM2 = 16;
constexpr int kD1A2Bytes = 2 * kD1W * 128;          // stage-2 A operand: 2 images x 64 rows (w) x 32 floats

struct D1Params {
  float* X;                 // [B][C][2 m1][m2] complex
  float2* X2;               // [2 m1 m2][B][CinP] complex or null
  const float* lscale;      // [m2] Hermitian weights (backward of irfft2) or null
  const float2* twh;        // [H] (cos, sin)(2 pi j / H)
  const float2* tw2;        // [m2][W + 1] (cos, sin)(2 pi l w / W)
  int B, C0, C1, CinP, H, m1, m2, nimg;
};

#ifdef PDES_K1_TRACE
__device__ long long g_k1_trace[5 * 16 * 4];
#define TR(role, it, k) do { if (blockIdx.x == 0 && lane == 0 && (it) < 16) g_k1_trace[((role) * 16 + (it)) * 4 + (k)] = clock64(); } while (0)
#else
#define TR(role, it, k) do { } while (0)
#endif

struct D1Bars {
  unsigned long long full[kD1Stages], empty[kD1Stages], lo_full[2], d1_full[2], d1_empty[2], a2_full[2], d2_full[2], d2_empty[2];
};

__global__ void __launch_bounds__(kD1Threads, 1)
k_dft_fwd_tc(D1Params p, const __grid_constant__ CUtensorMap tmap_x0, const __grid_constant__ CUtensorMap tmap_x1) {
  PDES_DYN_SMEM(unsigned char, smem_raw);
  unsigned char* base = ptx::align_smem_1024(smem_raw);
  const int H = p.H, NK = p.m1 + 1, M1 = p.m1, M2 = p.m2, C = p.C0 + p.C1;
  const uint32_t blk_bytes = (uint32_t)H * 128u;                 // one column half of one image: H rows of 32 floats
  const uint32_t stage_bytes = 4 * blk_bytes;                    // 2 images x 2 halves
  unsigned char* stages = base;
  unsigned char* a2 = stages + kD1Stages * stage_bytes;          // [2 bufs][hi | lo][2 images][64 w][128 B], 1 KB aligned
  float* Ehi = reinterpret_cast<float*>(a2 + 2 * 2 * kD1A2Bytes);               // canonical K-major [H][32]
  float* Elo = Ehi + (size_t)H * kD1N;
  float* Fhi = Elo + (size_t)H * kD1N;                           // stage-2 table, canonical K-major [64 w][32]
  float* Flo = Fhi + (size_t)kD1W * kD1N;
  __shared__ __align__(8) D1Bars bars;
  __shared__ uint32_t tmem_slot;

  const int tid = threadIdx.x, warp = ptx::uniform_warp_idx(), lane = tid & 31;
  const int npairs = (p.nimg + 1) >> 1;

  if (tid == 0) {
    for (int i = 0; i < kD1Stages; ++i) { ptx::mbar_init(&bars.full[i], 1); ptx::mbar_init(&bars.empty[i], 1); }
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&bars.lo_full[i], 4);
      ptx::mbar_init(&bars.d1_full[i], 1);
      ptx::mbar_init(&bars.d1_empty[i], 4);
      ptx::mbar_init(&bars.a2_full[i], 4);
      ptx::mbar_init(&bars.d2_full[i], 1);
      ptx::mbar_init(&bars.d2_empty[i], 4);
    }
    ptx::fence_mbar_init();
  }
  if (warp == 3) {
    ptx::tmem_alloc(&tmem_slot, 512);
    ptx::tmem_relinquish();
  }
  // E[h][n]: n < 16 -> cos(2 pi n h / H), n >= 16 -> sin(2 pi (n - 16) h / H), zero for kx > m1;
  // F[w][n]: n = 2 l -> cos(2 pi l w / W), n = 2 l + 1 -> sin(2 pi l w / W), zero for l >= m2.
  // hi / lo halves in the canonical no-swizzle K-major layout (8 n x 4 k core matrices; k-group stride 512 bytes).
  for (int idx = tid; idx < (H + kD1W) * kD1N; idx += kD1Threads) {
    const int k = idx / kD1N, n = idx - k * kD1N;
    float v = 0.0f;
    float* hi_t = Ehi;
    float* lo_t = Elo;
    int kk = k;
    if (k < H) {
      const int kx = n & 15;
      if (kx < NK) {
        const float2 t = __ldg(p.twh + (kx * k) % H);
        v = (n < 16) ? t.x : t.y;
      }
    } else {
      kk = k - H;
      hi_t = Fhi;
      lo_t = Flo;
      const int l = n >> 1;
      if (l < M2) {
        const float2 t = __ldg(p.tw2 + l * (kD1W + 1) + kk);
        v = (n & 1) ? t.y : t.x;
      }
    }
    const float hi = __uint_as_float(__float_as_uint(v) & 0xffffe000u);
    const uint32_t off = (uint32_t)(kk >> 2) * 128u + (uint32_t)(n >> 3) * 32u + (uint32_t)(n & 7) * 4u + (uint32_t)(kk & 3);
    hi_t[off] = hi;
    lo_t[off] = v - hi;
  }
  ptx::fence_proxy_async();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  const uint32_t d1_col = 0, lo_col = 64, d2_col = 256;          // D1 at 0 / 32, lo at 64 / 64 + H, D2 at 256 / 288

  if (warp == 0) {
    // ================================================================== TMA producer
    const uint64_t pol = ptx::l2_policy_evict_last();
    uint32_t s = 0, eph = 1, it = 0;
    for (int j = blockIdx.x; j < npairs; j += gridDim.x, ++it) {
      TR(0, it, 0);
      if (it >= (uint32_t)kD1Stages) ptx::mbar_wait(&bars.empty[s], eph);
      TR(0, it, 1);
      const int nim = (2 * j + 1 < p.nimg) ? 2 : 1;
      ptx::mbar_arrive_expect_tx_ws(&bars.full[s], (uint32_t)nim * 2u * blk_bytes);
      for (int i = 0; i < nim; ++i) {
        const int im = 2 * j + i, b = im / C, c = im - b * C;
        const bool first = c < p.C0;
        const int idx = first ? b * p.C0 + c : b * p.C1 + (c - p.C0);
        unsigned char* dst = stages + s * stage_bytes + (uint32_t)(2 * i) * blk_bytes;
        if (first) {
          ptx::tma_load_3d_ws_hint(dst, &tmap_x0, 0, 0, idx, &bars.full[s], pol);
          ptx::tma_load_3d_ws_hint(dst + blk_bytes, &tmap_x0, 32, 0, idx, &bars.full[s], pol);
        } else {
          ptx::tma_load_3d_ws_hint(dst, &tmap_x1, 0, 0, idx, &bars.full[s], pol);
          ptx::tma_load_3d_ws_hint(dst + blk_bytes, &tmap_x1, 32, 0, idx, &bars.full[s], pol);
        }
      }
      if (++s == (uint32_t)kD1Stages) { s = 0; eph ^= 1u; }
    }
  } else if (warp == 1 || warp == 2) {
    // ================================================================== stage-1 MMA issue: two issuers take alternate pairs
    // (one issuing thread sustains one small-N tcgen05.mma per ~50 cycles; the two warps feed the two accumulator buffers)
    const uint32_t me = (uint32_t)(warp - 1);
    const uint32_t idesc_mn = ptx::idesc_tf32_a_mn(128, kD1N), idesc_k = ptx::idesc_tf32(128, kD1N);
    const uint32_t lbo_b = (kD1N / 8) * 128;                      // 512 bytes between k-groups of 4
    const uint64_t ehi0 = ptx::smem_desc_noswizzle(ptx::smem_u32(Ehi), lbo_b, 128);
    const uint64_t elo0 = ptx::smem_desc_noswizzle(ptx::smem_u32(Elo), lbo_b, 128);
    const uint64_t a0 = ptx::smem_desc_mn_sw128b32(ptx::smem_u32(stages), blk_bytes, 512);
    const int nks = H / 8;
    uint32_t s = 0, fph = 0, it = 0;
    for (int j = blockIdx.x; j < npairs; j += gridDim.x, ++it) {
      if ((it & 1u) == me) {
        const uint32_t a = it & 1, ph = (it >> 1) & 1;
        if (me == 0) TR(1, it, 0);
        ptx::mbar_wait(&bars.full[s], fph);
        ptx::mbar_wait(&bars.lo_full[a], ph);
        if (it >= 2) ptx::mbar_wait(&bars.d1_empty[a], ph ^ 1u);
        if (me == 0) TR(1, it, 1);
        ptx::tc_fence_after();
        const uint32_t dcol = tmem_base + d1_col + a * kD1N;
        const uint32_t lcol = tmem_base + lo_col + a * (uint32_t)H;
        const uint64_t as = a0 + (uint64_t)((s * stage_bytes) >> 4);
#pragma unroll 4
        for (int ks = 0; ks < nks; ++ks) {
          const uint64_t ad = as + (uint64_t)((ks * 1024) >> 4);                // next group of 8 rows h
          const uint64_t kb = (uint64_t)((ks * 2 * lbo_b) >> 4);
          ptx::mma_tf32_ws(dcol, ad, elo0 + kb, idesc_mn, ks != 0 ? 1u : 0u);   // trunc(x) * E_lo
          ptx::mma_tf32_ws(dcol, ad, ehi0 + kb, idesc_mn, 1u);                  // trunc(x) * E_hi
        }
#pragma unroll 4
        for (int ks = 0; ks < nks; ++ks)                                        // (x - trunc(x)) * E_hi, A from tensor memory
          ptx::mma_tf32_ta_ws(dcol, lcol + ks * 8, ehi0 + (uint64_t)((ks * 2 * lbo_b) >> 4), idesc_k, 1u);
        ptx::tc_commit_ws(&bars.empty[s]);                        // the stage may be overwritten
        ptx::tc_commit_ws(&bars.d1_full[a]);                      // accumulator ready; the lo buffer may be rewritten
        if (me == 0) TR(1, it, 2);
      }
      if (++s == (uint32_t)kD1Stages) { s = 0; fph ^= 1u; }
    }
  } else if (warp == 3) {
    // ================================================================== stage-2 MMA issue
    const uint32_t idesc_mn = ptx::idesc_tf32_a_mn(128, kD1N);
    const uint32_t lbo_b = (kD1N / 8) * 128;
    const uint64_t fhi0 = ptx::smem_desc_noswizzle(ptx::smem_u32(Fhi), lbo_b, 128);
    const uint64_t flo0 = ptx::smem_desc_noswizzle(ptx::smem_u32(Flo), lbo_b, 128);
    // stage-2 A: the four 32-lane blocks of M = 128 are image 0, image 1 and (never read back) whatever follows them
    const uint64_t a2d0 = ptx::smem_desc_mn_sw128b32(ptx::smem_u32(a2), kD1W * 128, 512);
    uint32_t it = 0;
    for (int j = blockIdx.x; j < npairs; j += gridDim.x, ++it) {
      const uint32_t a = it & 1, ph = (it >> 1) & 1;
      ptx::mbar_wait(&bars.a2_full[a], ph);
      if (it >= 2) ptx::mbar_wait(&bars.d2_empty[a], ph ^ 1u);
      ptx::tc_fence_after();
      const uint32_t dcol = tmem_base + d2_col + a * kD1N;
      const uint64_t hi_d = a2d0 + (uint64_t)((a * 2u * kD1A2Bytes) >> 4), lo_d = hi_d + (uint64_t)(kD1A2Bytes >> 4);
#pragma unroll
      for (int ks = 0; ks < kD1W / 8; ++ks) {
        const uint64_t ko = (uint64_t)((ks * 1024) >> 4), kb = (uint64_t)((ks * 2 * lbo_b) >> 4);
        ptx::mma_tf32_ws(dcol, hi_d + ko, flo0 + kb, idesc_mn, ks != 0 ? 1u : 0u);
        ptx::mma_tf32_ws(dcol, lo_d + ko, fhi0 + kb, idesc_mn, 1u);
        ptx::mma_tf32_ws(dcol, hi_d + ko, fhi0 + kb, idesc_mn, 1u);
      }
      ptx::tc_commit_ws(&bars.d2_full[a]);
      TR(1, it, 3);
    }
  } else if (warp < 8) {
    // ================================================================== lo = x - trunc(x) -> tensor memory
    const int q = warp & 3, m = q * 32 + lane, i = m >> 6, w = m & 63, wb = w >> 5, wl = w & 31;
    const uint32_t col_off = (uint32_t)(2 * i + wb) * blk_bytes + (uint32_t)(wl & 7) * 4u;
    const uint32_t chunk = (uint32_t)(wl >> 3);                  // 32-byte chunk of the 128-byte row (swizzle 128B, 32B atom)
    uint32_t s = 0, fph = 0, it = 0;
    for (int j = blockIdx.x; j < npairs; j += gridDim.x, ++it) {
      const uint32_t a = it & 1, ph = (it >> 1) & 1;
      if (warp == 4) TR(2, it, 0);
      ptx::mbar_wait(&bars.full[s], fph);
      if (warp == 4) TR(2, it, 1);
      if (it >= 2) ptx::mbar_wait(&bars.d1_full[a], ph ^ 1u);     // the MMAs that read this lo buffer have completed
      if (warp == 4) TR(2, it, 2);
      ptx::tc_fence_after();
      const unsigned char* colp = stages + s * stage_bytes + col_off;
      const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16) + lo_col + a * (uint32_t)H;
      for (int h0 = 0; h0 < H; h0 += 16) {
        uint32_t lo[16];
#pragma unroll
        for (int e = 0; e < 16; ++e) {
          const int h = h0 + e;
          float v = 0.0f;
          if (h < H) v = *reinterpret_cast<const float*>(colp + (uint32_t)h * 128u + ((chunk ^ (uint32_t)(h & 3)) << 5));
          lo[e] = __float_as_uint(v - __uint_as_float(__float_as_uint(v) & 0xffffe000u));
        }
        ptx::tmem_st16(trow + (uint32_t)h0, lo);
      }
      ptx::tmem_st_wait();
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&bars.lo_full[a]);
      if (warp == 4) TR(2, it, 3);
      if (++s == (uint32_t)kD1Stages) { s = 0; fph ^= 1u; }
    }
  } else if (warp < 12) {
    // ================================================================== stage-1 accumulator -> stage-2 A operand
    // Lane (image, w) holds row w of D1 = [P[0..15] | Q[0..15]] for its image: that IS row k = w of the MN-major stage-2
    // operand A2[m = (kx, P|Q)][k = w].  It is written raw (the tensor core truncates it: the hi half) and as its residual.
    const int q = warp & 3, m = q * 32 + lane, i = m >> 6, w = m & 63;
    uint32_t it = 0;
    for (int j = blockIdx.x; j < npairs; j += gridDim.x, ++it) {
      const uint32_t a = it & 1, ph = (it >> 1) & 1;
      if (warp == 8) TR(3, it, 0);
      ptx::mbar_wait(&bars.d1_full[a], ph);
      if (warp == 8) TR(3, it, 1);
      if (it >= 2) ptx::mbar_wait(&bars.d2_full[a], ph ^ 1u);     // the stage-2 MMAs that read this A2 buffer have completed
      if (warp == 8) TR(3, it, 2);
      ptx::tc_fence_after();
      uint32_t r[32];
      ptx::tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + d1_col + a * kD1N, r);
      ptx::tmem_ld_wait();
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&bars.d1_empty[a]);
      unsigned char* hi_row = a2 + a * 2u * kD1A2Bytes + (uint32_t)i * (kD1W * 128) + (uint32_t)w * 128u;
      unsigned char* lo_row = hi_row + kD1A2Bytes;
#pragma unroll
      for (int c = 0; c < 4; ++c) {                                // 32-byte chunk c of the row goes to chunk c ^ (w % 4)
        const uint32_t off = ((uint32_t)c ^ (uint32_t)(w & 3)) << 5;
#pragma unroll
        for (int hlf = 0; hlf < 2; ++hlf) {
          float4 hv, lv;
          const int e = c * 8 + hlf * 4;
          hv.x = __uint_as_float(r[e]); hv.y = __uint_as_float(r[e + 1]); hv.z = __uint_as_float(r[e + 2]); hv.w = __uint_as_float(r[e + 3]);
          lv.x = hv.x - __uint_as_float(r[e] & 0xffffe000u);
          lv.y = hv.y - __uint_as_float(r[e + 1] & 0xffffe000u);
          lv.z = hv.z - __uint_as_float(r[e + 2] & 0xffffe000u);
          lv.w = hv.w - __uint_as_float(r[e + 3] & 0xffffe000u);
          *reinterpret_cast<float4*>(hi_row + off + hlf * 16) = hv;
          *reinterpret_cast<float4*>(lo_row + off + hlf * 16) = lv;
        }
      }
      ptx::fence_proxy_async();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&bars.a2_full[a]);
      if (warp == 8) TR(3, it, 3);
    }
  } else if (warp == 12 || warp == 13 || warp == 16 || warp == 17) {
    // ================================================================== output: D2 lanes 0..31 = image 0, 32..63 = image 1
    // lane n < 16: (a, bs) = sum_w P[kx = n][w] (cos, sin)(l w); lane n >= 16: (cq, d) = the same for Q[kx = n - 16].
    const int i = warp & 1;                                      // warps 12 / 16 -> TMEM lanes 0..31, warps 13 / 17 -> lanes 32..63
    const int lpar = (warp >> 4) & 1;                            // warps 12 / 13 write the even l, 16 / 17 the odd l
    const int kxi = lane & 15;
    const bool neg = lane >= 16;                                 // lanes 16..31 write the -kx rows
    uint32_t it = 0;
    for (int j = blockIdx.x; j < npairs; j += gridDim.x, ++it) {
      const uint32_t a = it & 1, ph = (it >> 1) & 1;
      if (warp == 12) TR(4, it, 0);
      ptx::mbar_wait(&bars.d2_full[a], ph);
      if (warp == 12) TR(4, it, 1);
      ptx::tc_fence_after();
      uint32_t r[32];
      ptx::tmem_ld32(tmem_base + ((uint32_t)(i * 32) << 16) + d2_col + a * kD1N, r);
      ptx::tmem_ld_wait();
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&bars.d2_empty[a]);
      const int im = 2 * j + i;
      const bool live = im < p.nimg && kxi < NK && (neg ? kxi > 0 : kxi < M1);
      const int b = im / C, c = im - b * C;
      const int krow = neg ? 2 * M1 - kxi : kxi;
      float* o = p.X + ((size_t)im * (2 * M1) + krow) * M2 * 2;
#pragma unroll
      for (int l = 0; l < kD1MaxM2; ++l) {
        const float mine_c = __uint_as_float(r[2 * l]), mine_s = __uint_as_float(r[2 * l + 1]);
        const float oth_c = __shfl_xor_sync(0xffffffffu, mine_c, 16), oth_s = __shfl_xor_sync(0xffffffffu, mine_s, 16);
        if (l < M2 && live && (l & 1) == lpar) {
          const float sa = neg ? oth_c : mine_c, sb = neg ? oth_s : mine_s;      // sums over P
          const float sc = neg ? mine_c : oth_c, sd = neg ? mine_s : oth_s;      // sums over Q
          const float scl = (p.lscale != nullptr) ? __ldg(p.lscale + l) : 1.0f;
          // +kx -> row k = kx: sum (P - iQ)(c - is);   -kx -> row k = 2*M1 - kx: sum (P + iQ)(c - is)
          const float vr = neg ? scl * (sa + sd) : scl * (sa - sd);
          const float vi = neg ? scl * (sc - sb) : -scl * (sb + sc);
          *reinterpret_cast<float2*>(o + 2 * l) = make_float2(vr, vi);
          if (p.X2 != nullptr) p.X2[((size_t)(krow * M2 + l) * p.B + b) * p.CinP + c] = make_float2(vr, vi);
        }
      }
      if (warp == 12) TR(4, it, 2);
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 3) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, 512);
  }
}

typedef CUresult (*D1EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int g_d1_sms = 0;

bool d1_encode(D1EncodeFn enc, CUtensorMap* tm, const float* x, int H, long nimg) {
  memset(tm, 0, sizeof(*tm));
  const cuuint64_t gdim[3] = {(cuuint64_t)kD1W, (cuuint64_t)H, (cuuint64_t)nimg};
  const cuuint64_t gstr[2] = {(cuuint64_t)kD1W * 4, (cuuint64_t)H * kD1W * 4};
  const cuuint32_t box[3] = {32, (cuuint32_t)H, 1};
  const cuuint32_t estr[3] = {1, 1, 1};
  return enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(x), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace
#endif  // !PDES_CPU_EMU

/* Returns -1 when the shape / build is outside this kernel (the caller then uses the FFMA kernels), else a PDES_* code. */
int dft_fwd_tc_try(const float* x0, int C0, const float* x1, int C1, int B, int H, int W, int m1, int m2, const float* tables,
                   int herm_scale, float* X, float* X2, int CinP, void* stream) {
#ifdef PDES_CPU_EMU
  (void)x0; (void)C0; (void)x1; (void)C1; (void)B; (void)H; (void)W; (void)m1; (void)m2; (void)tables; (void)herm_scale;
  (void)X; (void)X2; (void)CinP; (void)stream;
  return -1;
#else
  if (W != kD1W || H % 8 != 0 || H > kD1MaxH || H < 8 || m1 + 1 > kD1MaxNK || 2 * m1 > H || m2 > kD1MaxM2) return -1;
  if (pdes_get_tensor_core_mode() < 2 || tensor_map_encoder() == nullptr) return -1;
  if (!aligned16(x0) || (x1 != nullptr && !aligned16(x1)) || !aligned16(tables)) return -1;
  if ((long)B * (C0 + C1) < 2) return -1;
  {
    const char* e = getenv("PDES_K1_TC");                          // opt-in: slower than the FFMA kernel (see the header)
    if (e == nullptr || e[0] != '1') return -1;
  }
  D1EncodeFn enc = reinterpret_cast<D1EncodeFn>(tensor_map_encoder());
  alignas(64) CUtensorMap t0, t1;
  if (!d1_encode(enc, &t0, x0, H, (long)B * C0)) return -1;
  if (C1 > 0) { if (!d1_encode(enc, &t1, x1, H, (long)B * C1)) return -1; } else t1 = t0;
  const TableLayout t = table_layout(H, W, m1, m2);
  D1Params p;
  p.X = X; p.X2 = reinterpret_cast<float2*>(X2); p.lscale = herm_scale ? tables + t.herm : nullptr;
  p.twh = reinterpret_cast<const float2*>(tables + t.twh);
  p.tw2 = reinterpret_cast<const float2*>(tables + t.tw2);
  p.B = B; p.C0 = C0; p.C1 = C1; p.CinP = CinP; p.H = H; p.m1 = m1; p.m2 = m2; p.nimg = B * (C0 + C1);
  const size_t smem = (size_t)kD1Stages * 4 * H * 128 + (size_t)2 * 2 * kD1A2Bytes + (size_t)2 * (H + kD1W) * kD1N * 4 + 1024;
  if (smem > 227 * 1024) return -1;
  if (g_d1_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_d1_sms, cudaDevAttrMultiProcessorCount, dev);
    if (g_d1_sms <= 0) g_d1_sms = 148;
  }
  const int npairs = (p.nimg + 1) / 2;
  auto kfn = k_dft_fwd_tc;
  PDES_SET_SMEM(kfn, smem);
  PDES_LAUNCH(kfn, dim3((unsigned)(npairs < g_d1_sms ? npairs : g_d1_sms)), dim3(kD1Threads), smem, stream, p, t0, t1);
  return check_launch("pdes_dft_fwd(tcgen05)");
#endif
}

}  // namespace pdes

#if defined(PDES_K1_TRACE) && !defined(PDES_CPU_EMU)
extern "C" int pdes_debug_k1_trace(long long* host_out) {
  return (int)cudaMemcpyFromSymbol(host_out, pdes::g_k1_trace, sizeof(long long) * 5 * 16 * 4);
}
#endif
