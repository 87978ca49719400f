// K3b on the 5th-generation tensor cores (tcgen05 + TMEM), fp32-faithful via the 3xTF32 split.
//
//   pre[b][n][p] = sum_k xin[b][k][p] * Wt[k][n]  (+ spectral term + bias + residual),  out = act(pre)
//
// GEMM orientation: M = 128 pixels of one image (TMEM lanes), N = output channels (TMEM columns, <= 256),
// K = input channels.  With pixels on the TMEM lanes every epilogue thread owns one pixel, so for a fixed channel a
// warp stores 32 consecutive pixels: fully coalesced NCHW stores with no shared-memory transpose.
//
// fp32 parity: tcgen05 has no fp32 kind.  Every operand x is split as hi = tf32(x) (low 13 mantissa bits cleared) and
// lo = x - hi (exact); D += A_hi*B_hi + A_lo*B_hi + A_hi*B_lo with fp32 accumulation in TMEM leaves a relative error
// of ~2^-21 per product, well inside the 1e-5 budget (plain TF32 would be ~1e-3).  Tensor work triples, but
// K3b is HBM-bound on the tensor pipe even at 3x (about 4 us of MMA per 128-pixel tile vs 6 us of HBM time).
//
// Operand staging (no swizzle, K-major canonical "core matrix" layout, 8 rows x 16 bytes):
//   * A (activations): 128 producer threads, one pixel each, read 16 channels with warp-coalesced loads, split
//     hi/lo in registers and write 16-byte rows of the core matrices (conflict-free) -- the split needs the data in
//     registers anyway, so this is where the staging belongs;
//   * B (weights): packed once per weight version by `k_pack_b_tf32` into per-chunk canonical blocks (hi, lo), so
//     a chunk is ONE contiguous block fetched by a bulk async copy (TMA 1-D, SASS UBLKCP) that completes on the
//     stage's mbarrier together with the producers' arrivals.
// Pipeline: 2 stages of 16 channels; warp 4 lane 0 issues the MMAs and frees stages with tcgen05.commit; two CTAs
// per SM (2 x 256 TMEM columns, 2 x ~82 KB shared memory) so one CTA's epilogue overlaps the other's main loop.
#include "pdes_common.cuh"
#include "pdes_ptx.cuh"
#ifndef PDES_CPU_EMU
#include <cuda.h>      // CUtensorMap types only; the encoder is fetched with cudaGetDriverEntryPoint (no libcuda link)
#include <cstring>
#include <cstdlib>
#endif

namespace pdes {

constexpr int kTcBK = 16;          // channels per pipeline stage
constexpr int kTcM = 128;          // pixels per CTA tile
constexpr int kTcMaxN = 256;
constexpr int kWgtMaxCtas = 160;   // upper bound on the CTAs (= partial sums) of the tensor-core weight-gradient kernel

__host__ __device__ inline int tc_npad(int N) { return (N + 15) & ~15; }
__host__ __device__ inline int tc_nchunks(int K) { return (K + kTcBK - 1) / kTcBK; }
// floats of one canonical block (N_pad x 16); a chunk is two blocks: hi then lo
__host__ __device__ inline size_t tc_block_floats(int N) { return (size_t)tc_npad(N) * kTcBK; }

namespace {

__device__ __forceinline__ float tf32_hi(float x) { return __uint_as_float(__float_as_uint(x) & 0xffffe000u); }

// W(k, n) = Wt[k * sk + n * sn] -> packed[chunk][hi|lo][canonical K-major block of N_pad rows x 16 k]
__global__ void __launch_bounds__(256)
k_pack_b_tf32(const float* __restrict__ Wt, size_t sk, size_t sn, int K, int N, int npad, int nchunks, float* __restrict__ out) {
  const int total = nchunks * npad * kTcBK;
  const int lbo_f = (npad / 8) * 32;                        // floats between K-adjacent core matrices
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const int n = idx % npad;                               // n fastest: coalesced reads of Wt rows
    const int kk = (idx / npad) % kTcBK;
    const int c = idx / (npad * kTcBK);
    const int k = c * kTcBK + kk;
    const float w = (k < K && n < N) ? __ldg(Wt + (size_t)k * sk + (size_t)n * sn) : 0.0f;
    const float hi = tf32_hi(w);
    const size_t blk = (size_t)npad * kTcBK;
    const size_t off = (size_t)(kk / 4) * lbo_f + (size_t)(n / 8) * 32 + (n % 8) * 4 + (kk % 4);
    out[(size_t)c * 2 * blk + off] = hi;
    out[(size_t)c * 2 * blk + blk + off] = w - hi;
  }
}

#ifndef PDES_CPU_EMU

struct TcParams {
  const float* Z;       // [B][H][2*m2][N] or null
  const float* wpack;   // packed weights
  const float* x0; int C0;
  const float* x1; int C1;
  const float* bias;
  const float* res;
  const float* T;       // [2*m2][W]
  float* out;
  float* pre;
  int N, npad, K, H, W, m2, act;
  int single_pass;      // 1 = plain TF32 (hi*hi only): the separately reported reduced-precision mode
  size_t out_bs;        // batch stride (floats) of out / res / pre: N*H*W, or larger when writing a channel sub-range
};

constexpr int kTcRaw = 3;          // depth of the raw activation ring fed by bulk async copies
constexpr int kTcThreads = 192;    // warps 0-3: convert + epilogue, warp 4: MMA issue + TMEM, warp 5: bulk-copy issue

// Pipeline of one CTA (= 128 pixels x all output channels of one image):
//   warp 5 lane 0 : raw activations, 16 channel rows of 512 B per chunk, bulk async copies into a 3-deep ring
//   warps 0-3     : ring -> registers -> hi/lo TF32 split -> canonical K-major A tile (2 stages); thread 0 also
//                   launches the bulk copy of the pre-packed weight chunk into the same stage
//   warp 4 lane 0 : 3 x tcgen05.mma per K=8 step into TMEM, tcgen05.commit frees the stage
//   spectral term : appended as extra K chunks: A = T[j][w] masked by the pixel's row, B = Z[b][row][j][:]
//   warps 0-3     : epilogue, thread = pixel = TMEM lane; residual prefetched one 32-column group ahead
__global__ void __launch_bounds__(kTcThreads, 2)
k_inv_w_gemm_tc(TcParams p) {
  PDES_DYN_SMEM(unsigned char, smem_raw);
  unsigned char* base = ptx::align_smem_1024(smem_raw);
  const int npad = p.npad;
  const uint32_t a_blk = kTcM * kTcBK * 4;                 // 8 KB
  const uint32_t b_blk = (uint32_t)npad * kTcBK * 4;
  const uint32_t stage_bytes = 2 * a_blk + 2 * b_blk;
  float* raw = reinterpret_cast<float*>(base + 2 * (size_t)stage_bytes);      // [kTcRaw][16][128]
  __shared__ __align__(8) unsigned long long full_bar[2], empty_bar[2], raw_full[kTcRaw], raw_empty[kTcRaw], done_bar;
  __shared__ uint32_t tmem_slot;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int HW = p.H * p.W, W = p.W;
  const int p0 = blockIdx.x * kTcM, b = blockIdx.y;
  const int nx = tc_nchunks(p.K);
  const int J = 2 * p.m2;
  const int plast = (p0 + kTcM - 1 < HW - 1) ? (p0 + kTcM - 1) : (HW - 1);
  const int h0 = p0 / W;
  const int kspec = (p.Z != nullptr) ? (plast / W - h0 + 1) * J : 0;
  const int nsp = tc_nchunks(kspec);
  const int nchunks = nx + nsp;
  const int npx = (HW - p0 < kTcM) ? (HW - p0) : kTcM;
  const uint32_t tmem_cols = npad <= 32 ? 32 : (npad <= 64 ? 64 : (npad <= 128 ? 128 : 256));

  if (tid == 0) {
    for (int s = 0; s < 2; ++s) {
      ptx::mbar_init(&full_bar[s], 129);
      ptx::mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < kTcRaw; ++s) {
      ptx::mbar_init(&raw_full[s], 1);
      ptx::mbar_init(&raw_empty[s], 128);
    }
    ptx::mbar_init(&done_bar, 1);
    ptx::fence_mbar_init();
  }
  if (warp == 4) {
    ptx::tmem_alloc(&tmem_slot, tmem_cols);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = tmem_slot;

  if (warp < 4) {
    // ------------------------------------------------------------------ convert: one pixel per thread
    const int pp = p0 + tid;
    const bool pvalid = pp < HW;
    const int hh = pvalid ? pp / W : 0, ww = pvalid ? pp % W : 0;
    const uint32_t row_off = (uint32_t)(tid >> 3) * 128 + (uint32_t)(tid & 7) * 16;   // (m/8)*SBO + (m%8)*16
    const uint32_t lbo_b = (uint32_t)(npad / 8) * 128;
    for (int c = 0; c < nchunks; ++c) {
      const int s = c & 1, u = c >> 1;
      if (c >= 2) ptx::mbar_wait(&empty_bar[s], (uint32_t)((u - 1) & 1));
      unsigned char* st = base + (size_t)s * stage_bytes;
      float v[kTcBK];
      int release_slot = -1;
      if (c >= nsp) {
        const int cx = c - nsp;                               // activation chunk index
        if (tid == 0) {
          ptx::mbar_arrive_expect_tx(&full_bar[s], 2 * b_blk);
          ptx::bulk_g2s(st + 2 * a_blk, p.wpack + (size_t)cx * 2 * (b_blk / 4), 2 * b_blk, &full_bar[s]);
        }
        const int slot = cx % kTcRaw;
        ptx::mbar_wait(&raw_full[slot], (uint32_t)((cx / kTcRaw) & 1));
        const float* rw = raw + (size_t)slot * kTcBK * kTcM + tid;
#pragma unroll
        for (int kk = 0; kk < kTcBK; ++kk) v[kk] = (pvalid && cx * kTcBK + kk < p.K) ? rw[kk * kTcM] : 0.0f;
        release_slot = slot;                                  // released below, once the loaded values have been consumed
      } else {
        const int cs = c;                                     // spectral chunks come first: their L2 latency
                                                              // overlaps the first bulk copies of the activations
        if (tid == 0) ptx::mbar_arrive(&full_bar[s]);          // stands in for the weight copy's arrival
        // A: T[j][w] if this pixel sits in row h0 + r, else 0   (block-diagonal over the rows of the tile)
#pragma unroll
        for (int kk = 0; kk < kTcBK; ++kk) {
          const int k = cs * kTcBK + kk;
          const int r = k / J, j = k - r * J;
          v[kk] = (pvalid && k < kspec && hh - h0 == r) ? __ldg(p.T + (size_t)j * W + ww) : 0.0f;
        }
        // B: Z[b][h0 + r][j][n] -> canonical (n, kk); 8 independent loads in flight per thread
        constexpr int UN = 8;
        const int nelem = npad * kTcBK;
        for (int base0 = tid; base0 < nelem; base0 += 128 * UN) {
          float z[UN];
#pragma unroll
          for (int uu = 0; uu < UN; ++uu) {
            const int idx = base0 + uu * 128;
            const int kk = idx / npad, n = idx - kk * npad;
            const int k = cs * kTcBK + kk;
            const int r = k / J, j = k - r * J;
            z[uu] = (idx < nelem && k < kspec && n < p.N)
                        ? __ldg(p.Z + (((size_t)b * p.H + h0 + r) * J + j) * p.N + n) : 0.0f;
          }
#pragma unroll
          for (int uu = 0; uu < UN; ++uu) {
            const int idx = base0 + uu * 128;
            if (idx < nelem) {
              const int kk = idx / npad, n = idx - kk * npad;
              const float zh = tf32_hi(z[uu]);
              const uint32_t off = (uint32_t)(kk >> 2) * lbo_b + (uint32_t)(n >> 3) * 128 + (uint32_t)(n & 7) * 16 + (uint32_t)(kk & 3) * 4;
              *reinterpret_cast<float*>(st + 2 * a_blk + off) = zh;
              *reinterpret_cast<float*>(st + 2 * a_blk + b_blk + off) = z[uu] - zh;
            }
          }
        }
      }
#pragma unroll
      for (int q = 0; q < kTcBK / 4; ++q) {
        float4 hi, lo;
        hi.x = tf32_hi(v[4 * q + 0]); lo.x = v[4 * q + 0] - hi.x;
        hi.y = tf32_hi(v[4 * q + 1]); lo.y = v[4 * q + 1] - hi.y;
        hi.z = tf32_hi(v[4 * q + 2]); lo.z = v[4 * q + 2] - hi.z;
        hi.w = tf32_hi(v[4 * q + 3]); lo.w = v[4 * q + 3] - hi.w;
        const uint32_t off = (uint32_t)q * (kTcM / 8) * 128 + row_off;                // (k/4)*LBO_A + row
        *reinterpret_cast<float4*>(st + off) = hi;
        *reinterpret_cast<float4*>(st + a_blk + off) = lo;
      }
      // The raw slot is handed back only AFTER its values went through the stores above: an arrive placed right behind
      // the loads is issued while they are still in flight (ptxas puts no scoreboard wait in front of SYNCS.ARRIVE), and
      // the producer's next copy into the slot then overtakes them.
      if (release_slot >= 0) ptx::mbar_arrive(&raw_empty[release_slot]);
      ptx::fence_proxy_async();
      ptx::mbar_arrive(&full_bar[s]);
    }

    // ------------------------------------------------------------------ epilogue: thread = pixel (TMEM lane)
    const int N = p.N;
    const size_t obase = (size_t)b * N * HW + pp;
    auto load_res = [&](float (&rr)[32], int n0) {
#pragma unroll
      for (int e = 0; e < 32; ++e) {
        const int n = n0 + e;
        rr[e] = (p.res != nullptr && pvalid && n < N) ? __ldg(p.res + obase + (size_t)n * HW) : 0.0f;
      }
    };
    auto finish = [&](const float (&rr)[32], int n0) {
      uint32_t r[32];
      ptx::tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)n0, r);
      ptx::tmem_ld_wait();
      if (!pvalid) return;
#pragma unroll
      for (int e = 0; e < 32; ++e) {
        const int n = n0 + e;
        if (n < N) {
          float v = __uint_as_float(r[e]) + rr[e];
          if (p.bias != nullptr) v += __ldg(p.bias + n);
          if (p.pre != nullptr) p.pre[obase + (size_t)n * HW] = v;
          if (p.act == PDES_ACT_GELU) v = gelu_f(v);
          p.out[obase + (size_t)n * HW] = v;
        }
      }
    };
    float ra[32], rb[32];
    load_res(ra, 0);                                   // issued before the accumulator is complete
    ptx::mbar_wait(&done_bar, 0);
    ptx::tc_fence_after();
    for (int n0 = 0; n0 < npad; n0 += 64) {
      const bool second = n0 + 32 < npad;
      if (second) load_res(rb, n0 + 32);
      finish(ra, n0);
      if (second) {
        if (n0 + 64 < npad) load_res(ra, n0 + 64);
        finish(rb, n0 + 32);
      }
    }
    ptx::tc_fence_before();
  } else if (warp == 4) {
    if (lane == 0) {
      // ---------------------------------------------------------------- MMA issuer (one thread)
      const uint32_t idesc = ptx::idesc_tf32(kTcM, npad);
      const uint32_t lbo_a = (kTcM / 8) * 128, lbo_b = (uint32_t)(npad / 8) * 128, sbo = 128;
      for (int c = 0; c < nchunks; ++c) {
        const int s = c & 1, u = c >> 1;
        ptx::mbar_wait(&full_bar[s], (uint32_t)(u & 1));
        ptx::tc_fence_after();
        const uint32_t sa = ptx::smem_u32(base + (size_t)s * stage_bytes);
        const uint32_t sb = sa + 2 * a_blk;
#pragma unroll
        for (int ks = 0; ks < kTcBK / 8; ++ks) {
          const uint64_t a_hi = ptx::smem_desc_noswizzle(sa + ks * 2 * lbo_a, lbo_a, sbo);
          const uint64_t a_lo = ptx::smem_desc_noswizzle(sa + a_blk + ks * 2 * lbo_a, lbo_a, sbo);
          const uint64_t b_hi = ptx::smem_desc_noswizzle(sb + ks * 2 * lbo_b, lbo_b, sbo);
          const uint64_t b_lo = ptx::smem_desc_noswizzle(sb + b_blk + ks * 2 * lbo_b, lbo_b, sbo);
          ptx::mma_tf32(tmem_base, a_lo, b_hi, idesc, (c | ks) != 0 ? 1u : 0u);   // small terms first
          ptx::mma_tf32(tmem_base, a_hi, b_lo, idesc, 1u);
          ptx::mma_tf32(tmem_base, a_hi, b_hi, idesc, 1u);
        }
        ptx::tc_commit(&empty_bar[s]);
      }
      ptx::tc_commit(&done_bar);
    }
  } else if (lane == 0) {
    // ------------------------------------------------------------------ bulk-copy issuer (raw activation rows)
    const uint32_t rowbytes = (uint32_t)npx * 4;
    for (int c = 0; c < nx; ++c) {
      const int slot = c % kTcRaw, use = c / kTcRaw;
      if (c >= kTcRaw) ptx::mbar_wait(&raw_empty[slot], (uint32_t)((use - 1) & 1));
      const int nk = (p.K - c * kTcBK < kTcBK) ? (p.K - c * kTcBK) : kTcBK;
      ptx::mbar_arrive_expect_tx(&raw_full[slot], (uint32_t)nk * rowbytes);
      float* dst = raw + (size_t)slot * kTcBK * kTcM;
      for (int kk = 0; kk < nk; ++kk) {
        const int k = c * kTcBK + kk;
        const float* src = (k < p.C0) ? p.x0 + ((size_t)b * p.C0 + k) * HW : p.x1 + ((size_t)b * p.C1 + (k - p.C0)) * HW;
        ptx::bulk_g2s(dst + kk * kTcM, src + p0, rowbytes, &raw_full[slot]);
      }
    }
  }
  __syncthreads();
  if (warp == 4) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, tmem_cols);
  }
}


// ----------------------------------------------------------------------------------------------------------------
// v3: persistent, warp-specialised version of the same GEMM (mode 2, default).
//   one CTA per SM loops over 128-pixel tiles; 896 threads (see kK3* below):
//     warps 0-7   convert   two groups taking alternate 16-channel chunks: raw ring -> hi/lo split -> the A operand in
//                           TENSOR MEMORY (tcgen05.st) and, for the spectral chunks, the canonical B stage
//     warps 8-23  epilogue  TMEM -> registers -> bias / residual / packed-FFMA2 GELU -> coalesced stores, overlapping
//                           the next tile's main loop through the second accumulator
//     warp  24    MMA       one thread issues tcgen05.mma (A from TMEM, B from shared memory) into one of TWO TMEM
//                           accumulators and frees the stage with one tcgen05.commit
//     warp  25    raw issue 2-D TMA boxes of activation rows / bulk copies of Z rows into a raw ring (up to 8 slots)
//     warp  26    B issue   bulk async copies of the packed weight chunks into the stage ring, running ahead across
//                           tile boundaries
#ifdef PDES_TC_TRACE
__device__ long long g_trace[4096];
#if PDES_TC_TRACE == 2       // MMA-issue thread only: 10 stamps per chunk (start, a_full, b_full, 6 MMAs, commits)
#define TRACE(slot) do { } while (0)
#define TRACE2(slot) do { if (blockIdx.x == 0 && (slot) < 4096) g_trace[(slot)] = clock64(); } while (0)
#else
#define TRACE(slot) do { if (blockIdx.x == 0 && (slot) < 4096) g_trace[(slot)] = clock64(); } while (0)
#define TRACE2(slot) do { } while (0)
#endif
#else
#define TRACE(slot) do { } while (0)
#define TRACE2(slot) do { } while (0)
#endif
constexpr int kV3Threads = 768;       // 8 service warps + 16 epilogue warps
constexpr int kV3EpiParts = 4;         // epilogue warps per TMEM lane quadrant
constexpr int kV3ASt = 3, kV3BSt = 4, kV3Raw = 3;   // (ring depths of the 3x3-conv kernel below)
constexpr int kV3MaxSt = 4, kV3MaxRaw = 8;

struct V3Bars {
  unsigned long long a_full[kV3ASt], a_empty[kV3ASt], b_full[kV3BSt], b_empty[kV3BSt], raw_full[kV3Raw],
      raw_empty[kV3Raw], acc_full[2], acc_empty[2];
};

// One ring of NST operand stages (A and B share the index): full[s] = 128 convert arrivals + 1 arrival (carrying the byte
// count of the weight copy) of the B-issue thread, empty[s] = ONE tcgen05.commit.  The MMA thread therefore pays one
// wait and one commit per chunk: its serial overhead between chunks is what the tensor pipe idles on.
struct V3RingBars {
  unsigned long long full[kV3MaxSt], empty[kV3MaxSt], raw_full[kV3MaxRaw], raw_empty[kV3MaxRaw], acc_full[2], acc_empty[2];
};

// Warp roles of the K3b kernel (896 threads): warps 0-7 convert (two groups of 4 that take alternate chunks), 8-23
// epilogue (4 per TMEM lane quadrant), 24 MMA issue, 25 activation copies, 26 weight copies.
constexpr int kK3Threads = 896, kK3EpiParts = 4, kK3MmaWarp = 24, kK3RawWarp = 25, kK3WgtWarp = 26;

// kTA: the activation operand is written by the convert warps straight into TENSOR MEMORY (tcgen05.st, lane = pixel,
// one column per k) and the MMAs read it from there, so it never crosses shared memory a second and third time:
// per 16-channel chunk the shared-memory traffic drops from 116 KB to 76 KB (the mainloop is shared-memory-bandwidth
// bound).  Needs 2 * npad + 2 * 16 * NST <= 512 TMEM columns: NST = 4 stages for npad <= 192, 3 for npad <= 208; wider
// outputs keep (3) A stages in shared memory.  The hi halves sit in the gap below column 256 (after accumulator 0), the
// lo halves below column 512 (after accumulator 1).
constexpr int kV3TaMaxN = 208;
template <bool kTA>
__global__ void __launch_bounds__(kK3Threads, 1)
k_inv_w_gemm_tc_v3(TcParams p, int B, int ntiles_per_img, const __grid_constant__ CUtensorMap tmap_x0, int ntmap_chunks,
                   int NST, int NRAW, int nfull, int tail_shift) {
  PDES_DYN_SMEM(unsigned char, smem_raw);
  unsigned char* base = ptx::align_smem_1024(smem_raw);
  const int npad = p.npad;
  const uint32_t a_blk = kTcM * kTcBK * 4;                       // 8 KB
  const uint32_t b_blk = (uint32_t)npad * kTcBK * 4;             // 12 KB at N = 192
  const uint32_t a_stage = 2 * a_blk, b_stage = 2 * b_blk;
  const int slot_f = kTcBK * (npad > kTcM ? npad : kTcM);         // floats per raw slot
  unsigned char* sA = base;
  unsigned char* sB = sA + (kTA ? 0 : NST * a_stage);
  float* raw = reinterpret_cast<float*>(sB + NST * b_stage);
  const uint32_t ta_hi = 256u - 16u * NST, ta_lo = 512u - 16u * NST;   // TMEM columns of the A stages (kTA)
  float* tsp = raw + (size_t)NRAW * slot_f;                     // [nsp_max*16][128]: fp32 A operand of the spectral chunks
  __shared__ __align__(8) V3RingBars bars;
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(16) float bias_s[kTcMaxN];      // the epilogue stalled on per-group __ldg(bias) (ncu: 18 % of all samples)

  // warp index as a warp-uniform value: the single-issuer roles below run with the whole warp converged and one elected
  // lane issuing (ptx::*_ws), which keeps their operands in uniform registers (no waterfall loop per tcgen05 / TMA issue)
  const int tid = threadIdx.x, warp = ptx::uniform_warp_idx(), lane = tid & 31;
  const int HW = p.H * p.W, W = p.W, J = 2 * p.m2;
  const int nx = tc_nchunks(p.K);
  const int ntiles = B * ntiles_per_img;
#ifdef PDES_TC_TRACE
  const int dbg = p.act >> 8;        // ablation mask (trace builds only): 1 no MMA, 2 no weight copy, 4 no raw copy, 8 no convert
  p.act &= 255;
#else
  constexpr int dbg = 0;
#endif

  if (tid == 0) {
    // ONE arrival per warp (after __syncwarp), not one per thread: 128 arrivals on one barrier word serialise at a few
    // cycles each, and two such barriers per chunk were ~800 cycles of the ~1250-cycle chunk period.
    for (int i = 0; i < NST; ++i) { ptx::mbar_init(&bars.full[i], 4 + 1); ptx::mbar_init(&bars.empty[i], 1); }
    for (int i = 0; i < NRAW; ++i) { ptx::mbar_init(&bars.raw_full[i], 1); ptx::mbar_init(&bars.raw_empty[i], 4); }
    for (int i = 0; i < 2; ++i) { ptx::mbar_init(&bars.acc_full[i], 1); ptx::mbar_init(&bars.acc_empty[i], 4 * kK3EpiParts); }
    ptx::fence_mbar_init();
  }
  if (warp == kK3MmaWarp) {
    ptx::tmem_alloc(&tmem_slot, 512);
    ptx::tmem_relinquish();
  }
  // Launched as a programmatic dependent of the previous kernel in the stream: everything above overlaps its tail; no
  // global memory is read before this point.
  PDES_GRID_DEP_WAIT();
  for (int i = threadIdx.x; i < kTcMaxN; i += blockDim.x) bias_s[i] = (p.bias != nullptr && i < p.N) ? __ldg(p.bias + i) : 0.0f;
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = tmem_slot;

  // Work items (identical in every role).  Items [0, nfull) are whole 128-pixel x npad-column tiles, nfull a multiple of
  // the grid.  The ntiles - nfull tiles of the last, partly filled round are cut into 2^tail_shift column ranges each so
  // that (almost) every CTA gets a piece: at B = 16 the 768 tiles ran as 6 rounds on 148 CTAs for 5.19 rounds of work.
  const int nitems = nfull + ((ntiles - nfull) << tail_shift);
  auto item_geom = [&](int i, int& t, int& n0, int& ncols) {
    if (i < nfull) { t = i; n0 = 0; ncols = npad; return; }
    const int j = i - nfull;
    t = nfull + (j >> tail_shift);
    ncols = npad >> tail_shift;
    n0 = (j & ((1 << tail_shift) - 1)) * ncols;
  };
  auto tile_geom = [&](int t, int& b, int& p0, int& h0, int& kspec, int& nsp) {
    b = t / ntiles_per_img;
    p0 = (t - b * ntiles_per_img) * kTcM;
    const int plast = (p0 + kTcM - 1 < HW - 1) ? (p0 + kTcM - 1) : (HW - 1);
    h0 = p0 / W;
    kspec = (p.Z != nullptr) ? (plast / W - h0 + 1) * J : 0;
    nsp = tc_nchunks(kspec);
  };

  if (warp < 8) {
    // ================================================================== convert (group = warp / 4 takes chunks g % 2 == group)
    const int grp = warp >> 2, cwarp = warp & 3;
    const int tid = threadIdx.x & 127;                             // pixel / TMEM lane of this thread (shadows the CTA-wide tid)
    auto cvt_sync = [] { asm volatile("bar.sync 1, 256;" ::: "memory"); };
    const uint32_t row_off = (uint32_t)(tid >> 3) * 128 + (uint32_t)(tid & 7) * 16;
    const uint32_t lbo_b = (uint32_t)(npad / 8) * 128;
    uint32_t g = 0, s = 0, sph = 1;                              // global chunk counter, ring slot and its "empty" parity
    uint32_t r = 0, rph = 0;                                     // raw-ring slot and its "full" parity
    // Spectral A operand: A[m][k] = T[j][w(m)] if pixel m sits in row h0 + k/J of the tile (block diagonal over the
    // tile's rows).  It depends on the tile only through p0 % W, so when every tile starts at column 0
    // (128 % W == 0, e.g. the shipped W = 64) it is built ONCE per CTA; otherwise it is rebuilt per tile.
    const bool uniform_geom = (kTcM % W == 0);
    // (counters instead of k / J and loads batched four at a time: as a division + dependent __ldg per k this table
    // took ~11 k cycles = 5.6 us before the first chunk of the kernel could be converted)
    auto build_tsp = [&](int p0, int h0, int kspec, int nsp) {
      const int pp = p0 + tid;
      const bool pv = pp < HW;
      const int hh = pv ? pp / W : 0, ww = pv ? pp % W : 0;
      const int myr = pv ? hh - h0 : -1;                             // the one row block in which this pixel has non-zeros
      const int kend = nsp * kTcBK;
      for (int k = 0; k < kend; ++k) tsp[k * kTcM + tid] = 0.0f;
      if (myr >= 0 && myr * J < kspec) {
        const float* tcol = p.T + ww;
        float* dst = tsp + (size_t)(myr * J) * kTcM + tid;
        int j = 0;
        for (; j + 4 <= J; j += 4) {
          const float v0 = __ldg(tcol + (size_t)j * W), v1 = __ldg(tcol + (size_t)(j + 1) * W);
          const float v2 = __ldg(tcol + (size_t)(j + 2) * W), v3 = __ldg(tcol + (size_t)(j + 3) * W);
          dst[j * kTcM] = v0; dst[(j + 1) * kTcM] = v1; dst[(j + 2) * kTcM] = v2; dst[(j + 3) * kTcM] = v3;
        }
        for (; j < J; ++j) dst[j * kTcM] = __ldg(tcol + (size_t)j * W);
      }
    };
    if (p.Z != nullptr && uniform_geom && (int)blockIdx.x < nitems) {
      int b, p0, h0, kspec, nsp, t0, n0, ncols;
      item_geom(blockIdx.x, t0, n0, ncols);
      tile_geom(t0, b, p0, h0, kspec, nsp);
      if (grp == 0) build_tsp(p0, h0, kspec, nsp);                // column tid is read by thread tid of BOTH groups
      cvt_sync();
    }
    for (int i = blockIdx.x; i < nitems; i += gridDim.x) {
      int b, p0, h0, kspec, nsp, t, n0, ncols;
      item_geom(i, t, n0, ncols);
      tile_geom(t, b, p0, h0, kspec, nsp);
      const int pp = p0 + tid;
      const bool pvalid = pp < HW;
      if (p.Z != nullptr && !uniform_geom) {
        cvt_sync();                                                // both groups are done with the previous tile's table
        if (grp == 0) build_tsp(p0, h0, kspec, nsp);
        cvt_sync();
      }
      for (int c = 0; c < nsp + nx; ++c, ++g) {
       if ((int)(g & 1) == grp) {
        if (tid == 0) TRACE(0 * 512 + g * 4 + 0);
        if (g >= (uint32_t)NST) ptx::mbar_wait(&bars.empty[s], sph);
        if (tid == 0) TRACE(0 * 512 + g * 4 + 1);
        unsigned char* st = sA + s * a_stage;
        float v[kTcBK];
        bool release_raw = false;
        if (kTA) ptx::tc_fence_after();
        if (c >= nsp) {
          const int cx = c - nsp;
          ptx::mbar_wait(&bars.raw_full[r], rph);
          if (tid == 0) TRACE(0 * 512 + g * 4 + 2);
          const float* rw = raw + (size_t)r * slot_f + tid;               // activation rows: dense [16][128]
#pragma unroll
          for (int kk = 0; kk < kTcBK; ++kk) v[kk] = (pvalid && cx * kTcBK + kk < p.K && !(dbg & 8)) ? rw[kk * kTcM] : 0.0f;
          // raw_empty[r] is arrived on BELOW, after v[] went through tcgen05.st / st.shared.  Arriving here released
          // the slot with the 16 loads still in flight (SASS: LD x16, SYNCS.ARRIVE, first use of the data after it); at
          // B = 16 the producer's next TMA box overtook the last rows in ~1 of 8000 tiles (one k-row of a 32-pixel
          // quadrant taken from the chunk 8 ahead): found by tools/stress_k3b.py.
          release_raw = true;
        } else {
          // spectral chunk: A = T[j][w] on this pixel's row, B = Z rows (staged in the raw slot) -> canonical
#pragma unroll
          for (int kk = 0; kk < kTcBK; ++kk) v[kk] = tsp[(c * kTcBK + kk) * kTcM + tid];
          ptx::mbar_wait(&bars.raw_full[r], rph);
          const float* rw = raw + (size_t)r * slot_f;                     // Z rows: dense [16][N]
          unsigned char* sb = sB + s * b_stage;
          // one (n, 4 consecutive k) item = one 16-byte row of a core matrix: conflict-free LDS and STS.128
          for (int n = n0 + tid; n < n0 + ncols; n += 128) {               // only the item's column range is multiplied
#pragma unroll
            for (int kq = 0; kq < kTcBK / 4; ++kq) {
              float z[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const int kk = kq * 4 + e;
                z[e] = (c * kTcBK + kk < kspec && n < p.N) ? rw[kk * p.N + n] : 0.0f;
              }
              float4 hi, lo;
              hi.x = tf32_hi(z[0]); lo.x = z[0] - hi.x;
              hi.y = tf32_hi(z[1]); lo.y = z[1] - hi.y;
              hi.z = tf32_hi(z[2]); lo.z = z[2] - hi.z;
              hi.w = tf32_hi(z[3]); lo.w = z[3] - hi.w;
              const uint32_t off = (uint32_t)kq * lbo_b + (uint32_t)(n >> 3) * 128 + (uint32_t)(n & 7) * 16;
              *reinterpret_cast<float4*>(sb + off) = hi;
              *reinterpret_cast<float4*>(sb + b_blk + off) = lo;
            }
          }
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive(&bars.raw_empty[r]);
        }
        if (dbg & 8) {
          ptx::tc_fence_before();
        } else if constexpr (kTA) {
          uint32_t uh[kTcBK], ul[kTcBK];
#pragma unroll
          for (int kk = 0; kk < kTcBK; ++kk) {
            const float h = tf32_hi(v[kk]);
            uh[kk] = __float_as_uint(h);
            ul[kk] = __float_as_uint(v[kk] - h);
          }
          const uint32_t trow = tmem_base + ((uint32_t)(cwarp * 32) << 16) + s * kTcBK;
          ptx::tmem_st16(trow + ta_hi, uh);
          ptx::tmem_st16(trow + ta_lo, ul);
          ptx::tmem_st_wait();
          if (release_raw) {
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(&bars.raw_empty[r]);
          }
          ptx::tc_fence_before();
          if (c < nsp) ptx::fence_proxy_async();                     // the spectral B block went through st.shared
        } else {
#pragma unroll
          for (int qd = 0; qd < kTcBK / 4; ++qd) {
            float4 hi, lo;
            hi.x = tf32_hi(v[4 * qd + 0]); lo.x = v[4 * qd + 0] - hi.x;
            hi.y = tf32_hi(v[4 * qd + 1]); lo.y = v[4 * qd + 1] - hi.y;
            hi.z = tf32_hi(v[4 * qd + 2]); lo.z = v[4 * qd + 2] - hi.z;
            hi.w = tf32_hi(v[4 * qd + 3]); lo.w = v[4 * qd + 3] - hi.w;
            const uint32_t off = (uint32_t)qd * (kTcM / 8) * 128 + row_off;
            *reinterpret_cast<float4*>(st + off) = hi;
            *reinterpret_cast<float4*>(st + a_blk + off) = lo;
          }
          if (release_raw) {
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(&bars.raw_empty[r]);
          }
          ptx::fence_proxy_async();
        }
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&bars.full[s]);
        if (tid == 0) TRACE(0 * 512 + g * 4 + 3);
       }
        if (++s == (uint32_t)NST) { s = 0; sph ^= 1; }
        if (++r == (uint32_t)NRAW) { r = 0; rph ^= 1; }
      }
    }
  } else if (warp == kK3MmaWarp) {
    {
      // ================================================================ MMA issue (warp-converged, one elected lane issues)
      const uint32_t lbo_a = (kTcM / 8) * 128, lbo_b = (uint32_t)(npad / 8) * 128, sbo = 128;
      const uint64_t a_hi0 = ptx::smem_desc_noswizzle(ptx::smem_u32(sA), lbo_a, sbo);
      const uint64_t a_lo0 = ptx::smem_desc_noswizzle(ptx::smem_u32(sA) + a_blk, lbo_a, sbo);
      const uint64_t b_hi0 = ptx::smem_desc_noswizzle(ptx::smem_u32(sB), lbo_b, sbo);
      const uint64_t b_lo0 = ptx::smem_desc_noswizzle(ptx::smem_u32(sB) + b_blk, lbo_b, sbo);
      uint32_t g = 0, it = 0, s = 0, sph = 0;
      for (int i = blockIdx.x; i < nitems; i += gridDim.x, ++it) {
        int b, p0, h0, kspec, nsp, t, n0, ncols;
        item_geom(i, t, n0, ncols);
        tile_geom(t, b, p0, h0, kspec, nsp);
        const uint32_t idesc = ptx::idesc_tf32(kTcM, ncols);
        const uint64_t bcol = (uint64_t)(((uint32_t)(n0 >> 3) * 128u) >> 4);  // the item's first 8-column block of the B stage
        const uint32_t a = it & 1;
        TRACE(3 * 512 + 256 + it * 2 + 0);
        if (it >= 2) ptx::mbar_wait(&bars.acc_empty[a], ((it / 2) - 1) & 1);
        TRACE(3 * 512 + 256 + it * 2 + 1);
        ptx::tc_fence_after();
        const uint32_t dcol = tmem_base + a * 256;
        const int nch = nsp + nx;
        for (int c = 0; c < nch; ++c, ++g) {
          TRACE(1 * 512 + g * 4 + 0);
          TRACE2(g * 10 + 0);
          ptx::mbar_wait(&bars.full[s], sph);
          TRACE(1 * 512 + g * 4 + 1);
          TRACE2(g * 10 + 1);
          TRACE(1 * 512 + g * 4 + 2);
          TRACE2(g * 10 + 2);
          ptx::tc_fence_after();
          // only the 14-bit start-address field of a descriptor changes between stages / K steps
          const uint64_t da = (uint64_t)((s * a_stage) >> 4), db = (uint64_t)((s * b_stage) >> 4) + bcol;
#pragma unroll
          for (int ks = 0; ks < kTcBK / 8; ++ks) {
            const uint64_t ka = da + (uint64_t)((ks * 2 * lbo_a) >> 4), kb = db + (uint64_t)((ks * 2 * lbo_b) >> 4);
            if (dbg & 1) {
            } else if constexpr (kTA) {
              const uint32_t ta = tmem_base + s * kTcBK + ks * 8;
              if (p.single_pass) {
                ptx::mma_tf32_ta_ws(dcol, ta + ta_hi, b_hi0 + kb, idesc, (c | ks) != 0 ? 1u : 0u);
              } else {
                ptx::mma_tf32_ta_ws(dcol, ta + ta_lo, b_hi0 + kb, idesc, (c | ks) != 0 ? 1u : 0u);
                TRACE2(g * 10 + 3 + ks * 3);
                ptx::mma_tf32_ta_ws(dcol, ta + ta_hi, b_lo0 + kb, idesc, 1u);
                TRACE2(g * 10 + 4 + ks * 3);
                ptx::mma_tf32_ta_ws(dcol, ta + ta_hi, b_hi0 + kb, idesc, 1u);
                TRACE2(g * 10 + 5 + ks * 3);
              }
            } else if (p.single_pass) {
              ptx::mma_tf32_ws(dcol, a_hi0 + ka, b_hi0 + kb, idesc, (c | ks) != 0 ? 1u : 0u);
            } else {
              ptx::mma_tf32_ws(dcol, a_lo0 + ka, b_hi0 + kb, idesc, (c | ks) != 0 ? 1u : 0u);
              ptx::mma_tf32_ws(dcol, a_hi0 + ka, b_lo0 + kb, idesc, 1u);
              ptx::mma_tf32_ws(dcol, a_hi0 + ka, b_hi0 + kb, idesc, 1u);
            }
          }
          if (dbg & 16) ptx::mbar_arrive_ws(&bars.empty[s]); else
          ptx::tc_commit_ws(&bars.empty[s]);
          TRACE(1 * 512 + g * 4 + 3);
          TRACE2(g * 10 + 9);
          if (++s == (uint32_t)NST) { s = 0; sph ^= 1; }
        }
        ptx::tc_commit_ws(&bars.acc_full[a]);
      }
    }
  } else if (warp == kK3RawWarp) {
    {
      // ================================================================ raw ring issue (warp-converged) (activation rows / Z rows)
      // The activations come from HBM (not L2): at ~1.2 us loaded latency a 3-slot ring (24 KB in flight per SM) paced
      // the whole kernel at ~1250 cycles per chunk; the ring is now as deep as shared memory allows (up to 8 slots).
      uint32_t g = 0, r = 0, rph = 1;
      // last read of the activations in the forward chain (K1 kept them in L2 with evict-last): evict-first frees the
      // lines for the part of the input that has not been read yet
      const uint64_t pol_last_use = ptx::l2_policy_evict_first();
      for (int i = blockIdx.x; i < nitems; i += gridDim.x) {
        int b, p0, h0, kspec, nsp, t, n0, ncols;
        item_geom(i, t, n0, ncols);
        tile_geom(t, b, p0, h0, kspec, nsp);
        const int npx = (HW - p0 < kTcM) ? (HW - p0) : kTcM;
        // (a TMA L2 prefetch of this tile's residual [N][128 px], issued here, was measured on B200: 72.3 -> 77.2 us, like
        // round 1's prefetch by an idle warp -- the epilogue is not waiting on HBM latency alone; removed)
        for (int c = 0; c < nsp + nx; ++c, ++g) {
          TRACE(2 * 512 + g * 4 + 0);
          if (g >= (uint32_t)NRAW) ptx::mbar_wait(&bars.raw_empty[r], rph);
          TRACE(2 * 512 + g * 4 + 1);
          float* dst = raw + (size_t)r * slot_f;
          if (c < nsp) {
            // 16 consecutive (row, j) lines of Z are contiguous in memory: ONE bulk copy per spectral chunk
            const int nk = (kspec - c * kTcBK < kTcBK) ? (kspec - c * kTcBK) : kTcBK;
            ptx::mbar_arrive_expect_tx_ws(&bars.raw_full[r], (uint32_t)nk * p.N * 4);
            ptx::bulk_g2s_ws(dst, p.Z + (((size_t)b * p.H + h0) * J + (size_t)c * kTcBK) * p.N, (uint32_t)nk * p.N * 4,
                          &bars.raw_full[r]);
          } else {
            const int cx = c - nsp;
            const int nk = (p.K - cx * kTcBK < kTcBK) ? (p.K - cx * kTcBK) : kTcBK;
            if (cx < ntmap_chunks) {
              // one 2-D TMA box: 16 channel rows x 128 pixels (SASS UTMALDG)
              if (dbg & 4) {
                ptx::mbar_arrive_ws(&bars.raw_full[r]);
              } else {
              ptx::mbar_arrive_expect_tx_ws(&bars.raw_full[r], (uint32_t)kTcBK * kTcM * 4);
              ptx::tma_load_2d_ws_hint(dst, &tmap_x0, p0, b * p.C0 + cx * kTcBK, &bars.raw_full[r], pol_last_use);
              }
            } else {
              ptx::mbar_arrive_expect_tx_ws(&bars.raw_full[r], (uint32_t)nk * npx * 4);
              for (int kk = 0; kk < nk; ++kk) {
                const int k = cx * kTcBK + kk;
                const float* src = (k < p.C0) ? p.x0 + ((size_t)b * p.C0 + k) * HW : p.x1 + ((size_t)b * p.C1 + (k - p.C0)) * HW;
                ptx::bulk_g2s_ws(dst + kk * kTcM, src + p0, (uint32_t)npx * 4, &bars.raw_full[r]);
              }
            }
          }
          TRACE(2 * 512 + g * 4 + 2);
          if (++r == (uint32_t)NRAW) { r = 0; rph ^= 1; }
        }
      }
    }
  } else if (warp == kK3WgtWarp) {
    {
      // ================================================================ B ring issue (warp-converged) (packed weight chunks)
      uint32_t g = 0, s = 0, sph = 1;
      for (int i = blockIdx.x; i < nitems; i += gridDim.x) {
        int b, p0, h0, kspec, nsp, t, n0, ncols;
        item_geom(i, t, n0, ncols);
        tile_geom(t, b, p0, h0, kspec, nsp);
        for (int c = 0; c < nsp + nx; ++c, ++g) {
          if (g >= (uint32_t)NST) ptx::mbar_wait(&bars.empty[s], sph);
          if (c < nsp) {
            ptx::mbar_arrive_ws(&bars.full[s]);            // B of a spectral chunk is written by the convert warps
          } else if (dbg & 2) {
            ptx::mbar_arrive_ws(&bars.full[s]);
          } else {
            ptx::mbar_arrive_expect_tx_ws(&bars.full[s], b_stage);
            ptx::bulk_g2s_ws(sB + s * b_stage, p.wpack + (size_t)(c - nsp) * (b_stage / 4), b_stage, &bars.full[s]);
          }
          if (++s == (uint32_t)NST) { s = 0; sph ^= 1; }
        }
      }
    }
  } else if (warp >= 8 && warp < 8 + 4 * kK3EpiParts) {
    // ==================================================================== epilogue: 4 warps per TMEM lane quadrant
    // Code-size discipline: the exact-erf GELU is ~40 instructions, so the column loop is NOT unrolled beyond 8
    // (a fully unrolled 32-column body was > 30 KB of SASS and ran out of the instruction cache: 270 cycles/output).
    const int quad = warp & 3, part = (warp - 8) >> 2;
    const int N = p.N;
    uint32_t it = 0;
    for (int i = blockIdx.x; i < nitems; i += gridDim.x, ++it) {
      int b, p0, h0, kspec, nsp, t, nbase, ncols;
      item_geom(i, t, nbase, ncols);
      tile_geom(t, b, p0, h0, kspec, nsp);
      const int nq = ncols / 8;                                      // 8-column groups of this item (TMEM columns [0, ncols))
      const int per = (nq + kK3EpiParts - 1) / kK3EpiParts;
      const int qbeg = (part * per < nq) ? part * per : nq, qend = (qbeg + per < nq) ? qbeg + per : nq;
      const int pp = p0 + quad * 32 + lane;
      const bool pvalid = pp < HW;
      // 32-bit element offsets (the host guarantees B * out_bs < 2^32): one IMAD.WIDE per access instead of a 64-bit
      // multiply-add chain, and the offset of channel n is shared by the residual load and the two stores
      const uint32_t obase = (uint32_t)b * (uint32_t)p.out_bs + (uint32_t)pp;
      const uint32_t uHW = (uint32_t)HW;
      const uint32_t a = it & 1;
      const bool has_res = p.res != nullptr && pvalid;
      const bool has_pre = p.pre != nullptr, do_gelu = p.act == PDES_ACT_GELU;
      // the residual (U-Net branch) is prefetched TWO column groups ahead: the epilogue is the pacing role of this kernel
      // and, with one group in flight per warp (16 KB per SM), it ran at the latency of its own loads (~2900 cycles per group)
      float cur[8], nxt[8], nx2[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const int n = nbase + qbeg * 8 + e;
        cur[e] = (has_res && qbeg < qend && n < N) ? __ldcs(p.res + (obase + (uint32_t)n * uHW)) : 0.0f;
        nxt[e] = (has_res && qbeg + 1 < qend && n + 8 < N) ? __ldcs(p.res + (obase + (uint32_t)(n + 8) * uHW)) : 0.0f;
      }
      if (tid == 256) TRACE(3 * 512 + it * 4 + 0);
      ptx::mbar_wait(&bars.acc_full[a], (it / 2) & 1);
      if (tid == 256) TRACE(3 * 512 + it * 4 + 1);
      ptx::tc_fence_after();
#ifdef PDES_TC_TRACE
      if (p.act == 77) {                                            // debug: no epilogue work at all
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&bars.acc_empty[a]);
        continue;
      }
#endif
      const uint32_t tbase = tmem_base + a * 256 + ((uint32_t)(quad * 32) << 16);
#pragma unroll 1
      for (int qi = qbeg; qi < qend; ++qi) {
        const int c0 = qi * 8, n0 = nbase + c0;                      // TMEM column and output channel of the group
#pragma unroll
        for (int e = 0; e < 8; ++e) {                                // prefetch the residual of the group after the next
          const int n = n0 + 16 + e;
          nx2[e] = (has_res && qi + 2 < qend && n < N) ? __ldcs(p.res + (obase + (uint32_t)n * uHW)) : 0.0f;
        }
        uint32_t r[8];
        ptx::tmem_ld8(tbase + (uint32_t)c0, r);
        ptx::tmem_ld_wait();
        if (qi + 1 == qend) {                                        // accumulator fully read: hand it back to the MMA warp
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive(&bars.acc_empty[a]);
          if (tid == 256) TRACE(3 * 512 + it * 4 + 2);
        }
#ifdef PDES_TC_TRACE
        if (p.act == 78) { for (int e = 0; e < 8; ++e) cur[e] = nxt[e] + __uint_as_float(r[e]); continue; }   // debug: TMEM loads only
#endif
        if (pvalid) {
          const uint32_t o0 = obase + (uint32_t)n0 * uHW;
          float* const po = p.out;
          float* const pq = p.pre;
          if (n0 + 8 <= N) {                                         // whole group valid: no per-element guards
            float bz[8];
            {
              const float4 b0 = *reinterpret_cast<const float4*>(bias_s + n0);          // warp-broadcast shared-memory loads
              const float4 b1 = *reinterpret_cast<const float4*>(bias_s + n0 + 4);
              bz[0] = b0.x; bz[1] = b0.y; bz[2] = b0.z; bz[3] = b0.w; bz[4] = b1.x; bz[5] = b1.y; bz[6] = b1.z; bz[7] = b1.w;
            }
#pragma unroll
            for (int e = 0; e < 8; e += 2) {                         // two outputs per step: packed-FFMA2 GELU
              float2 v = make_float2(cur[e] + bz[e], cur[e + 1] + bz[e + 1]);
              ffma2(v, make_float2(__uint_as_float(r[e]), __uint_as_float(r[e + 1])), make_float2(1.0f, 1.0f));
              const uint32_t oa = o0 + (uint32_t)e * uHW, ob = oa + uHW;
              if (has_pre) { __stcs(pq + oa, v.x); __stcs(pq + ob, v.y); }
              if (do_gelu) v = gelu_fast2_f(v);
              __stcs(po + oa, v.x);                                  // streaming: the output must not evict the input
              __stcs(po + ob, v.y);
            }
          } else {
            for (int e = 0; e < 8 && n0 + e < N; ++e) {
              float v = __uint_as_float(r[e]) + cur[e];
              v += bias_s[n0 + e];
              if (has_pre) __stcs(pq + (o0 + (uint32_t)e * uHW), v);
              if (do_gelu) v = gelu_fast_f(v);
              __stcs(po + (o0 + (uint32_t)e * uHW), v);
            }
          }
        }
#pragma unroll
        for (int e = 0; e < 8; ++e) { cur[e] = nxt[e]; nxt[e] = nx2[e]; }
      }
      if (tid == 256) TRACE(3 * 512 + it * 4 + 3);
      if (qbeg >= qend) {                                            // no column group for this warp (N <= 8)
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&bars.acc_empty[a]);         // still only after acc_full: keeps phases in step
      }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == kK3MmaWarp) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, 512);
  }
}


// ----------------------------------------------------------------------------------------------------------------
// 1x1-conv weight gradient on tcgen05:  dW[o][i] = sum_{b,p} g[b][o][p] * xin[b][i][p]   (+ ones column = dbias)
// GEMM: M = output channels (two overlapping 128-row tiles), N = input channels + 1, K = pixels.  Both operands are
// pixel-contiguous in HBM, i.e. natively K-major: 2-D TMA boxes of [channels x 16 pixels] land in a raw ring, 256
// convert threads split hi/lo and write 16-byte core-matrix rows (K-adjacent core matrices are placed LBO = n*128+32
// bytes apart so the 4 k-quads of a row fall in different banks), one thread issues 12 MMAs per 16-pixel chunk.
// A CTA accumulates a contiguous range of pixel chunks in TMEM and writes ONE partial [M][N]; a small kernel reduces
// the partials in a fixed order (deterministic).
constexpr int kWgtThreads = 320;      // warps 0-7 convert (+ epilogue), warp 8 MMA, warp 9 TMA issue
constexpr int kWgtStages = 3, kWgtRaw = 2;
__host__ __device__ inline int wgt_off_x0(int M) { return (M * kTcBK + 31) & ~31; }                 // floats, 128-byte aligned
__host__ __device__ inline int wgt_raw_floats(int M, int K) { return (wgt_off_x0(M) + K * kTcBK + 31) & ~31; }

struct WgtParams {
  const float* x1; int C1;
  float* part;
  int B, M, C0, K, HW, npadN, mrows, total_chunks, per_cta;
  int x0_ld, x0_off;    // x0 is the channel range [x0_off, x0_off + C0) of a tensor with x0_ld channels per sample
};

__global__ void __launch_bounds__(kWgtThreads, 1)
k_wgrad_tc(WgtParams p, const __grid_constant__ CUtensorMap tmap_g, const __grid_constant__ CUtensorMap tmap_x0) {
  PDES_DYN_SMEM(unsigned char, smem_raw);
  unsigned char* base = ptx::align_smem_1024(smem_raw);
  const uint32_t lbo_a = (uint32_t)(p.mrows / 8) * 128 + 32, lbo_b = (uint32_t)(p.npadN / 8) * 128 + 32;
  const uint32_t a_blk = 4 * lbo_a, b_blk = 4 * lbo_b;
  const uint32_t stage_bytes = 2 * a_blk + 2 * b_blk;
  const int raw_f = wgt_raw_floats(p.M, p.K);                     // floats per raw slot
  const int off_x0 = wgt_off_x0(p.M);
  float* raw = reinterpret_cast<float*>(base + (size_t)kWgtStages * stage_bytes);
  __shared__ __align__(8) unsigned long long full_bar[kWgtStages], empty_bar[kWgtStages], raw_full[kWgtRaw],
      raw_empty[kWgtRaw], done_bar;
  __shared__ uint32_t tmem_slot;

  // (warp index as a warp-uniform value: the MMA and TMA roles run warp-converged with one elected lane issuing, see
  // DESIGN.md 3.1-2; as `if (lane == 0)` roles every tcgen05.mma / TMA issue sat in a waterfall loop)
  const int tid = threadIdx.x, warp = ptx::uniform_warp_idx(), lane = tid & 31;
  const int cpi = p.HW / kTcBK;                                   // chunks per image
  const int c_beg = blockIdx.x * p.per_cta;
  const int c_end = (c_beg + p.per_cta < p.total_chunks) ? (c_beg + p.per_cta) : p.total_chunks;
  const int nch = c_end - c_beg;

  if (tid == 0) {
    // one arrival per convert WARP (after __syncwarp), not per thread: 256 arrivals on one barrier word serialise
    for (int i = 0; i < kWgtStages; ++i) { ptx::mbar_init(&full_bar[i], 8); ptx::mbar_init(&empty_bar[i], 1); }
    for (int i = 0; i < kWgtRaw; ++i) { ptx::mbar_init(&raw_full[i], 1); ptx::mbar_init(&raw_empty[i], 8); }
    ptx::mbar_init(&done_bar, 1);
    ptx::fence_mbar_init();
  }
  if (warp == 8) {
    ptx::tmem_alloc(&tmem_slot, 512);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = tmem_slot;

  if (warp < 8) {
    // ================================================================== convert
    const int nitems = (p.mrows + p.npadN) * 4;
    for (int c = 0; c < nch; ++c) {
      const uint32_t s = c % kWgtStages, r = c % kWgtRaw;
      if (c >= kWgtStages) ptx::mbar_wait(&empty_bar[s], ((c / kWgtStages) - 1) & 1);
      ptx::mbar_wait(&raw_full[r], (c / kWgtRaw) & 1);
      const float* rw = raw + (size_t)r * raw_f;
      unsigned char* st = base + (size_t)s * stage_bytes;
      for (int it = tid; it < nitems; it += 256) {
        const int row = it >> 2, kq = it & 3;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        unsigned char* dst;
        if (row < p.mrows) {                                       // A: g rows
          if (row < p.M) v = *reinterpret_cast<const float4*>(rw + row * kTcBK + kq * 4);
          dst = st + (uint32_t)kq * lbo_a + (uint32_t)(row >> 3) * 128 + (uint32_t)(row & 7) * 16;
        } else {                                                   // B: input rows, then the all-ones row
          const int n = row - p.mrows;
          if (n < p.K) v = *reinterpret_cast<const float4*>(rw + off_x0 + n * kTcBK + kq * 4);
          else if (n == p.K) v = make_float4(1.f, 1.f, 1.f, 1.f);
          dst = st + 2 * a_blk + (uint32_t)kq * lbo_b + (uint32_t)(n >> 3) * 128 + (uint32_t)(n & 7) * 16;
        }
        float4 hi, lo;
        hi.x = tf32_hi(v.x); lo.x = v.x - hi.x;
        hi.y = tf32_hi(v.y); lo.y = v.y - hi.y;
        hi.z = tf32_hi(v.z); lo.z = v.z - hi.z;
        hi.w = tf32_hi(v.w); lo.w = v.w - hi.w;
        *reinterpret_cast<float4*>(dst) = hi;
        *reinterpret_cast<float4*>(dst + (row < p.mrows ? a_blk : b_blk)) = lo;
      }
      // (both arrivals come after the st.shared of the converted data: the raw slot is released only once its values
      // have been consumed)
      ptx::fence_proxy_async();
      __syncwarp();
      if (lane == 0) {
        ptx::mbar_arrive(&raw_empty[r]);
        ptx::mbar_arrive(&full_bar[s]);
      }
    }
    // ================================================================== epilogue: one partial per CTA
    if (nch > 0) {
      ptx::mbar_wait(&done_bar, 0);
      ptx::tc_fence_after();
    }
    const int quad = warp & 3, tile = warp >> 2;                   // warps 0-3: rows [0,128), warps 4-7: second tile
    const int o = (tile == 0 ? 0 : p.mrows - 128) + quad * 32 + lane;
    const bool wr = (tile == 0) ? (o < p.M) : (p.M > 128 && o >= 128 && o < p.M);
    float* dstrow = p.part + ((size_t)blockIdx.x * p.M + (wr ? o : 0)) * p.npadN;
    const uint32_t tb = tmem_base + (uint32_t)tile * 256 + ((uint32_t)(quad * 32) << 16);
#pragma unroll 1
    for (int n0 = 0; n0 < p.npadN; n0 += 8) {
      uint32_t rr[8];
      if (nch > 0) {
        ptx::tmem_ld8(tb + (uint32_t)n0, rr);
        ptx::tmem_ld_wait();
      } else {
#pragma unroll
        for (int e = 0; e < 8; ++e) rr[e] = 0u;
      }
      if (wr) {
        *reinterpret_cast<float4*>(dstrow + n0) = make_float4(__uint_as_float(rr[0]), __uint_as_float(rr[1]), __uint_as_float(rr[2]), __uint_as_float(rr[3]));
        *reinterpret_cast<float4*>(dstrow + n0 + 4) = make_float4(__uint_as_float(rr[4]), __uint_as_float(rr[5]), __uint_as_float(rr[6]), __uint_as_float(rr[7]));
      }
    }
    ptx::tc_fence_before();
  } else if (warp == 8) {
    if (nch > 0) {
      // ================================================================ MMA issue (warp-converged, one elected lane issues)
      const uint32_t idesc = ptx::idesc_tf32(128, p.npadN);
      const uint32_t row1 = (uint32_t)((p.mrows - 128) / 8) * 128;           // byte offset of the second 128-row tile
      const uint32_t sbase = ptx::smem_u32(base);
      const uint64_t a_hi0 = ptx::smem_desc_noswizzle(sbase, lbo_a, 128);
      const uint64_t a_lo0 = ptx::smem_desc_noswizzle(sbase + a_blk, lbo_a, 128);
      const uint64_t b_hi0 = ptx::smem_desc_noswizzle(sbase + 2 * a_blk, lbo_b, 128);
      const uint64_t b_lo0 = ptx::smem_desc_noswizzle(sbase + 2 * a_blk + b_blk, lbo_b, 128);
      const int ntile = p.mrows > 128 ? 2 : 1;
      for (int c = 0; c < nch; ++c) {
        const uint32_t s = c % kWgtStages;
        ptx::mbar_wait(&full_bar[s], (c / kWgtStages) & 1);
        ptx::tc_fence_after();
        const uint64_t ds = (uint64_t)((s * stage_bytes) >> 4);
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
          const uint64_t kb = ds + (uint64_t)((ks * 2 * lbo_b) >> 4);
          for (int t = 0; t < ntile; ++t) {
            const uint64_t ka = ds + (uint64_t)((ks * 2 * lbo_a + (t ? row1 : 0u)) >> 4);
            const uint32_t dcol = tmem_base + (uint32_t)t * 256;
            const uint32_t accf = (c | ks) != 0 ? 1u : 0u;
            ptx::mma_tf32_ws(dcol, a_lo0 + ka, b_hi0 + kb, idesc, accf);
            ptx::mma_tf32_ws(dcol, a_hi0 + ka, b_lo0 + kb, idesc, 1u);
            ptx::mma_tf32_ws(dcol, a_hi0 + ka, b_hi0 + kb, idesc, 1u);
          }
        }
        ptx::tc_commit_ws(&empty_bar[s]);
      }
      ptx::tc_commit_ws(&done_bar);
    }
  } else {
    // ================================================================== TMA issue (warp-converged)
    int b = c_beg / cpi, pc = c_beg - b * cpi;                    // counters instead of a division per chunk
    for (int c = 0; c < nch; ++c) {
      const uint32_t r = c % kWgtRaw;
      if (c >= kWgtRaw) ptx::mbar_wait(&raw_empty[r], ((c / kWgtRaw) - 1) & 1);
      const int px = pc * kTcBK;
      float* dst = raw + (size_t)r * raw_f;
      ptx::mbar_arrive_expect_tx_ws(&raw_full[r], (uint32_t)(p.M + p.K) * kTcBK * 4);
      ptx::tma_load_2d_ws(dst, &tmap_g, px, b * p.M, &raw_full[r]);
      ptx::tma_load_2d_ws(dst + off_x0, &tmap_x0, px, b * p.x0_ld + p.x0_off, &raw_full[r]);
      for (int k = 0; k < p.C1; ++k)
        ptx::bulk_g2s_ws(dst + off_x0 + (p.C0 + k) * kTcBK, p.x1 + ((size_t)b * p.C1 + k) * p.HW + px, kTcBK * 4, &raw_full[r]);
      if (++pc == cpi) { pc = 0; ++b; }
    }
  }
  __syncthreads();
  if (warp == 8) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, 512);
  }
}

__global__ void __launch_bounds__(256)
k_wgrad_reduce_ld(const float* __restrict__ part, int nslab, int M, int K, int ld, float* __restrict__ dW, int ldw,
                  float* __restrict__ dbias) {
  const int n = M * (K + 1);
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n) return;
  const int o = idx / (K + 1), i = idx % (K + 1);
  float sum = 0.0f;
  for (int s = 0; s < nslab; ++s) sum += __ldg(part + ((size_t)s * M + o) * ld + i);
  if (i < K) {
    if (dW != nullptr) dW[(size_t)o * ldw + i] = sum;
  } else if (dbias != nullptr) {
    dbias[o] = sum;
  }
}


// ----------------------------------------------------------------------------------------------------------------
// U-Net branch: 3x3 *valid* convolution forward as an implicit GEMM on the same tcgen05
// pipeline (3xTF32).  M = 128 output pixels (a tile of 8 rows x 16 columns), N = Cout, K = Cin * 9.
//   raw ring : one 3-D TMA box [16 channels][10 rows][20 columns] per channel chunk (the tile's input patch)
//   convert  : for each of the 9 taps (ky,kx) the 128 threads gather their pixel's 16 channels from the patch,
//              split hi/lo and write the canonical A stage -- the patch is read from HBM once and reused 9 times
//   B ring   : weights packed per (channel chunk, tap) into canonical (hi, lo) blocks, one bulk copy per step
//   MMA / epilogue as in the GEMM above (double-buffered TMEM accumulator, bias add, coalesced stores).
constexpr int kCvTH = 8, kCvTW = 16, kCvBH = kCvTH + 2, kCvBW = 20;   // output tile 8x16, input box 10x20 (80-byte rows)
constexpr int kCvRawF = kTcBK * kCvBH * kCvBW;                            // 3200 floats = 12.8 KB per raw slot

struct ConvParams {
  const float* wpack;
  const float* bias;
  float* out;
  int B, Cin, N, npad, H, W, Ho, Wo, tiles_y, tiles_x;
};

__global__ void __launch_bounds__(256)
k_pack_conv3x3_tf32(const float* __restrict__ Wt, int N, int Cin, int npad, int ncc, float* __restrict__ out) {
  const int total = ncc * 9 * npad * kTcBK;
  const int lbo_f = (npad / 8) * 32;
  const size_t blk = (size_t)npad * kTcBK;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const int n = idx % npad;
    const int kk = (idx / npad) % kTcBK;
    const int tap = (idx / (npad * kTcBK)) % 9;
    const int cc = idx / (npad * kTcBK * 9);
    const int c = cc * kTcBK + kk;
    const float w = (c < Cin && n < N) ? __ldg(Wt + ((size_t)n * Cin + c) * 9 + tap) : 0.0f;
    const float hi = tf32_hi(w);
    const size_t off = (size_t)(kk / 4) * lbo_f + (size_t)(n / 8) * 32 + (n % 8) * 4 + (kk % 4);
    const size_t chunk = (size_t)cc * 9 + tap;
    out[chunk * 2 * blk + off] = hi;
    out[chunk * 2 * blk + blk + off] = w - hi;
  }
}

__global__ void __launch_bounds__(kV3Threads, 1)
k_conv3x3_tc(ConvParams p, const __grid_constant__ CUtensorMap tmap_x) {
  PDES_DYN_SMEM(unsigned char, smem_raw);
  unsigned char* base = ptx::align_smem_1024(smem_raw);
  const int npad = p.npad;
  const uint32_t a_blk = kTcM * kTcBK * 4;
  const uint32_t b_blk = (uint32_t)npad * kTcBK * 4;
  const uint32_t a_stage = 2 * a_blk, b_stage = 2 * b_blk;
  unsigned char* sA = base;
  unsigned char* sB = sA + kV3ASt * a_stage;
  float* raw = reinterpret_cast<float*>(sB + kV3BSt * b_stage);
  __shared__ __align__(8) V3Bars bars;
  __shared__ uint32_t tmem_slot;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int ncc = tc_nchunks(p.Cin);
  const int nch = ncc * 9;                                        // MMA chunks per tile
  const int tiles_img = p.tiles_y * p.tiles_x;
  const int ntiles = p.B * tiles_img;

  if (tid == 0) {
    for (int i = 0; i < kV3ASt; ++i) { ptx::mbar_init(&bars.a_full[i], 128); ptx::mbar_init(&bars.a_empty[i], 1); }
    for (int i = 0; i < kV3BSt; ++i) { ptx::mbar_init(&bars.b_full[i], 1); ptx::mbar_init(&bars.b_empty[i], 1); }
    for (int i = 0; i < kV3Raw; ++i) { ptx::mbar_init(&bars.raw_full[i], 1); ptx::mbar_init(&bars.raw_empty[i], 128); }
    for (int i = 0; i < 2; ++i) { ptx::mbar_init(&bars.acc_full[i], 1); ptx::mbar_init(&bars.acc_empty[i], 32 * 4 * kV3EpiParts); }
    ptx::fence_mbar_init();
  }
  if (warp == 4) {
    ptx::tmem_alloc(&tmem_slot, 512);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = tmem_slot;

  auto tile_geom = [&](int t, int& b, int& y0, int& x0) {
    b = t / tiles_img;
    const int r = t - b * tiles_img;
    y0 = (r / p.tiles_x) * kCvTH;
    x0 = (r % p.tiles_x) * kCvTW;
  };

  if (warp < 4) {
    // ================================================================== convert (gather one tap per A stage)
    const uint32_t row_off = (uint32_t)(tid >> 3) * 128 + (uint32_t)(tid & 7) * 16;
    const int pr = tid >> 4, pc = tid & 15;                        // pixel of this thread inside the 8x16 tile
    uint32_t g = 0, gr = 0;
    for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
      for (int cc = 0; cc < ncc; ++cc, ++gr) {
        const uint32_t r = gr % kV3Raw;
        ptx::mbar_wait(&bars.raw_full[r], (gr / kV3Raw) & 1);
        const float* patch = raw + (size_t)r * kCvRawF + pr * kCvBW + pc;
        const int nvalid = (p.Cin - cc * kTcBK < kTcBK) ? (p.Cin - cc * kTcBK) : kTcBK;
#pragma unroll 1
        for (int tap = 0; tap < 9; ++tap, ++g) {
          const uint32_t s = g % kV3ASt;
          if (g >= kV3ASt) ptx::mbar_wait(&bars.a_empty[s], ((g / kV3ASt) - 1) & 1);
          unsigned char* st = sA + s * a_stage;
          const float* src = patch + (tap / 3) * kCvBW + (tap % 3);
          float v[kTcBK];
#pragma unroll
          for (int kk = 0; kk < kTcBK; ++kk) v[kk] = (kk < nvalid) ? src[kk * (kCvBH * kCvBW)] : 0.0f;
#pragma unroll
          for (int qd = 0; qd < kTcBK / 4; ++qd) {
            float4 hi, lo;
            hi.x = tf32_hi(v[4 * qd + 0]); lo.x = v[4 * qd + 0] - hi.x;
            hi.y = tf32_hi(v[4 * qd + 1]); lo.y = v[4 * qd + 1] - hi.y;
            hi.z = tf32_hi(v[4 * qd + 2]); lo.z = v[4 * qd + 2] - hi.z;
            hi.w = tf32_hi(v[4 * qd + 3]); lo.w = v[4 * qd + 3] - hi.w;
            const uint32_t off = (uint32_t)qd * (kTcM / 8) * 128 + row_off;
            *reinterpret_cast<float4*>(st + off) = hi;
            *reinterpret_cast<float4*>(st + a_blk + off) = lo;
          }
          ptx::fence_proxy_async();
          ptx::mbar_arrive(&bars.a_full[s]);
        }
        ptx::mbar_arrive(&bars.raw_empty[r]);
      }
    }
  } else if (warp == 4) {
    if (lane == 0) {
      // ================================================================ MMA issue
      const uint32_t idesc = ptx::idesc_tf32(kTcM, npad);
      const uint32_t lbo_a = (kTcM / 8) * 128, lbo_b = (uint32_t)(npad / 8) * 128, sbo = 128;
      const uint64_t a_hi0 = ptx::smem_desc_noswizzle(ptx::smem_u32(sA), lbo_a, sbo);
      const uint64_t a_lo0 = ptx::smem_desc_noswizzle(ptx::smem_u32(sA) + a_blk, lbo_a, sbo);
      const uint64_t b_hi0 = ptx::smem_desc_noswizzle(ptx::smem_u32(sB), lbo_b, sbo);
      const uint64_t b_lo0 = ptx::smem_desc_noswizzle(ptx::smem_u32(sB) + b_blk, lbo_b, sbo);
      uint32_t g = 0, it = 0;
      for (int t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
        const uint32_t a = it & 1;
        if (it >= 2) ptx::mbar_wait(&bars.acc_empty[a], ((it / 2) - 1) & 1);
        ptx::tc_fence_after();
        const uint32_t dcol = tmem_base + a * 256;
        for (int c = 0; c < nch; ++c, ++g) {
          const uint32_t s = g % kV3ASt, q = g % kV3BSt;
          ptx::mbar_wait(&bars.a_full[s], (g / kV3ASt) & 1);
          ptx::mbar_wait(&bars.b_full[q], (g / kV3BSt) & 1);
          ptx::tc_fence_after();
          const uint64_t da = (uint64_t)((s * a_stage) >> 4), db = (uint64_t)((q * b_stage) >> 4);
#pragma unroll
          for (int ks = 0; ks < kTcBK / 8; ++ks) {
            const uint64_t ka = da + (uint64_t)((ks * 2 * lbo_a) >> 4), kb = db + (uint64_t)((ks * 2 * lbo_b) >> 4);
            ptx::mma_tf32(dcol, a_lo0 + ka, b_hi0 + kb, idesc, (c | ks) != 0 ? 1u : 0u);
            ptx::mma_tf32(dcol, a_hi0 + ka, b_lo0 + kb, idesc, 1u);
            ptx::mma_tf32(dcol, a_hi0 + ka, b_hi0 + kb, idesc, 1u);
          }
          ptx::tc_commit(&bars.a_empty[s]);
          ptx::tc_commit(&bars.b_empty[q]);
        }
        ptx::tc_commit(&bars.acc_full[a]);
      }
    }
  } else if (warp == 5) {
    if (lane == 0) {
      // ================================================================ raw ring: one 3-D TMA box per channel chunk
      uint32_t gr = 0;
      for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
        int b, y0, x0;
        tile_geom(t, b, y0, x0);
        for (int cc = 0; cc < ncc; ++cc, ++gr) {
          const uint32_t r = gr % kV3Raw;
          if (gr >= kV3Raw) ptx::mbar_wait(&bars.raw_empty[r], ((gr / kV3Raw) - 1) & 1);
          ptx::mbar_arrive_expect_tx(&bars.raw_full[r], (uint32_t)kCvRawF * 4);
          ptx::tma_load_3d(raw + (size_t)r * kCvRawF, &tmap_x, x0, y0, b * p.Cin + cc * kTcBK, &bars.raw_full[r]);
        }
      }
    }
  } else if (warp == 6) {
    if (lane == 0) {
      // ================================================================ B ring: packed weights of (chunk, tap)
      uint32_t g = 0;
      for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
        for (int c = 0; c < nch; ++c, ++g) {
          const uint32_t q = g % kV3BSt;
          if (g >= kV3BSt) ptx::mbar_wait(&bars.b_empty[q], ((g / kV3BSt) - 1) & 1);
          ptx::mbar_arrive_expect_tx(&bars.b_full[q], b_stage);
          ptx::bulk_g2s(sB + q * b_stage, p.wpack + (size_t)c * (b_stage / 4), b_stage, &bars.b_full[q]);
        }
      }
    }
  } else if (warp >= 8) {
    // ==================================================================== epilogue
    const int quad = warp & 3, part = (warp - 8) >> 2;
    const int N = p.N;
    const int nq = (npad + 7) / 8;
    const int per = (nq + kV3EpiParts - 1) / kV3EpiParts;
    const int qbeg = (part * per < nq) ? part * per : nq, qend = (qbeg + per < nq) ? qbeg + per : nq;
    const int m = quad * 32 + lane, pr = m >> 4, pc = m & 15;
    const size_t plane = (size_t)p.Ho * p.Wo;
    uint32_t it = 0;
    for (int t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
      int b, y0, x0;
      tile_geom(t, b, y0, x0);
      const int y = y0 + pr, x = x0 + pc;
      const bool pvalid = y < p.Ho && x < p.Wo;
      const uint32_t a = it & 1;
      ptx::mbar_wait(&bars.acc_full[a], (it / 2) & 1);
      ptx::tc_fence_after();
      const uint32_t tbase = tmem_base + a * 256 + ((uint32_t)(quad * 32) << 16);
      float* obase = p.out + (size_t)b * N * plane + (size_t)y * p.Wo + x;
#pragma unroll 1
      for (int qi = qbeg; qi < qend; ++qi) {
        const int n0 = qi * 8;
        uint32_t r[8];
        ptx::tmem_ld8(tbase + (uint32_t)n0, r);
        ptx::tmem_ld_wait();
        if (qi + 1 == qend) {
          ptx::tc_fence_before();
          ptx::mbar_arrive(&bars.acc_empty[a]);
        }
        if (pvalid) {
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const int n = n0 + e;
            if (n < N) obase[(size_t)n * plane] = __uint_as_float(r[e]) + (p.bias != nullptr ? __ldg(p.bias + n) : 0.0f);
          }
        }
      }
      if (qbeg >= qend) {
        ptx::tc_fence_before();
        ptx::mbar_arrive(&bars.acc_empty[a]);
      }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, 512);
  }
}

#endif  // !PDES_CPU_EMU

int g_tc_mode =
#ifdef PDES_CPU_EMU
    0;
#else
    2;
#endif
int g_num_sms = 0;

#ifndef PDES_CPU_EMU
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn g_encode_tiled() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(sym);
  }
  return fn;
}
#endif

}  // namespace
}  // namespace pdes

extern "C" {

#ifdef PDES_TC_TRACE
int pdes_tc_trace_read(long long* host, int n) {
  return (int)cudaMemcpyFromSymbol(host, pdes::g_trace, sizeof(long long) * n);
}
#endif

void pdes_set_tensor_core_mode(int mode) {
#ifdef PDES_CPU_EMU
  (void)mode;
  pdes::g_tc_mode = 0;
#else
  pdes::g_tc_mode = (mode >= 1 && mode <= 3) ? mode : 0;
#endif
}

int pdes_get_tensor_core_mode(void) { return pdes::g_tc_mode; }

int pdes_gemm_tc_supported(int N, int K) { return (N >= 1 && N <= pdes::kTcMaxN && K >= 1) ? 1 : 0; }

/* The tensor-core kernel streams activation rows with 16-byte-granular bulk copies: needs H*W % 4 == 0 and 16-byte
 * aligned activation pointers; the spectral term is folded in as (rows per 128-pixel tile) * 2*m2 extra K steps, so
 * very narrow grids (many rows per tile) stay on the FFMA kernel. */
int pdes_inv_w_gemm_tc_ok(int N, int K, int H, int W, int m2, const float* x0, const float* x1) {
  if (!pdes_gemm_tc_supported(N, K)) return 0;
  if ((H * W) % 4 != 0) return 0;
  if ((reinterpret_cast<uintptr_t>(x0) & 15u) != 0 || (x1 != nullptr && (reinterpret_cast<uintptr_t>(x1) & 15u) != 0)) return 0;
  const int rows = (pdes::kTcM + W - 1) / W + 1;
  if (m2 > 0 && rows * 2 * m2 > 8 * pdes::kTcBK) return 0;
  return 1;
}

size_t pdes_wgrad_tc_workspace_floats(int M, int K) {
  if (M <= 0 || K <= 0 || M > 256 || K + 1 > 256) return 0;
  return (size_t)pdes::kWgtMaxCtas * M * pdes::tc_npad(K + 1);
}

/* dW / dbias of the 1x1 conv on tcgen05 (3xTF32).  Returns PDES_ERR_UNSUPPORTED when the shape does not fit the
 * tensor-core kernel (the caller then uses pdes_wgrad). */
}  // extern "C"

namespace pdes {
namespace {
int wgrad_tc_impl(const float* g, const float* x0, int x0_ld, int x0_off, int C0, const float* x1, int C1, float* dW, int ldw,
                  float* dbias, float* ws, int B, int M, int HW, void* stream) {
#ifdef PDES_CPU_EMU
  (void)x0_ld; (void)x0_off; (void)ldw;
  (void)g; (void)x0; (void)C0; (void)x1; (void)C1; (void)dW; (void)dbias; (void)ws; (void)B; (void)M; (void)HW; (void)stream;
  set_error("pdes_wgrad_tc: tcgen05 path is not available in the CPU emulation build");
  return PDES_ERR_UNSUPPORTED;
#else
  PDES_REQUIRE(g && x0 && ws, PDES_ERR_ARG, "pdes_wgrad_tc: null pointer");
  const int K = C0 + C1;
  const bool ok = B > 0 && M >= 8 && M <= 256 && C0 >= 1 && C0 <= 256 && C1 >= 0 && K + 1 <= 256 && HW % kTcBK == 0 &&
                  aligned16(g) && aligned16(x0) && (x1 == nullptr || aligned16(x1)) && (C1 == 0) == (x1 == nullptr) &&
                  g_encode_tiled() != nullptr;
  PDES_REQUIRE(ok, PDES_ERR_UNSUPPORTED, "pdes_wgrad_tc: shape not supported by the tensor-core kernel");
  if (g_num_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (g_num_sms <= 0) g_num_sms = 148;
  }
  WgtParams p;
  p.x1 = x1; p.C1 = C1; p.part = ws; p.B = B; p.M = M; p.C0 = C0; p.K = K; p.HW = HW;
  p.x0_ld = x0_ld; p.x0_off = x0_off;
  PDES_REQUIRE(x0_ld >= C0 && x0_off >= 0 && x0_off + C0 <= x0_ld && ldw >= K, PDES_ERR_ARG, "pdes_wgrad_tc: bad channel range / ldw");
  p.npadN = tc_npad(K + 1);
  const int mpad = (M + 7) & ~7;
  p.mrows = mpad > 128 ? mpad : 128;
  p.total_chunks = B * (HW / kTcBK);
  int G = g_num_sms < kWgtMaxCtas ? g_num_sms : kWgtMaxCtas;
  if (G > p.total_chunks) G = p.total_chunks;
  p.per_cta = ceil_div(p.total_chunks, G);
  G = ceil_div(p.total_chunks, p.per_cta);
  const uint32_t lbo_a = (uint32_t)(p.mrows / 8) * 128 + 32, lbo_b = (uint32_t)(p.npadN / 8) * 128 + 32;
  const size_t smem = (size_t)kWgtStages * 8 * (lbo_a + lbo_b) + (size_t)kWgtRaw * wgt_raw_floats(M, K) * 4 + 1024;
  PDES_REQUIRE(smem <= 227 * 1024, PDES_ERR_UNSUPPORTED, "pdes_wgrad_tc: needs %zu B of shared memory", smem);
  alignas(64) CUtensorMap tg, tx;
  memset(&tg, 0, sizeof(tg));
  memset(&tx, 0, sizeof(tx));
  {
    const cuuint64_t gdim[2] = {(cuuint64_t)HW, (cuuint64_t)B * (cuuint64_t)M};
    const cuuint64_t gstr[1] = {(cuuint64_t)HW * 4};
    const cuuint32_t box[2] = {(cuuint32_t)kTcBK, (cuuint32_t)M};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = g_encode_tiled()(&tg, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(g), gdim, gstr, box, estr,
                                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                        CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    PDES_REQUIRE(r == CUDA_SUCCESS, PDES_ERR_UNSUPPORTED, "pdes_wgrad_tc: cuTensorMapEncodeTiled(g) failed (%d)", (int)r);
  }
  {
    const cuuint64_t gdim[2] = {(cuuint64_t)HW, (cuuint64_t)B * (cuuint64_t)x0_ld};
    const cuuint64_t gstr[1] = {(cuuint64_t)HW * 4};
    const cuuint32_t box[2] = {(cuuint32_t)kTcBK, (cuuint32_t)C0};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = g_encode_tiled()(&tx, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(x0), gdim, gstr, box, estr,
                                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                        CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    PDES_REQUIRE(r == CUDA_SUCCESS, PDES_ERR_UNSUPPORTED, "pdes_wgrad_tc: cuTensorMapEncodeTiled(x0) failed (%d)", (int)r);
  }
  auto kfn = k_wgrad_tc;
  PDES_SET_SMEM(kfn, smem);
  PDES_LAUNCH(kfn, dim3((unsigned)G), dim3(kWgtThreads), smem, stream, p, tg, tx);
  if (int e = check_launch("pdes_wgrad_tc")) return e;
  auto rfn = k_wgrad_reduce_ld;
  const int n = M * (K + 1);
  PDES_LAUNCH(rfn, dim3((unsigned)ceil_div(n, 256)), dim3(256), 0, stream, ws, G, M, K, p.npadN, dW, ldw, dbias);
  return check_launch("pdes_wgrad_tc_reduce");
#endif
}
}  // namespace
}  // namespace pdes

extern "C" {

int pdes_wgrad_tc(const float* g, const float* x0, int C0, const float* x1, int C1, float* dW, float* dbias, float* ws,
                  int B, int M, int HW, void* stream) {
  return pdes::wgrad_tc_impl(g, x0, C0, 0, C0, x1, C1, dW, C0 + C1, dbias, ws, B, M, HW, stream);
}

/* Weight / bias gradient of a 1x1 conv for the channel range [c_off, c_off + Cn) of x [B][x_ld][HW]:
 *   dW[o * ldw + i] = sum_{b,p} g[b,o,p] * x[b, c_off + i, p]  (i < Cn <= 255),  dbias[o] = sum g  (optional).
 * Same tcgen05 kernel and workspace as pdes_wgrad_tc; wide inputs are covered by several calls. */
int pdes_wgrad_tc_range(const float* g, const float* x, int x_ld, int c_off, int Cn, float* dW, int ldw, float* dbias,
                        float* ws, int B, int M, int HW, void* stream) {
  return pdes::wgrad_tc_impl(g, x, x_ld, c_off, Cn, nullptr, 0, dW, ldw, dbias, ws, B, M, HW, stream);
}

size_t pdes_conv3x3_tc_pack_floats(int Cin, int N) {
  if (Cin <= 0 || N <= 0 || N > pdes::kTcMaxN) return 0;
  return (size_t)pdes::tc_nchunks(Cin) * 9 * 2 * pdes::tc_block_floats(N);
}

int pdes_conv3x3_tc_ok(int B, int Cin, int N, int H, int W, const float* x) {
  if (B <= 0 || Cin <= 0 || N <= 0 || N > pdes::kTcMaxN || N % 4 != 0) return 0;
  if (H < 10 || W < 20 || W % 4 != 0) return 0;                     // 3-D TMA box 10 x 20, 16-byte global strides
  if ((reinterpret_cast<uintptr_t>(x) & 15u) != 0) return 0;
  return 1;
}

/* out[B][N][H-2][W-2] = valid 3x3 cross-correlation of x[B][Cin][H][W] with w[N][Cin][3][3] (+ bias), 3xTF32 on
 * tcgen05.  wpack needs pdes_conv3x3_tc_pack_floats() floats and is rewritten on every call (weights change every
 * optimizer step).  Returns PDES_ERR_UNSUPPORTED when pdes_conv3x3_tc_ok() is false. */
int pdes_conv3x3_tc(const float* x, const float* w, const float* bias, float* wpack, float* out, int B, int Cin, int N,
                    int H, int W, void* stream) {
  using namespace pdes;
#ifdef PDES_CPU_EMU
  (void)x; (void)w; (void)bias; (void)wpack; (void)out; (void)B; (void)Cin; (void)N; (void)H; (void)W; (void)stream;
  set_error("pdes_conv3x3_tc: tcgen05 path is not available in the CPU emulation build");
  return PDES_ERR_UNSUPPORTED;
#else
  PDES_REQUIRE(x && w && wpack && out, PDES_ERR_ARG, "pdes_conv3x3_tc: null pointer");
  PDES_REQUIRE(pdes_conv3x3_tc_ok(B, Cin, N, H, W, x) && g_encode_tiled() != nullptr, PDES_ERR_UNSUPPORTED,
               "pdes_conv3x3_tc: shape/alignment not supported (Cin=%d N=%d H=%d W=%d)", Cin, N, H, W);
  if (g_num_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (g_num_sms <= 0) g_num_sms = 148;
  }
  ConvParams p;
  p.wpack = wpack; p.bias = bias; p.out = out; p.B = B; p.Cin = Cin; p.N = N; p.npad = tc_npad(N); p.H = H; p.W = W;
  p.Ho = H - 2; p.Wo = W - 2;
  p.tiles_y = ceil_div(p.Ho, kCvTH); p.tiles_x = ceil_div(p.Wo, kCvTW);
  const int ncc = tc_nchunks(Cin);
  {
    const int total = ncc * 9 * p.npad * kTcBK;
    auto pk = k_pack_conv3x3_tf32;
    PDES_LAUNCH(pk, dim3((unsigned)ceil_div(total, 256)), dim3(256), 0, stream, w, N, Cin, p.npad, ncc, wpack);
    if (int e = check_launch("pdes_conv3x3_tc(pack)")) return e;
  }
  alignas(64) CUtensorMap tm;
  memset(&tm, 0, sizeof(tm));
  const cuuint64_t gdim[3] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B * (cuuint64_t)Cin};
  const cuuint64_t gstr[2] = {(cuuint64_t)W * 4, (cuuint64_t)H * W * 4};
  const cuuint32_t box[3] = {(cuuint32_t)kCvBW, (cuuint32_t)kCvBH, (cuuint32_t)kTcBK};
  const cuuint32_t estr[3] = {1, 1, 1};
  const CUresult r = g_encode_tiled()(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(x), gdim, gstr, box, estr,
                                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                      CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  PDES_REQUIRE(r == CUDA_SUCCESS, PDES_ERR_UNSUPPORTED, "pdes_conv3x3_tc: cuTensorMapEncodeTiled failed (%d)", (int)r);
  const size_t smem = (size_t)kV3ASt * 2 * kTcM * kTcBK * 4 + (size_t)kV3BSt * 2 * p.npad * kTcBK * 4 +
                      (size_t)kV3Raw * kCvRawF * 4 + 1024;
  const int ntiles = B * p.tiles_y * p.tiles_x;
  auto kfn = k_conv3x3_tc;
  PDES_SET_SMEM(kfn, smem);
  PDES_LAUNCH(kfn, dim3((unsigned)(ntiles < g_num_sms ? ntiles : g_num_sms)), dim3(kV3Threads), smem, stream, p, tm);
  return check_launch("pdes_conv3x3_tc");
#endif
}

size_t pdes_gemm_tc_pack_floats(int K, int N) {
  if (!pdes_gemm_tc_supported(N, K)) return 0;
  return (size_t)pdes::tc_nchunks(K) * 2 * pdes::tc_block_floats(N);
}

static int gemm_tc_pack_impl(const float* Wt, size_t sk, size_t sn, int K, int N, float* packed, void* stream) {
  using namespace pdes;
  PDES_REQUIRE(pdes_gemm_tc_supported(N, K), PDES_ERR_UNSUPPORTED, "pdes_gemm_tc_pack: N=%d > %d", N, kTcMaxN);
  const int npad = tc_npad(N), nchunks = tc_nchunks(K);
  const int total = nchunks * npad * kTcBK;
  auto kfn = k_pack_b_tf32;
  PDES_LAUNCH(kfn, dim3((unsigned)ceil_div(total, 256)), dim3(256), 0, stream, Wt, sk, sn, K, N, npad, nchunks, packed);
  return check_launch("pdes_gemm_tc_pack");
}

int pdes_gemm_tc_pack(const float* Wt, int lda, int K, int N, float* packed, void* stream) {
  using namespace pdes;
  PDES_REQUIRE(Wt && packed && K > 0 && N > 0 && lda >= N, PDES_ERR_ARG, "pdes_gemm_tc_pack: bad arguments");
  return gemm_tc_pack_impl(Wt, (size_t)lda, 1, K, N, packed, stream);
}

/* Same packed operand from the transposed storage W[n][k] (row stride ldw >= K): this is how an nn.Conv2d(k=1) weight
 * [Cout][Cin] feeds the forward GEMM (K = Cin, N = Cout) without a separate transpose pass. */
int pdes_gemm_tc_pack_t(const float* W, int ldw, int K, int N, float* packed, void* stream) {
  using namespace pdes;
  PDES_REQUIRE(W && packed && K > 0 && N > 0 && ldw >= K, PDES_ERR_ARG, "pdes_gemm_tc_pack_t: bad arguments");
  return gemm_tc_pack_impl(W, 1, (size_t)ldw, K, N, packed, stream);
}

}  // extern "C"

namespace pdes {
namespace {
int inv_w_gemm_tc_impl(const float* Z, const float* wpack, const float* x0, int C0, const float* x1, int C1,
                       const float* bias, const float* res, const float* tables, int backward_scale, float* out,
                       float* pre, size_t out_bs, int B, int N, int H, int W, int m1, int m2, int act, void* stream) {
#ifdef PDES_CPU_EMU
  (void)out_bs;
  (void)Z; (void)wpack; (void)x0; (void)C0; (void)x1; (void)C1; (void)bias; (void)res; (void)tables;
  (void)backward_scale; (void)out; (void)pre; (void)B; (void)N; (void)H; (void)W; (void)m1; (void)m2; (void)act;
  (void)stream;
  set_error("pdes_inv_w_gemm_tc: tcgen05 path is not available in the CPU emulation build");
  return PDES_ERR_UNSUPPORTED;
#else
  PDES_REQUIRE(wpack && x0 && out, PDES_ERR_ARG, "pdes_inv_w_gemm_tc: null pointer");
  PDES_REQUIRE(B > 0 && N > 0 && H > 0 && W > 0 && C0 > 0 && C1 >= 0, PDES_ERR_ARG, "pdes_inv_w_gemm_tc: bad sizes");
  PDES_REQUIRE((C1 == 0) == (x1 == nullptr), PDES_ERR_ARG, "pdes_inv_w_gemm_tc: x1/C1 mismatch");
  PDES_REQUIRE(pdes_inv_w_gemm_tc_ok(N, C0 + C1, H, W, Z != nullptr ? m2 : 0, x0, x1), PDES_ERR_UNSUPPORTED,
               "pdes_inv_w_gemm_tc: shape/alignment not supported by the tensor-core kernel (N=%d, HW=%d, W=%d)", N, H * W, W);
  PDES_REQUIRE(B <= 65535, PDES_ERR_UNSUPPORTED, "pdes_inv_w_gemm_tc: grid too large");
#ifndef PDES_TC_TRACE
  PDES_REQUIRE(act == PDES_ACT_NONE || act == PDES_ACT_GELU, PDES_ERR_ARG, "pdes_inv_w_gemm_tc: unknown activation");
#endif
  PDES_REQUIRE((reinterpret_cast<uintptr_t>(wpack) & 15u) == 0, PDES_ERR_ARG, "pdes_inv_w_gemm_tc: wpack must be 16-byte aligned");
  TcParams p;
  p.Z = Z; p.wpack = wpack; p.x0 = x0; p.C0 = C0; p.x1 = x1; p.C1 = C1; p.bias = bias; p.res = res; p.T = nullptr;
  if (Z != nullptr) {
    PDES_REQUIRE(tables != nullptr && m1 > 0 && m2 > 0 && m2 <= W / 2 + 1, PDES_ERR_ARG, "pdes_inv_w_gemm_tc: spectral term needs tables/modes");
    const TableLayout t = table_layout(H, W, m1, m2);
    p.T = tables + (backward_scale ? t.tinv_b : t.tinv_f);
  }
  p.out = out; p.pre = pre; p.N = N; p.npad = tc_npad(N); p.K = C0 + C1; p.H = H; p.W = W; p.m2 = m2; p.act = act;
  p.single_pass = (g_tc_mode == 3) ? 1 : 0;
  p.out_bs = out_bs;
  PDES_REQUIRE(out_bs >= (size_t)N * H * W, PDES_ERR_ARG, "pdes_inv_w_gemm_tc: output batch stride smaller than N*H*W");
  const int v3_rows = (kTcM + W - 1) / W + 1;
  const bool v3_ta = tc_npad(N) <= kV3TaMaxN;        // A operand through tensor memory (no shared-memory A stages)
  const int v3_nst = (v3_ta && tc_npad(N) <= 192) ? 4 : 3;
  const size_t v3_slot = (size_t)kTcBK * (tc_npad(N) > kTcM ? tc_npad(N) : kTcM) * 4;           // bytes per raw slot
  const size_t v3_fixed = (size_t)(v3_ta ? 0 : v3_nst) * 2 * kTcM * kTcBK * 4 + (size_t)v3_nst * 2 * tc_npad(N) * kTcBK * 4 +
                          (size_t)(Z != nullptr ? tc_nchunks(v3_rows * 2 * m2) : 0) * kTcBK * kTcM * 4 + 2048;
  const bool v3_fits = v3_fixed + 3 * v3_slot <= 227 * 1024 && (unsigned long long)B * out_bs < (1ull << 32);   // 32-bit epilogue offsets
  int v3_nraw = v3_fits ? (int)((227 * 1024 - v3_fixed) / v3_slot) : 0;
  if (v3_nraw > kV3MaxRaw) v3_nraw = kV3MaxRaw;
  if (g_tc_mode >= 2 && (Z == nullptr || N % 4 == 0) && v3_fits) {
    if (g_num_sms == 0) {
      int dev = 0;
      cudaGetDevice(&dev);
      cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
      if (g_num_sms <= 0) g_num_sms = 148;
    }
    const int rawld = p.npad > kTcM ? p.npad : kTcM;
    const int rows_max = (kTcM + W - 1) / W + 1;
    const int nsp_max = (Z != nullptr) ? tc_nchunks(rows_max * 2 * m2) : 0;
    const size_t smem3 = (size_t)(v3_ta ? 0 : v3_nst) * 2 * kTcM * kTcBK * 4 + (size_t)v3_nst * 2 * p.npad * kTcBK * 4 +
                         (size_t)v3_nraw * kTcBK * rawld * 4 + (size_t)nsp_max * kTcBK * kTcM * 4 + 1024;
    const int tiles_per_img = ceil_div(H * W, kTcM);
    const int ntiles = B * tiles_per_img;
    // 2-D tensor map over x0 viewed as [B*C0 rows][HW pixels]; used for the chunks that lie entirely inside x0 when
    // every tile is a full 128-pixel box (otherwise the kernel falls back to per-row bulk copies)
    alignas(64) CUtensorMap tmap;
    memset(&tmap, 0, sizeof(tmap));
    int ntmap_chunks = 0;
    // (a partial last tile of an image is fine: the box columns beyond H*W are out of bounds and arrive as zeros)
    if ((H * W) % 4 == 0 && g_encode_tiled() != nullptr) {
      const cuuint64_t gdim[2] = {(cuuint64_t)(H * W), (cuuint64_t)B * (cuuint64_t)C0};
      const cuuint64_t gstr[1] = {(cuuint64_t)(H * W) * 4};
      const cuuint32_t box[2] = {(cuuint32_t)kTcM, (cuuint32_t)kTcBK};
      const cuuint32_t estr[2] = {1, 1};
      const CUresult r = g_encode_tiled()(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(x0), gdim, gstr, box,
                                          estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                          CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r == CUDA_SUCCESS) ntmap_chunks = C0 / kTcBK;
    }
#ifdef PDES_TC_TRACE
    int v3_nst_used = v3_nst, v3_nraw_used = v3_nraw;
    if (const char* e = getenv("PDES_V3_NST")) v3_nst_used = atoi(e);
    if (const char* e = getenv("PDES_V3_NRAW")) v3_nraw_used = atoi(e);
#define v3_nst v3_nst_used
#define v3_nraw v3_nraw_used
#endif
    auto k3 = v3_ta ? k_inv_w_gemm_tc_v3<true> : k_inv_w_gemm_tc_v3<false>;
    PDES_SET_SMEM(k3, smem3);
    // the last, partly filled round of tiles is cut into 2 or 4 column ranges when that gives every tile piece its own CTA
    const int grid3 = ntiles < g_num_sms ? ntiles : g_num_sms;
    const int nfull = (ntiles / grid3) * grid3, ntail = ntiles - nfull;
    int tail_shift = 0;
    if (ntail > 0 && nfull > 0 && getenv("PDES_K3B_NO_TAIL_SPLIT") == nullptr) {
      if (ntail * 4 <= grid3 && p.npad % 64 == 0) tail_shift = 2;
      else if (ntail * 2 <= grid3 && p.npad % 32 == 0) tail_shift = 1;
    }
    PDES_LAUNCH_PDL(k3, dim3((unsigned)grid3), dim3(kK3Threads), smem3, stream, p, B, tiles_per_img, tmap, ntmap_chunks,
                    v3_nst, v3_nraw, nfull, tail_shift);
#ifdef PDES_TC_TRACE
#undef v3_nst
#undef v3_nraw
#endif
    return check_launch("pdes_inv_w_gemm_tc(v3)");
  }
  PDES_REQUIRE(out_bs == (size_t)N * H * W, PDES_ERR_UNSUPPORTED, "pdes_inv_w_gemm_tc: strided output needs tensor-core mode >= 2");
  const size_t stage = (size_t)2 * kTcM * kTcBK * 4 + (size_t)2 * p.npad * kTcBK * 4;
  const size_t smem = 2 * stage + (size_t)kTcRaw * kTcBK * kTcM * 4 + 1024;
  auto kfn = k_inv_w_gemm_tc;
  PDES_SET_SMEM(kfn, smem);
  PDES_LAUNCH(kfn, dim3((unsigned)ceil_div(H * W, kTcM), (unsigned)B), dim3(kTcThreads), smem, stream, p);
  return check_launch("pdes_inv_w_gemm_tc");
#endif
}
}  // namespace
}  // namespace pdes

extern "C" {

int pdes_inv_w_gemm_tc(const float* Z, const float* wpack, const float* x0, int C0, const float* x1, int C1,
                       const float* bias, const float* res, const float* tables, int backward_scale, float* out,
                       float* pre, int B, int N, int H, int W, int m1, int m2, int act, void* stream) {
  return pdes::inv_w_gemm_tc_impl(Z, wpack, x0, C0, x1, C1, bias, res, tables, backward_scale, out, pre,
                                  (size_t)N * H * W, B, N, H, W, m1, m2, act, stream);
}

int pdes_conv1x1_tc_ok(int B, int Cin, int N, int HW, const float* x) {
  if (B <= 0 || B > 65535 || Cin <= 0 || HW <= 0 || pdes_get_tensor_core_mode() < 2) return 0;
  return pdes_inv_w_gemm_tc_ok(N, Cin, 1, HW, 0, x, nullptr);
}

/* 1x1 convolution (U-Net shortcut / projection convs) on the K3b tensor-core pipeline:
 *   out[b, n, p] = act(sum_k Wt[k][n] * x[b, k, p] + bias[n] + res[b, n, p]),   x [B][Cin][HW] contiguous,
 * out / res with batch stride `out_bstride` floats (>= N*HW; lets the caller write a channel sub-range of a wider
 * tensor, e.g. the two halves of a 385-channel input gradient).  wpack from pdes_gemm_tc_pack(Wt, lda, Cin, N). */
int pdes_conv1x1_tc(const float* x, int Cin, const float* wpack, const float* bias, const float* res, float* out,
                    size_t out_bstride, int B, int N, int HW, int act, void* stream) {
  using namespace pdes;
  PDES_REQUIRE(HW > 0 && pdes_conv1x1_tc_ok(B, Cin, N, HW, x), PDES_ERR_UNSUPPORTED,
               "pdes_conv1x1_tc: shape/alignment/mode not supported (Cin=%d N=%d HW=%d)", Cin, N, HW);
  return inv_w_gemm_tc_impl(nullptr, wpack, x, Cin, nullptr, 0, bias, res, nullptr, 0, out, nullptr, out_bstride, B, N, 1, HW,
                            0, 0, act, stream);
}

}  // extern "C"
