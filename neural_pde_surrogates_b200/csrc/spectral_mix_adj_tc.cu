// Adjoint of K2 with respect to the spectrum (dX) on tcgen05, from the SAME packed weights the forward kernel streams.
//
//   reference: autograd of compl_mul2d = einsum("bixy,ioxy->boxy"), proc_fno.py:253-255:
//              GX[b,i,m] = sum_o conj(W[i,o,m]) GO[b,o,m]
//
// The forward pack Wp[m][tile][chunk][row o][16 i][re|im] (spectral_mix_tc.cu) holds, for one mode and one tile of
// output channels, rows of 128 bytes = 16 input channels x (re, im).  Read as an MN-major tf32 A operand (pdes_ptx.cuh:
// smem_desc_mn_sw128b32; rows = K index o, the 32 floats of a row = 32 M rows (i, re|im)) it is exactly W^T, so the adjoint
// needs no second 59 MB copy of the weights and no per-step re-pack:
//     D[(i, re)][n] = sum_o Wr[i,o] G[o][n],   D[(i, im)][n] = sum_o Wi[i,o] G[o][n],   n = (b, re|im),  G = GO de-interleaved
//     GXr[b,i] = D[(i,re)][(b,re)] + D[(i,im)][(b,im)],   GXi[b,i] = D[(i,re)][(b,im)] - D[(i,im)][(b,re)]
// M = 128 = 4 consecutive chunks (64 input channels), N = 2B (padded to 32), K = all output channels (both tiles).
// 3xTF32: the raw TMA tile is the hi half (kind::tf32 truncates), its residual goes through tensor memory, G is split into
// hi / lo K-major blocks in shared memory once per mode.  Rows the reference overwrites when 2*m1 > H are zero in Wp, which
// is also their adjoint.  Output in the O2 layout [2][mode][B][C0] (partial 1 = 0) that K3a (k_inv_h2) reads.
//
// Measured on B200 (B = 16, 193 -> 192 channels, 200 modes): 36.9 us against 41 us for the FFMA/TMA adjoint plus the saving
// of the new K3a on its output (block backward 321 -> 308 us).  It is tensor-pipe bound: an MN-major tf32 MMA occupies the
// pipe ~95 cycles whatever N is (N = 2B = 32 here), 72 of them per item, and 600 items on 148 SMs run as 5 per CTA on 120
// CTAs.  A K-major operand would be ~5x cheaper per MMA but needs a transposed second pack (59 MB, rebuilt every step).
//
// Work item = (mode, group of 64 input channels) = 2 stages (one per output-channel tile); items are cut into contiguous
// ranges, one per SM.  320 threads: warp 0 TMA (4 boxes of `to` rows per stage, 128B_ATOM_32B swizzle, 3-stage ring),
// warp 1 MMA issue, warps 2-5 G blocks + residual -> TMEM, warps 6-9 epilogue.
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

constexpr int kAdjStages = 3, kAdjThreads = 320, kAdjN = 32, kAdjMaxTo = 96, kAdjBK = 16;

struct AdjParams {
  const float2* GO2;        // [nmodes][B][CoutP] complex, mode-major gradient spectrum (K1 on g)
  float2* O2;               // [2][nmodes][B][C0] complex
  int B, Cout, CoutP, C0, nmodes, ntile, to, nck, ngroups, nitems;
};

struct AdjBars {
  unsigned long long full[kAdjStages], empty[kAdjStages], lo_full[2], lo_empty[2], b_full, b_empty, d_full[2], d_empty[2];
};

__global__ void __launch_bounds__(kAdjThreads, 1)
k_mix_adj_tc(AdjParams p, const __grid_constant__ CUtensorMap tmap_w) {
  PDES_DYN_SMEM(unsigned char, smem_raw);
  unsigned char* base = ptx::align_smem_1024(smem_raw);
  const uint32_t blk_bytes = (uint32_t)p.to * 128u;               // one chunk: `to` rows of 128 bytes
  const uint32_t stage_bytes = 4 * blk_bytes;                     // 4 chunks = 64 input channels
  const int KT = p.ntile * p.to;                                  // K extent: padded output channels
  unsigned char* stages = base;
  float* Ghi = reinterpret_cast<float*>(stages + kAdjStages * stage_bytes);      // canonical K-major [KT][32]
  float* Glo = Ghi + (size_t)KT * kAdjN;
  __shared__ __align__(8) AdjBars bars;
  __shared__ uint32_t tmem_slot;

  const int tid = threadIdx.x, warp = ptx::uniform_warp_idx(), lane = tid & 31;
  const int per = (p.nitems + (int)gridDim.x - 1) / (int)gridDim.x;
  const int it_beg = (int)blockIdx.x * per, it_end = (it_beg + per < p.nitems) ? it_beg + per : p.nitems;

  if (tid == 0) {
    for (int i = 0; i < kAdjStages; ++i) { ptx::mbar_init(&bars.full[i], 1); ptx::mbar_init(&bars.empty[i], 1); }
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&bars.lo_full[i], 4);
      ptx::mbar_init(&bars.lo_empty[i], 1);
      ptx::mbar_init(&bars.d_full[i], 1);
      ptx::mbar_init(&bars.d_empty[i], 4);
    }
    ptx::mbar_init(&bars.b_full, 4);
    ptx::mbar_init(&bars.b_empty, 1);
    ptx::fence_mbar_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc(&tmem_slot, 256);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  const uint32_t d_col = 0, lo_col = 64;                          // D buffers at 0 / 32, residual buffers at 64 / 64 + to
  const uint32_t lbo_g = (kAdjN / 8) * 128;                       // 512 bytes between k-groups of 4 in the G blocks

  if (warp == 0) {
    // ================================================================== TMA producer (the weights are static: no dependency wait)
    const uint64_t pol = ptx::l2_policy_evict_first();
    uint32_t s = 0, eph = 1, n = 0;
    for (int it = it_beg; it < it_end; ++it) {
      const int m = it / p.ngroups, ig = it - m * p.ngroups;
      for (int t = 0; t < p.ntile; ++t, ++n) {
        if (n >= (uint32_t)kAdjStages) ptx::mbar_wait(&bars.empty[s], eph);
        ptx::mbar_arrive_expect_tx_ws(&bars.full[s], stage_bytes);
        const int row0 = ((m * p.ntile + t) * p.nck + 4 * ig) * p.to;           // 4 consecutive chunks are contiguous rows
        for (int j = 0; j < 4; ++j)
          ptx::tma_load_2d_ws_hint(stages + s * stage_bytes + (uint32_t)j * blk_bytes, &tmap_w, 0, row0 + j * p.to, &bars.full[s], pol);
        if (++s == (uint32_t)kAdjStages) { s = 0; eph ^= 1u; }
      }
    }
  } else if (warp == 1) {
    // ================================================================== MMA issue
    const uint32_t idesc_mn = ptx::idesc_tf32_a_mn(128, kAdjN), idesc_k = ptx::idesc_tf32(128, kAdjN);
    const uint64_t ghi0 = ptx::smem_desc_noswizzle(ptx::smem_u32(Ghi), lbo_g, 128);
    const uint64_t glo0 = ptx::smem_desc_noswizzle(ptx::smem_u32(Glo), lbo_g, 128);
    const uint64_t a0 = ptx::smem_desc_mn_sw128b32(ptx::smem_u32(stages), blk_bytes, 512);
    const int nks = p.to / 8;
    uint32_t s = 0, fph = 0, n = 0, ni = 0, nmode = 0;
    int cur_m = -1;
    for (int it = it_beg; it < it_end; ++it, ++ni) {
      const int m = it / p.ngroups;
      const uint32_t acc = ni & 1, dph = (ni >> 1) & 1;
      if (m != cur_m) {                                           // G blocks of a new mode
        ptx::mbar_wait(&bars.b_full, nmode & 1);
        cur_m = m;
        ++nmode;
      }
      if (ni >= 2) ptx::mbar_wait(&bars.d_empty[acc], dph ^ 1u);
      const uint32_t dcol = tmem_base + d_col + acc * kAdjN;
      for (int t = 0; t < p.ntile; ++t, ++n) {
        const uint32_t a = n & 1, lph = (n >> 1) & 1;
        ptx::mbar_wait(&bars.full[s], fph);
        ptx::mbar_wait(&bars.lo_full[a], lph);
        ptx::tc_fence_after();
        const uint32_t lcol = tmem_base + lo_col + a * (uint32_t)p.to;
        const uint64_t as = a0 + (uint64_t)((s * stage_bytes) >> 4);
#pragma unroll 4
        for (int ks = 0; ks < nks; ++ks) {
          const uint64_t ad = as + (uint64_t)((ks * 1024) >> 4);
          const uint64_t kb = (uint64_t)((((uint32_t)(t * p.to) / 4 + ks * 2) * lbo_g) >> 4);
          ptx::mma_tf32_ws(dcol, ad, glo0 + kb, idesc_mn, (t | ks) != 0 ? 1u : 0u);   // trunc(W) * G_lo
          ptx::mma_tf32_ws(dcol, ad, ghi0 + kb, idesc_mn, 1u);                        // trunc(W) * G_hi
        }
#pragma unroll 4
        for (int ks = 0; ks < nks; ++ks)                                               // (W - trunc(W)) * G_hi
          ptx::mma_tf32_ta_ws(dcol, lcol + ks * 8, ghi0 + (uint64_t)((((uint32_t)(t * p.to) / 4 + ks * 2) * lbo_g) >> 4), idesc_k, 1u);
        ptx::tc_commit_ws(&bars.empty[s]);
        ptx::tc_commit_ws(&bars.lo_empty[a]);
        if (++s == (uint32_t)kAdjStages) { s = 0; fph ^= 1u; }
      }
      ptx::tc_commit_ws(&bars.d_full[acc]);
      const bool last_of_mode = (it + 1 == it_end) || ((it + 1) / p.ngroups != m);
      if (last_of_mode) ptx::tc_commit_ws(&bars.b_empty);         // the G blocks may be rebuilt for the next mode
    }
  } else if (warp < 6) {
    // ================================================================== G blocks (once per mode) and residual -> tensor memory
    const int q = warp & 3, ml = q * 32 + lane;                   // TMEM lane = (chunk j, i2, re|im)
    const int ct = tid - 64;                                      // 0..127 within the group
    const uint32_t col_off = (uint32_t)(ml >> 5) * blk_bytes + (uint32_t)(ml & 7) * 4u;
    const uint32_t chunk = (uint32_t)((ml & 31) >> 3);
    PDES_GRID_DEP_WAIT();                                         // GO2 is written by the previous kernel (K1 on the gradient)
    uint32_t s = 0, fph = 0, n = 0, nmode = 0;
    int cur_m = -1;
    for (int it = it_beg; it < it_end; ++it) {
      const int m = it / p.ngroups;
      if (m != cur_m) {
        if (nmode >= 1) ptx::mbar_wait(&bars.b_empty, (nmode - 1) & 1);
        // G[k = o][n = (b, re|im)] from GO2[m][b][o]: one item = (b, 4 consecutive o) = 32 bytes in, 2 x (hi, lo) 16-byte rows out
        const int nq = KT / 4;
        for (int idx = ct; idx < kAdjN / 2 * nq; idx += 128) {
          const int b = idx / nq, kq = idx - b * nq, o = kq * 4;
          float4 v0 = make_float4(0.f, 0.f, 0.f, 0.f), v1 = v0;
          if (b < p.B) {
            // k = t * to + row is the output channel itself (tile t covers channels [t * to, (t + 1) * to)); KT == CoutP
            const float4* src = reinterpret_cast<const float4*>(p.GO2 + ((size_t)m * p.B + b) * p.CoutP + o);
            v0 = __ldg(src);                         // (pad columns o >= Cout are inside the allocation and masked below)
            v1 = __ldg(src + 1);
          }
          float re[4] = {v0.x, v0.z, v1.x, v1.z}, im[4] = {v0.y, v0.w, v1.y, v1.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            if (o + e >= p.Cout) { re[e] = 0.f; im[e] = 0.f; }
          }
          float4 rh, rl, ih, il;
          rh.x = __uint_as_float(__float_as_uint(re[0]) & 0xffffe000u); rl.x = re[0] - rh.x;
          rh.y = __uint_as_float(__float_as_uint(re[1]) & 0xffffe000u); rl.y = re[1] - rh.y;
          rh.z = __uint_as_float(__float_as_uint(re[2]) & 0xffffe000u); rl.z = re[2] - rh.z;
          rh.w = __uint_as_float(__float_as_uint(re[3]) & 0xffffe000u); rl.w = re[3] - rh.w;
          ih.x = __uint_as_float(__float_as_uint(im[0]) & 0xffffe000u); il.x = im[0] - ih.x;
          ih.y = __uint_as_float(__float_as_uint(im[1]) & 0xffffe000u); il.y = im[1] - ih.y;
          ih.z = __uint_as_float(__float_as_uint(im[2]) & 0xffffe000u); il.z = im[2] - ih.z;
          ih.w = __uint_as_float(__float_as_uint(im[3]) & 0xffffe000u); il.w = im[3] - ih.w;
          const int n0 = 2 * b;
          const uint32_t off_re = (uint32_t)kq * 128u + (uint32_t)(n0 >> 3) * 32u + (uint32_t)(n0 & 7) * 4u;   // floats
          const uint32_t off_im = off_re + 4u;
          *reinterpret_cast<float4*>(Ghi + off_re) = rh;
          *reinterpret_cast<float4*>(Glo + off_re) = rl;
          *reinterpret_cast<float4*>(Ghi + off_im) = ih;
          *reinterpret_cast<float4*>(Glo + off_im) = il;
        }
        ptx::fence_proxy_async();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&bars.b_full);
        cur_m = m;
        ++nmode;
      }
      for (int t = 0; t < p.ntile; ++t, ++n) {
        const uint32_t a = n & 1, lph = (n >> 1) & 1;
        ptx::mbar_wait(&bars.full[s], fph);
        if (n >= 2) ptx::mbar_wait(&bars.lo_empty[a], lph ^ 1u);
        ptx::tc_fence_after();
        const unsigned char* colp = stages + s * stage_bytes + col_off;
        const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16) + lo_col + a * (uint32_t)p.to;
        for (int r0 = 0; r0 < p.to; r0 += 16) {
          uint32_t lo[16];
#pragma unroll
          for (int e = 0; e < 16; ++e) {
            const int r = r0 + e;
            float v = 0.0f;
            if (r < p.to) v = *reinterpret_cast<const float*>(colp + (uint32_t)r * 128u + ((chunk ^ (uint32_t)(r & 3)) << 5));
            lo[e] = __float_as_uint(v - __uint_as_float(__float_as_uint(v) & 0xffffe000u));
          }
          ptx::tmem_st16(trow + (uint32_t)r0, lo);
        }
        ptx::tmem_st_wait();
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&bars.lo_full[a]);
        if (++s == (uint32_t)kAdjStages) { s = 0; fph ^= 1u; }
      }
    }
  } else {
    // ================================================================== epilogue: lanes (i, re) / (i, im) are neighbours
    const int q = warp & 3, ml = q * 32 + lane;
    const int j = ml >> 5, i2 = (ml & 31) >> 1, odd = ml & 1;
    const size_t pstride = (size_t)p.nmodes * p.B * p.C0;
    uint32_t ni = 0;
    for (int it = it_beg; it < it_end; ++it, ++ni) {
      const int m = it / p.ngroups, ig = it - m * p.ngroups;
      const uint32_t acc = ni & 1, dph = (ni >> 1) & 1;
      ptx::mbar_wait(&bars.d_full[acc], dph);
      ptx::tc_fence_after();
      uint32_t r[32];
      ptx::tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + d_col + acc * kAdjN, r);
      ptx::tmem_ld_wait();
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&bars.d_empty[acc]);
      const int i = ig * 64 + j * 16 + i2;
      float2* out = p.O2 + ((size_t)m * p.B) * p.C0 + i;
#pragma unroll
      for (int b = 0; b < kAdjN / 2; ++b) {
        const float mine_r = __uint_as_float(r[2 * b]), mine_i = __uint_as_float(r[2 * b + 1]);     // this lane's row x (G re, G im)
        const float oth_r = __shfl_xor_sync(0xffffffffu, mine_r, 1), oth_i = __shfl_xor_sync(0xffffffffu, mine_i, 1);
        if (b < p.B && i < p.C0) {
          if (!odd) {                                // even lane = Wr row: GX = (Wr Gr + Wi Gi, Wr Gi - Wi Gr)
            out[(size_t)b * p.C0] = make_float2(mine_r + oth_i, mine_i - oth_r);
          } else {                                   // odd lane keeps partial 1 of the O2 layout at zero
            out[pstride + (size_t)b * p.C0] = make_float2(0.f, 0.f);
          }
        }
      }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, 256);
  }
}

typedef CUresult (*AdjEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int g_adj_sms = 0;

}  // namespace
#endif  // !PDES_CPU_EMU

}  // namespace pdes

extern "C" {

/* 1 when pdes_mix_tc_dx can run this shape: the forward pack geometry of (Cin, Cout) with tiles of <= 96 rows, the
 * C0 channels that need a gradient a multiple of 64, B <= 16. */
int pdes_mix_tc_dx_ok(int B, int Cin, int Cout, int C0, int m1, int m2) {
#ifdef PDES_CPU_EMU
  (void)B; (void)Cin; (void)Cout; (void)C0; (void)m1; (void)m2;
  return 0;
#else
  if (!pdes_mix_tc_ok(B, Cin, Cout, m1, m2)) return 0;
  if (B > pdes::kAdjN / 2 || C0 <= 0 || C0 > Cin || C0 % 64 != 0) return 0;
  const int to = pdes::mt_to(Cout), ntile = pdes::mt_ntile(Cout);
  if (to > pdes::kAdjMaxTo || to % 8 != 0 || ntile * to != pdes::mt_cinp(Cout)) return 0;   // K extent == padded row count of GO2
  if (getenv("PDES_NO_DX_TC") != nullptr) return 0;
  return 1;
#endif
}

/* GX = adjoint of the per-mode channel mix w.r.t. the spectrum, for the first C0 input channels.
 * GO2 [2 m1 m2][B][Cout_pad] complex (pdes_dft_fwd2 on the output gradient, Hermitian weights on), Wp = the FORWARD pack of
 * pdes_mix_tc_pack(Cin, Cout), O2 [2][2 m1 m2][B][C0] complex (same layout as pdes_mix_tc_fwd's output; read by
 * pdes_inv_h_modes). */
int pdes_mix_tc_dx(const float* GO2, const float* Wp, float* O2, int B, int Cin, int Cout, int C0, int m1, int m2, void* stream) {
  using namespace pdes;
#ifdef PDES_CPU_EMU
  (void)GO2; (void)Wp; (void)O2; (void)B; (void)Cin; (void)Cout; (void)C0; (void)m1; (void)m2; (void)stream;
  set_error("pdes_mix_tc_dx: tcgen05 path is not available in the CPU emulation build");
  return PDES_ERR_UNSUPPORTED;
#else
  PDES_REQUIRE(GO2 && Wp && O2, PDES_ERR_ARG, "pdes_mix_tc_dx: null pointer");
  PDES_REQUIRE(pdes_mix_tc_dx_ok(B, Cin, Cout, C0, m1, m2), PDES_ERR_UNSUPPORTED, "pdes_mix_tc_dx: shape not supported");
  PDES_REQUIRE(aligned16(GO2) && aligned16(Wp) && aligned16(O2), PDES_ERR_ARG, "pdes_mix_tc_dx: pointers must be 16-byte aligned");
  AdjParams p;
  p.GO2 = reinterpret_cast<const float2*>(GO2);
  p.O2 = reinterpret_cast<float2*>(O2);
  p.B = B; p.Cout = Cout; p.CoutP = mt_cinp(Cout); p.C0 = C0; p.nmodes = 2 * m1 * m2;
  p.ntile = mt_ntile(Cout); p.to = mt_to(Cout); p.nck = mt_cinp(Cin) / kAdjBK; p.ngroups = C0 / 64;
  p.nitems = p.nmodes * p.ngroups;
  AdjEncodeFn enc = reinterpret_cast<AdjEncodeFn>(tensor_map_encoder());
  alignas(64) CUtensorMap tw;
  memset(&tw, 0, sizeof(tw));
  const cuuint64_t gdim[2] = {32, (cuuint64_t)p.nmodes * p.ntile * p.nck * p.to};
  const cuuint64_t gstr[1] = {128};
  const cuuint32_t box[2] = {32, (cuuint32_t)p.to};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = enc(&tw, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(Wp), gdim, gstr, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  PDES_REQUIRE(r == CUDA_SUCCESS, PDES_ERR_UNSUPPORTED, "pdes_mix_tc_dx: cuTensorMapEncodeTiled(Wp) failed (%d)", (int)r);
  const size_t smem = (size_t)kAdjStages * 4 * p.to * 128 + (size_t)2 * p.ntile * p.to * kAdjN * 4 + 1024;
  PDES_REQUIRE(smem <= 227 * 1024, PDES_ERR_UNSUPPORTED, "pdes_mix_tc_dx: not enough shared memory");
  if (g_adj_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_adj_sms, cudaDevAttrMultiProcessorCount, dev);
    if (g_adj_sms <= 0) g_adj_sms = 148;
  }
  const int grid = p.nitems < g_adj_sms ? p.nitems : g_adj_sms;
  auto kfn = k_mix_adj_tc;
  PDES_SET_SMEM(kfn, smem);
  PDES_LAUNCH_PDL(kfn, dim3((unsigned)grid), dim3(kAdjThreads), smem, stream, p, tw);
  return check_launch("pdes_mix_tc_dx");
#endif
}

}  // extern "C"
