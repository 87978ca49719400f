// K2 (forward mix and its dX adjoint), TMA-fed version for sm_100a.
//
//   P[ks][b][c][m] = sum_{r in split ks} Xin[b][r][m] * (CONJ ? conj : id)(W[r,c,m])
//   (einsum("bixy,ioxy->boxy"), reference proc_fno.py:253-255,266-269, weights in the parameter layout
//   [Cin][Cout][m1][m2] complex, mode index fastest, proc_fno.py:240-243)
//
// The op is a pure weight stream (59 MB per launch at the twophase config, B flop per weight byte), so the kernel is
// built around keeping HBM busy rather than around the FFMAs:
//   * one CTA = (mode tile of <= 120 modes of ONE weight block) x (TCc output channels) x (BT samples) x (one split of
//     the reduction); per reduction step ONE elected thread issues two 3-D TMA boxes (SASS UTMALDG) into a ring of NST
//     stages:   W box  [TCc channels][MT modes]   (a contiguous TCc*MM*8-byte run of the parameter for the forward mix)
//               X box  [BT samples][MT modes]
//     ~100-150 KB are in flight per SM, independent of occupancy and registers;
//   * compute threads = (mode, channel group, sample group): CPT x BPT complex accumulators each, operands read with
//     conflict-free LDS.64 (lanes = consecutive modes), 4*CPT*BPT FFMA per CPT + BPT shared loads;
//   * X is read once per TCc channels (through L2), each weight byte once.
// The same kernel serves the dX adjoint: only the tensor-map box changes ([1 o][TCc i] instead of [TCc o][1 i]).
// Not compiled for the CPU emulation build (the generic kernels in spectral_mix.cu remain the fallback there and for
// shapes the tensor maps cannot describe: odd m1*m2 or unaligned pointers).
#include "pdes_common.cuh"
#include "pdes_ptx.cuh"
#ifndef PDES_CPU_EMU
#include <cuda.h>
#include <cstring>
#endif

namespace pdes {

#ifdef PDES_CPU_EMU
int mix_tma_splits(int, int, int, int, int) { return 0; }
int mix_tma_launch(bool, const float*, const float*, const float*, float*, int, int, int, int, int, int, int, int, int,
                   void*) {
  return PDES_ERR_UNSUPPORTED;
}
int mix_dw_tma_launch(const float*, const float*, float*, float*, int, int, int, int, int, int, void*) { return PDES_ERR_UNSUPPORTED; }
#else

namespace {

constexpr int kMixTmaMaxStages = 8;
constexpr int kMixTmaMaxMT = 120;      // modes per tile: 4 thread groups x 120 + the producer warp = 512 threads (128 registers)
constexpr size_t kMixTmaSmem = 200 * 1024;

struct MixTmaCfg {
  int BT, TCc;      // samples / channels per CTA
};
inline MixTmaCfg mix_tma_cfg(int B) {
  if (B > 8) return {16, 8};
  if (B > 4) return {8, 16};
  return {4, 16};
}
struct MixTmaGeom {
  int ntm, MT;      // mode tiles per weight block, modes per tile (even)
};
inline MixTmaGeom mix_tma_geom(int MM) {
  MixTmaGeom g;
  g.ntm = ceil_div(MM, kMixTmaMaxMT);
  g.MT = (ceil_div(MM, g.ntm) + 1) & ~1;
  return g;
}

struct MixTmaParams {
  float2* P;
  int B, Cred, Cn, MM, m1, m2, H, MT, ntm, red_per_split, nsplit, nstages, conj_box;
};

// MTC: modes per tile as a compile-time constant (100 = the shipped 10x10 modes: shared-memory offsets become
// immediates), 0 = run-time value.
template <int CPT, int BPT, int GC, int GB, bool CONJ, int MTC>
__global__ void __launch_bounds__(kMixTmaMaxMT * GC * GB + 32, 1)
k_mix_tma(const __grid_constant__ CUtensorMap tm_w1, const __grid_constant__ CUtensorMap tm_w2,
          const __grid_constant__ CUtensorMap tm_x, MixTmaParams p) {
  constexpr int TCc = CPT * GC, BT = BPT * GB;
  PDES_DYN_SMEM(unsigned char, smem_raw);
  // (offset arithmetic on the shared pointer, not an integer round trip: keeps the loads LDS instead of generic LD)
  unsigned char* base = smem_raw + ((128u - (ptx::smem_u32(smem_raw) & 127u)) & 127u);
  __shared__ __align__(8) unsigned long long full[kMixTmaMaxStages], empty[kMixTmaMaxStages];

  const int MT = MTC ? MTC : p.MT, MM = p.MM, M2 = 2 * MM;
  const int ncomp = MT * GC * GB;                           // compute threads; the warp after them is the producer
  const int ncomp_warps = (ncomp + 31) / 32;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int blk = blockIdx.x / p.ntm;                       // weight block (0: rows kx < m1, 1: the negative-frequency rows)
  const int m0 = (blockIdx.x - blk * p.ntm) * MT;           // first mode of the tile inside its block
  const int c0 = blockIdx.y * TCc;
  const int ks = blockIdx.z % p.nsplit;
  const int b0 = (blockIdx.z / p.nsplit) * BT;
  const int r0 = ks * p.red_per_split;
  const int r1 = (r0 + p.red_per_split < p.Cred) ? (r0 + p.red_per_split) : p.Cred;
  const int NST = p.nstages;
  const uint32_t w_bytes = (uint32_t)TCc * MT * 8, x_bytes = (uint32_t)BT * MT * 8;
  const uint32_t stage_bytes = (w_bytes + x_bytes + 127) & ~127u;

  if (tid == 0) {
    for (int s = 0; s < NST; ++s) {
      ptx::mbar_init(&full[s], 1);
      ptx::mbar_init(&empty[s], ncomp_warps);
    }
    ptx::fence_mbar_init();
  }
  __syncthreads();

  if (warp == ncomp_warps) {
    if (lane == 0) {
      // ---------------------------------------------------------------- producer: two TMA boxes per reduction step
      const CUtensorMap* tw = blk ? &tm_w2 : &tm_w1;
      int s = 0;
      uint32_t ph = 1;                                               // first pass over the ring: slots are free
      for (int r = r0; r < r1; ++r) {
        if (r - r0 >= NST) ptx::mbar_wait(&empty[s], ph);
        unsigned char* st = base + (size_t)s * stage_bytes;
        ptx::mbar_arrive_expect_tx(&full[s], w_bytes + x_bytes);
        // W viewed as [Cin][Cout][2*MM floats]: forward reduces over Cin (box [1][TCc][2MT]), dX over Cout (box [TCc][1][2MT])
        if (CONJ)
          ptx::tma_load_3d(st, tw, 2 * m0, r, c0, &full[s]);
        else
          ptx::tma_load_3d(st, tw, 2 * m0, c0, r, &full[s]);
        ptx::tma_load_3d(st + w_bytes, &tm_x, 2 * (blk * MM + m0), r, b0, &full[s]);
        if (++s == NST) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp < ncomp_warps) {
    // ------------------------------------------------------------------ consumers
    const bool active = tid < ncomp;
    const int m = active ? tid % MT : 0;
    const int grp = active ? tid / MT : 0;
    const int gc = grp % GC, gb = grp / GC;
    float2 acc[CPT][BPT];
#pragma unroll
    for (int t = 0; t < CPT; ++t)
#pragma unroll
      for (int bb = 0; bb < BPT; ++bb) acc[t][bb] = make_float2(0.0f, 0.0f);
    const uint32_t w_off = (uint32_t)((gc * CPT) * MT + m) * 8, x_off = w_bytes + (uint32_t)((gb * BPT) * MT + m) * 8;
    int s = 0;
    uint32_t ph = 0;
    for (int r = r0; r < r1; ++r) {
      ptx::mbar_wait(&full[s], ph);
      const unsigned char* st = base + (size_t)s * stage_bytes;
      const float2* ws = reinterpret_cast<const float2*>(st + w_off);
      const float2* xs = reinterpret_cast<const float2*>(st + x_off);
      float2 wv[CPT], wq[CPT];                                       // w and i*w: x*w = x.x * w + x.y * (i*w)
#pragma unroll
      for (int t = 0; t < CPT; ++t) {
        wv[t] = ws[t * MT];
        if (CONJ) wv[t].y = -wv[t].y;
        wq[t] = make_float2(-wv[t].y, wv[t].x);
      }
#pragma unroll
      for (int bb = 0; bb < BPT; ++bb) {
        const float2 xv = xs[bb * MT];
        const float2 xx = make_float2(xv.x, xv.x), yy = make_float2(xv.y, xv.y);   // scalar-broadcast operands (free)
#pragma unroll
        for (int t = 0; t < CPT; ++t) {
          ffma2(acc[t][bb], xx, wv[t]);
          ffma2(acc[t][bb], yy, wq[t]);
        }
      }
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&empty[s]);
      if (++s == NST) { s = 0; ph ^= 1; }
    }
    if (active && m0 + m < MM) {
      const int mg = blk * MM + m0 + m;                     // mode index in [0, 2*MM)
      const bool dead = row_dead(mg / p.m2, p.m1, p.H);
#pragma unroll
      for (int bb = 0; bb < BPT; ++bb) {
        const int b = b0 + gb * BPT + bb;
        if (b >= p.B) continue;
#pragma unroll
        for (int t = 0; t < CPT; ++t) {
          const int c = c0 + gc * CPT + t;
          if (c < p.Cn)
            p.P[(((size_t)ks * p.B + b) * p.Cn + c) * M2 + mg] = dead ? make_float2(0.f, 0.f) : acc[t][bb];
        }
      }
    }
  }
}

// Weight gradient  GW[i][o][m] = sum_b conj(X[b][i][m]) * GO[b][o][m]  (written in the parameter layout), same structure:
// one CTA = (100 modes of one weight block) x TI input channels x TO output channels, the reduction runs over the
// samples: per sample two TMA boxes (X [TI][MT], GO [TO][MT]) land in the ring, thread = (mode, IPT x OPT channel
// tile) keeps IPT*OPT complex accumulators and does two FFMA2 per complex MAC (conj(x) g = g.re (x.re, -x.im) +
// g.im (x.im, x.re)).
struct MixDwParams {
  float2* gw1;
  float2* gw2;
  int B, Cin, Cout, MM, m1, m2, H, MT, ntm, nstages;
};

template <int IPT, int OPT, int GI, int GO_, int MTC>
__global__ void __launch_bounds__(kMixTmaMaxMT * GI * GO_ + 32, 1)
k_mix_dw_tma(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_go, MixDwParams p) {
  constexpr int TI = IPT * GI, TO = OPT * GO_;
  PDES_DYN_SMEM(unsigned char, smem_raw);
  unsigned char* base = smem_raw + ((128u - (ptx::smem_u32(smem_raw) & 127u)) & 127u);
  __shared__ __align__(8) unsigned long long full[kMixTmaMaxStages], empty[kMixTmaMaxStages];

  const int MT = MTC ? MTC : p.MT, MM = p.MM;
  const int ncomp = MT * GI * GO_;
  const int ncomp_warps = (ncomp + 31) / 32;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int blk = blockIdx.x / p.ntm;
  const int m0 = (blockIdx.x - blk * p.ntm) * MT;
  const int o0 = blockIdx.y * TO, i0 = blockIdx.z * TI;
  const int NST = p.nstages;
  const uint32_t x_bytes = (uint32_t)TI * MT * 8, g_bytes = (uint32_t)TO * MT * 8;
  const uint32_t stage_bytes = (x_bytes + g_bytes + 127) & ~127u;

  if (tid == 0) {
    for (int s = 0; s < NST; ++s) {
      ptx::mbar_init(&full[s], 1);
      ptx::mbar_init(&empty[s], ncomp_warps);
    }
    ptx::fence_mbar_init();
  }
  __syncthreads();

  if (warp == ncomp_warps) {
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 1;
      for (int b = 0; b < p.B; ++b) {
        if (b >= NST) ptx::mbar_wait(&empty[s], ph);
        unsigned char* st = base + (size_t)s * stage_bytes;
        ptx::mbar_arrive_expect_tx(&full[s], x_bytes + g_bytes);
        ptx::tma_load_3d(st, &tm_x, 2 * (blk * MM + m0), i0, b, &full[s]);
        ptx::tma_load_3d(st + x_bytes, &tm_go, 2 * (blk * MM + m0), o0, b, &full[s]);
        if (++s == NST) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp < ncomp_warps) {
    const bool active = tid < ncomp;
    const int m = active ? tid % MT : 0;
    const int grp = active ? tid / MT : 0;
    const int gi = grp % GI, go = grp / GI;
    float2 acc[IPT][OPT];
#pragma unroll
    for (int t = 0; t < IPT; ++t)
#pragma unroll
      for (int u = 0; u < OPT; ++u) acc[t][u] = make_float2(0.0f, 0.0f);
    const uint32_t x_off = (uint32_t)((gi * IPT) * MT + m) * 8, g_off = x_bytes + (uint32_t)((go * OPT) * MT + m) * 8;
    int s = 0;
    uint32_t ph = 0;
    for (int b = 0; b < p.B; ++b) {
      ptx::mbar_wait(&full[s], ph);
      const unsigned char* st = base + (size_t)s * stage_bytes;
      const float2* xs = reinterpret_cast<const float2*>(st + x_off);
      const float2* gs = reinterpret_cast<const float2*>(st + g_off);
      float2 xc[IPT], xw[IPT];
#pragma unroll
      for (int t = 0; t < IPT; ++t) {
        const float2 x = xs[t * MT];
        xc[t] = make_float2(x.x, -x.y);
        xw[t] = make_float2(x.y, x.x);
      }
#pragma unroll
      for (int u = 0; u < OPT; ++u) {
        const float2 g = gs[u * MT];
        const float2 gx = make_float2(g.x, g.x), gy = make_float2(g.y, g.y);
#pragma unroll
        for (int t = 0; t < IPT; ++t) {
          ffma2(acc[t][u], gx, xc[t]);
          ffma2(acc[t][u], gy, xw[t]);
        }
      }
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&empty[s]);
      if (++s == NST) { s = 0; ph ^= 1; }
    }
    if (active && m0 + m < MM) {
      const bool dead = row_dead((blk * MM + m0 + m) / p.m2, p.m1, p.H);
      float2* gw = (blk ? p.gw2 : p.gw1) + m0 + m;
#pragma unroll
      for (int t = 0; t < IPT; ++t) {
        const int i = i0 + gi * IPT + t;
        if (i >= p.Cin) continue;
#pragma unroll
        for (int u = 0; u < OPT; ++u) {
          const int o = o0 + go * OPT + u;
          if (o < p.Cout) gw[((size_t)i * p.Cout + o) * MM] = dead ? make_float2(0.f, 0.f) : acc[t][u];
        }
      }
    }
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

bool encode3(EncodeTiledFn enc, CUtensorMap* tm, const float* base, cuuint64_t d0, cuuint64_t d1, cuuint64_t d2, cuuint64_t s1_bytes,
             cuuint64_t s2_bytes, cuuint32_t b0, cuuint32_t b1, cuuint32_t b2) {
  memset(tm, 0, sizeof(*tm));
  const cuuint64_t gdim[3] = {d0, d1, d2};
  const cuuint64_t gstr[2] = {s1_bytes, s2_bytes};
  const cuuint32_t box[3] = {b0, b1, b2};
  const cuuint32_t estr[3] = {1, 1, 1};
  return enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

inline bool mix_tma_shape_ok(int B, int Cred, int Cn, int m1, int m2) {
  const long MM = (long)m1 * m2;
  return B > 0 && Cred > 0 && Cn > 0 && MM > 0 && MM % 2 == 0 && 2 * MM < (1L << 30);
}

}  // namespace

// number of reduction splits that fills the GPU once with this kernel's grid; 0 = shape not supported
int mix_tma_splits(int B, int Cred, int Cn, int m1, int m2) {
  if (!mix_tma_shape_ok(B, Cred, Cn, m1, m2)) return 0;
  const MixTmaCfg cfg = mix_tma_cfg(B);
  const MixTmaGeom g = mix_tma_geom(m1 * m2);
  const long base = 2L * g.ntm * ceil_div(Cn, cfg.TCc) * ceil_div(B, cfg.BT);
  long ns = 148 / base;                                   // one CTA per SM, a single wave
  const long cap = Cred / 8 > 0 ? Cred / 8 : 1;           // keep >= 8 reduction steps per split
  if (ns > cap) ns = cap;
  if (ns > 16) ns = 16;
  if (ns < 1) ns = 1;
  return (int)ns;
}

// conj = false: forward mix (reduce over Cin, W strides [r = i][c = o]); conj = true: dX adjoint (reduce over Cout).
// Cw_in / Cw_out: the physical weight tensor dims [Cw_in][Cw_out][MM].
int mix_tma_launch(bool conj, const float* Xin, const float* w1, const float* w2, float* P, int nsplit, int B, int Cred,
                   int Cn, int Cw_in, int Cw_out, int m1, int m2, int H, void* stream) {
  if (!mix_tma_shape_ok(B, Cred, Cn, m1, m2)) return PDES_ERR_UNSUPPORTED;
  if (!aligned16(Xin) || !aligned16(w1) || !aligned16(w2) || !aligned16(P)) return PDES_ERR_UNSUPPORTED;
  EncodeTiledFn enc = reinterpret_cast<EncodeTiledFn>(tensor_map_encoder());
  if (enc == nullptr) return PDES_ERR_UNSUPPORTED;
  const int MM = m1 * m2;
  const MixTmaCfg cfg = mix_tma_cfg(B);
  const MixTmaGeom g = mix_tma_geom(MM);
  const dim3 grid((unsigned)(2 * g.ntm), (unsigned)ceil_div(Cn, cfg.TCc), (unsigned)(nsplit * ceil_div(B, cfg.BT)));
  if (grid.y > 65535 || grid.z > 65535) return PDES_ERR_UNSUPPORTED;

  alignas(64) CUtensorMap tw1, tw2, tx;
  const cuuint64_t row = (cuuint64_t)MM * 8;
  const cuuint32_t bi = (cuuint32_t)(2 * g.MT);
  bool ok;
  if (conj) {   // box [TCc rows of Cw_in][1 row of Cw_out][2MT]
    ok = encode3(enc, &tw1, w1, 2 * (cuuint64_t)MM, (cuuint64_t)Cw_out, (cuuint64_t)Cw_in, row, row * Cw_out, bi, 1, (cuuint32_t)cfg.TCc) &&
         encode3(enc, &tw2, w2, 2 * (cuuint64_t)MM, (cuuint64_t)Cw_out, (cuuint64_t)Cw_in, row, row * Cw_out, bi, 1, (cuuint32_t)cfg.TCc);
  } else {      // box [1 row of Cw_in][TCc rows of Cw_out][2MT]
    ok = encode3(enc, &tw1, w1, 2 * (cuuint64_t)MM, (cuuint64_t)Cw_out, (cuuint64_t)Cw_in, row, row * Cw_out, bi, (cuuint32_t)cfg.TCc, 1) &&
         encode3(enc, &tw2, w2, 2 * (cuuint64_t)MM, (cuuint64_t)Cw_out, (cuuint64_t)Cw_in, row, row * Cw_out, bi, (cuuint32_t)cfg.TCc, 1);
  }
  // Xin [B][Cred][2*MM modes] complex
  ok = ok && encode3(enc, &tx, Xin, 4 * (cuuint64_t)MM, (cuuint64_t)Cred, (cuuint64_t)B, 2 * row, 2 * row * Cred, bi, 1, (cuuint32_t)cfg.BT);
  if (!ok) return PDES_ERR_UNSUPPORTED;

  MixTmaParams p;
  p.P = reinterpret_cast<float2*>(P);
  p.B = B; p.Cred = Cred; p.Cn = Cn; p.MM = MM; p.m1 = m1; p.m2 = m2; p.H = H; p.MT = g.MT; p.ntm = g.ntm;
  p.red_per_split = ceil_div(Cred, nsplit); p.nsplit = nsplit; p.conj_box = conj ? 1 : 0;
  const size_t stage = (((size_t)(cfg.TCc + cfg.BT) * g.MT * 8) + 127) & ~size_t(127);
  int nst = (int)((kMixTmaSmem - 256) / stage);
  if (nst > kMixTmaMaxStages) nst = kMixTmaMaxStages;
  if (nst < 2) return PDES_ERR_UNSUPPORTED;
  p.nstages = nst;
  const size_t smem = (size_t)nst * stage + 256;

#define PDES_MIXT_LAUNCH(CPT, BPT, GC, GB)                                                          \
  do {                                                                                              \
    const unsigned threads = (unsigned)(((g.MT * GC * GB + 31) / 32 + 1) * 32);                     \
    if (conj && g.MT == 100) {                                                                      \
      auto kfn = k_mix_tma<CPT, BPT, GC, GB, true, 100>;                                            \
      PDES_SET_SMEM(kfn, smem);                                                                     \
      PDES_LAUNCH(kfn, grid, dim3(threads), smem, stream, tw1, tw2, tx, p);                         \
    } else if (conj) {                                                                              \
      auto kfn = k_mix_tma<CPT, BPT, GC, GB, true, 0>;                                              \
      PDES_SET_SMEM(kfn, smem);                                                                     \
      PDES_LAUNCH(kfn, grid, dim3(threads), smem, stream, tw1, tw2, tx, p);                         \
    } else if (g.MT == 100) {                                                                       \
      auto kfn = k_mix_tma<CPT, BPT, GC, GB, false, 100>;                                           \
      PDES_SET_SMEM(kfn, smem);                                                                     \
      PDES_LAUNCH(kfn, grid, dim3(threads), smem, stream, tw1, tw2, tx, p);                         \
    } else {                                                                                        \
      auto kfn = k_mix_tma<CPT, BPT, GC, GB, false, 0>;                                             \
      PDES_SET_SMEM(kfn, smem);                                                                     \
      PDES_LAUNCH(kfn, grid, dim3(threads), smem, stream, tw1, tw2, tx, p);                         \
    }                                                                                               \
  } while (0)
  if (cfg.BT == 16) PDES_MIXT_LAUNCH(4, 8, 2, 2);
  else if (cfg.BT == 8) PDES_MIXT_LAUNCH(4, 8, 4, 1);
  else PDES_MIXT_LAUNCH(4, 4, 4, 1);
#undef PDES_MIXT_LAUNCH
  return check_launch(conj ? "pdes_mix_dx(tma)" : "pdes_mix_fwd(tma)");
}
// weight gradient of the mix; PDES_ERR_UNSUPPORTED = caller falls back to k_mix_dw
int mix_dw_tma_launch(const float* X, const float* GO, float* gw1, float* gw2, int B, int Cin, int Cout, int m1, int m2,
                      int H, void* stream) {
  if (!mix_tma_shape_ok(B, Cin, Cout, m1, m2) || B < 8) return PDES_ERR_UNSUPPORTED;    // short sample loops: old kernel
  if (!aligned16(X) || !aligned16(GO) || !aligned16(gw1) || !aligned16(gw2)) return PDES_ERR_UNSUPPORTED;
  EncodeTiledFn enc = reinterpret_cast<EncodeTiledFn>(tensor_map_encoder());
  if (enc == nullptr) return PDES_ERR_UNSUPPORTED;
  constexpr int IPT = 4, OPT = 8, GI = 2, GO_ = 2, TI = IPT * GI, TO = OPT * GO_;
  const int MM = m1 * m2;
  const MixTmaGeom g = mix_tma_geom(MM);
  const dim3 grid((unsigned)(2 * g.ntm), (unsigned)ceil_div(Cout, TO), (unsigned)ceil_div(Cin, TI));
  if (grid.y > 65535 || grid.z > 65535) return PDES_ERR_UNSUPPORTED;
  alignas(64) CUtensorMap tx, tg;
  const cuuint64_t row = (cuuint64_t)MM * 16;                       // bytes of one [2*MM modes] complex row
  const cuuint32_t bi = (cuuint32_t)(2 * g.MT);
  if (!encode3(enc, &tx, X, 4 * (cuuint64_t)MM, (cuuint64_t)Cin, (cuuint64_t)B, row, row * Cin, bi, (cuuint32_t)TI, 1) ||
      !encode3(enc, &tg, GO, 4 * (cuuint64_t)MM, (cuuint64_t)Cout, (cuuint64_t)B, row, row * Cout, bi, (cuuint32_t)TO, 1))
    return PDES_ERR_UNSUPPORTED;
  MixDwParams p;
  p.gw1 = reinterpret_cast<float2*>(gw1); p.gw2 = reinterpret_cast<float2*>(gw2);
  p.B = B; p.Cin = Cin; p.Cout = Cout; p.MM = MM; p.m1 = m1; p.m2 = m2; p.H = H; p.MT = g.MT; p.ntm = g.ntm;
  const size_t stage = (((size_t)(TI + TO) * g.MT * 8) + 127) & ~size_t(127);
  int nst = (int)((kMixTmaSmem - 256) / stage);
  if (nst > kMixTmaMaxStages) nst = kMixTmaMaxStages;
  if (nst > B) nst = B;
  if (nst < 2) return PDES_ERR_UNSUPPORTED;
  p.nstages = nst;
  const size_t smem = (size_t)nst * stage + 256;
  const unsigned threads = (unsigned)(((g.MT * GI * GO_ + 31) / 32 + 1) * 32);
  if (g.MT == 100) {
    auto kfn = k_mix_dw_tma<IPT, OPT, GI, GO_, 100>;
    PDES_SET_SMEM(kfn, smem);
    PDES_LAUNCH(kfn, grid, dim3(threads), smem, stream, tx, tg, p);
  } else {
    auto kfn = k_mix_dw_tma<IPT, OPT, GI, GO_, 0>;
    PDES_SET_SMEM(kfn, smem);
    PDES_LAUNCH(kfn, grid, dim3(threads), smem, stream, tx, tg, p);
  }
  return check_launch("pdes_mix_dw(tma)");
}
#endif  // PDES_CPU_EMU

}  // namespace pdes
