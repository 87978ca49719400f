// K2: per-mode complex channel mixing and its two adjoints.
//
// Replaces einsum("bixy,ioxy->boxy") x2 (reference proc_fno.py:253-255,266-269).  The weights stay in the
// reference's parameter layout [Cin][Cout][m1][m2] complex (mode index fastest, proc_fno.py:240-243): a lane owns ONE
// mode, a warp 32 consecutive modes, so every weight load is a contiguous 256-byte segment and each weight byte is
// read once.  No packed copy, no permute, no bmm.
//
// Version 2 (round 1, after profiling): the first version let every channel-group CTA re-read its X slice from L2
// (240 MB of L2->SM traffic at B=16 for a 59 MB weight stream).  Now a CTA owns a block of 32 modes and a range of
// output channels, stages its X[r][b][32 modes] slice in shared memory ONCE, and its warps sweep the channel groups:
// per reduction step a warp issues TC coalesced weight loads (the HBM stream) and BT conflict-free LDS.64 for
// 4*TC*BT FFMA.  Plain FFMA on purpose: the op is weight-bandwidth bound (B flop per weight byte).
#include "pdes_common.cuh"

namespace pdes {
namespace {

constexpr int kTC = 4;            // channels per warp pass
constexpr int kMixWarps = 8;      // warps per CTA

struct cplx { float x, y; };

// P[ks][b][c][m] = sum_{r in split ks} Xin[b][r][m] * (CONJ ? conj : id)(W[r,c,m])
template <int BT, int TC, bool CONJ>
__global__ void __launch_bounds__(32 * kMixWarps)
k_mix(const float2* __restrict__ Xin, const float2* __restrict__ W1, const float2* __restrict__ W2,
      float2* __restrict__ P, int B, int Cred, int Cn, int MM, int m1, int m2, int H, int wr_stride,
      int wo_stride, int red_per_split, int nsplit, int chan_per_cta) {
  PDES_DYN_SMEM(float2, xs);                                  // [nr][BT][32]
  const int M2 = 2 * MM;
  const int lane = threadIdx.x, warp = threadIdx.y;
  const int m = blockIdx.x * 32 + lane;
  const int ks = blockIdx.z % nsplit;
  const int b0 = (blockIdx.z / nsplit) * BT;
  const int r0 = ks * red_per_split;
  const int r1 = (r0 + red_per_split < Cred) ? (r0 + red_per_split) : Cred;
  const int nr = r1 - r0;
  const int nb = (B - b0 < BT) ? (B - b0) : BT;
  const bool mvalid = m < M2;

  // stage X[r0..r1)[b0..b0+BT)[32 modes]: 256-byte coalesced rows
  for (int row = warp; row < nr * BT; row += kMixWarps) {
    const int r = row / BT, bb = row - r * BT;
    xs[row * 32 + lane] = (mvalid && bb < nb) ? __ldg(Xin + ((size_t)(b0 + bb) * Cred + r0 + r) * M2 + m)
                                              : make_float2(0.f, 0.f);
  }
  __syncthreads();

  const bool second = m >= MM;
  const int mm = second ? m - MM : m;
  const float2* wbase = (second ? W2 : W1) + mm + (size_t)r0 * wr_stride;
  const bool dead = row_dead(m / m2, m1, H);
  const int cbeg = blockIdx.y * chan_per_cta;
  const int cend = (cbeg + chan_per_cta < Cn) ? (cbeg + chan_per_cta) : Cn;

  for (int c0 = cbeg + warp * TC; c0 < cend; c0 += kMixWarps * TC) {
    const int nc = (cend - c0 < TC) ? (cend - c0) : TC;
    cplx acc[TC][BT];
#pragma unroll
    for (int t = 0; t < TC; ++t)
#pragma unroll
      for (int bb = 0; bb < BT; ++bb) acc[t][bb].x = acc[t][bb].y = 0.0f;
    if (mvalid) {
      const float2* wr = wbase + (size_t)c0 * wo_stride;
#pragma unroll 4
      for (int r = 0; r < nr; ++r) {
        float2 wv[TC];
#pragma unroll
        for (int t = 0; t < TC; ++t) wv[t] = (t < nc) ? __ldg(wr + t * wo_stride) : make_float2(0.f, 0.f);
        const float2* xr = xs + (size_t)r * BT * 32 + lane;
#pragma unroll
        for (int bb = 0; bb < BT; ++bb) {
          const float2 xv = xr[bb * 32];
#pragma unroll
          for (int t = 0; t < TC; ++t) {
            const float wx = wv[t].x, wy = CONJ ? -wv[t].y : wv[t].y;
            acc[t][bb].x = fmaf(xv.x, wx, fmaf(-xv.y, wy, acc[t][bb].x));
            acc[t][bb].y = fmaf(xv.x, wy, fmaf(xv.y, wx, acc[t][bb].y));
          }
        }
        wr += wr_stride;
      }
#pragma unroll
      for (int t = 0; t < TC; ++t)
#pragma unroll
        for (int bb = 0; bb < BT; ++bb)
          if (t < nc && bb < nb)
            P[(((size_t)ks * B + b0 + bb) * Cn + c0 + t) * M2 + m] =
                dead ? make_float2(0.f, 0.f) : make_float2(acc[t][bb].x, acc[t][bb].y);
    }
  }
}

constexpr int kMixMaxX = 256;    // threads along the mode axis per CTA


struct MixGeom { int bx, by; };
inline MixGeom mix_geom(int M2) {
  MixGeom g;
  g.bx = ((M2 + 31) / 32) * 32;
  if (g.bx > kMixMaxX) g.bx = kMixMaxX;
  g.by = kMixMaxX / g.bx;
  if (g.by < 1) g.by = 1;
  return g;
}

// Streaming variant (first version): thread = mode, a CTA spans ALL modes of TC channels, so per reduction step it
// reads one contiguous TC*m1*m2 run of each weight tensor (best DRAM locality); X comes from L1/L2.  Measured best for
// the forward mix (78 vs 98 us at B=16), while the shared-memory variant below wins for the dX adjoint (73 vs 123 us).
// P[ks][b][c][m] = sum_{r in split ks} Xin[b][r][m] * (CONJ ? conj : id)(W[r,c,m])
template <int BT, int TC, bool CONJ>
__global__ void __launch_bounds__(kMixMaxX)
k_mix_stream(const float2* __restrict__ Xin, const float2* __restrict__ W1, const float2* __restrict__ W2,
      float2* __restrict__ P, int B, int Cred, int Cn, int MM, int m1, int m2, int H, int wr_stride,
      int wo_stride, int red_per_split, int nsplit) {
  const int M2 = 2 * MM;
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  const int c0 = (blockIdx.y * blockDim.y + threadIdx.y) * TC;
  const int ks = blockIdx.z % nsplit;
  const int b0 = (blockIdx.z / nsplit) * BT;
  if (m >= M2 || c0 >= Cn) return;
  const bool second = m >= MM;
  const int mm = second ? m - MM : m;
  const int r0 = ks * red_per_split;
  const int r1 = (r0 + red_per_split < Cred) ? (r0 + red_per_split) : Cred;
  const int nb = (B - b0 < BT) ? (B - b0) : BT;
  const int nc = (Cn - c0 < TC) ? (Cn - c0) : TC;
  const int xs = Cred * M2;                                  // batch stride of Xin (elements)

  cplx acc[TC][BT];
#pragma unroll
  for (int t = 0; t < TC; ++t)
#pragma unroll
    for (int bb = 0; bb < BT; ++bb) acc[t][bb].x = acc[t][bb].y = 0.0f;

  const float2* xr = Xin + ((size_t)b0 * Cred + r0) * M2 + m;
  const float2* wr = (second ? W2 : W1) + mm + (size_t)r0 * wr_stride + (size_t)c0 * wo_stride;
#pragma unroll 2
  for (int r = r0; r < r1; ++r) {
    float2 xv[BT], wv[TC];
#pragma unroll
    for (int bb = 0; bb < BT; ++bb) xv[bb] = (bb < nb) ? __ldg(xr + bb * xs) : make_float2(0.f, 0.f);
#pragma unroll
    for (int t = 0; t < TC; ++t) wv[t] = (t < nc) ? __ldg(wr + t * wo_stride) : make_float2(0.f, 0.f);
#pragma unroll
    for (int t = 0; t < TC; ++t) {
      const float wx = wv[t].x, wy = CONJ ? -wv[t].y : wv[t].y;
#pragma unroll
      for (int bb = 0; bb < BT; ++bb) {
        acc[t][bb].x = fmaf(xv[bb].x, wx, fmaf(-xv[bb].y, wy, acc[t][bb].x));
        acc[t][bb].y = fmaf(xv[bb].x, wy, fmaf(xv[bb].y, wx, acc[t][bb].y));
      }
    }
    xr += M2;
    wr += wr_stride;
  }
  const bool dead = row_dead(m / m2, m1, H);
#pragma unroll
  for (int t = 0; t < TC; ++t)
#pragma unroll
    for (int bb = 0; bb < BT; ++bb)
      if (t < nc && bb < nb)
        P[(((size_t)ks * B + b0 + bb) * Cn + c0 + t) * M2 + m] =
            dead ? make_float2(0.f, 0.f) : make_float2(acc[t][bb].x, acc[t][bb].y);
}

// GW[i][o][m] = sum_b conj(X[b][i][m]) * GO[b][o][m]   (written in the parameter layout)
// CTA = (32 modes, IR input channels, OR output channels): both slices staged in shared memory once, warps sweep the
// 4x4 (i, o) sub-tiles.
constexpr int kDwIR = 16, kDwOR = 16, kDwT = 4;

__global__ void __launch_bounds__(32 * kMixWarps)
k_mix_dw(const float2* __restrict__ X, const float2* __restrict__ GO, float2* __restrict__ gW1,
         float2* __restrict__ gW2, int B, int Cin, int Cout, int MM, int m1, int m2, int H) {
  PDES_DYN_SMEM(float2, sm);
  float2* xs = sm;                                           // [B][kDwIR][32]
  float2* gs = sm + (size_t)B * kDwIR * 32;                  // [B][kDwOR][32]
  const int M2 = 2 * MM;
  const int lane = threadIdx.x, warp = threadIdx.y;
  const int m = blockIdx.x * 32 + lane;
  const int i0 = blockIdx.z * kDwIR, o0 = blockIdx.y * kDwOR;
  const bool mvalid = m < M2;
  for (int row = warp; row < B * kDwIR; row += kMixWarps) {
    const int b = row / kDwIR, i = i0 + row % kDwIR;
    xs[row * 32 + lane] = (mvalid && i < Cin) ? __ldg(X + ((size_t)b * Cin + i) * M2 + m) : make_float2(0.f, 0.f);
  }
  for (int row = warp; row < B * kDwOR; row += kMixWarps) {
    const int b = row / kDwOR, o = o0 + row % kDwOR;
    gs[row * 32 + lane] = (mvalid && o < Cout) ? __ldg(GO + ((size_t)b * Cout + o) * M2 + m) : make_float2(0.f, 0.f);
  }
  __syncthreads();
  if (!mvalid) return;
  const bool dead = row_dead(m / m2, m1, H);
  const bool second = m >= MM;
  const int mm = second ? m - MM : m;
  float2* gW = (second ? gW2 : gW1) + mm;
  constexpr int NSI = kDwIR / kDwT, NSO = kDwOR / kDwT;
  for (int sub = warp; sub < NSI * NSO; sub += kMixWarps) {
    const int si = (sub / NSO) * kDwT, so = (sub % NSO) * kDwT;
    cplx acc[kDwT][kDwT];
#pragma unroll
    for (int t = 0; t < kDwT; ++t)
#pragma unroll
      for (int u = 0; u < kDwT; ++u) acc[t][u].x = acc[t][u].y = 0.0f;
#pragma unroll 2
    for (int b = 0; b < B; ++b) {
      float2 xi[kDwT], go[kDwT];
#pragma unroll
      for (int t = 0; t < kDwT; ++t) xi[t] = xs[((size_t)b * kDwIR + si + t) * 32 + lane];
#pragma unroll
      for (int u = 0; u < kDwT; ++u) go[u] = gs[((size_t)b * kDwOR + so + u) * 32 + lane];
#pragma unroll
      for (int t = 0; t < kDwT; ++t)
#pragma unroll
        for (int u = 0; u < kDwT; ++u) {
          acc[t][u].x = fmaf(xi[t].x, go[u].x, fmaf(xi[t].y, go[u].y, acc[t][u].x));
          acc[t][u].y = fmaf(xi[t].x, go[u].y, fmaf(-xi[t].y, go[u].x, acc[t][u].y));
        }
    }
#pragma unroll
    for (int t = 0; t < kDwT; ++t)
#pragma unroll
      for (int u = 0; u < kDwT; ++u) {
        const int i = i0 + si + t, o = o0 + so + u;
        if (i < Cin && o < Cout)
          gW[((size_t)i * Cout + o) * MM] = dead ? make_float2(0.f, 0.f) : make_float2(acc[t][u].x, acc[t][u].y);
      }
  }
}

inline int pick_bt(int B) { return B >= 8 ? 8 : (B >= 4 ? 4 : (B >= 2 ? 2 : 1)); }

struct MixPlan { int nsplit, red_per_split, nranges, chan_per_cta; };

// nsplit is given (the caller sized the partial-sum buffer); pick the channel ranges so that the grid has ~2 CTAs/SM
inline MixPlan mix_plan(int B, int Cred, int Cn, int M2, int nsplit) {
  MixPlan pl;
  pl.nsplit = nsplit;
  pl.red_per_split = ceil_div(Cred, nsplit);
  const long base = (long)ceil_div(M2, 32) * nsplit * ceil_div(B, pick_bt(B));
  const int groups = ceil_div(Cn, kTC);
  int nr = (int)((2L * 148 + base - 1) / base);
  const int max_nr = ceil_div(groups, kMixWarps);          // at least one pass of all warps per CTA
  if (nr > max_nr) nr = max_nr;
  if (nr < 1) nr = 1;
  pl.chan_per_cta = ceil_div(groups, nr) * kTC;
  pl.nranges = ceil_div(Cn, pl.chan_per_cta);
  return pl;
}

template <bool CONJ>
int launch_mix(const float* Xin, const float* w1, const float* w2, float* P, int nsplit, int B, int Cred, int Cn,
               int MM, int m1, int m2, int H, long wr_stride, long wo_stride, void* stream, const char* what) {
  const int M2 = 2 * MM;
  PDES_REQUIRE(wr_stride < (1L << 31) && wo_stride < (1L << 31) && (long)Cred * M2 < (1L << 31), PDES_ERR_UNSUPPORTED,
               "%s: tensor too large for 32-bit strides", what);
  const MixPlan pl = mix_plan(B, Cred, Cn, M2, nsplit);
  const int BT = pick_bt(B);
  const int nbt = ceil_div(B, BT);
  const size_t smem = (size_t)pl.red_per_split * BT * 32 * sizeof(float2);
  PDES_REQUIRE(smem <= (size_t)kMaxDynSmem, PDES_ERR_UNSUPPORTED, "%s: reduction split of %d channels needs %zu B of shared memory",
               what, pl.red_per_split, smem);
  const dim3 grid((unsigned)ceil_div(M2, 32), (unsigned)pl.nranges, (unsigned)(nsplit * nbt));
  PDES_REQUIRE(grid.y <= 65535 && grid.z <= 65535, PDES_ERR_UNSUPPORTED, "%s: grid too large", what);
  const dim3 block(32, kMixWarps);
  const float2* X2 = reinterpret_cast<const float2*>(Xin);
  const float2* A = reinterpret_cast<const float2*>(w1);
  const float2* Bw = reinterpret_cast<const float2*>(w2);
  float2* P2 = reinterpret_cast<float2*>(P);
#define PDES_MIX_CASE(bt)                                                                                     \
  case bt: {                                                                                                  \
    auto kfn = k_mix<bt, kTC, CONJ>;                                                                          \
    if (smem > 48 * 1024) PDES_SET_SMEM(kfn, smem);                                                           \
    PDES_LAUNCH(kfn, grid, block, smem, stream, X2, A, Bw, P2, B, Cred, Cn, MM, m1, m2, H, (int)wr_stride,    \
                (int)wo_stride, pl.red_per_split, nsplit, pl.chan_per_cta);                                   \
  } break;
  switch (BT) {
    PDES_MIX_CASE(8)
    PDES_MIX_CASE(4)
    PDES_MIX_CASE(2)
    default:
      PDES_MIX_CASE(1)
  }
#undef PDES_MIX_CASE
  return check_launch(what);
}


template <bool CONJ>
int launch_mix_stream(const float* Xin, const float* w1, const float* w2, float* P, int nsplit, int B, int Cred, int Cn,
               int MM, int m1, int m2, int H, long wr_stride, long wo_stride, void* stream, const char* what) {
  const int M2 = 2 * MM;
  PDES_REQUIRE(wr_stride < (1L << 31) && wo_stride < (1L << 31) && (long)Cred * M2 < (1L << 31), PDES_ERR_UNSUPPORTED,
               "%s: tensor too large for 32-bit strides", what);
  const int red_per_split = ceil_div(Cred, nsplit);
  const int BT = pick_bt(B);
  const int nbt = ceil_div(B, BT);
  const MixGeom g = mix_geom(M2);
  const dim3 grid((unsigned)ceil_div(M2, g.bx), (unsigned)ceil_div(Cn, kTC * g.by), (unsigned)(nsplit * nbt));
  PDES_REQUIRE(grid.y <= 65535 && grid.z <= 65535, PDES_ERR_UNSUPPORTED, "%s: grid too large", what);
  const dim3 block(g.bx, g.by);
  const float2* X2 = reinterpret_cast<const float2*>(Xin);
  const float2* A = reinterpret_cast<const float2*>(w1);
  const float2* Bw = reinterpret_cast<const float2*>(w2);
  float2* P2 = reinterpret_cast<float2*>(P);
#define PDES_MIXS_CASE(bt)                                                                                     \
  case bt: {                                                                                                  \
    auto kfn = k_mix_stream<bt, kTC, CONJ>;                                                                          \
    PDES_LAUNCH(kfn, grid, block, 0, stream, X2, A, Bw, P2, B, Cred, Cn, MM, m1, m2, H, (int)wr_stride,       \
                (int)wo_stride, red_per_split, nsplit);                                                       \
  } break;
  switch (BT) {
    PDES_MIXS_CASE(8)
    PDES_MIXS_CASE(4)
    PDES_MIXS_CASE(2)
    default:
      PDES_MIXS_CASE(1)
  }
#undef PDES_MIXS_CASE
  return check_launch(what);
}

int check_mix_args(const char* what, const void* a, const void* b, const void* c, const void* d, int B, int Cin,
                   int Cout, int H, int m1, int m2) {
  PDES_REQUIRE(a && b && c && d, PDES_ERR_ARG, "%s: null pointer", what);
  PDES_REQUIRE(B > 0 && Cin > 0 && Cout > 0 && H > 0, PDES_ERR_ARG, "%s: non-positive size", what);
  PDES_REQUIRE(m1 > 0 && m2 > 0 && m1 <= H, PDES_ERR_ARG, "%s: modes (%d,%d) out of range for H=%d", what, m1, m2, H);
  return PDES_OK;
}

}  // namespace
}  // namespace pdes

extern "C" {

int pdes_mix_suggest_splits(int B, int Cred, int Cout, int m1, int m2) {
  using namespace pdes;
  if (B <= 0 || Cred <= 0 || Cout <= 0 || m1 <= 0 || m2 <= 0) return 1;
  if (const int ns = mix_tma_splits(B, Cred, Cout, m1, m2)) return ns;
  const int M2 = 2 * m1 * m2;
  const MixGeom g = mix_geom(M2);
  const long blocks = (long)ceil_div(M2, g.bx) * ceil_div(Cout, kTC * g.by) * ceil_div(B, pick_bt(B));
  long ns = (3L * 148 + blocks - 1) / blocks;       // aim for >= 3 CTAs (of up to 256 threads) per SM
  const long cap = Cred / 8 > 0 ? Cred / 8 : 1;     // keep >= 8 reduction channels per split
  if (ns > cap) ns = cap;
  if (ns > 16) ns = 16;
  if (ns < 1) ns = 1;
  return (int)ns;
}

int pdes_mix_fwd(const float* X, const float* w1, const float* w2, float* P, int nsplit, int B, int Cin, int Cout,
                 int H, int m1, int m2, void* stream) {
  using namespace pdes;
  if (int e = check_mix_args("pdes_mix_fwd", X, w1, w2, P, B, Cin, Cout, H, m1, m2)) return e;
  PDES_REQUIRE(nsplit >= 1 && nsplit <= Cin, PDES_ERR_ARG, "pdes_mix_fwd: bad nsplit %d", nsplit);
  const int MM = m1 * m2;
  if (mix_tma_launch(false, X, w1, w2, P, nsplit, B, Cin, Cout, Cin, Cout, m1, m2, H, stream) == PDES_OK) return PDES_OK;
  return launch_mix_stream<false>(X, w1, w2, P, nsplit, B, Cin, Cout, MM, m1, m2, H, (long)Cout * MM, (long)MM, stream,
                                  "pdes_mix_fwd");
}

int pdes_mix_dx(const float* GO, const float* w1, const float* w2, float* P, int nsplit, int B, int Cin, int Cout,
                int Cgrad, int H, int m1, int m2, void* stream) {
  using namespace pdes;
  if (int e = check_mix_args("pdes_mix_dx", GO, w1, w2, P, B, Cin, Cout, H, m1, m2)) return e;
  PDES_REQUIRE(Cgrad > 0 && Cgrad <= Cin, PDES_ERR_ARG, "pdes_mix_dx: Cgrad %d not in (0,%d]", Cgrad, Cin);
  PDES_REQUIRE(nsplit >= 1 && nsplit <= Cout, PDES_ERR_ARG, "pdes_mix_dx: bad nsplit %d", nsplit);
  const int MM = m1 * m2;
  if (mix_tma_launch(true, GO, w1, w2, P, nsplit, B, Cout, Cgrad, Cin, Cout, m1, m2, H, stream) == PDES_OK) return PDES_OK;
  const size_t tile = (size_t)ceil_div(Cout, nsplit) * pick_bt(B) * 32 * sizeof(float2);
  if (tile > (size_t)kMaxDynSmem)      // X slice of one split does not fit in shared memory: streaming variant
    return launch_mix_stream<true>(GO, w1, w2, P, nsplit, B, Cout, Cgrad, MM, m1, m2, H, (long)MM, (long)Cout * MM, stream,
                                   "pdes_mix_dx");
  return launch_mix<true>(GO, w1, w2, P, nsplit, B, Cout, Cgrad, MM, m1, m2, H, (long)MM, (long)Cout * MM, stream,
                          "pdes_mix_dx");
}

int pdes_mix_dw(const float* X, const float* GO, float* gw1, float* gw2, int B, int Cin, int Cout, int H, int m1,
                int m2, void* stream) {
  using namespace pdes;
  if (int e = check_mix_args("pdes_mix_dw", X, GO, gw1, gw2, B, Cin, Cout, H, m1, m2)) return e;
  const int MM = m1 * m2;
  if (mix_dw_tma_launch(X, GO, gw1, gw2, B, Cin, Cout, m1, m2, H, stream) == PDES_OK) return PDES_OK;
  const size_t smem = (size_t)B * (kDwIR + kDwOR) * 32 * sizeof(float2);
  PDES_REQUIRE(smem <= (size_t)kMaxDynSmem, PDES_ERR_UNSUPPORTED, "pdes_mix_dw: batch %d needs %zu B of shared memory", B, smem);
  const dim3 grid((unsigned)ceil_div(2 * MM, 32), (unsigned)ceil_div(Cout, kDwOR), (unsigned)ceil_div(Cin, kDwIR));
  PDES_REQUIRE(grid.y <= 65535 && grid.z <= 65535, PDES_ERR_UNSUPPORTED, "pdes_mix_dw: grid too large");
  auto kfn = k_mix_dw;
  if (smem > 48 * 1024) PDES_SET_SMEM(kfn, smem);
  PDES_LAUNCH(kfn, grid, dim3(32, kMixWarps), smem, stream, reinterpret_cast<const float2*>(X),
              reinterpret_cast<const float2*>(GO), reinterpret_cast<float2*>(gw1), reinterpret_cast<float2*>(gw2), B,
              Cin, Cout, MM, m1, m2, H);
  return check_launch("pdes_mix_dw");
}

}  // extern "C"
