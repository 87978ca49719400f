// K2: per-mode complex channel mixing and its two adjoints.
//
// Replaces einsum("bixy,ioxy->boxy") x2 (reference proc_fno.py:253-255,266-269).  The weights stay in the
// reference's parameter layout [Cin][Cout][m1][m2] complex (mode index fastest, proc_fno.py:240-243):
// a thread owns ONE mode and a CTA spans all modes of a few channels, so for every reduction channel the CTA reads
// one contiguous run of TC*m1*m2 complex weights and each weight byte is read from HBM exactly once.  No packed
// copy, no permute, no bmm.  The layer is weight-bandwidth bound (B flop per weight byte), hence plain FFMA.
#include "pdes_common.cuh"

namespace pdes {
namespace {

constexpr int kMixMaxX = 256;    // threads along the mode axis per CTA
constexpr int kTC = 4;           // channels per thread

struct cplx { float x, y; };

struct MixGeom { int bx, by; };
inline MixGeom mix_geom(int M2) {
  MixGeom g;
  g.bx = ((M2 + 31) / 32) * 32;
  if (g.bx > kMixMaxX) g.bx = kMixMaxX;
  g.by = kMixMaxX / g.bx;
  if (g.by < 1) g.by = 1;
  return g;
}

// P[ks][b][c][m] = sum_{r in split ks} Xin[b][r][m] * (CONJ ? conj : id)(W[r,c,m])
template <int BT, int TC, bool CONJ>
__global__ void __launch_bounds__(kMixMaxX)
k_mix(const float2* __restrict__ Xin, const float2* __restrict__ W1, const float2* __restrict__ W2,
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
template <int TI, int TO>
__global__ void __launch_bounds__(kMixMaxX)
k_mix_dw(const float2* __restrict__ X, const float2* __restrict__ GO, float2* __restrict__ gW1,
         float2* __restrict__ gW2, int B, int Cin, int Cout, int MM, int m1, int m2, int H) {
  const int M2 = 2 * MM;
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  const int o0 = (blockIdx.y * blockDim.y + threadIdx.y) * TO;
  const int i0 = blockIdx.z * TI;
  if (m >= M2 || o0 >= Cout) return;
  const int ni = (Cin - i0 < TI) ? (Cin - i0) : TI;
  const int no = (Cout - o0 < TO) ? (Cout - o0) : TO;
  cplx acc[TI][TO];
#pragma unroll
  for (int t = 0; t < TI; ++t)
#pragma unroll
    for (int u = 0; u < TO; ++u) acc[t][u].x = acc[t][u].y = 0.0f;
  const float2* xp = X + (size_t)i0 * M2 + m;
  const float2* gp = GO + (size_t)o0 * M2 + m;
#pragma unroll 2
  for (int b = 0; b < B; ++b) {
    float2 xi[TI], go[TO];
#pragma unroll
    for (int t = 0; t < TI; ++t) xi[t] = (t < ni) ? __ldg(xp + t * M2) : make_float2(0.f, 0.f);
#pragma unroll
    for (int u = 0; u < TO; ++u) go[u] = (u < no) ? __ldg(gp + u * M2) : make_float2(0.f, 0.f);
#pragma unroll
    for (int t = 0; t < TI; ++t)
#pragma unroll
      for (int u = 0; u < TO; ++u) {
        acc[t][u].x = fmaf(xi[t].x, go[u].x, fmaf(xi[t].y, go[u].y, acc[t][u].x));
        acc[t][u].y = fmaf(xi[t].x, go[u].y, fmaf(-xi[t].y, go[u].x, acc[t][u].y));
      }
    xp += (size_t)Cin * M2;
    gp += (size_t)Cout * M2;
  }
  const bool dead = row_dead(m / m2, m1, H);
  const bool second = m >= MM;
  const int mm = second ? m - MM : m;
  float2* gW = (second ? gW2 : gW1) + mm;
#pragma unroll
  for (int t = 0; t < TI; ++t)
#pragma unroll
    for (int u = 0; u < TO; ++u)
      if (t < ni && u < no)
        gW[((size_t)(i0 + t) * Cout + o0 + u) * MM] =
            dead ? make_float2(0.f, 0.f) : make_float2(acc[t][u].x, acc[t][u].y);
}

inline int pick_bt(int B) { return B >= 8 ? 8 : (B >= 4 ? 4 : (B >= 2 ? 2 : 1)); }

template <bool CONJ>
int launch_mix(const float* Xin, const float* w1, const float* w2, float* P, int nsplit, int B, int Cred, int Cn,
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
#define PDES_MIX_CASE(bt)                                                                                     \
  case bt: {                                                                                                  \
    auto kfn = k_mix<bt, kTC, CONJ>;                                                                          \
    PDES_LAUNCH(kfn, grid, block, 0, stream, X2, A, Bw, P2, B, Cred, Cn, MM, m1, m2, H, (int)wr_stride,       \
                (int)wo_stride, red_per_split, nsplit);                                                       \
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
  return launch_mix<false>(X, w1, w2, P, nsplit, B, Cin, Cout, MM, m1, m2, H, (long)Cout * MM, (long)MM, stream,
                           "pdes_mix_fwd");
}

int pdes_mix_dx(const float* GO, const float* w1, const float* w2, float* P, int nsplit, int B, int Cin, int Cout,
                int Cgrad, int H, int m1, int m2, void* stream) {
  using namespace pdes;
  if (int e = check_mix_args("pdes_mix_dx", GO, w1, w2, P, B, Cin, Cout, H, m1, m2)) return e;
  PDES_REQUIRE(Cgrad > 0 && Cgrad <= Cin, PDES_ERR_ARG, "pdes_mix_dx: Cgrad %d not in (0,%d]", Cgrad, Cin);
  PDES_REQUIRE(nsplit >= 1 && nsplit <= Cout, PDES_ERR_ARG, "pdes_mix_dx: bad nsplit %d", nsplit);
  const int MM = m1 * m2;
  return launch_mix<true>(GO, w1, w2, P, nsplit, B, Cout, Cgrad, MM, m1, m2, H, (long)MM, (long)Cout * MM, stream,
                          "pdes_mix_dx");
}

int pdes_mix_dw(const float* X, const float* GO, float* gw1, float* gw2, int B, int Cin, int Cout, int H, int m1,
                int m2, void* stream) {
  using namespace pdes;
  if (int e = check_mix_args("pdes_mix_dw", X, GO, gw1, gw2, B, Cin, Cout, H, m1, m2)) return e;
  const int MM = m1 * m2;
  constexpr int TI = 4, TO = 4;
  const MixGeom g = mix_geom(2 * MM);
  const dim3 grid((unsigned)ceil_div(2 * MM, g.bx), (unsigned)ceil_div(Cout, TO * g.by), (unsigned)ceil_div(Cin, TI));
  PDES_REQUIRE(grid.y <= 65535 && grid.z <= 65535, PDES_ERR_UNSUPPORTED, "pdes_mix_dw: grid too large");
  auto kfn = k_mix_dw<TI, TO>;
  PDES_LAUNCH(kfn, grid, dim3(g.bx, g.by), 0, stream, reinterpret_cast<const float2*>(X),
              reinterpret_cast<const float2*>(GO), reinterpret_cast<float2*>(gw1), reinterpret_cast<float2*>(gw2), B,
              Cin, Cout, MM, m1, m2, H);
  return check_launch("pdes_mix_dw");
}

}  // extern "C"
