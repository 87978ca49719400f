// Fused launch chains for one FNO_Layer / U-FNO block tail (forward and backward).
// These are the two entry points the nn.Module binding calls; each enqueues a fixed sequence of kernels on
// the caller's stream (no allocation, no synchronisation => CUDA-graph capturable).
#include "pdes_common.cuh"

namespace pdes {
namespace {

struct FwdWs { size_t P, Z, PK, WT, X2, O2, total; int nsplit; };
FwdWs fwd_ws(int B, int Cin, int Cout, int H, int W, int m1, int m2) {
  (void)W;
  FwdWs w;
  w.nsplit = pdes_mix_suggest_splits(B, Cin, Cout, m1, m2);
  const size_t M2 = (size_t)2 * m1 * m2;
  w.P = 0;
  w.Z = w.P + round4((size_t)w.nsplit * B * Cout * M2 * 2);
  w.PK = w.Z + round4((size_t)B * H * 2 * m2 * Cout);
  w.WT = w.PK + round4(pdes_gemm_tc_pack_floats(Cin, Cout));         // packed 1x1 weights (tensor-core mode)
  w.X2 = w.WT + round4((size_t)Cin * Cout);                            // transposed 1x1 weights (FFMA mode)
  w.O2 = w.X2 + round4(pdes_mix_tc_x2_floats(B, Cin, m1, m2));         // mode-major spectrum for K2 on tcgen05
  w.total = w.O2 + round4(pdes_mix_tc_o2_floats(B, Cout, m1, m2));     // its output (two partials)
  return w;
}

struct BwdWs { size_t GO, PX, Zg, WG, PK, GO2, O2g, total; int nsplit; };
BwdWs bwd_ws(int B, int Cin, int C0, int Cout, int H, int W, int m1, int m2, bool has_conv) {
  BwdWs w;
  w.nsplit = pdes_mix_suggest_splits(B, Cout, C0, m1, m2);
  const size_t M2 = (size_t)2 * m1 * m2;
  w.GO = 0;
  w.PX = w.GO + round4((size_t)B * Cout * M2 * 2);
  w.Zg = w.PX + round4((size_t)w.nsplit * B * C0 * M2 * 2);
  w.WG = w.Zg + round4((size_t)B * H * 2 * m2 * C0);
  size_t wg = has_conv ? pdes_wgrad_workspace_floats(B, Cout, Cin, H * W) : 0;
  if (has_conv && pdes_wgrad_tc_workspace_floats(Cout, Cin) > wg) wg = pdes_wgrad_tc_workspace_floats(Cout, Cin);
  w.PK = w.WG + round4(wg);
  w.GO2 = w.PK + (has_conv ? round4(pdes_gemm_tc_pack_floats(Cout, C0)) : 0);
  // dX on tcgen05 from the forward pack: mode-major gradient spectrum and its output in K2's O2 layout
  const bool dx_tc = pdes_mix_tc_dx_ok(B, Cin, Cout, C0, m1, m2) != 0;
  w.O2g = w.GO2 + (dx_tc ? round4(pdes_mix_tc_x2_floats(B, Cout, m1, m2)) : 0);
  w.total = w.O2g + (dx_tc ? round4(pdes_mix_tc_o2_floats(B, C0, m1, m2)) : 0);
  return w;
}

}  // namespace
}  // namespace pdes

extern "C" {

size_t pdes_block_fwd_workspace_floats(int B, int Cin, int Cout, int H, int W, int m1, int m2) {
  if (B <= 0 || Cin <= 0 || Cout <= 0 || H <= 0 || W <= 0 || m1 <= 0 || m2 <= 0) return 0;
  return pdes::fwd_ws(B, Cin, Cout, H, W, m1, m2).total;
}

int pdes_block_forward(const float* h, int C0, const float* vb, int C1, const float* w1, const float* w2,
                       const float* wspec, const float* wc, const float* wpack, const float* bias, const float* res,
                       const float* tables,
                       float* Xsave, float* ws, float* out, float* pre, int B, int Cout, int H, int W, int m1, int m2,
                       int act, void* stream) {
  using namespace pdes;
  PDES_REQUIRE(h && w1 && w2 && tables && Xsave && ws && out, PDES_ERR_ARG, "pdes_block_forward: null pointer");
  PDES_REQUIRE(wpack == nullptr || wc != nullptr, PDES_ERR_ARG, "pdes_block_forward: wpack without wc");
  const int Cin = C0 + C1;
  const FwdWs w = fwd_ws(B, Cin, Cout, H, W, m1, m2);
  float* P = ws + w.P;
  float* Z = ws + w.Z;
  if (wspec != nullptr && pdes_mix_tc_ok(B, Cin, Cout, m1, m2)) {
    // K2 on tcgen05 from the packed master copy: K1 also writes the mode-major spectrum, K3a reads K2's layout
    float* X2 = ws + w.X2;
    float* O2 = ws + w.O2;
    if (int e = pdes_dft_fwd2(h, C0, vb, C1, B, H, W, m1, m2, tables, 0, Xsave, X2, stream)) return e;
    if (int e = pdes_mix_tc_fwd(X2, wspec, O2, B, Cin, Cout, m1, m2, stream)) return e;
    if (int e = pdes_inv_h_modes(O2, B, Cout, H, m1, m2, tables, Z, stream)) return e;
  } else {
    if (int e = pdes_dft_fwd(h, C0, vb, C1, B, H, W, m1, m2, tables, 0, Xsave, stream)) return e;
    if (int e = pdes_mix_fwd(Xsave, w1, w2, P, w.nsplit, B, Cin, Cout, H, m1, m2, stream)) return e;
    if (int e = pdes_inv_h(P, w.nsplit, B, Cout, H, m1, m2, tables, Z, stream)) return e;
  }
  if (wc != nullptr && pdes_get_tensor_core_mode() && pdes_inv_w_gemm_tc_ok(Cout, Cin, H, W, m2, h, vb)) {
    const float* PK = wpack;
    if (PK == nullptr) {                       // no cached operand: pack wc[o][i] as At[k = i][n = o] on the fly
      if (int e = pdes_gemm_tc_pack_t(wc, Cin, Cin, Cout, ws + w.PK, stream)) return e;
      PK = ws + w.PK;
    }
    return pdes_inv_w_gemm_tc(Z, PK, h, C0, vb, C1, bias, res, tables, 0, out, pre, B, Cout, H, W, m1, m2, act, stream);
  }
  const float* wct = nullptr;
  if (wc != nullptr) {                         // FFMA kernel reads At[i][o]
    if (int e = pdes_transpose(wc, ws + w.WT, Cout, Cin, stream)) return e;
    wct = ws + w.WT;
  }
  return pdes_inv_w_gemm(Z, wct, Cout, h, C0, vb, C1, bias, res, tables, 0, out, pre, B, Cout, H, W, m1, m2, act,
                         stream);
}

size_t pdes_block_bwd_workspace_floats(int B, int C0, int C1, int Cout, int H, int W, int m1, int m2) {
  if (B <= 0 || C0 <= 0 || C1 < 0 || Cout <= 0 || H <= 0 || W <= 0 || m1 <= 0 || m2 <= 0) return 0;
  return pdes::bwd_ws(B, C0 + C1, C0, Cout, H, W, m1, m2, true).total;
}

int pdes_block_backward(const float* g_out, const float* pre, const float* h, int C0, const float* vb, int C1,
                        const float* Xsave, const float* w1, const float* w2, const float* wspec, const float* wc,
                        const float* wpack, const float* tables, float* ws, float* g_pre, float* dh, float* gw1, float* gw2,
                        float* dwc, float* dbias, int B, int Cout, int H, int W, int m1, int m2, int act, void* stream) {
  using namespace pdes;
  PDES_REQUIRE(g_out && h && Xsave && w1 && w2 && tables && ws && dh && gw1 && gw2, PDES_ERR_ARG,
               "pdes_block_backward: null pointer");
  PDES_REQUIRE(act == PDES_ACT_NONE || (pre != nullptr && g_pre != nullptr), PDES_ERR_ARG,
               "pdes_block_backward: activation backward needs pre and g_pre");
  const int Cin = C0 + C1;
  const bool has_conv = wc != nullptr;
  const BwdWs w = bwd_ws(B, Cin, C0, Cout, H, W, m1, m2, has_conv);
  float* GO = ws + w.GO;
  float* PX = ws + w.PX;
  float* Zg = ws + w.Zg;
  float* WG = ws + w.WG;
  const float* gp = g_out;
  if (act != PDES_ACT_NONE) {
    if (int e = pdes_act_bwd(g_out, pre, g_pre, (size_t)B * Cout * H * W, act, stream)) return e;
    gp = g_pre;
  }
  // dX of the channel mix on tcgen05 straight from the FORWARD pack (read as an MN-major operand): needs the pack and the
  // mode-major copy of the gradient spectrum
  const bool dx_tc = wspec != nullptr && pdes_mix_tc_dx_ok(B, Cin, Cout, C0, m1, m2) != 0;
  // GO = c_l/(HW) * DFT(g_pre)                     (adjoint of K3)
  if (dx_tc) {
    if (int e = pdes_dft_fwd2(gp, Cout, nullptr, 0, B, H, W, m1, m2, tables, 1, GO, ws + w.GO2, stream)) return e;
  } else if (int e = pdes_dft_fwd(gp, Cout, nullptr, 0, B, H, W, m1, m2, tables, 1, GO, stream)) {
    return e;
  }
  // weight gradients in the parameter layout       (adjoint of K2 w.r.t. W)
  if (int e = pdes_mix_dw(Xsave, GO, gw1, gw2, B, Cin, Cout, H, m1, m2, stream)) return e;
  // GX for the C0 channels that need a gradient    (adjoint of K2 w.r.t. X), then its H-axis inverse (unit weights)
  if (dx_tc) {
    if (int e = pdes_mix_tc_dx(ws + w.GO2, wspec, ws + w.O2g, B, Cin, Cout, C0, m1, m2, stream)) return e;
    if (int e = pdes_inv_h_modes(ws + w.O2g, B, C0, H, m1, m2, tables, Zg, stream)) return e;
  } else {
    if (int e = pdes_mix_dx(GO, w1, w2, PX, w.nsplit, B, Cin, Cout, C0, H, m1, m2, stream)) return e;
    if (int e = pdes_inv_h(PX, w.nsplit, B, C0, H, m1, m2, tables, Zg, stream)) return e;
  }
  // dh = Re(pruned inverse of GX, unit weights) + wc^T g_pre      (adjoint of K1 fused with the 1x1 dX)
  if (has_conv && pdes_get_tensor_core_mode() && pdes_inv_w_gemm_tc_ok(C0, Cout, H, W, m2, gp, nullptr)) {
    const float* PK = wpack;                   // At[k = o][n = i] = wc[o][i], lda = Cin
    if (PK == nullptr) {
      if (int e = pdes_gemm_tc_pack(wc, Cin, Cout, C0, ws + w.PK, stream)) return e;
      PK = ws + w.PK;
    }
    if (int e = pdes_inv_w_gemm_tc(Zg, PK, gp, Cout, nullptr, 0, nullptr, nullptr, tables, 1, dh, nullptr, B, C0, H, W,
                                   m1, m2, PDES_ACT_NONE, stream))
      return e;
  } else if (int e = pdes_inv_w_gemm(Zg, wc, Cin, has_conv ? gp : nullptr, Cout, nullptr, 0, nullptr, nullptr, tables, 1,
                                     dh, nullptr, B, C0, H, W, m1, m2, PDES_ACT_NONE, stream)) {
    return e;
  }
  if (has_conv && (dwc != nullptr || dbias != nullptr)) {
    int e = PDES_ERR_UNSUPPORTED;
    if (pdes_get_tensor_core_mode() >= 2) e = pdes_wgrad_tc(gp, h, C0, vb, C1, dwc, dbias, WG, B, Cout, H * W, stream);
    if (e == PDES_ERR_UNSUPPORTED) e = pdes_wgrad(gp, h, C0, vb, C1, dwc, dbias, WG, B, Cout, H * W, stream);
    if (e) return e;
  }
  return PDES_OK;
}

}  // extern "C"
