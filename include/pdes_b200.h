/* pdes_b200.h -- C ABI of the B200-native U-FNO / FNO spectral block.
 *
 * This is the drop-in boundary for the hot path of yoeripoels/neural-pde-surrogates.  The reference has
 * no FFI of its own (it is pure PyTorch); each entry point below cites the reference lines (relative to
 * /root/reference/src) whose work it replaces.  The reference-side binding is a `torch.autograd.Function`
 * that passes `tensor.data_ptr()` and `torch.cuda.current_stream().cuda_stream`.
 *
 * Conventions
 *   - all buffers are DEVICE pointers owned by the caller (PyTorch); the library never allocates,
 *     frees or synchronises, so every call is CUDA-graph capturable;
 *   - float32 everywhere; complex values are interleaved (re, im) pairs == torch.complex64 memory;
 *   - activations are NCHW contiguous; `stream` is a cudaStream_t passed as void*;
 *   - return value 0 = ok, otherwise a PDES_ERR_* code; `pdes_last_error()` gives the message
 *     (the Python binding raises RuntimeError/ValueError from it);
 *   - retained modes: rows kx in [0,m1) U [H-m1,H) (index k in [0,2*m1)), columns ky in [0,m2);
 *     "MM" = m1*m2; spectra are stored [B][C][2*m1][m2] complex (mode index fastest), which is also the
 *     mode order of the reference parameters weights1/weights2 [Cin][Cout][m1][m2] (proc_fno.py:240-243).
 */
#ifndef PDES_B200_H_
#define PDES_B200_H_

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PDES_OK 0
#define PDES_ERR_ARG 1         /* null pointer / non-positive size / modes out of range (proc_fno.py:135-139 asserts) */
#define PDES_ERR_UNSUPPORTED 2 /* shape needs more shared memory than one sm_100a CTA has */
#define PDES_ERR_LAUNCH 3      /* cudaGetLastError() != cudaSuccess after a launch */

#define PDES_ACT_NONE 0
#define PDES_ACT_GELU 1        /* exact erf GELU == nn.GELU() default (proc_ufno.py:44,118; proc_fno.py:94,153) */

int pdes_version(void);
const char* pdes_last_error(void);
/* 1 when built by nvcc for sm_100a, 0 for the CPU emulation build used by the CPU tests. */
int pdes_is_cuda_build(void);

/* ---- twiddle tables (host side, float64 math rounded once to float32) -------------------------------
 * The caller fills a host buffer once per (H, W, m1, m2), uploads it, and passes the device copy as
 * `tables` to the kernels below. */
size_t pdes_tables_floats(int H, int W, int m1, int m2);
int pdes_tables_fill(int H, int W, int m1, int m2, float* host_buf);

/* ---- K1: pruned forward DFT -------------------------------------------------------------------------
 * Replaces torch.fft.rfft2 + the two mode slices (proc_fno.py:261,267,269) and, with herm_scale=1, the
 * FftC2RBackward of irfft2 (GO = c_l/(HW) * DFT(g)).  The input is the channel concatenation of x0
 * [B,C0,H,W] and x1 [B,C1,H,W] (x1 may be NULL with C1=0): this removes torch.cat (proc_ufno.py:111).
 * X: [B][C0+C1][2*m1][m2] complex. */
int pdes_dft_fwd(const float* x0, int C0, const float* x1, int C1, int B, int H, int W, int m1, int m2,
                 const float* tables, int herm_scale, float* X, void* stream);

/* ---- K2: per-mode complex channel mixing ------------------------------------------------------------
 * Replaces compl_mul2d = einsum("bixy,ioxy->boxy") x2 (proc_fno.py:253-255,266-269) reading weights1 /
 * weights2 in their native parameter layout.  Output is `nsplit` partial sums P[nsplit][B][Cout][2MM]
 * (split over the reduction channel; the consumer K3a adds them, so the result is deterministic).
 * Rows of the first block that the reference overwrites when 2*m1 > H are written as zero. */
int pdes_mix_suggest_splits(int B, int Cred, int Cout, int m1, int m2);
int pdes_mix_fwd(const float* X, const float* w1, const float* w2, float* P, int nsplit,
                 int B, int Cin, int Cout, int H, int m1, int m2, void* stream);
/* adjoint w.r.t. X: GX[b,i,m] = sum_o GO[b,o,m] * conj(W[i,o,m]) for i < Cgrad (<= Cin).
 * P: [nsplit][B][Cgrad][2MM] partial sums. (BmmBackward of proc_fno.py:255) */
int pdes_mix_dx(const float* GO, const float* w1, const float* w2, float* P, int nsplit,
                int B, int Cin, int Cout, int Cgrad, int H, int m1, int m2, void* stream);
/* adjoint w.r.t. the weights, written in the parameter layout: gw1/gw2 [Cin][Cout][m1][m2] complex. */
int pdes_mix_dw(const float* X, const float* GO, float* gw1, float* gw2,
                int B, int Cin, int Cout, int H, int m1, int m2, void* stream);

/* ---- K2 on the 5th-generation tensor cores (tcgen05 + TMEM, TMA-staged tiles, 3xTF32 => fp32-faithful) -----------------
 * The same einsum (proc_fno.py:253-255,266-269) as a per-mode real 2x2-block GEMM: output channel on the M side (TMEM
 * lanes), N = (sample, re|im), K = input channel; D = Wr*[Xr|Xi] + Wi*[-Xi|Xr], so the weights are streamed once.
 *   pdes_mix_tc_pack : packed master copy Wp[m][tile][chunk][row][16 i][re|im] (<= 128 output channels per tile, 16
 *                      input channels per chunk, zero padded: every chunk is one contiguous block of 128-byte rows in
 *                      the order the kernel streams them; rows the reference overwrites when 2*m1 > H stored as
 *                      zero) of weights1/weights2; the caller rebuilds it
 *                      when the parameters change (once per optimizer step; never during a rollout).  The parameters,
 *                      their gradients, Adam and the all-reduce keep the reference layout.
 *   pdes_dft_fwd2    : K1 that additionally writes the mode-major spectrum X2[m][b][i_pad] complex.
 *   pdes_mix_tc_fwd  : O2[2][m][b][o] complex = mix(X2, Wp); partial 1 holds the part of a work item that a second CTA
 *                      finished (the chunk stream is cut into equal ranges, one per SM) and zeros otherwise.
 *   pdes_inv_h_modes : K3a for that layout: Z[b][h][2l+ri][o] = sum_k e^{+2 pi i kx_k h/H} (O2[0] + O2[1]).
 * pdes_mix_tc_ok() == 0 (B > 32, tensor-core mode < 2, no TMA driver entry point) => use pdes_mix_fwd + pdes_inv_h. */
int pdes_mix_tc_ok(int B, int Cin, int Cout, int m1, int m2);
size_t pdes_mix_tc_pack_floats(int Cin, int Cout, int m1, int m2);
size_t pdes_mix_tc_x2_floats(int B, int Cin, int m1, int m2);
size_t pdes_mix_tc_o2_floats(int B, int Cout, int m1, int m2);
int pdes_mix_tc_pack(const float* w1, const float* w2, float* Wp, int Cin, int Cout, int H, int m1, int m2, void* stream);
int pdes_dft_fwd2(const float* x0, int C0, const float* x1, int C1, int B, int H, int W, int m1, int m2,
                  const float* tables, int herm_scale, float* X, float* X2, void* stream);
int pdes_mix_tc_fwd(const float* X2, const float* Wp, float* O2, int B, int Cin, int Cout, int m1, int m2, void* stream);
int pdes_inv_h_modes(const float* O2, int B, int C, int H, int m1, int m2, const float* tables, float* Z, void* stream);

/* ---- K3a: inverse DFT along H -----------------------------------------------------------------------
 * Z[b][h][j][c] (c fastest, j = 2*l + {re,im}) = sum_k e^{+2 pi i kx_k h/H} * sum_s P[s][b][c][k][l].
 * First half of torch.fft.irfft2 (proc_fno.py:287), without ever building the zero-padded spectrum
 * (proc_fno.py:265). */
int pdes_inv_h(const float* P, int nsplit, int B, int C, int H, int m1, int m2, const float* tables,
               float* Z, void* stream);

/* ---- K3b: inverse DFT along W fused with the 1x1 conv, bias, residual and activation ------------------
 *   pre[b,o,h,w] = sum_j Z[b,h,j,o] * T[j,w]  +  sum_i At[i,o] * xin[b,i,h,w]  + bias[o] + res[b,o,h,w]
 *   out = act(pre)
 * Replaces the second half of irfft2 (proc_fno.py:287), self.w(x) (proc_fno.py:143), x1 + x2 (:146),
 * h_fno + h_unet and the GELU (proc_ufno.py:118 / proc_fno.py:153-154).  At is the 1x1 weight stored
 * [K][lda] with the output channel contiguous (forward: transpose of w.weight; backward dX: w.weight
 * itself).  Z / At / bias / res / pre may each be NULL (term skipped).  `backward_scale` selects
 * T with s_l = 1 (adjoint of K1) instead of c_l/(HW).  M = output channels, K = C0 + C1 input channels. */
int pdes_inv_w_gemm(const float* Z, const float* At, int lda, const float* x0, int C0, const float* x1, int C1,
                    const float* bias, const float* res, const float* tables, int backward_scale,
                    float* out, float* pre, int B, int M, int H, int W, int m1, int m2, int act, void* stream);

/* ---- K3b on the tensor cores (tcgen05 + TMEM, 3xTF32 split => fp32-faithful, rel. error ~2^-21 per product) -----
 * Same contract as pdes_inv_w_gemm but the 1x1-conv weights come pre-packed by pdes_gemm_tc_pack (hi/lo TF32 split
 * in the UMMA canonical K-major layout, one contiguous block per 16-channel chunk so it can be fetched with a bulk
 * async copy).  Requires N <= 256 output channels.  pdes_set_tensor_core_mode: 0 = every multiply on fp32 FFMA;
 * 1 = tcgen05 3xTF32, one tile per CTA (first version, kept for comparison); 2 = tcgen05 3xTF32, persistent
 * warp-specialised kernel (default on sm_100a); 3 = the same kernel with a single TF32 pass (hi*hi only): NOT
 * fp32-faithful (rel. error ~5e-4), offered as the separately reported reduced-precision mode. */
void pdes_set_tensor_core_mode(int mode);
int pdes_get_tensor_core_mode(void);
int pdes_gemm_tc_supported(int N, int K);
int pdes_inv_w_gemm_tc_ok(int N, int K, int H, int W, int m2, const float* x0, const float* x1);
size_t pdes_gemm_tc_pack_floats(int K, int N);
int pdes_gemm_tc_pack(const float* Wt, int lda, int K, int N, float* packed, void* stream);
/* the same operand from the transposed storage W[n][k] (row stride ldw): an nn.Conv2d(k=1) weight [Cout][Cin]
 * (proc_fno.py:114-117) feeds the forward GEMM directly, no transpose pass */
int pdes_gemm_tc_pack_t(const float* W, int ldw, int K, int N, float* packed, void* stream);
int pdes_inv_w_gemm_tc(const float* Z, const float* wpack, const float* x0, int C0, const float* x1, int C1,
                       const float* bias, const float* res, const float* tables, int backward_scale,
                       float* out, float* pre, int B, int N, int H, int W, int m1, int m2, int act, void* stream);

/* 1x1-conv weight / bias gradient on tcgen05 (3xTF32), same contract as pdes_wgrad below; ws needs
 * pdes_wgrad_tc_workspace_floats() floats.  Returns PDES_ERR_UNSUPPORTED for shapes it does not cover
 * (M > 256, K + 1 > 256, H*W % 16 != 0, unaligned pointers): the caller then uses pdes_wgrad. */
size_t pdes_wgrad_tc_workspace_floats(int M, int K);
int pdes_wgrad_tc(const float* g, const float* x0, int C0, const float* x1, int C1, float* dW, float* dbias,
                  float* ws, int B, int M, int HW, void* stream);

/* ---- decoder: fused per-pixel temporal Conv1d stack of TimeConvDense ------------------------
 * Replaces `self.decoder(z)` of reference dec_grid.py:97-146 (permute + Conv1d(1->2,k=13,s=2) + act + Conv1d(2->1,k=8)
 * per pixel) for time_window = 25, one field.  z [B][75][HW] (output of the 1x1 pre-decoder, no permute),
 * out / gy [B][25][HW], w1 [2][1][13], b1 [2], w2 [1][2][8], b2 [1].  Backward writes dz and all weight gradients
 * (deterministic two-stage reduction); ws needs pdes_timeconv_bwd_workspace_floats() floats. */
int pdes_timeconv_ok(int time_window, int num_c);
int pdes_timeconv_forward(const float* z, const float* w1, const float* b1, const float* w2, const float* b2, float* out,
                          int B, int HW, int time_window, int act, void* stream);
size_t pdes_timeconv_bwd_workspace_floats(int B, int HW, int time_window);
int pdes_timeconv_backward(const float* z, const float* gy, const float* w1, const float* b1, const float* w2,
                           const float* b2, float* dz, float* dw1, float* db1, float* dw2, float* db2, float* ws, int B,
                           int HW, int time_window, int act, void* stream);

/* ---- per-step wrapper: fused output constraints of one model application (no autograd) -----------
 * out[b,0,t,:] = mask(volume_rescale(mask(tanh(x[b,0,tw-1,:] + steps[t] * delta[b,0,t,:]))))   replaces add_delta
 * (dec_grid.py:8-23), the final activation, the obstacle masking and the 'individual_static' approximate volume
 * preservation of activation_wrapper.py:33-106 (~25 element-wise / reduction launches) for one field (num_c = 1).
 * delta, x, out: [B][tw][HW]; mask: [B][...] with batch stride mask_bstride, channel 0 used (NULL if !use_mask);
 * steps, cap: [tw] (the reference's cumulative sums of dt and of max_pct_dif). */
int pdes_constrain_forward(const float* delta, const float* x, const float* mask, int mask_bstride, const float* steps,
                           const float* cap, float* out, int B, int tw, int HW, int use_tanh, int use_mask,
                           int use_volume, void* stream);

/* ---- U-Net branch: 1x1 convolutions on the K3b / wgrad tensor-core kernels ---------------
 * Replaces forward and backward of the nn.Conv2d(k=1) layers of the reference's ResidualBlock shortcut
 * (proc_unet_modern.py:219-222) -- cuDNN runs them as SIMT sgemm at ~27 TFLOP/s.
 *   pdes_conv1x1_tc:      out[b,n,p] = act(sum_k Wt[k][n] * x[b,k,p] + bias[n] + res[b,n,p]); x [B][Cin][HW]
 *                         contiguous, out/res with batch stride out_bstride floats (>= N*HW, so a call can write a
 *                         channel sub-range of a wider tensor: the input gradient of a 385-channel conv is two
 *                         calls with N <= 256 each).  wpack = pdes_gemm_tc_pack(Wt, lda, Cin, N).
 *   pdes_wgrad_tc_range:  dW[o*ldw + i] = sum_{b,p} g[b,o,p] * x[b, c_off + i, p] for i < Cn <= 255 of a tensor with
 *                         x_ld channels, dbias[o] = sum g (optional); workspace as pdes_wgrad_tc.
 * Need H*W % 4 == 0 (forward / dX) and H*W % 16 == 0 (weight gradient), 16-byte aligned pointers, tensor-core
 * mode >= 2; otherwise PDES_ERR_UNSUPPORTED and the caller keeps cuDNN. */
int pdes_conv1x1_tc_ok(int B, int Cin, int N, int HW, const float* x);
int pdes_conv1x1_tc(const float* x, int Cin, const float* wpack, const float* bias, const float* res, float* out,
                    size_t out_bstride, int B, int N, int HW, int act, void* stream);
int pdes_wgrad_tc_range(const float* g, const float* x, int x_ld, int c_off, int Cn, float* dW, int ldw, float* dbias,
                        float* ws, int B, int M, int HW, void* stream);

/* ---- U-Net branch: 3x3 valid convolution forward on tcgen05 (3xTF32 implicit GEMM) ------
 * Replaces the forward of the nn.Conv2d(k=3, padding=0) layers of the reference's ResidualBlock
 * (proc_unet_modern.py:217-218, the "circular without padding" valid convs).  Needs W % 4 == 0, H >= 10, W >= 20,
 * N <= 256, N % 4 == 0; otherwise returns PDES_ERR_UNSUPPORTED and the caller keeps cuDNN. */
size_t pdes_conv3x3_tc_pack_floats(int Cin, int N);
int pdes_conv3x3_tc_ok(int B, int Cin, int N, int H, int W, const float* x);
int pdes_conv3x3_tc(const float* x, const float* w, const float* bias, float* wpack, float* out, int B, int Cin, int N,
                    int H, int W, void* stream);

/* ---- pointwise / small helpers ------------------------------------------------------------------------
 * g_pre = g_out * act'(pre)  (GeluBackward of proc_ufno.py:118) */
int pdes_act_bwd(const float* g_out, const float* pre, float* g_pre, size_t n, int act, void* stream);
/* out[k][m] = in[m][k]  (w.weight [Cout][Cin] -> At [Cin][Cout]) */
int pdes_transpose(const float* in, float* out, int M, int K, void* stream);
/* 1x1-conv weight and bias gradients (ConvolutionBackward of proc_fno.py:143):
 *   dW[o][i] = sum_{b,p} g[b,o,p] * xin[b,i,p],  dbias[o] = sum_{b,p} g[b,o,p]
 * ws needs pdes_wgrad_workspace_floats() floats. */
size_t pdes_wgrad_workspace_floats(int B, int M, int K, int HW);
int pdes_wgrad(const float* g, const float* x0, int C0, const float* x1, int C1, float* dW, float* dbias,
               float* ws, int B, int M, int HW, void* stream);

/* ---- U-Net branch helper: fused GroupNorm + activation ----------------------------------
 * y = act(GroupNorm_G(x) * gamma + beta) and its backward (reference proc_unet_modern.py:234-245: norm -> activation
 * in front of every 3x3 conv; :155,:194 before the final conv).  x, y: [B][C][HW]; stats [B*G][2] = (mean, rstd) is
 * written by the forward and read by the backward; ws: pdes_gn_workspace_bytes() bytes (8-byte aligned). */
size_t pdes_gn_workspace_bytes(int B, int C, int HW, int G);
int pdes_gn_act_forward(const float* x, const float* gamma, const float* beta, float eps, float* y, float* stats,
                        void* ws, int B, int C, int HW, int G, int act, void* stream);
int pdes_gn_act_backward(const float* dy, const float* x, const float* gamma, const float* beta, const float* stats,
                         float* dx, float* dgamma, float* dbeta, void* ws, int B, int C, int HW, int G, int act,
                         void* stream);

/* ---- fused chains (what the nn.Module binding calls) ----------------------------------------------------
 * One FNO_Layer / U-FNO block tail, forward:  K1 -> K2 -> K3a -> K3b.
 *   h [B,C0,H,W], vb [B,C1,H,W] or NULL, w1/w2 complex [Cin][Cout][m1][m2], wc [Cout][Cin] = w.weight in its
 *   parameter layout (NULL = no 1x1), wspec = pdes_mix_tc_pack(w1, w2) cached by the caller per weight version or NULL
 *   (then K2 runs from the parameter layout on the FFMA kernels), wpack = pdes_gemm_tc_pack_t(wc, Cin, Cin, Cout) cached by the caller per weight
 *   version or NULL (then it is packed into the workspace on every call), bias [Cout] or NULL, res [B,Cout,H,W] or
 *   NULL (the U-Net branch), out [B,Cout,H,W], pre (NULL unless the backward will need it), Xsave [B][Cin][2MM]
 *   complex (kept for the backward), ws: pdes_block_fwd_workspace_floats() floats. */
size_t pdes_block_fwd_workspace_floats(int B, int Cin, int Cout, int H, int W, int m1, int m2);
int pdes_block_forward(const float* h, int C0, const float* vb, int C1, const float* w1, const float* w2,
                       const float* wspec, const float* wc, const float* wpack, const float* bias, const float* res,
                       const float* tables, float* Xsave, float* ws, float* out, float* pre,
                       int B, int Cout, int H, int W, int m1, int m2, int act, void* stream);
/* backward of the same block.  Inputs: g_out, pre (if act != none), h, vb, Xsave, w1, w2, wc [Cout][Cin],
 * wpack = pdes_gemm_tc_pack(wc, Cin, Cout, C0) or NULL.
 * Outputs: g_pre [B,Cout,H,W] (also the gradient of `res`), dh [B,C0,H,W], gw1/gw2 (parameter layout),
 * dwc [Cout][Cin] and dbias [Cout] (both may be NULL when there is no 1x1 conv). */
size_t pdes_block_bwd_workspace_floats(int B, int C0, int C1, int Cout, int H, int W, int m1, int m2);
int pdes_block_backward(const float* g_out, const float* pre, const float* h, int C0, const float* vb, int C1,
                        const float* Xsave, const float* w1, const float* w2, const float* wspec, const float* wc,
                        const float* wpack, const float* tables, float* ws, float* g_pre, float* dh, float* gw1, float* gw2,
                        float* dwc, float* dbias,
                        int B, int Cout, int H, int W, int m1, int m2, int act, void* stream);

/* Adjoint of the per-mode channel mix w.r.t. the spectrum on tcgen05, from the FORWARD pack of pdes_mix_tc_pack (read as an
 * MN-major operand: no second copy of the weights).  GO2 = mode-major gradient spectrum (pdes_dft_fwd2 on the output
 * gradient with herm_scale = 1), O2 in the layout of pdes_mix_tc_fwd's output for the first C0 input channels.
 * Replaces the autograd of compl_mul2d w.r.t. its input, proc_fno.py:253-255. */
int pdes_mix_tc_dx_ok(int B, int Cin, int Cout, int C0, int m1, int m2);
int pdes_mix_tc_dx(const float* GO2, const float* Wp, float* O2, int B, int Cin, int Cout, int C0, int m1, int m2,
                   void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PDES_B200_H_ */
