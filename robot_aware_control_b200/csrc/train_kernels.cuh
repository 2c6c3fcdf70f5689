// Launchers of the non-GEMM kernels of the SVG training step (train_kernels.cu). The GEMM-shaped work of the
// backward pass (dgrad, wgrad) reuses conv_tc_kernel with the fp32 segmented epilogue (EPI_F32).
// Reference: PredictionTrainer._train_step (src/prediction/trainer.py:326-465), losses.py:13-50,97-106,
// vgg_layer in train mode (vgg_64.py:8-18), ConvLSTMCell (lstm.py:129-149), reparameterize (lstm.py:276-279).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>

namespace rac {

// ---- weights: fp32 master (flat) <-> packed bf16 GEMM operands
// Wp[n][tap][c] = params[row_off[n] + col_off[c] + (flip ? taps-1-tap : tap)]  (0 where an offset is negative)
cudaError_t launch_pack_weights(const float* params, const long long* row_off, const int* col_off, int n_packed,
                                int taps, int ctot, int flip, __nv_bfloat16* wp, cudaStream_t s, int tiled = 0);
// dgrad operand: Wd[c][tap][n] = Wp[n][taps-1-tap][c], n padded with zeros to kpad
cudaError_t launch_transpose_flip(const __nv_bfloat16* wp, int n_packed, int taps, int ctot, int kpad,
                                  __nv_bfloat16* wd, cudaStream_t s);
// grads[row_off[n] + col_off[c] + tap'] = dWp[n][tap][c]
cudaError_t launch_unpack_grads(const float* dwp, const long long* row_off, const int* col_off, int n_packed, int taps,
                                int ctot, int flip, float* grads, cudaStream_t s);
// first layer (encoder.c1.0): master w[64][cin][3][3] <-> fp32 [9*cin][64]
cudaError_t launch_pack_first(const float* w, int cin, float* wf, cudaStream_t s);
cudaError_t launch_first_wgrad(const float* img4, const float* mask_a, const float* mask_b, long long mask_bstride,
                               const float* draw /* [pix][64] fp32 */, float* gw /* master layout, += */, int B,
                               int H, int W, int cin, float* part /* scratch [max_blocks][9 * cin * 64] */,
                               int max_blocks, cudaStream_t s);

// ---- BatchNorm (train mode) + LeakyReLU(0.2)
// raw: [M, C] fp32 (pre-BN conv output). mean / rstd: [C]. running stats updated `updates` times (momentum 0.1).
// groups > 1: the tensor holds `groups` time steps of M rows each; statistics (mean / rstd [groups][C]) are per step and
// the running statistics are updated step after step
cudaError_t launch_bn_stats(const float* raw, int M, int C, float* mean, float* rstd, float* running_mean,
                            float* running_var, int updates, cudaStream_t s, int groups = 1);
cudaError_t launch_bn_act(const float* raw, const float* mean, const float* rstd, const float* gamma,
                          const float* beta, int B, int H, int W, int C, __nv_bfloat16* out, int cstride, int coff,
                          int upsample, cudaStream_t s, int groups = 1);
// dy: gradient w.r.t. the layer output, fp32 rows of `dy_cstride` floats starting at channel dy_coff; with
// upsample=1 it lives on the 2H x 2W grid and the four replicated positions are summed.
// Produces draw [M, C] bf16 (gradient w.r.t. the raw conv output) and accumulates dgamma / dbeta.
cudaError_t launch_bn_bwd(const float* dy, int dy_cstride, int dy_coff, int upsample, const float* raw,
                          const float* mean, const float* rstd, const float* gamma, const float* beta, int B, int H,
                          int W, int C, float* scratch /* [groups][2*C] */, __nv_bfloat16* draw, float* draw_f32_or_null,
                          float* dgamma, float* dbeta, cudaStream_t s, int groups = 1);

// ---- NormConvLSTMCell (cfg.lstm_group_norm), pointwise + GroupNorm part: train_gn_kernels.cu
struct GnCellArgs {
  const float* raw_ih;  // [B * P, 4 * hid] fp32 convolution outputs (bias included), packed columns (channel, gate)
  const float* raw_hh;
  const float* params;  // flat parameter buffer; the six affine vectors in the reference's order
  long long g_ih, b_ih, g_hh, b_hh;  // GroupNorm(16, 4 * hid) weight / bias of ih_gates.1 and hh_gates.1 (gate * hid + ch)
  long long g_c, b_c;                // c_norm weight / bias (ch)
  const float* c_prev;  // [B * P, hid] normalised cell of the previous step (zeros at the first)
  float* gates;         // [B * P, 4 * hid] post-activation (i, f, o, g)
  float* c_raw;         // [B * P, hid] cell before c_norm
  float* c_out;         // [B * P, hid] normalised cell
  __nv_bfloat16* h_out; // [B * P, hid]
  float* stats;         // [B][48][2] mean / rstd of the 32 gate groups and 16 cell groups
  // backward only
  const float* dh;      // [B * P, hid] gradient w.r.t. h
  float* dc;            // [B * P, hid] in: gradient w.r.t. c_out from the later step; out: w.r.t. c_prev
  float* dy;            // scratch [B * P, 4 * hid] gradient w.r.t. the gate pre-activations
  __nv_bfloat16* d_ih;  // [B * P, 4 * hid] gradient w.r.t. raw_ih (GEMM operand of the ih convolution's backward)
  __nv_bfloat16* d_hh;
  float* part;          // scratch [B][14 * hid] per-sample partial sums of the affine-parameter gradients
  int B, P, hid;
};
cudaError_t launch_gn_cell_fwd(const GnCellArgs& a, cudaStream_t s);
// also adds the affine-parameter gradients to `grads` (flat, same offsets as `params`)
cudaError_t launch_gn_cell_bwd(const GnCellArgs& a, float* grads, cudaStream_t s);

// ---- ConvLSTM cell backward (elementwise). gates: saved post-activation (i,f,o,g) interleaved [M, 4*hid];
// dgates: bf16 [M, 4*hid] w.r.t. the gate pre-activations; dc is updated in place (dc_next -> dc_prev).
// ConvLSTM cell on gate pre-activations delivered as gx (or null) + nsplit split-K slices + bias (packed columns)
cudaError_t launch_lstm_cell_fwd(const float* gx, const float* part, int nsplit, long long split_stride,
                                 const float* bias, const float* c_prev_or_null, float* c_out, __nv_bfloat16* h_out,
                                 float* gates_out, int M, int hid, cudaStream_t s);
cudaError_t launch_lstm_bwd(const float* dh, float* dc, const float* gates, const float* c_prev_or_null,
                            const float* c_new, int M, int hid, __nv_bfloat16* dgates, cudaStream_t s);
// grads[bias_off[n]] += sum_m dy[m][n] for n < nvalid   (dy bf16 [M, ncols])
cudaError_t launch_bias_grad(const __nv_bfloat16* dy, int M, int ncols, int nvalid, const long long* bias_off,
                             float* grads, cudaStream_t s);

// ---- reparameterisation + KL backward. dz: fp32 [M, 64] (gradient w.r.t. z from the frame predictor input conv).
// mu/lv/eps tensors are NCHW (B, z, hw). Outputs bf16 [M, 128] interleaved (z channel, {mu, logvar}).
cudaError_t launch_gauss_bwd(const float* dz, const float* mu, const float* lv, const float* eps, const float* mu_p,
                             const float* lv_p, int B, int z_dim, int hw, float kl_weight, int bs,
                             __nv_bfloat16* dpost, __nv_bfloat16* dprior, cudaStream_t s);

// ---- frame loss (forward value + gradient w.r.t. the pre-sigmoid decoder output)
// x4: (B,4,H,W) sigmoid outputs; xj, xi: (B,3,H,W); mask: (B,1,H,W) or null. kind 0 = l1, 1 = dontcare_l1, 2 = mse,
// 3 = dontcare_mse; batch_weight (B) or null: per-sample factor of the l1 kinds (movement weighting).
// loss_out[b] = per-sample contribution (already scaled so that the step loss is sum_b loss_out[b]).
// step API (losses in the caller's autograd graph): external KL / frame gradients
cudaError_t launch_gauss_bwd_ext(const float* dz, const float* lv, const float* eps, const float* dmu, const float* dlv,
                                 const float* dmu_p, const float* dlv_p, int B, int z_dim, int hw, __nv_bfloat16* dpost,
                                 __nv_bfloat16* dprior, cudaStream_t s);
cudaError_t launch_sigmoid_bwd(const float* x4, const float* dx4, __nv_bfloat16* dlogit, int B, int HW, cudaStream_t s);
// gp_in: extra gradient w.r.t. the composited prediction (from the next step when that step consumed this
// prediction as its input, scheduled sampling) or null; gxj_out: gradient w.r.t. x_j through the composite (=) or null.
cudaError_t launch_frame_loss(const float* x4, const float* xj, const float* xi, const float* mask, int kind,
                              float robot_weight, int B, int HW, float* loss_out, __nv_bfloat16* dlogit /* [B*HW, 64] */,
                              const float* gp_in, float* gxj_out, cudaStream_t s, const float* batch_weight = nullptr,
                              int Bdiv = 0 /* batch size of one time step when B holds several; 0 = B */);
// logged robot / world MSE (losses.py:52-78) and the KL term (losses.py:97-106) over several time steps at once;
// part: scratch (2 * n resp. 64 floats); results are ADDED to out2[0..1] / out_accum[0]
cudaError_t launch_robot_world_mse_batched(const float* pred, const float* target, const float* mask, float* part,
                                           float* out2, int n, int Bdiv, int HW, cudaStream_t s);
cudaError_t launch_kl_loss_batched(const float* mu1, const float* lv1, const float* mu2, const float* lv2, float* part,
                                   float* out_accum, long long n, int bs, cudaStream_t s);
// x_pred = (1 - m) * x_j + m * rgb (trainer.py:406-407), NCHW fp32 (B,3,H,W)
cudaError_t launch_composite(const float* x4, const float* xj, float* xp, int B, int HW, cudaStream_t s);
// gradient of encoder.c1.0 w.r.t. its rgb input channels, accumulated (+=) into gimg (B,3,H,W); robot pixels of
// `zero_mask` get no gradient (zero_robot_region). draw: [pix][64] fp32, wf: [9*cin][64] fp32.
cudaError_t launch_first_dgrad(const float* draw, const float* wf, int cin, const float* zero_mask, float* gimg, int B,
                               int H, int W, cudaStream_t s);

// ---- 2x2 max-pool backward (first maximum in scan order, as torch): din (+)= route(dout)
cudaError_t launch_pool_bwd(const __nv_bfloat16* in, int in_cstride, int in_coff, const float* dout, int B, int H,
                            int W, int C, float* din, int din_cstride, int din_coff, int accumulate, cudaStream_t s);

// ---- layout helper
cudaError_t launch_cast_bf16(const float* src, long long n, __nv_bfloat16* dst, cudaStream_t s);

// ---- optimiser (torch.optim.Adam, no weight decay) and noise
// one convolution: packed weight gradient -> Adam on its flat parameter / moment slices -> bf16 operand (one pass)
cudaError_t launch_adam_pack(float* params, float* m, float* v, const float* dwp, const long long* row_off,
                             const int* col_off, int n_packed, int taps, int ctot, int flip, int tiled,
                             __nv_bfloat16* wp, float lr, float b1, float b2, float eps, int t, float grad_scale,
                             cudaStream_t s, int vec_ok);
// one-time check of a layer's packing tables (synchronous): *ok_host = 1 if launch_adam_pack may take its float4 path
cudaError_t adam_pack_vec_ok(const long long* row_off, const int* col_off, int n_packed, int taps, int ctot, int* ok_host);
cudaError_t launch_adam(float* p, const float* g, float* m, float* v, long long n, float lr, float b1, float b2,
                        float eps, int t, cudaStream_t s, float grad_scale = 1.f);
cudaError_t launch_normal_fill(float* dst, long long n, unsigned long long seed, unsigned int ctr, cudaStream_t s);
cudaError_t launch_sum_f32(const float* src, int n, float* dst_accum, cudaStream_t s);
// dst[i] = off[i] >= 0 ? params[off[i]] : 0   (packed bias vector of a convolution)
cudaError_t launch_gather_f32(const float* params, const long long* off, int n, float* dst, cudaStream_t s);
// (B,3,H,W) fp32 -> [B,H,W,4]; robot pixels of `mask` (B,1,H,W) zeroed when given (trainer.py:365-368)
cudaError_t launch_img_prep_train(const float* img_nchw, const float* mask, float* img4, int B, int HW, cudaStream_t s);

}  // namespace rac
