// Host-side launchers of the small non-GEMM kernels (misc_kernels.cu, cem_kernels.cu).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>

#include "../../include/racb200.h"

namespace rac {

// encoder.c1.0 (reference vgg_64.py:99-101 via vgg_layer :8-18): 3x3 conv over [rgb | mask_t | mask_t+1], folded
// eval BatchNorm, LeakyReLU(0.2). K = 27..45 is too small for a tensor-core tile; HBM-write bound.
// mask planes of consecutive candidates are `mask_bstride` floats apart.
cudaError_t launch_first_conv(const float* img4, const float* mask_a, const float* mask_b, long long mask_bstride,
                              const float* w, const float* bias, __nv_bfloat16* out, int B, int H, int W, int cin,
                              cudaStream_t s, float* raw_out = nullptr);
// the same layer on the tensor cores (first_conv_tc.cu): per-thread im2col rows written in the swizzled UMMA layout,
// 6 MMAs per 256-pixel tile; inference only (bf16 inputs, fused activation)
cudaError_t launch_first_conv_tc(const float* img4, const float* mask_a, const float* mask_b, long long mask_bstride,
                                 const float* w, const float* bias, __nv_bfloat16* out, int B, int H, int W, int cin,
                                 int num_sms, cudaStream_t s);
cudaError_t first_conv_tc_set_attributes();
// nn.MaxPool2d(2,2) (reference vgg_64.py:120,126-128) over a channel slice of an NHWC buffer
cudaError_t launch_maxpool2(const __nv_bfloat16* in, int in_cstride, int in_coff, __nv_bfloat16* out, int B, int H,
                            int W, int C, cudaStream_t s);
// tiled action / robot-state channels (reference dynamics.py:591-603) as one zero-padded 64-channel k-block
// action rows are `astride` floats apart (a time slice of an (n, steps, action_dim) tensor); action may be null
// (posterior input: robot state only, reference dynamics.py:621-623)
cudaError_t launch_aux_tile(const float* action, int astride, int adim, const float* r, const float* r2, int rdim,
                            __nv_bfloat16* aux, int B, int HW, cudaStream_t s);
// start image uint8 HWC -> per-candidate fp32 NHWC4 / 255, robot pixels of mask_0 zeroed
// (reference trajectory_sampler.py:130-131,141-142)
cudaError_t launch_img_prep_u8(const uint8_t* img_hwc, const float* mask0, int zero_robot, float* img4, int B, int H,
                               int W, cudaStream_t s);
// NCHW fp32 (B,3,H,W) -> NHWC4
cudaError_t launch_img_prep_nchw(const float* img_nchw, float* img4, int B, int H, int W, cudaStream_t s);
// goal images uint8 (G,H,W,3) -> fp32 (G,H,W,4) / 255 (reference trajectory_sampler.py:77-79)
cudaError_t launch_goal_prep(const uint8_t* goal_hwc, float* goal4, int G, int H, int W, cudaStream_t s);
// per-candidate cost from the per-tile partials of the frame epilogue (reference losses.py:224-235,244-263,307-335;
// fp64 accumulation over steps as trajectory_sampler.py:74,169)
// peers != null (last step of a sharded plan): the finished cost is also stored into every rank's gathered vector
cudaError_t launch_cost_finish(const float* cost_part, int nparts, int dontcare, float weight, int accumulate,
                               double* sum_cost, float* step_cost, int B, cudaStream_t s,
                               double* const* peers = nullptr, int peer_world = 0, long long peer_offset = 0);
cudaError_t launch_peer_barrier(uint32_t* const* pads, int base, int rank, int world, uint32_t seq, cudaStream_t s);
// stand-alone planning cost in the reference's own tensor layout (NCHW fp32): ImgL2Cost / ImgDontcareCost
cudaError_t launch_masked_cost(const float* curr, const float* goal, const float* curr_mask, const float* goal_mask,
                               int dontcare, float* out, int B, int HW, cudaStream_t s);
// training criteria, forward value only (reference losses.py:13-19,35-50,97-106)
cudaError_t launch_l1_loss(const float* pred, const float* target, float* out, int64_t n, cudaStream_t s);
cudaError_t launch_dontcare_l1_loss(const float* pred, const float* target, const float* mask, float robot_weight,
                                    float* out, int B, int HW, cudaStream_t s);
// l1 / dontcare_l1 / mse / dontcare_mse (trainer.py:149-161) with optional per-sample batch weight
cudaError_t launch_recon_loss(const float* pred, const float* target, const float* mask, const float* batch_weight,
                              int kind, float robot_weight, float* per_sample, float* out, int B, int HW,
                              cudaStream_t s);
// out2[0] += robot_mse_criterion, out2[1] += world_mse_criterion (losses.py:52-78)
cudaError_t launch_robot_world_mse(const float* pred, const float* target, const float* mask, float* out2, int B,
                                   int HW, cudaStream_t s);
cudaError_t launch_kl_loss(const float* mu1, const float* lv1, const float* mu2, const float* lv2, float* out,
                           int64_t n, int bs, cudaStream_t s);

// ---- evaluation metrics (metric_kernels.cu; reference src/utils/metrics.py:13-78, losses.py:80-94) ----
// mode 0: psnr of metrics.py (both images through (x + 1) / 2), optional robot-region zeroing + clamp; mode 1: world_psnr
cudaError_t launch_psnr(const float* est, const float* tgt, const float* mask, int mode, int clamp01, float* out,
                        int B, int C, int HW, cudaStream_t s);
// 11x11 gaussian SSIM map (B,C,H,W) and / or its per-plane means (B*C); optional robot-region zeroing of both images
cudaError_t launch_ssim(const float* img1, const float* img2, const float* mask, float* map_out, float* plane_mean,
                        int B, int C, int H, int W, cudaStream_t s);

// (1 - m) * x_j + m * rgb from the 4-channel decoder output, NCHW fp32 (trainer.py:653-654)
cudaError_t launch_composite_nchw(const float* x4, const float* xj, float* out, int B, int HW, cudaStream_t s);

// ---- NormConvLSTMCell pointwise part (norm_lstm.cu; reference lstm.py:177-198) ----
// ih / hh: raw gate convolutions [B, P, 4*hid] fp32, packed column (channel, gate); gn_params: packed GroupNorm affine
// [ih gamma | ih beta | hh gamma | hh beta] (4*hid each, packed column order) + [cell gamma | cell beta] (hid each)
// (ih is also scratch: its `out` gate slot carries the output gate from pass 2 to pass 3)
cudaError_t launch_norm_lstm_cell(float* ih, const float* hh, const float* gn_params, float* c_state,
                                  __nv_bfloat16* h_out, int B, int P, int hid, cudaStream_t s);
// gate convolutions ran with EPI_GATES: GroupNorm partial sums gn_part[B][2][16][6][2] come from their epilogues
// (nslots used per group), obuf = fp32 scratch [B, P, hid] for the output gate
cudaError_t launch_norm_lstm_cell_fused(const float* ih, const float* hh, const float* gn_part, int nslots,
                                        const float* gn_params, float* c_state, float* obuf, __nv_bfloat16* h_out,
                                        int B, int P, int hid, cudaStream_t s);
cudaError_t norm_lstm_set_attributes();

// ---- robot state / mask producer (robot_kernels.cu; reference wx250s_model.py:57-182, franka_model.py:30-80) ----
cudaError_t launch_robot_states(const rac_robot_model* m, const float* start_state, const float* actions, int n, int L,
                                int adim, float* states, long long t_stride, cudaStream_t s);
cudaError_t launch_robot_masks(const rac_robot_model* m, const float* states, long long st_stride, int n, int T1, int H,
                               int W, float extra_radius, float* masks, long long m_stride, cudaStream_t s);

// ---- CEM (cem_kernels.cu) ----
cudaError_t launch_cem_sample(const float* mean, const float* stdv, const float* noise, unsigned long long seed,
                              int iter, int n_total, int L, int adim_model, int cand_offset, int n_local,
                              float clampv, float* act2, float* act5, cudaStream_t s);
cudaError_t launch_topk(const double* costs, int n, int k, int64_t* idx_out, double* val_out, cudaStream_t s);
cudaError_t launch_refit(const float* act2, int L2, const int64_t* idx, int k, float std_floor, float* mean_out,
                         float* std_out, cudaStream_t s);
cudaError_t cem_set_attributes();

// encoder outputs of candidate 0 -> candidates first .. B-1 (channels [coff, coff + C) of an NHWC bf16 tensor)
cudaError_t launch_broadcast_candidate(__nv_bfloat16* buf, int HW, int cstride, int coff, int C, int first, int B,
                                       cudaStream_t s);

}  // namespace rac
