// Shared definitions for the implicit-GEMM convolution kernels (tcgen05 product path and the SIMT
// cross-check kernel) and their fused epilogues.
//
// Activations are NHWC bf16: tensor [B, H, W, C] with C a multiple of 64. A convolution is the GEMM
//   D[m, n] = sum_{tap, c} A[(b, y + kh - pad, x + kw - pad), c] * Wp[n, tap * Ctot + c]
// with m = (b, y, x). One M-tile is BLOCK_M (128 or 256) rows = a TMA box {64 ch, W, BH, NB} (W * BH * NB == BLOCK_M); zero padding
// comes from TMA out-of-bounds fill, the channel concat of several inputs (reference: torch.cat at
// src/prediction/models/dynamics.py:600-641, lstm.py:132, vgg_64.py:236-240) is a loop over source tensors.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>

namespace rac {

// EPI_LSTM = inference cell (MUFU math, epilogue-private cell-state layout); EPI_LSTM_TRAIN = training cell (libm math,
// NHWC cell state read from c_in, gates saved for the backward pass) -- separate instantiations keep each one's code small
// EPI_GATES = raw fp32 gate pre-activations of a NormConvLSTMCell convolution + per-sample GroupNorm partial sums
// EPI_F32_BT = EPI_F32 with the weight operand read "transposed" straight from the FORWARD packing Wp[n][tap][c] (dgrad
// of the training step: B[c][(tap, n)] = Wp[n][taps-1-tap][c] is an MN-major operand of tcgen05 -- no transposed,
// flipped copy of the weights is ever materialised)
enum EpiMode : int { EPI_ACT = 0, EPI_LSTM = 1, EPI_GAUSS = 2, EPI_FRAME = 3, EPI_F32 = 4, EPI_LSTM_TRAIN = 5, EPI_GATES = 6,
                     EPI_F32_BT = 7 };

// One destination of the fp32 epilogue: packed columns [n_begin, n_end) of the GEMM go to dst (row-major NHWC rows,
// `cstride` floats per row, starting at channel `coff`); accumulate: += instead of =. Bounds are multiples of 32.
struct F32Seg {
  int n_begin, n_end;
  float* dst;
  int cstride, coff, accumulate;
};

constexpr int kMaxSrc = 3;
constexpr int kTileM = 128;
constexpr int kBlockK = 64;  // bf16 channels per k-block = one 128-byte swizzle row

struct ConvGeom {
  int B, H, W;          // candidates, output (= input) spatial size
  int BH, NB;           // rows / candidates per M-tile (W * BH * NB == BLOCK_M)
  int ks, pad;          // square filter, 'same' padding
  int nsrc;             // concatenated inputs
  int src_kb[kMaxSrc];  // k-blocks (channels / 64) per input
  int src_dead[kMaxSrc];  // 1: this input is known to be all zero (LSTM h_prev right after init_hidden) -> its k-blocks are skipped
  int ctot;             // total input channels (sum of padded source channels)
  int num_m_tiles, num_n_tiles, tiles_per_img;  // tiles_per_img = H / BH
  int w_shift, bhw_shift;                       // log2(W), log2(BH * W)
  int no_split_tail;                            // 1: disable the tail splitting of conv_tc_kernel (A/B measurements)
  int epi_staged;                               // 1 / 2: fp32 epilogues store through the per-warp transpose tile (32- / 128-bit)
  // 1: the rows of an M-tile are ordered (row of the map, candidate, column) instead of (candidate, row, column): the TMA
  // box is {64, W, NB, BH} on a (C, W, B, H) view of the tensor. With NB * W == 128 each 128-row MMA sub-tile is then
  // ONE row of the 6x8 latent map, and the MMAs of a sub-tile whose input row for a filter tap lies in the zero
  // padding are not issued at all (the ConvLSTM gate convolutions run at the power cap: fewer MMAs = faster).
  int y_major;
  int nbw_shift;                                // log2(NB * W)
  int ksplit;                                   // > 1: split-K work items (conv_tc_kernel, fp32 epilogues only)
  unsigned long long* timeline;                 // debug (RAC_TRAIN_TIMELINE): per-CTA %globaltimer stamps [grid][8], or null
  int w_prefetch;                               // > 0: a spare thread prefetches the weight boxes into L2 this many k-blocks ahead (conv_tc_kernel)
  int w_tiled;                                  // 1: weights in the k-block-major packing [kb][n][64] (training; rac_api.cu::encode_w_map_tiled)
};

// row r of an M-tile -> (candidate, y, x)
__device__ __forceinline__ void tile_row_to_pixel(const ConvGeom& g, int grp, int yb, int r, int& b, int& y, int& x) {
  if (g.y_major) {
    b = grp * g.NB + ((r >> g.w_shift) & (g.NB - 1));
    y = yb * g.BH + (r >> g.nbw_shift);
  } else {
    b = grp * g.NB + (r >> g.bhw_shift);
    y = yb * g.BH + ((r >> g.w_shift) & (g.BH - 1));
  }
  x = r & (g.W - 1);
}

struct EpiParams {
  const float* bias;  // [num_n_tiles * BLOCK_N], packed column order
  int cout;           // valid packed output columns
  // EPI_ACT: y = act(acc + bias) -> bf16 NHWC, optional channel offset into a concat buffer and 2x nearest upsample
  __nv_bfloat16* out;
  int out_cstride, out_coff, upsample, lrelu;
  // EPI_LSTM (reference lstm.py:135-149): columns are (channel, gate) interleaved, gate order in/remember/out/cell
  float* c_state;        // [B,H,W,hid] fp32, updated in place
  __nv_bfloat16* h_out;  // [B,H,W,hid]
  int hid;
  const float* c_in;     // training: previous cell state read from here (null: c_state, in place)
  float* gates_out;      // training: post-activation gates fp32 [B*H*W, 4*hid] in packed column order (null: not saved)
  int c_tiled;           // EPI_LSTM: c_state uses the epilogue-private tiled layout (epilogue.cuh::lstm_load_c)
  int exact_math;        // (set by the training path; EPI_LSTM_TRAIN always uses libm tanh / exp: gradients are checked against fp32 autograd)
  // EPI_F32 (training: raw pre-BatchNorm conv output, dgrad into gradient accumulators, wgrad into packed dW)
  F32Seg seg[3];
  int nseg;
  // split-K (ConvGeom::ksplit > 1): raw accumulators go to split_part[split][pixel][num_n_tiles * BLOCK_N] instead, and
  // launch_splitk_reduce applies bias / segments afterwards
  float* split_part;
  long long split_stride;
  // EPI_GATES (lstm_group_norm, reference lstm.py:163-171,181-183): acc + bias -> seg[0].dst (fp32 [rows, 4*hid], packed
  // (channel, gate) columns) and GroupNorm(16, 4*hid) partial sums of this tile: gn_part[b][gn_tensor][gate*4 + quarter]
  // [slot][{sum, sumsq}], slot = row tile * gn_ntq + column tile inside the quarter (6 slots reserved)
  float* gn_part;
  int gn_tensor, gn_ntq;
  // EPI_GAUSS (reference lstm.py:276-286): columns are (z channel, {mu, logvar}) interleaved
  const float* eps;         // NCHW (B, z_dim, H, W) fp32 noise, or null -> Philox
  float* mu_out;            // NCHW fp32 or null
  float* logvar_out;        // NCHW fp32 or null
  __nv_bfloat16* z_out;     // [B,H,W,64] (channels >= z_dim written as 0)
  int z_dim, sample_mean;
  unsigned long long seed;
  unsigned int noise_ctr;   // (iteration, step) counter for Philox
  int cand_offset;          // global id of candidate 0 of this rank
  // EPI_FRAME (reference trajectory_sampler.py:148-168, losses.py:224-263, image.py:5-20)
  const float* curr_img;   // [B,H,W,4] fp32 (rgb + pad)
  float* next_img;         // [B,H,W,4] fp32
  const float* mask_next;  // [B,H,W] fp32 {0,1} or null
  const float* goal_img;   // [H,W,4] fp32
  const float* goal_mask;  // [H,W] fp32 or null
  float* xpred_out;        // NCHW (B,4,H,W) fp32 or null
  float* cost_part;        // [B][cost_nparts][2] (sum of squares, world pixel count), one partial per warp, or null
  int cost_nparts;         // partials per candidate: H*W/32 (generic kernel) or 128 (halo kernel: 16 tiles x 8 warps)
  int zero_robot, dontcare;
};

struct ConvTmaps {
  CUtensorMap a[kMaxSrc];
  CUtensorMap w;
};

// raw pointers for the SIMT cross-check kernel
struct ConvRaw {
  const __nv_bfloat16* src[kMaxSrc];
  const __nv_bfloat16* w;  // [n_pad][ks*ks*ctot]
};

struct ConvOp {
  ConvGeom g;
  EpiParams e;
  ConvTmaps tm;
  ConvRaw raw;
  ConvTmaps tm_mc;       // multicast kernel (conv_tc_mc.cu): 128-row y-major activation boxes, 256-row weight box
  int mc;                // 1: launch conv_tc_mc_kernel (2-CTA cluster sharing the activation tile by TMA multicast)
  ConvTmaps tm2;         // CTA-pair kernel (conv_tc2.cu): 128-row activation boxes, 128-row weight box
  int two_cta;           // 1: launch conv_tc2_kernel (cta_group::2) instead of conv_tc_kernel
  CUtensorMap tm_halo;   // halo kernel (conv_halo.cu): activation map with a column-major box
  int halo;              // 1: launch conv_halo_kernel instead of conv_tc_kernel
  int halo_column_loads; // 1: tm_halo is the per-column fallback map
  int block_m;
  int block_n;
  int epi;
  const char* name;
};

// host launchers (conv_tc.cu)
cudaError_t launch_conv_tc(const ConvOp& op, int num_sms, cudaStream_t stream);
cudaError_t launch_conv_simt(const ConvOp& op, cudaStream_t stream);
cudaError_t conv_tc_set_attributes();
// sums the split-K slices in order and applies the fp32 epilogue (bias, destination segments, accumulate flags)
cudaError_t launch_splitk_reduce(const EpiParams& e, int ksplit, long long rows, int ncols, cudaStream_t stream);
// halo-tile kernel for the 64-wide full-resolution 3x3 layers (conv_halo.cu)
bool conv_halo_supported(const ConvOp& op);
cudaError_t launch_conv_halo(const ConvOp& op, const CUtensorMap& tm_a, int column_loads, int use_base_offset,
                             int num_sms, cudaStream_t stream);
cudaError_t conv_halo_set_attributes();
// LSTM gate convolution with the activation tile multicast inside a 2-CTA cluster (conv_tc_mc.cu)
bool conv_tc_mc_supported(const ConvOp& op);
cudaError_t launch_conv_tc_mc(const ConvOp& op, const ConvTmaps& tm_mc, int num_sms, cudaStream_t stream);
cudaError_t conv_tc_mc_set_attributes();
// CTA-pair (cta_group::2) kernel for the 256 x 256 tiles (conv_tc2.cu)
bool conv_tc2_supported(const ConvOp& op);
cudaError_t launch_conv_tc2(const ConvOp& op, const ConvTmaps& tm2, int num_sms, cudaStream_t stream);
cudaError_t conv_tc2_set_attributes();

}  // namespace rac
