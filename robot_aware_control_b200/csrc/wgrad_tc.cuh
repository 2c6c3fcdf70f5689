// Implicit-GEMM weight gradient on tcgen05 with MN-major operands read straight from the NHWC tape (wgrad_tc.cu).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace rac {

constexpr int kWgMaxSrc = 3;
constexpr int kWgMaxCTiles = 8;

struct WgradTmaps {
  CUtensorMap dy;            // (kpad, W, H, B, S) bf16: output gradient of every time step
  CUtensorMap x[kWgMaxSrc];  // (C_s, W, H, B, S) bf16: the concatenated inputs of the forward convolution
};

struct WgradGeom {
  int ks, pad;                 // square filter, 'same' padding
  int BH, NB, rows;            // positions per k-block = W * BH * NB (48 or 64)
  int hgroups, bgroups;        // H / BH, B / NB
  int kb_total;                // steps * bgroups * hgroups
  int splits, kb_per_split;    // split-K: grid.y slices of the contraction, summed afterwards in order
  int n_tiles;                 // ceil(kpad / 256) output-channel tiles
  int kpad;                    // packed output channels (rows of dWp)
  int taps, ctot;              // dWp row = taps * ctot floats
  int num_ctiles;              // input-channel tiles: (source, first channel, width in {64, 128, 192, 256})
  int ct_src[kWgMaxCTiles], ct_c0[kWgMaxCTiles], ct_w[kWgMaxCTiles];
  int src_coff[kWgMaxSrc];     // channel offset of source s inside ctot
  int src_tshift[kWgMaxSrc];   // 1: source s is read at step t - 1 (h_{t-1} of a ConvLSTM; step -1 = zeros)
  int max_stages;              // 0 = as many pipeline stages as the shared-memory ring holds (set by launch_wgrad_tc)
  int stage_a_boxes, stage_b_boxes;  // boxes per stage: the largest dY / X tile of this launch (set by launch_wgrad_tc)
  int producers;               // 1: one thread issues all TMA boxes; 2: warp 0 the dY boxes, warp 3 the X boxes; 3: same, one lane per box
  int epi_staged;              // 1: the epilogue stores through per-warp shared-memory transpose tiles (set by launch_wgrad_tc)
  int experiment;              // timing experiments only (results are garbage): 1 = no MMAs issued, 2 = no TMA loads issued
  float* out;                  // [splits][kpad][taps * ctot]
  long long out_split_stride;  // floats between two split partials
};

cudaError_t wgrad_tc_set_attributes();
cudaError_t launch_wgrad_tc(const WgradTmaps& tm, const WgradGeom& g, cudaStream_t s);
cudaError_t launch_wgrad_reduce(const float* part, int splits, long long n, long long split_stride, float* out,
                                cudaStream_t s);

}  // namespace rac
