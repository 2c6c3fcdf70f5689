// C ABI (include/racb200.h): handle, packed weights, activation workspace, TMA descriptors and the launch sequences
// for one SVG prediction step, the autoregressive rollout with fused planning cost, and the device-resident CEM loop.
#include "../../include/racb200.h"

#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <unordered_map>
#include <vector>

#include "conv.cuh"
#include "misc_kernels.cuh"

namespace rac {  // data_kernels.cu
cudaError_t launch_process_batch(const uint8_t* frames, const void* masks, int mask_u8, int B, int T, int Hs, int Ws,
                                 const rac_augment* aug, float* img_out, float* mask_out, cudaStream_t s);
cudaError_t launch_preprocess_states(const float* states, const float* actions, const rac_clip_calib* calib, int B, int T,
                                     int R, int A_in, int A_out, float* states_out, float* actions_out, cudaStream_t s);
}

namespace {

using namespace rac;
typedef __nv_bfloat16 bf16;

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

struct LayerSpec {
  int ks = 3;
  int ctot = 0;     // padded input channels per tap
  int n_packed = 0; // padded output columns
  int cout = 0;     // valid output columns (packed order)
  int block_n = 128;   // packing granularity of the output columns (pack.py); the launch tile may be wider
  bool first = false;  // fp32 first layer
};

struct Layer {
  void* w = nullptr;
  float* bias = nullptr;
  bool loaded = false;
};

struct Workspace {
  int B = 0;
  void* arena = nullptr;
  size_t arena_bytes = 0;
  float* img = nullptr;
  bf16 *a1 = nullptr, *cat5 = nullptr, *p1 = nullptr, *a2 = nullptr, *cat4 = nullptr, *p2 = nullptr, *a3a = nullptr,
       *a3b = nullptr, *cat3 = nullptr, *p3 = nullptr, *a4a = nullptr, *a4b = nullptr, *h4 = nullptr;
  bf16 *s1 = nullptr, *s2 = nullptr, *s3 = nullptr;  // keep_skip alternates (lazy)
  bf16 *aux = nullptr, *auxp = nullptr, *pin = nullptr, *postin = nullptr, *fin = nullptr, *z = nullptr,
       *zpost = nullptr;
  bf16* hs[3][2][2] = {};  // [prior, post, frame predictor][layer][ping-pong]
  float* cs[3][2] = {};
  bf16 *d2a = nullptr, *d2b = nullptr, *d3a = nullptr, *d3b = nullptr, *d4a = nullptr, *d5 = nullptr;
  float* cost_part = nullptr;
  float* goal4 = nullptr;
  // CEM scratch
  float *act2 = nullptr, *act5 = nullptr, *mean = nullptr, *stdv = nullptr;
  float *rob_states = nullptr, *rob_masks = nullptr;  // device-side robot model outputs of rac_cem_plan
  double* sum_cost = nullptr;
  int64_t* elite = nullptr;
  int cem_n = 0, cem_steps = 0;
  // ops
  ConvOp enc[2][10];     // [keep_skip][layer 1..9]
  bool enc_built[2] = {false, false};
  ConvOp in_conv[3];     // prior_in, post_in, fp_in
  ConvOp lstm[3][2][2];  // [which][layer][parity]
  ConvOp lstm_hh[3][2][2];  // lstm_group_norm: hh_gates convolution (lstm[][][] is then the ih_gates one)
  float *gn_ih = nullptr, *gn_hh = nullptr;  // lstm_group_norm: raw gate convolutions [B, 48, 4g] fp32
  float *gn_part = nullptr, *gn_o = nullptr; // fused GroupNorm statistics: per-tile partial sums, output-gate scratch
  ConvOp gauss[2][2];    // [prior, post][parity]
  ConvOp dec[10][2];     // [layer][parity] (only dec[0] depends on parity)
  std::unordered_map<std::string, std::pair<void*, std::pair<int64_t, int>>> named;
};

constexpr int kMaxGoals = 16;

}  // namespace

struct rac_handle {
  rac_config cfg;
  int device = 0;
  int num_sms = 148;
  int H = 48, W = 64, hl = 6, wl = 8;
  int enc_cin = 3;
  LayerSpec spec[RAC_L_COUNT_GN];
  Layer layer[RAC_L_COUNT_GN];
  float* gn_params[3][2] = {};  // lstm_group_norm: packed GroupNorm affine per cell [prior, post, fp][layer]
  int num_layers = RAC_L_COUNT;  // RAC_L_COUNT_GN with lstm_group_norm
  Workspace ws;
  int tile_m = 256;        // rows per CTA tile (RAC_TILE_M=128 selects the 128-row tiles, for A/B measurements)
  int cur[3] = {0, 0, 0};  // ping-pong index of the live hidden state per LSTM stack
  bool hidden_zero[3] = {false, false, false};  // h == 0 since init_hidden: the h_prev half of K is skipped
  int skip_zero_hidden = 1;  // RAC_SKIP_ZERO_H=0 disables the skip (A/B measurements)
  // RAC_2CTA bit 1: CTA-pair (cta_group::2) kernel for the 256-wide BN+LeakyReLU layers (default on: +2.5 % of the
  // rollout, their exposed epilogue is hidden by the second TMEM stage); bit 0: for the LSTM gate convolutions too
  // (default off: measured neutral to -1 % -- those kernels run at the power cap, hiding their epilogue only lowers
  // the clock of the main loop)
  int two_cta = 2;
  int gn_fuse_stats = 1;     // RAC_GN_FUSE=0: lstm_group_norm statistics by the 3-pass cell kernel instead of the gate-conv epilogue
  bool gn_fused = false;
  int enc_dedup = 1;         // RAC_ENC_DEDUP=0: run the encoder for every candidate at the first rollout step too
  int first_conv_tc = 1;     // RAC_FIRST_TC=0: encoder.c1.0 on the CUDA cores (fp32 inputs) instead of the tensor-core kernel
  int lstm_mc = 1;           // RAC_LSTM_MC=0: LSTM gate convs without the 2-CTA cluster that shares the activation tile by TMA multicast
  int y_major = 1;           // RAC_YMAJOR=0: candidate-major LSTM tiles (no per-sub-tile skipping of padding taps)
  int split_tail = 1;        // RAC_SPLIT_TAIL=0: no tail splitting in conv_tc_kernel (A/B measurements)
  int act_block_n = 256;     // RAC_ACT_BN=128: 256x128 tiles (double-buffered TMEM) for the BN+LeakyReLU layers (A/B measurements)
  int use_halo = 1;          // RAC_HALO=0: generic kernel for the 64-wide full-resolution layers too (A/B measurements)
  int halo_base_offset = 0;  // RAC_HALO_BASE_OFFSET=1 sets the descriptor base-offset field for the row-shifted operands: WRONG on
                             // B200 (measured: the 128B swizzle phase follows the absolute smem address bits [7:9])
  int halo_force_columns = 0;  // RAC_HALO_COLUMNS=1: per-column TMA loads even if the permuted-stride map encodes
  int c_tiled = 1;           // RAC_C_TILED=0: cell state in NHWC instead of the epilogue-private tiled layout (diagnosis)
  EncodeTiledFn encode = nullptr;
  int64_t launches = 0;
  char err[512] = {0};
  // live per-kernel timing (bench.py roofline): CUDA event pairs around every launch whose op name has this prefix
  std::string prof_prefix;
  std::vector<cudaEvent_t> prof_ev;
  size_t prof_used = 0;
  int64_t prof_dropped = 0;
  void* train = nullptr;  // TrainState (rac_train.inc.cu)
};

namespace {

int fail(rac_handle* h, int code, const char* fmt, ...) {
  if (h) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(h->err, sizeof(h->err), fmt, ap);
    va_end(ap);
  }
  return code;
}

#define CK(call)                                                                                   \
  do {                                                                                             \
    cudaError_t e_ = (call);                                                                       \
    if (e_ != cudaSuccess)                                                                         \
      return fail(h, RAC_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
  } while (0)
#define CKR(call)                  \
  do {                             \
    int r_ = (call);               \
    if (r_ != RAC_OK) return r_;   \
  } while (0)

inline int round_up(int a, int b) { return (a + b - 1) / b * b; }

void fill_specs(rac_handle* h) {
  const rac_config& c = h->cfg;
  const int g = c.g_dim;
  auto set = [&](int id, int ks, int ctot, int cout, int block_n) {
    LayerSpec& s = h->spec[id];
    s.ks = ks; s.ctot = ctot; s.cout = cout; s.block_n = block_n; s.n_packed = round_up(cout, block_n);
  };
  h->enc_cin = 3 + (c.use_mask ? 1 : 0) + ((c.use_mask && c.use_future_mask) ? 1 : 0);
  h->spec[RAC_L_ENC_C1_0].first = true;
  h->spec[RAC_L_ENC_C1_0].ks = 3; h->spec[RAC_L_ENC_C1_0].ctot = h->enc_cin;
  h->spec[RAC_L_ENC_C1_0].cout = 64; h->spec[RAC_L_ENC_C1_0].n_packed = 64; h->spec[RAC_L_ENC_C1_0].block_n = 64;
  set(RAC_L_ENC_C1_1, 3, 64, 64, 64);
  set(RAC_L_ENC_C2_0, 3, 64, 128, 128);
  set(RAC_L_ENC_C2_1, 3, 128, 128, 128);
  set(RAC_L_ENC_C3_0, 3, 128, 256, 128);
  set(RAC_L_ENC_C3_1, 3, 256, 256, 128);
  set(RAC_L_ENC_C3_2, 3, 256, 256, 128);
  set(RAC_L_ENC_C4_0, 3, 256, 512, 128);
  set(RAC_L_ENC_C4_1, 3, 512, 512, 128);
  set(RAC_L_ENC_C4_2, 3, 512, g, 128);
  set(RAC_L_PRIOR_IN, 3, 64 + g, g, 128);
  set(RAC_L_PRIOR_LSTM0, 5, 2 * g, 4 * g, 128);
  set(RAC_L_PRIOR_LSTM1, 3, 2 * g, 4 * g, 128);
  set(RAC_L_PRIOR_GAUSS, 3, g, 128, 128);
  set(RAC_L_FP_IN, 3, 64 + g + 64, g, 128);
  set(RAC_L_FP_LSTM0, 5, 2 * g, 4 * g, 128);
  set(RAC_L_FP_LSTM1, 3, 2 * g, 4 * g, 128);
  set(RAC_L_DEC_UPC2_0, 3, g, 512, 128);
  set(RAC_L_DEC_UPC2_1, 3, 512, 512, 128);
  set(RAC_L_DEC_UPC2_2, 3, 512, 256, 128);
  set(RAC_L_DEC_UPC3_0, 3, 512, 256, 128);
  set(RAC_L_DEC_UPC3_1, 3, 256, 256, 128);
  set(RAC_L_DEC_UPC3_2, 3, 256, 128, 128);
  set(RAC_L_DEC_UPC4_0, 3, 256, 128, 128);
  set(RAC_L_DEC_UPC4_1, 3, 128, 64, 64);
  set(RAC_L_DEC_UPC5_0, 3, 128, 64, 64);
  set(RAC_L_DEC_UPC5_1, 3, 64, 4, 16);
  set(RAC_L_POST_IN, 3, (c.use_robot_state ? 64 : 0) + g, g, 128);
  set(RAC_L_POST_LSTM0, 5, 2 * g, 4 * g, 128);
  set(RAC_L_POST_LSTM1, 3, 2 * g, 4 * g, 128);
  set(RAC_L_POST_GAUSS, 3, g, 128, 128);
  if (c.lstm_group_norm) {
    // NormConvLSTMCell (lstm.py:163-171): separate ih / hh gate convolutions, each g -> 4g
    const int ih[6] = {RAC_L_PRIOR_LSTM0, RAC_L_PRIOR_LSTM1, RAC_L_FP_LSTM0, RAC_L_FP_LSTM1, RAC_L_POST_LSTM0, RAC_L_POST_LSTM1};
    for (int i = 0; i < 6; ++i) {
      const int ks = (i & 1) ? 3 : 5;
      set(ih[i], ks, g, 4 * g, 128);
      set(RAC_L_PRIOR_LSTM0_HH + i, ks, g, 4 * g, 128);
    }
    h->num_layers = RAC_L_COUNT_GN;
  }
}

int64_t spec_w_elems(const LayerSpec& s) {
  if (s.first) return static_cast<int64_t>(9) * s.ctot * 64;
  return static_cast<int64_t>(s.n_packed) * s.ks * s.ks * s.ctot;
}

// ------------------------------------------------------------------ tensor maps
int encode_act_map(rac_handle* h, CUtensorMap* m, const bf16* ptr, int C, int B, int H, int W, int BH, int NB) {
  cuuint64_t dims[4] = {static_cast<cuuint64_t>(C), static_cast<cuuint64_t>(W), static_cast<cuuint64_t>(H),
                        static_cast<cuuint64_t>(B)};
  cuuint64_t strides[3] = {static_cast<cuuint64_t>(C) * 2, static_cast<cuuint64_t>(W) * C * 2,
                           static_cast<cuuint64_t>(H) * W * C * 2};
  cuuint32_t box[4] = {static_cast<cuuint32_t>(kBlockK), static_cast<cuuint32_t>(W), static_cast<cuuint32_t>(BH),
                       static_cast<cuuint32_t>(NB)};
  cuuint32_t es[4] = {1, 1, 1, 1};
  CUresult r = h->encode(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<bf16*>(ptr), dims, strides, box, es,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(h, RAC_ERR_CUDA, "cuTensorMapEncodeTiled(activation C=%d) failed: %d", C, (int)r);
  return RAC_OK;
}
int encode_w_map(rac_handle* h, CUtensorMap* m, const bf16* ptr, int K, int N, int block_n) {
  cuuint64_t dims[2] = {static_cast<cuuint64_t>(K), static_cast<cuuint64_t>(N)};
  cuuint64_t strides[1] = {static_cast<cuuint64_t>(K) * 2};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(kBlockK), static_cast<cuuint32_t>(block_n)};
  cuuint32_t es[2] = {1, 1};
  CUresult r = h->encode(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<bf16*>(ptr), dims, strides, box, es,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(h, RAC_ERR_CUDA, "cuTensorMapEncodeTiled(weights K=%d N=%d) failed: %d", K, N, (int)r);
  return RAC_OK;
}

// Forward-packed weights Wp[n][tap][c] as the MN-major B operand of a dgrad GEMM (EPI_F32_BT): 3-D view (c, tap, n),
// box {64 c, 1 tap, 64 n}; rows n >= n_packed (the K padding of the output-gradient operand) are out of bounds = zero
int encode_w_map_bt(rac_handle* h, CUtensorMap* m, const bf16* ptr, int ctot, int taps, int n_packed) {
  cuuint64_t dims[3] = {static_cast<cuuint64_t>(ctot), static_cast<cuuint64_t>(taps), static_cast<cuuint64_t>(n_packed)};
  cuuint64_t strides[2] = {static_cast<cuuint64_t>(ctot) * 2, static_cast<cuuint64_t>(taps) * ctot * 2};
  cuuint32_t box[3] = {64, 1, 64};
  cuuint32_t es[3] = {1, 1, 1};
  CUresult r = h->encode(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<bf16*>(ptr), dims, strides, box, es,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(h, RAC_ERR_CUDA, "cuTensorMapEncodeTiled(forward weights as dgrad operand) failed: %d", (int)r);
  return RAC_OK;
}

// Training weights in the k-block-major ("tiled") packing Wt[kb][n][64] (kb = tap * ctot / 64 + c / 64): dims
// (64, n_packed, kb_total). A forward box {64, block_n, 1} and a dgrad box {64 c, 64 n, 1} (MN-major operand, rows n
// beyond n_packed are out of bounds = zero) are both ONE contiguous chunk of memory. With the row-major packing
// Wp[n][tap * ctot + c] the 128-byte rows of a box lie taps * ctot * 2 bytes apart (a multiple of 2 KB for every layer),
// and the weight-streaming GEMMs of the training step (768 rows, the weights reused by 3 CTAs only) ran at 0.8 TB/s of
// DRAM reads with ~3 us per box (profiles/r02_dgrad5x5_ncu_s7.txt).
int encode_w_map_tiled(rac_handle* h, CUtensorMap* m, const bf16* ptr, int n_packed, int kb_total, int box_rows) {
  cuuint64_t dims[3] = {64, static_cast<cuuint64_t>(n_packed), static_cast<cuuint64_t>(kb_total)};
  cuuint64_t strides[2] = {128, static_cast<cuuint64_t>(n_packed) * 128};
  cuuint32_t box[3] = {64, static_cast<cuuint32_t>(box_rows), 1};
  cuuint32_t es[3] = {1, 1, 1};
  CUresult r = h->encode(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<bf16*>(ptr), dims, strides, box, es,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(h, RAC_ERR_CUDA, "cuTensorMapEncodeTiled(tiled weights n=%d kb=%d) failed: %d", n_packed, kb_total, (int)r);
  return RAC_OK;
}

// Halo kernel activation map. Preferred: dims (C, H, W, B) -- H and W swapped through the strides -- so that ONE box
// {64, 8, 34, 1} lands column-major (8 rows of a column = one swizzle atom). If the driver rejects the non-monotonic
// strides: dims (C, W, H, B) with a one-column box {64, 1, 8, 1}, loaded 34 times per tile.
int encode_halo_map(rac_handle* h, CUtensorMap* m, const bf16* ptr, int C, int B, int H, int W, int* column_loads) {
  cuuint32_t es[4] = {1, 1, 1, 1};
  if (!h->halo_force_columns) {
    cuuint64_t dims[4] = {static_cast<cuuint64_t>(C), static_cast<cuuint64_t>(H), static_cast<cuuint64_t>(W),
                          static_cast<cuuint64_t>(B)};
    cuuint64_t strides[3] = {static_cast<cuuint64_t>(W) * C * 2, static_cast<cuuint64_t>(C) * 2,
                             static_cast<cuuint64_t>(H) * W * C * 2};
    cuuint32_t box[4] = {static_cast<cuuint32_t>(kBlockK), 8, 34, 1};
    CUresult r = h->encode(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<bf16*>(ptr), dims, strides, box, es,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r == CUDA_SUCCESS) { *column_loads = 0; return RAC_OK; }
  }
  cuuint64_t dims[4] = {static_cast<cuuint64_t>(C), static_cast<cuuint64_t>(W), static_cast<cuuint64_t>(H),
                        static_cast<cuuint64_t>(B)};
  cuuint64_t strides[3] = {static_cast<cuuint64_t>(C) * 2, static_cast<cuuint64_t>(W) * C * 2,
                           static_cast<cuuint64_t>(H) * W * C * 2};
  cuuint32_t box[4] = {static_cast<cuuint32_t>(kBlockK), 1, 8, 1};
  CUresult r = h->encode(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<bf16*>(ptr), dims, strides, box, es,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(h, RAC_ERR_CUDA, "cuTensorMapEncodeTiled(halo column map C=%d) failed: %d", C, (int)r);
  *column_loads = 1;
  return RAC_OK;
}

// (C, W, B, H) view of an activation tensor: box {64, W, NB, BH} -> tile rows ordered (row of the map, candidate, column)
int encode_act_map_ymajor(rac_handle* h, CUtensorMap* m, const bf16* ptr, int C, int B, int H, int W, int BH, int NB) {
  cuuint64_t dims[4] = {static_cast<cuuint64_t>(C), static_cast<cuuint64_t>(W), static_cast<cuuint64_t>(B),
                        static_cast<cuuint64_t>(H)};
  cuuint64_t strides[3] = {static_cast<cuuint64_t>(C) * 2, static_cast<cuuint64_t>(H) * W * C * 2,
                           static_cast<cuuint64_t>(W) * C * 2};
  cuuint32_t box[4] = {static_cast<cuuint32_t>(kBlockK), static_cast<cuuint32_t>(W), static_cast<cuuint32_t>(NB),
                       static_cast<cuuint32_t>(BH)};
  cuuint32_t es[4] = {1, 1, 1, 1};
  CUresult r = h->encode(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<bf16*>(ptr), dims, strides, box, es,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(h, RAC_ERR_CUDA, "cuTensorMapEncodeTiled(y-major activation C=%d) failed: %d", C, (int)r);
  return RAC_OK;
}

struct Src {
  const bf16* p;
  int C;
  // training tape only (implicit-GEMM weight gradient over all time steps): the tensor this source is at step 0 of the
  // tape (nullptr: p itself), and 1 when the source is that tensor one step EARLIER (h_{t-1} of a ConvLSTM)
  const bf16* base0 = nullptr;
  int tshift = 0;
};

// (C, W, H, B, S) view of per-step NHWC tensors that lie `step_bytes` apart: box {64, W, BH, NB, 1}
int encode_act_map5(rac_handle* h, CUtensorMap* m, const bf16* ptr, int C, int B, int H, int W, int S,
                    unsigned long long step_bytes, int BH, int NB) {
  cuuint64_t dims[5] = {static_cast<cuuint64_t>(C), static_cast<cuuint64_t>(W), static_cast<cuuint64_t>(H),
                        static_cast<cuuint64_t>(B), static_cast<cuuint64_t>(S)};
  cuuint64_t strides[4] = {static_cast<cuuint64_t>(C) * 2, static_cast<cuuint64_t>(W) * C * 2,
                           static_cast<cuuint64_t>(H) * W * C * 2, static_cast<cuuint64_t>(step_bytes)};
  cuuint32_t box[5] = {static_cast<cuuint32_t>(kBlockK), static_cast<cuuint32_t>(W), static_cast<cuuint32_t>(BH),
                       static_cast<cuuint32_t>(NB), 1};
  cuuint32_t es[5] = {1, 1, 1, 1, 1};
  CUresult r = h->encode(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<bf16*>(ptr), dims, strides, box, es,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return fail(h, RAC_ERR_CUDA, "cuTensorMapEncodeTiled(5-D tape view C=%d S=%d stride=%llu) failed: %d", C, S, step_bytes, (int)r);
  return RAC_OK;
}

int ilog2(int v) {
  int s = 0;
  while ((1 << s) < v) ++s;
  return s;
}

// Build one convolution op: geometry, TMA descriptors, raw pointers, constant epilogue fields.
int make_conv(rac_handle* h, ConvOp* op, const char* name, int layer, int H, int W, std::vector<Src> srcs, int epi) {
  const LayerSpec& s = h->spec[layer];
  const int B = h->ws.B;
  memset(op, 0, sizeof(*op));
  op->name = name;
  op->block_m = h->tile_m;
  op->block_n = s.block_n;
  if (h->tile_m == 256 && s.n_packed % 256 == 0 && (epi == EPI_ACT || epi == EPI_LSTM || epi == EPI_F32 || epi == EPI_GATES)) op->block_n = 256;
  if (epi == EPI_ACT && h->act_block_n == 128 && op->block_n == 256) op->block_n = 128;
  op->epi = epi;
  ConvGeom& g = op->g;
  g.B = B; g.H = H; g.W = W; g.ks = s.ks; g.pad = s.ks / 2;
  const bool big = h->tile_m == 256;
  switch (W) {  // TMA box {64 ch, W, BH, NB} with W * BH * NB == tile_m
    case 64: g.BH = big ? 4 : 2; g.NB = 1; break;
    case 32: g.BH = big ? 8 : 4; g.NB = 1; break;
    case 16: g.BH = 4; g.NB = big ? 4 : 2; break;
    case 8: g.BH = 2; g.NB = big ? 16 : 8; break;
    default: return fail(h, RAC_ERR_INVALID, "unsupported feature-map width %d", W);
  }
  if (H % g.BH != 0) return fail(h, RAC_ERR_INVALID, "feature-map height %d not divisible by tile rows %d", H, g.BH);
  g.nsrc = static_cast<int>(srcs.size());
  g.ctot = 0;
  // ConvLSTM gate convolutions on the 6x8 maps: one map row per 128-row MMA sub-tile (see ConvGeom::y_major)
  g.nbw_shift = ilog2(g.NB * W);
  g.y_major = (h->y_major && h->cfg.conv_impl == 0 && epi == EPI_LSTM && big && g.BH == 2 && g.NB * W == 128 &&
               !(h->two_cta & 1)) ? 1 : 0;
  for (int i = 0; i < g.nsrc; ++i) {
    if (srcs[i].C % kBlockK != 0) return fail(h, RAC_ERR_INVALID, "%s: source channels %d not a multiple of 64", name, srcs[i].C);
    g.src_kb[i] = srcs[i].C / kBlockK;
    g.ctot += srcs[i].C;
    op->raw.src[i] = srcs[i].p;
    if (g.y_major) CKR(encode_act_map_ymajor(h, &op->tm.a[i], srcs[i].p, srcs[i].C, B, H, W, g.BH, g.NB));
    else CKR(encode_act_map(h, &op->tm.a[i], srcs[i].p, srcs[i].C, B, H, W, g.BH, g.NB));
  }
  if (g.ctot != s.ctot) return fail(h, RAC_ERR_INVALID, "%s: channel mismatch %d vs packed %d", name, g.ctot, s.ctot);
  g.tiles_per_img = H / g.BH;
  g.num_m_tiles = ((B + g.NB - 1) / g.NB) * g.tiles_per_img;
  g.num_n_tiles = s.n_packed / op->block_n;
  g.w_shift = ilog2(W);
  g.bhw_shift = ilog2(g.BH * W);
  g.no_split_tail = h->split_tail ? 0 : 1;
  const bf16* wp = static_cast<const bf16*>(h->layer[layer].w);
  op->raw.w = wp;
  CKR(encode_w_map(h, &op->tm.w, wp, s.ks * s.ks * s.ctot, s.n_packed, op->block_n));
  op->e.bias = h->layer[layer].bias;
  op->e.cout = s.cout;
  op->e.cost_nparts = H * W / 32;
  if (h->cfg.conv_impl == 0 && conv_tc2_supported(*op) &&
      (((h->two_cta & 1) && epi == EPI_LSTM) || ((h->two_cta & 2) && epi == EPI_ACT))) {
    for (int i = 0; i < g.nsrc; ++i)
      CKR(encode_act_map(h, &op->tm2.a[i], srcs[i].p, srcs[i].C, B, H, W, g.BH, g.NB / 2));
    CKR(encode_w_map(h, &op->tm2.w, wp, s.ks * s.ks * s.ctot, s.n_packed, 128));
    op->two_cta = 1;
  }
  if (h->lstm_mc && h->cfg.conv_impl == 0 && conv_tc_mc_supported(*op)) {
    for (int i = 0; i < g.nsrc; ++i)
      CKR(encode_act_map_ymajor(h, &op->tm_mc.a[i], srcs[i].p, srcs[i].C, B, H, W, 1, g.NB));
    op->tm_mc.w = op->tm.w;
    op->mc = 1;
  }
  if (h->use_halo && h->cfg.conv_impl == 0 && h->tile_m == 256 && conv_halo_supported(*op)) {
    CKR(encode_halo_map(h, &op->tm_halo, srcs[0].p, srcs[0].C, B, H, W, &op->halo_column_loads));
    op->halo = 1;
    op->e.cost_nparts = 128;
  }
  return RAC_OK;
}

void set_act(ConvOp* op, bf16* out, int cstride, int coff, int upsample, int lrelu) {
  op->e.out = out; op->e.out_cstride = cstride; op->e.out_coff = coff; op->e.upsample = upsample; op->e.lrelu = lrelu;
}

// live timing of the small non-GEMM kernels through the same hook (rac_profile_begin with their name)
struct ProfScope {
  rac_handle* h;
  cudaStream_t st;
  bool timed = false;
  ProfScope(rac_handle* h_, const char* name, cudaStream_t st_) : h(h_), st(st_) {
    if (!h->prof_prefix.empty() && strstr(name, h->prof_prefix.c_str()) != nullptr) {
      if (h->prof_used + 2 <= h->prof_ev.size()) {
        timed = true;
        cudaEventRecord(h->prof_ev[h->prof_used], st);
      } else {
        h->prof_dropped++;
      }
    }
  }
  ~ProfScope() {
    if (timed) {
      cudaEventRecord(h->prof_ev[h->prof_used + 1], st);
      h->prof_used += 2;
    }
  }
};

int launch(rac_handle* h, const ConvOp& op, cudaStream_t st) {
  bool timed = false;
  if (!h->prof_prefix.empty() && strstr(op.name, h->prof_prefix.c_str()) != nullptr) {
    if (h->prof_used + 2 <= h->prof_ev.size()) {
      timed = true;
      cudaEventRecord(h->prof_ev[h->prof_used], st);
    } else {
      h->prof_dropped++;
    }
  }
  cudaError_t e = h->cfg.conv_impl == 1 ? launch_conv_simt(op, st)
                  : op.mc ? launch_conv_tc_mc(op, op.tm_mc, h->num_sms, st)
                  : op.two_cta ? launch_conv_tc2(op, op.tm2, h->num_sms, st)
                  : op.halo ? launch_conv_halo(op, op.tm_halo, op.halo_column_loads, h->halo_base_offset, h->num_sms, st)
                            : launch_conv_tc(op, h->num_sms, st);
  if (timed) {
    cudaEventRecord(h->prof_ev[h->prof_used + 1], st);
    h->prof_used += 2;
  }
  if (e != cudaSuccess) return fail(h, RAC_ERR_CUDA, "launch of conv '%s' failed: %s", op.name, cudaGetErrorString(e));
  h->launches++;
  return RAC_OK;
}

// ------------------------------------------------------------------ workspace
struct Bump {
  size_t off = 0;
  char* base = nullptr;
  template <typename T>
  T* take(size_t n) {
    off = (off + 1023) / 1024 * 1024;
    T* p = base ? reinterpret_cast<T*>(base + off) : nullptr;
    off += n * sizeof(T);
    return p;
  }
};

void carve(rac_handle* h, Bump& bp, int B) {
  Workspace& w = h->ws;
  const size_t n = static_cast<size_t>(B);
  const int g = h->cfg.g_dim;
  const size_t P0 = 48 * 64, P1 = 24 * 32, P2 = 12 * 16, P3 = 6 * 8;
  w.img = bp.take<float>(n * P0 * 4);
  w.a1 = bp.take<bf16>(n * P0 * 64);
  w.cat5 = bp.take<bf16>(n * P0 * 128);
  w.p1 = bp.take<bf16>(n * P1 * 64);
  w.a2 = bp.take<bf16>(n * P1 * 128);
  w.cat4 = bp.take<bf16>(n * P1 * 256);
  w.p2 = bp.take<bf16>(n * P2 * 128);
  w.a3a = bp.take<bf16>(n * P2 * 256);
  w.a3b = bp.take<bf16>(n * P2 * 256);
  w.cat3 = bp.take<bf16>(n * P2 * 512);
  w.p3 = bp.take<bf16>(n * P3 * 256);
  w.a4a = bp.take<bf16>(n * P3 * 512);
  w.a4b = bp.take<bf16>(n * P3 * 512);
  w.h4 = bp.take<bf16>(n * P3 * g);
  w.aux = bp.take<bf16>(n * P3 * 64);
  w.auxp = bp.take<bf16>(n * P3 * 64);
  w.pin = bp.take<bf16>(n * P3 * g);
  w.postin = bp.take<bf16>(n * P3 * g);
  w.fin = bp.take<bf16>(n * P3 * g);
  w.z = bp.take<bf16>(n * P3 * 64);
  w.zpost = bp.take<bf16>(n * P3 * 64);
  for (int l = 0; l < 3; ++l)
    for (int k = 0; k < 2; ++k) {
      for (int p = 0; p < 2; ++p) w.hs[l][k][p] = bp.take<bf16>(n * P3 * g);
      w.cs[l][k] = bp.take<float>(((n + 15) / 16 * 16) * P3 * g);  // whole latent tiles: the tiled layout pads the batch
    }
  w.d2a = bp.take<bf16>(n * P3 * 512);
  w.d2b = bp.take<bf16>(n * P3 * 512);
  w.d3a = bp.take<bf16>(n * P2 * 256);
  w.d3b = bp.take<bf16>(n * P2 * 256);
  w.d4a = bp.take<bf16>(n * P1 * 128);
  w.d5 = bp.take<bf16>(n * P0 * 64);
  if (h->cfg.lstm_group_norm) {
    w.gn_ih = bp.take<float>(n * P3 * 4 * g);
    w.gn_hh = bp.take<float>(n * P3 * 4 * g);
    w.gn_part = bp.take<float>(((n + 15) / 16 * 16) * 2 * 16 * 6 * 2);
    w.gn_o = bp.take<float>(n * P3 * g);
  }
  w.cost_part = bp.take<float>(n * 128 * 2);
  w.goal4 = bp.take<float>(static_cast<size_t>(kMaxGoals) * P0 * 4);
}

int build_enc_ops(rac_handle* h, int ks) {
  Workspace& w = h->ws;
  ConvOp* e = w.enc[ks];
  const int g = h->cfg.g_dim;
  // h1/h2/h3 normally land in the second channel half of the decoder concat buffers (torch.cat([up, skip]),
  // vgg_64.py:236-240); with keep_skip they go to side buffers so that the held skip tensors survive.
  bf16* o1 = ks ? w.s1 : w.cat5; const int c1s = ks ? 64 : 128, c1o = ks ? 0 : 64;
  bf16* o2 = ks ? w.s2 : w.cat4; const int c2s = ks ? 128 : 256, c2o = ks ? 0 : 128;
  bf16* o3 = ks ? w.s3 : w.cat3; const int c3s = ks ? 256 : 512, c3o = ks ? 0 : 256;
  CKR(make_conv(h, &e[1], "encoder.c1.1", RAC_L_ENC_C1_1, 48, 64, {{w.a1, 64}}, EPI_ACT)); set_act(&e[1], o1, c1s, c1o, 0, 1);
  CKR(make_conv(h, &e[2], "encoder.c2.0", RAC_L_ENC_C2_0, 24, 32, {{w.p1, 64}}, EPI_ACT)); set_act(&e[2], w.a2, 128, 0, 0, 1);
  CKR(make_conv(h, &e[3], "encoder.c2.1", RAC_L_ENC_C2_1, 24, 32, {{w.a2, 128}}, EPI_ACT)); set_act(&e[3], o2, c2s, c2o, 0, 1);
  CKR(make_conv(h, &e[4], "encoder.c3.0", RAC_L_ENC_C3_0, 12, 16, {{w.p2, 128}}, EPI_ACT)); set_act(&e[4], w.a3a, 256, 0, 0, 1);
  CKR(make_conv(h, &e[5], "encoder.c3.1", RAC_L_ENC_C3_1, 12, 16, {{w.a3a, 256}}, EPI_ACT)); set_act(&e[5], w.a3b, 256, 0, 0, 1);
  CKR(make_conv(h, &e[6], "encoder.c3.2", RAC_L_ENC_C3_2, 12, 16, {{w.a3b, 256}}, EPI_ACT)); set_act(&e[6], o3, c3s, c3o, 0, 1);
  CKR(make_conv(h, &e[7], "encoder.c4.0", RAC_L_ENC_C4_0, 6, 8, {{w.p3, 256}}, EPI_ACT)); set_act(&e[7], w.a4a, 512, 0, 0, 1);
  CKR(make_conv(h, &e[8], "encoder.c4.1", RAC_L_ENC_C4_1, 6, 8, {{w.a4a, 512}}, EPI_ACT)); set_act(&e[8], w.a4b, 512, 0, 0, 1);
  CKR(make_conv(h, &e[9], "encoder.c4.2", RAC_L_ENC_C4_2, 6, 8, {{w.a4b, 512}}, EPI_ACT)); set_act(&e[9], w.h4, g, 0, 0, 1);
  w.enc_built[ks] = true;
  return RAC_OK;
}

int build_ops(rac_handle* h) {
  Workspace& w = h->ws;
  const rac_config& c = h->cfg;
  const int g = c.g_dim;
  CKR(build_enc_ops(h, 0));
  // input convolutions (dynamics.py:496-498,512-513): tiled action/state channels arrive as the 64-channel `aux` block
  CKR(make_conv(h, &w.in_conv[0], "prior_input_conv", RAC_L_PRIOR_IN, 6, 8, {{w.aux, 64}, {w.h4, g}}, EPI_ACT));
  set_act(&w.in_conv[0], w.pin, g, 0, 0, 0);
  if (c.use_robot_state) {
    CKR(make_conv(h, &w.in_conv[1], "posterior_input_conv", RAC_L_POST_IN, 6, 8, {{w.auxp, 64}, {w.h4, g}}, EPI_ACT));
  } else {
    CKR(make_conv(h, &w.in_conv[1], "posterior_input_conv", RAC_L_POST_IN, 6, 8, {{w.h4, g}}, EPI_ACT));
  }
  set_act(&w.in_conv[1], w.postin, g, 0, 0, 0);
  CKR(make_conv(h, &w.in_conv[2], "frame_pred_input_conv", RAC_L_FP_IN, 6, 8, {{w.aux, 64}, {w.h4, g}, {w.z, 64}}, EPI_ACT));
  set_act(&w.in_conv[2], w.fin, g, 0, 0, 0);
  const int l0[3] = {RAC_L_PRIOR_LSTM0, RAC_L_POST_LSTM0, RAC_L_FP_LSTM0};
  const int l1[3] = {RAC_L_PRIOR_LSTM1, RAC_L_POST_LSTM1, RAC_L_FP_LSTM1};
  bf16* xin[3] = {w.pin, w.postin, w.fin};
  const char* n0[3] = {"prior.lstm.0", "posterior.lstm.0", "frame_predictor.lstm.0"};
  const char* n1[3] = {"prior.lstm.1", "posterior.lstm.1", "frame_predictor.lstm.1"};
  if (c.lstm_group_norm) {
    const int hh0[3] = {RAC_L_PRIOR_LSTM0_HH, RAC_L_POST_LSTM0_HH, RAC_L_FP_LSTM0_HH};
    // GroupNorm sums fused into the gate-conv epilogue when a 256-column tile lies inside one GroupNorm quarter
    h->gn_fused = h->gn_fuse_stats && c.conv_impl == 0 && h->tile_m == 256 && (g == 256 || g == 512);
    const int gates_epi = h->gn_fused ? EPI_GATES : EPI_F32;
    auto f32_out = [&](ConvOp* op, float* dst, int tensor) {
      op->e.nseg = 1;
      op->e.seg[0] = F32Seg{0, 4 * g, dst, 4 * g, 0, 0};
      op->e.gn_part = w.gn_part;
      op->e.gn_tensor = tensor;
      op->e.gn_ntq = g / 256;  // 256-column tiles (64 channels) per GroupNorm quarter (g / 4 channels)
    };
    for (int l = 0; l < 3; ++l)
      for (int p = 0; p < 2; ++p) {
        CKR(make_conv(h, &w.lstm[l][0][p], n0[l], l0[l], 6, 8, {{xin[l], g}}, gates_epi));
        CKR(make_conv(h, &w.lstm_hh[l][0][p], n0[l], hh0[l], 6, 8, {{w.hs[l][0][p], g}}, gates_epi));
        CKR(make_conv(h, &w.lstm[l][1][p], n1[l], l1[l], 6, 8, {{w.hs[l][0][p ^ 1], g}}, gates_epi));
        CKR(make_conv(h, &w.lstm_hh[l][1][p], n1[l], hh0[l] + 1, 6, 8, {{w.hs[l][1][p], g}}, gates_epi));
        for (int k = 0; k < 2; ++k) {
          f32_out(&w.lstm[l][k][p], w.gn_ih, 0);
          f32_out(&w.lstm_hh[l][k][p], w.gn_hh, 1);
        }
      }
  }
  for (int l = 0; l < 3 && !c.lstm_group_norm; ++l)
    for (int p = 0; p < 2; ++p) {
      ConvOp* a = &w.lstm[l][0][p];
      CKR(make_conv(h, a, n0[l], l0[l], 6, 8, {{xin[l], g}, {w.hs[l][0][p], g}}, EPI_LSTM));
      a->e.c_state = w.cs[l][0]; a->e.h_out = w.hs[l][0][p ^ 1]; a->e.hid = g; a->e.c_tiled = h->c_tiled;
      ConvOp* b = &w.lstm[l][1][p];
      CKR(make_conv(h, b, n1[l], l1[l], 6, 8, {{w.hs[l][0][p ^ 1], g}, {w.hs[l][1][p], g}}, EPI_LSTM));
      b->e.c_state = w.cs[l][1]; b->e.h_out = w.hs[l][1][p ^ 1]; b->e.hid = g; b->e.c_tiled = h->c_tiled;
    }
  for (int p = 0; p < 2; ++p) {
    CKR(make_conv(h, &w.gauss[0][p], "prior.mu_net|logvar_net", RAC_L_PRIOR_GAUSS, 6, 8, {{w.hs[0][1][p ^ 1], g}}, EPI_GAUSS));
    CKR(make_conv(h, &w.gauss[1][p], "posterior.mu_net|logvar_net", RAC_L_POST_GAUSS, 6, 8, {{w.hs[1][1][p ^ 1], g}}, EPI_GAUSS));
    for (int q = 0; q < 2; ++q) { w.gauss[q][p].e.z_dim = c.z_dim; w.gauss[q][p].e.z_out = w.z; }
    CKR(make_conv(h, &w.dec[0][p], "decoder.upc2.0", RAC_L_DEC_UPC2_0, 6, 8, {{w.hs[2][1][p ^ 1], g}}, EPI_ACT));
    set_act(&w.dec[0][p], w.d2a, 512, 0, 0, 1);
  }
  ConvOp(*d)[2] = w.dec;
  CKR(make_conv(h, &d[1][0], "decoder.upc2.1", RAC_L_DEC_UPC2_1, 6, 8, {{w.d2a, 512}}, EPI_ACT)); set_act(&d[1][0], w.d2b, 512, 0, 0, 1);
  CKR(make_conv(h, &d[2][0], "decoder.upc2.2", RAC_L_DEC_UPC2_2, 6, 8, {{w.d2b, 512}}, EPI_ACT)); set_act(&d[2][0], w.cat3, 512, 0, 1, 1);
  CKR(make_conv(h, &d[3][0], "decoder.upc3.0", RAC_L_DEC_UPC3_0, 12, 16, {{w.cat3, 512}}, EPI_ACT)); set_act(&d[3][0], w.d3a, 256, 0, 0, 1);
  CKR(make_conv(h, &d[4][0], "decoder.upc3.1", RAC_L_DEC_UPC3_1, 12, 16, {{w.d3a, 256}}, EPI_ACT)); set_act(&d[4][0], w.d3b, 256, 0, 0, 1);
  CKR(make_conv(h, &d[5][0], "decoder.upc3.2", RAC_L_DEC_UPC3_2, 12, 16, {{w.d3b, 256}}, EPI_ACT)); set_act(&d[5][0], w.cat4, 256, 0, 1, 1);
  CKR(make_conv(h, &d[6][0], "decoder.upc4.0", RAC_L_DEC_UPC4_0, 24, 32, {{w.cat4, 256}}, EPI_ACT)); set_act(&d[6][0], w.d4a, 128, 0, 0, 1);
  CKR(make_conv(h, &d[7][0], "decoder.upc4.1", RAC_L_DEC_UPC4_1, 24, 32, {{w.d4a, 128}}, EPI_ACT)); set_act(&d[7][0], w.cat5, 128, 0, 1, 1);
  CKR(make_conv(h, &d[8][0], "decoder.upc5.0", RAC_L_DEC_UPC5_0, 48, 64, {{w.cat5, 128}}, EPI_ACT)); set_act(&d[8][0], w.d5, 64, 0, 0, 1);
  CKR(make_conv(h, &d[9][0], "decoder.upc5.1", RAC_L_DEC_UPC5_1, 48, 64, {{w.d5, 64}}, EPI_FRAME));
  return RAC_OK;
}

void name_buffers(rac_handle* h) {
  Workspace& w = h->ws;
  const size_t n = static_cast<size_t>(w.B);
  const int g = h->cfg.g_dim;
  auto put = [&](const char* k, void* p, size_t elems, int eb) { w.named[k] = {p, {static_cast<int64_t>(elems), eb}}; };
  put("img", w.img, n * 3072 * 4, 4);
  put("a1", w.a1, n * 3072 * 64, 2);
  put("cat5", w.cat5, n * 3072 * 128, 2);
  put("cat4", w.cat4, n * 768 * 256, 2);
  put("cat3", w.cat3, n * 192 * 512, 2);
  put("p1", w.p1, n * 768 * 64, 2);
  put("a2", w.a2, n * 768 * 128, 2);
  put("h4", w.h4, n * 48 * g, 2);
  put("aux", w.aux, n * 48 * 64, 2);
  put("prior_in", w.pin, n * 48 * g, 2);
  put("post_in", w.postin, n * 48 * g, 2);
  put("frame_in", w.fin, n * 48 * g, 2);
  put("z", w.z, n * 48 * 64, 2);
  put("d2a", w.d2a, n * 48 * 512, 2);
  put("d2b", w.d2b, n * 48 * 512, 2);
  put("d3a", w.d3a, n * 192 * 256, 2);
  put("d4a", w.d4a, n * 768 * 128, 2);
  put("d5", w.d5, n * 3072 * 64, 2);
  put("cost_part", w.cost_part, n * 128 * 2, 4);
  const char* hn[3] = {"prior", "post", "fp"};
  for (int l = 0; l < 3; ++l)
    for (int k = 0; k < 2; ++k) {
      for (int p = 0; p < 2; ++p) {
        char key[64];
        snprintf(key, sizeof(key), "%s.h%d.%d", hn[l], k, p);
        put(key, w.hs[l][k][p], n * 48 * g, 2);
      }
      char key[64];
      snprintf(key, sizeof(key), "%s.c%d", hn[l], k);
      put(key, w.cs[l][k], n * 48 * g, 4);
    }
}

void free_ws(rac_handle* h) {
  Workspace& w = h->ws;
  if (w.arena) cudaFree(w.arena);
  if (w.s1) cudaFree(w.s1);
  if (w.s2) cudaFree(w.s2);
  if (w.s3) cudaFree(w.s3);
  if (w.act2) cudaFree(w.act2);
  if (w.act5) cudaFree(w.act5);
  if (w.mean) cudaFree(w.mean);
  if (w.sum_cost) cudaFree(w.sum_cost);
  if (w.elite) cudaFree(w.elite);
  if (w.rob_states) cudaFree(w.rob_states);
  if (w.rob_masks) cudaFree(w.rob_masks);
  w = Workspace();
}

int ensure_keep_skip(rac_handle* h) {
  Workspace& w = h->ws;
  if (w.enc_built[1]) return RAC_OK;
  const size_t n = static_cast<size_t>(w.B);
  CK(cudaMalloc(&w.s1, n * 3072 * 64 * 2));
  CK(cudaMalloc(&w.s2, n * 768 * 128 * 2));
  CK(cudaMalloc(&w.s3, n * 192 * 256 * 2));
  return build_enc_ops(h, 1);
}

struct StepArgs {
  int n;
  const float* mask_a;  // mask_t      (n,H,W) or null
  const float* mask_b;  // mask_{t+1}  (n,H,W) or null
  long long mask_bstride;  // floats between the mask planes of consecutive candidates
  const float* robot;
  const float* robot_next;
  const float* action;
  int action_stride;
  const float* eps;
  unsigned long long seed;
  unsigned int noise_ctr;
  int cand_offset;
  int sample_mean;
  int use_posterior;
  const float* next_robot;
  const float* eps_post;
  int force_use_prior;
  int keep_skip;
  float *mu_p, *logvar_p, *mu, *logvar;
  // frame epilogue
  const float* curr_img;  // composite source (null: raw x_pred only)
  float* next_img;
  const float* mask_next;
  const float* goal_img;
  const float* goal_mask;
  float* xpred_out;
  float* cost_part;
  int zero_robot, dontcare;
  // 1: every candidate's input frame is the same image (first step of a rollout from one start frame): with an
  // image-only encoder input the encoder runs for one tile group of candidates and its outputs are broadcast
  int shared_frame;
};

// ConvLSTM stack l (two cells). Right after init_hidden h_prev is all zero, so its half of the K loop is skipped (exact).
int launch_lstm(rac_handle* h, int l, cudaStream_t st) {
  Workspace& w = h->ws;
  const int p = h->cur[l];
  if (h->cfg.lstm_group_norm) {
    // NormConvLSTMCell (lstm.py:177-198): ih / hh gate convolutions -> GroupNorm + cell kernel, twice
    const int g = h->cfg.g_dim;
    for (int k = 0; k < 2; ++k) {
      CKR(launch(h, w.lstm[l][k][p], st));
      CKR(launch(h, w.lstm_hh[l][k][p], st));
      {
        ProfScope ps(h, "norm_lstm_cell", st);
        if (h->gn_fused)
          CK(launch_norm_lstm_cell_fused(w.gn_ih, w.gn_hh, w.gn_part, 3 * (g / 256), h->gn_params[l][k], w.cs[l][k],
                                         w.gn_o, w.hs[l][k][p ^ 1], w.B, 48, g, st));
        else
          CK(launch_norm_lstm_cell(w.gn_ih, w.gn_hh, h->gn_params[l][k], w.cs[l][k], w.hs[l][k][p ^ 1], w.B, 48, g, st));
      }
      h->launches++;
    }
    h->hidden_zero[l] = false;
    return RAC_OK;
  }
  ConvOp a = w.lstm[l][0][p], b = w.lstm[l][1][p];
  if (h->hidden_zero[l] && h->skip_zero_hidden) a.g.src_dead[1] = b.g.src_dead[1] = 1;
  CKR(launch(h, a, st));
  CKR(launch(h, b, st));
  h->hidden_zero[l] = false;
  return RAC_OK;
}

// One SVGConvModel.forward (dynamics.py:544-644) + compositing / cost epilogue (trajectory_sampler.py:148-168).
int run_step(rac_handle* h, const StepArgs& a, cudaStream_t st) {
  Workspace& w = h->ws;
  const rac_config& c = h->cfg;
  const int B = w.B;
  const int ks = a.keep_skip ? 1 : 0;
  if (ks) CKR(ensure_keep_skip(h));
  // ---- encoder (vgg_64.py:122-129)
  // First step of a rollout: all B candidates start from the SAME frame, and without mask channels the encoder sees
  // nothing else -- its outputs (h and the three skips) are the same 737 KB for every candidate. It then runs for one
  // tile group of kSharedBatch candidates (the prebuilt ops with a smaller candidate count: same tensor maps, fewer
  // tiles) and candidate 0's outputs are copied to the others: exact (every output row is computed from identical
  // inputs by the same instructions), 1.7 of 18.3 GFLOP per frame less at that step. RAC_ENC_DEDUP=0 switches it off.
  constexpr int kSharedBatch = 16;
  const bool shared = a.shared_frame && h->enc_dedup && !c.use_mask && c.conv_impl == 0 && B > kSharedBatch;
  const int Be = shared ? kSharedBatch : B;
  {
    ProfScope ps(h, "encoder.c1.0", st);
    if (h->first_conv_tc && c.conv_impl == 0)
      CK(launch_first_conv_tc(w.img, c.use_mask ? a.mask_a : nullptr, (c.use_mask && c.use_future_mask) ? a.mask_b : nullptr,
                              a.mask_bstride, static_cast<const float*>(h->layer[RAC_L_ENC_C1_0].w), h->layer[RAC_L_ENC_C1_0].bias,
                              w.a1, Be, 48, 64, h->enc_cin, h->num_sms, st));
    else
      CK(launch_first_conv(w.img, c.use_mask ? a.mask_a : nullptr, (c.use_mask && c.use_future_mask) ? a.mask_b : nullptr,
                           a.mask_bstride, static_cast<const float*>(h->layer[RAC_L_ENC_C1_0].w), h->layer[RAC_L_ENC_C1_0].bias, w.a1, Be,
                           48, 64, h->enc_cin, st));
  }
  h->launches++;
  ConvOp* e = w.enc[ks];
  auto enc = [&](int i) -> int {
    if (!shared) return launch(h, e[i], st);
    ConvOp op = e[i];
    op.g.B = Be;
    op.g.num_m_tiles = ((Be + op.g.NB - 1) / op.g.NB) * op.g.tiles_per_img;
    return launch(h, op, st);
  };
  auto spread = [&](int i) -> int {  // candidate 0's output of layer i -> candidates Be .. B-1
    if (!shared) return RAC_OK;
    const ConvOp& op = e[i];
    CK(launch_broadcast_candidate(op.e.out, op.g.H * op.g.W, op.e.out_cstride, op.e.out_coff, op.e.cout, Be, B, st));
    h->launches++;
    return RAC_OK;
  };
  CKR(enc(1));
  {
    ProfScope ps(h, "maxpool.1", st);
    CK(launch_maxpool2(e[1].e.out, e[1].e.out_cstride, e[1].e.out_coff, w.p1, Be, 48, 64, 64, st));
  }
  CKR(spread(1));
  CKR(enc(2));
  CKR(enc(3));
  {
    ProfScope ps(h, "maxpool.2", st);
    CK(launch_maxpool2(e[3].e.out, e[3].e.out_cstride, e[3].e.out_coff, w.p2, Be, 24, 32, 128, st));
  }
  CKR(spread(3));
  CKR(enc(4));
  CKR(enc(5));
  CKR(enc(6));
  {
    ProfScope ps(h, "maxpool.3", st);
    CK(launch_maxpool2(e[6].e.out, e[6].e.out_cstride, e[6].e.out_coff, w.p3, Be, 12, 16, 256, st));
  }
  CKR(spread(6));
  CKR(enc(7));
  CKR(enc(8));
  CKR(enc(9));
  CKR(spread(9));
  // ---- tiled action / state channels (dynamics.py:591-603)
  CK(launch_aux_tile(a.action, a.action_stride, c.action_dim, c.use_robot_state ? a.robot : nullptr,
                     (c.use_robot_state && c.use_future_robot_state) ? a.robot_next : nullptr, c.robot_dim, w.aux, B,
                     48, st));
  h->launches += 4;
  // ---- learned prior (dynamics.py:594-610)
  CKR(launch(h, w.in_conv[0], st));
  {
    const int p = h->cur[0];
    CKR(launch_lstm(h, 0, st));
    ConvOp gop = w.gauss[0][p];
    gop.e.eps = a.eps; gop.e.mu_out = a.mu_p; gop.e.logvar_out = a.logvar_p; gop.e.sample_mean = a.sample_mean;
    gop.e.seed = a.seed; gop.e.noise_ctr = a.noise_ctr; gop.e.cand_offset = a.cand_offset; gop.e.z_out = w.z;
    CKR(launch(h, gop, st));
    h->cur[0] ^= 1;
  }
  // ---- posterior (dynamics.py:613-629); h_target re-encodes the CURRENT frame (reference quirk, :619) == h4
  if (a.use_posterior) {
    if (c.use_robot_state) {
      CK(launch_aux_tile(nullptr, 0, 0, a.next_robot, nullptr, c.robot_dim, w.auxp, B, 48, st));
      h->launches++;
    }
    CKR(launch(h, w.in_conv[1], st));
    const int p = h->cur[1];
    CKR(launch_lstm(h, 1, st));
    ConvOp gop = w.gauss[1][p];
    gop.e.eps = a.eps_post; gop.e.mu_out = a.mu; gop.e.logvar_out = a.logvar; gop.e.sample_mean = 0;
    gop.e.seed = a.seed ^ 0x9e3779b97f4a7c15ull; gop.e.noise_ctr = a.noise_ctr; gop.e.cand_offset = a.cand_offset;
    gop.e.z_out = a.force_use_prior ? w.zpost : w.z;
    CKR(launch(h, gop, st));
    h->cur[1] ^= 1;
  }
  // ---- frame predictor (dynamics.py:631-641)
  CKR(launch(h, w.in_conv[2], st));
  const int pf = h->cur[2];
  CKR(launch_lstm(h, 2, st));
  h->cur[2] ^= 1;
  // ---- decoder (vgg_64.py:223-241)
  CKR(launch(h, w.dec[0][pf], st));
  for (int i = 1; i < 9; ++i) CKR(launch(h, w.dec[i][0], st));
  ConvOp fop = w.dec[9][0];
  fop.e.curr_img = a.curr_img; fop.e.next_img = a.next_img; fop.e.mask_next = a.mask_next;
  fop.e.goal_img = a.goal_img; fop.e.goal_mask = a.goal_mask; fop.e.xpred_out = a.xpred_out;
  fop.e.cost_part = a.cost_part; fop.e.zero_robot = a.zero_robot; fop.e.dontcare = a.dontcare;
  CKR(launch(h, fop, st));
  return RAC_OK;
}

int check_ready(rac_handle* h, int n) {
  if (!h) return RAC_ERR_INVALID;
  if (h->ws.B != n || h->ws.arena == nullptr)
    return fail(h, RAC_ERR_STATE, "workspace prepared for batch %d, call needs %d (rac_prepare first)", h->ws.B, n);
  return RAC_OK;
}

__global__ void fill_kernel(float* p, float v, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

}  // namespace

// =================================================================================================== C ABI
extern "C" {

int rac_abi_version(void) { return RAC_ABI_VERSION; }

int rac_create(const rac_config* cfg, rac_handle** out) {
  if (!cfg || !out) return RAC_ERR_INVALID;
  *out = nullptr;
  rac_handle* h = new rac_handle();
  h->cfg = *cfg;
  *out = h;  // returned even on failure so that rac_last_error() is readable; caller must rac_destroy it
  if (cfg->image_height != 48 || cfg->image_width != 64)
    return fail(h, RAC_ERR_INVALID, "only 48x64 images are supported (got %dx%d)", cfg->image_height, cfg->image_width);
  if (cfg->g_dim < 64 || cfg->g_dim % 64 != 0) return fail(h, RAC_ERR_INVALID, "g_dim must be a positive multiple of 64");
  if (cfg->z_dim < 1 || cfg->z_dim > 64) return fail(h, RAC_ERR_INVALID, "z_dim must be in [1, 64]");
  const int aux_c = cfg->action_dim + (cfg->use_robot_state ? cfg->robot_dim : 0) +
                    ((cfg->use_robot_state && cfg->use_future_robot_state) ? cfg->robot_dim : 0);
  if (cfg->action_dim < 1 || aux_c > 64) return fail(h, RAC_ERR_INVALID, "action_dim + robot dims must be in [1, 64]");
  if (cfg->use_future_mask && !cfg->use_mask) return fail(h, RAC_ERR_INVALID, "use_future_mask requires use_mask");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return fail(h, RAC_ERR_CUDA, "no CUDA device: racb200 has no CPU path");
  CK(cudaGetDevice(&h->device));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, h->device));
  if (prop.major != 10) return fail(h, RAC_ERR_CUDA, "racb200 is built for sm_100a only; device is sm_%d%d", prop.major, prop.minor);
  h->num_sms = prop.multiProcessorCount;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
  if (!fn || qres != cudaDriverEntryPointSuccess) return fail(h, RAC_ERR_CUDA, "cuTensorMapEncodeTiled not available");
  h->encode = reinterpret_cast<EncodeTiledFn>(fn);
  if (const char* sz = getenv("RAC_SKIP_ZERO_H")) h->skip_zero_hidden = atoi(sz) != 0;
  if (const char* ct = getenv("RAC_C_TILED")) h->c_tiled = atoi(ct) != 0;
  if (const char* v = getenv("RAC_HALO")) h->use_halo = atoi(v) != 0;
  if (const char* v = getenv("RAC_ACT_BN")) h->act_block_n = atoi(v);
  if (const char* v = getenv("RAC_SPLIT_TAIL")) h->split_tail = atoi(v) != 0;
  if (const char* v = getenv("RAC_YMAJOR")) h->y_major = atoi(v) != 0;
  if (const char* v = getenv("RAC_LSTM_MC")) h->lstm_mc = atoi(v) != 0;
  CK(conv_tc_mc_set_attributes());
  if (const char* v = getenv("RAC_2CTA")) h->two_cta = atoi(v);
  if (const char* v = getenv("RAC_FIRST_TC")) h->first_conv_tc = atoi(v) != 0;
  if (const char* v = getenv("RAC_ENC_DEDUP")) h->enc_dedup = atoi(v) != 0;
  if (const char* v = getenv("RAC_GN_FUSE")) h->gn_fuse_stats = atoi(v) != 0;
  CK(first_conv_tc_set_attributes());
  CK(conv_tc2_set_attributes());
  if (const char* v = getenv("RAC_HALO_BASE_OFFSET")) h->halo_base_offset = atoi(v) != 0;
  if (const char* v = getenv("RAC_HALO_COLUMNS")) h->halo_force_columns = atoi(v) != 0;
  CK(conv_halo_set_attributes());
  if (const char* tm = getenv("RAC_TILE_M")) {
    const int v = atoi(tm);
    if (v != 128 && v != 256) return fail(h, RAC_ERR_INVALID, "RAC_TILE_M must be 128 or 256");
    h->tile_m = v;
  }
  CK(conv_tc_set_attributes());
  CK(cem_set_attributes());
  if (cfg->lstm_group_norm) {
    if (cfg->g_dim > 512) return fail(h, RAC_ERR_UNSUPPORTED, "lstm_group_norm supports g_dim <= 512");
    CK(norm_lstm_set_attributes());
  }
  fill_specs(h);
  return RAC_OK;
}

int rac_train_destroy(rac_handle* h);

int rac_destroy(rac_handle* h) {
  if (!h) return RAC_OK;
  rac_train_destroy(h);
  free_ws(h);
  for (cudaEvent_t ev : h->prof_ev) cudaEventDestroy(ev);
  for (int i = 0; i < RAC_L_COUNT_GN; ++i) {
    if (h->layer[i].w) cudaFree(h->layer[i].w);
    if (h->layer[i].bias) cudaFree(h->layer[i].bias);
  }
  for (int l = 0; l < 3; ++l)
    for (int k = 0; k < 2; ++k)
      if (h->gn_params[l][k]) cudaFree(h->gn_params[l][k]);
  delete h;
  return RAC_OK;
}

const char* rac_last_error(const rac_handle* h) { return h ? h->err : "null handle"; }

int rac_layer_shape(const rac_handle* h, int layer, int64_t* w_elems, int64_t* bias_elems, int* k_packed,
                    int* n_packed) {
  if (!h || layer < 0 || layer >= h->num_layers) return RAC_ERR_INVALID;
  const LayerSpec& s = h->spec[layer];
  if (w_elems) *w_elems = spec_w_elems(s);
  if (bias_elems) *bias_elems = s.n_packed;
  if (k_packed) *k_packed = s.ks * s.ks * s.ctot;
  if (n_packed) *n_packed = s.n_packed;
  return RAC_OK;
}

int rac_load_layer(rac_handle* h, int layer, const void* wsrc, int64_t w_elems, const float* bias, int64_t bias_elems) {
  if (!h || layer < 0 || layer >= h->num_layers || !wsrc || !bias) return fail(h, RAC_ERR_INVALID, "bad layer argument");
  const LayerSpec& s = h->spec[layer];
  if (w_elems != spec_w_elems(s) || bias_elems != s.n_packed)
    return fail(h, RAC_ERR_INVALID, "layer %d: packed size mismatch (w %lld vs %lld, bias %lld vs %d)", layer,
                (long long)w_elems, (long long)spec_w_elems(s), (long long)bias_elems, s.n_packed);
  Layer& L = h->layer[layer];
  const size_t wb = static_cast<size_t>(w_elems) * (s.first ? 4 : 2);
  if (!L.w) CK(cudaMalloc(&L.w, wb));
  if (!L.bias) CK(cudaMalloc(reinterpret_cast<void**>(&L.bias), static_cast<size_t>(bias_elems) * 4));
  CK(cudaMemcpy(L.w, wsrc, wb, cudaMemcpyDefault));
  CK(cudaMemcpy(L.bias, bias, static_cast<size_t>(bias_elems) * 4, cudaMemcpyDefault));
  L.loaded = true;
  return RAC_OK;
}

int rac_load_lstm_norm(rac_handle* h, int layer, const float* packed, int64_t elems) {
  if (!h || !packed) return RAC_ERR_INVALID;
  if (!h->cfg.lstm_group_norm) return fail(h, RAC_ERR_STATE, "rac_load_lstm_norm: the handle was created without lstm_group_norm");
  int l = -1, k = -1;
  switch (layer) {
    case RAC_L_PRIOR_LSTM0: l = 0; k = 0; break;
    case RAC_L_PRIOR_LSTM1: l = 0; k = 1; break;
    case RAC_L_POST_LSTM0: l = 1; k = 0; break;
    case RAC_L_POST_LSTM1: l = 1; k = 1; break;
    case RAC_L_FP_LSTM0: l = 2; k = 0; break;
    case RAC_L_FP_LSTM1: l = 2; k = 1; break;
    default: return fail(h, RAC_ERR_INVALID, "rac_load_lstm_norm: layer %d is not an LSTM cell", layer);
  }
  const int64_t need = static_cast<int64_t>(18) * h->cfg.g_dim;
  if (elems != need) return fail(h, RAC_ERR_INVALID, "rac_load_lstm_norm: %lld floats given, %lld expected", (long long)elems, (long long)need);
  if (!h->gn_params[l][k]) CK(cudaMalloc(reinterpret_cast<void**>(&h->gn_params[l][k]), static_cast<size_t>(need) * 4));
  CK(cudaMemcpy(h->gn_params[l][k], packed, static_cast<size_t>(need) * 4, cudaMemcpyDefault));
  return RAC_OK;
}

int rac_prepare(rac_handle* h, int batch) {
  if (!h || batch < 1) return fail(h, RAC_ERR_INVALID, "batch must be >= 1");
  for (int i = 0; i < h->num_layers; ++i)
    if (!h->layer[i].loaded) return fail(h, RAC_ERR_STATE, "layer %d not loaded (rac_load_layer)", i);
  if (h->cfg.lstm_group_norm)
    for (int l = 0; l < 3; ++l)
      for (int k = 0; k < 2; ++k)
        if (!h->gn_params[l][k]) return fail(h, RAC_ERR_STATE, "GroupNorm parameters of LSTM stack %d cell %d not loaded (rac_load_lstm_norm)", l, k);
  if (h->ws.B == batch && h->ws.arena) return RAC_OK;
  CK(cudaDeviceSynchronize());
  free_ws(h);
  h->ws.B = batch;
  Bump count;
  carve(h, count, batch);
  h->ws.arena_bytes = count.off + 1024;
  CK(cudaMalloc(&h->ws.arena, h->ws.arena_bytes));
  CK(cudaMemset(h->ws.arena, 0, h->ws.arena_bytes));
  Bump bp;
  bp.base = static_cast<char*>(h->ws.arena);
  carve(h, bp, batch);
  CKR(build_ops(h));
  name_buffers(h);
  h->cur[0] = h->cur[1] = h->cur[2] = 0;
  return RAC_OK;
}

int rac_init_hidden(rac_handle* h, int batch, void* stream) {
  CKR(check_ready(h, batch));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  Workspace& w = h->ws;
  const size_t ne = static_cast<size_t>(batch) * 48 * h->cfg.g_dim;
  for (int l = 0; l < 3; ++l)
    for (int k = 0; k < 2; ++k) {
      for (int p = 0; p < 2; ++p) CK(cudaMemsetAsync(w.hs[l][k][p], 0, ne * 2, st));
      CK(cudaMemsetAsync(w.cs[l][k], 0, static_cast<size_t>((batch + 15) / 16 * 16) * 48 * h->cfg.g_dim * 4, st));
    }
  h->cur[0] = h->cur[1] = h->cur[2] = 0;
  h->hidden_zero[0] = h->hidden_zero[1] = h->hidden_zero[2] = true;
  return RAC_OK;
}

int rac_forward(rac_handle* h, const rac_step* s, void* stream) {
  if (!h || !s) return RAC_ERR_INVALID;
  CKR(check_ready(h, s->n));
  const rac_config& c = h->cfg;
  if (!s->image || !s->action || !s->x_pred) return fail(h, RAC_ERR_INVALID, "image, action and x_pred are required");
  if (c.use_mask && !s->mask) return fail(h, RAC_ERR_INVALID, "model_use_mask is set but mask is NULL");
  if (c.use_robot_state && !s->robot) return fail(h, RAC_ERR_INVALID, "model_use_robot_state is set but robot is NULL");
  if (c.use_robot_state && c.use_future_robot_state && !s->robot_next)
    return fail(h, RAC_ERR_INVALID, "model_use_future_robot_state is set but robot_next is NULL");
  if (s->use_posterior && c.use_robot_state && !s->next_robot)
    return fail(h, RAC_ERR_INVALID, "posterior branch needs next_robot");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  CK(launch_img_prep_nchw(s->image, h->ws.img, s->n, 48, 64, st));
  h->launches++;
  StepArgs a{};
  a.n = s->n;
  const int mask_ch = (c.use_mask && c.use_future_mask) ? 2 : 1;  // mask is (n, mask_ch, H, W)
  a.mask_a = s->mask;
  a.mask_b = (s->mask && mask_ch == 2) ? s->mask + 48 * 64 : nullptr;
  a.mask_bstride = static_cast<long long>(mask_ch) * 48 * 64;
  a.robot = s->robot; a.robot_next = s->robot_next;
  a.action = s->action; a.action_stride = c.action_dim;
  a.eps = s->eps; a.seed = s->seed; a.noise_ctr = s->noise_ctr; a.cand_offset = 0; a.sample_mean = s->sample_mean;
  a.use_posterior = s->use_posterior; a.next_robot = s->next_robot; a.eps_post = s->eps_post;
  a.force_use_prior = s->force_use_prior; a.keep_skip = s->keep_skip;
  a.mu_p = s->mu_p; a.logvar_p = s->logvar_p; a.mu = s->mu; a.logvar = s->logvar;
  a.xpred_out = s->x_pred;
  return run_step(h, a, st);
}

int rac_rollout_cost(rac_handle* h, const rac_rollout* r, void* stream) {
  if (!h || !r) return RAC_ERR_INVALID;
  CKR(check_ready(h, r->n));
  const rac_config& c = h->cfg;
  Workspace& w = h->ws;
  if (r->steps < 1 || !r->actions || !r->start_img || !r->goal_imgs || !r->sum_cost || r->num_goals < 1)
    return fail(h, RAC_ERR_INVALID, "rollout: steps, actions, start_img, goal_imgs, sum_cost are required");
  if (r->num_goals > kMaxGoals) return fail(h, RAC_ERR_INVALID, "at most %d goal images", kMaxGoals);
  const bool need_masks = c.use_mask || r->zero_robot || r->dontcare_cost;
  if (need_masks && !r->masks)
    // reference: zero_robot_region(None, img) raises
    return fail(h, RAC_ERR_INVALID, "robot masks are required by this configuration but masks is NULL");
  if (c.use_robot_state && !r->states) return fail(h, RAC_ERR_INVALID, "model_use_robot_state is set but states is NULL");
  if (r->dontcare_cost && !r->goal_masks) return fail(h, RAC_ERR_INVALID, "dontcare cost needs goal masks");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int n = r->n;
  const size_t P0 = 48 * 64;
  CKR(rac_init_hidden(h, n, stream));
  CK(launch_goal_prep(r->goal_imgs, w.goal4, r->num_goals, 48, 64, st));
  CK(launch_img_prep_u8(r->start_img, r->masks, r->zero_robot, w.img, n, 48, 64, st));
  CK(cudaMemsetAsync(r->sum_cost, 0, sizeof(double) * n, st));
  h->launches += 2;
  const int zc = c.z_dim * 48;
  for (int t = 0; t < r->steps; ++t) {
    StepArgs a{};
    a.n = n;
    a.mask_a = r->masks ? r->masks + static_cast<size_t>(t) * r->mask_t_stride : nullptr;
    a.mask_b = r->masks ? r->masks + static_cast<size_t>(t + 1) * r->mask_t_stride : nullptr;
    a.mask_bstride = 48 * 64;
    a.robot = r->states ? r->states + static_cast<size_t>(t) * r->state_t_stride : nullptr;
    a.robot_next = r->states ? r->states + static_cast<size_t>(t + 1) * r->state_t_stride : nullptr;
    a.action = r->actions + static_cast<size_t>(t) * c.action_dim;
    a.action_stride = r->steps * c.action_dim;
    a.eps = r->eps ? r->eps + static_cast<size_t>(t) * n * zc : nullptr;
    a.seed = r->seed; a.noise_ctr = r->noise_ctr_base + t; a.cand_offset = r->cand_offset;
    a.sample_mean = r->sample_mean;
    a.curr_img = w.img;
    a.next_img = w.img;  // in place: every thread reads and writes only its own pixel
    a.mask_next = (r->zero_robot || r->dontcare_cost) ? a.mask_b : nullptr;
    const int gi = t < r->num_goals ? t : r->num_goals - 1;  // trajectory_sampler.py:154
    a.goal_img = w.goal4 + static_cast<size_t>(gi) * P0 * 4;
    a.goal_mask = r->goal_masks ? r->goal_masks + static_cast<size_t>(gi) * P0 : nullptr;
    a.cost_part = w.cost_part;
    a.zero_robot = r->zero_robot; a.dontcare = r->dontcare_cost;
    a.shared_frame = (t == 0 && !r->zero_robot) ? 1 : 0;  // img_prep_u8 gave every candidate the same start frame
    CKR(run_step(h, a, st));
    if (r->obs_out)
      CK(cudaMemcpyAsync(r->obs_out + static_cast<size_t>(t) * n * P0 * 4, w.img, sizeof(float) * n * P0 * 4,
                         cudaMemcpyDeviceToDevice, st));
    const int use = (!r->sparse_cost || t == r->steps - 1) ? 1 : 0;  // trajectory_sampler.py:167
    {
      ProfScope ps(h, "cost_finish", st);
      const bool push = r->peer_world > 0 && r->peer_cost_bufs && t == r->steps - 1;
      CK(launch_cost_finish(w.cost_part, w.dec[9][0].e.cost_nparts, r->dontcare_cost, r->world_cost_weight, use, r->sum_cost,
                            r->step_cost_out ? r->step_cost_out + static_cast<size_t>(t) * n : nullptr, n, st,
                            push ? r->peer_cost_bufs : nullptr, r->peer_world, r->peer_offset));
    }
    h->launches++;
  }
  return RAC_OK;
}

int rac_peer_barrier(uint32_t* const* signal_pads, int slot_base, int rank, int peer_world, uint32_t seq, void* stream) {
  if (!signal_pads) return RAC_ERR_INVALID;
  return launch_peer_barrier(signal_pads, slot_base, rank, peer_world, seq, static_cast<cudaStream_t>(stream)) == cudaSuccess ? RAC_OK
                                                                                                                  : RAC_ERR_INVALID;
}

int rac_cem_sample(const float* mean, const float* stdv, const float* noise, unsigned long long seed, int iter,
                   int n_total, int steps, int action_dim_model, int cand_offset, int n_local, float clamp,
                   float* act2_out, float* act_model_out, void* stream) {
  if (!mean || !stdv || !act2_out || !act_model_out || n_total < 1 || steps < 1 || action_dim_model < 2)
    return RAC_ERR_INVALID;
  cudaError_t e = launch_cem_sample(mean, stdv, noise, seed, iter, n_total, steps, action_dim_model, cand_offset,
                                    n_local, clamp, act2_out, act_model_out, static_cast<cudaStream_t>(stream));
  return e == cudaSuccess ? RAC_OK : RAC_ERR_CUDA;
}

int rac_topk(const double* costs, int n, int k, int64_t* idx_out, double* val_out, void* stream) {
  if (!costs || !idx_out || n < 1 || k < 1 || k > n || k > 4096) return RAC_ERR_INVALID;
  static bool attr = false;
  if (!attr) {
    if (cem_set_attributes() != cudaSuccess) return RAC_ERR_CUDA;
    attr = true;
  }
  cudaError_t e = launch_topk(costs, n, k, idx_out, val_out, static_cast<cudaStream_t>(stream));
  return e == cudaSuccess ? RAC_OK : RAC_ERR_CUDA;
}

int rac_cem_refit(const float* act2, int steps, const int64_t* elite_idx, int k, float std_floor, float* mean_out,
                  float* std_out, void* stream) {
  if (!act2 || !elite_idx || !mean_out || !std_out || steps < 1 || k < 1) return RAC_ERR_INVALID;
  cudaError_t e = launch_refit(act2, steps * 2, elite_idx, k, std_floor, mean_out, std_out,
                               static_cast<cudaStream_t>(stream));
  return e == cudaSuccess ? RAC_OK : RAC_ERR_CUDA;
}

int rac_cem_plan(rac_handle* h, const rac_cem* c, float* mean_out, float* std_out, int64_t* elite_idx_out,
                 double* last_costs_out, void* stream) {
  if (!h || !c || !mean_out) return RAC_ERR_INVALID;
  CKR(check_ready(h, c->n));
  if (c->steps < 1 || c->iters < 1 || c->topk < 1 || c->topk > c->n || c->topk > 4096)
    return fail(h, RAC_ERR_INVALID, "cem: need steps>=1, iters>=1, 1<=topk<=min(n,4096)");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  Workspace& w = h->ws;
  const int n = c->n, L = c->steps, A = h->cfg.action_dim;
  if (w.cem_n != n || w.cem_steps != L) {
    CK(cudaStreamSynchronize(st));
    if (w.act2) { cudaFree(w.act2); cudaFree(w.act5); cudaFree(w.mean); cudaFree(w.sum_cost); cudaFree(w.elite); }
    if (w.rob_states) { cudaFree(w.rob_states); w.rob_states = nullptr; }
    if (w.rob_masks) { cudaFree(w.rob_masks); w.rob_masks = nullptr; }
    CK(cudaMalloc(reinterpret_cast<void**>(&w.act2), sizeof(float) * n * L * 2));
    CK(cudaMalloc(reinterpret_cast<void**>(&w.act5), sizeof(float) * n * L * A));
    CK(cudaMalloc(reinterpret_cast<void**>(&w.mean), sizeof(float) * 64 * 2));
    w.stdv = w.mean + 64;
    CK(cudaMalloc(reinterpret_cast<void**>(&w.sum_cost), sizeof(double) * n));
    CK(cudaMalloc(reinterpret_cast<void**>(&w.elite), sizeof(int64_t) * 4096));
    w.cem_n = n; w.cem_steps = L;
  }
  if (2 * L > 32) return fail(h, RAC_ERR_INVALID, "cem: steps must be <= 16");
  if (c->robot) {
    if (!c->robot_start_state) return fail(h, RAC_ERR_INVALID, "cem: robot model without robot_start_state");
    if (c->robot->kind != RAC_ROBOT_WX250S && c->robot->kind != RAC_ROBOT_FRANKA)
      return fail(h, RAC_ERR_UNSUPPORTED, "cem: robot model kind %d", c->robot->kind);
    if (!w.rob_states) CK(cudaMalloc(reinterpret_cast<void**>(&w.rob_states), sizeof(float) * (L + 1) * n * 5));
    if (c->robot_render_masks && !w.rob_masks)
      CK(cudaMalloc(reinterpret_cast<void**>(&w.rob_masks), sizeof(float) * (L + 1) * n * 48 * 64));
  }
  // mean = 0, std = init_std (cem.py:71-73)
  CK(cudaMemsetAsync(w.mean, 0, sizeof(float) * 2 * L, st));
  fill_kernel<<<1, 64, 0, st>>>(w.stdv, c->init_std, 2 * L);
  CK(cudaGetLastError());
  h->launches++;
  for (int it = 0; it < c->iters; ++it) {
    const float* nz = c->noise ? c->noise + static_cast<size_t>(it) * n * L * 2 : nullptr;
    CK(launch_cem_sample(w.mean, w.stdv, nz, c->rollout.seed, it, n, L, A, 0, n, c->clamp, w.act2, w.act5, st));
    rac_rollout r = c->rollout;
    r.n = n; r.steps = L; r.cand_offset = 0; r.actions = w.act5; r.sum_cost = w.sum_cost;
    r.noise_ctr_base = c->rollout.noise_ctr_base + static_cast<unsigned>(it * L);
    if (c->rollout.eps) r.eps = c->rollout.eps + static_cast<size_t>(it) * L * n * h->cfg.z_dim * 48;
    if (c->robot) {
      // robot_model.predict_batch on the device (trajectory_sampler.py:100-109): no act5.cpu() per iteration
      CK(launch_robot_states(c->robot, c->robot_start_state, w.act5, n, L, A, w.rob_states, static_cast<long long>(n) * 5, st));
      r.states = w.rob_states; r.state_t_stride = static_cast<int64_t>(n) * 5;
      h->launches++;
      if (c->robot_render_masks) {
        CK(launch_robot_masks(c->robot, w.rob_states, static_cast<long long>(n) * 5, n, L + 1, 48, 64, c->robot_extra_radius,
                              w.rob_masks, static_cast<long long>(n) * 48 * 64, st));
        r.masks = w.rob_masks; r.mask_t_stride = static_cast<int64_t>(n) * 48 * 64;
        h->launches++;
      }
      if (!h->cfg.use_robot_state) { r.states = nullptr; }
    }
    CKR(rac_rollout_cost(h, &r, stream));
    CK(launch_topk(w.sum_cost, n, c->topk, w.elite, nullptr, st));
    CK(launch_refit(w.act2, 2 * L, w.elite, c->topk, c->std_floor, w.mean, w.stdv, st));
    h->launches += 3;
  }
  CK(cudaMemcpyAsync(mean_out, w.mean, sizeof(float) * 2 * L, cudaMemcpyDeviceToDevice, st));
  if (std_out) CK(cudaMemcpyAsync(std_out, w.stdv, sizeof(float) * 2 * L, cudaMemcpyDeviceToDevice, st));
  if (elite_idx_out) CK(cudaMemcpyAsync(elite_idx_out, w.elite, sizeof(int64_t) * c->topk, cudaMemcpyDeviceToDevice, st));
  if (last_costs_out) CK(cudaMemcpyAsync(last_costs_out, w.sum_cost, sizeof(double) * n, cudaMemcpyDeviceToDevice, st));
  return RAC_OK;
}

int rac_masked_cost(const float* curr, const float* goal, const float* curr_mask, const float* goal_mask, int dontcare,
                    float* out, int n, int hw, void* stream) {
  if (!curr || !goal || !out || n < 0) return RAC_ERR_INVALID;
  if (dontcare && (!curr_mask || !goal_mask)) return RAC_ERR_INVALID;
  cudaError_t e = launch_masked_cost(curr, goal, curr_mask, goal_mask, dontcare, out, n, hw,
                                     static_cast<cudaStream_t>(stream));
  return e == cudaSuccess ? RAC_OK : RAC_ERR_CUDA;
}

int rac_l1_loss(const float* pred, const float* target, float* out, int64_t numel, void* stream) {
  if (!pred || !target || !out || numel < 1) return RAC_ERR_INVALID;
  return launch_l1_loss(pred, target, out, numel, static_cast<cudaStream_t>(stream)) == cudaSuccess ? RAC_OK : RAC_ERR_CUDA;
}
int rac_dontcare_l1_loss(const float* pred, const float* target, const float* mask, float robot_weight, float* out,
                         int n, int hw, void* stream) {
  if (!pred || !target || !mask || !out || n < 1) return RAC_ERR_INVALID;
  return launch_dontcare_l1_loss(pred, target, mask, robot_weight, out, n, hw, static_cast<cudaStream_t>(stream)) == cudaSuccess ? RAC_OK : RAC_ERR_CUDA;
}
int rac_recon_loss(const float* pred, const float* target, const float* mask, const float* batch_weight, int kind,
                   float robot_weight, float* per_sample, float* out, int n, int hw, void* stream) {
  if (!pred || !target || !per_sample || !out || n < 1 || hw < 1 || kind < 0 || kind > 3) return RAC_ERR_INVALID;
  if ((kind & 1) && !mask) return RAC_ERR_INVALID;
  return launch_recon_loss(pred, target, mask, batch_weight, kind, robot_weight, per_sample, out, n, hw,
                           static_cast<cudaStream_t>(stream)) == cudaSuccess ? RAC_OK : RAC_ERR_CUDA;
}
int rac_predict_states(const rac_robot_model* m, const float* start_state, const float* actions, int n, int steps,
                       int action_dim, float* states_out, int64_t state_t_stride, void* stream) {
  if (!m || !start_state || !actions || !states_out || n < 0 || steps < 1 || action_dim < 2) return RAC_ERR_INVALID;
  if (m->kind != RAC_ROBOT_WX250S && m->kind != RAC_ROBOT_FRANKA) return RAC_ERR_UNSUPPORTED;
  return launch_robot_states(m, start_state, actions, n, steps, action_dim, states_out, state_t_stride,
                             static_cast<cudaStream_t>(stream)) == cudaSuccess ? RAC_OK : RAC_ERR_CUDA;
}
int rac_render_masks(const rac_robot_model* m, const float* states, int64_t state_t_stride, int n, int steps, int H,
                     int W, float extra_radius, float* masks_out, int64_t mask_t_stride, void* stream) {
  if (!m || !states || !masks_out || n < 0 || steps < 0 || H < 1 || W < 1 || extra_radius < 0.f) return RAC_ERR_INVALID;
  return launch_robot_masks(m, states, state_t_stride, n, steps + 1, H, W, extra_radius, masks_out, mask_t_stride,
                            static_cast<cudaStream_t>(stream)) == cudaSuccess ? RAC_OK : RAC_ERR_CUDA;
}
int rac_robot_world_mse(const float* pred, const float* target, const float* mask, float* out2, int n, int hw,
                        void* stream) {
  if (!pred || !target || !mask || !out2 || n < 1) return RAC_ERR_INVALID;
  return launch_robot_world_mse(pred, target, mask, out2, n, hw, static_cast<cudaStream_t>(stream)) == cudaSuccess ? RAC_OK : RAC_ERR_CUDA;
}
int rac_kl_loss(const float* mu1, const float* logvar1, const float* mu2, const float* logvar2, float* out,
                int64_t numel, int batch, void* stream) {
  if (!mu1 || !logvar1 || !mu2 || !logvar2 || !out || numel < 1 || batch < 1) return RAC_ERR_INVALID;
  return launch_kl_loss(mu1, logvar1, mu2, logvar2, out, numel, batch, static_cast<cudaStream_t>(stream)) == cudaSuccess ? RAC_OK : RAC_ERR_CUDA;
}

int rac_psnr(const float* est, const float* target, const float* mask, int clamp01, float* out, int n, int c, int hw,
             void* stream) {
  if (!est || !target || !out || n < 0 || c < 1 || hw < 1) return RAC_ERR_INVALID;
  return launch_psnr(est, target, mask, 0, clamp01, out, n, c, hw, static_cast<cudaStream_t>(stream)) == cudaSuccess ? RAC_OK : RAC_ERR_CUDA;
}
int rac_world_psnr(const float* pred, const float* target, const float* mask, float* out, int n, int hw, void* stream) {
  if (!pred || !target || !mask || !out || n < 0 || hw < 1) return RAC_ERR_INVALID;
  return launch_psnr(pred, target, mask, 1, 0, out, n, 3, hw, static_cast<cudaStream_t>(stream)) == cudaSuccess ? RAC_OK : RAC_ERR_CUDA;
}
int rac_ssim(const float* img1, const float* img2, const float* mask, float* map_out, float* plane_mean_out, int n,
             int c, int h, int w, void* stream) {
  if (!img1 || !img2 || (!map_out && !plane_mean_out) || n < 0 || c < 1 || h < 1 || w < 1) return RAC_ERR_INVALID;
  if (static_cast<size_t>(7) * h * w * 4 > 226 * 1024) return RAC_ERR_UNSUPPORTED;
  return launch_ssim(img1, img2, mask, map_out, plane_mean_out, n, c, h, w, static_cast<cudaStream_t>(stream)) == cudaSuccess ? RAC_OK : RAC_ERR_CUDA;
}

int rac_process_batch(const uint8_t* frames, const void* masks, int mask_is_u8, int B, int T, int H, int W,
                      const rac_augment* aug, float* images_out, float* masks_out, void* stream) {
  if (!frames || !images_out || B < 0 || T < 0 || (masks && !masks_out)) return RAC_ERR_INVALID;
  if (H < 1 || W < 1) return RAC_ERR_INVALID;
  const bool stored48 = H == 48 && W == 64;  // (16-byte staging loads on that path only)
  if ((stored48 && ((reinterpret_cast<uintptr_t>(frames) & 15) || (masks && !mask_is_u8 && (reinterpret_cast<uintptr_t>(masks) & 15)))) ||
      (aug && (reinterpret_cast<uintptr_t>(aug) & 7)))
    return RAC_ERR_INVALID;
  return rac::launch_process_batch(frames, masks, mask_is_u8, B, T, H, W, aug, images_out, masks_out,
                                   static_cast<cudaStream_t>(stream)) == cudaSuccess ? RAC_OK : RAC_ERR_CUDA;
}

int rac_preprocess_states(const float* states, const float* actions, const rac_clip_calib* calib, int B, int T, int R,
                          int A_in, int A_out, float* states_out, float* actions_out, void* stream) {
  if (!states || !calib || !states_out || B < 0 || T < 1 || R < 5 || (actions && (!actions_out || A_in < 1 || A_out < A_in || A_out > A_in + 1)))
    return RAC_ERR_INVALID;
  return rac::launch_preprocess_states(states, actions, calib, B, T, R, A_in, A_out, states_out, actions_out,
                                       static_cast<cudaStream_t>(stream)) == cudaSuccess ? RAC_OK : RAC_ERR_CUDA;
}

int rac_composite(const float* x_pred4, const float* x_j, float* out, int n, int hw, void* stream) {
  if (!x_pred4 || !x_j || !out || n < 0 || hw < 1) return RAC_ERR_INVALID;
  return launch_composite_nchw(x_pred4, x_j, out, n, hw, static_cast<cudaStream_t>(stream)) == cudaSuccess ? RAC_OK : RAC_ERR_CUDA;
}

int rac_debug_buffer(rac_handle* h, const char* name, void** ptr, int64_t* elems, int* elem_bytes) {
  if (!h || !name || !ptr) return RAC_ERR_INVALID;
  auto it = h->ws.named.find(name);
  if (it == h->ws.named.end()) return fail(h, RAC_ERR_INVALID, "no buffer named '%s'", name);
  *ptr = it->second.first;
  if (elems) *elems = it->second.second.first;
  if (elem_bytes) *elem_bytes = it->second.second.second;
  return RAC_OK;
}

int64_t rac_launch_count(const rac_handle* h) { return h ? h->launches : 0; }

int rac_profile_begin(rac_handle* h, const char* name_substr, int max_launches) {
  if (!h || !name_substr || max_launches < 1) return RAC_ERR_INVALID;
  for (cudaEvent_t ev : h->prof_ev) cudaEventDestroy(ev);
  h->prof_ev.assign(static_cast<size_t>(max_launches) * 2, nullptr);
  for (auto& ev : h->prof_ev) CK(cudaEventCreate(&ev));
  h->prof_prefix = name_substr;
  h->prof_used = 0;
  h->prof_dropped = 0;
  return RAC_OK;
}

int rac_profile_end(rac_handle* h, int64_t* launches, double* total_ms) {
  if (!h || !launches || !total_ms) return RAC_ERR_INVALID;
  double acc = 0.0;
  for (size_t i = 0; i + 1 < h->prof_used; i += 2) {
    CK(cudaEventSynchronize(h->prof_ev[i + 1]));
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, h->prof_ev[i], h->prof_ev[i + 1]));
    acc += ms;
  }
  *launches = static_cast<int64_t>(h->prof_used / 2);
  *total_ms = acc;
  for (cudaEvent_t ev : h->prof_ev) cudaEventDestroy(ev);
  h->prof_ev.clear();
  h->prof_prefix.clear();
  h->prof_used = 0;
  return RAC_OK;
}

}  // extern "C"

#include "rac_train.inc.cu"
