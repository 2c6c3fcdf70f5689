// Training step of the SVG model (reference PredictionTrainer._train_step, src/prediction/trainer.py:326-465):
// train-mode forward with posterior, BPTT backward over all steps, Adam. Included at the end of rac_api.cu (one
// translation unit: it uses the handle and the tensor-map helpers defined there).
//
// Every GEMM-shaped piece -- forward convolutions, dgrad (convolution of the output gradient with the transposed,
// flipped weights) and wgrad (dW = dY^T x im2col(X), contraction over the B*H*W rows) -- runs on conv_tc_kernel
// (tcgen05 / TMA); the rest are the small kernels of train_kernels.cu. Parameters, gradients and Adam moments are
// flat fp32 buffers owned by the caller (torch tensors): the gradient buffer can be all-reduced as one message.
//
// Scheduled sampling (trainer.py:132-147,353-356): per step the caller says whether the input frame is the ground
// truth or the model's previous composited prediction; in the latter case the gradient flows back through the
// composite and the encoder input into the previous step, as in the reference (x_pred.clone(), not detached).
//
// Decoder skips: with cfg.last_frame_skip True every step decodes with the skips of its own input frame; with False
// (the config default, src/config/__init__.py:217-222) the reference keeps the skips of the FIRST frame for the whole
// clip (trainer.py:370-371,409-411 with dynamics.py:586-588,644: the model returns the skip it used, so `skip` never
// changes after i == 1, whatever n_past is). `fixed_skip` is that mode: steps t > 0 decode from their own concat
// buffers (dcat*) whose skip halves are copies of step 0's, and the skip-half gradients of all steps are summed in
// G_skip* and enter the encoder backward of step 0 only.
// Step API (rac_train_step_begin / _forward / _backward): the same tape and kernels driven one time step per call, for a
// train-mode SVGConvModel.forward inside the caller's torch autograd graph (the reference's own _train_step body then
// runs unchanged: compositing, losses and the KL term stay in torch, their gradients come back in through _backward).
// Not implemented: heatmaps, multiview.

#include <algorithm>

#include "train_kernels.cuh"
#include "wgrad_tc.cuh"

namespace {

struct TLayer {
  rac_train_layer d{};
  bf16* wp = nullptr;    // [n_packed][taps*ctot]
  bf16* wd = nullptr;    // [ctot][taps*kpad]
  float* dwp = nullptr;  // [kpad][taps*ctot]
  float* bias = nullptr; // [n_packed]
  // output gradients of ALL time steps [steps][M][kpad] (bf16): written by the layer's backward producers, read by the
  // dgrad GEMM of each step and, all steps at once, by the implicit-GEMM weight gradient (the sum over time is part of
  // its contraction)
  bf16* dyT = nullptr;
  int taps = 0, ctot = 0, n_packed = 0, kpad = 0;
  int vec_ok = 0;        // the fused optimizer step may walk the flat side as float4 (adam_pack_vec_ok)
};

struct VggRt {  // one vgg_layer (conv3x3 no bias + BatchNorm + LeakyReLU) at one time step
  float* raw = nullptr;
  float* mean = nullptr;
  float* rstd = nullptr;
};

// Everything the backward pass of one time step needs. The tape is TIME-MAJOR: every field below is a slice of ONE
// allocation [steps][...], so a tensor of `n` consecutive steps is also one contiguous NHWC tensor of n * B images.
// That is what lets (a) every non-recurrent layer run ONCE per training step over all time steps (teacher-forced
// clips) and (b) the weight-gradient kernel walk all steps through one 5-D tensor map.
struct Tape {
  float* img4;
  bf16 *a1, *cat5, *p1, *a2, *cat4, *p2, *a3a, *a3b, *cat3, *p3, *a4a, *a4b, *h4;
  bf16 *aux, *auxp, *pin, *postin, *fin, *z, *zprior;
  bf16* hs[3][2];
  float* cs[3][2];
  float* gates[3][2];
  float *raw_ih[3][2], *raw_hh[3][2], *c_raw[3][2], *gn_stats[3][2];  // lstm_group_norm only
  bf16 *d2a, *d2b, *d3a, *d3b, *d4a, *d5;
  bf16 *dcat5, *dcat4, *dcat3;  // decoder concat buffers: == cat* unless fixed_skip
  VggRt vgg[19];
  float *mu_p, *lv_p, *mu, *lv, *x4;
  float *eps_p, *eps_q;
  float* xp;          // composited prediction (B,3,H,W): the next step's input under scheduled sampling
  const float* xj;    // this step's input frame (ground truth or the previous step's xp)
  int sampled;        // 1: xj is the model's own previous prediction (gradient flows back, trainer.py:354)
};

// the per-step inputs of the model: slices of the time-first batch, or (step API) the tensors of one forward() call
struct StepIO {
  const float *x_j, *m_j, *m_i, *r_j, *r_i, *a_j, *x_i;
  long long mask_bstride;  // floats between the mask planes of consecutive samples
};

// time steps [t0, t0 + n) handled by ONE launch per layer: n = steps for a teacher-forced clip, n = 1 when a step
// consumes the previous step's prediction (scheduled sampling) or with fixed_skip
struct Span {
  int t0, n;
};

// a gradient accumulator (fp32 NHWC) with one slot per time step, time-major like the tape
struct GBuf {
  float* p = nullptr;
  size_t n = 0;
  float* at(int t) const { return p + static_cast<size_t>(t) * n; }
};

struct TrainState {
  rac_train_config cfg{};
  TLayer L[RAC_L_COUNT_GN];
  int nlayers = RAC_L_COUNT;  // RAC_L_COUNT_GN with lstm_group_norm (separate ih / hh gate convolutions)
  bool gn = false;
  float* gn_dy = nullptr;     // [M3, 4g] gate pre-activation gradients of the cell being processed
  float* gn_part = nullptr;   // [B][14 g] per-sample partials of the GroupNorm affine gradients
  float *params = nullptr, *buffers = nullptr, *grads = nullptr, *m = nullptr, *v = nullptr;
  // step API (train-mode SVGConvModel.forward under torch autograd): one forward / backward call per time step, the
  // losses and the compositing live in the caller's autograd graph
  int step_api = 0;
  StepIO io{};
  int active_steps = 0;               // steps on the tape (step API: forward calls so far; else cfg.steps)
  const float *ext_dx4 = nullptr, *ext_dmu = nullptr, *ext_dlv = nullptr, *ext_dmu_p = nullptr, *ext_dlv_p = nullptr;
  float* ext_dimg = nullptr;
  int bwd_next = -1;                  // step API: the time step rac_train_step_backward must be called for next
  const rac_train_batch* cur_bt = nullptr;  // the batch of the running rac_train_forward_backward (grads_ready callback)
  bool unpacked[RAC_L_COUNT_GN] = {};       // layers whose weight gradient is already in the flat buffer
  bool deferred[RAC_L_COUNT_GN] = {};       // layers whose gradient stays packed for the fused optimizer step
  bool packed_valid[RAC_L_COUNT_GN] = {};   // layers whose bf16 operand is current (written by the fused optimizer step)
  float grad_scale = 1.f;
  int per_step = 0;                   // RAC_TRAIN_PER_STEP=1: never batch the time steps (A/B measurements, cross-checks)
  int dgrad_bt = 1;                   // dgrad reads the forward weight packing as an MN-major operand (0: transposed copy Wd)
  float* wg_part = nullptr;           // split-K partials of the layer being processed
  int lstm_fused = 0;                 // RAC_TRAIN_LSTM_FUSED=1: gate convolution with the fused cell epilogue, one launch per cell and step (A/B)
  float* lstm_gx = nullptr;           // [steps][M3][4g] input-half gate pre-activations of the cell being processed
  int w_prefetch = 0;                 // k-blocks of L2 weight prefetch distance in the weight-streaming GEMMs (RAC_TRAIN_W_PREFETCH; off: measured no effect at 4-32)
  int w_tiled = 1;                    // bf16 weights in the k-block-major packing (0: row-major Wp[n][tap][c]; RAC_TRAIN_W_TILED=0, SIMT, Wd)
  int tile_model = 1;                 // RAC_TRAIN_TILE_MODEL=0: round-1 tile rule (256-row tiles unless < 120 of them), no split-K
  int splitk = 1;                     // RAC_TRAIN_SPLITK=0: no split-K in the small-M forward / dgrad GEMMs (A/B measurements)
  float* sk_part = nullptr;           // their slices [ksplit][rows][N]
  size_t sk_part_elems = 0;
  size_t wg_part_elems = 0;
  float* wfirst = nullptr;  // [9*cin][64]
  float* zero64 = nullptr;
  std::vector<Tape> tape;
  void* arena = nullptr;
  GBuf G_d5, G_cat5, G_d4a, G_cat4, G_d3b, G_d3a, G_cat3, G_d2b, G_d2a, G_fin, G_pin, G_postin, G_z, G_h4;
  GBuf G_a4b, G_a4a, G_p3, G_a3b, G_a3a, G_p2, G_a2, G_p1, G_a1;
  float *G_skip5, *G_skip4, *G_skip3;  // fixed_skip: skip-half gradients summed over the steps ([M, C] fp32)
  // dL/dh_t of every ConvLSTM cell, one slot per step, zeroed before the backward pass; every contribution is added:
  // decoder / gaussian heads / the cell above at step t, and the cell's own gate convolution of step t + 1
  GBuf DH[3][2];
  float* G_dc[3][2];
  float* G_img[2];       // gradient w.r.t. a sampled input frame, ping-pong across steps
  float* dbg_draw32 = nullptr;  // RAC_TRAIN_DEBUG_KEEP=1: first-layer raw gradient of the last SAMPLED step (tests)
  int dbg_keep = 0;
  float* bn_scratch;     // [steps][2 * 2048]
  GBuf draw32;           // fp32 copy of the first layer's raw gradient
  float* fw_part;        // per-CTA partial sums of the first layer's weight gradient [kFwBlocks][45 * 64]
  bf16* hzero;
  float* czero;
  float* loss_part;      // [steps * B]
  float* metric_part;    // [max(2 * steps * B, 64)]
  int adam_t = 0;
  int M[4];
};

constexpr int kFwBlocks = 296;  // CTAs of the first layer's weight-gradient kernel (one partial sum each; two per SM)

inline int pick_bn(int n) { return n % 256 == 0 ? 256 : (n % 128 == 0 ? 128 : 64); }

struct GemmGeom {
  int B, H, W, ks;
  bool plain;  // plain GEMM view: rows = B*64, H = 1
};

// Build + launch one conv_tc GEMM with an explicit operand description (tensor maps encoded on the fly).
struct GemmOpt {
  const int* src_dead = nullptr;  // per source: 1 = its k-blocks are skipped (known zero, or handled by another GEMM)
  int partial_only = 0;           // 1: leave the raw split-K slices in T->sk_part (the caller's own kernel reduces them)
  int* ksplit_out = nullptr;      //    ... and how many there are
};

// Tile shape and split-K factor of an fp32-epilogue training GEMM. Two regimes, both settled by measurement
// (profiles/r02_train_tile_sweep.txt, per-CTA phase timelines in profiles/r02_train_timeline_s10.txt):
// * weight-streaming GEMMs -- at most 8 M-tiles of 256 rows, i.e. the per-step ConvLSTM gate convolutions and their
//   dgrads (768 rows, 200-800 k-blocks, 50-100 MB of weights used once): 256-row tiles (a k-block of a 256 x 256 tile
//   runs at the MMA rate, 0.61 us; 128 x 128 tiles took 0.5 us per k-block for a quarter of the work) and as many
//   split-K slices as fill ONE wave of the SMs (12 slices of the 12 tiles of a dgrad: a second wave or fewer, longer
//   slices both measured slower);
// * everything else: 256-row tiles unless there are fewer than 120 of them, then 128 x <= 128 tiles; no split-K.
struct TilePlan { int bm, bn, ksplit; };
TilePlan plan_tiles(int mt256, int mt128, int n_cols, int max_bn, int min_kb, long long rows, bool allow_split,
                    bool partial, int sms, size_t part_capacity) {
  (void)partial;
  const int tiles256 = mt256 * (n_cols / max_bn);
  if (mt256 <= 8 && allow_split) {
    int k = std::max(1, sms / std::max(tiles256, 1));
    k = std::min(k, std::max(1, min_kb / 8));
    while (k > 1 && static_cast<size_t>(k) * rows * n_cols > part_capacity) --k;
    return {256, max_bn, k};
  }
  if (tiles256 < 120) return {128, std::min(max_bn, 128), 1};
  (void)mt128;
  return {256, max_bn, 1};
}

int t_gemm(rac_handle* h, const char* name, const GemmGeom& gg, const std::vector<Src>& srcs, const bf16* w, int ktotal,
           int n_rows_w, int block_n, int epi, const EpiParams& ep, cudaStream_t st, int bt_rows = 0,
           const GemmOpt& opt = GemmOpt()) {
  ConvOp op;
  memset(&op, 0, sizeof(op));
  op.name = name;
  op.epi = epi;
  ConvGeom& g = op.g;
  g.B = gg.B; g.H = gg.H; g.W = gg.W; g.ks = gg.ks; g.pad = gg.ks / 2;
  // batch-16 training GEMMs are small: when 256-row tiles would leave most of the 148 SMs idle, use 128 x <=128 tiles
  auto geom = [&](int bm) -> bool {
    const bool big = bm == 256;
    if (gg.plain) { g.BH = 1; g.NB = big ? 4 : 2; return true; }
    switch (gg.W) {
      case 64: g.BH = big ? 4 : 2; g.NB = 1; return true;
      case 32: g.BH = big ? 8 : 4; g.NB = 1; return true;
      case 16: g.BH = 4; g.NB = big ? 4 : 2; return true;
      case 8: g.BH = 2; g.NB = big ? 16 : 8; return true;
      default: return false;
    }
  };
  if (!geom(256)) return fail(h, RAC_ERR_INVALID, "train gemm %s: bad width %d", name, gg.W);
  TrainState* T = static_cast<TrainState*>(h->train);
  const bool can_split = T && (epi == EPI_F32 || epi == EPI_F32_BT) && h->cfg.conv_impl != 1;
  if (opt.partial_only && !can_split) return fail(h, RAC_ERR_STATE, "train gemm %s: split-K slices need the tcgen05 path", name);
  const long long rows = static_cast<long long>(gg.B) * gg.H * gg.W;
  const int mt256 = ((gg.B + g.NB - 1) / g.NB) * (gg.H / g.BH);
  int ksplit = 1;
  op.block_m = 256;
  op.block_n = block_n;
  if (can_split && T->tile_model) {
    geom(128);
    const int mt128 = ((gg.B + g.NB - 1) / g.NB) * (gg.H / g.BH);
    int live_c = 0;
    for (size_t i = 0; i < srcs.size(); ++i) live_c += (opt.src_dead && opt.src_dead[i]) ? 0 : srcs[i].C / kBlockK;
    const int min_kb = ((gg.ks + 1) / 2) * gg.ks * live_c;  // fewest live k-blocks of any tile (border rows)
    const TilePlan tp = plan_tiles(mt256, mt128, n_rows_w, block_n, min_kb, rows, T->splitk != 0, opt.partial_only != 0,
                                   h->num_sms, T->sk_part_elems);
    TilePlan use = tp;
    // RAC_TRAIN_FORCE_TILE="<name substring>:<block_m>:<ksplit>": overrides the plan of the matching small-M GEMMs (sweeps)
    static const char* force = getenv("RAC_TRAIN_FORCE_TILE");
    for (const char* f = (force && rows <= 1024) ? force : nullptr; f && *f;) {  // comma-separated entries
      char sub[64]; int fbm = 0, fk = 0;
      if (sscanf(f, "%63[^:]:%d:%d", sub, &fbm, &fk) == 3 && strstr(name, sub) && (fbm == 128 || fbm == 256) && fk >= 1 &&
          static_cast<size_t>(fk) * rows * n_rows_w <= T->sk_part_elems && min_kb / fk >= 1)
        use = {fbm, fbm == 256 ? block_n : std::min(block_n, 128), fk};
      f = strchr(f, ',');
      if (f) ++f;
    }
    op.block_m = use.bm; op.block_n = block_n = use.bn; ksplit = use.ksplit;
    geom(use.bm);
  } else {
    // batch-16 training GEMMs are small: when 256-row tiles would leave most of the 148 SMs idle, use 128 x <=128 tiles
    const long long tiles = static_cast<long long>(mt256) * (n_rows_w / block_n);
    if (tiles < 120 || opt.partial_only) {
      op.block_m = 128;
      geom(128);
      if (op.block_n > 128) op.block_n = 128;
      block_n = op.block_n;
    }
  }
  // fp32 epilogues store through the per-warp transpose tile, 128 bits both ways (epilogue.cuh, warp_store_rows_f32_v4):
  // 13.50 -> 13.22 ms per training step. RAC_EPI_STAGED=0: the direct per-row stores, =1: the 32-bit read-back (slower
  // than the direct stores, 14.35 ms: 32 scalar LDS / STG pairs per chunk) -- profiles/r02_train_ab_s24.txt
  static const int epi_staged = [] { const char* v = getenv("RAC_EPI_STAGED"); return v ? atoi(v) : 2; }();
  g.epi_staged = (epi_staged >= 0 && epi_staged <= 2) ? epi_staged : 2;
  g.nsrc = static_cast<int>(srcs.size());
  for (int i = 0; i < g.nsrc; ++i) {
    if (srcs[i].C % kBlockK) return fail(h, RAC_ERR_INVALID, "train gemm %s: source channels %d", name, srcs[i].C);
    g.src_kb[i] = srcs[i].C / kBlockK;
    g.src_dead[i] = (opt.src_dead && opt.src_dead[i]) ? 1 : 0;
    g.ctot += srcs[i].C;
    op.raw.src[i] = srcs[i].p;
    CKR(encode_act_map(h, &op.tm.a[i], srcs[i].p, srcs[i].C, gg.B, gg.H, gg.W, g.BH, g.NB));
  }
  if (gg.ks * gg.ks * g.ctot != ktotal) return fail(h, RAC_ERR_INVALID, "train gemm %s: K mismatch %d vs %d", name, gg.ks * gg.ks * g.ctot, ktotal);
  if (n_rows_w % block_n) return fail(h, RAC_ERR_INVALID, "train gemm %s: N %d not a multiple of %d", name, n_rows_w, block_n);
  g.tiles_per_img = gg.H / g.BH;
  g.num_m_tiles = ((gg.B + g.NB - 1) / g.NB) * g.tiles_per_img;
  g.num_n_tiles = n_rows_w / block_n;
  g.w_shift = ilog2(gg.W);
  g.bhw_shift = ilog2(g.BH * gg.W);
  op.raw.w = w;
  // (EPI_F32_BT: the GEMM's N = n_rows_w is the layer's input-channel count, its K rows are the layer's bt_rows packed
  // output columns; the weight buffer is the forward packing either way)
  const bool tiled = T && T->w_tiled;
  g.w_tiled = tiled ? 1 : 0;
  // few M-tiles = every weight box is fetched from DRAM for 1-8 CTAs: let the spare thread run ahead in L2
  if (T && T->w_prefetch > 0 && g.num_m_tiles <= 8 && h->cfg.conv_impl != 1) g.w_prefetch = T->w_prefetch;
  if (epi == EPI_F32_BT) {
    if (tiled) CKR(encode_w_map_tiled(h, &op.tm.w, w, bt_rows, gg.ks * gg.ks * (n_rows_w / kBlockK), 64));
    else CKR(encode_w_map_bt(h, &op.tm.w, w, n_rows_w, gg.ks * gg.ks, bt_rows));
  } else {
    if (tiled) CKR(encode_w_map_tiled(h, &op.tm.w, w, n_rows_w, ktotal / kBlockK, block_n));
    else CKR(encode_w_map(h, &op.tm.w, w, ktotal, n_rows_w, block_n));
  }
  op.e = ep;
  // RAC_TRAIN_TIMELINE=<substring of the GEMM name>: per-CTA phase stamps of matching launches, printed to stderr
  static const char* tl_name = getenv("RAC_TRAIN_TIMELINE");
  unsigned long long* tl_buf = nullptr;
  if (tl_name && *tl_name && strstr(name, tl_name)) {
    CK(cudaMalloc(&tl_buf, sizeof(unsigned long long) * 8 * 160));
    CK(cudaMemset(tl_buf, 0, sizeof(unsigned long long) * 8 * 160));
    op.g.timeline = tl_buf;
  }
  struct TlPrint {
    unsigned long long* buf; const char* name; const ConvOp* op; cudaStream_t st; int ksplit;
    ~TlPrint() {
      if (!buf) return;
      cudaStreamSynchronize(st);
      std::vector<unsigned long long> hst(8 * 160);
      cudaMemcpy(hst.data(), buf, hst.size() * 8, cudaMemcpyDeviceToHost);
      cudaFree(buf);
      unsigned long long t0 = ~0ull, t5 = 0;
      int n = 0;
      double ph[5] = {0, 0, 0, 0, 0}, mx[5] = {0, 0, 0, 0, 0};
      for (int c = 0; c < 160; ++c) {
        const unsigned long long* t = &hst[c * 8];
        if (!t[0] || !t[5]) continue;
        ++n; t0 = std::min(t0, t[0]); t5 = std::max(t5, t[5]);
        for (int i = 0; i < 5; ++i) { const double d = (double)((long long)t[i + 1] - (long long)t[i]) * 1e-3; ph[i] += d; mx[i] = std::max(mx[i], d); }
      }
      if (n) fprintf(stderr, "[timeline] %s tile %dx%d grid %d ksplit %d span %.1f us | avg(max) us: start->data %.1f(%.1f) mainloop %.1f(%.1f) "
                     "mma->epi %.1f(%.1f) epilogue %.1f(%.1f) tail %.1f(%.1f)\n", name, op->block_m, op->block_n, n, ksplit,
                     (double)(t5 - t0) * 1e-3, ph[0] / n, mx[0], ph[1] / n, mx[1], ph[2] / n, mx[2], ph[3] / n, mx[3], ph[4] / n, mx[4]);
    }
  } tl_print{tl_buf, name, &op, st, ksplit};
  // Split-K for the small-M GEMMs (the per-step ConvLSTM gate convolutions and their dgrads: 768 rows, 200-800 k-blocks):
  // see conv_tc_kernel. Only the fp32 epilogues; the slices are reduced in a fixed order by splitk_reduce_kernel, or by
  // the caller's kernel (partial_only: the ConvLSTM cell).
  if (can_split && (ksplit > 1 || opt.partial_only)) {
    if (static_cast<size_t>(ksplit) * rows * n_rows_w > T->sk_part_elems)
      return fail(h, RAC_ERR_STATE, "train gemm %s: split-K scratch too small", name);
    op.g.ksplit = ksplit;
    op.e.split_part = T->sk_part;
    op.e.split_stride = rows * n_rows_w;
    CKR(launch(h, op, st));
    if (opt.partial_only) {
      if (opt.ksplit_out) *opt.ksplit_out = ksplit;
      return RAC_OK;
    }
    CK(launch_splitk_reduce(op.e, ksplit, rows, n_rows_w, st));
    h->launches++;
    return RAC_OK;
  }
  return launch(h, op, st);
}

int tile_block_n(int n_packed, int epi) {
  if (epi == EPI_FRAME) return 16;
  if (epi == EPI_GAUSS) return 128;
  if (n_packed % 256 == 0) return 256;
  return n_packed % 128 == 0 ? 128 : 64;
}

// ------------------------------------------------------------------ forward pieces
struct VggDef { int layer, H, W, cin, cout; };

// one vgg_layer over the n * B images of a span: convolution -> per-step BatchNorm statistics -> normalise + LeakyReLU
int vgg_forward(rac_handle* h, TrainState* T, VggRt& rt, const VggDef& d, const bf16* in, bf16* out, int cstride,
                int coff, int up, int updates, Span sp, cudaStream_t st) {
  const TLayer& L = T->L[d.layer];
  const int B = T->cfg.batch, nb = sp.n * B;
  EpiParams e{};
  e.cout = L.n_packed;
  e.nseg = 1;
  e.seg[0] = {0, L.n_packed, rt.raw, d.cout, 0, 0};
  if (L.n_packed != d.cout) return fail(h, RAC_ERR_INVALID, "vgg layer %d: cout %d must be a multiple of the pack granularity", d.layer, d.cout);
  CKR(t_gemm(h, "train.vgg.fwd", {nb, d.H, d.W, 3, false}, {{in, d.cin}}, L.wp, 9 * d.cin, L.n_packed,
             pick_bn(L.n_packed), EPI_F32, e, st));
  const int M = B * d.H * d.W;
  CK(launch_bn_stats(rt.raw, M, d.cout, rt.mean, rt.rstd, T->buffers + L.d.rmean_off, T->buffers + L.d.rvar_off, updates, st, sp.n));
  CK(launch_bn_act(rt.raw, rt.mean, rt.rstd, T->params + L.d.gamma_off, T->params + L.d.beta_off, nb, d.H, d.W, d.cout,
                   out, cstride, coff, up, st, sp.n));
  return RAC_OK;
}

// Tiling of one layer's implicit-GEMM weight gradient (everything but the tensor maps / sources)
WgradGeom wg_plan(const TLayer& L, int B, int H, int W, int S) {
  WgradGeom g{};
  g.ks = (L.taps == 25) ? 5 : 3;
  g.pad = g.ks / 2;
  switch (W) {
    case 64: g.BH = 1; break;
    case 32: g.BH = 2; break;
    case 16: g.BH = 4; break;
    default: g.BH = H; break;  // the 6 x 8 latent map: one whole image (48 positions) per k-block
  }
  g.NB = 1;
  g.rows = W * g.BH * g.NB;
  g.hgroups = H / g.BH;
  g.bgroups = B / g.NB;
  g.kb_total = S * g.bgroups * g.hgroups;
  g.kpad = L.kpad;
  g.n_tiles = (L.kpad + 255) / 256;
  g.taps = L.taps;
  g.ctot = L.ctot;
  return g;
}

void wg_split(WgradGeom& g, int num_sms) {
  const int tiles = g.n_tiles * g.num_ctiles * g.taps;
  int splits = std::max(1, (2 * num_sms) / std::max(tiles, 1));
  splits = std::min(splits, std::max(1, g.kb_total / 4));
  g.kb_per_split = (g.kb_total + splits - 1) / splits;
  g.splits = (g.kb_total + g.kb_per_split - 1) / g.kb_per_split;
  g.out_split_stride = static_cast<long long>(g.kpad) * g.taps * g.ctot;
}

int wg_ctiles(WgradGeom& g, const int* src_c, int nsrc) {
  int coff = 0, n = 0;
  for (int s = 0; s < nsrc; ++s) {
    g.src_coff[s] = coff;
    for (int c0 = 0; c0 < src_c[s]; c0 += 256) {
      if (n == kWgMaxCTiles) return -1;
      g.ct_src[n] = s; g.ct_c0[n] = c0; g.ct_w[n] = std::min(256, src_c[s] - c0);
      ++n;
    }
    coff += src_c[s];
  }
  g.num_ctiles = n;
  return coff;
}

// dWp of one layer from the tape of ALL time steps (called once, after the last processed BPTT step)
int wgrad_implicit(rac_handle* h, TrainState* T, TLayer& L, int H, int W, const std::vector<Src>& xs, cudaStream_t st) {
  const int B = T->cfg.batch, S = T->active_steps;
  WgradGeom g = wg_plan(L, B, H, W, S);
  int src_c[kWgMaxSrc];
  if (xs.size() > kWgMaxSrc) return fail(h, RAC_ERR_INVALID, "wgrad: %zu sources", xs.size());
  for (size_t i = 0; i < xs.size(); ++i) { src_c[i] = xs[i].C; g.src_tshift[i] = xs[i].tshift; }
  if (wg_ctiles(g, src_c, static_cast<int>(xs.size())) != L.ctot) return fail(h, RAC_ERR_INVALID, "wgrad: channel mismatch");
  wg_split(g, h->num_sms);
  const size_t out_elems = static_cast<size_t>(g.out_split_stride);
  if (g.splits > 1 && out_elems * g.splits > T->wg_part_elems) return fail(h, RAC_ERR_STATE, "wgrad: split-K scratch too small");
  g.out = g.splits > 1 ? T->wg_part : L.dwp;
  WgradTmaps tm;
  memset(&tm, 0, sizeof(tm));
  const size_t M = static_cast<size_t>(B) * H * W;
  CKR(encode_act_map5(h, &tm.dy, L.dyT, L.kpad, B, H, W, S, M * L.kpad * 2ull, g.BH, g.NB));
  for (size_t i = 0; i < xs.size(); ++i) {
    // time-major tape: step t of a tensor lies B * H * W * C elements after step t - 1
    const bf16* base = xs[i].base0 ? xs[i].base0 : xs[i].p;
    CKR(encode_act_map5(h, &tm.x[i], base, xs[i].C, B, H, W, S, M * xs[i].C * 2ull, g.BH, g.NB));
  }
  CK(launch_wgrad_tc(tm, g, st));
  h->launches++;
  if (g.splits > 1) {
    CK(launch_wgrad_reduce(T->wg_part, g.splits, static_cast<long long>(out_elems), g.out_split_stride, L.dwp, st));
    h->launches++;
  }
  return RAC_OK;
}

// slot of time step t in a layer's all-steps output-gradient buffer [steps][M][kpad] (bf16 GEMM operand)
bf16* dy_slot(TrainState* T, int layer, int H, int W, int t) {
  const TLayer& L = T->L[layer];
  return L.dyT + static_cast<size_t>(t) * T->cfg.batch * H * W * L.kpad;
}

// Backward of one convolution for the steps of `sp`. dY (bf16 [n * M, kpad], packed column order) normally IS the
// layer's slot (its producer wrote it there). xs = the concatenated inputs at step sp.t0 (time-major: the following
// steps are contiguous); for the weight gradient, which runs ONCE per training step after the last processed time step
// (t0 == 0) and contracts over the rows of all steps, xs[i].base0 / tshift say where step 0 of that input lives.
int conv_backward(rac_handle* h, TrainState* T, int layer, int H, int W, const std::vector<Src>& xs, const bf16* dY,
                  const F32Seg* segs, int nseg, Span sp, cudaStream_t st) {
  TLayer& L = T->L[layer];
  const int B = T->cfg.batch, nb = sp.n * B;
  const size_t M = static_cast<size_t>(B) * H * W;
  const int ks = (L.taps == 25) ? 5 : 3;
  const int S = T->active_steps;
  bf16* slot = dy_slot(T, layer, H, W, sp.t0);
  if (dY != slot)
    CK(cudaMemcpyAsync(slot, dY, sizeof(bf16) * M * sp.n * L.kpad, cudaMemcpyDeviceToDevice, st));
  if (sp.t0 == 0) {
    std::vector<Src> x0 = xs;
    for (Src& s : x0)
      if (!s.base0) s.base0 = s.p;  // xs are the inputs of step t0 == 0
    CKR(wgrad_implicit(h, T, L, H, W, x0, st));
    if (L.d.bias_off) CK(launch_bias_grad(L.dyT, static_cast<int>(S * M), L.kpad, L.n_packed, L.d.bias_off, T->grads, st));
    // data-parallel overlap: a large layer's gradient goes to the flat buffer now and is handed to the caller's all-reduce
    const rac_train_batch* bt = T->cur_bt;
    if (bt && bt->defer_unpack && L.d.w_count > 0 && L.d.w_count >= bt->defer_min) {
      // fused optimizer step: the gradient stays packed; (data parallel) all-reduced in place in that form
      T->deferred[layer] = true;
      if (bt->grads_ready)
        bt->grads_ready(bt->grads_ready_user, L.dwp, static_cast<long long>(L.kpad) * L.taps * L.ctot, L.d.w_off, L.d.w_count);
    } else if (bt && bt->grads_ready && L.d.grad_count > 0 && L.d.grad_count >= bt->grads_ready_min) {
      CK(launch_unpack_grads(L.dwp, L.d.row_off, L.d.col_off, L.n_packed, L.taps, L.ctot, L.d.flip, T->grads, st));
      h->launches++;
      T->unpacked[layer] = true;
      bt->grads_ready(bt->grads_ready_user, T->grads + L.d.grad_off, L.d.grad_count, L.d.grad_off, L.d.grad_count);
    }
  }
  // ---- dgrad: dX = conv(dY, Wd)
  if (nseg > 0) {
    EpiParams e{};
    e.cout = L.ctot;
    e.nseg = nseg;
    for (int i = 0; i < nseg; ++i) e.seg[i] = segs[i];
    // B operand = the forward packing Wp read as an MN-major operand (conv_tc.cu, EPI_F32_BT); the transposed + flipped
    // copy Wd only exists for the SIMT cross-check configuration
    if (T->dgrad_bt)
      CKR(t_gemm(h, "train.dgrad", {nb, H, W, ks, false}, {{slot, L.kpad}}, L.wp, L.taps * L.kpad, L.ctot, pick_bn(L.ctot),
                 EPI_F32_BT, e, st, L.n_packed));
    else
      CKR(t_gemm(h, "train.dgrad", {nb, H, W, ks, false}, {{slot, L.kpad}}, L.wd, L.taps * L.kpad, L.ctot, pick_bn(L.ctot),
                 EPI_F32, e, st));
  }
  return RAC_OK;
}

int vgg_backward(rac_handle* h, TrainState* T, VggRt& rt, const VggDef& d, const bf16* in, const float* dy,
                 int dy_cstride, int dy_coff, int up, const F32Seg* segs, int nseg, Span sp, cudaStream_t st) {
  const TLayer& L = T->L[d.layer];
  bf16* slot = dy_slot(T, d.layer, d.H, d.W, sp.t0);
  CK(launch_bn_bwd(dy, dy_cstride, dy_coff, up, rt.raw, rt.mean, rt.rstd, T->params + L.d.gamma_off,
                   T->params + L.d.beta_off, sp.n * T->cfg.batch, d.H, d.W, d.cout, T->bn_scratch, slot, nullptr,
                   T->grads + L.d.gamma_off, T->grads + L.d.beta_off, st, sp.n));
  return conv_backward(h, T, d.layer, d.H, d.W, {{in, d.cin}}, slot, segs, nseg, sp, st);
}

const int kLstm0[3] = {RAC_L_PRIOR_LSTM0, RAC_L_POST_LSTM0, RAC_L_FP_LSTM0};
const int kLstm1[3] = {RAC_L_PRIOR_LSTM1, RAC_L_POST_LSTM1, RAC_L_FP_LSTM1};

const int kLstmHH[3] = {RAC_L_PRIOR_LSTM0_HH, RAC_L_POST_LSTM0_HH, RAC_L_FP_LSTM0_HH};  // + layer

GnCellArgs gn_args(rac_handle* h, TrainState* T, int s, int l, int t) {
  const int B = T->cfg.batch, g = h->cfg.g_dim;
  Tape& tp = T->tape[t];
  const TLayer& Li = T->L[l == 0 ? kLstm0[s] : kLstm1[s]];
  const TLayer& Lh = T->L[kLstmHH[s] + l];
  GnCellArgs a{};
  a.raw_ih = tp.raw_ih[s][l]; a.raw_hh = tp.raw_hh[s][l]; a.params = T->params;
  a.g_ih = Li.d.gamma_off; a.b_ih = Li.d.beta_off; a.g_hh = Lh.d.gamma_off; a.b_hh = Lh.d.beta_off;
  a.g_c = Li.d.cnorm_gamma_off; a.b_c = Li.d.cnorm_beta_off;
  a.c_prev = t > 0 ? T->tape[t - 1].cs[s][l] : T->czero;
  a.gates = tp.gates[s][l]; a.c_raw = tp.c_raw[s][l]; a.c_out = tp.cs[s][l]; a.h_out = tp.hs[s][l];
  a.stats = tp.gn_stats[s][l];
  a.B = B; a.P = 48; a.hid = g;
  return a;
}

// NormConvLSTMCell stack (lstm.py:151-198): two plain convolutions per cell, then the GroupNorm + cell kernel
int lstm_forward_gn(rac_handle* h, TrainState* T, int s, int t, const bf16* xin, cudaStream_t st) {
  const int B = T->cfg.batch, g = h->cfg.g_dim;
  Tape& tp = T->tape[t];
  const bf16* x = xin;
  for (int l = 0; l < 2; ++l) {
    const bf16* hprev = t > 0 ? T->tape[t - 1].hs[s][l] : T->hzero;
    const int ks = l == 0 ? 5 : 3;
    const int ids[2] = {l == 0 ? kLstm0[s] : kLstm1[s], kLstmHH[s] + l};
    const bf16* in[2] = {x, hprev};
    float* out[2] = {tp.raw_ih[s][l], tp.raw_hh[s][l]};
    for (int k = 0; k < 2; ++k) {
      const TLayer& L = T->L[ids[k]];
      EpiParams e{};
      e.bias = L.bias; e.cout = L.n_packed; e.nseg = 1;
      e.seg[0] = {0, L.n_packed, out[k], 4 * g, 0, 0};
      CKR(t_gemm(h, k == 0 ? "train.lstm.ih.fwd" : "train.lstm.hh.fwd", {B, 6, 8, ks, false}, {{in[k], g}}, L.wp,
                 ks * ks * g, L.n_packed, pick_bn(L.n_packed), EPI_F32, e, st));
    }
    CK(launch_gn_cell_fwd(gn_args(h, T, s, l, t), st));
    x = tp.hs[s][l];
  }
  return RAC_OK;
}

// backward of one NormConvLSTMCell stack at step t; dL/dh_t of both cells is complete in DH[s][l][t] by now
int lstm_backward_gn(rac_handle* h, TrainState* T, int s, int t, const bf16* xin, float* g_in, cudaStream_t st) {
  const int g = h->cfg.g_dim;
  Tape& tp = T->tape[t];
  const Span one{t, 1};
  for (int l = 1; l >= 0; --l) {
    const int li = l == 0 ? kLstm0[s] : kLstm1[s], lh = kLstmHH[s] + l;
    GnCellArgs a = gn_args(h, T, s, l, t);
    a.dh = T->DH[s][l].at(t); a.dc = T->G_dc[s][l]; a.dy = T->gn_dy;
    a.d_ih = dy_slot(T, li, 6, 8, t); a.d_hh = dy_slot(T, lh, 6, 8, t);
    a.part = T->gn_part;
    CK(launch_gn_cell_bwd(a, T->grads, st));
    const bf16* x = l == 0 ? xin : tp.hs[s][0];
    const bf16* hprev = t > 0 ? T->tape[t - 1].hs[s][l] : T->hzero;
    // input: layer 1 feeds layer 0's dh (+=), layer 0 feeds the stack input (=)
    F32Seg seg = {0, g, l == 1 ? T->DH[s][0].at(t) : g_in, g, 0, l == 1 ? 1 : 0};
    CKR(conv_backward(h, T, li, 6, 8, {{x, g}}, a.d_ih, &seg, 1, one, st));
    // h_prev: the same cell at step t - 1 (+=); nothing before the first step
    F32Seg segh = {0, g, t > 0 ? T->DH[s][l].at(t - 1) : nullptr, g, 0, 1};
    CKR(conv_backward(h, T, lh, 6, 8, {{hprev, g, T->tape[0].hs[s][l], 1}}, a.d_hh, &segh, t > 0 ? 1 : 0, one, st));
  }
  return RAC_OK;
}

int lstm_forward(rac_handle* h, TrainState* T, int s, int t, const bf16* xin, cudaStream_t st) {
  if (T->gn) return lstm_forward_gn(h, T, s, t, xin, st);
  const int B = T->cfg.batch, g = h->cfg.g_dim;
  Tape& tp = T->tape[t];
  const bf16* x = xin;
  for (int l = 0; l < 2; ++l) {
    const int layer = l == 0 ? kLstm0[s] : kLstm1[s];
    const TLayer& L = T->L[layer];
    const bf16* hprev = t > 0 ? T->tape[t - 1].hs[s][l] : T->hzero;
    EpiParams e{};
    e.bias = L.bias; e.cout = 4 * g; e.hid = g;
    e.c_in = t > 0 ? T->tape[t - 1].cs[s][l] : T->czero;
    e.c_state = tp.cs[s][l]; e.h_out = tp.hs[s][l]; e.gates_out = tp.gates[s][l]; e.exact_math = 1;
    const int ks = l == 0 ? 5 : 3;
    CKR(t_gemm(h, "train.lstm.fwd", {B, 6, 8, ks, false}, {{x, g}, {hprev, g}}, L.wp, ks * ks * 2 * g, L.n_packed,
               tile_block_n(L.n_packed, EPI_LSTM_TRAIN), EPI_LSTM_TRAIN, e, st));
    x = tp.hs[s][l];
  }
  return RAC_OK;
}

// ConvLSTM stack s over the steps of `sp`, layer-major (layer 1 only needs layer 0's outputs, never the other way).
// The gate convolution never runs with its fused cell epilogue here: with 768 rows per step it is 96 tiles for 148 SMs,
// so (a) for a teacher-forced clip the INPUT half of K (x_t: known for all steps up front -- the input convolution's
// output for layer 0, layer 0's h of all steps for layer 1) is one GEMM over all n * B images into lstm_gx, (b) the
// recurrent half (h_{t-1}; nothing at t == 0) runs as split-K work items whose slices (c) the cell kernel sums.
int lstm_forward_span(rac_handle* h, TrainState* T, int s, Span sp, const bf16* xin0, cudaStream_t st) {
  if (T->gn || T->lstm_fused) {
    const size_t xn = static_cast<size_t>(T->cfg.batch) * 48 * h->cfg.g_dim;
    for (int t = sp.t0; t < sp.t0 + sp.n; ++t) CKR(lstm_forward(h, T, s, t, xin0 + static_cast<size_t>(t - sp.t0) * xn, st));
    return RAC_OK;
  }
  const int B = T->cfg.batch, g = h->cfg.g_dim, M = B * 48;
  const size_t xn = static_cast<size_t>(M) * g;
  const bf16* x0 = xin0;
  const bool pre = sp.n > 1;
  for (int l = 0; l < 2; ++l) {
    const int layer = l == 0 ? kLstm0[s] : kLstm1[s];
    const TLayer& L = T->L[layer];
    const int ks = l == 0 ? 5 : 3;
    if (L.n_packed != 4 * g) return fail(h, RAC_ERR_INVALID, "lstm gates: %d packed columns for hid %d", L.n_packed, g);
    if (pre) {
      const int dead[2] = {0, 1};
      GemmOpt o; o.src_dead = dead;
      EpiParams e{};
      e.cout = L.n_packed; e.nseg = 1;
      e.seg[0] = {0, L.n_packed, T->lstm_gx, 4 * g, 0, 0};
      CKR(t_gemm(h, "train.lstm.x.fwd", {sp.n * B, 6, 8, ks, false}, {{x0, g}, {x0, g}}, L.wp, ks * ks * 2 * g, L.n_packed,
                 pick_bn(L.n_packed), EPI_F32, e, st, 0, o));
    }
    for (int t = sp.t0; t < sp.t0 + sp.n; ++t) {
      Tape& tp = T->tape[t];
      const bf16* hprev = t > 0 ? T->tape[t - 1].hs[s][l] : T->hzero;
      const bf16* x = x0 + static_cast<size_t>(t - sp.t0) * xn;
      int nsplit = 0;
      if (!pre || t > 0) {
        const int dead[2] = {pre ? 1 : 0, t > 0 ? 0 : 1};
        GemmOpt o; o.src_dead = dead; o.partial_only = 1; o.ksplit_out = &nsplit;
        EpiParams e{};
        e.cout = L.n_packed;
        CKR(t_gemm(h, "train.lstm.h.fwd", {B, 6, 8, ks, false}, {{x, g}, {hprev, g}}, L.wp, ks * ks * 2 * g, L.n_packed,
                   pick_bn(L.n_packed), EPI_F32, e, st, 0, o));
      }
      CK(launch_lstm_cell_fwd(pre ? T->lstm_gx + static_cast<size_t>(t - sp.t0) * M * 4 * g : nullptr, T->sk_part, nsplit,
                              static_cast<long long>(M) * L.n_packed, L.bias, t > 0 ? T->tape[t - 1].cs[s][l] : nullptr,
                              tp.cs[s][l], tp.hs[s][l], tp.gates[s][l], M, g, st));
      h->launches++;
    }
    x0 = T->tape[sp.t0].hs[s][l];  // time-major tape: the h of the following steps is contiguous
  }
  return RAC_OK;
}

// backward of one ConvLSTM stack at step t; the gradient w.r.t. the stack input lands in `g_in` (=)
int lstm_backward(rac_handle* h, TrainState* T, int s, int t, const bf16* xin, float* g_in, cudaStream_t st) {
  if (T->gn) return lstm_backward_gn(h, T, s, t, xin, g_in, st);
  const int B = T->cfg.batch, g = h->cfg.g_dim, M = B * 48;
  Tape& tp = T->tape[t];
  const Span one{t, 1};
  for (int l = 1; l >= 0; --l) {
    const int layer = l == 0 ? kLstm0[s] : kLstm1[s];
    const float* cprev = t > 0 ? T->tape[t - 1].cs[s][l] : nullptr;
    bf16* slot = dy_slot(T, layer, 6, 8, t);
    CK(launch_lstm_bwd(T->DH[s][l].at(t), T->G_dc[s][l], tp.gates[s][l], cprev, tp.cs[s][l], M, g, slot, st));
    const bf16* x = l == 0 ? xin : tp.hs[s][0];
    const bf16* hprev = t > 0 ? T->tape[t - 1].hs[s][l] : T->hzero;
    F32Seg segs[2];
    // input half: layer 1 feeds layer 0's dh (+=), layer 0 feeds the stack input (=)
    segs[0] = {0, g, l == 1 ? T->DH[s][0].at(t) : g_in, g, 0, l == 1 ? 1 : 0};
    // h_prev half: the same cell at step t - 1 (+=)
    segs[1] = {g, 2 * g, t > 0 ? T->DH[s][l].at(t - 1) : nullptr, g, 0, 1};
    CKR(conv_backward(h, T, layer, 6, 8, {{x, g}, {hprev, g, T->tape[0].hs[s][l], 1}}, slot, segs, 2, one, st));
  }
  return RAC_OK;
}

const VggDef kEnc[10] = {{RAC_L_ENC_C1_0, 48, 64, 0, 64},   {RAC_L_ENC_C1_1, 48, 64, 64, 64},   {RAC_L_ENC_C2_0, 24, 32, 64, 128},
                         {RAC_L_ENC_C2_1, 24, 32, 128, 128}, {RAC_L_ENC_C3_0, 12, 16, 128, 256}, {RAC_L_ENC_C3_1, 12, 16, 256, 256},
                         {RAC_L_ENC_C3_2, 12, 16, 256, 256}, {RAC_L_ENC_C4_0, 6, 8, 256, 512},   {RAC_L_ENC_C4_1, 6, 8, 512, 512},
                         {RAC_L_ENC_C4_2, 6, 8, 512, 0}};
const VggDef kDec[9] = {{RAC_L_DEC_UPC2_0, 6, 8, 0, 512},     {RAC_L_DEC_UPC2_1, 6, 8, 512, 512},   {RAC_L_DEC_UPC2_2, 6, 8, 512, 256},
                        {RAC_L_DEC_UPC3_0, 12, 16, 512, 256}, {RAC_L_DEC_UPC3_1, 12, 16, 256, 256}, {RAC_L_DEC_UPC3_2, 12, 16, 256, 128},
                        {RAC_L_DEC_UPC4_0, 24, 32, 256, 128}, {RAC_L_DEC_UPC4_1, 24, 32, 128, 64},  {RAC_L_DEC_UPC5_0, 48, 64, 128, 64}};

VggDef enc_def(const rac_handle* h, int i) { VggDef d = kEnc[i]; if (i == 9) d.cout = h->cfg.g_dim; return d; }
VggDef dec_def(const rac_handle* h, int i) { VggDef d = kDec[i]; if (i == 0) d.cin = h->cfg.g_dim; return d; }

// ------------------------------------------------------------------ forward phases
// Each phase handles the time steps of `sp` with ONE launch per layer (n * B images). A teacher-forced clip runs every
// non-recurrent phase once for all steps; only the ConvLSTM cells are stepped through time.
StepIO step_io(const rac_handle* h, const TrainState* T, const rac_train_batch* bt, int t) {
  if (T->step_api) return T->io;  // rac_train_step_forward / _backward: the caller's tensors of this one step
  const rac_config& c = h->cfg;
  const size_t B = T->cfg.batch, HW = 48 * 64;
  StepIO io{};
  io.mask_bstride = static_cast<long long>(HW);
  io.x_j = T->tape[t].xj;
  io.x_i = bt->images + static_cast<size_t>(t + 1) * B * 3 * HW;
  io.m_j = bt->masks ? bt->masks + static_cast<size_t>(t) * B * HW : nullptr;
  io.m_i = bt->masks ? bt->masks + static_cast<size_t>(t + 1) * B * HW : nullptr;
  io.r_j = bt->states ? bt->states + static_cast<size_t>(t) * B * c.robot_dim : nullptr;
  io.r_i = bt->states ? bt->states + static_cast<size_t>(t + 1) * B * c.robot_dim : nullptr;
  io.a_j = bt->actions + static_cast<size_t>(t) * B * c.action_dim;
  return io;
}

// encoder (run twice per step in the reference: identical values, BatchNorm running stats updated twice) + the tiled
// action / state channels + the prior / posterior input convolutions
int forward_encoder(rac_handle* h, TrainState* T, const rac_train_batch* bt, Span sp, cudaStream_t st) {
  const rac_config& c = h->cfg;
  const int B = T->cfg.batch, nb = sp.n * B, g = c.g_dim;
  const size_t HW = 48 * 64;
  Tape& tp = T->tape[sp.t0];
  for (int t = sp.t0; t < sp.t0 + sp.n; ++t) {
    Tape& q = T->tape[t];
    if (T->step_api) { q.sampled = 0; q.xj = T->io.x_j; continue; }
    q.sampled = (t > 0 && bt->true_token && !bt->true_token[t]) ? 1 : 0;
    q.xj = q.sampled ? T->tape[t - 1].xp : bt->images + static_cast<size_t>(t) * B * 3 * HW;
  }
  const StepIO io = step_io(h, T, bt, sp.t0);
  // (step API: the caller hands over the frame the model sees, robot pixels already zeroed, trainer.py:365-368)
  CK(launch_img_prep_train(io.x_j, (T->cfg.zero_robot && !T->step_api) ? io.m_j : nullptr, tp.img4, nb, static_cast<int>(HW), st));
  CK(launch_first_conv(tp.img4, c.use_mask ? io.m_j : nullptr, (c.use_mask && c.use_future_mask) ? io.m_i : nullptr,
                       io.mask_bstride, T->wfirst, T->zero64, nullptr, nb, 48, 64, h->enc_cin, st, tp.vgg[0].raw));
  {
    const TLayer& L = T->L[RAC_L_ENC_C1_0];
    CK(launch_bn_stats(tp.vgg[0].raw, B * static_cast<int>(HW), 64, tp.vgg[0].mean, tp.vgg[0].rstd,
                       T->buffers + L.d.rmean_off, T->buffers + L.d.rvar_off, 2, st, sp.n));
    CK(launch_bn_act(tp.vgg[0].raw, tp.vgg[0].mean, tp.vgg[0].rstd, T->params + L.d.gamma_off, T->params + L.d.beta_off,
                     nb, 48, 64, 64, tp.a1, 64, 0, 0, st, sp.n));
  }
  CKR(vgg_forward(h, T, tp.vgg[1], enc_def(h, 1), tp.a1, tp.cat5, 128, 64, 0, 2, sp, st));
  CK(launch_maxpool2(tp.cat5, 128, 64, tp.p1, nb, 48, 64, 64, st));
  CKR(vgg_forward(h, T, tp.vgg[2], enc_def(h, 2), tp.p1, tp.a2, 128, 0, 0, 2, sp, st));
  CKR(vgg_forward(h, T, tp.vgg[3], enc_def(h, 3), tp.a2, tp.cat4, 256, 128, 0, 2, sp, st));
  CK(launch_maxpool2(tp.cat4, 256, 128, tp.p2, nb, 24, 32, 128, st));
  CKR(vgg_forward(h, T, tp.vgg[4], enc_def(h, 4), tp.p2, tp.a3a, 256, 0, 0, 2, sp, st));
  CKR(vgg_forward(h, T, tp.vgg[5], enc_def(h, 5), tp.a3a, tp.a3b, 256, 0, 0, 2, sp, st));
  CKR(vgg_forward(h, T, tp.vgg[6], enc_def(h, 6), tp.a3b, tp.cat3, 512, 256, 0, 2, sp, st));
  CK(launch_maxpool2(tp.cat3, 512, 256, tp.p3, nb, 12, 16, 256, st));
  CKR(vgg_forward(h, T, tp.vgg[7], enc_def(h, 7), tp.p3, tp.a4a, 512, 0, 0, 2, sp, st));
  CKR(vgg_forward(h, T, tp.vgg[8], enc_def(h, 8), tp.a4a, tp.a4b, 512, 0, 0, 2, sp, st));
  CKR(vgg_forward(h, T, tp.vgg[9], enc_def(h, 9), tp.a4b, tp.h4, g, 0, 0, 2, sp, st));
  CK(launch_aux_tile(io.a_j, c.action_dim, c.action_dim, c.use_robot_state ? io.r_j : nullptr,
                     (c.use_robot_state && c.use_future_robot_state) ? io.r_i : nullptr, c.robot_dim, tp.aux, nb, 48, st));
  return RAC_OK;
}

int input_conv_fwd(rac_handle* h, TrainState* T, int layer, std::vector<Src> srcs, bf16* out, int nb, cudaStream_t st) {
  const TLayer& L = T->L[layer];
  const int g = h->cfg.g_dim;
  EpiParams e{};
  e.bias = L.bias; e.cout = g; e.out = out; e.out_cstride = g; e.out_coff = 0; e.upsample = 0; e.lrelu = 0;
  return t_gemm(h, "train.input_conv.fwd", {nb, 6, 8, 3, false}, srcs, L.wp, 9 * L.ctot, L.n_packed,
                tile_block_n(L.n_packed, EPI_ACT), EPI_ACT, e, st);
}

// prior_input_conv / posterior_input_conv (dynamics.py:600-602,621-623; h_target == h4, :619)
int forward_input_convs(rac_handle* h, TrainState* T, const rac_train_batch* bt, Span sp, cudaStream_t st) {
  const rac_config& c = h->cfg;
  const int nb = sp.n * T->cfg.batch, g = c.g_dim;
  Tape& tp = T->tape[sp.t0];
  const StepIO io = step_io(h, T, bt, sp.t0);
  CKR(input_conv_fwd(h, T, RAC_L_PRIOR_IN, {{tp.aux, 64}, {tp.h4, g}}, tp.pin, nb, st));
  if (c.use_robot_state) {
    CK(launch_aux_tile(nullptr, 0, 0, io.r_i, nullptr, c.robot_dim, tp.auxp, nb, 48, st));
    CKR(input_conv_fwd(h, T, RAC_L_POST_IN, {{tp.auxp, 64}, {tp.h4, g}}, tp.postin, nb, st));
  } else {
    CKR(input_conv_fwd(h, T, RAC_L_POST_IN, {{tp.h4, g}}, tp.postin, nb, st));
  }
  return RAC_OK;
}

// mu_net / logvar_net + reparameterisation (lstm.py:276-286) of stack s (0 prior, 1 posterior)
int forward_gauss(rac_handle* h, TrainState* T, int s, Span sp, cudaStream_t st) {
  const int nb = sp.n * T->cfg.batch, g = h->cfg.g_dim;
  Tape& tp = T->tape[sp.t0];
  const TLayer& L = T->L[s == 0 ? RAC_L_PRIOR_GAUSS : RAC_L_POST_GAUSS];
  EpiParams e{};
  e.bias = L.bias; e.cout = 128; e.z_dim = h->cfg.z_dim;
  e.eps = s == 0 ? tp.eps_p : tp.eps_q;
  e.mu_out = s == 0 ? tp.mu_p : tp.mu;
  e.logvar_out = s == 0 ? tp.lv_p : tp.lv;
  e.z_out = s == 0 ? tp.zprior : tp.z;
  return t_gemm(h, "train.gauss.fwd", {nb, 6, 8, 3, false}, {{tp.hs[s][1], g}}, L.wp, 9 * g, 128, 128, EPI_GAUSS, e, st);
}

int forward_fp_in(rac_handle* h, TrainState* T, Span sp, cudaStream_t st) {
  Tape& tp = T->tape[sp.t0];
  return input_conv_fwd(h, T, RAC_L_FP_IN, {{tp.aux, 64}, {tp.h4, h->cfg.g_dim}, {tp.z, 64}}, tp.fin, sp.n * T->cfg.batch, st);
}

// decoder + frame head + compositing + the step's logged metrics / KL value
int forward_decoder(rac_handle* h, TrainState* T, const rac_train_batch* bt, Span sp, cudaStream_t st) {
  const rac_config& c = h->cfg;
  const int B = T->cfg.batch, nb = sp.n * B, z = c.z_dim;
  const size_t HW = 48 * 64;
  Tape& tp = T->tape[sp.t0];
  const StepIO io = step_io(h, T, bt, sp.t0);
  CKR(vgg_forward(h, T, tp.vgg[10], dec_def(h, 0), tp.hs[2][1], tp.d2a, 512, 0, 0, 1, sp, st));
  CKR(vgg_forward(h, T, tp.vgg[11], dec_def(h, 1), tp.d2a, tp.d2b, 512, 0, 0, 1, sp, st));
  if (tp.dcat5 != tp.cat5) {
    // fixed_skip (one step at a time): the skip halves are those of the first frame
    const Tape& t0 = T->tape[0];
    CK(cudaMemcpy2DAsync(tp.dcat3 + 256, 512 * sizeof(bf16), t0.cat3 + 256, 512 * sizeof(bf16), 256 * sizeof(bf16),
                         static_cast<size_t>(B) * 192, cudaMemcpyDeviceToDevice, st));
    CK(cudaMemcpy2DAsync(tp.dcat4 + 128, 256 * sizeof(bf16), t0.cat4 + 128, 256 * sizeof(bf16), 128 * sizeof(bf16),
                         static_cast<size_t>(B) * 768, cudaMemcpyDeviceToDevice, st));
    CK(cudaMemcpy2DAsync(tp.dcat5 + 64, 128 * sizeof(bf16), t0.cat5 + 64, 128 * sizeof(bf16), 64 * sizeof(bf16),
                         static_cast<size_t>(B) * 3072, cudaMemcpyDeviceToDevice, st));
  }
  CKR(vgg_forward(h, T, tp.vgg[12], dec_def(h, 2), tp.d2b, tp.dcat3, 512, 0, 1, 1, sp, st));
  CKR(vgg_forward(h, T, tp.vgg[13], dec_def(h, 3), tp.dcat3, tp.d3a, 256, 0, 0, 1, sp, st));
  CKR(vgg_forward(h, T, tp.vgg[14], dec_def(h, 4), tp.d3a, tp.d3b, 256, 0, 0, 1, sp, st));
  CKR(vgg_forward(h, T, tp.vgg[15], dec_def(h, 5), tp.d3b, tp.dcat4, 256, 0, 1, 1, sp, st));
  CKR(vgg_forward(h, T, tp.vgg[16], dec_def(h, 6), tp.dcat4, tp.d4a, 128, 0, 0, 1, sp, st));
  CKR(vgg_forward(h, T, tp.vgg[17], dec_def(h, 7), tp.d4a, tp.dcat5, 128, 0, 1, 1, sp, st));
  CKR(vgg_forward(h, T, tp.vgg[18], dec_def(h, 8), tp.dcat5, tp.d5, 64, 0, 0, 1, sp, st));
  {
    const TLayer& L = T->L[RAC_L_DEC_UPC5_1];
    EpiParams e{};
    e.bias = L.bias; e.cout = 4; e.xpred_out = tp.x4;
    CKR(t_gemm(h, "train.frame.fwd", {nb, 48, 64, 3, false}, {{tp.d5, 64}}, L.wp, 9 * 64, 16, 16, EPI_FRAME, e, st));
  }
  if (T->step_api) return RAC_OK;  // compositing, losses and metrics belong to the caller's autograd graph
  CK(launch_composite(tp.x4, io.x_j, tp.xp, nb, static_cast<int>(HW), st));
  if (io.m_i)  // logging metrics of the reference step (trainer.py:436-439)
    CK(launch_robot_world_mse_batched(tp.xp, io.x_i, io.m_i, T->metric_part, bt->losses + 2, nb, B, static_cast<int>(HW), st));
  // KL(posterior || prior) value (trainer.py:454-458)
  CK(launch_kl_loss_batched(tp.mu, tp.lv, tp.mu_p, tp.lv_p, T->metric_part, bt->losses + 1,
                            static_cast<long long>(nb) * z * 48, B, st));
  return RAC_OK;
}

// ------------------------------------------------------------------ backward phases
// reconstruction loss + decoder; the gradient w.r.t. the frame predictor's output h lands in DH[2][1][t] (+=)
int backward_decoder(rac_handle* h, TrainState* T, const rac_train_batch* bt, Span sp, cudaStream_t st) {
  const rac_config& c = h->cfg;
  const int B = T->cfg.batch, nb = sp.n * B, g = c.g_dim, t0 = sp.t0;
  const size_t HW = 48 * 64;
  const int S = T->cfg.steps;
  Tape& tp = T->tape[t0];
  const StepIO io = step_io(h, T, bt, t0);
  // scheduled sampling (one step at a time): gradient handed over by / to the neighbouring steps through the frame
  const int cur = (S - 1 - t0) & 1;
  const float* gp_in = (!T->step_api && sp.n == 1 && t0 + 1 < S && T->tape[t0 + 1].sampled) ? T->G_img[cur] : nullptr;
  float* gxj_out = (sp.n == 1 && tp.sampled) ? T->G_img[cur ^ 1] : nullptr;
  // ---- reconstruction loss and its gradient w.r.t. the decoder logits (trainer.py:406-433)
  bf16* dlogit = dy_slot(T, RAC_L_DEC_UPC5_1, 48, 64, t0);
  if (T->step_api) {
    // dL/d(x_pred) comes from the caller's autograd graph: through the sigmoid to the logits
    CK(launch_sigmoid_bwd(tp.x4, T->ext_dx4, dlogit, nb, static_cast<int>(HW), st));
  } else {
    CK(launch_frame_loss(tp.x4, io.x_j, io.x_i, io.m_i, T->cfg.recon_kind, T->cfg.robot_pixel_weight, nb, static_cast<int>(HW),
                         T->loss_part, dlogit, gp_in, gxj_out, st, bt->batch_weight, B));
    CK(launch_sum_f32(T->loss_part, nb, bt->losses + 0, st));
  }
  F32Seg seg[3];
  seg[0] = {0, 64, T->G_d5.at(t0), 64, 0, 0};
  CKR(conv_backward(h, T, RAC_L_DEC_UPC5_1, 48, 64, {{tp.d5, 64}}, dlogit, seg, 1, sp, st));
  // gradient of a concat buffer: [decoder half | skip half]. fixed_skip: the skip half is summed over the steps
  // (first writer = the first processed step, t == S - 1) instead of going to this step's encoder
  const bool fixed = T->cfg.fixed_skip != 0;
  const int skip_acc = (t0 == T->active_steps - 1) ? 0 : 1;
  auto cat_segs = [&](float* G_cat, float* G_skip, int half) -> int {
    if (!fixed) { seg[0] = {0, 2 * half, G_cat, 2 * half, 0, 0}; return 1; }
    seg[0] = {0, half, G_cat, 2 * half, 0, 0};
    seg[1] = {half, 2 * half, G_skip, half, 0, skip_acc};
    return 2;
  };
  int ncat = cat_segs(T->G_cat5.at(t0), T->G_skip5, 64);
  CKR(vgg_backward(h, T, tp.vgg[18], dec_def(h, 8), tp.dcat5, T->G_d5.at(t0), 64, 0, 0, seg, ncat, sp, st));
  seg[0] = {0, 128, T->G_d4a.at(t0), 128, 0, 0};
  CKR(vgg_backward(h, T, tp.vgg[17], dec_def(h, 7), tp.d4a, T->G_cat5.at(t0), 128, 0, 1, seg, 1, sp, st));
  ncat = cat_segs(T->G_cat4.at(t0), T->G_skip4, 128);
  CKR(vgg_backward(h, T, tp.vgg[16], dec_def(h, 6), tp.dcat4, T->G_d4a.at(t0), 128, 0, 0, seg, ncat, sp, st));
  seg[0] = {0, 256, T->G_d3b.at(t0), 256, 0, 0};
  CKR(vgg_backward(h, T, tp.vgg[15], dec_def(h, 5), tp.d3b, T->G_cat4.at(t0), 256, 0, 1, seg, 1, sp, st));
  seg[0] = {0, 256, T->G_d3a.at(t0), 256, 0, 0};
  CKR(vgg_backward(h, T, tp.vgg[14], dec_def(h, 4), tp.d3a, T->G_d3b.at(t0), 256, 0, 0, seg, 1, sp, st));
  ncat = cat_segs(T->G_cat3.at(t0), T->G_skip3, 256);
  CKR(vgg_backward(h, T, tp.vgg[13], dec_def(h, 3), tp.dcat3, T->G_d3a.at(t0), 256, 0, 0, seg, ncat, sp, st));
  seg[0] = {0, 512, T->G_d2b.at(t0), 512, 0, 0};
  CKR(vgg_backward(h, T, tp.vgg[12], dec_def(h, 2), tp.d2b, T->G_cat3.at(t0), 512, 0, 1, seg, 1, sp, st));
  seg[0] = {0, 512, T->G_d2a.at(t0), 512, 0, 0};
  CKR(vgg_backward(h, T, tp.vgg[11], dec_def(h, 1), tp.d2a, T->G_d2b.at(t0), 512, 0, 0, seg, 1, sp, st));
  seg[0] = {0, g, T->DH[2][1].at(t0), g, 0, 1};
  CKR(vgg_backward(h, T, tp.vgg[10], dec_def(h, 0), tp.hs[2][1], T->G_d2a.at(t0), 512, 0, 0, seg, 1, sp, st));
  return RAC_OK;
}

// frame_pred_input_conv: G_fin (from the frame predictor stack) -> G_h4 (=, first writer), G_z (=)
int backward_fp_in(rac_handle* h, TrainState* T, Span sp, cudaStream_t st) {
  const int nb = sp.n * T->cfg.batch, g = h->cfg.g_dim, t0 = sp.t0;
  Tape& tp = T->tape[t0];
  bf16* slot = dy_slot(T, RAC_L_FP_IN, 6, 8, t0);
  CK(launch_cast_bf16(T->G_fin.at(t0), static_cast<long long>(nb) * 48 * g, slot, st));
  F32Seg seg[3];
  seg[0] = {0, 64, nullptr, 0, 0, 0};
  seg[1] = {64, 64 + g, T->G_h4.at(t0), g, 0, 0};
  seg[2] = {64 + g, 128 + g, T->G_z.at(t0), 64, 0, 0};
  return conv_backward(h, T, RAC_L_FP_IN, 6, 8, {{tp.aux, 64}, {tp.h4, g}, {tp.z, 64}}, slot, seg, 3, sp, st);
}

// z sample + KL -> posterior and prior heads; their input gradients land in DH[1][1] / DH[0][1] (+=)
int backward_gauss(rac_handle* h, TrainState* T, Span sp, cudaStream_t st) {
  const int B = T->cfg.batch, nb = sp.n * B, g = h->cfg.g_dim, z = h->cfg.z_dim, t0 = sp.t0;
  Tape& tp = T->tape[t0];
  bf16* dpost = dy_slot(T, RAC_L_POST_GAUSS, 6, 8, t0);
  bf16* dprior = dy_slot(T, RAC_L_PRIOR_GAUSS, 6, 8, t0);
  if (T->step_api)  // the KL term lives in the caller's graph: its gradients w.r.t. (mu, logvar, mu_p, logvar_p) are inputs
    CK(launch_gauss_bwd_ext(T->G_z.at(t0), tp.lv, tp.eps_q, T->ext_dmu, T->ext_dlv, T->ext_dmu_p, T->ext_dlv_p, nb, z, 48,
                            dpost, dprior, st));
  else
    CK(launch_gauss_bwd(T->G_z.at(t0), tp.mu, tp.lv, tp.eps_q, tp.mu_p, tp.lv_p, nb, z, 48, T->cfg.kl_beta, B, dpost, dprior, st));
  F32Seg seg = {0, g, T->DH[1][1].at(t0), g, 0, 1};
  CKR(conv_backward(h, T, RAC_L_POST_GAUSS, 6, 8, {{tp.hs[1][1], g}}, dpost, &seg, 1, sp, st));
  seg = {0, g, T->DH[0][1].at(t0), g, 0, 1};
  CKR(conv_backward(h, T, RAC_L_PRIOR_GAUSS, 6, 8, {{tp.hs[0][1], g}}, dprior, &seg, 1, sp, st));
  return RAC_OK;
}

// posterior_input_conv (s == 1) / prior_input_conv (s == 0): G_postin / G_pin -> G_h4 (+=)
int backward_input_conv(rac_handle* h, TrainState* T, int s, Span sp, cudaStream_t st) {
  const rac_config& c = h->cfg;
  const int nb = sp.n * T->cfg.batch, g = c.g_dim, t0 = sp.t0;
  Tape& tp = T->tape[t0];
  const int layer = s == 1 ? RAC_L_POST_IN : RAC_L_PRIOR_IN;
  bf16* slot = dy_slot(T, layer, 6, 8, t0);
  CK(launch_cast_bf16((s == 1 ? T->G_postin : T->G_pin).at(t0), static_cast<long long>(nb) * 48 * g, slot, st));
  F32Seg seg[2];
  if (s == 1 && !c.use_robot_state) {
    seg[0] = {0, g, T->G_h4.at(t0), g, 0, 1};
    return conv_backward(h, T, layer, 6, 8, {{tp.h4, g}}, slot, seg, 1, sp, st);
  }
  seg[0] = {0, 64, nullptr, 0, 0, 0};
  seg[1] = {64, 64 + g, T->G_h4.at(t0), g, 0, 1};
  return conv_backward(h, T, layer, 6, 8, {{s == 1 ? tp.auxp : tp.aux, 64}, {tp.h4, g}}, slot, seg, 2, sp, st);
}

// encoder (the gradients of the two reference passes are summed in G_h4 / the skip halves)
int backward_encoder(rac_handle* h, TrainState* T, const rac_train_batch* bt, Span sp, cudaStream_t st) {
  const rac_config& c = h->cfg;
  const int B = T->cfg.batch, nb = sp.n * B, g = c.g_dim, t0 = sp.t0;
  const size_t HW = 48 * 64;
  Tape& tp = T->tape[t0];
  const StepIO io = step_io(h, T, bt, t0);
  const bool fixed = T->cfg.fixed_skip != 0;
  F32Seg seg[3];
  seg[0] = {0, 512, T->G_a4b.at(t0), 512, 0, 0};
  CKR(vgg_backward(h, T, tp.vgg[9], enc_def(h, 9), tp.a4b, T->G_h4.at(t0), g, 0, 0, seg, 1, sp, st));
  seg[0] = {0, 512, T->G_a4a.at(t0), 512, 0, 0};
  CKR(vgg_backward(h, T, tp.vgg[8], enc_def(h, 8), tp.a4a, T->G_a4b.at(t0), 512, 0, 0, seg, 1, sp, st));
  seg[0] = {0, 256, T->G_p3.at(t0), 256, 0, 0};
  CKR(vgg_backward(h, T, tp.vgg[7], enc_def(h, 7), tp.p3, T->G_a4a.at(t0), 512, 0, 0, seg, 1, sp, st));
  // gradient of the encoder outputs that double as skips: pooling path + decoder skip path. fixed_skip: steps t > 0
  // get the pooling path only; step 0 adds it to the all-steps skip sum
  struct SkipGrad { float* p; int cstride, coff, acc; };
  auto skip_grad = [&](float* G_cat, float* G_skip, int half) -> SkipGrad {
    if (!fixed) return {G_cat, 2 * half, half, 1};
    if (t0 > 0) return {G_cat, 2 * half, half, 0};
    return {G_skip, half, 0, 1};
  };
  SkipGrad sk = skip_grad(T->G_cat3.at(t0), T->G_skip3, 256);
  CK(launch_pool_bwd(tp.cat3, 512, 256, T->G_p3.at(t0), nb, 12, 16, 256, sk.p, sk.cstride, sk.coff, sk.acc, st));
  seg[0] = {0, 256, T->G_a3b.at(t0), 256, 0, 0};
  CKR(vgg_backward(h, T, tp.vgg[6], enc_def(h, 6), tp.a3b, sk.p, sk.cstride, sk.coff, 0, seg, 1, sp, st));
  seg[0] = {0, 256, T->G_a3a.at(t0), 256, 0, 0};
  CKR(vgg_backward(h, T, tp.vgg[5], enc_def(h, 5), tp.a3a, T->G_a3b.at(t0), 256, 0, 0, seg, 1, sp, st));
  seg[0] = {0, 128, T->G_p2.at(t0), 128, 0, 0};
  CKR(vgg_backward(h, T, tp.vgg[4], enc_def(h, 4), tp.p2, T->G_a3a.at(t0), 256, 0, 0, seg, 1, sp, st));
  sk = skip_grad(T->G_cat4.at(t0), T->G_skip4, 128);
  CK(launch_pool_bwd(tp.cat4, 256, 128, T->G_p2.at(t0), nb, 24, 32, 128, sk.p, sk.cstride, sk.coff, sk.acc, st));
  seg[0] = {0, 128, T->G_a2.at(t0), 128, 0, 0};
  CKR(vgg_backward(h, T, tp.vgg[3], enc_def(h, 3), tp.a2, sk.p, sk.cstride, sk.coff, 0, seg, 1, sp, st));
  seg[0] = {0, 64, T->G_p1.at(t0), 64, 0, 0};
  CKR(vgg_backward(h, T, tp.vgg[2], enc_def(h, 2), tp.p1, T->G_a2.at(t0), 128, 0, 0, seg, 1, sp, st));
  sk = skip_grad(T->G_cat5.at(t0), T->G_skip5, 64);
  CK(launch_pool_bwd(tp.cat5, 128, 64, T->G_p1.at(t0), nb, 48, 64, 64, sk.p, sk.cstride, sk.coff, sk.acc, st));
  seg[0] = {0, 64, T->G_a1.at(t0), 64, 0, 0};
  CKR(vgg_backward(h, T, tp.vgg[1], enc_def(h, 1), tp.a1, sk.p, sk.cstride, sk.coff, 0, seg, 1, sp, st));
  {
    // encoder.c1.0: BatchNorm backward, then the small-K weight gradient on CUDA cores (no input gradient needed)
    const TLayer& L = T->L[RAC_L_ENC_C1_0];
    float* draw32 = T->draw32.at(t0);
    CK(launch_bn_bwd(T->G_a1.at(t0), 64, 0, 0, tp.vgg[0].raw, tp.vgg[0].mean, tp.vgg[0].rstd, T->params + L.d.gamma_off,
                     T->params + L.d.beta_off, nb, 48, 64, 64, T->bn_scratch, nullptr, draw32,
                     T->grads + L.d.gamma_off, T->grads + L.d.beta_off, st, sp.n));
    CK(launch_first_wgrad(tp.img4, c.use_mask ? io.m_j : nullptr, (c.use_mask && c.use_future_mask) ? io.m_i : nullptr,
                          io.mask_bstride, draw32, T->grads + L.d.w_off, nb, 48, 64, h->enc_cin, T->fw_part,
                          kFwBlocks, st));
    if (T->step_api && T->ext_dimg) {  // gradient w.r.t. the input frame (a model-sampled frame feeds the next step)
      CK(cudaMemsetAsync(T->ext_dimg, 0, sizeof(float) * static_cast<size_t>(B) * 3 * HW, st));
      CK(launch_first_dgrad(draw32, T->wfirst, h->enc_cin, nullptr, T->ext_dimg, B, 48, 64, st));
    }
    // a model-sampled input frame also receives gradient through the encoder (zeroed robot pixels get none)
    if (sp.n == 1 && tp.sampled) {
      float* gxj_out = T->G_img[((T->cfg.steps - 1 - t0) & 1) ^ 1];
      if (T->dbg_keep)
        CK(cudaMemcpyAsync(T->dbg_draw32, draw32, sizeof(float) * static_cast<size_t>(B) * HW * 64, cudaMemcpyDeviceToDevice, st));
      CK(launch_first_dgrad(draw32, T->wfirst, h->enc_cin, T->cfg.zero_robot ? io.m_j : nullptr, gxj_out, B, 48, 64, st));
    }
  }
  return RAC_OK;
}

// parameters -> packed bf16 operands (forward + dgrad), packed biases; zero the flat gradient buffer
int train_prologue(rac_handle* h, TrainState* T, cudaStream_t st) {
  // ---- parameters -> packed bf16 operands (forward + dgrad), packed biases; zero the gradient accumulators
  for (int i = 1; i < T->nlayers; ++i) {
    TLayer& L = T->L[i];
    // (a layer the fused optimizer step has just written keeps its operand; its bias is re-gathered below)
    if (!T->packed_valid[i])
      CK(launch_pack_weights(T->params, L.d.row_off, L.d.col_off, L.n_packed, L.taps, L.ctot, L.d.flip, L.wp, st, T->w_tiled));
    if (!T->dgrad_bt) CK(launch_transpose_flip(L.wp, L.n_packed, L.taps, L.ctot, L.kpad, L.wd, st));
    if (L.d.bias_off) CK(launch_gather_f32(T->params, L.d.bias_off, L.n_packed, L.bias, st));
  }
  CK(launch_pack_first(T->params + T->L[RAC_L_ENC_C1_0].d.w_off, h->enc_cin, T->wfirst, st));
  CK(cudaMemsetAsync(T->grads, 0, sizeof(float) * T->cfg.n_params, st));
  return RAC_OK;
}

void train_free(rac_handle* h) {
  TrainState* T = static_cast<TrainState*>(h->train);
  if (!T) return;
  if (T->arena) cudaFree(T->arena);
  delete T;
  h->train = nullptr;
}

}  // namespace

extern "C" {

int rac_train_create(rac_handle* h, const rac_train_config* cfg, const rac_train_layer* layers, float* params,
                     float* buffers, float* grads, float* adam_m, float* adam_v) {
  if (!h || !cfg || !layers || !params || !buffers || !grads || !adam_m || !adam_v)
    return fail(h, RAC_ERR_INVALID, "rac_train_create: null argument");
  if (cfg->batch < 1 || cfg->steps < 1 || (cfg->batch * 48) % 64 != 0)
    return fail(h, RAC_ERR_INVALID, "training batch must be a positive multiple of 4 (rows per map must be 64-aligned)");
  if (h->cfg.g_dim % 128 != 0) return fail(h, RAC_ERR_INVALID, "training needs g_dim %% 128 == 0");
  train_free(h);
  TrainState* T = new TrainState();
  h->train = T;
  T->cfg = *cfg;
  T->params = params; T->buffers = buffers; T->grads = grads; T->m = adam_m; T->v = adam_v;
  if (const char* dk = getenv("RAC_TRAIN_DEBUG_KEEP")) T->dbg_keep = atoi(dk);
  if (const char* ps = getenv("RAC_TRAIN_PER_STEP")) T->per_step = atoi(ps);
  if (h->cfg.conv_impl == 1) T->dgrad_bt = 0;
  if (const char* wd = getenv("RAC_DGRAD_WD")) T->dgrad_bt = atoi(wd) ? 0 : 1;
  if (const char* sk = getenv("RAC_TRAIN_SPLITK")) T->splitk = atoi(sk);
  if (const char* lf = getenv("RAC_TRAIN_LSTM_FUSED")) T->lstm_fused = atoi(lf);
  if (const char* tm = getenv("RAC_TRAIN_TILE_MODEL")) T->tile_model = atoi(tm);
  if (const char* wt = getenv("RAC_TRAIN_W_TILED")) T->w_tiled = atoi(wt);
  if (const char* wp = getenv("RAC_TRAIN_W_PREFETCH")) T->w_prefetch = atoi(wp);
  if (!T->dgrad_bt || h->cfg.conv_impl == 1) T->w_tiled = 0;  // (the transposed copy and the SIMT kernels index Wp[n][tap][c])
  if (!T->tile_model) T->lstm_fused = 1;
  if (h->cfg.conv_impl == 1) T->lstm_fused = 1;  // (the SIMT cross-check kernels have no split-K items)
  {
    static bool attr = false;
    if (!attr) { CK(wgrad_tc_set_attributes()); attr = true; }
  }
  T->gn = h->cfg.lstm_group_norm != 0;
  T->nlayers = T->gn ? RAC_L_COUNT_GN : RAC_L_COUNT;
  const int B = cfg->batch, S = cfg->steps, g = h->cfg.g_dim, z = h->cfg.z_dim;
  const size_t M0 = static_cast<size_t>(B) * 3072, M1 = static_cast<size_t>(B) * 768, M2 = static_cast<size_t>(B) * 192,
               M3 = static_cast<size_t>(B) * 48;
  T->M[0] = static_cast<int>(M0); T->M[1] = static_cast<int>(M1); T->M[2] = static_cast<int>(M2); T->M[3] = static_cast<int>(M3);
  for (int i = 0; i < T->nlayers; ++i) {
    TLayer& L = T->L[i];
    L.d = layers[i];
    const LayerSpec& sp = h->spec[i];
    L.taps = sp.ks * sp.ks; L.ctot = sp.ctot; L.n_packed = sp.n_packed; L.kpad = round_up(sp.n_packed, 64);
    if (i >= 1 && L.d.w_count > 0 && L.d.row_off && L.d.col_off)
      CK(adam_pack_vec_ok(L.d.row_off, L.d.col_off, L.n_packed, L.taps, L.ctot, &L.vec_ok));
  }
  // geometry (rows) per layer for scratch sizing
  auto rows_of = [&](int layer) -> size_t {
    switch (layer) {
      case RAC_L_ENC_C1_0: case RAC_L_ENC_C1_1: case RAC_L_DEC_UPC5_0: case RAC_L_DEC_UPC5_1: return M0;
      case RAC_L_ENC_C2_0: case RAC_L_ENC_C2_1: case RAC_L_DEC_UPC4_0: case RAC_L_DEC_UPC4_1: return M1;
      case RAC_L_ENC_C3_0: case RAC_L_ENC_C3_1: case RAC_L_ENC_C3_2: case RAC_L_DEC_UPC3_0: case RAC_L_DEC_UPC3_1:
      case RAC_L_DEC_UPC3_2: return M2;
      default: return M3;
    }
  };
  for (int pass = 0; pass < 2; ++pass) {
    Bump bp;
    bp.base = pass ? static_cast<char*>(T->arena) : nullptr;
    for (int i = 1; i < T->nlayers; ++i) {
      TLayer& L = T->L[i];
      L.wp = bp.take<bf16>(static_cast<size_t>(L.n_packed) * L.taps * L.ctot);
      // dyT: the output gradients of all time steps [steps][M][kpad] -- the dgrad operand of each step and, all steps
      // together, the A operand of the weight-gradient GEMM
      L.dyT = bp.take<bf16>(static_cast<size_t>(L.kpad) * S * rows_of(i));
      if (!T->dgrad_bt) L.wd = bp.take<bf16>(static_cast<size_t>(L.ctot) * L.taps * L.kpad);
      L.dwp = bp.take<float>(static_cast<size_t>(L.kpad) * L.taps * L.ctot);
      L.bias = bp.take<float>(L.n_packed);
    }
    T->wfirst = bp.take<float>(45 * 64);
    T->zero64 = bp.take<float>(64);
    T->tape.resize(S);
    // time-major tape: one allocation [S][n] per tensor, step t = slice t
    auto TA = [&](auto get, size_t n) {
      using P = std::remove_reference_t<decltype(get(T->tape[0]))>;
      using E = std::remove_pointer_t<P>;
      E* b = bp.take<E>(static_cast<size_t>(S) * n);
      for (int t = 0; t < S; ++t) get(T->tape[t]) = b ? b + static_cast<size_t>(t) * n : nullptr;
    };
#define TAPE(field, n) TA([&](Tape& tp) -> decltype(tp.field)& { return tp.field; }, (n))
    TAPE(img4, M0 * 4);
    TAPE(a1, M0 * 64); TAPE(cat5, M0 * 128); TAPE(p1, M1 * 64);
    TAPE(a2, M1 * 128); TAPE(cat4, M1 * 256); TAPE(p2, M2 * 128);
    TAPE(a3a, M2 * 256); TAPE(a3b, M2 * 256); TAPE(cat3, M2 * 512);
    TAPE(p3, M3 * 256); TAPE(a4a, M3 * 512); TAPE(a4b, M3 * 512);
    TAPE(h4, M3 * g); TAPE(aux, M3 * 64); TAPE(auxp, M3 * 64);
    TAPE(pin, M3 * g); TAPE(postin, M3 * g); TAPE(fin, M3 * g);
    TAPE(z, M3 * 64); TAPE(zprior, M3 * 64);
    for (int s = 0; s < 3; ++s)
      for (int l = 0; l < 2; ++l) {
        TAPE(hs[s][l], M3 * g);
        TAPE(cs[s][l], M3 * g);
        TAPE(gates[s][l], M3 * 4 * g);
        if (T->gn) {
          TAPE(raw_ih[s][l], M3 * 4 * g); TAPE(raw_hh[s][l], M3 * 4 * g);
          TAPE(c_raw[s][l], M3 * g); TAPE(gn_stats[s][l], static_cast<size_t>(B) * 96);
        }
      }
    TAPE(d2a, M3 * 512); TAPE(d2b, M3 * 512); TAPE(d3a, M2 * 256);
    TAPE(d3b, M2 * 256); TAPE(d4a, M1 * 128); TAPE(d5, M0 * 64);
    if (cfg->fixed_skip) {
      // every step decodes from its own concat buffers (skip halves = copies of step 0's encoder outputs); step 0 too,
      // so that the decoder inputs of all steps are one time-major tensor for the weight gradient
      TAPE(dcat5, M0 * 128); TAPE(dcat4, M1 * 256); TAPE(dcat3, M2 * 512);
    } else {
      for (int t = 0; t < S; ++t) {
        Tape& tp = T->tape[t];
        tp.dcat5 = tp.cat5; tp.dcat4 = tp.cat4; tp.dcat3 = tp.cat3;
      }
    }
    for (int i = 0; i < 19; ++i) {
      const VggDef d = i < 10 ? enc_def(h, i) : dec_def(h, i - 10);
      TAPE(vgg[i].raw, static_cast<size_t>(B) * d.H * d.W * d.cout);
      TAPE(vgg[i].mean, static_cast<size_t>(d.cout));
      TAPE(vgg[i].rstd, static_cast<size_t>(d.cout));
    }
    const size_t zn = static_cast<size_t>(B) * z * 48;
    TAPE(mu_p, zn); TAPE(lv_p, zn); TAPE(mu, zn); TAPE(lv, zn);
    TAPE(eps_p, zn); TAPE(eps_q, zn);
    TAPE(x4, M0 * 4);
    TAPE(xp, M0 * 3);
#undef TAPE
    auto GB = [&](GBuf& b, size_t n) { b.n = n; b.p = bp.take<float>(static_cast<size_t>(S) * n); };
    GB(T->G_d5, M0 * 64); GB(T->G_cat5, M0 * 128); GB(T->G_d4a, M1 * 128);
    GB(T->G_cat4, M1 * 256); GB(T->G_d3b, M2 * 256); GB(T->G_d3a, M2 * 256);
    GB(T->G_cat3, M2 * 512); GB(T->G_d2b, M3 * 512); GB(T->G_d2a, M3 * 512);
    GB(T->G_fin, M3 * g); GB(T->G_pin, M3 * g); GB(T->G_postin, M3 * g);
    GB(T->G_z, M3 * 64); GB(T->G_h4, M3 * g);
    GB(T->G_a4b, M3 * 512); GB(T->G_a4a, M3 * 512); GB(T->G_p3, M3 * 256);
    GB(T->G_a3b, M2 * 256); GB(T->G_a3a, M2 * 256); GB(T->G_p2, M2 * 128);
    GB(T->G_a2, M1 * 128); GB(T->G_p1, M1 * 64); GB(T->G_a1, M0 * 64);
    GB(T->draw32, M0 * 64);
    T->G_skip5 = bp.take<float>(M0 * 64); T->G_skip4 = bp.take<float>(M1 * 128); T->G_skip3 = bp.take<float>(M2 * 256);
    for (int s = 0; s < 3; ++s)
      for (int l = 0; l < 2; ++l) {
        GB(T->DH[s][l], M3 * g);
        T->G_dc[s][l] = bp.take<float>(M3 * g);
      }
    T->bn_scratch = bp.take<float>(static_cast<size_t>(S) * 2 * 2048);
    T->fw_part = bp.take<float>(static_cast<size_t>(kFwBlocks) * 45 * 64);
    T->hzero = bp.take<bf16>(M3 * g); T->czero = bp.take<float>(M3 * g);
    T->G_img[0] = bp.take<float>(M0 * 3); T->G_img[1] = bp.take<float>(M0 * 3);
    T->dbg_draw32 = bp.take<float>(M0 * 64);
    T->loss_part = bp.take<float>(static_cast<size_t>(S) * B);
    T->metric_part = bp.take<float>(std::max<size_t>(static_cast<size_t>(2) * S * B, 64));
    if (T->gn) {
      T->gn_dy = bp.take<float>(M3 * 4 * g);
      T->gn_part = bp.take<float>(static_cast<size_t>(B) * 14 * g);
    }
    {
      // split-K scratch of the weight gradient: the largest splits x |dWp| over the layers that need more than one slice
      size_t need = 4;
      for (int i = 1; i < T->nlayers; ++i) {
        const TLayer& L = T->L[i];
        const size_t rows = rows_of(i);
        const int W = rows == M0 ? 64 : rows == M1 ? 32 : rows == M2 ? 16 : 8;
        WgradGeom wg = wg_plan(L, B, W * 3 / 4, W, S);
        wg.num_ctiles = std::max(1, L.ctot / 256);  // (a lower bound on the tile count: more tiles = fewer splits)
        wg_split(wg, h->num_sms);
        if (wg.splits > 1) need = std::max(need, static_cast<size_t>(wg.splits) * static_cast<size_t>(wg.out_split_stride));
      }
      T->wg_part_elems = need;
      T->wg_part = bp.take<float>(need);
    }
    // split-K slices of the small-M GEMMs: up to 16 slices of one step's widest output (the 4g gate pre-activations)
    T->sk_part_elems = static_cast<size_t>(16) * M3 * 4 * g;
    T->sk_part = bp.take<float>(T->sk_part_elems);
    T->lstm_gx = bp.take<float>(static_cast<size_t>(S) * M3 * 4 * g);
    if (!pass) {
      CK(cudaMalloc(&T->arena, bp.off + 1024));
      CK(cudaMemset(T->arena, 0, bp.off + 1024));
    }
  }
  return RAC_OK;
}

int rac_train_destroy(rac_handle* h) {
  if (!h) return RAC_ERR_INVALID;
  train_free(h);
  return RAC_OK;
}

int rac_train_forward_backward(rac_handle* h, const rac_train_batch* bt, void* stream) {
  if (!h || !bt || !h->train) return fail(h, RAC_ERR_STATE, "rac_train_create first");
  TrainState* T = static_cast<TrainState*>(h->train);
  const rac_config& c = h->cfg;
  if (!bt->images || !bt->actions || !bt->losses) return fail(h, RAC_ERR_INVALID, "images, actions, losses are required");
  if (T->cfg.recon_kind < 0 || T->cfg.recon_kind > 3) return fail(h, RAC_ERR_INVALID, "recon_kind %d", T->cfg.recon_kind);
  if ((c.use_mask || T->cfg.zero_robot || (T->cfg.recon_kind & 1)) && !bt->masks)
    return fail(h, RAC_ERR_INVALID, "this configuration needs masks");
  if (c.use_robot_state && !bt->states) return fail(h, RAC_ERR_INVALID, "model_use_robot_state needs states");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int B = T->cfg.batch, S = T->cfg.steps, g = c.g_dim, z = c.z_dim;
  const size_t M3 = static_cast<size_t>(B) * 48;
  CKR(train_prologue(h, T, st));
  CK(cudaMemsetAsync(bt->losses, 0, sizeof(float) * 4, st));
  T->cur_bt = bt;
  for (bool& u : T->unpacked) u = false;
  for (bool& u : T->deferred) u = false;
  T->step_api = 0;
  T->active_steps = S;
  // ---- reparameterisation noise of all steps (the tape is time-major: one copy / one fill per tensor)
  const size_t zn = static_cast<size_t>(B) * z * 48;
  // Philox counter = the caller's global training step (persisted in checkpoints), NOT a count kept in this state:
  // rac_train_create runs again after a resume or a new batch shape and must not replay the noise of steps 0..k
  const unsigned int noise_step = static_cast<unsigned int>(bt->noise_step);
  if (bt->eps_prior) CK(cudaMemcpyAsync(T->tape[0].eps_p, bt->eps_prior, sizeof(float) * zn * S, cudaMemcpyDeviceToDevice, st));
  else CK(launch_normal_fill(T->tape[0].eps_p, static_cast<long long>(zn) * S, bt->seed, 2u * noise_step, st));
  if (bt->eps_post) CK(cudaMemcpyAsync(T->tape[0].eps_q, bt->eps_post, sizeof(float) * zn * S, cudaMemcpyDeviceToDevice, st));
  else CK(launch_normal_fill(T->tape[0].eps_q, static_cast<long long>(zn) * S, bt->seed, 2u * noise_step + 1u, st));
  // Teacher-forced clip (every input frame is ground truth) with per-step skips: nothing but the ConvLSTM cells depends
  // on the previous time step, so every other layer runs ONCE over all S * B images. A step that consumes the model's
  // own prediction (scheduled sampling), or fixed_skip (whose skip gradients are summed over time), goes step by step
  bool batched = !T->per_step && !T->cfg.fixed_skip;
  if (bt->true_token)
    for (int t = 1; t < S; ++t) batched = batched && bt->true_token[t] != 0;
  const Span all{0, S};
  // ---- forward
  if (batched) {
    CKR(forward_encoder(h, T, bt, all, st));
    CKR(forward_input_convs(h, T, bt, all, st));
    CKR(lstm_forward_span(h, T, 0, all, T->tape[0].pin, st));
    CKR(forward_gauss(h, T, 0, all, st));
    CKR(lstm_forward_span(h, T, 1, all, T->tape[0].postin, st));
    CKR(forward_gauss(h, T, 1, all, st));
    CKR(forward_fp_in(h, T, all, st));
    CKR(lstm_forward_span(h, T, 2, all, T->tape[0].fin, st));
    CKR(forward_decoder(h, T, bt, all, st));
  } else {
    for (int t = 0; t < S; ++t) {
      const Span one{t, 1};
      CKR(forward_encoder(h, T, bt, one, st));
      CKR(forward_input_convs(h, T, bt, one, st));
      CKR(lstm_forward_span(h, T, 0, one, T->tape[t].pin, st));
      CKR(forward_gauss(h, T, 0, one, st));
      CKR(lstm_forward_span(h, T, 1, one, T->tape[t].postin, st));
      CKR(forward_gauss(h, T, 1, one, st));
      CKR(forward_fp_in(h, T, one, st));
      CKR(lstm_forward_span(h, T, 2, one, T->tape[t].fin, st));
      CKR(forward_decoder(h, T, bt, one, st));
    }
  }
  // ---- BPTT
  for (int s = 0; s < 3; ++s)
    for (int l = 0; l < 2; ++l) {
      CK(cudaMemsetAsync(T->DH[s][l].p, 0, sizeof(float) * M3 * g * S, st));
      CK(cudaMemsetAsync(T->G_dc[s][l], 0, sizeof(float) * M3 * g, st));
    }
  if (batched) {
    CKR(backward_decoder(h, T, bt, all, st));
    for (int t = S - 1; t >= 0; --t) CKR(lstm_backward(h, T, 2, t, T->tape[t].fin, T->G_fin.at(t), st));
    CKR(backward_fp_in(h, T, all, st));
    CKR(backward_gauss(h, T, all, st));
    for (int t = S - 1; t >= 0; --t) CKR(lstm_backward(h, T, 1, t, T->tape[t].postin, T->G_postin.at(t), st));
    CKR(backward_input_conv(h, T, 1, all, st));
    for (int t = S - 1; t >= 0; --t) CKR(lstm_backward(h, T, 0, t, T->tape[t].pin, T->G_pin.at(t), st));
    CKR(backward_input_conv(h, T, 0, all, st));
    CKR(backward_encoder(h, T, bt, all, st));
  } else {
    for (int t = S - 1; t >= 0; --t) {
      const Span one{t, 1};
      CKR(backward_decoder(h, T, bt, one, st));
      CKR(lstm_backward(h, T, 2, t, T->tape[t].fin, T->G_fin.at(t), st));
      CKR(backward_fp_in(h, T, one, st));
      CKR(backward_gauss(h, T, one, st));
      CKR(lstm_backward(h, T, 1, t, T->tape[t].postin, T->G_postin.at(t), st));
      CKR(backward_input_conv(h, T, 1, one, st));
      CKR(lstm_backward(h, T, 0, t, T->tape[t].pin, T->G_pin.at(t), st));
      CKR(backward_input_conv(h, T, 0, one, st));
      CKR(backward_encoder(h, T, bt, one, st));
    }
  }
  // ---- packed weight gradients -> flat parameter layout (layers not handed over earlier)
  for (int i = 1; i < T->nlayers; ++i) {
    TLayer& L = T->L[i];
    if (T->unpacked[i] || T->deferred[i]) continue;
    CK(launch_unpack_grads(L.dwp, L.d.row_off, L.d.col_off, L.n_packed, L.taps, L.ctot, L.d.flip, T->grads, st));
  }
  T->cur_bt = nullptr;
  return RAC_OK;
}

// ------------------------------------------------------------------ step API
namespace {
int step_io_from(rac_handle* h, TrainState* T, const rac_train_step* io) {
  const rac_config& c = h->cfg;
  const long long HW = 48 * 64;
  if (!io || !io->image || !io->action) return fail(h, RAC_ERR_INVALID, "train step: image and action are required");
  if (c.use_mask && !io->mask) return fail(h, RAC_ERR_INVALID, "train step: model_use_mask needs mask");
  if (c.use_robot_state && (!io->robot || !io->next_robot))
    return fail(h, RAC_ERR_INVALID, "train step: model_use_robot_state needs robot and next_robot");
  StepIO s{};
  s.x_j = io->image;
  s.m_j = io->mask;
  s.m_i = (io->mask && c.use_future_mask) ? io->mask + HW : nullptr;  // mask = cat([m_j, m_i], 1) (trainer.py:373-375)
  s.mask_bstride = (c.use_mask && c.use_future_mask) ? 2 * HW : HW;
  s.r_j = io->robot;
  s.r_i = io->next_robot;
  s.a_j = io->action;
  s.x_i = nullptr;
  T->io = s;
  return RAC_OK;
}
}  // namespace

int rac_train_step_begin(rac_handle* h, void* stream) {
  if (!h || !h->train) return fail(h, RAC_ERR_STATE, "rac_train_create first");
  TrainState* T = static_cast<TrainState*>(h->train);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  for (bool& u : T->packed_valid) u = false;  // (an external optimizer has stepped the parameters)
  CKR(train_prologue(h, T, st));
  T->step_api = 1;
  T->active_steps = 0;
  T->bwd_next = -1;
  return RAC_OK;
}

int rac_train_step_forward(rac_handle* h, const rac_train_step* io, void* stream) {
  if (!h || !h->train) return fail(h, RAC_ERR_STATE, "rac_train_create first");
  TrainState* T = static_cast<TrainState*>(h->train);
  if (!T->step_api) return fail(h, RAC_ERR_STATE, "rac_train_step_begin first");
  if (T->bwd_next >= 0) return fail(h, RAC_ERR_STATE, "train step: forward after the backward pass has started (rac_train_step_begin first)");
  const int t = T->active_steps;
  if (t >= T->cfg.steps)
    return fail(h, RAC_ERR_STATE, "train step: %d forward calls since rac_train_step_begin, the tape holds %d (n_past + n_future - 1)", t + 1, T->cfg.steps);
  // last_frame_skip False: step 0 decodes with its own skips, every later step with step 0's (dynamics.py:586-588)
  if ((io && io->keep_skip != 0) != (T->cfg.fixed_skip != 0 && t > 0))
    return fail(h, RAC_ERR_INVALID, "train step %d: keep_skip %d does not match fixed_skip %d", t, io ? io->keep_skip : 0, T->cfg.fixed_skip);
  CKR(step_io_from(h, T, io));
  if (!io->x_pred || !io->mu || !io->logvar || !io->mu_p || !io->logvar_p) return fail(h, RAC_ERR_INVALID, "train step: null output");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const rac_config& c = h->cfg;
  const size_t zn = static_cast<size_t>(T->cfg.batch) * c.z_dim * 48, M0 = static_cast<size_t>(T->cfg.batch) * 3072;
  Tape& tp = T->tape[t];
  const unsigned int ns = static_cast<unsigned int>(io->noise_step);
  if (io->eps_prior) CK(cudaMemcpyAsync(tp.eps_p, io->eps_prior, sizeof(float) * zn, cudaMemcpyDeviceToDevice, st));
  else CK(launch_normal_fill(tp.eps_p, static_cast<long long>(zn), io->seed, 2u * ns, st));
  if (io->eps_post) CK(cudaMemcpyAsync(tp.eps_q, io->eps_post, sizeof(float) * zn, cudaMemcpyDeviceToDevice, st));
  else CK(launch_normal_fill(tp.eps_q, static_cast<long long>(zn), io->seed, 2u * ns + 1u, st));
  const Span one{t, 1};
  CKR(forward_encoder(h, T, nullptr, one, st));
  CKR(forward_input_convs(h, T, nullptr, one, st));
  CKR(lstm_forward_span(h, T, 0, one, tp.pin, st));
  CKR(forward_gauss(h, T, 0, one, st));
  CKR(lstm_forward_span(h, T, 1, one, tp.postin, st));
  CKR(forward_gauss(h, T, 1, one, st));
  CKR(forward_fp_in(h, T, one, st));
  CKR(lstm_forward_span(h, T, 2, one, tp.fin, st));
  CKR(forward_decoder(h, T, nullptr, one, st));
  CK(cudaMemcpyAsync(io->x_pred, tp.x4, sizeof(float) * M0 * 4, cudaMemcpyDeviceToDevice, st));
  CK(cudaMemcpyAsync(io->mu, tp.mu, sizeof(float) * zn, cudaMemcpyDeviceToDevice, st));
  CK(cudaMemcpyAsync(io->logvar, tp.lv, sizeof(float) * zn, cudaMemcpyDeviceToDevice, st));
  CK(cudaMemcpyAsync(io->mu_p, tp.mu_p, sizeof(float) * zn, cudaMemcpyDeviceToDevice, st));
  CK(cudaMemcpyAsync(io->logvar_p, tp.lv_p, sizeof(float) * zn, cudaMemcpyDeviceToDevice, st));
  T->active_steps = t + 1;
  return RAC_OK;
}

int rac_train_step_backward(rac_handle* h, int t, const rac_train_step* io, const float* d_x_pred, const float* d_mu,
                            const float* d_logvar, const float* d_mu_p, const float* d_logvar_p, float* d_image,
                            void* stream) {
  if (!h || !h->train) return fail(h, RAC_ERR_STATE, "rac_train_create first");
  TrainState* T = static_cast<TrainState*>(h->train);
  if (!T->step_api || T->active_steps < 1) return fail(h, RAC_ERR_STATE, "train step backward: no forward step on the tape");
  if (T->bwd_next == -2) return fail(h, RAC_ERR_STATE, "train step backward: this tape has been consumed (one backward pass per rac_train_step_begin)");
  if (T->bwd_next == -1) {
    // first call of the backward pass: the latest step whose outputs reached the loss; later steps got no gradient,
    // so the tape is cut there (their contribution to every parameter gradient is zero)
    if (t < 0 || t >= T->active_steps) return fail(h, RAC_ERR_INVALID, "train step backward: step %d of %d", t, T->active_steps);
    T->active_steps = t + 1;
  } else if (t != T->bwd_next) {
    return fail(h, RAC_ERR_STATE, "train step backward: step %d requested, step %d is next (BPTT runs from the last step down to 0)", t, T->bwd_next);
  }
  const int S = T->active_steps;
  CKR(step_io_from(h, T, io));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int B = T->cfg.batch, g = h->cfg.g_dim;
  const size_t M3 = static_cast<size_t>(B) * 48;
  if (t == S - 1)
    for (int s = 0; s < 3; ++s)
      for (int l = 0; l < 2; ++l) {
        CK(cudaMemsetAsync(T->DH[s][l].p, 0, sizeof(float) * M3 * g * T->cfg.steps, st));
        CK(cudaMemsetAsync(T->G_dc[s][l], 0, sizeof(float) * M3 * g, st));
      }
  T->ext_dx4 = d_x_pred; T->ext_dmu = d_mu; T->ext_dlv = d_logvar; T->ext_dmu_p = d_mu_p; T->ext_dlv_p = d_logvar_p;
  T->ext_dimg = d_image;
  Tape& tp = T->tape[t];
  const Span one{t, 1};
  CKR(backward_decoder(h, T, nullptr, one, st));
  CKR(lstm_backward(h, T, 2, t, tp.fin, T->G_fin.at(t), st));
  CKR(backward_fp_in(h, T, one, st));
  CKR(backward_gauss(h, T, one, st));
  CKR(lstm_backward(h, T, 1, t, tp.postin, T->G_postin.at(t), st));
  CKR(backward_input_conv(h, T, 1, one, st));
  CKR(lstm_backward(h, T, 0, t, tp.pin, T->G_pin.at(t), st));
  CKR(backward_input_conv(h, T, 0, one, st));
  CKR(backward_encoder(h, T, nullptr, one, st));
  T->ext_dimg = nullptr;
  if (t == 0) {
    for (int i = 1; i < T->nlayers; ++i) {
      TLayer& L = T->L[i];
      CK(launch_unpack_grads(L.dwp, L.d.row_off, L.d.col_off, L.n_packed, L.taps, L.ctot, L.d.flip, T->grads, st));
    }
    T->bwd_next = -2;
  } else {
    T->bwd_next = t - 1;
  }
  return RAC_OK;
}

int rac_train_debug_buffer(rac_handle* h, const char* name, int step, void** ptr) {
  if (!h || !h->train || !name || !ptr) return RAC_ERR_INVALID;
  TrainState* T = static_cast<TrainState*>(h->train);
  if (step < 0 || step >= static_cast<int>(T->tape.size())) return fail(h, RAC_ERR_INVALID, "bad step %d", step);
  Tape& tp = T->tape[step];
  // gradient accumulators: the slot of time step `step`
  struct { const char* n; void* p; } tab[] = {
      {"G_d5", T->G_d5.at(step)}, {"G_cat5", T->G_cat5.at(step)}, {"G_d4a", T->G_d4a.at(step)}, {"G_cat4", T->G_cat4.at(step)}, {"G_d3b", T->G_d3b.at(step)},
      {"G_d3a", T->G_d3a.at(step)}, {"G_cat3", T->G_cat3.at(step)}, {"G_d2b", T->G_d2b.at(step)}, {"G_d2a", T->G_d2a.at(step)}, {"G_fin", T->G_fin.at(step)},
      {"G_pin", T->G_pin.at(step)}, {"G_postin", T->G_postin.at(step)}, {"G_z", T->G_z.at(step)}, {"G_h4", T->G_h4.at(step)}, {"G_a4b", T->G_a4b.at(step)},
      {"G_a4a", T->G_a4a.at(step)}, {"G_p3", T->G_p3.at(step)}, {"G_a3b", T->G_a3b.at(step)}, {"G_a3a", T->G_a3a.at(step)}, {"G_p2", T->G_p2.at(step)},
      {"G_a2", T->G_a2.at(step)}, {"G_p1", T->G_p1.at(step)}, {"G_a1", T->G_a1.at(step)},
      {"img4", tp.img4}, {"a1", tp.a1}, {"cat5", tp.cat5}, {"p1", tp.p1}, {"a2", tp.a2}, {"cat4", tp.cat4},
      {"p2", tp.p2}, {"a3a", tp.a3a}, {"a3b", tp.a3b}, {"cat3", tp.cat3}, {"p3", tp.p3}, {"a4a", tp.a4a},
      {"a4b", tp.a4b}, {"h4", tp.h4}, {"d2a", tp.d2a}, {"d2b", tp.d2b}, {"d3a", tp.d3a}, {"d3b", tp.d3b},
      {"d4a", tp.d4a}, {"d5", tp.d5}, {"dcat5", tp.dcat5}, {"dcat4", tp.dcat4}, {"dcat3", tp.dcat3},
      {"G_skip5", T->G_skip5}, {"G_skip4", T->G_skip4}, {"G_skip3", T->G_skip3}, {"x4", tp.x4}, {"hfp1", tp.hs[2][1]}, {"xp", tp.xp},
      {"G_img0", T->G_img[0]}, {"G_img1", T->G_img[1]}, {"dbg_draw32", T->dbg_draw32}};
  for (auto& e : tab)
    if (!strcmp(e.n, name)) { *ptr = e.p; return RAC_OK; }
  if (!strncmp(name, "raw", 3)) {
    const int i = atoi(name + 3);
    if (i >= 0 && i < 19) { *ptr = tp.vgg[i].raw; return RAC_OK; }
  }
  return fail(h, RAC_ERR_INVALID, "no training buffer named '%s'", name);
}

int rac_train_set_adam_step(rac_handle* h, int steps_taken) {
  if (!h || !h->train) return fail(h, RAC_ERR_STATE, "rac_train_create first");
  if (steps_taken < 0) return fail(h, RAC_ERR_INVALID, "negative Adam step count %d", steps_taken);
  static_cast<TrainState*>(h->train)->adam_t = steps_taken;
  return RAC_OK;
}

int rac_train_adam_step(rac_handle* h, void* stream) {
  if (!h || !h->train) return fail(h, RAC_ERR_STATE, "rac_train_create first");
  TrainState* T = static_cast<TrainState*>(h->train);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  T->adam_t += 1;
  // layers whose gradient stayed packed (rac_train_batch.defer_unpack): one fused pass each; everything else -- the
  // complement of their weight ranges in the flat buffers -- through the flat kernel
  std::vector<std::pair<long long, long long>> fused;
  for (int i = 1; i < T->nlayers; ++i) {
    TLayer& L = T->L[i];
    if (!T->deferred[i]) { T->packed_valid[i] = false; continue; }
    CK(launch_adam_pack(T->params, T->m, T->v, L.dwp, L.d.row_off, L.d.col_off, L.n_packed, L.taps, L.ctot, L.d.flip,
                        T->w_tiled, L.wp, T->cfg.lr, T->cfg.beta1, T->cfg.beta2, T->cfg.adam_eps, T->adam_t, T->grad_scale, st,
                        L.vec_ok));
    h->launches++;
    T->packed_valid[i] = true;
    T->deferred[i] = false;
    fused.push_back({L.d.w_off, L.d.w_count});
  }
  std::sort(fused.begin(), fused.end());
  long long pos = 0;
  fused.push_back({T->cfg.n_params, 0});
  for (const auto& f : fused) {
    if (f.first < pos) return fail(h, RAC_ERR_STATE, "overlapping weight ranges in the fused optimizer step");
    if (f.first > pos)
      CK(launch_adam(T->params + pos, T->grads + pos, T->m + pos, T->v + pos, f.first - pos, T->cfg.lr, T->cfg.beta1,
                     T->cfg.beta2, T->cfg.adam_eps, T->adam_t, st, T->grad_scale));
    pos = f.first + f.second;
  }
  return RAC_OK;
}

int rac_train_unpack_deferred(rac_handle* h, void* stream) {
  if (!h || !h->train) return fail(h, RAC_ERR_STATE, "rac_train_create first");
  TrainState* T = static_cast<TrainState*>(h->train);
  for (int i = 1; i < T->nlayers; ++i) {
    TLayer& L = T->L[i];
    if (!T->deferred[i]) continue;
    CK(launch_unpack_grads(L.dwp, L.d.row_off, L.d.col_off, L.n_packed, L.taps, L.ctot, L.d.flip, T->grads,
                           static_cast<cudaStream_t>(stream)));
  }
  return RAC_OK;
}

int rac_train_invalidate_packed(rac_handle* h) {
  if (!h) return RAC_ERR_INVALID;
  if (!h->train) return RAC_OK;
  for (bool& u : static_cast<TrainState*>(h->train)->packed_valid) u = false;
  return RAC_OK;
}

int rac_train_set_grad_scale(rac_handle* h, float scale) {
  if (!h || !h->train) return fail(h, RAC_ERR_STATE, "rac_train_create first");
  if (!(scale > 0.f)) return fail(h, RAC_ERR_INVALID, "gradient scale %g", scale);
  static_cast<TrainState*>(h->train)->grad_scale = scale;
  return RAC_OK;
}

}  // extern "C"
