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
// Not implemented: heatmaps, multiview, batch_weight.

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
  // wgrad operands of ALL time steps, transposed: xcolT [taps*ctot][steps*mpad], dyT [kpad][steps*mpad]. The weight
  // gradient is a sum over time, so the contraction simply runs over steps * rows: ONE GEMM per layer with K = steps *
  // mpad after the last BPTT step instead of one K = mpad GEMM (+ a read-modify-write of dW) per step.
  bf16* xcolT = nullptr;
  bf16* dyT = nullptr;
  int taps = 0, ctot = 0, n_packed = 0, kpad = 0;
};

struct VggRt {  // one vgg_layer (conv3x3 no bias + BatchNorm + LeakyReLU) at one time step
  float* raw = nullptr;
  float* mean = nullptr;
  float* rstd = nullptr;
};

struct Tape {  // everything the backward pass of one time step needs
  float* img4;
  bf16 *a1, *cat5, *p1, *a2, *cat4, *p2, *a3a, *a3b, *cat3, *p3, *a4a, *a4b, *h4;
  bf16 *aux, *auxp, *pin, *postin, *fin, *z, *zprior;
  bf16* hs[3][2];
  float* cs[3][2];
  float* gates[3][2];
  float *raw_ih[3][2], *raw_hh[3][2], *c_raw[3][2], *gn_stats[3][2];  // lstm_group_norm only
  bf16 *d2a, *d2b, *d3a, *d3b, *d4a, *d5;
  bf16 *dcat5, *dcat4, *dcat3;  // decoder concat buffers: == cat* unless fixed_skip and t > 0
  VggRt vgg[19];
  float *mu_p, *lv_p, *mu, *lv, *x4;
  float *eps_p, *eps_q;
  float* xp;          // composited prediction (B,3,H,W): the next step's input under scheduled sampling
  const float* xj;    // this step's input frame (ground truth or the previous step's xp)
  int sampled;        // 1: xj is the model's own previous prediction (gradient flows back, trainer.py:354)
};

struct TrainState {
  rac_train_config cfg{};
  TLayer L[RAC_L_COUNT_GN];
  int nlayers = RAC_L_COUNT;  // RAC_L_COUNT_GN with lstm_group_norm (separate ih / hh gate convolutions)
  bool gn = false;
  float* gn_dy = nullptr;     // [M3, 4g] gate pre-activation gradients of the cell being processed
  float* gn_part = nullptr;   // [B][14 g] per-sample partials of the GroupNorm affine gradients
  float *params = nullptr, *buffers = nullptr, *grads = nullptr, *m = nullptr, *v = nullptr;
  // weight gradients: 1 = implicit GEMM over all time steps straight from the tape (wgrad_tc.cu); 0 = round-1 path
  // (materialised im2col(X)^T / dY^T + plain GEMM; RAC_WGRAD_IM2COL=1, kept for A/B measurements and cross-checks)
  int wgrad_implicit = 1;
  unsigned long long step_bytes = 0;  // distance between the same tape buffer of two consecutive time steps
  float* wg_part = nullptr;           // split-K partials of the layer being processed
  size_t wg_part_elems = 0;
  float* wfirst = nullptr;  // [9*cin][64]
  float* zero64 = nullptr;
  std::vector<Tape> tape;
  void* arena = nullptr;
  // gradient accumulators (fp32 NHWC)
  float *G_d5, *G_cat5, *G_d4a, *G_cat4, *G_d3b, *G_d3a, *G_cat3, *G_d2b, *G_d2a, *G_fin, *G_pin, *G_postin, *G_z, *G_h4;
  float *G_a4b, *G_a4a, *G_p3, *G_a3b, *G_a3a, *G_p2, *G_a2, *G_p1, *G_a1;
  float *G_skip5, *G_skip4, *G_skip3;  // fixed_skip: skip-half gradients summed over the steps ([M, C] fp32)
  float* G_hs[3][2][2];
  float* G_dc[3][2];
  float* G_img[2];       // gradient w.r.t. a sampled input frame, ping-pong across steps
  float* dbg_draw32 = nullptr;  // RAC_TRAIN_DEBUG_KEEP=1: first-layer raw gradient of the last SAMPLED step (tests)
  int dbg_keep = 0;
  bf16 *dy_a, *dy_b;     // bf16 gradient operands (largest [M, C])
  int cur_t = 0;         // time step being processed by the backward pass (column block of the wgrad operands)
  float* bn_scratch;
  float* draw32;         // fp32 copy of the first layer's raw gradient
  float* fw_part;        // per-CTA partial sums of the first layer's weight gradient [kFwBlocks][45 * 64]
  bf16* hzero;
  float* czero;
  float* loss_part;      // [B]
  float* kl_tmp;
  int adam_t = 0;
  int M[4];
};

constexpr int kFwBlocks = 148;  // CTAs of the first layer's weight-gradient kernel (one partial sum each)

inline int pick_bn(int n) { return n % 256 == 0 ? 256 : (n % 128 == 0 ? 128 : 64); }

struct GemmGeom {
  int B, H, W, ks;
  bool plain;  // plain GEMM view: rows = B*64, H = 1
};

// Build + launch one conv_tc GEMM with an explicit operand description (tensor maps encoded on the fly).
int t_gemm(rac_handle* h, const char* name, const GemmGeom& gg, const std::vector<Src>& srcs, const bf16* w, int ktotal,
           int n_rows_w, int block_n, int epi, const EpiParams& ep, cudaStream_t st) {
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
  op.block_m = 256;
  op.block_n = block_n;
  {
    const long long tiles = static_cast<long long>((gg.B + g.NB - 1) / g.NB) * (gg.H / g.BH) * (n_rows_w / block_n);
    if (tiles < 120) {
      op.block_m = 128;
      geom(128);
      if (op.block_n > 128) op.block_n = 128;
      block_n = op.block_n;
    }
  }
  g.nsrc = static_cast<int>(srcs.size());
  for (int i = 0; i < g.nsrc; ++i) {
    if (srcs[i].C % kBlockK) return fail(h, RAC_ERR_INVALID, "train gemm %s: source channels %d", name, srcs[i].C);
    g.src_kb[i] = srcs[i].C / kBlockK;
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
  CKR(encode_w_map(h, &op.tm.w, w, ktotal, n_rows_w, block_n));
  op.e = ep;
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

int vgg_forward(rac_handle* h, TrainState* T, VggRt& rt, const VggDef& d, const bf16* in, bf16* out, int cstride,
                int coff, int up, int updates, cudaStream_t st) {
  const TLayer& L = T->L[d.layer];
  const int B = T->cfg.batch;
  EpiParams e{};
  e.cout = L.n_packed;
  e.nseg = 1;
  e.seg[0] = {0, L.n_packed, rt.raw, d.cout, 0, 0};
  if (L.n_packed != d.cout) return fail(h, RAC_ERR_INVALID, "vgg layer %d: cout %d must be a multiple of the pack granularity", d.layer, d.cout);
  CKR(t_gemm(h, "train.vgg.fwd", {B, d.H, d.W, 3, false}, {{in, d.cin}}, L.wp, 9 * d.cin, L.n_packed,
             pick_bn(L.n_packed), EPI_F32, e, st));
  const int M = B * d.H * d.W;
  CK(launch_bn_stats(rt.raw, M, d.cout, rt.mean, rt.rstd, T->buffers + L.d.rmean_off, T->buffers + L.d.rvar_off, updates, st));
  CK(launch_bn_act(rt.raw, rt.mean, rt.rstd, T->params + L.d.gamma_off, T->params + L.d.beta_off, B, d.H, d.W, d.cout,
                   out, cstride, coff, up, st));
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
  g.n_tiles = (L.kpad + 127) / 128;
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
  const int B = T->cfg.batch, S = T->cfg.steps;
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
    const bf16* base = xs[i].base0 ? xs[i].base0 : xs[i].p;
    CKR(encode_act_map5(h, &tm.x[i], base, xs[i].C, B, H, W, S, T->step_bytes, g.BH, g.NB));
  }
  CK(launch_wgrad_tc(tm, g, st));
  h->launches++;
  if (g.splits > 1) {
    CK(launch_wgrad_reduce(T->wg_part, g.splits, static_cast<long long>(out_elems), g.out_split_stride, L.dwp, st));
    h->launches++;
  }
  return RAC_OK;
}

// dY (bf16 [M, kpad], packed column order) -> weight gradient (accumulated into L.dwp) and input gradient (segments)
int conv_backward(rac_handle* h, TrainState* T, int layer, int H, int W, const std::vector<Src>& xs, const bf16* dY,
                  const F32Seg* segs, int nseg, cudaStream_t st) {
  TLayer& L = T->L[layer];
  const int B = T->cfg.batch;
  const int M = B * H * W, mpad = round_up(M, 64);
  const int ks = (L.taps == 25) ? 5 : 3;
  const int S = T->cfg.steps;
  if (T->wgrad_implicit) {
    // ---- wgrad, implicit GEMM: this step's dY is kept in slot t of the layer's all-steps buffer [S][M][kpad]; the
    // inputs X_t already live on the tape. ONE launch per layer, after the last processed step (t == 0), contracts over
    // the rows of all time steps (the weight gradient is a sum over time)
    CK(cudaMemcpyAsync(L.dyT + static_cast<size_t>(T->cur_t) * M * L.kpad, dY, sizeof(bf16) * static_cast<size_t>(M) * L.kpad,
                       cudaMemcpyDeviceToDevice, st));
    if (T->cur_t == 0) {
      CKR(wgrad_implicit(h, T, L, H, W, xs, st));
      if (L.d.bias_off) CK(launch_bias_grad(L.dyT, S * M, L.kpad, L.n_packed, L.d.bias_off, T->grads, st));
    }
  } else {
    // ---- wgrad (round-1 path): dWp[n][tap*ctot + c] = sum_t sum_m dY_t[m][n] * X_t[shift_tap(m)][c]: this step's
    // operands go to column block t of the layer's transposed buffers; the GEMM runs once, after the last processed step
    const int ld = S * mpad;
    int coff = 0;
    for (const Src& s : xs) {
      CK(launch_im2col_t(s.p, B, H, W, s.C, ks, L.ctot, coff, mpad, L.xcolT + static_cast<size_t>(T->cur_t) * mpad, st, ld));
      coff += s.C;
    }
    CK(launch_transpose_bf16(dY, M, L.kpad, mpad, L.kpad, L.dyT + static_cast<size_t>(T->cur_t) * mpad, st, ld));
    if (T->cur_t == 0) {
      EpiParams e{};
      const int ncols = L.taps * L.ctot;
      e.cout = ncols;
      e.nseg = 1;
      e.seg[0] = {0, ncols, L.dwp, ncols, 0, 1};
      CKR(t_gemm(h, "train.wgrad", {L.kpad / 64, 1, 64, 1, true}, {{L.dyT, ld}}, L.xcolT, ld, ncols, pick_bn(ncols),
                 EPI_F32, e, st));
      // bias gradient = row sums of the all-time-steps dY^T: once, after the last processed step
      if (L.d.bias_off) CK(launch_bias_grad_rows(L.dyT, ld, L.n_packed, L.d.bias_off, T->grads, st));
    }
  }
  // ---- dgrad: dX = conv(dY, Wd)
  if (nseg > 0) {
    EpiParams e{};
    e.cout = L.ctot;
    e.nseg = nseg;
    for (int i = 0; i < nseg; ++i) e.seg[i] = segs[i];
    CKR(t_gemm(h, "train.dgrad", {B, H, W, ks, false}, {{dY, L.kpad}}, L.wd, L.taps * L.kpad, L.ctot, pick_bn(L.ctot),
               EPI_F32, e, st));
  }
  return RAC_OK;
}

int vgg_backward(rac_handle* h, TrainState* T, VggRt& rt, const VggDef& d, const bf16* in, const float* dy,
                 int dy_cstride, int dy_coff, int up, const F32Seg* segs, int nseg, cudaStream_t st) {
  const TLayer& L = T->L[d.layer];
  CK(launch_bn_bwd(dy, dy_cstride, dy_coff, up, rt.raw, rt.mean, rt.rstd, T->params + L.d.gamma_off,
                   T->params + L.d.beta_off, T->cfg.batch, d.H, d.W, d.cout, T->bn_scratch, T->dy_a, nullptr,
                   T->grads + L.d.gamma_off, T->grads + L.d.beta_off, st));
  return conv_backward(h, T, d.layer, d.H, d.W, {{in, d.cin}}, T->dy_a, segs, nseg, st);
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

int lstm_backward_gn(rac_handle* h, TrainState* T, int s, int t, const bf16* xin, float* g_in, int cur, cudaStream_t st) {
  const int g = h->cfg.g_dim;
  Tape& tp = T->tape[t];
  for (int l = 1; l >= 0; --l) {
    GnCellArgs a = gn_args(h, T, s, l, t);
    a.dh = T->G_hs[s][l][cur]; a.dc = T->G_dc[s][l]; a.dy = T->gn_dy; a.d_ih = T->dy_a; a.d_hh = T->dy_b;
    a.part = T->gn_part;
    CK(launch_gn_cell_bwd(a, T->grads, st));
    const bf16* x = l == 0 ? xin : tp.hs[s][0];
    const bf16* hprev = t > 0 ? T->tape[t - 1].hs[s][l] : T->hzero;
    // input: layer 1 feeds layer 0's dh (+=), layer 0 feeds the stack input (=)
    F32Seg seg = {0, g, l == 1 ? T->G_hs[s][0][cur] : g_in, g, 0, l == 1 ? 1 : 0};
    CKR(conv_backward(h, T, l == 0 ? kLstm0[s] : kLstm1[s], 6, 8, {{x, g}}, T->dy_a, &seg, 1, st));
    // h_prev: the same layer at step t-1 (first writer of that buffer for this step); nothing before the first step
    F32Seg segh = {0, g, t > 0 ? T->G_hs[s][l][cur ^ 1] : nullptr, g, 0, 0};
    CKR(conv_backward(h, T, kLstmHH[s] + l, 6, 8, {{hprev, g, T->tape[0].hs[s][l], 1}}, T->dy_b, &segh, t > 0 ? 1 : 0, st));
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

// backward of one ConvLSTM stack at step t; the gradient w.r.t. the stack input lands in `g_in` (=)
int lstm_backward(rac_handle* h, TrainState* T, int s, int t, const bf16* xin, float* g_in, int cur, cudaStream_t st) {
  if (T->gn) return lstm_backward_gn(h, T, s, t, xin, g_in, cur, st);
  const int B = T->cfg.batch, g = h->cfg.g_dim, M = B * 48;
  Tape& tp = T->tape[t];
  for (int l = 1; l >= 0; --l) {
    const int layer = l == 0 ? kLstm0[s] : kLstm1[s];
    const float* cprev = t > 0 ? T->tape[t - 1].cs[s][l] : nullptr;
    CK(launch_lstm_bwd(T->G_hs[s][l][cur], T->G_dc[s][l], tp.gates[s][l], cprev, tp.cs[s][l], M, g, T->dy_a, st));
    const bf16* x = l == 0 ? xin : tp.hs[s][0];
    const bf16* hprev = t > 0 ? T->tape[t - 1].hs[s][l] : T->hzero;
    F32Seg segs[2];
    // input half: layer 1 feeds layer 0's dh (+=), layer 0 feeds the stack input (=)
    segs[0] = {0, g, l == 1 ? T->G_hs[s][0][cur] : g_in, g, 0, l == 1 ? 1 : 0};
    // h_prev half: the same layer at step t-1 (first writer of that buffer for this step)
    segs[1] = {g, 2 * g, t > 0 ? T->G_hs[s][l][cur ^ 1] : nullptr, g, 0, 0};
    CKR(conv_backward(h, T, layer, 6, 8, {{x, g}, {hprev, g, T->tape[0].hs[s][l], 1}}, T->dy_a, segs, 2, st));
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

int train_forward_step(rac_handle* h, TrainState* T, const rac_train_batch* bt, int t, cudaStream_t st) {
  const rac_config& c = h->cfg;
  const int B = T->cfg.batch, g = c.g_dim, z = c.z_dim;
  const size_t HW = 48 * 64;
  Tape& tp = T->tape[t];
  tp.sampled = (t > 0 && bt->true_token && !bt->true_token[t]) ? 1 : 0;
  const float* x_j = tp.sampled ? T->tape[t - 1].xp : bt->images + static_cast<size_t>(t) * B * 3 * HW;
  tp.xj = x_j;
  const float* m_j = bt->masks ? bt->masks + static_cast<size_t>(t) * B * HW : nullptr;
  const float* m_i = bt->masks ? bt->masks + static_cast<size_t>(t + 1) * B * HW : nullptr;
  const float* r_j = bt->states ? bt->states + static_cast<size_t>(t) * B * c.robot_dim : nullptr;
  const float* r_i = bt->states ? bt->states + static_cast<size_t>(t + 1) * B * c.robot_dim : nullptr;
  const float* a_j = bt->actions + static_cast<size_t>(t) * B * c.action_dim;
  CK(launch_img_prep_train(x_j, T->cfg.zero_robot ? m_j : nullptr, tp.img4, B, static_cast<int>(HW), st));
  // ---- encoder (run twice per step in the reference: identical values, BatchNorm running stats updated twice)
  CK(launch_first_conv(tp.img4, c.use_mask ? m_j : nullptr, (c.use_mask && c.use_future_mask) ? m_i : nullptr,
                       static_cast<long long>(HW), T->wfirst, T->zero64, nullptr, B, 48, 64, h->enc_cin, st, tp.vgg[0].raw));
  {
    const TLayer& L = T->L[RAC_L_ENC_C1_0];
    CK(launch_bn_stats(tp.vgg[0].raw, B * static_cast<int>(HW), 64, tp.vgg[0].mean, tp.vgg[0].rstd,
                       T->buffers + L.d.rmean_off, T->buffers + L.d.rvar_off, 2, st));
    CK(launch_bn_act(tp.vgg[0].raw, tp.vgg[0].mean, tp.vgg[0].rstd, T->params + L.d.gamma_off, T->params + L.d.beta_off,
                     B, 48, 64, 64, tp.a1, 64, 0, 0, st));
  }
  CKR(vgg_forward(h, T, tp.vgg[1], enc_def(h, 1), tp.a1, tp.cat5, 128, 64, 0, 2, st));
  CK(launch_maxpool2(tp.cat5, 128, 64, tp.p1, B, 48, 64, 64, st));
  CKR(vgg_forward(h, T, tp.vgg[2], enc_def(h, 2), tp.p1, tp.a2, 128, 0, 0, 2, st));
  CKR(vgg_forward(h, T, tp.vgg[3], enc_def(h, 3), tp.a2, tp.cat4, 256, 128, 0, 2, st));
  CK(launch_maxpool2(tp.cat4, 256, 128, tp.p2, B, 24, 32, 128, st));
  CKR(vgg_forward(h, T, tp.vgg[4], enc_def(h, 4), tp.p2, tp.a3a, 256, 0, 0, 2, st));
  CKR(vgg_forward(h, T, tp.vgg[5], enc_def(h, 5), tp.a3a, tp.a3b, 256, 0, 0, 2, st));
  CKR(vgg_forward(h, T, tp.vgg[6], enc_def(h, 6), tp.a3b, tp.cat3, 512, 256, 0, 2, st));
  CK(launch_maxpool2(tp.cat3, 512, 256, tp.p3, B, 12, 16, 256, st));
  CKR(vgg_forward(h, T, tp.vgg[7], enc_def(h, 7), tp.p3, tp.a4a, 512, 0, 0, 2, st));
  CKR(vgg_forward(h, T, tp.vgg[8], enc_def(h, 8), tp.a4a, tp.a4b, 512, 0, 0, 2, st));
  CKR(vgg_forward(h, T, tp.vgg[9], enc_def(h, 9), tp.a4b, tp.h4, g, 0, 0, 2, st));
  // ---- prior
  CK(launch_aux_tile(a_j, c.action_dim, c.action_dim, c.use_robot_state ? r_j : nullptr,
                     (c.use_robot_state && c.use_future_robot_state) ? r_i : nullptr, c.robot_dim, tp.aux, B, 48, st));
  auto input_conv = [&](int layer, std::vector<Src> srcs, bf16* out) -> int {
    const TLayer& L = T->L[layer];
    EpiParams e{};
    e.bias = L.bias; e.cout = g; e.out = out; e.out_cstride = g; e.out_coff = 0; e.upsample = 0; e.lrelu = 0;
    return t_gemm(h, "train.input_conv.fwd", {B, 6, 8, 3, false}, srcs, L.wp, 9 * L.ctot, L.n_packed,
                  tile_block_n(L.n_packed, EPI_ACT), EPI_ACT, e, st);
  };
  auto gauss = [&](int layer, const bf16* hin, const float* eps, float* mu, float* lv, bf16* zout) -> int {
    const TLayer& L = T->L[layer];
    EpiParams e{};
    e.bias = L.bias; e.cout = 128; e.eps = eps; e.mu_out = mu; e.logvar_out = lv; e.z_out = zout; e.z_dim = z;
    return t_gemm(h, "train.gauss.fwd", {B, 6, 8, 3, false}, {{hin, g}}, L.wp, 9 * g, 128, 128, EPI_GAUSS, e, st);
  };
  CKR(input_conv(RAC_L_PRIOR_IN, {{tp.aux, 64}, {tp.h4, g}}, tp.pin));
  CKR(lstm_forward(h, T, 0, t, tp.pin, st));
  CKR(gauss(RAC_L_PRIOR_GAUSS, tp.hs[0][1], tp.eps_p, tp.mu_p, tp.lv_p, tp.zprior));
  // ---- posterior (h_target == h4, dynamics.py:619)
  if (c.use_robot_state) {
    CK(launch_aux_tile(nullptr, 0, 0, r_i, nullptr, c.robot_dim, tp.auxp, B, 48, st));
    CKR(input_conv(RAC_L_POST_IN, {{tp.auxp, 64}, {tp.h4, g}}, tp.postin));
  } else {
    CKR(input_conv(RAC_L_POST_IN, {{tp.h4, g}}, tp.postin));
  }
  CKR(lstm_forward(h, T, 1, t, tp.postin, st));
  CKR(gauss(RAC_L_POST_GAUSS, tp.hs[1][1], tp.eps_q, tp.mu, tp.lv, tp.z));
  // ---- frame predictor + decoder
  CKR(input_conv(RAC_L_FP_IN, {{tp.aux, 64}, {tp.h4, g}, {tp.z, 64}}, tp.fin));
  CKR(lstm_forward(h, T, 2, t, tp.fin, st));
  CKR(vgg_forward(h, T, tp.vgg[10], dec_def(h, 0), tp.hs[2][1], tp.d2a, 512, 0, 0, 1, st));
  CKR(vgg_forward(h, T, tp.vgg[11], dec_def(h, 1), tp.d2a, tp.d2b, 512, 0, 0, 1, st));
  if (tp.dcat5 != tp.cat5) {
    // fixed_skip, t > 0: the skip halves are those of the first frame
    const Tape& t0 = T->tape[0];
    CK(cudaMemcpy2DAsync(tp.dcat3 + 256, 512 * sizeof(bf16), t0.cat3 + 256, 512 * sizeof(bf16), 256 * sizeof(bf16),
                         static_cast<size_t>(B) * 192, cudaMemcpyDeviceToDevice, st));
    CK(cudaMemcpy2DAsync(tp.dcat4 + 128, 256 * sizeof(bf16), t0.cat4 + 128, 256 * sizeof(bf16), 128 * sizeof(bf16),
                         static_cast<size_t>(B) * 768, cudaMemcpyDeviceToDevice, st));
    CK(cudaMemcpy2DAsync(tp.dcat5 + 64, 128 * sizeof(bf16), t0.cat5 + 64, 128 * sizeof(bf16), 64 * sizeof(bf16),
                         static_cast<size_t>(B) * 3072, cudaMemcpyDeviceToDevice, st));
  }
  CKR(vgg_forward(h, T, tp.vgg[12], dec_def(h, 2), tp.d2b, tp.dcat3, 512, 0, 1, 1, st));
  CKR(vgg_forward(h, T, tp.vgg[13], dec_def(h, 3), tp.dcat3, tp.d3a, 256, 0, 0, 1, st));
  CKR(vgg_forward(h, T, tp.vgg[14], dec_def(h, 4), tp.d3a, tp.d3b, 256, 0, 0, 1, st));
  CKR(vgg_forward(h, T, tp.vgg[15], dec_def(h, 5), tp.d3b, tp.dcat4, 256, 0, 1, 1, st));
  CKR(vgg_forward(h, T, tp.vgg[16], dec_def(h, 6), tp.dcat4, tp.d4a, 128, 0, 0, 1, st));
  CKR(vgg_forward(h, T, tp.vgg[17], dec_def(h, 7), tp.d4a, tp.dcat5, 128, 0, 1, 1, st));
  CKR(vgg_forward(h, T, tp.vgg[18], dec_def(h, 8), tp.dcat5, tp.d5, 64, 0, 0, 1, st));
  {
    const TLayer& L = T->L[RAC_L_DEC_UPC5_1];
    EpiParams e{};
    e.bias = L.bias; e.cout = 4; e.xpred_out = tp.x4;
    CKR(t_gemm(h, "train.frame.fwd", {B, 48, 64, 3, false}, {{tp.d5, 64}}, L.wp, 9 * 64, 16, 16, EPI_FRAME, e, st));
  }
  CK(launch_composite(tp.x4, x_j, tp.xp, B, static_cast<int>(HW), st));
  if (m_i)  // logging metrics of the reference step (trainer.py:436-439)
    CK(launch_robot_world_mse(tp.xp, bt->images + static_cast<size_t>(t + 1) * B * 3 * HW, m_i, bt->losses + 2, B,
                              static_cast<int>(HW), st));
  // KL(posterior || prior) value (trainer.py:454-458)
  CK(launch_kl_loss(tp.mu, tp.lv, tp.mu_p, tp.lv_p, T->kl_tmp, static_cast<int64_t>(B) * z * 48, B, st));
  CK(launch_sum_f32(T->kl_tmp, 1, bt->losses + 1, st));
  return RAC_OK;
}

int train_backward_step(rac_handle* h, TrainState* T, const rac_train_batch* bt, int t, int cur, cudaStream_t st) {
  const rac_config& c = h->cfg;
  const int B = T->cfg.batch, g = c.g_dim, z = c.z_dim;
  const size_t HW = 48 * 64;
  Tape& tp = T->tape[t];
  const float* x_j = tp.xj;
  const float* x_i = bt->images + static_cast<size_t>(t + 1) * B * 3 * HW;
  const int S = T->cfg.steps;
  const float* gp_in = (t + 1 < S && T->tape[t + 1].sampled) ? T->G_img[cur] : nullptr;
  float* gxj_out = tp.sampled ? T->G_img[cur ^ 1] : nullptr;
  const float* m_j = bt->masks ? bt->masks + static_cast<size_t>(t) * B * HW : nullptr;
  const float* m_i = bt->masks ? bt->masks + static_cast<size_t>(t + 1) * B * HW : nullptr;
  // ---- reconstruction loss and its gradient w.r.t. the decoder logits (trainer.py:406-433)
  CK(launch_frame_loss(tp.x4, x_j, x_i, m_i, T->cfg.recon_kind, T->cfg.robot_pixel_weight, B, static_cast<int>(HW),
                       T->loss_part, T->dy_b, gp_in, gxj_out, st, bt->batch_weight));
  CK(launch_sum_f32(T->loss_part, B, bt->losses + 0, st));
  F32Seg seg[3];
  // ---- decoder
  seg[0] = {0, 64, T->G_d5, 64, 0, 0};
  CKR(conv_backward(h, T, RAC_L_DEC_UPC5_1, 48, 64, {{tp.d5, 64}}, T->dy_b, seg, 1, st));
  // gradient of a concat buffer: [decoder half | skip half]. fixed_skip: the skip half is summed over the steps
  // (first writer = the first processed step, t == S - 1) instead of going to this step's encoder
  const bool fixed = T->cfg.fixed_skip != 0;
  const int skip_acc = (t == S - 1) ? 0 : 1;
  auto cat_segs = [&](float* G_cat, float* G_skip, int half) -> int {
    if (!fixed) { seg[0] = {0, 2 * half, G_cat, 2 * half, 0, 0}; return 1; }
    seg[0] = {0, half, G_cat, 2 * half, 0, 0};
    seg[1] = {half, 2 * half, G_skip, half, 0, skip_acc};
    return 2;
  };
  int ncat = cat_segs(T->G_cat5, T->G_skip5, 64);
  CKR(vgg_backward(h, T, tp.vgg[18], dec_def(h, 8), tp.dcat5, T->G_d5, 64, 0, 0, seg, ncat, st));
  seg[0] = {0, 128, T->G_d4a, 128, 0, 0};
  CKR(vgg_backward(h, T, tp.vgg[17], dec_def(h, 7), tp.d4a, T->G_cat5, 128, 0, 1, seg, 1, st));
  ncat = cat_segs(T->G_cat4, T->G_skip4, 128);
  CKR(vgg_backward(h, T, tp.vgg[16], dec_def(h, 6), tp.dcat4, T->G_d4a, 128, 0, 0, seg, ncat, st));
  seg[0] = {0, 256, T->G_d3b, 256, 0, 0};
  CKR(vgg_backward(h, T, tp.vgg[15], dec_def(h, 5), tp.d3b, T->G_cat4, 256, 0, 1, seg, 1, st));
  seg[0] = {0, 256, T->G_d3a, 256, 0, 0};
  CKR(vgg_backward(h, T, tp.vgg[14], dec_def(h, 4), tp.d3a, T->G_d3b, 256, 0, 0, seg, 1, st));
  ncat = cat_segs(T->G_cat3, T->G_skip3, 256);
  CKR(vgg_backward(h, T, tp.vgg[13], dec_def(h, 3), tp.dcat3, T->G_d3a, 256, 0, 0, seg, ncat, st));
  seg[0] = {0, 512, T->G_d2b, 512, 0, 0};
  CKR(vgg_backward(h, T, tp.vgg[12], dec_def(h, 2), tp.d2b, T->G_cat3, 512, 0, 1, seg, 1, st));
  seg[0] = {0, 512, T->G_d2a, 512, 0, 0};
  CKR(vgg_backward(h, T, tp.vgg[11], dec_def(h, 1), tp.d2a, T->G_d2b, 512, 0, 0, seg, 1, st));
  seg[0] = {0, g, T->G_hs[2][1][cur], g, 0, 1};
  CKR(vgg_backward(h, T, tp.vgg[10], dec_def(h, 0), tp.hs[2][1], T->G_d2a, 512, 0, 0, seg, 1, st));
  // ---- frame predictor
  CKR(lstm_backward(h, T, 2, t, tp.fin, T->G_fin, cur, st));
  CK(launch_cast_bf16(T->G_fin, static_cast<long long>(B) * 48 * g, T->dy_b, st));
  seg[0] = {0, 64, nullptr, 0, 0, 0};
  seg[1] = {64, 64 + g, T->G_h4, g, 0, 0};
  seg[2] = {64 + g, 128 + g, T->G_z, 64, 0, 0};
  CKR(conv_backward(h, T, RAC_L_FP_IN, 6, 8, {{tp.aux, 64}, {tp.h4, g}, {tp.z, 64}}, T->dy_b, seg, 3, st));
  // ---- z sample + KL -> posterior and prior heads
  CK(launch_gauss_bwd(T->G_z, tp.mu, tp.lv, tp.eps_q, tp.mu_p, tp.lv_p, B, z, 48, T->cfg.kl_beta, B, T->dy_a, T->dy_b, st));
  // (conv_backward uses dy_a only through its dY argument; the two heads are processed one after the other)
  {
    // posterior head first: its operand lives in dy_a, which lstm_backward will overwrite later
    seg[0] = {0, g, T->G_hs[1][1][cur], g, 0, 1};
    CKR(conv_backward(h, T, RAC_L_POST_GAUSS, 6, 8, {{tp.hs[1][1], g}}, T->dy_a, seg, 1, st));
    seg[0] = {0, g, T->G_hs[0][1][cur], g, 0, 1};
    CKR(conv_backward(h, T, RAC_L_PRIOR_GAUSS, 6, 8, {{tp.hs[0][1], g}}, T->dy_b, seg, 1, st));
  }
  // ---- posterior stack
  CKR(lstm_backward(h, T, 1, t, tp.postin, T->G_postin, cur, st));
  CK(launch_cast_bf16(T->G_postin, static_cast<long long>(B) * 48 * g, T->dy_b, st));
  if (c.use_robot_state) {
    seg[0] = {0, 64, nullptr, 0, 0, 0};
    seg[1] = {64, 64 + g, T->G_h4, g, 0, 1};
    CKR(conv_backward(h, T, RAC_L_POST_IN, 6, 8, {{tp.auxp, 64}, {tp.h4, g}}, T->dy_b, seg, 2, st));
  } else {
    seg[0] = {0, g, T->G_h4, g, 0, 1};
    CKR(conv_backward(h, T, RAC_L_POST_IN, 6, 8, {{tp.h4, g}}, T->dy_b, seg, 1, st));
  }
  // ---- prior stack
  CKR(lstm_backward(h, T, 0, t, tp.pin, T->G_pin, cur, st));
  CK(launch_cast_bf16(T->G_pin, static_cast<long long>(B) * 48 * g, T->dy_b, st));
  seg[0] = {0, 64, nullptr, 0, 0, 0};
  seg[1] = {64, 64 + g, T->G_h4, g, 0, 1};
  CKR(conv_backward(h, T, RAC_L_PRIOR_IN, 6, 8, {{tp.aux, 64}, {tp.h4, g}}, T->dy_b, seg, 2, st));
  // ---- encoder (the gradients of the two reference passes are summed in G_h4 / the skip halves)
  seg[0] = {0, 512, T->G_a4b, 512, 0, 0};
  CKR(vgg_backward(h, T, tp.vgg[9], enc_def(h, 9), tp.a4b, T->G_h4, g, 0, 0, seg, 1, st));
  seg[0] = {0, 512, T->G_a4a, 512, 0, 0};
  CKR(vgg_backward(h, T, tp.vgg[8], enc_def(h, 8), tp.a4a, T->G_a4b, 512, 0, 0, seg, 1, st));
  seg[0] = {0, 256, T->G_p3, 256, 0, 0};
  CKR(vgg_backward(h, T, tp.vgg[7], enc_def(h, 7), tp.p3, T->G_a4a, 512, 0, 0, seg, 1, st));
  // gradient of the encoder outputs that double as skips: pooling path + decoder skip path. fixed_skip: steps t > 0
  // get the pooling path only; step 0 adds it to the all-steps skip sum
  struct SkipGrad { float* p; int cstride, coff, acc; };
  auto skip_grad = [&](float* G_cat, float* G_skip, int half) -> SkipGrad {
    if (!fixed) return {G_cat, 2 * half, half, 1};
    if (t > 0) return {G_cat, 2 * half, half, 0};
    return {G_skip, half, 0, 1};
  };
  SkipGrad sk = skip_grad(T->G_cat3, T->G_skip3, 256);
  CK(launch_pool_bwd(tp.cat3, 512, 256, T->G_p3, B, 12, 16, 256, sk.p, sk.cstride, sk.coff, sk.acc, st));
  seg[0] = {0, 256, T->G_a3b, 256, 0, 0};
  CKR(vgg_backward(h, T, tp.vgg[6], enc_def(h, 6), tp.a3b, sk.p, sk.cstride, sk.coff, 0, seg, 1, st));
  seg[0] = {0, 256, T->G_a3a, 256, 0, 0};
  CKR(vgg_backward(h, T, tp.vgg[5], enc_def(h, 5), tp.a3a, T->G_a3b, 256, 0, 0, seg, 1, st));
  seg[0] = {0, 128, T->G_p2, 128, 0, 0};
  CKR(vgg_backward(h, T, tp.vgg[4], enc_def(h, 4), tp.p2, T->G_a3a, 256, 0, 0, seg, 1, st));
  sk = skip_grad(T->G_cat4, T->G_skip4, 128);
  CK(launch_pool_bwd(tp.cat4, 256, 128, T->G_p2, B, 24, 32, 128, sk.p, sk.cstride, sk.coff, sk.acc, st));
  seg[0] = {0, 128, T->G_a2, 128, 0, 0};
  CKR(vgg_backward(h, T, tp.vgg[3], enc_def(h, 3), tp.a2, sk.p, sk.cstride, sk.coff, 0, seg, 1, st));
  seg[0] = {0, 64, T->G_p1, 64, 0, 0};
  CKR(vgg_backward(h, T, tp.vgg[2], enc_def(h, 2), tp.p1, T->G_a2, 128, 0, 0, seg, 1, st));
  sk = skip_grad(T->G_cat5, T->G_skip5, 64);
  CK(launch_pool_bwd(tp.cat5, 128, 64, T->G_p1, B, 48, 64, 64, sk.p, sk.cstride, sk.coff, sk.acc, st));
  seg[0] = {0, 64, T->G_a1, 64, 0, 0};
  CKR(vgg_backward(h, T, tp.vgg[1], enc_def(h, 1), tp.a1, sk.p, sk.cstride, sk.coff, 0, seg, 1, st));
  {
    // encoder.c1.0: BatchNorm backward, then the small-K weight gradient on CUDA cores (no input gradient needed)
    const TLayer& L = T->L[RAC_L_ENC_C1_0];
    CK(launch_bn_bwd(T->G_a1, 64, 0, 0, tp.vgg[0].raw, tp.vgg[0].mean, tp.vgg[0].rstd, T->params + L.d.gamma_off,
                     T->params + L.d.beta_off, B, 48, 64, 64, T->bn_scratch, T->dy_a, T->draw32,
                     T->grads + L.d.gamma_off, T->grads + L.d.beta_off, st));
    CK(launch_first_wgrad(tp.img4, c.use_mask ? m_j : nullptr, (c.use_mask && c.use_future_mask) ? m_i : nullptr,
                          static_cast<long long>(HW), T->draw32, T->grads + L.d.w_off, B, 48, 64, h->enc_cin, T->fw_part,
                          kFwBlocks, st));
    // a model-sampled input frame also receives gradient through the encoder (zeroed robot pixels get none)
    if (tp.sampled && T->dbg_keep)
      CK(cudaMemcpyAsync(T->dbg_draw32, T->draw32, sizeof(float) * static_cast<size_t>(B) * HW * 64, cudaMemcpyDeviceToDevice, st));
    if (tp.sampled)
      CK(launch_first_dgrad(T->draw32, T->wfirst, h->enc_cin, T->cfg.zero_robot ? m_j : nullptr, gxj_out, B, 48, 64, st));
  }
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
  if (const char* wg = getenv("RAC_WGRAD_IM2COL")) T->wgrad_implicit = atoi(wg) ? 0 : 1;
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
      {
        // dyT: the output gradients of all time steps -- [S][M][kpad] for the implicit-GEMM weight gradient, or the
        // transposed [kpad][S * mpad] (+ the im2col operand xcolT) of the round-1 path
        const size_t ld = static_cast<size_t>(S) * round_up(static_cast<int>(rows_of(i)), 64);
        if (!T->wgrad_implicit) L.xcolT = bp.take<bf16>(static_cast<size_t>(L.taps) * L.ctot * ld);
        L.dyT = bp.take<bf16>(static_cast<size_t>(L.kpad) * ld);
      }
      L.wd = bp.take<bf16>(static_cast<size_t>(L.ctot) * L.taps * L.kpad);
      L.dwp = bp.take<float>(static_cast<size_t>(L.kpad) * L.taps * L.ctot);
      L.bias = bp.take<float>(L.n_packed);
    }
    T->wfirst = bp.take<float>(45 * 64);
    T->zero64 = bp.take<float>(64);
    T->tape.resize(S);
    for (int t = 0; t < S; ++t) {
      Tape& tp = T->tape[t];
      tp.img4 = bp.take<float>(M0 * 4);
      tp.a1 = bp.take<bf16>(M0 * 64); tp.cat5 = bp.take<bf16>(M0 * 128); tp.p1 = bp.take<bf16>(M1 * 64);
      tp.a2 = bp.take<bf16>(M1 * 128); tp.cat4 = bp.take<bf16>(M1 * 256); tp.p2 = bp.take<bf16>(M2 * 128);
      tp.a3a = bp.take<bf16>(M2 * 256); tp.a3b = bp.take<bf16>(M2 * 256); tp.cat3 = bp.take<bf16>(M2 * 512);
      tp.p3 = bp.take<bf16>(M3 * 256); tp.a4a = bp.take<bf16>(M3 * 512); tp.a4b = bp.take<bf16>(M3 * 512);
      tp.h4 = bp.take<bf16>(M3 * g); tp.aux = bp.take<bf16>(M3 * 64); tp.auxp = bp.take<bf16>(M3 * 64);
      tp.pin = bp.take<bf16>(M3 * g); tp.postin = bp.take<bf16>(M3 * g); tp.fin = bp.take<bf16>(M3 * g);
      tp.z = bp.take<bf16>(M3 * 64); tp.zprior = bp.take<bf16>(M3 * 64);
      for (int s = 0; s < 3; ++s)
        for (int l = 0; l < 2; ++l) {
          tp.hs[s][l] = bp.take<bf16>(M3 * g);
          tp.cs[s][l] = bp.take<float>(M3 * g);
          tp.gates[s][l] = bp.take<float>(M3 * 4 * g);
          if (T->gn) {
            tp.raw_ih[s][l] = bp.take<float>(M3 * 4 * g); tp.raw_hh[s][l] = bp.take<float>(M3 * 4 * g);
            tp.c_raw[s][l] = bp.take<float>(M3 * g); tp.gn_stats[s][l] = bp.take<float>(static_cast<size_t>(B) * 96);
          }
        }
      tp.d2a = bp.take<bf16>(M3 * 512); tp.d2b = bp.take<bf16>(M3 * 512); tp.d3a = bp.take<bf16>(M2 * 256);
      tp.d3b = bp.take<bf16>(M2 * 256); tp.d4a = bp.take<bf16>(M1 * 128); tp.d5 = bp.take<bf16>(M0 * 64);
      tp.dcat5 = tp.cat5; tp.dcat4 = tp.cat4; tp.dcat3 = tp.cat3;
      // fixed_skip: every step decodes from its own concat buffers (skip halves = copies of step 0's encoder outputs).
      // Step 0 could use cat* directly, but the all-time-steps weight gradient reads the decoder inputs through ONE
      // tensor map with a constant step stride, so step 0 gets its own buffers too
      if (cfg->fixed_skip && (t > 0 || T->wgrad_implicit)) {
        tp.dcat5 = bp.take<bf16>(M0 * 128); tp.dcat4 = bp.take<bf16>(M1 * 256); tp.dcat3 = bp.take<bf16>(M2 * 512);
      }
      for (int i = 0; i < 19; ++i) {
        const VggDef d = i < 10 ? enc_def(h, i) : dec_def(h, i - 10);
        tp.vgg[i].raw = bp.take<float>(static_cast<size_t>(B) * d.H * d.W * d.cout);
        tp.vgg[i].mean = bp.take<float>(d.cout);
        tp.vgg[i].rstd = bp.take<float>(d.cout);
      }
      const size_t zn = static_cast<size_t>(B) * z * 48;
      tp.mu_p = bp.take<float>(zn); tp.lv_p = bp.take<float>(zn); tp.mu = bp.take<float>(zn); tp.lv = bp.take<float>(zn);
      tp.eps_p = bp.take<float>(zn); tp.eps_q = bp.take<float>(zn);
      tp.x4 = bp.take<float>(M0 * 4);
      tp.xp = bp.take<float>(M0 * 3);
    }
    T->G_d5 = bp.take<float>(M0 * 64); T->G_cat5 = bp.take<float>(M0 * 128); T->G_d4a = bp.take<float>(M1 * 128);
    T->G_cat4 = bp.take<float>(M1 * 256); T->G_d3b = bp.take<float>(M2 * 256); T->G_d3a = bp.take<float>(M2 * 256);
    T->G_cat3 = bp.take<float>(M2 * 512); T->G_d2b = bp.take<float>(M3 * 512); T->G_d2a = bp.take<float>(M3 * 512);
    T->G_fin = bp.take<float>(M3 * g); T->G_pin = bp.take<float>(M3 * g); T->G_postin = bp.take<float>(M3 * g);
    T->G_z = bp.take<float>(M3 * 64); T->G_h4 = bp.take<float>(M3 * g);
    T->G_a4b = bp.take<float>(M3 * 512); T->G_a4a = bp.take<float>(M3 * 512); T->G_p3 = bp.take<float>(M3 * 256);
    T->G_a3b = bp.take<float>(M2 * 256); T->G_a3a = bp.take<float>(M2 * 256); T->G_p2 = bp.take<float>(M2 * 128);
    T->G_a2 = bp.take<float>(M1 * 128); T->G_p1 = bp.take<float>(M1 * 64); T->G_a1 = bp.take<float>(M0 * 64);
    T->G_skip5 = bp.take<float>(M0 * 64); T->G_skip4 = bp.take<float>(M1 * 128); T->G_skip3 = bp.take<float>(M2 * 256);
    for (int s = 0; s < 3; ++s)
      for (int l = 0; l < 2; ++l) {
        T->G_hs[s][l][0] = bp.take<float>(M3 * g);
        T->G_hs[s][l][1] = bp.take<float>(M3 * g);
        T->G_dc[s][l] = bp.take<float>(M3 * g);
      }
    const size_t dy_elems = std::max(M0 * 64, M3 * static_cast<size_t>(4 * g));
    T->dy_a = bp.take<bf16>(dy_elems); T->dy_b = bp.take<bf16>(dy_elems);
    T->bn_scratch = bp.take<float>(2 * 2048);
    T->draw32 = bp.take<float>(M0 * 64);
    T->fw_part = bp.take<float>(static_cast<size_t>(kFwBlocks) * 45 * 64);
    T->hzero = bp.take<bf16>(M3 * g); T->czero = bp.take<float>(M3 * g);
    T->G_img[0] = bp.take<float>(M0 * 3); T->G_img[1] = bp.take<float>(M0 * 3);
    T->dbg_draw32 = bp.take<float>(M0 * 64);
    T->loss_part = bp.take<float>(B); T->kl_tmp = bp.take<float>(4);
    if (T->gn) {
      T->gn_dy = bp.take<float>(M3 * 4 * g);
      T->gn_part = bp.take<float>(static_cast<size_t>(B) * 14 * g);
    }
    if (T->wgrad_implicit) {
      // split-K scratch: the largest splits x |dWp| over the layers that need more than one slice
      size_t need = 4;
      for (int i = 1; i < T->nlayers; ++i) {
        const TLayer& L = T->L[i];
        const size_t rows = rows_of(i);
        const int W = rows == M0 ? 64 : rows == M1 ? 32 : rows == M2 ? 16 : 8;
        WgradGeom wg = wg_plan(L, B, W * 3 / 4, W, S);
        int one_src[1] = {L.ctot};
        wg_ctiles(wg, one_src, 1);  // (an upper bound on the tile count is enough: more tiles = fewer splits)
        wg.num_ctiles = std::max(1, L.ctot / 256);
        wg_split(wg, h->num_sms);
        if (wg.splits > 1) need = std::max(need, static_cast<size_t>(wg.splits) * static_cast<size_t>(wg.out_split_stride));
      }
      T->wg_part_elems = need;
      T->wg_part = bp.take<float>(need);
    }
    if (!pass) {
      CK(cudaMalloc(&T->arena, bp.off + 1024));
      CK(cudaMemset(T->arena, 0, bp.off + 1024));
    }
  }
  // every step of the tape allocates the same sequence of buffers: one constant stride between time steps
  T->step_bytes = S > 1 ? static_cast<unsigned long long>(reinterpret_cast<const char*>(T->tape[1].img4) -
                                                          reinterpret_cast<const char*>(T->tape[0].img4))
                        : (1ull << 20);
  for (int t = 1; t < S; ++t) {
    const Tape &a = T->tape[t - 1], &b = T->tape[t];
    const char* pa[] = {reinterpret_cast<const char*>(a.img4), reinterpret_cast<const char*>(a.hs[2][1]),
                        reinterpret_cast<const char*>(a.d5), reinterpret_cast<const char*>(a.dcat3), reinterpret_cast<const char*>(a.xp)};
    const char* pb[] = {reinterpret_cast<const char*>(b.img4), reinterpret_cast<const char*>(b.hs[2][1]),
                        reinterpret_cast<const char*>(b.d5), reinterpret_cast<const char*>(b.dcat3), reinterpret_cast<const char*>(b.xp)};
    for (int k = 0; k < 5; ++k)
      if (static_cast<unsigned long long>(pb[k] - pa[k]) != T->step_bytes && T->wgrad_implicit)
        return fail(h, RAC_ERR_STATE, "training tape: time steps are not equally spaced");
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
  // ---- parameters -> packed bf16 operands (forward + dgrad), packed biases; zero the gradient accumulators
  for (int i = 1; i < T->nlayers; ++i) {
    TLayer& L = T->L[i];
    CK(launch_pack_weights(T->params, L.d.row_off, L.d.col_off, L.n_packed, L.taps, L.ctot, L.d.flip, L.wp, st));
    CK(launch_transpose_flip(L.wp, L.n_packed, L.taps, L.ctot, L.kpad, L.wd, st));
    if (L.d.bias_off) CK(launch_gather_f32(T->params, L.d.bias_off, L.n_packed, L.bias, st));
    // (the implicit-GEMM weight gradient writes every element of dWp; only the round-1 GEMM accumulates into it)
    if (!T->wgrad_implicit) CK(cudaMemsetAsync(L.dwp, 0, sizeof(float) * L.kpad * L.taps * L.ctot, st));
  }
  CK(launch_pack_first(T->params + T->L[RAC_L_ENC_C1_0].d.w_off, h->enc_cin, T->wfirst, st));
  CK(cudaMemsetAsync(T->grads, 0, sizeof(float) * T->cfg.n_params, st));
  CK(cudaMemsetAsync(bt->losses, 0, sizeof(float) * 4, st));
  const size_t zn = static_cast<size_t>(B) * z * 48;
  // Philox counter of the reparameterisation noise = the caller's global training step (persisted in checkpoints),
  // NOT a count kept in this state: rac_train_create runs again after a resume or a new batch shape and must not
  // replay the noise of steps 0..k
  const unsigned int noise_step = static_cast<unsigned int>(bt->noise_step);
  for (int t = 0; t < S; ++t) {
    Tape& tp = T->tape[t];
    if (bt->eps_prior) CK(cudaMemcpyAsync(tp.eps_p, bt->eps_prior + t * zn, sizeof(float) * zn, cudaMemcpyDeviceToDevice, st));
    else CK(launch_normal_fill(tp.eps_p, static_cast<long long>(zn), bt->seed, 2u * (noise_step * S + t), st));
    if (bt->eps_post) CK(cudaMemcpyAsync(tp.eps_q, bt->eps_post + t * zn, sizeof(float) * zn, cudaMemcpyDeviceToDevice, st));
    else CK(launch_normal_fill(tp.eps_q, static_cast<long long>(zn), bt->seed, 2u * (noise_step * S + t) + 1u, st));
    CKR(train_forward_step(h, T, bt, t, st));
  }
  // ---- BPTT
  for (int s = 0; s < 3; ++s)
    for (int l = 0; l < 2; ++l) {
      CK(cudaMemsetAsync(T->G_hs[s][l][0], 0, sizeof(float) * M3 * g, st));
      CK(cudaMemsetAsync(T->G_hs[s][l][1], 0, sizeof(float) * M3 * g, st));
      CK(cudaMemsetAsync(T->G_dc[s][l], 0, sizeof(float) * M3 * g, st));
    }
  int cur = 0;
  for (int t = S - 1; t >= 0; --t) {
    T->cur_t = t;
    CKR(train_backward_step(h, T, bt, t, cur, st));
    cur ^= 1;
  }
  // ---- packed weight gradients -> flat parameter layout
  for (int i = 1; i < T->nlayers; ++i) {
    TLayer& L = T->L[i];
    CK(launch_unpack_grads(L.dwp, L.d.row_off, L.d.col_off, L.n_packed, L.taps, L.ctot, L.d.flip, T->grads, st));
  }
  return RAC_OK;
}

int rac_train_debug_buffer(rac_handle* h, const char* name, int step, void** ptr) {
  if (!h || !h->train || !name || !ptr) return RAC_ERR_INVALID;
  TrainState* T = static_cast<TrainState*>(h->train);
  if (step < 0 || step >= static_cast<int>(T->tape.size())) return fail(h, RAC_ERR_INVALID, "bad step %d", step);
  Tape& tp = T->tape[step];
  struct { const char* n; void* p; } tab[] = {
      {"G_d5", T->G_d5}, {"G_cat5", T->G_cat5}, {"G_d4a", T->G_d4a}, {"G_cat4", T->G_cat4}, {"G_d3b", T->G_d3b},
      {"G_d3a", T->G_d3a}, {"G_cat3", T->G_cat3}, {"G_d2b", T->G_d2b}, {"G_d2a", T->G_d2a}, {"G_fin", T->G_fin},
      {"G_pin", T->G_pin}, {"G_postin", T->G_postin}, {"G_z", T->G_z}, {"G_h4", T->G_h4}, {"G_a4b", T->G_a4b},
      {"G_a4a", T->G_a4a}, {"G_p3", T->G_p3}, {"G_a3b", T->G_a3b}, {"G_a3a", T->G_a3a}, {"G_p2", T->G_p2},
      {"G_a2", T->G_a2}, {"G_p1", T->G_p1}, {"G_a1", T->G_a1},
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
  T->adam_t += 1;
  CK(launch_adam(T->params, T->grads, T->m, T->v, T->cfg.n_params, T->cfg.lr, T->cfg.beta1, T->cfg.beta2,
                 T->cfg.adam_eps, T->adam_t, static_cast<cudaStream_t>(stream)));
  return RAC_OK;
}

}  // extern "C"
