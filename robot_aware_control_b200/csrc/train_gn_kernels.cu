// NormConvLSTMCell in the training step (reference src/prediction/models/lstm.py:151-198, cfg.lstm_group_norm):
//   gates = GroupNorm16(ih_conv(x)) + GroupNorm16(hh_conv(h_prev));  i, f, o = sigmoid, g = tanh
//   c = GroupNorm16(f * c_prev + i * g);  h = o * tanh(c)
// The two convolutions (forward, dgrad, wgrad) are conv_tc_kernel GEMMs like every other layer; this file holds the
// pointwise + GroupNorm part, forward (train mode: everything the backward pass needs is saved) and backward.
//
// Layout: raw convolution outputs and gate tensors are [M = B * P, 4 * hid] fp32 with packed columns (channel, gate);
// cells are [M, hid] fp32. GroupNorm(16, 4 * hid) groups of the reference = (gate, quarter of the channels);
// GroupNorm(16, hid) groups = sixteenths of the channels. Statistics are per sample, so one CTA owns (sample, quarter):
// its 4 x 2 gate groups and 4 cell groups reduce inside the CTA, in a fixed order (no atomics: deterministic). Thread t
// owns channel t % (hid / 4) of the quarter and walks the positions p = t / (hid / 4), + 256 / (hid / 4), ...; every
// element is re-read only by the thread that wrote it. The affine-parameter gradients are sums over samples and
// positions: each CTA writes per-sample partial sums, gn_affine_fold_kernel adds them over the samples in order.
// The affine vectors are read from the flat parameter buffer in the reference's own order (gate * hid + channel).
#include "train_kernels.cuh"
#include "epilogue.cuh"
#include "ptx.cuh"

namespace rac {

namespace {

constexpr int kGnThreads = 256;
constexpr float kGnEps = 1e-5f;

__device__ __forceinline__ float sigmoid_libm(float x) { return 1.0f / (1.0f + expf(-x)); }

// Sum of every v[n] over the CTA; all threads get all totals. sh: [kGnThreads / 32][N] floats.
template <int N>
__device__ __forceinline__ void block_sum_vec(float (&v)[N], float* sh) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int n = 0; n < N; ++n)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v[n] += __shfl_xor_sync(0xffffffffu, v[n], o);
  __syncthreads();  // earlier readers of sh are done
  if (lane == 0)
#pragma unroll
    for (int n = 0; n < N; ++n) sh[warp * N + n] = v[n];
  __syncthreads();
#pragma unroll
  for (int n = 0; n < N; ++n) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < kGnThreads / 32; ++w) t += sh[w * N + n];
    v[n] = t;
  }
}

__device__ __forceinline__ void mean_rstd(float sum, float sumsq, float count, float& mean, float& rstd) {
  const double m = static_cast<double>(sum) / count;
  double var = static_cast<double>(sumsq) / count - m * m;
  if (var < 0.0) var = 0.0;
  mean = static_cast<float>(m);
  rstd = static_cast<float>(1.0 / sqrt(var + static_cast<double>(kGnEps)));
}

__device__ __forceinline__ float comp(const float4& v, int k) { return k == 0 ? v.x : (k == 1 ? v.y : (k == 2 ? v.z : v.w)); }

// stats per sample: 48 (mean, rstd) pairs: [tensor (ih, hh)][gate][quarter] = 32, then [quarter][sixteenth in quarter] = 16
__device__ __forceinline__ int stat_gate(int tensor, int gate, int quarter) { return (tensor * 4 + gate) * 4 + quarter; }
__device__ __forceinline__ int stat_cell(int quarter, int k) { return 32 + quarter * 4 + k; }

__global__ void __launch_bounds__(kGnThreads)
gn_cell_fwd_kernel(const GnCellArgs a) {
  pdl_entry();
  __shared__ float sh[(kGnThreads / 32) * 16];
  const int hid = a.hid, P = a.P, QW = hid / 4, R = kGnThreads / QW;
  const int b = blockIdx.x, quarter = blockIdx.y;
  const int chl = threadIdx.x % QW, r0 = threadIdx.x / QW, ch = quarter * QW + chl;
  const int kc = chl / (QW / 4);  // cell group inside the quarter
  const size_t row0 = static_cast<size_t>(b) * P;
  const float count_g = static_cast<float>(P) * QW, count_c = static_cast<float>(P) * (QW / 4);

  // ---- pass 1: statistics of the 2 x 4 gate groups
  float v[16];
#pragma unroll
  for (int n = 0; n < 16; ++n) v[n] = 0.f;
  for (int p = r0; p < P; p += R) {
    const size_t o = (row0 + p) * 4 * hid + 4 * ch;
    const float4 x = *reinterpret_cast<const float4*>(a.raw_ih + o);
    const float4 y = *reinterpret_cast<const float4*>(a.raw_hh + o);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float xv = comp(x, k), yv = comp(y, k);
      v[k] += xv; v[4 + k] += xv * xv; v[8 + k] += yv; v[12 + k] += yv * yv;
    }
  }
  block_sum_vec<16>(v, sh);
  float mu_i[4], rs_i[4], mu_h[4], rs_h[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    mean_rstd(v[k], v[4 + k], count_g, mu_i[k], rs_i[k]);
    mean_rstd(v[8 + k], v[12 + k], count_g, mu_h[k], rs_h[k]);
  }
  float g_i[4], b_i[4], g_h[4], b_h[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    g_i[k] = a.params[a.g_ih + k * hid + ch]; b_i[k] = a.params[a.b_ih + k * hid + ch];
    g_h[k] = a.params[a.g_hh + k * hid + ch]; b_h[k] = a.params[a.b_hh + k * hid + ch];
  }

  // ---- pass 2: gates, pre-norm cell, statistics of the 4 cell groups
  float sc = 0.f, qc = 0.f;
  for (int p = r0; p < P; p += R) {
    const size_t o = (row0 + p) * 4 * hid + 4 * ch;
    const size_t oc = (row0 + p) * hid + ch;
    const float4 x = *reinterpret_cast<const float4*>(a.raw_ih + o);
    const float4 y = *reinterpret_cast<const float4*>(a.raw_hh + o);
    float pre[4];
#pragma unroll
    for (int k = 0; k < 4; ++k)
      pre[k] = (g_i[k] * ((comp(x, k) - mu_i[k]) * rs_i[k]) + b_i[k]) + (g_h[k] * ((comp(y, k) - mu_h[k]) * rs_h[k]) + b_h[k]);
    const float ig = sigmoid_libm(pre[0]), fg = sigmoid_libm(pre[1]), og = sigmoid_libm(pre[2]), gg = tanhf(pre[3]);
    const float cr = fg * a.c_prev[oc] + ig * gg;
    *reinterpret_cast<float4*>(a.gates + o) = make_float4(ig, fg, og, gg);
    a.c_raw[oc] = cr;
    sc += cr; qc += cr * cr;
  }
  float w[8];
#pragma unroll
  for (int k = 0; k < 4; ++k) { w[2 * k] = (k == kc) ? sc : 0.f; w[2 * k + 1] = (k == kc) ? qc : 0.f; }
  block_sum_vec<8>(w, sh);
  float mu_c[4], rs_c[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) mean_rstd(w[2 * k], w[2 * k + 1], count_c, mu_c[k], rs_c[k]);
  float my_mu = mu_c[0], my_rs = rs_c[0];
#pragma unroll
  for (int k = 1; k < 4; ++k) if (k == kc) { my_mu = mu_c[k]; my_rs = rs_c[k]; }
  const float g_c = a.params[a.g_c + ch], b_c = a.params[a.b_c + ch];

  // ---- pass 3: normalised cell, hidden state
  for (int p = r0; p < P; p += R) {
    const size_t o = (row0 + p) * 4 * hid + 4 * ch;
    const size_t oc = (row0 + p) * hid + ch;
    const float cn = g_c * ((a.c_raw[oc] - my_mu) * my_rs) + b_c;
    const float og = a.gates[o + 2];
    a.c_out[oc] = cn;
    a.h_out[oc] = __float2bfloat16(og * tanhf(cn));
  }
  if (threadIdx.x == 0) {
    float* st = a.stats + static_cast<size_t>(b) * 96;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      st[2 * stat_gate(0, k, quarter)] = mu_i[k]; st[2 * stat_gate(0, k, quarter) + 1] = rs_i[k];
      st[2 * stat_gate(1, k, quarter)] = mu_h[k]; st[2 * stat_gate(1, k, quarter) + 1] = rs_h[k];
      st[2 * stat_cell(quarter, k)] = mu_c[k]; st[2 * stat_cell(quarter, k) + 1] = rs_c[k];
    }
  }
}

__global__ void __launch_bounds__(kGnThreads)
gn_cell_bwd_kernel(const GnCellArgs a) {
  pdl_entry();
  __shared__ float sh[(kGnThreads / 32) * 16];
  __shared__ float s_par[14][kGnThreads];
  const int hid = a.hid, P = a.P, QW = hid / 4, R = kGnThreads / QW;
  const int b = blockIdx.x, quarter = blockIdx.y;
  const int chl = threadIdx.x % QW, r0 = threadIdx.x / QW, ch = quarter * QW + chl;
  const int kc = chl / (QW / 4);
  const size_t row0 = static_cast<size_t>(b) * P;
  const float count_g = static_cast<float>(P) * QW, count_c = static_cast<float>(P) * (QW / 4);
  const float* st = a.stats + static_cast<size_t>(b) * 96;
  float mu_i[4], rs_i[4], mu_h[4], rs_h[4], g_i[4], g_h[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    mu_i[k] = st[2 * stat_gate(0, k, quarter)]; rs_i[k] = st[2 * stat_gate(0, k, quarter) + 1];
    mu_h[k] = st[2 * stat_gate(1, k, quarter)]; rs_h[k] = st[2 * stat_gate(1, k, quarter) + 1];
    g_i[k] = a.params[a.g_ih + k * hid + ch];
    g_h[k] = a.params[a.g_hh + k * hid + ch];
  }
  const float mu_c = st[2 * stat_cell(quarter, kc)], rs_c = st[2 * stat_cell(quarter, kc) + 1];
  const float g_c = a.params[a.g_c + ch];

  // ---- pass A: cell-norm backward reductions; d gamma_c / d beta_c of this sample
  float sA1 = 0.f, sA2 = 0.f, pgc = 0.f, pbc = 0.f;
  for (int p = r0; p < P; p += R) {
    const size_t o = (row0 + p) * 4 * hid + 4 * ch;
    const size_t oc = (row0 + p) * hid + ch;
    const float og = a.gates[o + 2];
    const float tc = tanhf(a.c_out[oc]);
    const float dcn = a.dc[oc] + a.dh[oc] * og * (1.f - tc * tc);
    const float chat = (a.c_raw[oc] - mu_c) * rs_c;
    sA1 += g_c * dcn; sA2 += g_c * dcn * chat;
    pgc += dcn * chat; pbc += dcn;
  }
  float w[8];
#pragma unroll
  for (int k = 0; k < 4; ++k) { w[2 * k] = (k == kc) ? sA1 : 0.f; w[2 * k + 1] = (k == kc) ? sA2 : 0.f; }
  block_sum_vec<8>(w, sh);
  float m1 = w[0], m2 = w[1];
#pragma unroll
  for (int k = 1; k < 4; ++k) if (k == kc) { m1 = w[2 * k]; m2 = w[2 * k + 1]; }
  m1 /= count_c; m2 /= count_c;

  // ---- pass B: gate pre-activation gradients (= dY of both GroupNorms), their reductions, d affine partials
  float v[16];
#pragma unroll
  for (int n = 0; n < 16; ++n) v[n] = 0.f;
  float pg_i[4] = {0.f, 0.f, 0.f, 0.f}, pg_h[4] = {0.f, 0.f, 0.f, 0.f}, pb[4] = {0.f, 0.f, 0.f, 0.f};
  for (int p = r0; p < P; p += R) {
    const size_t o = (row0 + p) * 4 * hid + 4 * ch;
    const size_t oc = (row0 + p) * hid + ch;
    const float4 gt = *reinterpret_cast<const float4*>(a.gates + o);  // i, f, o, g (post-activation)
    const float tc = tanhf(a.c_out[oc]);
    const float dhv = a.dh[oc];
    const float dcn = a.dc[oc] + dhv * gt.z * (1.f - tc * tc);
    const float chat = (a.c_raw[oc] - mu_c) * rs_c;
    const float dcr = rs_c * (g_c * dcn - m1 - chat * m2);
    float dy[4];
    dy[0] = dcr * gt.w * gt.x * (1.f - gt.x);
    dy[1] = dcr * a.c_prev[oc] * gt.y * (1.f - gt.y);
    dy[2] = dhv * tc * gt.z * (1.f - gt.z);
    dy[3] = dcr * gt.x * (1.f - gt.w * gt.w);
    a.dc[oc] = dcr * gt.y;  // gradient w.r.t. the previous step's (normalised) cell
    *reinterpret_cast<float4*>(a.dy + o) = make_float4(dy[0], dy[1], dy[2], dy[3]);
    const float4 x = *reinterpret_cast<const float4*>(a.raw_ih + o);
    const float4 y = *reinterpret_cast<const float4*>(a.raw_hh + o);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float xh = (comp(x, k) - mu_i[k]) * rs_i[k], yh = (comp(y, k) - mu_h[k]) * rs_h[k];
      v[k] += g_i[k] * dy[k]; v[4 + k] += g_i[k] * dy[k] * xh;
      v[8 + k] += g_h[k] * dy[k]; v[12 + k] += g_h[k] * dy[k] * yh;
      pg_i[k] += dy[k] * xh; pg_h[k] += dy[k] * yh; pb[k] += dy[k];
    }
  }
  block_sum_vec<16>(v, sh);
#pragma unroll
  for (int n = 0; n < 16; ++n) v[n] /= count_g;

  // ---- pass C: gradients w.r.t. the raw convolution outputs (bf16 GEMM operands of the two conv backward passes)
  for (int p = r0; p < P; p += R) {
    const size_t o = (row0 + p) * 4 * hid + 4 * ch;
    const float4 d4 = *reinterpret_cast<const float4*>(a.dy + o);
    const float4 x = *reinterpret_cast<const float4*>(a.raw_ih + o);
    const float4 y = *reinterpret_cast<const float4*>(a.raw_hh + o);
    float di[4], dh4[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float dyk = comp(d4, k);
      const float xh = (comp(x, k) - mu_i[k]) * rs_i[k], yh = (comp(y, k) - mu_h[k]) * rs_h[k];
      di[k] = rs_i[k] * (g_i[k] * dyk - v[k] - xh * v[4 + k]);
      dh4[k] = rs_h[k] * (g_h[k] * dyk - v[8 + k] - yh * v[12 + k]);
    }
    *reinterpret_cast<uint2*>(a.d_ih + o) = make_uint2(pack_bf16x2(di[0], di[1]), pack_bf16x2(di[2], di[3]));
    *reinterpret_cast<uint2*>(a.d_hh + o) = make_uint2(pack_bf16x2(dh4[0], dh4[1]), pack_bf16x2(dh4[2], dh4[3]));
  }

  // ---- affine-parameter gradient partials of this sample: sum over the position lanes of each channel, in order
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    s_par[k][threadIdx.x] = pg_i[k]; s_par[4 + k][threadIdx.x] = pb[k]; s_par[8 + k][threadIdx.x] = pg_h[k];
  }
  s_par[12][threadIdx.x] = pgc; s_par[13][threadIdx.x] = pbc;
  __syncthreads();
  if (r0 == 0) {
    float* part = a.part + static_cast<size_t>(b) * 14 * hid;
#pragma unroll
    for (int j = 0; j < 14; ++j) {
      float t = 0.f;
      for (int r = 0; r < R; ++r) t += s_par[j][r * QW + chl];
      if (j < 12) part[(j >> 2) * 4 * hid + (j & 3) * hid + ch] = t;  // [d gamma_ih | d beta | d gamma_hh][gate * hid + ch]
      else part[12 * hid + (j - 12) * hid + ch] = t;                    // [d gamma_c | d beta_c][ch]
    }
  }
}

// grads += sum over samples of the per-sample partials, in sample order
__global__ void __launch_bounds__(256)
gn_affine_fold_kernel(const float* __restrict__ part, int B, int hid, float* __restrict__ grads, long long g_ih,
                      long long b_ih, long long g_hh, long long b_hh, long long g_c, long long b_c) {
  pdl_entry();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 14 * hid) return;
  float t = 0.f;
  for (int b = 0; b < B; ++b) t += part[static_cast<size_t>(b) * 14 * hid + i];
  if (i < 4 * hid) grads[g_ih + i] += t;
  else if (i < 8 * hid) { grads[b_ih + (i - 4 * hid)] += t; grads[b_hh + (i - 4 * hid)] += t; }
  else if (i < 12 * hid) grads[g_hh + (i - 8 * hid)] += t;
  else if (i < 13 * hid) grads[g_c + (i - 12 * hid)] += t;
  else grads[b_c + (i - 13 * hid)] += t;
}

bool gn_shape_ok(const GnCellArgs& a) {
  const int qw = a.hid / 4;
  return a.hid % 16 == 0 && qw >= 4 && qw <= kGnThreads && kGnThreads % qw == 0 && a.B > 0 && a.P > 0;
}

}  // namespace

cudaError_t launch_gn_cell_fwd(const GnCellArgs& a, cudaStream_t s) {
  if (!gn_shape_ok(a)) return cudaErrorInvalidValue;
  if (cudaError_t e_ = launch_pdl_small(gn_cell_fwd_kernel, dim3(a.B, 4), dim3(kGnThreads), 0, s, a); e_ != cudaSuccess) return e_;
  return cudaGetLastError();
}

cudaError_t launch_gn_cell_bwd(const GnCellArgs& a, float* grads, cudaStream_t s) {
  if (!gn_shape_ok(a) || !a.part || !a.dy || !a.d_ih || !a.d_hh || !a.dh || !a.dc) return cudaErrorInvalidValue;
  if (cudaError_t e_ = launch_pdl_small(gn_cell_bwd_kernel, dim3(a.B, 4), dim3(kGnThreads), 0, s, a); e_ != cudaSuccess) return e_;
  if (cudaError_t e_ = launch_pdl_small(gn_affine_fold_kernel, dim3((14 * a.hid + 255) / 256), dim3(256), 0, s, a.part, a.B, a.hid, grads, a.g_ih, a.b_ih, a.g_hh,
                                                                a.b_hh, a.g_c, a.b_c); e_ != cudaSuccess) return e_;
  return cudaGetLastError();
}

}  // namespace rac
