// NormConvLSTMCell pointwise part (reference src/prediction/models/lstm.py:177-198, lstm_group_norm=True):
//   gates = GroupNorm16(ih_conv(x)) + GroupNorm16(hh_conv(h_prev));  i,f,o = sigmoid, g = tanh
//   c = GroupNorm16(f * c_prev + i * g);  h = o * tanh(c)
// The two gate convolutions run on the tensor cores (conv_tc_kernel<.., EPI_F32>) and leave their raw fp32 outputs
// [B, P, 4*hid] in HBM with packed column order (channel, gate). GroupNorm statistics are per SAMPLE over
// (channels of the group) x (P positions), so one CTA owns one sample: pass 1 reduces the 2 x 16 gate-group statistics,
// pass 2 re-reads the gates, forms the pre-norm cell state and the output gate and reduces the 16 cell-group
// statistics, pass 3 normalises the cell and writes c (fp32) / h (bf16). Between pass 2 and 3 the pre-norm cell state
// lives in c_state itself and the output gate in the (already consumed) `out` slot of the ih buffer -- every thread
// re-reads only what it wrote, and the kernel needs 40 KB of shared memory instead of 196 KB, so 2 CTAs share an SM.
// All reductions have a fixed order (deterministic).
// Measured at 2000 samples, g = 512: 1.33 ms per launch (1.49 ms with the 196 KB single-CTA staging). The kernel is
// latency-bound, not bandwidth-bound (3.6 GB per launch = 0.6 ms): a variant with one 8-CTA cluster per sample that
// keeps the gates in registers and reduces the statistics through distributed shared memory reads them only once but
// measured 1.59 ms (two cluster-wide barriers per 6-position slice, one resident CTA per SM); fusing the gate
// statistics into the convolution epilogues is the remaining fix.
#include "misc_kernels.cuh"

namespace rac {

namespace {

__device__ __forceinline__ float sigmoid_exact(float x) { return 1.0f / (1.0f + expf(-x)); }

// Sum of the `n` floats scr[i * stride], i in [0, n), by one warp in a fixed order.
__device__ __forceinline__ double warp_sum_strided(const float* scr, int n, int stride, int lane) {
  double acc = 0.0;
  for (int i = lane; i < n; i += 32) acc += static_cast<double>(scr[static_cast<size_t>(i) * stride]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  return acc;
}

// block = hid * R threads (R = rows of positions processed concurrently); thread t owns channel t % hid
__global__ void __launch_bounds__(512, 2)
norm_lstm_cell_kernel(float* __restrict__ ih, const float* __restrict__ hh, const float* __restrict__ gnp,
                      float* __restrict__ c_state, __nv_bfloat16* __restrict__ h_out, int P, int hid, int stats_off,
                      float eps) {
  extern __shared__ float smem[];
  const int T = blockDim.x;
  const int R = T / hid;
  const int ch = threadIdx.x % hid;
  const int r0 = threadIdx.x / hid;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = T >> 5;
  float* stats = smem + stats_off;                     // mean[48], rstd[48]: ih groups 0..15, hh 16..31, cell 32..47
  float* scr = smem;                                   // pass-1 scratch: [16][T]
  const size_t b = blockIdx.x;
  float4* ih4 = reinterpret_cast<float4*>(ih + b * P * 4 * hid);
  const float4* hh4 = reinterpret_cast<const float4*>(hh + b * P * 4 * hid);

  // ---- pass 1: per-thread partial sums of the 4 gates of its channel, both tensors
  float s[2][4] = {}, q[2][4] = {};
#pragma unroll 4  // 8 independent 16-byte loads in flight per thread: the pass is latency-bound otherwise (1 CTA / SM)
  for (int p = r0; p < P; p += R) {
    const float4 a = ih4[static_cast<size_t>(p) * hid + ch];
    const float4 c = hh4[static_cast<size_t>(p) * hid + ch];
    s[0][0] += a.x; s[0][1] += a.y; s[0][2] += a.z; s[0][3] += a.w;
    q[0][0] += a.x * a.x; q[0][1] += a.y * a.y; q[0][2] += a.z * a.z; q[0][3] += a.w * a.w;
    s[1][0] += c.x; s[1][1] += c.y; s[1][2] += c.z; s[1][3] += c.w;
    q[1][0] += c.x * c.x; q[1][1] += c.y * c.y; q[1][2] += c.z * c.z; q[1][3] += c.w * c.w;
  }
  // scratch layout: value v = (tensor * 4 + gate) * 2 + {sum, sumsq}; entry index e = r0 * hid + ch
#pragma unroll
  for (int t = 0; t < 2; ++t)
#pragma unroll
    for (int gte = 0; gte < 4; ++gte) {
      scr[static_cast<size_t>((t * 4 + gte) * 2 + 0) * T + threadIdx.x] = s[t][gte];
      scr[static_cast<size_t>((t * 4 + gte) * 2 + 1) * T + threadIdx.x] = q[t][gte];
    }
  __syncthreads();
  // group (tensor, gate, quarter) = reference GroupNorm group gate * 4 + quarter of channels [quarter * hid/4, ...)
  const int qw = hid / 4;
  for (int job = warp; job < 32; job += nwarps) {
    const int t = job >> 4, gte = (job >> 2) & 3, quarter = job & 3;
    double sum = 0.0, sq = 0.0;
    for (int r = 0; r < R; ++r) {  // entries of row r: channels quarter*qw .. +qw-1 are contiguous
      sum += warp_sum_strided(scr + static_cast<size_t>((t * 4 + gte) * 2 + 0) * T + r * hid + quarter * qw, qw, 1, lane);
      sq += warp_sum_strided(scr + static_cast<size_t>((t * 4 + gte) * 2 + 1) * T + r * hid + quarter * qw, qw, 1, lane);
    }
    if (lane == 0) {
      const double n = static_cast<double>(P) * qw;
      const double mean = sum / n;
      double var = sq / n - mean * mean;
      if (var < 0.0) var = 0.0;
      stats[job] = static_cast<float>(mean);
      stats[48 + job] = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
    }
  }
  __syncthreads();

  // ---- pass 2: normalise + affine, gates, pre-norm cell state
  const int quarter = ch / qw;
  const float4* gp4 = reinterpret_cast<const float4*>(gnp);
  const float4 ga = gp4[ch], ba = gp4[hid + ch], gb = gp4[2 * hid + ch], bb = gp4[3 * hid + ch];
  float ma[4], ra[4], mb[4], rb[4];
#pragma unroll
  for (int gte = 0; gte < 4; ++gte) {
    ma[gte] = stats[gte * 4 + quarter]; ra[gte] = stats[48 + gte * 4 + quarter];
    mb[gte] = stats[16 + gte * 4 + quarter]; rb[gte] = stats[48 + 16 + gte * 4 + quarter];
  }
  float cs = 0.f, cq = 0.f;
  float* cst = c_state + b * P * hid;
#pragma unroll 4
  for (int p = r0; p < P; p += R) {
    const float4 a = ih4[static_cast<size_t>(p) * hid + ch];
    const float4 c = hh4[static_cast<size_t>(p) * hid + ch];
    const float gi = ((a.x - ma[0]) * ra[0] * ga.x + ba.x) + ((c.x - mb[0]) * rb[0] * gb.x + bb.x);
    const float gf = ((a.y - ma[1]) * ra[1] * ga.y + ba.y) + ((c.y - mb[1]) * rb[1] * gb.y + bb.y);
    const float go = ((a.z - ma[2]) * ra[2] * ga.z + ba.z) + ((c.z - mb[2]) * rb[2] * gb.z + bb.z);
    const float gc = ((a.w - ma[3]) * ra[3] * ga.w + ba.w) + ((c.w - mb[3]) * rb[3] * gb.w + bb.w);
    const float cp = sigmoid_exact(gf) * cst[static_cast<size_t>(p) * hid + ch] + sigmoid_exact(gi) * tanhf(gc);
    cst[static_cast<size_t>(p) * hid + ch] = cp;                                              // pre-norm cell state
    reinterpret_cast<float*>(ih4 + static_cast<size_t>(p) * hid + ch)[2] = sigmoid_exact(go);  // output gate
    cs += cp;
    cq += cp * cp;
  }
  // cell GroupNorm(16, hid): group = ch / (hid / 16); per-thread partials -> stats scratch behind the stats block
  float* cscr = stats + 96;  // [2][T]
  cscr[threadIdx.x] = cs;
  cscr[T + threadIdx.x] = cq;
  __syncthreads();
  const int cw = hid / 16;
  for (int job = warp; job < 16; job += nwarps) {
    double sum = 0.0, sq = 0.0;
    for (int r = 0; r < R; ++r) {
      sum += warp_sum_strided(cscr + r * hid + job * cw, cw, 1, lane);
      sq += warp_sum_strided(cscr + T + r * hid + job * cw, cw, 1, lane);
    }
    if (lane == 0) {
      const double n = static_cast<double>(P) * cw;
      const double mean = sum / n;
      double var = sq / n - mean * mean;
      if (var < 0.0) var = 0.0;
      stats[32 + job] = static_cast<float>(mean);
      stats[48 + 32 + job] = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
    }
  }
  __syncthreads();

  // ---- pass 3: cell GroupNorm, hidden state
  const float cg = gnp[16 * hid + ch], cb = gnp[17 * hid + ch];
  const float mc = stats[32 + ch / cw], rc = stats[48 + 32 + ch / cw];
  __nv_bfloat16* ho = h_out + b * P * hid;
#pragma unroll 4
  for (int p = r0; p < P; p += R) {
    const size_t i = static_cast<size_t>(p) * hid + ch;
    const float c = (cst[i] - mc) * rc * cg + cb;
    cst[i] = c;
    ho[i] = __float2bfloat16(reinterpret_cast<const float*>(ih4 + i)[2] * tanhf(c));
  }
}

// Variant for gate convolutions that ran with the EPI_GATES epilogue: the GroupNorm sums of both gate tensors arrive
// as per-tile partials (gn_part[b][tensor][group][slot][2], written by conv_tc_kernel), so the 786 KB of gates of a
// sample are read ONCE. Phase A forms the pre-norm cell state (-> c_state) and the output gate (-> obuf, fp32) and
// reduces the cell statistics; phase B re-reads those 2 x 98 KB (L1 / L2 resident) and writes c / h.
__global__ void __launch_bounds__(512, 2)
norm_lstm_cell_fused_kernel(const float* __restrict__ ih, const float* __restrict__ hh, const float* __restrict__ part,
                            int nslots, const float* __restrict__ gnp, float* __restrict__ c_state,
                            float* __restrict__ obuf, __nv_bfloat16* __restrict__ h_out, int P, int hid, float eps) {
  extern __shared__ float smem[];
  float* stats = smem;       // mean[48], rstd[48]
  float* cscr = smem + 96;   // [2][T]
  const int T = blockDim.x;
  const int R = T / hid;
  const int ch = threadIdx.x % hid;
  const int r0 = threadIdx.x / hid;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = T >> 5;
  const size_t b = blockIdx.x;
  const float4* ih4 = reinterpret_cast<const float4*>(ih + b * P * 4 * hid);
  const float4* hh4 = reinterpret_cast<const float4*>(hh + b * P * 4 * hid);
  const int qw = hid / 4;
  if (threadIdx.x < 32) {  // (tensor, group): sum of the tile partials in slot order
    const float* p = part + (b * 32 + threadIdx.x) * 6 * 2;
    double sum = 0.0, sq = 0.0;
    for (int s = 0; s < nslots; ++s) {
      sum += static_cast<double>(p[2 * s]);
      sq += static_cast<double>(p[2 * s + 1]);
    }
    const double n = static_cast<double>(P) * qw;
    const double mean = sum / n;
    double var = sq / n - mean * mean;
    if (var < 0.0) var = 0.0;
    stats[threadIdx.x] = static_cast<float>(mean);
    stats[48 + threadIdx.x] = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
  }
  __syncthreads();
  const int quarter = ch / qw;
  const float4* gp4 = reinterpret_cast<const float4*>(gnp);
  const float4 ga = gp4[ch], ba = gp4[hid + ch], gb = gp4[2 * hid + ch], bb = gp4[3 * hid + ch];
  float ma[4], ra[4], mb[4], rb[4];
#pragma unroll
  for (int gte = 0; gte < 4; ++gte) {
    ma[gte] = stats[gte * 4 + quarter]; ra[gte] = stats[48 + gte * 4 + quarter];
    mb[gte] = stats[16 + gte * 4 + quarter]; rb[gte] = stats[48 + 16 + gte * 4 + quarter];
  }
  float cs = 0.f, cq = 0.f;
  float* cst = c_state + b * P * hid;
  float* ob = obuf + b * P * hid;
#pragma unroll 4
  for (int p = r0; p < P; p += R) {
    const size_t i = static_cast<size_t>(p) * hid + ch;
    const float4 a = __ldcs(ih4 + i);
    const float4 c = __ldcs(hh4 + i);
    const float gi = ((a.x - ma[0]) * ra[0] * ga.x + ba.x) + ((c.x - mb[0]) * rb[0] * gb.x + bb.x);
    const float gf = ((a.y - ma[1]) * ra[1] * ga.y + ba.y) + ((c.y - mb[1]) * rb[1] * gb.y + bb.y);
    const float go = ((a.z - ma[2]) * ra[2] * ga.z + ba.z) + ((c.z - mb[2]) * rb[2] * gb.z + bb.z);
    const float gc = ((a.w - ma[3]) * ra[3] * ga.w + ba.w) + ((c.w - mb[3]) * rb[3] * gb.w + bb.w);
    const float cp = sigmoid_exact(gf) * cst[i] + sigmoid_exact(gi) * tanhf(gc);
    cst[i] = cp;
    ob[i] = sigmoid_exact(go);
    cs += cp;
    cq += cp * cp;
  }
  cscr[threadIdx.x] = cs;
  cscr[T + threadIdx.x] = cq;
  __syncthreads();
  const int cw = hid / 16;
  for (int job = warp; job < 16; job += nwarps) {
    double sum = 0.0, sq = 0.0;
    for (int r = 0; r < R; ++r) {
      sum += warp_sum_strided(cscr + r * hid + job * cw, cw, 1, lane);
      sq += warp_sum_strided(cscr + T + r * hid + job * cw, cw, 1, lane);
    }
    if (lane == 0) {
      const double n = static_cast<double>(P) * cw;
      const double mean = sum / n;
      double var = sq / n - mean * mean;
      if (var < 0.0) var = 0.0;
      stats[32 + job] = static_cast<float>(mean);
      stats[48 + 32 + job] = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
    }
  }
  __syncthreads();
  const float cg = gnp[16 * hid + ch], cb = gnp[17 * hid + ch];
  const float mc = stats[32 + ch / cw], rc = stats[48 + 32 + ch / cw];
  __nv_bfloat16* ho = h_out + b * P * hid;
#pragma unroll 4
  for (int p = r0; p < P; p += R) {
    const size_t i = static_cast<size_t>(p) * hid + ch;
    const float c = (cst[i] - mc) * rc * cg + cb;
    cst[i] = c;
    ho[i] = __float2bfloat16(ob[i] * tanhf(c));
  }
}

}  // namespace

static size_t norm_lstm_stats_off(int, int, int T) { return static_cast<size_t>(16) * T; }  // after the pass-1 scratch

cudaError_t norm_lstm_set_attributes() {
  return cudaFuncSetAttribute(norm_lstm_cell_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
}

cudaError_t launch_norm_lstm_cell_fused(const float* ih, const float* hh, const float* gn_part, int nslots,
                                        const float* gn_params, float* c_state, float* obuf, __nv_bfloat16* h_out,
                                        int B, int P, int hid, cudaStream_t s) {
  if (hid % 64 != 0 || hid > 512 || nslots < 1 || nslots > 6) return cudaErrorInvalidValue;
  const int R = hid >= 512 ? 1 : 512 / hid;
  const int T = hid * R;
  if (T > 512) return cudaErrorInvalidValue;
  const size_t smem = (96 + 2 * static_cast<size_t>(T)) * sizeof(float);
  norm_lstm_cell_fused_kernel<<<B, T, smem, s>>>(ih, hh, gn_part, nslots, gn_params, c_state, obuf, h_out, P, hid, 1e-5f);
  return cudaGetLastError();
}

cudaError_t launch_norm_lstm_cell(float* ih, const float* hh, const float* gn_params, float* c_state,
                                  __nv_bfloat16* h_out, int B, int P, int hid, cudaStream_t s) {
  if (hid % 64 != 0 || hid > 512) return cudaErrorInvalidValue;
  const int R = hid >= 512 ? 1 : 512 / hid;
  const int T = hid * R;
  const size_t off = norm_lstm_stats_off(P, hid, T);
  const size_t smem = (off + 96 + 2 * static_cast<size_t>(T)) * 4;
  if (T > 512 || smem > 227 * 1024) return cudaErrorInvalidValue;
  norm_lstm_cell_kernel<<<B, T, smem, s>>>(ih, hh, gn_params, c_state, h_out, P, hid, static_cast<int>(off), 1e-5f);
  return cudaGetLastError();
}

}  // namespace rac
