// Fused epilogues of the convolution kernels. One thread owns one output row (b, y, x) and receives CH consecutive
// fp32 accumulator columns starting at packed column n0. The same functions serve the tcgen05 kernel (accumulators
// read from TMEM) and the SIMT cross-check kernel.
#pragma once
#include "conv.cuh"
#include "ptx.cuh"

namespace rac {

// ---------------------------------------------------------------- Philox4x32-10 counter RNG
struct Philox4 {
  uint32_t v[4];
};
__host__ __device__ __forceinline__ Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                          uint32_t k0, uint32_t k1) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint64_t p0 = static_cast<uint64_t>(M0) * c0;
    const uint64_t p1 = static_cast<uint64_t>(M1) * c2;
    const uint32_t n0 = static_cast<uint32_t>(p1 >> 32) ^ c1 ^ k0;
    const uint32_t n1 = static_cast<uint32_t>(p1);
    const uint32_t n2 = static_cast<uint32_t>(p0 >> 32) ^ c3 ^ k1;
    const uint32_t n3 = static_cast<uint32_t>(p0);
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += W0; k1 += W1;
  }
  Philox4 o;
  o.v[0] = c0; o.v[1] = c1; o.v[2] = c2; o.v[3] = c3;
  return o;
}
// two uniform 32-bit words -> two standard normals (Box-Muller)
__device__ __forceinline__ void box_muller(uint32_t a, uint32_t b, float& n0, float& n1) {
  const float u1 = (static_cast<float>(a) + 1.0f) * 2.3283064365386963e-10f;  // (0, 1]
  const float u2 = static_cast<float>(b) * 2.3283064365386963e-10f;           // [0, 1)
  const float r = sqrtf(-2.0f * logf(u1));
  float s, c;
  sincosf(6.283185307179586f * u2, &s, &c);
  n0 = r * c;
  n1 = r * s;
}

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }
// MUFU.TANH: max relative error 2^-11, well below the bf16 rounding (2^-9) of the hidden state it feeds
__device__ __forceinline__ float tanh_fast(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float sigmoid_fast(float x) { return fmaf(0.5f, tanh_fast(0.5f * x), 0.5f); }

__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

// ---------------------------------------------------------------- EPI_ACT
// bias + LeakyReLU(0.2) + bf16 pack of 8 consecutive columns: the epilogue of the 256-column tile is exposed and
// instruction-issue bound (8 warps), so this is 2 x LDG.128 + per value FADD, FMUL, FMNMX (+ half a pack)
__device__ __forceinline__ uint4 act_pack8(const EpiParams& e, int n, const float* a) {
  const float4 b0 = __ldg(reinterpret_cast<const float4*>(e.bias + n));
  const float4 b1 = __ldg(reinterpret_cast<const float4*>(e.bias + n + 4));
  float v[8] = {a[0] + b0.x, a[1] + b0.y, a[2] + b0.z, a[3] + b0.w, a[4] + b1.x, a[5] + b1.y, a[6] + b1.z, a[7] + b1.w};
  if (e.lrelu) {
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = fmaxf(v[j], 0.2f * v[j]);  // == v > 0 ? v : 0.2 v
  }
  return make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
}

template <int CH>
__device__ __forceinline__ void epi_act(const ConvGeom& g, const EpiParams& e, int b, int y, int x, bool valid, int n0,
                                        const float* acc) {
  if (!valid || n0 + CH > e.cout) return;
  uint32_t pk[CH / 2];
#pragma unroll
  for (int q = 0; q < CH / 8; ++q) {
    const uint4 t = act_pack8(e, n0 + 8 * q, acc + 8 * q);
    pk[4 * q] = t.x; pk[4 * q + 1] = t.y; pk[4 * q + 2] = t.z; pk[4 * q + 3] = t.w;
  }
  if (!e.upsample) {
    __nv_bfloat16* dst = e.out + (static_cast<size_t>(b * g.H + y) * g.W + x) * e.out_cstride + e.out_coff + n0;
#pragma unroll
    for (int q = 0; q < CH / 8; ++q)
      reinterpret_cast<uint4*>(dst)[q] = make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
  } else {
    // nearest 2x upsample (reference vgg_64.py:221,235-239) folded into the store
#pragma unroll
    for (int dy = 0; dy < 2; ++dy)
#pragma unroll
      for (int dx = 0; dx < 2; ++dx) {
        __nv_bfloat16* dst =
            e.out + (static_cast<size_t>(b * 2 * g.H + 2 * y + dy) * (2 * g.W) + 2 * x + dx) * e.out_cstride +
            e.out_coff + n0;
#pragma unroll
        for (int q = 0; q < CH / 8; ++q)
          reinterpret_cast<uint4*>(dst)[q] = make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
      }
  }
}

// ---------------------------------------------------------------- EPI_ACT, coalesced variant (tcgen05 kernels)
// One thread owns one output row, so a plain per-thread store sends the 32 lanes of a warp to 32 different 128-byte
// lines with 16 B each: the L1 processes 32 requests per instruction and the (exposed) epilogue of the 256-column tile
// became store-issue bound. Here 64 channels (two 32-column chunks = 8 units of 16 B) are first transposed inside
// groups of 8 lanes (3 butterfly stages of warp shuffles), after which lane j of a group holds unit j of all 8 rows of
// the group: every store instruction then writes 4 complete 128-byte segments (one per group).
__device__ __forceinline__ void act_pack32(const EpiParams& e, int n0, const float* acc, uint4* u) {
#pragma unroll
  for (int q = 0; q < 4; ++q) u[q] = act_pack8(e, n0 + 8 * q, acc + 8 * q);
}
__device__ __forceinline__ uint4 shfl_xor_u4(uint4 v, int m) {
  return make_uint4(__shfl_xor_sync(0xffffffffu, v.x, m), __shfl_xor_sync(0xffffffffu, v.y, m),
                    __shfl_xor_sync(0xffffffffu, v.z, m), __shfl_xor_sync(0xffffffffu, v.w, m));
}
// u[k] of lane (8G + i)  ->  u[i] of lane (8G + k)
__device__ __forceinline__ void transpose8x8_u4(uint4* u, int lane) {
#pragma unroll
  for (int bit = 1; bit < 8; bit <<= 1) {
    const bool hi = (lane & bit) != 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      if (k & bit) continue;
      const uint4 send = hi ? u[k] : u[k | bit];
      const uint4 recv = shfl_xor_u4(send, bit);
      if (hi) u[k] = recv; else u[k | bit] = recv;
    }
  }
}
// RowFn: (row index inside the tile) -> (b, y, x, valid). Must be called by all 32 lanes (shuffles). u[0..7] = the 64
// packed channels n0 .. n0+63 of this thread's row r.
template <typename RowFn>
__device__ __forceinline__ void epi_act_store64(const ConvGeom& g, const EpiParams& e, int r, int n0, uint4* u,
                                                RowFn row_of) {
  const int lane = threadIdx.x & 31;
  transpose8x8_u4(u, lane);
  if (n0 + 64 > e.cout) return;
  const int rbase = r - (lane & 7);
  const int unit = lane & 7;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    int b, y, x;
    bool valid;
    row_of(rbase + i, b, y, x, valid);
    if (!valid) continue;
    if (!e.upsample) {
      __nv_bfloat16* dst = e.out + (static_cast<size_t>(b * g.H + y) * g.W + x) * e.out_cstride + e.out_coff + n0 + unit * 8;
      *reinterpret_cast<uint4*>(dst) = u[i];
    } else {
#pragma unroll
      for (int dy = 0; dy < 2; ++dy)
#pragma unroll
        for (int dx = 0; dx < 2; ++dx) {
          __nv_bfloat16* dst =
              e.out + (static_cast<size_t>(b * 2 * g.H + 2 * y + dy) * (2 * g.W) + 2 * x + dx) * e.out_cstride +
              e.out_coff + n0 + unit * 8;
          *reinterpret_cast<uint4*>(dst) = u[i];
        }
    }
  }
}

// ---------------------------------------------------------------- EPI_LSTM
// columns n0 .. n0+31 = 8 channels x (in, remember, out, cell). The previous cell state of a chunk is fetched with
// lstm_load_c one chunk ahead so that its global-load latency hides behind the gate math of the current chunk.
__device__ __forceinline__ bool lstm_chunk_live(const EpiParams& e, bool valid, int n0) {
  return valid && n0 + 32 <= e.cout;
}
// Cell-state layout. Training keeps c in NHWC [B,H,W,hid] (its backward kernels read it). Inference (e.c_tiled) uses a
// layout private to this epilogue -- c is read and written by nothing else: [m_tile][8-channel group][half][row][4],
// so that the 32 lanes of a warp (32 consecutive tile rows) touch 512 contiguous bytes per 16-byte access instead of
// 32 different 128-byte lines (row stride hid * 4 B), which made the exposed epilogue of the 256x256 tile store-bound.
// `ctile` = float offset of (m_tile, row) in that layout: (m_tile * (hid / 8) * 2 * BLOCK_M + row) * 4; bm = BLOCK_M.
template <bool TRAIN>
__device__ __forceinline__ void lstm_load_c(const ConvGeom& g, const EpiParams& e, int b, int y, int x, bool valid,
                                            int n0, float* cprev, size_t ctile, int bm) {
  if (!lstm_chunk_live(e, valid, n0)) return;
  if (!TRAIN && e.c_tiled) {
    const float* src = e.c_state + ctile + static_cast<size_t>(n0 >> 5) * (8 * bm);
    *reinterpret_cast<float4*>(cprev) = *reinterpret_cast<const float4*>(src);
    *reinterpret_cast<float4*>(cprev + 4) = *reinterpret_cast<const float4*>(src + 4 * bm);
    return;
  }
  const size_t base = (static_cast<size_t>(b * g.H + y) * g.W + x) * e.hid + (n0 >> 2);
  const float* src = (TRAIN && e.c_in) ? e.c_in : e.c_state;
  *reinterpret_cast<float4*>(cprev) = *reinterpret_cast<const float4*>(src + base);
  *reinterpret_cast<float4*>(cprev + 4) = *reinterpret_cast<const float4*>(src + base + 4);
}
// bias: the 32 bias values of this chunk (global or shared memory)
template <bool TRAIN>
__device__ __forceinline__ void epi_lstm(const ConvGeom& g, const EpiParams& e, int b, int y, int x, bool valid, int n0,
                                         const float* acc, const float* cprev, size_t ctile, int bm,
                                         const float* bias) {
  if (!lstm_chunk_live(e, valid, n0)) return;
  const size_t base = (static_cast<size_t>(b * g.H + y) * g.W + x) * e.hid + (n0 >> 2);
  const float4* bias4 = reinterpret_cast<const float4*>(bias);
  float cn[8], hn[8];
  float4* gsave = (TRAIN && e.gates_out)
                      ? reinterpret_cast<float4*>(e.gates_out + (static_cast<size_t>(b * g.H + y) * g.W + x) * (4 * e.hid) + n0)
                      : nullptr;
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const float4 bq = bias4[q];
    float ig, fg, og, cg;
    if (TRAIN) {
      ig = sigmoidf_(acc[4 * q + 0] + bq.x); fg = sigmoidf_(acc[4 * q + 1] + bq.y);
      og = sigmoidf_(acc[4 * q + 2] + bq.z); cg = tanhf(acc[4 * q + 3] + bq.w);
      cn[q] = fg * cprev[q] + ig * cg;
      hn[q] = og * tanhf(cn[q]);
    } else {
      ig = sigmoid_fast(acc[4 * q + 0] + bq.x); fg = sigmoid_fast(acc[4 * q + 1] + bq.y);
      og = sigmoid_fast(acc[4 * q + 2] + bq.z); cg = tanh_fast(acc[4 * q + 3] + bq.w);
      cn[q] = fg * cprev[q] + ig * cg;
      hn[q] = og * tanh_fast(cn[q]);
    }
    if (TRAIN && gsave) gsave[q] = make_float4(ig, fg, og, cg);
  }
  if (!TRAIN && e.c_tiled) {
    float* dst = e.c_state + ctile + static_cast<size_t>(n0 >> 5) * (8 * bm);
    *reinterpret_cast<float4*>(dst) = *reinterpret_cast<float4*>(cn);
    *reinterpret_cast<float4*>(dst + 4 * bm) = *reinterpret_cast<float4*>(cn + 4);
  } else {
    *reinterpret_cast<float4*>(e.c_state + base) = *reinterpret_cast<float4*>(cn);
    *reinterpret_cast<float4*>(e.c_state + base + 4) = *reinterpret_cast<float4*>(cn + 4);
  }
  *reinterpret_cast<uint4*>(e.h_out + base) = make_uint4(pack_bf16x2(hn[0], hn[1]), pack_bf16x2(hn[2], hn[3]),
                                                         pack_bf16x2(hn[4], hn[5]), pack_bf16x2(hn[6], hn[7]));
}

// ---------------------------------------------------------------- EPI_F32
// fp32 store / accumulate of CH columns into one of up to three destinations (training path)
template <int CH>
__device__ __forceinline__ void epi_f32(const ConvGeom& g, const EpiParams& e, int b, int y, int x, bool valid, int n0,
                                        const float* acc) {
  if (!valid || n0 + CH > e.cout) return;
#pragma unroll
  for (int s = 0; s < 3; ++s) {
    if (s >= e.nseg) break;
    const F32Seg& sg = e.seg[s];
    if (n0 < sg.n_begin || n0 >= sg.n_end) continue;
    if (sg.dst == nullptr) return;
    float* dst = sg.dst + (static_cast<size_t>(b * g.H + y) * g.W + x) * sg.cstride + sg.coff + (n0 - sg.n_begin);
    float4* d4 = reinterpret_cast<float4*>(dst);
#pragma unroll
    for (int q = 0; q < CH / 4; ++q) {
      float4 v = make_float4(acc[4 * q], acc[4 * q + 1], acc[4 * q + 2], acc[4 * q + 3]);
      if (e.bias) {
        v.x += __ldg(e.bias + n0 + 4 * q); v.y += __ldg(e.bias + n0 + 4 * q + 1);
        v.z += __ldg(e.bias + n0 + 4 * q + 2); v.w += __ldg(e.bias + n0 + 4 * q + 3);
      }
      if (sg.accumulate) {
        const float4 o = d4[q];
        v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w;
      }
      d4[q] = v;
    }
    return;
  }
}

// (the warp-collective transposed stores warp_store_rows_f32 / _v4 live in ptx.cuh: wgrad_tc.cu uses them too)
// epi_f32 / epi_split through the staged store: same destinations, same values (bias and accumulation included)
template <int CH>
__device__ __forceinline__ void epi_f32_staged(const ConvGeom& g, const EpiParams& e, int b, int y, int x, bool valid, int n0,
                                               int ncols, int split, const float* acc, float* stage, int lane) {
  static_assert(CH == 32, "the staged store handles 32-column chunks");
  const long long pix = static_cast<long long>(b * g.H + y) * g.W + x;
  if (g.epi_staged == 2) {  // (uniform) 128-bit variant, same control flow as below
    if (e.split_part != nullptr) {
      warp_store_rows_f32_v4(stage, e.split_part + static_cast<size_t>(split) * e.split_stride + n0,
                             valid ? pix * ncols : -1, acc, nullptr, false, lane);
      return;
    }
    if (n0 + CH > e.cout) return;
#pragma unroll
    for (int s = 0; s < 3; ++s) {
      if (s >= e.nseg) break;
      const F32Seg& sg = e.seg[s];
      if (n0 < sg.n_begin || n0 >= sg.n_end) continue;
      if (sg.dst == nullptr) return;
      warp_store_rows_f32_v4(stage, sg.dst + sg.coff + (n0 - sg.n_begin), valid ? pix * sg.cstride : -1, acc,
                             e.bias ? e.bias + n0 : nullptr, sg.accumulate != 0, lane);
      return;
    }
    return;
  }
  if (e.split_part != nullptr) {
    warp_store_rows_f32(stage, e.split_part + static_cast<size_t>(split) * e.split_stride + n0,
                        valid ? pix * ncols : -1, acc, nullptr, false, lane);
    return;
  }
  if (n0 + CH > e.cout) return;  // (uniform over the warp, like everything below that depends on n0 only)
#pragma unroll
  for (int s = 0; s < 3; ++s) {
    if (s >= e.nseg) break;
    const F32Seg& sg = e.seg[s];
    if (n0 < sg.n_begin || n0 >= sg.n_end) continue;
    if (sg.dst == nullptr) return;
    warp_store_rows_f32(stage, sg.dst + sg.coff + (n0 - sg.n_begin), valid ? pix * sg.cstride : -1, acc,
                        e.bias ? e.bias + n0 : nullptr, sg.accumulate != 0, lane);
    return;
  }
}

// split-K work item: raw accumulator chunk -> this split's slice (the reduce kernel applies bias and segments)
template <int CH>
__device__ __forceinline__ void epi_split(const ConvGeom& g, const EpiParams& e, int b, int y, int x, bool valid, int n0,
                                          int ncols, int split, const float* acc) {
  if (!valid) return;
  float4* d4 = reinterpret_cast<float4*>(e.split_part + static_cast<size_t>(split) * e.split_stride +
                                         (static_cast<size_t>(b * g.H + y) * g.W + x) * ncols + n0);
#pragma unroll
  for (int q = 0; q < CH / 4; ++q) d4[q] = make_float4(acc[4 * q], acc[4 * q + 1], acc[4 * q + 2], acc[4 * q + 3]);
}

// ---------------------------------------------------------------- EPI_GATES
// 32 columns = 8 channels x (in, remember, out, cell) raw gate pre-activations of one NormConvLSTMCell convolution:
// bias added, stored fp32, and summed per gate for the GroupNorm statistics (gs / gq live in registers across the
// column chunks of a tile: all 64 channels of a 256-column tile belong to the same GroupNorm quarter).
__device__ __forceinline__ void epi_gates(const ConvGeom& g, const EpiParams& e, int b, int y, int x, bool valid, int n0,
                                          const float* acc, float* gs, float* gq) {
  if (!valid || n0 + 32 > e.cout) return;
  float* dst = e.seg[0].dst + (static_cast<size_t>(b * g.H + y) * g.W + x) * e.seg[0].cstride + n0;
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const float4 bq = __ldg(reinterpret_cast<const float4*>(e.bias + n0) + q);
    const float4 v = make_float4(acc[4 * q] + bq.x, acc[4 * q + 1] + bq.y, acc[4 * q + 2] + bq.z, acc[4 * q + 3] + bq.w);
    reinterpret_cast<float4*>(dst)[q] = v;
    gs[0] += v.x; gs[1] += v.y; gs[2] += v.z; gs[3] += v.w;
    gq[0] += v.x * v.x; gq[1] += v.y * v.y; gq[2] += v.z * v.z; gq[3] += v.w * v.w;
  }
}
// After the last chunk: the 16 rows of a sample inside the tile are 16 consecutive lanes; their sums go to the
// sample's slot of this (row tile, column tile) -- a fixed place, so the later reduction order is fixed too.
__device__ __forceinline__ void epi_gates_finish(const EpiParams& e, int b, bool valid, int n_tile, int yb, float* gs,
                                                 float* gq) {
#pragma unroll
  for (int o = 8; o > 0; o >>= 1)
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      gs[k] += __shfl_xor_sync(0xffffffffu, gs[k], o);
      gq[k] += __shfl_xor_sync(0xffffffffu, gq[k], o);
    }
  if ((threadIdx.x & 15) == 0 && valid) {
    const int quarter = n_tile / e.gn_ntq;
    const int slot = yb * e.gn_ntq + (n_tile - quarter * e.gn_ntq);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float* p = e.gn_part + ((((static_cast<size_t>(b) * 2 + e.gn_tensor) * 16 + (k * 4 + quarter)) * 6 + slot) * 2);
      p[0] = gs[k];
      p[1] = gq[k];
    }
  }
}

// ---------------------------------------------------------------- EPI_GAUSS
// columns n0 .. n0+31 = 16 z channels x (mu, logvar); z = mu + eps * exp(0.5 * logvar)
__device__ __forceinline__ void epi_gauss(const ConvGeom& g, const EpiParams& e, int b, int y, int x, bool valid,
                                          int n0, const float* acc) {
  if (!valid || n0 >= 128) return;
  const int zc0 = n0 >> 1;
  const int hw = g.H * g.W;
  const int pos = y * g.W + x;
  float zv[16];
#pragma unroll
  for (int q4 = 0; q4 < 4; ++q4) {
    float nz[4] = {0.f, 0.f, 0.f, 0.f};
    if (!e.sample_mean && e.eps == nullptr) {
      const Philox4 r = philox4x32_10(static_cast<uint32_t>(e.cand_offset + b),
                                      static_cast<uint32_t>(pos * 16 + (zc0 >> 2) + q4), e.noise_ctr, 0x5ac5u,
                                      static_cast<uint32_t>(e.seed), static_cast<uint32_t>(e.seed >> 32));
      box_muller(r.v[0], r.v[1], nz[0], nz[1]);
      box_muller(r.v[2], r.v[3], nz[2], nz[3]);
    }
#pragma unroll
    for (int qq = 0; qq < 4; ++qq) {
      const int q = q4 * 4 + qq;
      const int zc = zc0 + q;
      float z = 0.f;
      if (zc < e.z_dim) {
        const float mu = acc[2 * q] + __ldg(e.bias + n0 + 2 * q);
        const float lv = acc[2 * q + 1] + __ldg(e.bias + n0 + 2 * q + 1);
        const size_t nchw = (static_cast<size_t>(b) * e.z_dim + zc) * hw + pos;
        if (e.mu_out) e.mu_out[nchw] = mu;
        if (e.logvar_out) e.logvar_out[nchw] = lv;
        if (e.sample_mean) {
          z = mu;
        } else {
          const float ep = e.eps ? __ldg(e.eps + nchw) : nz[qq];
          z = ep * expf(0.5f * lv) + mu;
        }
      }
      zv[q] = z;
    }
  }
  __nv_bfloat16* dst = e.z_out + (static_cast<size_t>(b) * hw + pos) * 64 + zc0;
  reinterpret_cast<uint4*>(dst)[0] = make_uint4(pack_bf16x2(zv[0], zv[1]), pack_bf16x2(zv[2], zv[3]),
                                                pack_bf16x2(zv[4], zv[5]), pack_bf16x2(zv[6], zv[7]));
  reinterpret_cast<uint4*>(dst)[1] = make_uint4(pack_bf16x2(zv[8], zv[9]), pack_bf16x2(zv[10], zv[11]),
                                                pack_bf16x2(zv[12], zv[13]), pack_bf16x2(zv[14], zv[15]));
}

// ---------------------------------------------------------------- EPI_FRAME
// columns 0..3 = (r, g, b, compositing mask) pre-sigmoid. All 32 lanes of a warp belong to one candidate (NB == 1).
// Must be called by every lane of the warp (shuffle reduction).
__device__ __forceinline__ void epi_frame(const ConvGeom& g, const EpiParams& e, int b, int y, int x, bool valid,
                                          int part_idx, const float* acc) {
  float sq = 0.f, cnt = 0.f;
  if (valid) {
    float xp[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) xp[c] = sigmoidf_(acc[c] + __ldg(e.bias + c));
    const int pos = y * g.W + x;
    const size_t pix = static_cast<size_t>(b) * g.H * g.W + pos;
    if (e.xpred_out) {
#pragma unroll
      for (int c = 0; c < 4; ++c) e.xpred_out[(static_cast<size_t>(b) * 4 + c) * g.H * g.W + pos] = xp[c];
    }
    if (e.curr_img) {
      const float4 cur = *reinterpret_cast<const float4*>(e.curr_img + pix * 4);
      const float m = xp[3];
      float nx[3] = {(1.f - m) * cur.x + m * xp[0], (1.f - m) * cur.y + m * xp[1], (1.f - m) * cur.z + m * xp[2]};
      const bool robot = e.mask_next ? (__ldg(e.mask_next + pix) != 0.f) : false;
      if (e.zero_robot && robot) nx[0] = nx[1] = nx[2] = 0.f;
      *reinterpret_cast<float4*>(e.next_img + pix * 4) = make_float4(nx[0], nx[1], nx[2], 0.f);
      if (e.cost_part) {
        const float4 gl = __ldg(reinterpret_cast<const float4*>(e.goal_img) + pos);
        const float d0 = 255.f * (nx[0] - gl.x), d1 = 255.f * (nx[1] - gl.y), d2 = 255.f * (nx[2] - gl.z);
        sq = d0 * d0 + d1 * d1 + d2 * d2;
        if (e.dontcare) {
          const bool dc = robot || (e.goal_mask && __ldg(e.goal_mask + pos) != 0.f);
          if (dc) sq = 0.f;
          cnt = dc ? 0.f : 1.f;
        }
      }
    }
  }
  if (e.cost_part) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      sq += __shfl_xor_sync(0xffffffffu, sq, o);
      cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    }
    // all lanes of a warp belong to one candidate; lane 0 itself may be a discarded halo row (halo kernel)
    const bool any_valid = __any_sync(0xffffffffu, valid);
    if ((threadIdx.x & 31) == 0 && any_valid) {
      float* dst = e.cost_part + (static_cast<size_t>(b) * e.cost_nparts + part_idx) * 2;  // one partial per warp
      dst[0] = sq;
      dst[1] = cnt;
    }
  }
}

}  // namespace rac
