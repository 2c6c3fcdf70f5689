// Small memory-bound kernels around the convolution stack: first encoder layer, pooling, tiled action channels,
// image preparation, planning-cost reductions and the training criteria (forward values).
#include "misc_kernels.cuh"
#include "epilogue.cuh"
#include "ptx.cuh"

namespace rac {

// ---------------------------------------------------------------- encoder.c1.0
// w: [9 * cin][64] fp32 (tap-major, BN folded), bias [64]. K = 27..45 is too small for a tensor tile, so this layer is
// CUDA-core work next to an HBM-write stream (393 KB bf16 out per candidate). Register tile: one thread = 4
// consecutive pixels x 8 output channels, so every weight float4 read from shared memory feeds 16 FMAs and every
// input value 8. The 8 threads of a pixel group cover the 64 channels: their weight reads are one conflict-free
// 128-byte row (layout [tap][c][half][group][4]) and together they write each pixel's 128-byte output row; input
// pixels are broadcast loads shared through L1. Persistent over pixel groups (weights staged once per CTA).
template <int CIN>
__global__ void __launch_bounds__(256, 2)
first_conv_kernel(const float* __restrict__ img4, const float* __restrict__ mask_a, const float* __restrict__ mask_b,
                  long long mask_bstride, const float* __restrict__ w, const float* __restrict__ bias,
                  __nv_bfloat16* __restrict__ out, float* __restrict__ raw_out, int B, int H, int W) {
  pdl_entry();
  __shared__ __align__(16) float sw[9 * CIN * 64];
  __shared__ float sb[64];
  for (int i = threadIdx.x; i < 9 * CIN * 64; i += blockDim.x) {
    const int o = i & 63, tc = i >> 6;  // source [tap*CIN + c][o]; o = group * 8 + half * 4 + j
    sw[(tc * 2 + ((o >> 2) & 1)) * 32 + (o >> 3) * 4 + (o & 3)] = w[i];
  }
  if (threadIdx.x < 64) sb[threadIdx.x] = bias[threadIdx.x];
  __syncthreads();
  const int grp = threadIdx.x & 7;           // output channels grp*8 .. grp*8+7
  const int gpr = W >> 2;                    // pixel groups per image row
  const int total = B * H * gpr;             // pixel groups
  for (int pg = blockIdx.x * 32 + (threadIdx.x >> 3); pg < total; pg += gridDim.x * 32) {
    const int xg = pg % gpr;
    const int row = pg / gpr;                // b * H + y
    const int y = row % H;
    const int b = row / H;
    const int x0 = xg * 4;
    float acc[4][8];
#pragma unroll
    for (int p = 0; p < 4; ++p)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[p][j] = sb[grp * 8 + j];
#pragma unroll
    for (int kh = 0; kh < 3; ++kh) {
      const int yy = y + kh - 1;
      if (yy < 0 || yy >= H) continue;
      // input row segment x0-1 .. x0+4 (6 pixels), zero outside the image
      float in[6][CIN];
      const size_t rbase = (static_cast<size_t>(b) * H + yy) * W;
      const size_t mbase = static_cast<size_t>(b) * mask_bstride + static_cast<size_t>(yy) * W;
#pragma unroll
      for (int q = 0; q < 6; ++q) {
        const int xx = x0 + q - 1;
        const bool ok = xx >= 0 && xx < W;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (ok) v = __ldg(reinterpret_cast<const float4*>(img4) + rbase + xx);
        in[q][0] = v.x; in[q][1] = v.y; in[q][2] = v.z;
        if constexpr (CIN > 3) in[q][3] = ok ? __ldg(mask_a + mbase + xx) : 0.f;
        if constexpr (CIN > 4) in[q][4] = ok ? __ldg(mask_b + mbase + xx) : 0.f;
      }
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) {
#pragma unroll
        for (int c = 0; c < CIN; ++c) {
          const float* wt = &sw[(((kh * 3 + kw) * CIN + c) * 2) * 32 + grp * 4];
          const float4 w0 = *reinterpret_cast<const float4*>(wt);
          const float4 w1 = *reinterpret_cast<const float4*>(wt + 32);
#pragma unroll
          for (int p = 0; p < 4; ++p) {
            const float a = in[p + kw][c];
            acc[p][0] = fmaf(a, w0.x, acc[p][0]); acc[p][1] = fmaf(a, w0.y, acc[p][1]);
            acc[p][2] = fmaf(a, w0.z, acc[p][2]); acc[p][3] = fmaf(a, w0.w, acc[p][3]);
            acc[p][4] = fmaf(a, w1.x, acc[p][4]); acc[p][5] = fmaf(a, w1.y, acc[p][5]);
            acc[p][6] = fmaf(a, w1.z, acc[p][6]); acc[p][7] = fmaf(a, w1.w, acc[p][7]);
          }
        }
      }
    }
    const size_t pix0 = (static_cast<size_t>(b) * H + y) * W + x0;
    if (raw_out) {  // training: pre-BatchNorm fp32 output, activation applied later with batch statistics
#pragma unroll
      for (int p = 0; p < 4; ++p) {
        float* dst = raw_out + (pix0 + p) * 64 + grp * 8;
        *reinterpret_cast<float4*>(dst) = make_float4(acc[p][0], acc[p][1], acc[p][2], acc[p][3]);
        *reinterpret_cast<float4*>(dst + 4) = make_float4(acc[p][4], acc[p][5], acc[p][6], acc[p][7]);
      }
      continue;
    }
#pragma unroll
    for (int p = 0; p < 4; ++p) {
      float v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = acc[p][j] > 0.f ? acc[p][j] : 0.2f * acc[p][j];
      *reinterpret_cast<uint4*>(out + (pix0 + p) * 64 + grp * 8) =
          make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
    }
  }
}

cudaError_t launch_first_conv(const float* img4, const float* mask_a, const float* mask_b, long long mask_bstride,
                              const float* w, const float* bias, __nv_bfloat16* out, int B, int H, int W, int cin,
                              cudaStream_t s, float* raw_out) {
  if (cin < 3 || cin > 5 || W % 4 != 0) return cudaErrorInvalidValue;
  if ((cin > 3 && !mask_a) || (cin > 4 && !mask_b)) return cudaErrorInvalidValue;
  const long long groups = (static_cast<long long>(B) * H * (W / 4) + 31) / 32;  // 32 pixel groups per CTA iteration
  if (static_cast<long long>(B) * H * (W / 4) >= (1ll << 31)) return cudaErrorInvalidValue;
  const long long cap = 148 * 2;
  const unsigned grid = static_cast<unsigned>(groups < cap ? groups : cap);
  cudaError_t err;
  if (cin == 3)
    err = launch_pdl_small(first_conv_kernel<3>, dim3(grid), dim3(256), 0, s, img4, mask_a, mask_b, mask_bstride, w, bias, out, raw_out, B, H, W);
  else if (cin == 4)
    err = launch_pdl_small(first_conv_kernel<4>, dim3(grid), dim3(256), 0, s, img4, mask_a, mask_b, mask_bstride, w, bias, out, raw_out, B, H, W);
  else
    err = launch_pdl_small(first_conv_kernel<5>, dim3(grid), dim3(256), 0, s, img4, mask_a, mask_b, mask_bstride, w, bias, out, raw_out, B, H, W);
  if (err != cudaSuccess) return err;
  return cudaGetLastError();
}

// ---------------------------------------------------------------- 2x2 max pool
__device__ __forceinline__ uint4 bf16x8_max(uint4 a, uint4 b) {
  uint4 r;
  const __nv_bfloat162* pa = reinterpret_cast<const __nv_bfloat162*>(&a);
  const __nv_bfloat162* pb = reinterpret_cast<const __nv_bfloat162*>(&b);
  __nv_bfloat162* pr = reinterpret_cast<__nv_bfloat162*>(&r);
#pragma unroll
  for (int i = 0; i < 4; ++i) pr[i] = __hmax2(pa[i], pb[i]);
  return r;
}
__global__ void __launch_bounds__(256)
maxpool2_kernel(const __nv_bfloat16* __restrict__ in, int in_cstride, int in_coff, __nv_bfloat16* __restrict__ out,
                int B, int H, int W, int C) {
  pdl_entry();
  const int Ho = H / 2, Wo = W / 2, C8 = C / 8;
  const size_t total = static_cast<size_t>(B) * Ho * Wo * C8;
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int c8 = static_cast<int>(i % C8);
  const int xo = static_cast<int>((i / C8) % Wo);
  const int yo = static_cast<int>((i / (static_cast<size_t>(C8) * Wo)) % Ho);
  const size_t b = i / (static_cast<size_t>(C8) * Wo * Ho);
  const __nv_bfloat16* p = in + ((b * H + 2 * yo) * W + 2 * xo) * in_cstride + in_coff + c8 * 8;
  const uint4 v00 = *reinterpret_cast<const uint4*>(p);
  const uint4 v01 = *reinterpret_cast<const uint4*>(p + in_cstride);
  const uint4 v10 = *reinterpret_cast<const uint4*>(p + static_cast<size_t>(W) * in_cstride);
  const uint4 v11 = *reinterpret_cast<const uint4*>(p + static_cast<size_t>(W) * in_cstride + in_cstride);
  *reinterpret_cast<uint4*>(out + ((b * Ho + yo) * Wo + xo) * C + c8 * 8) =
      bf16x8_max(bf16x8_max(v00, v01), bf16x8_max(v10, v11));
}
cudaError_t launch_maxpool2(const __nv_bfloat16* in, int in_cstride, int in_coff, __nv_bfloat16* out, int B, int H,
                            int W, int C, cudaStream_t s) {
  const size_t total = static_cast<size_t>(B) * (H / 2) * (W / 2) * (C / 8);
  if (cudaError_t e_ = launch_pdl_small(maxpool2_kernel, dim3(static_cast<unsigned>((total + 255) / 256)), dim3(256), 0, s, in, in_cstride, in_coff, out, B, H, W, C); e_ != cudaSuccess) return e_;
  return cudaGetLastError();
}

// ---------------------------------------------------------------- tiled action / state channels
__global__ void __launch_bounds__(256)
aux_tile_kernel(const float* __restrict__ action, int astride, int adim, const float* __restrict__ r,
                const float* __restrict__ r2, int rdim, __nv_bfloat16* __restrict__ aux, int B, int HW) {
  pdl_entry();
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;  // (b, pos, c8)
  const size_t total = static_cast<size_t>(B) * HW * 8;
  if (i >= total) return;
  const int c8 = static_cast<int>(i & 7);
  const size_t b = i / (static_cast<size_t>(HW) * 8);
  float v[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int c = c8 * 8 + j;
    float x = 0.f;
    if (action && c < adim) x = action[b * astride + c];
    else if (r && c < adim + rdim) x = r[b * rdim + (c - adim)];
    else if (r && r2 && c < adim + 2 * rdim) x = r2[b * rdim + (c - adim - rdim)];
    v[j] = x;
  }
  *reinterpret_cast<uint4*>(aux + i * 8) =
      make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
}
cudaError_t launch_aux_tile(const float* action, int astride, int adim, const float* r, const float* r2, int rdim,
                            __nv_bfloat16* aux, int B, int HW, cudaStream_t s) {
  if (adim + (r ? rdim : 0) + (r2 ? rdim : 0) > 64) return cudaErrorInvalidValue;
  const size_t total = static_cast<size_t>(B) * HW * 8;
  if (cudaError_t e_ = launch_pdl_small(aux_tile_kernel, dim3(static_cast<unsigned>((total + 255) / 256)), dim3(256), 0, s, action, astride, adim, r, r2, rdim, aux, B,
                                                                             HW); e_ != cudaSuccess) return e_;
  return cudaGetLastError();
}

// ---------------------------------------------------------------- image preparation
__global__ void __launch_bounds__(256)
img_prep_u8_kernel(const uint8_t* __restrict__ img, const float* __restrict__ mask0, int zero_robot,
                   float* __restrict__ img4, int B, int HW) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= static_cast<size_t>(B) * HW) return;
  const int pos = static_cast<int>(i % HW);
  float r = static_cast<float>(img[pos * 3 + 0]) / 255.f;
  float g = static_cast<float>(img[pos * 3 + 1]) / 255.f;
  float b = static_cast<float>(img[pos * 3 + 2]) / 255.f;
  if (zero_robot && mask0 && mask0[i] != 0.f) r = g = b = 0.f;
  *reinterpret_cast<float4*>(img4 + i * 4) = make_float4(r, g, b, 0.f);
}
cudaError_t launch_img_prep_u8(const uint8_t* img_hwc, const float* mask0, int zero_robot, float* img4, int B, int H,
                               int W, cudaStream_t s) {
  const size_t total = static_cast<size_t>(B) * H * W;
  img_prep_u8_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, s>>>(img_hwc, mask0, zero_robot, img4, B,
                                                                                H * W);
  return cudaGetLastError();
}
__global__ void __launch_bounds__(256)
img_prep_nchw_kernel(const float* __restrict__ img, float* __restrict__ img4, int B, int HW) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= static_cast<size_t>(B) * HW) return;
  const size_t b = i / HW;
  const int pos = static_cast<int>(i % HW);
  const float* p = img + b * 3 * HW + pos;
  *reinterpret_cast<float4*>(img4 + i * 4) = make_float4(p[0], p[HW], p[2 * HW], 0.f);
}
cudaError_t launch_img_prep_nchw(const float* img_nchw, float* img4, int B, int H, int W, cudaStream_t s) {
  const size_t total = static_cast<size_t>(B) * H * W;
  img_prep_nchw_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, s>>>(img_nchw, img4, B, H * W);
  return cudaGetLastError();
}
__global__ void __launch_bounds__(256)
goal_prep_kernel(const uint8_t* __restrict__ g, float* __restrict__ g4, int total) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  *reinterpret_cast<float4*>(g4 + static_cast<size_t>(i) * 4) =
      make_float4(static_cast<float>(g[i * 3 + 0]) / 255.f, static_cast<float>(g[i * 3 + 1]) / 255.f,
                  static_cast<float>(g[i * 3 + 2]) / 255.f, 0.f);
}
cudaError_t launch_goal_prep(const uint8_t* goal_hwc, float* goal4, int G, int H, int W, cudaStream_t s) {
  const int total = G * H * W;
  goal_prep_kernel<<<(total + 255) / 256, 256, 0, s>>>(goal_hwc, goal4, total);
  return cudaGetLastError();
}

// ---------------------------------------------------------------- cost finish
__global__ void __launch_bounds__(128)
cost_finish_kernel(const float* __restrict__ part, int nparts, int dontcare, float weight, int accumulate,
                   double* __restrict__ sum_cost, float* __restrict__ step_cost, int B, double* const* peers,
                   int peer_world, long long peer_offset) {
  pdl_entry();
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const float* p = part + static_cast<size_t>(b) * nparts * 2;
  float sq = 0.f, cnt = 0.f;
  for (int i = 0; i < nparts; ++i) {  // fixed order: deterministic
    sq += p[2 * i];
    cnt += p[2 * i + 1];
  }
  float dist = sqrtf(sq);
  if (dontcare) dist = dist / cnt;  // reference divides without +1 (losses.py:259-260)
  const float cost = weight * (-dist);
  if (step_cost) step_cost[b] = accumulate ? cost : 0.f;  // unused steps report rew = 0 (trajectory_sampler.py:165-173)
  if (accumulate) sum_cost[b] += static_cast<double>(cost);
  if (peers) {
    // last rollout step on a sharded plan: the finished cost goes straight into every rank's gathered cost vector
    // (stores over NVLink peer memory; the all-gather of cem.py:96 without a collective call)
    const double v = sum_cost[b];
    for (int p = 0; p < peer_world; ++p) peers[p][peer_offset + b] = v;
    __threadfence_system();
  }
}

// Cross-GPU flag barrier: lane p publishes `seq` in slot [rank] of rank p's pad, then waits for slot [p] of its own pad.
__global__ void __launch_bounds__(32) peer_barrier_kernel(uint32_t* const* pads, int base, int rank, int world,
                                                          uint32_t seq) {
  const int p = threadIdx.x;
  if (p >= world) return;
  __threadfence_system();  // the peer stores of the preceding kernels on this stream are ordered before the flag
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(pads[p] + base + rank), "r"(seq) : "memory");
  const uint32_t* mine = pads[rank] + base + p;
  const long long t0 = clock64();
  for (;;) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(mine) : "memory");
    if (static_cast<int32_t>(v - seq) >= 0) break;
    // a slow rank (first-use allocations, host stalls) may legitimately lag by seconds; a missing one must not hang
    // the GPU for ever: give up after ~60 s
    if (clock64() - t0 > 100000000000ll) __trap();
  }
}
cudaError_t launch_peer_barrier(uint32_t* const* pads, int base, int rank, int world, uint32_t seq, cudaStream_t s) {
  if (world < 1 || world > 32 || rank < 0 || rank >= world || base < 0) return cudaErrorInvalidValue;
  peer_barrier_kernel<<<1, 32, 0, s>>>(pads, base, rank, world, seq);
  return cudaGetLastError();
}
cudaError_t launch_cost_finish(const float* cost_part, int nparts, int dontcare, float weight, int accumulate,
                               double* sum_cost, float* step_cost, int B, cudaStream_t s, double* const* peers,
                               int peer_world, long long peer_offset) {
  if (cudaError_t e_ = launch_pdl_small(cost_finish_kernel, dim3((B + 127) / 128), dim3(128), 0, s, cost_part, nparts, dontcare, weight, accumulate, sum_cost,
                                                     step_cost, B, peers, peer_world, peer_offset); e_ != cudaSuccess) return e_;
  return cudaGetLastError();
}

// ---------------------------------------------------------------- block reduction helper
template <typename T>
__device__ __forceinline__ T block_sum(T v, T* sh) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = (blockDim.x + 31) >> 5;
  __syncthreads();
  if (l == 0) sh[w] = v;
  __syncthreads();
  T t = (threadIdx.x < nw) ? sh[threadIdx.x] : T(0);
  if (w == 0) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    if (l == 0) sh[0] = t;
  }
  __syncthreads();
  return sh[0];
}

// ---------------------------------------------------------------- stand-alone planning cost (reference layout)
// One CTA per candidate: reads 3 planes of `curr` (+ mask) once with 128-bit loads; goal image/mask stay L2-resident.
__global__ void __launch_bounds__(256)
masked_cost_kernel(const float* __restrict__ curr, const float* __restrict__ goal, const float* __restrict__ cmask,
                   const float* __restrict__ gmask, int dontcare, float* __restrict__ out, int HW) {
  __shared__ float sh[32];
  const size_t b = blockIdx.x;
  const int q4 = HW / 4;
  const float4* c0 = reinterpret_cast<const float4*>(curr + b * 3 * HW);
  const float4* g0 = reinterpret_cast<const float4*>(goal);
  const float4* cm = cmask ? reinterpret_cast<const float4*>(cmask + b * HW) : nullptr;
  const float4* gm = gmask ? reinterpret_cast<const float4*>(gmask) : nullptr;
  float sq = 0.f, cnt = 0.f;
  for (int i = threadIdx.x; i < q4; i += blockDim.x) {
    float4 a[3], g[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      a[c] = __ldcs(c0 + c * q4 + i);
      g[c] = __ldg(g0 + c * q4 + i);
    }
    float keep[4] = {1.f, 1.f, 1.f, 1.f};
    if (dontcare) {
      float4 m1 = make_float4(0.f, 0.f, 0.f, 0.f), m2 = m1;
      if (cm) m1 = __ldcs(cm + i);
      if (gm) m2 = __ldg(gm + i);
      keep[0] = (m1.x != 0.f || m2.x != 0.f) ? 0.f : 1.f;
      keep[1] = (m1.y != 0.f || m2.y != 0.f) ? 0.f : 1.f;
      keep[2] = (m1.z != 0.f || m2.z != 0.f) ? 0.f : 1.f;
      keep[3] = (m1.w != 0.f || m2.w != 0.f) ? 0.f : 1.f;
      cnt += keep[0] + keep[1] + keep[2] + keep[3];
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float d0 = 255.f * (a[c].x - g[c].x), d1 = 255.f * (a[c].y - g[c].y);
      const float d2 = 255.f * (a[c].z - g[c].z), d3 = 255.f * (a[c].w - g[c].w);
      sq += keep[0] * d0 * d0 + keep[1] * d1 * d1 + keep[2] * d2 * d2 + keep[3] * d3 * d3;
    }
  }
  sq = block_sum(sq, sh);
  if (dontcare) cnt = block_sum(cnt, sh);
  if (threadIdx.x == 0) {
    float d = sqrtf(sq);
    if (dontcare) d = d / cnt;
    out[b] = -d;
  }
}
cudaError_t launch_masked_cost(const float* curr, const float* goal, const float* curr_mask, const float* goal_mask,
                               int dontcare, float* out, int B, int HW, cudaStream_t s) {
  if (HW % 4 != 0) return cudaErrorInvalidValue;
  if (B == 0) return cudaSuccess;
  masked_cost_kernel<<<B, 256, 0, s>>>(curr, goal, curr_mask, goal_mask, dontcare, out, HW);
  return cudaGetLastError();
}

// ---------------------------------------------------------------- training criteria (forward values)
__global__ void __launch_bounds__(1024) l1_loss_kernel(const float* __restrict__ p, const float* __restrict__ t,
                                                       float* __restrict__ out, long long n) {
  __shared__ double sh[32];
  double acc = 0.0;
  for (long long i = threadIdx.x; i < n; i += blockDim.x) acc += static_cast<double>(fabsf(t[i] - p[i]));
  acc = block_sum(acc, sh);
  if (threadIdx.x == 0) out[0] = static_cast<float>(acc / static_cast<double>(n));
}
cudaError_t launch_l1_loss(const float* pred, const float* target, float* out, int64_t n, cudaStream_t s) {
  l1_loss_kernel<<<1, 1024, 0, s>>>(pred, target, out, n);
  return cudaGetLastError();
}
__global__ void __launch_bounds__(1024)
dontcare_l1_kernel(const float* __restrict__ p, const float* __restrict__ t, const float* __restrict__ mask,
                   float robot_weight, float* __restrict__ out, int B, int HW) {
  __shared__ double sh[32];
  double total = 0.0;
  for (int b = 0; b < B; ++b) {
    double acc = 0.0, world = 0.0;
    for (int i = threadIdx.x; i < HW; i += blockDim.x) {
      const bool rb = mask[static_cast<size_t>(b) * HW + i] != 0.f;
      float sd = 0.f;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const size_t q = (static_cast<size_t>(b) * 3 + c) * HW + i;
        float d = t[q] - p[q];
        if (rb) d *= robot_weight;
        sd += fabsf(d);
      }
      acc += sd;
      world += rb ? 0.0 : 3.0;
    }
    acc = block_sum(acc, sh);
    world = block_sum(world, sh);
    total += acc / (world + 1.0);
  }
  if (threadIdx.x == 0) out[0] = static_cast<float>(total / B);
}
cudaError_t launch_dontcare_l1_loss(const float* pred, const float* target, const float* mask, float robot_weight,
                                    float* out, int B, int HW, cudaStream_t s) {
  dontcare_l1_kernel<<<1, 1024, 0, s>>>(pred, target, mask, robot_weight, out, B, HW);
  return cudaGetLastError();
}
// The four reconstruction criteria of the reference trainer (trainer.py:149-161; losses.py:11-50), one CTA per sample:
// kind 0 l1 (mean |d|, optional batch weight), 1 dontcare_l1 (robot pixels x robot_weight, / (#world elements + 1)),
// 2 mse (nn.MSELoss), 3 dontcare_mse. per_sample[b] = that sample's share of the batch mean; recon_loss_sum_kernel
// adds them in index order (deterministic).
__global__ void __launch_bounds__(512)
recon_loss_kernel(const float* __restrict__ p, const float* __restrict__ t, const float* __restrict__ mask,
                  const float* __restrict__ bw, int kind, float robot_weight, float* __restrict__ per_sample, int B,
                  int HW) {
  __shared__ double sh[32];
  const int b = blockIdx.x;
  const bool dontcare = (kind & 1) != 0, squared = kind >= 2;
  double acc = 0.0, world = 0.0;
  for (int i = threadIdx.x; i < HW; i += blockDim.x) {
    const bool rb = dontcare && mask[static_cast<size_t>(b) * HW + i] != 0.f;
    float sd = 0.f;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const size_t q = (static_cast<size_t>(b) * 3 + c) * HW + i;
      float d = t[q] - p[q];
      if (rb) d *= robot_weight;
      sd += squared ? d * d : fabsf(d);
    }
    acc += sd;
    world += rb ? 0.0 : 3.0;
  }
  acc = block_sum(acc, sh);
  world = block_sum(world, sh);
  if (threadIdx.x == 0) {
    const double denom = dontcare ? world + 1.0 : 3.0 * HW;
    const double w = (bw && !squared) ? static_cast<double>(bw[b]) : 1.0;
    per_sample[b] = static_cast<float>(w * acc / denom / B);
  }
}
__global__ void recon_loss_sum_kernel(const float* __restrict__ per_sample, int B, float* __restrict__ out) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    double s = 0.0;
    for (int b = 0; b < B; ++b) s += static_cast<double>(per_sample[b]);
    out[0] = static_cast<float>(s);
  }
}
cudaError_t launch_recon_loss(const float* pred, const float* target, const float* mask, const float* batch_weight,
                              int kind, float robot_weight, float* per_sample, float* out, int B, int HW,
                              cudaStream_t s) {
  recon_loss_kernel<<<B, 512, 0, s>>>(pred, target, mask, batch_weight, kind, robot_weight, per_sample, B, HW);
  recon_loss_sum_kernel<<<1, 32, 0, s>>>(per_sample, B, out);
  return cudaGetLastError();
}

// robot_mse_criterion / world_mse_criterion (losses.py:52-78): per sample sum(diff^2) over robot (world) pixels of
// all 3 channels / (#those elements + 1), mean over the batch. out[0] += robot, out[1] += world.
__global__ void __launch_bounds__(1024)
robot_world_mse_kernel(const float* __restrict__ p, const float* __restrict__ t, const float* __restrict__ mask,
                       float* __restrict__ out, int B, int HW) {
  __shared__ double sh[32];
  double tot_r = 0.0, tot_w = 0.0;
  for (int b = 0; b < B; ++b) {
    double sr = 0.0, sw = 0.0, cr = 0.0;
    for (int i = threadIdx.x; i < HW; i += blockDim.x) {
      const bool rb = mask[static_cast<size_t>(b) * HW + i] != 0.f;
      float sd = 0.f;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const size_t q = (static_cast<size_t>(b) * 3 + c) * HW + i;
        const float d = t[q] - p[q];
        sd += d * d;
      }
      if (rb) { sr += sd; cr += 3.0; } else { sw += sd; }
    }
    sr = block_sum(sr, sh);
    sw = block_sum(sw, sh);
    cr = block_sum(cr, sh);
    tot_r += sr / (cr + 1.0);
    tot_w += sw / (3.0 * HW - cr + 1.0);
  }
  if (threadIdx.x == 0) {
    out[0] += static_cast<float>(tot_r / B);
    out[1] += static_cast<float>(tot_w / B);
  }
}
cudaError_t launch_robot_world_mse(const float* pred, const float* target, const float* mask, float* out2, int B,
                                   int HW, cudaStream_t s) {
  robot_world_mse_kernel<<<1, 1024, 0, s>>>(pred, target, mask, out2, B, HW);
  return cudaGetLastError();
}

__global__ void __launch_bounds__(1024)
kl_loss_kernel(const float* __restrict__ mu1, const float* __restrict__ lv1, const float* __restrict__ mu2,
               const float* __restrict__ lv2, float* __restrict__ out, long long n, int bs) {
  __shared__ double sh[32];
  double acc = 0.0;
  for (long long i = threadIdx.x; i < n; i += blockDim.x) {
    const float s1 = expf(0.5f * lv1[i]), s2 = expf(0.5f * lv2[i]);
    const float dm = mu1[i] - mu2[i];
    const float k = logf(s2 / s1) + (expf(lv1[i]) + dm * dm) / (2.f * expf(lv2[i])) - 0.5f;
    acc += static_cast<double>(k);
  }
  acc = block_sum(acc, sh);
  if (threadIdx.x == 0) out[0] = static_cast<float>(acc / bs);
}
cudaError_t launch_kl_loss(const float* mu1, const float* lv1, const float* mu2, const float* lv2, float* out,
                           int64_t n, int bs, cudaStream_t s) {
  kl_loss_kernel<<<1, 1024, 0, s>>>(mu1, lv1, mu2, lv2, out, n, bs);
  return cudaGetLastError();
}

// Copies channels [coff, coff + C) of every pixel of candidate 0 to candidates first .. B-1 of an NHWC bf16 tensor
// [B][HW][cstride] (encoder outputs at the first rollout step: identical for all candidates, rac_api.cu::run_step).
// 16-byte accesses; the source (<= 393 KB) stays in L2.
__global__ void __launch_bounds__(256)
broadcast_candidate_kernel(__nv_bfloat16* __restrict__ buf, int HW, int cstride, int coff, int C8, int first, int B) {
  pdl_entry();
  const long long per = static_cast<long long>(HW) * C8;
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= per * (B - first)) return;
  const long long b = first + i / per;
  const long long r = i - (b - first) * per;
  const long long p = r / C8;
  const int c = static_cast<int>(r - p * C8) * 8;
  const uint4 v = *reinterpret_cast<const uint4*>(buf + p * cstride + coff + c);
  *reinterpret_cast<uint4*>(buf + (b * HW + p) * cstride + coff + c) = v;
}
cudaError_t launch_broadcast_candidate(__nv_bfloat16* buf, int HW, int cstride, int coff, int C, int first, int B,
                                       cudaStream_t s) {
  if (B <= first) return cudaSuccess;
  if (C % 8 || cstride % 8 || coff % 8) return cudaErrorInvalidValue;
  const long long total = static_cast<long long>(HW) * (C / 8) * (B - first);
  if (cudaError_t e_ = launch_pdl_small(broadcast_candidate_kernel, dim3(static_cast<unsigned>((total + 255) / 256)), dim3(256), 0, s, buf, HW, cstride, coff, C / 8, first, B); e_ != cudaSuccess) return e_;
  return cudaGetLastError();
}

}  // namespace rac
