// Training-data path on device. The reference prepares every clip on CPU loader workers
// (src/dataset/robonet/robonet_dataset.py:257-300 _preprocess_images_masks: ToTensor, optional random crop + bilinear
// resize back to H x W + shuffled colour jitter :546-573, masks re-binarised), collates batch-first float tensors and
// process_batch (:434-451) transposes them to time-first and copies them to the device: 16 bytes per pixel over PCIe.
// Here the loader hands over the raw uint8 frames (3 B / pixel + the mask) and ONE launch does the rest: a CTA per
// frame stages the frame through shared memory with 16-byte loads, every thread keeps its 12 pixels in registers
// through crop / resize / colour transforms (the contrast transform needs the frame's grey mean: one block reduction
// in a fixed order), and the time-first float planes are written with coalesced stores. HBM-bound and tiny
// (13 B read + 16 B written per pixel).
//
// The pixel arithmetic is torchvision's (transforms.functional on float tensors; the reference pins 0.9.1):
//   resize     ATen upsample_bilinear2d, align_corners False: src = max(scale * (dst + 0.5) - 0.5, 0), scale = in / out
//   blend      clamp(ratio * a + (1 - ratio) * b, 0, 1); brightness b = 0, contrast b = mean(grey), saturation b = grey
//   grey       0.2989 r + 0.587 g + 0.114 b
//   hue        rgb -> hsv, h = (h + factor) mod 1, hsv -> rgb
#include "misc_kernels.cuh"

#include <math.h>
#include <stdint.h>

#include "../../include/racb200.h"

namespace rac {

namespace {

constexpr int kH = 48, kW = 64, kHW = kH * kW;
constexpr int kThreads = 256, kPer = kHW / kThreads;  // 12 pixels per thread: p = tid + 256 * k

__device__ __forceinline__ float clamp01(float v) { return fminf(fmaxf(v, 0.f), 1.f); }
__device__ __forceinline__ float grey(float r, float g, float b) { return 0.2989f * r + 0.587f * g + 0.114f * b; }

struct Taps { int i0, i1; float l1; };
// one axis of the bilinear resize of a crop window [off, off + n_in) back to n_out samples
__device__ __forceinline__ Taps axis_taps(int d, int n_in, int n_out, int off) {
  const float scale = static_cast<float>(n_in) / static_cast<float>(n_out);
  float real = scale * (static_cast<float>(d) + 0.5f) - 0.5f;
  if (real < 0.f) real = 0.f;
  int k = static_cast<int>(floorf(real));
  if (k > n_in - 1) k = n_in - 1;
  Taps t;
  t.i0 = off + k;
  t.i1 = off + k + (k < n_in - 1 ? 1 : 0);
  t.l1 = fminf(fmaxf(real - static_cast<float>(k), 0.f), 1.f);
  return t;
}

__device__ __forceinline__ void hue_shift(float& r, float& g, float& b, float factor) {
  const float maxc = fmaxf(fmaxf(r, g), b), minc = fminf(fminf(r, g), b);
  const bool eqc = maxc == minc;
  const float cr = maxc - minc;
  const float s = cr / (eqc ? 1.f : maxc);
  const float div = eqc ? 1.f : cr;
  const float rc = (maxc - r) / div, gc = (maxc - g) / div, bc = (maxc - b) / div;
  float h = 0.f;
  if (maxc == r) h = bc - gc;
  else if (maxc == g) h = 2.f + rc - bc;
  else h = 4.f + gc - rc;
  h = fmodf(h / 6.f + 1.f, 1.f);
  h = h + factor;
  h = h - floorf(h);                      // python-style (h + factor) % 1.0
  if (h >= 1.f) h = 0.f;                  // -tiny % 1.0 rounds to 1.0 in float: same sector as 0 after the % 6 below
  const float v = maxc;
  const float h6 = h * 6.f;
  const float fl = floorf(h6);
  const float f = h6 - fl;
  const int i = static_cast<int>(fl) % 6;
  const float p = clamp01(v * (1.f - s));
  const float q = clamp01(v * (1.f - s * f));
  const float t = clamp01(v * (1.f - s * (1.f - f)));
  switch (i) {
    case 0: r = v; g = t; b = p; break;
    case 1: r = q; g = v; b = p; break;
    case 2: r = p; g = v; b = t; break;
    case 3: r = p; g = q; b = v; break;
    case 4: r = t; g = p; b = v; break;
    default: r = v; g = p; b = q; break;
  }
}

__global__ void __launch_bounds__(kThreads)
process_batch_kernel(const uint8_t* __restrict__ frames, const void* __restrict__ masks, int mask_u8, int B, int T,
                     const rac_augment* __restrict__ aug, float* __restrict__ img_out, float* __restrict__ mask_out) {
  __shared__ __align__(16) uint8_t s_raw[kHW * 3];
  __shared__ __align__(16) float s_mask[kHW];
  __shared__ float s_red[kThreads / 32];
  const int frame = blockIdx.x;           // batch-first input: frame = b * T + t
  const int b = frame / T, t = frame - b * T;
  const int tid = threadIdx.x;
  const size_t out_frame = static_cast<size_t>(t) * B + b;  // time-first output

  // ---- stage the frame and its mask (16-byte loads)
  {
    const uint4* src = reinterpret_cast<const uint4*>(frames + static_cast<size_t>(frame) * kHW * 3);
    for (int i = tid; i < kHW * 3 / 16; i += kThreads) reinterpret_cast<uint4*>(s_raw)[i] = __ldg(src + i);
    if (masks) {
      if (mask_u8) {
        const uint8_t* m = static_cast<const uint8_t*>(masks) + static_cast<size_t>(frame) * kHW;
        for (int i = tid; i < kHW; i += kThreads) s_mask[i] = static_cast<float>(__ldg(m + i));
      } else {
        const float4* m = reinterpret_cast<const float4*>(static_cast<const float*>(masks) + static_cast<size_t>(frame) * kHW);
        for (int i = tid; i < kHW / 4; i += kThreads) reinterpret_cast<float4*>(s_mask)[i] = __ldg(m + i);
      }
    }
  }
  __syncthreads();

  rac_augment a{};
  a.crop_h = kH; a.crop_w = kW;
  a.order[0] = a.order[1] = a.order[2] = a.order[3] = -1;
  if (aug) a = aug[b];
  const bool resize = !(a.crop_h == kH && a.crop_w == kW);

  float r[kPer], g[kPer], bl[kPer];
#pragma unroll
  for (int k = 0; k < kPer; ++k) {
    const int p = tid + kThreads * k;
    const int y = p >> 6, x = p & 63;
    if (!resize) {
      // ToTensor: value / 255 (IEEE division, bit-exact)
      r[k] = static_cast<float>(s_raw[p * 3]) / 255.f;
      g[k] = static_cast<float>(s_raw[p * 3 + 1]) / 255.f;
      bl[k] = static_cast<float>(s_raw[p * 3 + 2]) / 255.f;
      if (masks) mask_out[out_frame * kHW + p] = s_mask[p] != 0.f ? 1.f : 0.f;
    } else {
      const Taps ty = axis_taps(y, a.crop_h, kH, a.crop_i), tx = axis_taps(x, a.crop_w, kW, a.crop_j);
      const float ly1 = ty.l1, ly0 = 1.f - ty.l1, lx1 = tx.l1, lx0 = 1.f - tx.l1;
      const int q00 = ty.i0 * kW + tx.i0, q01 = ty.i0 * kW + tx.i1, q10 = ty.i1 * kW + tx.i0, q11 = ty.i1 * kW + tx.i1;
      auto px = [&](int c) -> float {
        const float p00 = static_cast<float>(s_raw[q00 * 3 + c]) / 255.f, p01 = static_cast<float>(s_raw[q01 * 3 + c]) / 255.f;
        const float p10 = static_cast<float>(s_raw[q10 * 3 + c]) / 255.f, p11 = static_cast<float>(s_raw[q11 * 3 + c]) / 255.f;
        return ly0 * (lx0 * p00 + lx1 * p01) + ly1 * (lx0 * p10 + lx1 * p11);
      };
      r[k] = px(0); g[k] = px(1); bl[k] = px(2);
      if (masks) {
        // "cast back to 0 or 1 value" (:288-290): any non-zero interpolated value is robot
        const float mv = ly0 * (lx0 * s_mask[q00] + lx1 * s_mask[q01]) + ly1 * (lx0 * s_mask[q10] + lx1 * s_mask[q11]);
        mask_out[out_frame * kHW + p] = mv != 0.f ? 1.f : 0.f;
      }
    }
  }

  // ---- colour jitter: the four transforms in the clip's shuffled order
  for (int o = 0; o < 4; ++o) {
    const int op = a.order[o];
    if (op < 0) continue;                  // uniform over the CTA
    const float f = static_cast<float>(a.factor[op]);
    const float f1 = static_cast<float>(1.0 - a.factor[op]);
    if (op == 0) {
#pragma unroll
      for (int k = 0; k < kPer; ++k) { r[k] = clamp01(f * r[k]); g[k] = clamp01(f * g[k]); bl[k] = clamp01(f * bl[k]); }
    } else if (op == 1) {
      float part = 0.f;
#pragma unroll
      for (int k = 0; k < kPer; ++k) part += grey(r[k], g[k], bl[k]);
#pragma unroll
      for (int s = 16; s > 0; s >>= 1) part += __shfl_xor_sync(0xffffffffu, part, s);
      __syncthreads();                     // the previous contrast pass (if any) has finished reading s_red
      if ((tid & 31) == 0) s_red[tid >> 5] = part;
      __syncthreads();
      float total = 0.f;
#pragma unroll
      for (int w = 0; w < kThreads / 32; ++w) total += s_red[w];
      const float m = f1 * (total / static_cast<float>(kHW));
#pragma unroll
      for (int k = 0; k < kPer; ++k) { r[k] = clamp01(f * r[k] + m); g[k] = clamp01(f * g[k] + m); bl[k] = clamp01(f * bl[k] + m); }
    } else if (op == 2) {
#pragma unroll
      for (int k = 0; k < kPer; ++k) {
        const float m = f1 * grey(r[k], g[k], bl[k]);
        r[k] = clamp01(f * r[k] + m); g[k] = clamp01(f * g[k] + m); bl[k] = clamp01(f * bl[k] + m);
      }
    } else {
#pragma unroll
      for (int k = 0; k < kPer; ++k) hue_shift(r[k], g[k], bl[k], f);
    }
  }

  float* dst = img_out + out_frame * 3 * kHW;
#pragma unroll
  for (int k = 0; k < kPer; ++k) {
    const int p = tid + kThreads * k;
    dst[p] = r[k];
    dst[kHW + p] = g[k];
    dst[2 * kHW + p] = bl[k];
  }
}

}  // namespace

cudaError_t launch_process_batch(const uint8_t* frames, const void* masks, int mask_u8, int B, int T,
                                 const rac_augment* aug, float* img_out, float* mask_out, cudaStream_t s) {
  if (B * T < 1) return cudaSuccess;
  process_batch_kernel<<<B * T, kThreads, 0, s>>>(frames, masks, mask_u8, B, T, aug, img_out, mask_out);
  return cudaGetLastError();
}

}  // namespace rac
