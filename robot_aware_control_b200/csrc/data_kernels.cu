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

// kStored: the clip is stored at another size (Hs x Ws, e.g. RoboNet's 240 x 320): the dataset's
// `tf.Compose([tf.ToTensor(), tf.Resize((h, w))])` (robonet_dataset.py:58,279-280,294-296) comes first. With the
// torchvision the reference pins (0.8.1 / 0.9.1) Resize of a float tensor is torch.nn.functional.interpolate(mode=
// "bilinear", align_corners=False) WITHOUT antialiasing: every 48 x 64 sample blends the 2 x 2 stored pixels around its
// centre. The CTA builds that float frame (and mask) in shared memory straight from the stored bytes; everything after
// it (crop / resize back, colour jitter) then reads floats instead of bytes.
template <bool kStored>
__global__ void __launch_bounds__(kThreads)
process_batch_kernel(const uint8_t* __restrict__ frames, const void* __restrict__ masks, int mask_u8, int B, int T,
                     int Hs, int Ws, const rac_augment* __restrict__ aug, float* __restrict__ img_out,
                     float* __restrict__ mask_out) {
  extern __shared__ __align__(16) uint8_t s_dyn[];
  // kStored: float planes [3][kHW] then the mask; else the raw bytes [kHW * 3] (padded to 16 B) then the mask
  uint8_t* s_raw = s_dyn;
  float* s_img = reinterpret_cast<float*>(s_dyn);
  float* s_mask = reinterpret_cast<float*>(s_dyn + (kStored ? kHW * 3 * 4 : kHW * 3));
  __shared__ float s_red[kThreads / 32];
  const int frame = blockIdx.x;           // batch-first input: frame = b * T + t
  const int b = frame / T, t = frame - b * T;
  const int tid = threadIdx.x;
  const size_t out_frame = static_cast<size_t>(t) * B + b;  // time-first output

  if constexpr (kStored) {
    const uint8_t* src = frames + static_cast<size_t>(frame) * Hs * Ws * 3;
    for (int p = tid; p < kHW; p += kThreads) {
      const int y = p >> 6, x = p & 63;
      const Taps ty = axis_taps(y, Hs, kH, 0), tx = axis_taps(x, Ws, kW, 0);
      const float ly1 = ty.l1, ly0 = 1.f - ty.l1, lx1 = tx.l1, lx0 = 1.f - tx.l1;
      const int q00 = ty.i0 * Ws + tx.i0, q01 = ty.i0 * Ws + tx.i1, q10 = ty.i1 * Ws + tx.i0, q11 = ty.i1 * Ws + tx.i1;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float p00 = static_cast<float>(__ldg(src + q00 * 3 + c)) / 255.f, p01 = static_cast<float>(__ldg(src + q01 * 3 + c)) / 255.f;
        const float p10 = static_cast<float>(__ldg(src + q10 * 3 + c)) / 255.f, p11 = static_cast<float>(__ldg(src + q11 * 3 + c)) / 255.f;
        s_img[c * kHW + p] = ly0 * (lx0 * p00 + lx1 * p01) + ly1 * (lx0 * p10 + lx1 * p11);
      }
      if (masks) {
        float m00, m01, m10, m11;
        const size_t mo = static_cast<size_t>(frame) * Hs * Ws;
        if (mask_u8) {
          const uint8_t* m = static_cast<const uint8_t*>(masks) + mo;
          m00 = m[q00]; m01 = m[q01]; m10 = m[q10]; m11 = m[q11];
        } else {
          const float* m = static_cast<const float*>(masks) + mo;
          m00 = m[q00]; m01 = m[q01]; m10 = m[q10]; m11 = m[q11];
        }
        s_mask[p] = ly0 * (lx0 * m00 + lx1 * m01) + ly1 * (lx0 * m10 + lx1 * m11);
      }
    }
  } else {
  // ---- stage the frame and its mask (16-byte loads)
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
  // one stored sample as the float ToTensor (+ Resize) produced it
  auto sample = [&](int q, int c) -> float {
    if constexpr (kStored) return s_img[c * kHW + q];
    else return static_cast<float>(s_raw[q * 3 + c]) / 255.f;  // ToTensor: value / 255 (IEEE division, bit-exact)
  };

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
      r[k] = sample(p, 0);
      g[k] = sample(p, 1);
      bl[k] = sample(p, 2);
      if (masks) mask_out[out_frame * kHW + p] = s_mask[p] != 0.f ? 1.f : 0.f;
    } else {
      const Taps ty = axis_taps(y, a.crop_h, kH, a.crop_i), tx = axis_taps(x, a.crop_w, kW, a.crop_j);
      const float ly1 = ty.l1, ly0 = 1.f - ty.l1, lx1 = tx.l1, lx0 = 1.f - tx.l1;
      const int q00 = ty.i0 * kW + tx.i0, q01 = ty.i0 * kW + tx.i1, q10 = ty.i1 * kW + tx.i0, q11 = ty.i1 * kW + tx.i1;
      auto px = [&](int c) -> float {
        const float p00 = sample(q00, c), p01 = sample(q01, c), p10 = sample(q10, c), p11 = sample(q11, c);
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

// RoboNetDataset._preprocess_states (robonet_dataset.py:302-334) + the autograsp action column of _load_actions
// (:173-194) + process_batch's time-first layout (:434-451), one thread per (clip, frame). The reference runs this in
// numpy per clip: float64 wherever a float64 operand takes part (file bounds, the camera matrices), the result is
// stored back into the float32 state array. Here: double throughout, one rounding to float at the end (<= 1 ulp off).
__global__ void __launch_bounds__(128)
preprocess_states_kernel(const float* __restrict__ states, const float* __restrict__ actions,
                         const rac_clip_calib* __restrict__ calib, int B, int T, int R, int A_in, int A_out,
                         float* __restrict__ states_out, float* __restrict__ actions_out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;  // (b, t)
  if (i >= B * T) return;
  const int b = i / T, t = i - b * T;
  const rac_clip_calib& c = calib[b];
  const float* s = states + static_cast<size_t>(i) * R;
  float* so = states_out + (static_cast<size_t>(t) * B + b) * R;
  double p[3] = {s[0], s[1], s[2]};
  if (c.kind == 2) {          // franka: onto the locobot frame (:318-322)
    p[0] = static_cast<double>(static_cast<float>(p[0] + c.frame_diff[0]));  // in-place += on the float32 array
    p[1] = static_cast<double>(static_cast<float>(p[1] + c.frame_diff[1]));
    p[2] = static_cast<double>(0.14f);
  } else if (c.kind == 0) {   // RoboNet file: states are normalised in the bounds (:323-324, denormalize :470-473)
    for (int k = 0; k < 3; ++k) p[k] = p[k] * (c.high[k] - c.low[k]) + c.low[k];
  }
  if (c.camera) {             // world -> camera frame (:326-333)
    double q[3];
    for (int r = 0; r < 3; ++r)
      q[r] = c.world2cam[r * 4 + 0] * p[0] + c.world2cam[r * 4 + 1] * p[1] + c.world2cam[r * 4 + 2] * p[2] + c.world2cam[r * 4 + 3];
    p[0] = q[0]; p[1] = q[1]; p[2] = q[2];
  }
  for (int k = 0; k < 3; ++k) so[k] = static_cast<float>((p[k] - c.low[k]) / (c.high[k] - c.low[k]));
  so[3] = s[3];
  so[4] = static_cast<float>((static_cast<double>(s[4]) - c.low[4]) / (c.high[4] - c.low[4]));
  for (int k = 5; k < R; ++k) so[k] = s[k];
  if (actions && t < T - 1) {
    const float* a = actions + (static_cast<size_t>(b) * (T - 1) + t) * A_in;
    float* ao = actions_out + (static_cast<size_t>(t) * B + b) * A_out;
    for (int k = 0; k < A_in; ++k) ao[k] = a[k];
    if (A_out == A_in + 1) {  // autograsp: the gripper command is read off the NEXT state's last column (:179-191)
      const double nxt = states[(static_cast<size_t>(b) * T + t + 1) * R + c.grip_col];
      const double mid = (c.grip_high + c.grip_low) / 2.0;
      ao[A_in] = static_cast<float>(nxt > mid ? c.grip_high : c.grip_low);
    }
  }
}

}  // namespace

cudaError_t launch_preprocess_states(const float* states, const float* actions, const rac_clip_calib* calib, int B, int T,
                                     int R, int A_in, int A_out, float* states_out, float* actions_out, cudaStream_t s) {
  if (B * T < 1) return cudaSuccess;
  preprocess_states_kernel<<<(B * T + 127) / 128, 128, 0, s>>>(states, actions, calib, B, T, R, A_in, A_out, states_out,
                                                               actions_out);
  return cudaGetLastError();
}

cudaError_t launch_process_batch(const uint8_t* frames, const void* masks, int mask_u8, int B, int T, int Hs, int Ws,
                                 const rac_augment* aug, float* img_out, float* mask_out, cudaStream_t s) {
  if (B * T < 1) return cudaSuccess;
  if (Hs == kH && Ws == kW) {
    process_batch_kernel<false><<<B * T, kThreads, kHW * 3 + kHW * 4, s>>>(frames, masks, mask_u8, B, T, Hs, Ws, aug, img_out, mask_out);
    return cudaGetLastError();
  }
  if (Hs < 1 || Ws < 1 || static_cast<long long>(Hs) * Ws * 3 > 0x7fffffffLL / 4) return cudaErrorInvalidValue;
  constexpr int kSmem = kHW * 3 * 4 + kHW * 4;  // 48 KB of float planes + mask: above the default dynamic limit
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(process_batch_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem);
    if (e != cudaSuccess) return e;
    attr = true;
  }
  process_batch_kernel<true><<<B * T, kThreads, kSmem, s>>>(frames, masks, mask_u8, B, T, Hs, Ws, aug, img_out, mask_out);
  return cudaGetLastError();
}

}  // namespace rac
