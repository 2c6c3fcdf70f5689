// Evaluation metrics of the reference on device (src/utils/metrics.py:13-78 psnr / ssim,
// src/prediction/losses.py:80-94 world_psnr_criterion), used by PredictionTrainer._eval_step (trainer.py:685-700),
// which blacks the robot region out with the true mask first (zero_robot_region, src/utils/image.py:5-20) -- that
// masking and the clamp(0, 1) are folded into the loads here. All are single-pass HBM-bound reductions: each image
// plane is read once (coalesced float4 / row loads), partial sums are combined in a fixed order.
#include "misc_kernels.cuh"

#include <math.h>

namespace rac {

namespace {

template <typename T>
__device__ __forceinline__ T block_sum_m(T v, T* sh) {
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

// mode 0: psnr(estimates, targets) of metrics.py:57-78 (data_dims = 3): both go through (x + 1) / 2 first.
//         mask != null: robot pixels of BOTH images are zeroed first; clamp01: clamp(0, 1) after the masking.
// mode 1: world_psnr_criterion(pred, target, mask): squared error over world pixels / (3 * #world pixels + 1).
__global__ void __launch_bounds__(256)
psnr_kernel(const float* __restrict__ est, const float* __restrict__ tgt, const float* __restrict__ mask, int mode,
            int clamp01, float* __restrict__ out, int C, int HW) {
  __shared__ double sh[32];
  const size_t b = blockIdx.x;
  const float* e = est + b * C * HW;
  const float* t = tgt + b * C * HW;
  const float* m = mask ? mask + b * HW : nullptr;
  double acc = 0.0, cnt = 0.0;
  for (int i = threadIdx.x; i < C * HW; i += blockDim.x) {
    const bool robot = m ? (m[i % HW] != 0.f) : false;
    float a = e[i], c = t[i];
    if (mode == 0) {
      if (robot) a = c = 0.f;
      if (clamp01) {
        a = fminf(fmaxf(a, 0.f), 1.f);
        c = fminf(fmaxf(c, 0.f), 1.f);
      }
      const float d = (a + 1.f) / 2.f - (c + 1.f) / 2.f;
      acc += static_cast<double>(d * d);
    } else {
      const float d = c - a;
      if (!robot) {
        acc += static_cast<double>(d * d);
        cnt += 1.0;
      }
    }
  }
  acc = block_sum_m(acc, sh);
  if (mode == 1) cnt = block_sum_m(cnt, sh);
  if (threadIdx.x == 0) {
    const float mse = static_cast<float>(mode == 0 ? acc / static_cast<double>(C * HW) : acc / (cnt + 1.0));
    out[b] = 10.f * logf(1.f / mse) / 2.302585092994046f;
  }
}

// x_pred = (1 - m) * x_j + m * rgb with (rgb, m) = the 4 channels of the decoder output (trainer.py:406-407,653-654;
// trajectory_sampler.py:149-150), NCHW fp32 as the reference interface carries them
__global__ void __launch_bounds__(256)
composite_nchw_kernel(const float* __restrict__ x4, const float* __restrict__ xj, float* __restrict__ out, int HW,
                      long long total) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const long long b = i / (3 * HW);
  const int r = static_cast<int>(i - b * 3 * HW);
  const int c = r / HW, p = r - c * HW;
  const float m = x4[(b * 4 + 3) * HW + p];
  out[i] = (1.f - m) * xj[i] + m * x4[(b * 4 + c) * HW + p];
}

struct SsimWindow {
  float w[11];
};

// One CTA per (sample, channel) plane. Shared memory: the two planes + 5 horizontally filtered planes
// (mu1, mu2, E[x1^2], E[x2^2], E[x1 x2]); the 11x11 gaussian window of the reference is separable (outer product of
// its normalised 1-D window, metrics.py:19-23), zero padding 5 (F.conv2d(padding=window_size // 2)).
__global__ void __launch_bounds__(256)
ssim_kernel(const float* __restrict__ img1, const float* __restrict__ img2, const float* __restrict__ mask,
            const SsimWindow win, float* __restrict__ map_out, float* __restrict__ plane_mean, int C, int H, int W) {
  extern __shared__ float sm[];
  __shared__ double sh[32];
  const int HW = H * W;
  float* p1 = sm;
  float* p2 = sm + HW;
  float* hf = sm + 2 * HW;  // [5][HW]
  const size_t plane = blockIdx.x;
  const float* a = img1 + plane * HW;
  const float* b = img2 + plane * HW;
  const float* m = mask ? mask + (plane / C) * HW : nullptr;
  for (int i = threadIdx.x; i < HW; i += blockDim.x) {
    const bool robot = m ? (m[i] != 0.f) : false;
    p1[i] = robot ? 0.f : a[i];
    p2[i] = robot ? 0.f : b[i];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < HW; i += blockDim.x) {
    const int y = i / W, x = i - y * W;
    float s1 = 0.f, s2 = 0.f, s11 = 0.f, s22 = 0.f, s12 = 0.f;
#pragma unroll
    for (int k = 0; k < 11; ++k) {
      const int xx = x + k - 5;
      if (xx < 0 || xx >= W) continue;
      const float u = p1[y * W + xx], v = p2[y * W + xx], wk = win.w[k];
      s1 += wk * u; s2 += wk * v; s11 += wk * (u * u); s22 += wk * (v * v); s12 += wk * (u * v);
    }
    hf[i] = s1; hf[HW + i] = s2; hf[2 * HW + i] = s11; hf[3 * HW + i] = s22; hf[4 * HW + i] = s12;
  }
  __syncthreads();
  double acc = 0.0;
  for (int i = threadIdx.x; i < HW; i += blockDim.x) {
    const int y = i / W, x = i - y * W;
    float mu1 = 0.f, mu2 = 0.f, e11 = 0.f, e22 = 0.f, e12 = 0.f;
#pragma unroll
    for (int k = 0; k < 11; ++k) {
      const int yy = y + k - 5;
      if (yy < 0 || yy >= H) continue;
      const int j = yy * W + x;
      const float wk = win.w[k];
      mu1 += wk * hf[j]; mu2 += wk * hf[HW + j]; e11 += wk * hf[2 * HW + j]; e22 += wk * hf[3 * HW + j];
      e12 += wk * hf[4 * HW + j];
    }
    const float mu1_sq = mu1 * mu1, mu2_sq = mu2 * mu2, mu12 = mu1 * mu2;
    const float sg1 = e11 - mu1_sq, sg2 = e22 - mu2_sq, sg12 = e12 - mu12;
    const float C1 = 0.01f * 0.01f, C2 = 0.03f * 0.03f;
    const float v = ((2.f * mu12 + C1) * (2.f * sg12 + C2)) / ((mu1_sq + mu2_sq + C1) * (sg1 + sg2 + C2));
    if (map_out) map_out[plane * HW + i] = v;
    acc += static_cast<double>(v);
  }
  if (plane_mean) {
    acc = block_sum_m(acc, sh);
    if (threadIdx.x == 0) plane_mean[plane] = static_cast<float>(acc / static_cast<double>(HW));
  }
}

}  // namespace

cudaError_t launch_psnr(const float* est, const float* tgt, const float* mask, int mode, int clamp01, float* out,
                        int B, int C, int HW, cudaStream_t s) {
  if (B == 0) return cudaSuccess;
  psnr_kernel<<<B, 256, 0, s>>>(est, tgt, mask, mode, clamp01, out, C, HW);
  return cudaGetLastError();
}

cudaError_t launch_composite_nchw(const float* x4, const float* xj, float* out, int B, int HW, cudaStream_t s) {
  const long long total = static_cast<long long>(B) * 3 * HW;
  if (total == 0) return cudaSuccess;
  composite_nchw_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, s>>>(x4, xj, out, HW, total);
  return cudaGetLastError();
}

cudaError_t launch_ssim(const float* img1, const float* img2, const float* mask, float* map_out, float* plane_mean,
                        int B, int C, int H, int W, cudaStream_t s) {
  if (B == 0) return cudaSuccess;
  const size_t smem = static_cast<size_t>(7) * H * W * sizeof(float);
  if (smem > 226 * 1024) return cudaErrorInvalidValue;  // + 256 B of static shared memory
  static size_t attr = 48 * 1024;
  if (smem > attr) {
    cudaError_t e = cudaFuncSetAttribute(ssim_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) return e;
    attr = smem;
  }
  // metrics.py:13-16: fp32 tensor of double-precision exponentials, divided by its fp32 sum
  SsimWindow win;
  float sum = 0.f;
  for (int x = 0; x < 11; ++x) {
    win.w[x] = static_cast<float>(exp(-static_cast<double>((x - 5) * (x - 5)) / (2.0 * 1.5 * 1.5)));
    sum += win.w[x];
  }
  for (int x = 0; x < 11; ++x) win.w[x] /= sum;
  ssim_kernel<<<B * C, 256, smem, s>>>(img1, img2, mask, win, map_out, plane_mean, C, H, W);
  return cudaGetLastError();
}

}  // namespace rac
