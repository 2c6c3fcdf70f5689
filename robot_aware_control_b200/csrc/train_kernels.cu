// Non-GEMM kernels of the SVG training step (see train_kernels.cuh). Batch 16 per GPU: these are small,
// memory-bound kernels; the FLOPs of the step live in conv_tc_kernel (forward, dgrad, wgrad).
#include <algorithm>

#include "train_kernels.cuh"
#include "epilogue.cuh"
#include "ptx.cuh"

namespace rac {

// ------------------------------------------------------------------------------------------------ weights
// Packed operand <-> flat parameter buffer, one CTA per (packed row n, 64-channel block). In the PyTorch layout
// (cout, cin, kh, kw) the `taps` weights of one (n, c) are contiguous and consecutive channels follow each other, so the
// flat side is walked linearly (j = c_local * taps + tap: coalesced) and the packed side [n][tap][c] is written /
// read through a shared-memory tile with c fastest (coalesced too). col_off[c] < 0 marks padding channels.
__global__ void __launch_bounds__(256)
pack_weights_kernel(const float* __restrict__ params, const long long* __restrict__ row_off,
                    const int* __restrict__ col_off, int n_packed, int taps, int ctot, int flip, int tiled,
                    __nv_bfloat16* __restrict__ wp) {
  pdl_entry();
  __shared__ float tile[25][65];
  const int c0 = blockIdx.y * 64;
  // a CTA walks several packed rows (a few thousand CTAs of 1600 elements each were launch-bound, not HBM-bound)
  for (int n = blockIdx.x; n < n_packed; n += gridDim.x) {
    const long long ro = row_off[n];
    for (int j = threadIdx.x; j < 64 * taps; j += blockDim.x) {
      const int cl = j / taps, tap = j - cl * taps;
      const int co = col_off[c0 + cl];
      float v = 0.f;
      if (ro >= 0 && co >= 0) v = __ldcs(params + ro + co + tap);
      tile[flip ? taps - 1 - tap : tap][cl] = v;
    }
    __syncthreads();
    for (int j = threadIdx.x; j < 32 * taps; j += blockDim.x) {
      const int tap = j >> 5, cl = (j & 31) * 2;
      // tiled: k-block-major panels [tap * ctot / 64 + c0 / 64][n][64] (every TMA weight box is one contiguous chunk)
      const long long at = tiled ? ((static_cast<long long>(tap) * (ctot >> 6) + blockIdx.y) * n_packed + n) * 64 + cl
                                 : (static_cast<long long>(n) * taps + tap) * ctot + c0 + cl;
      *reinterpret_cast<__nv_bfloat162*>(wp + at) = __floats2bfloat162_rn(tile[tap][cl], tile[tap][cl + 1]);
    }
    __syncthreads();
  }
}
cudaError_t launch_pack_weights(const float* params, const long long* row_off, const int* col_off, int n_packed,
                                int taps, int ctot, int flip, __nv_bfloat16* wp, cudaStream_t s, int tiled) {
  if (ctot % 64 != 0 || taps > 25) return cudaErrorInvalidValue;
  const int gx = std::min(n_packed, std::max(1, 1184 / (ctot / 64)));
  if (cudaError_t e_ = launch_pdl_small(pack_weights_kernel, dim3(gx, ctot / 64), dim3(256), 0, s, params, row_off, col_off, n_packed, taps, ctot, flip, tiled, wp); e_ != cudaSuccess) return e_;
  return cudaGetLastError();
}

// wd[c][tap][n] = wp[n][taps - 1 - tap][c] (n padded to kpad with zeros): 64 x 64 shared-memory transposes, one CTA per
// (64 n, 64 c) block looping over the taps
__global__ void __launch_bounds__(256)
transpose_flip_kernel(const __nv_bfloat16* __restrict__ wp, int n_packed, int taps, int ctot, int kpad,
                      __nv_bfloat16* __restrict__ wd) {
  pdl_entry();
  __shared__ __nv_bfloat16 tile[64][66];
  const int n0 = blockIdx.x * 64, c0 = blockIdx.y * 64;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  for (int tap = 0; tap < taps; ++tap) {
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      const int n = n0 + ty + r * 8;
      __nv_bfloat162 v = __floats2bfloat162_rn(0.f, 0.f);
      if (n < n_packed)
        v = *reinterpret_cast<const __nv_bfloat162*>(wp + (static_cast<long long>(n) * taps + (taps - 1 - tap)) * ctot + c0 + tx * 2);
      tile[ty + r * 8][tx * 2] = v.x;
      tile[ty + r * 8][tx * 2 + 1] = v.y;
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      const int c = c0 + ty + r * 8;
      const int n = n0 + tx * 2;
      if (n < kpad) {
        __nv_bfloat162 v;
        v.x = tile[tx * 2][ty + r * 8];
        v.y = tile[tx * 2 + 1][ty + r * 8];
        *reinterpret_cast<__nv_bfloat162*>(wd + (static_cast<long long>(c) * taps + tap) * kpad + n) = v;
      }
    }
    __syncthreads();
  }
}
cudaError_t launch_transpose_flip(const __nv_bfloat16* wp, int n_packed, int taps, int ctot, int kpad,
                                  __nv_bfloat16* wd, cudaStream_t s) {
  if (ctot % 64 != 0 || kpad % 2 != 0) return cudaErrorInvalidValue;
  if (cudaError_t e_ = launch_pdl_small(transpose_flip_kernel, dim3((kpad + 63) / 64, ctot / 64), dim3(256), 0, s, wp, n_packed, taps, ctot, kpad, wd); e_ != cudaSuccess) return e_;
  return cudaGetLastError();
}

__global__ void __launch_bounds__(256)
unpack_grads_kernel(const float* __restrict__ dwp, const long long* __restrict__ row_off,
                    const int* __restrict__ col_off, int n_packed, int taps, int ctot, int flip,
                    float* __restrict__ grads) {
  pdl_entry();
  __shared__ float tile[25][65];
  const int c0 = blockIdx.y * 64;
  for (int n = blockIdx.x; n < n_packed; n += gridDim.x) {
    const long long ro = row_off[n];
    if (ro < 0) continue;  // (uniform over the CTA)
    for (int j = threadIdx.x; j < 16 * taps; j += blockDim.x) {
      const int tap = j >> 4, cl = (j & 15) * 4;
      const float4 v = __ldcs(reinterpret_cast<const float4*>(dwp + (static_cast<long long>(n) * taps + tap) * ctot + c0 + cl));
      tile[tap][cl] = v.x; tile[tap][cl + 1] = v.y; tile[tap][cl + 2] = v.z; tile[tap][cl + 3] = v.w;
    }
    __syncthreads();
    for (int j = threadIdx.x; j < 64 * taps; j += blockDim.x) {
      const int cl = j / taps, tap = j - cl * taps;
      const int co = col_off[c0 + cl];
      if (co >= 0) grads[ro + co + tap] = tile[flip ? taps - 1 - tap : tap][cl];
    }
    __syncthreads();
  }
}
cudaError_t launch_unpack_grads(const float* dwp, const long long* row_off, const int* col_off, int n_packed, int taps,
                                int ctot, int flip, float* grads, cudaStream_t s) {
  if (ctot % 64 != 0 || taps > 25) return cudaErrorInvalidValue;
  const int gx = std::min(n_packed, std::max(1, 1184 / (ctot / 64)));
  if (cudaError_t e_ = launch_pdl_small(unpack_grads_kernel, dim3(gx, ctot / 64), dim3(256), 0, s, dwp, row_off, col_off, n_packed, taps, ctot, flip, grads); e_ != cudaSuccess) return e_;
  return cudaGetLastError();
}

// torch.optim.Adam's update of one element with the contraction spelled out (the flat kernel and the fused per-layer
// kernel must give the same bits whatever the compiler would fuse in either code shape)
__device__ __forceinline__ void adam_update(float& p, float g, float& m, float& v, float b1, float b2, float step,
                                            float bc2_sqrt, float eps) {
  m = __fmaf_rn(b1, m, __fmul_rn(1.f - b1, g));
  v = __fmaf_rn(b2, v, __fmul_rn(__fmul_rn(1.f - b2, g), g));
  p = __fsub_rn(p, __fdiv_rn(__fmul_rn(step, m), __fadd_rn(__fdiv_rn(sqrtf(v), bc2_sqrt), eps)));
}

// Fused optimizer step of ONE convolution: packed weight gradient (as the wgrad GEMM left it, dwp[n][tap][c]) -> Adam
// on the layer's slice of the flat fp32 parameters / moments -> the bf16 operand of the next step, in one pass. Replaces
// unpack_grads (4 B read + 4 B written per weight), the flat Adam's gradient read (4 B) and pack_weights (4 B read) for
// the layer: 30 instead of 46 bytes per weight. Same update formula, element by element, as adam_kernel (bit-equal
// parameters); the flat gradient buffer is NOT written for this layer (rac_train_unpack_deferred does it on demand).
// Tiling as pack_weights_kernel: the flat side is walked linearly (64 channels x taps contiguous floats per packed
// row), the packed side with c fastest, transposed through shared memory.
// VEC: the 64 channels of every block are 64 * taps CONTIGUOUS, 16-byte aligned floats of the flat buffers (one
// (cout, cin-range) slab of the PyTorch weight; adam_pack_vec_ok checks the layer's tables once): the flat side then
// moves as float4 (quad q = elements 4q .. 4q+3 of the slab) and the index arithmetic is paid once per quad. ncu of the
// scalar version on the 5x5 gate convolution (profiles/r02_train_top_ncu_s18.txt): 178 instructions per weight, issue
// slots 64 % busy, DRAM 45 % of peak -- instruction-bound, not memory-bound. Same update, element by element.
template <bool VEC>
__global__ void __launch_bounds__(256)
adam_pack_kernel(float* __restrict__ params, float* __restrict__ m, float* __restrict__ v, const float* __restrict__ dwp,
                 const long long* __restrict__ row_off, const int* __restrict__ col_off, int n_packed, int taps, int ctot,
                 int flip, int tiled, __nv_bfloat16* __restrict__ wp, float lr, float b1, float b2, float eps, float bc1,
                 float bc2_sqrt, float gscale) {
  pdl_entry();
  __shared__ float tg[25][65], tp[25][65];
  const int c0 = blockIdx.y * 64;
  const float step = lr / bc1;
  for (int n = blockIdx.x; n < n_packed; n += gridDim.x) {
    const long long ro = row_off[n];
    if constexpr (VEC) {
      constexpr int kQ = 2;  // ceil(16 * 25 / 256) quads per thread
      const int nq = 16 * taps;
      const long long base = ro >= 0 ? ro + col_off[c0] : 0;
      float4* p4 = reinterpret_cast<float4*>(params + base);
      float4* m4 = reinterpret_cast<float4*>(m + base);
      float4* v4 = reinterpret_cast<float4*>(v + base);
      float4 pi[kQ], mi[kQ], vi[kQ];
      if (ro >= 0) {
#pragma unroll
        for (int k = 0; k < kQ; ++k) {
          const int q = threadIdx.x + k * 256;
          if (q < nq) { pi[k] = __ldcs(p4 + q); mi[k] = __ldcs(m4 + q); vi[k] = __ldcs(v4 + q); }
        }
        for (int j = threadIdx.x; j < 16 * taps; j += blockDim.x) {
          const int tap = j >> 4, cl = (j & 15) * 4;
          const float4 g = __ldcs(reinterpret_cast<const float4*>(dwp + (static_cast<long long>(n) * taps + tap) * ctot + c0 + cl));
          tg[tap][cl] = g.x; tg[tap][cl + 1] = g.y; tg[tap][cl + 2] = g.z; tg[tap][cl + 3] = g.w;
        }
      }
      __syncthreads();  // (also: the previous row's operand write has finished reading tp)
#pragma unroll
      for (int k = 0; k < kQ; ++k) {
        const int q = threadIdx.x + k * 256;
        if (q >= nq) continue;
        int cl = (4 * q) / taps, tap = 4 * q - cl * taps;
        float pe[4] = {0.f, 0.f, 0.f, 0.f}, me[4], ve[4];
        if (ro >= 0) {
          pe[0] = pi[k].x; pe[1] = pi[k].y; pe[2] = pi[k].z; pe[3] = pi[k].w;
          me[0] = mi[k].x; me[1] = mi[k].y; me[2] = mi[k].z; me[3] = mi[k].w;
          ve[0] = vi[k].x; ve[1] = vi[k].y; ve[2] = vi[k].z; ve[3] = vi[k].w;
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int tf = flip ? taps - 1 - tap : tap;
          if (ro >= 0) adam_update(pe[i], tg[tf][cl] * gscale, me[i], ve[i], b1, b2, step, bc2_sqrt, eps);
          tp[tf][cl] = pe[i];  // (padding row: the operand gets zeros)
          if (++tap == taps) { tap = 0; ++cl; }
        }
        if (ro >= 0) {
          __stcs(m4 + q, make_float4(me[0], me[1], me[2], me[3]));
          __stcs(v4 + q, make_float4(ve[0], ve[1], ve[2], ve[3]));
          __stcs(p4 + q, make_float4(pe[0], pe[1], pe[2], pe[3]));
        }
      }
      __syncthreads();
    } else {
      // (1) all parameter / moment loads of this thread's (up to 7) elements go out first: three 4-byte streams per
      // element issued one element at a time left ~18 KB in flight per SM (the first version ran at 3 TB/s)
      constexpr int kPer = 7;  // ceil(64 * 25 / 256)
      float pi[kPer], mi[kPer], vi[kPer];
      long long fo[kPer];  // (kept in registers: recomputing them in (3) to gain a CTA per SM measured slower)
      int tfs[kPer], cls[kPer];
#pragma unroll
      for (int k = 0; k < kPer; ++k) {
        const int j = threadIdx.x + k * 256;
        fo[k] = -1;
        if (j < 64 * taps) {
          const int cl = j / taps, tap = j - cl * taps;
          const int co = col_off[c0 + cl];
          tfs[k] = flip ? taps - 1 - tap : tap;
          cls[k] = cl;
          if (ro >= 0 && co >= 0) {
            fo[k] = ro + co + tap;
            pi[k] = __ldcs(params + fo[k]); mi[k] = __ldcs(m + fo[k]); vi[k] = __ldcs(v + fo[k]);
          } else {
            fo[k] = -2;  // padding column / row: the operand gets a zero
          }
        }
      }
      // (2) the packed gradient of the row, transposed through shared memory
      if (ro >= 0) {
        for (int j = threadIdx.x; j < 16 * taps; j += blockDim.x) {
          const int tap = j >> 4, cl = (j & 15) * 4;
          const float4 g = __ldcs(reinterpret_cast<const float4*>(dwp + (static_cast<long long>(n) * taps + tap) * ctot + c0 + cl));
          tg[tap][cl] = g.x; tg[tap][cl + 1] = g.y; tg[tap][cl + 2] = g.z; tg[tap][cl + 3] = g.w;
        }
      }
      __syncthreads();  // (also: the previous row's operand write has finished reading tp)
      // (3) update, store, and the new parameter into the operand tile
#pragma unroll
      for (int k = 0; k < kPer; ++k) {
        if (fo[k] == -1) continue;
        float pn = 0.f;
        if (fo[k] >= 0) {
          adam_update(pi[k], tg[tfs[k]][cls[k]] * gscale, mi[k], vi[k], b1, b2, step, bc2_sqrt, eps);
          __stcs(m + fo[k], mi[k]);
          __stcs(v + fo[k], vi[k]);
          __stcs(params + fo[k], pi[k]);
          pn = pi[k];
        }
        tp[tfs[k]][cls[k]] = pn;
      }
      __syncthreads();
    }
    // (4) bf16 operand of the next step
    for (int j = threadIdx.x; j < 32 * taps; j += blockDim.x) {
      const int tap = j >> 5, cl = (j & 31) * 2;
      const long long at = tiled ? ((static_cast<long long>(tap) * (ctot >> 6) + blockIdx.y) * n_packed + n) * 64 + cl
                                 : (static_cast<long long>(n) * taps + tap) * ctot + c0 + cl;
      *reinterpret_cast<__nv_bfloat162*>(wp + at) = __floats2bfloat162_rn(tp[tap][cl], tp[tap][cl + 1]);
    }
  }
}
// 1 in *ok (preset by the caller) survives iff every 64-channel block of the layer is a contiguous slab (col_off[cb + t] ==
// col_off[cb] + t * taps >= 0) and every real row's slabs start on a 16-byte boundary: the conditions of adam_pack_kernel<true>
__global__ void adam_pack_vec_check_kernel(const long long* __restrict__ row_off, const int* __restrict__ col_off,
                                           int n_packed, int taps, int ctot, int* __restrict__ ok) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int nblk = ctot / 64;
  bool good = true;
  if (i < ctot) {
    const int cb = static_cast<int>(i) / 64 * 64;
    const int co0 = col_off[cb];
    good = co0 >= 0 && col_off[i] == co0 + (static_cast<int>(i) - cb) * taps;
  }
  if (i < static_cast<long long>(n_packed) * nblk) {
    const int n = static_cast<int>(i / nblk), cb = static_cast<int>(i % nblk) * 64;
    const long long ro = row_off[n];
    if (ro >= 0 && ((ro + col_off[cb]) & 3) != 0) good = false;
  }
  if (!good) *ok = 0;
}
cudaError_t adam_pack_vec_ok(const long long* row_off, const int* col_off, int n_packed, int taps, int ctot, int* ok_host) {
  *ok_host = 0;
  if (ctot % 64 != 0 || taps > 25 || n_packed < 1) return cudaSuccess;
  // RAC_ADAM_PACK_VEC=0: always the scalar kernel (A/B switch)
  if (const char* e = getenv("RAC_ADAM_PACK_VEC"))
    if (!atoi(e)) return cudaSuccess;
  int* d_ok = nullptr;
  cudaError_t err = cudaMalloc(&d_ok, sizeof(int));
  if (err != cudaSuccess) return err;
  const int one = 1;
  err = cudaMemcpy(d_ok, &one, sizeof(int), cudaMemcpyHostToDevice);
  if (err == cudaSuccess) {
    const long long total = std::max<long long>(ctot, static_cast<long long>(n_packed) * (ctot / 64));
    adam_pack_vec_check_kernel<<<static_cast<unsigned>((total + 255) / 256), 256>>>(row_off, col_off, n_packed, taps, ctot, d_ok);
    err = cudaGetLastError();
  }
  if (err == cudaSuccess) err = cudaMemcpy(ok_host, d_ok, sizeof(int), cudaMemcpyDeviceToHost);
  cudaFree(d_ok);
  return err;
}
cudaError_t launch_adam_pack(float* params, float* m, float* v, const float* dwp, const long long* row_off,
                             const int* col_off, int n_packed, int taps, int ctot, int flip, int tiled,
                             __nv_bfloat16* wp, float lr, float b1, float b2, float eps, int t, float grad_scale,
                             cudaStream_t s, int vec_ok) {
  if (ctot % 64 != 0 || taps > 25) return cudaErrorInvalidValue;
  const float bc1 = 1.f - powf(b1, static_cast<float>(t));
  const float bc2 = sqrtf(1.f - powf(b2, static_cast<float>(t)));
  const int gx = std::min(n_packed, std::max(1, 1184 / (ctot / 64)));
  if (vec_ok)
    return launch_pdl_small(adam_pack_kernel<true>, dim3(gx, ctot / 64), dim3(256), 0, s, params, m, v, dwp, row_off, col_off,
                            n_packed, taps, ctot, flip, tiled, wp, lr, b1, b2, eps, bc1, bc2, grad_scale);
  return launch_pdl_small(adam_pack_kernel<false>, dim3(gx, ctot / 64), dim3(256), 0, s, params, m, v, dwp, row_off, col_off,
                          n_packed, taps, ctot, flip, tiled, wp, lr, b1, b2, eps, bc1, bc2, grad_scale);
}

__global__ void pack_first_kernel(const float* __restrict__ w, int cin, float* __restrict__ wf) {
  pdl_entry();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;  // (tap, c, o)
  if (i >= 9 * cin * 64) return;
  const int o = i & 63;
  const int c = (i >> 6) % cin;
  const int tap = (i >> 6) / cin;
  wf[i] = w[(o * cin + c) * 9 + tap];
}
cudaError_t launch_pack_first(const float* w, int cin, float* wf, cudaStream_t s) {
  if (cudaError_t e_ = launch_pdl_small(pack_first_kernel, dim3((9 * cin * 64 + 255) / 256), dim3(256), 0, s, w, cin, wf); e_ != cudaSuccess) return e_;
  return cudaGetLastError();
}

// dW[o][c][tap] += sum_pix in[pix][tap, c] * draw[pix][o]. Deterministic: every CTA walks its 64-pixel chunks in a
// fixed order with its (k, o) products in registers and writes ONE partial per CTA; first_wgrad_fold_kernel then adds
// the partials to gw in CTA order (no atomics: the same bits every run).
// Register tile per thread: 3 filter elements (kg, kg + 16, kg + 32) x 4 neighbouring outputs, so one pixel costs three
// broadcast loads + one 128-bit load from shared memory for 12 FMAs (one (k, o) pair per thread and pixel meant two
// loads per FMA: the kernel sat on the shared-memory pipe, 0.33 ms for 1.4 GFLOP).
constexpr int kFwK = 48;  // 9 * cin <= 45 filter elements, padded with zero rows
__global__ void __launch_bounds__(256)
first_wgrad_kernel(const float* __restrict__ img4, const float* __restrict__ mask_a, const float* __restrict__ mask_b,
                   long long mask_bstride, const float* __restrict__ draw, float* __restrict__ part, int B, int H, int W,
                   int cin) {
  pdl_entry();
  __shared__ float sin_[kFwK][65];              // [filter element k = tap * cin + c][pixel]
  __shared__ __align__(16) float sdr[64][68];   // [pixel][output channel]
  const unsigned total = static_cast<unsigned>(B) * H * W;  // (< 2^31: checked by the launcher)
  const int K = 9 * cin;
  const int og = (threadIdx.x & 15) * 4, kg = threadIdx.x >> 4;
  const int fpx = threadIdx.x & 63, fk0 = threadIdx.x >> 6;  // fill role: one pixel, filter elements fk0 + 4 j
  float acc[3][4];
#pragma unroll
  for (int j = 0; j < 3; ++j)
#pragma unroll
    for (int i = 0; i < 4; ++i) acc[j][i] = 0.f;
  for (unsigned p0 = blockIdx.x * 64u; p0 < total; p0 += gridDim.x * 64u) {
    __syncthreads();  // the previous chunk has been consumed
    {
      const unsigned pix = p0 + fpx;
      const bool live = pix < total;
      const int x = static_cast<int>(pix % W), y = static_cast<int>((pix / W) % H);
      const size_t b = pix / (static_cast<unsigned>(W) * H);
#pragma unroll
      for (int j = 0; j < kFwK / 4; ++j) {
        const int k = fk0 + 4 * j;
        float v = 0.f;
        if (live && k < K) {
          const int tap = cin == 3 ? k / 3 : (cin == 4 ? k >> 2 : k / 5);
          const int c = k - tap * cin;
          const int yy = y + tap / 3 - 1, xx = x + tap % 3 - 1;
          if (yy >= 0 && yy < H && xx >= 0 && xx < W) {
            if (c < 3) v = img4[((b * H + yy) * W + xx) * 4 + c];
            else if (c == 3) v = mask_a[b * mask_bstride + static_cast<size_t>(yy) * W + xx];
            else v = mask_b[b * mask_bstride + static_cast<size_t>(yy) * W + xx];
          }
        }
        sin_[k][fpx] = v;
      }
    }
    for (int i = threadIdx.x; i < 64 * 16; i += blockDim.x) {
      const int px = i >> 4, q = i & 15;
      const unsigned pix = p0 + px;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (pix < total) v = __ldcs(reinterpret_cast<const float4*>(draw + static_cast<size_t>(pix) * 64) + q);
      *reinterpret_cast<float4*>(&sdr[px][q * 4]) = v;
    }
    __syncthreads();
#pragma unroll 4
    for (int px = 0; px < 64; ++px) {
      const float4 d = *reinterpret_cast<const float4*>(&sdr[px][og]);
      const float a0 = sin_[kg][px], a1 = sin_[kg + 16][px], a2 = sin_[kg + 32][px];
      acc[0][0] = fmaf(a0, d.x, acc[0][0]); acc[0][1] = fmaf(a0, d.y, acc[0][1]);
      acc[0][2] = fmaf(a0, d.z, acc[0][2]); acc[0][3] = fmaf(a0, d.w, acc[0][3]);
      acc[1][0] = fmaf(a1, d.x, acc[1][0]); acc[1][1] = fmaf(a1, d.y, acc[1][1]);
      acc[1][2] = fmaf(a1, d.z, acc[1][2]); acc[1][3] = fmaf(a1, d.w, acc[1][3]);
      acc[2][0] = fmaf(a2, d.x, acc[2][0]); acc[2][1] = fmaf(a2, d.y, acc[2][1]);
      acc[2][2] = fmaf(a2, d.z, acc[2][2]); acc[2][3] = fmaf(a2, d.w, acc[2][3]);
    }
  }
  // partial of this CTA: part[cta][k * 64 + o]
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    const int k = kg + 16 * j;
    if (k < K)
      *reinterpret_cast<float4*>(part + static_cast<size_t>(blockIdx.x) * K * 64 + k * 64 + og) =
          make_float4(acc[j][0], acc[j][1], acc[j][2], acc[j][3]);
  }
}
__global__ void __launch_bounds__(256)
first_wgrad_fold_kernel(const float* __restrict__ part, int nblk, float* __restrict__ gw, int cin) {
  pdl_entry();
  const int K = 9 * cin;
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= K * 64) return;
  float sum = 0.f;
  for (int b = 0; b < nblk; ++b) sum += part[static_cast<size_t>(b) * K * 64 + p];
  const int k = p >> 6, o = p & 63;
  const int tap = k / cin, c = k - tap * cin;
  gw[(o * cin + c) * 9 + tap] += sum;
}
cudaError_t launch_first_wgrad(const float* img4, const float* mask_a, const float* mask_b, long long mask_bstride,
                               const float* draw, float* gw, int B, int H, int W, int cin, float* part, int max_blocks,
                               cudaStream_t s) {
  const size_t total = static_cast<size_t>(B) * H * W;
  if (cin < 3 || cin > 5 || !part || max_blocks < 1 || total >= (1ull << 31)) return cudaErrorInvalidValue;
  int grid = static_cast<int>((total + 63) / 64);
  if (grid > max_blocks) grid = max_blocks;
  if (grid < 1) return cudaSuccess;
  if (cudaError_t e_ = launch_pdl_small(first_wgrad_kernel, dim3(grid), dim3(256), 0, s, img4, mask_a, mask_b, mask_bstride, draw, part, B, H, W, cin); e_ != cudaSuccess) return e_;
  if (cudaError_t e_ = launch_pdl_small(first_wgrad_fold_kernel, dim3((9 * cin * 64 + 255) / 256), dim3(256), 0, s, part, grid, gw, cin); e_ != cudaSuccess) return e_;
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------ BatchNorm
// Per-channel reductions over all M rows, fp64, deterministic: grid (C/32, R) blocks of (32 channels x 32 row lanes)
// write R partial sums per channel; a second tiny kernel folds them in a fixed order.
constexpr int kRedSplit = 64;   // row splits per group
constexpr int kRedSlots = 512;  // groups x splits (a group = the rows of one time step: BatchNorm statistics are per step)
__device__ double g_red_part[2 * kRedSlots * 2048];  // [which][group * splits + split][channel], channels <= 2048

template <typename F>
__device__ __forceinline__ void column_partial2(int M, int C, F load) {
  // M = rows per group; blockIdx.z = group (rows [z * M, (z + 1) * M) of the tensor)
  __shared__ double sh1[32][33], sh2[32][33];
  const int c = blockIdx.x * 32 + threadIdx.x;
  const int rows_per = (M + gridDim.y - 1) / gridDim.y;
  const int m_base = blockIdx.z * M;
  const int m_lo = m_base + blockIdx.y * rows_per;
  const int m_hi = min(m_base + M, m_lo + rows_per);
  double a = 0.0, b = 0.0;
  if (c < C) {
    for (int m = m_lo + threadIdx.y; m < m_hi; m += 32) {
      float v1, v2;
      load(m, c, v1, v2);
      a += v1;
      b += v2;
    }
  }
  sh1[threadIdx.y][threadIdx.x] = a;
  sh2[threadIdx.y][threadIdx.x] = b;
  __syncthreads();
  if (threadIdx.y == 0 && c < C) {
    double s1 = 0.0, s2 = 0.0;
    for (int i = 0; i < 32; ++i) {
      s1 += sh1[i][threadIdx.x];
      s2 += sh2[i][threadIdx.x];
    }
    const int slot = blockIdx.z * gridDim.y + blockIdx.y;
    g_red_part[(0 * kRedSlots + slot) * 2048 + c] = s1;
    g_red_part[(1 * kRedSlots + slot) * 2048 + c] = s2;
  }
}
__device__ __forceinline__ void column_fold2(int c, int nsplit, double& s1, double& s2, int group = 0) {
  s1 = s2 = 0.0;
  for (int r = group * nsplit; r < (group + 1) * nsplit; ++r) {
    s1 += __ldcg(&g_red_part[(0 * kRedSlots + r) * 2048 + c]);  // L2: written by other CTAs (of this launch, possibly)
    s2 += __ldcg(&g_red_part[(1 * kRedSlots + r) * 2048 + c]);
  }
}
inline int red_split(int M) {
  int r = M / 512;
  return r < 1 ? 1 : (r > kRedSplit ? kRedSplit : r);
}
// "last block folds": every CTA of a column_partial2 grid publishes its partial sums, then takes a ticket; the CTA that
// draws the last ticket of its channel block (blockIdx.x) sees all R partials of these 32 channels (release / acquire
// through the fences around the atomic) and folds them in the fixed order r = 0 .. R-1 -- the result does not depend on
// which CTA happens to be last, so the reduction stays deterministic while the separate fold launch disappears.
__device__ unsigned int g_red_ticket[64];  // one counter per 32-channel block (C <= 2048), self-resetting
__device__ __forceinline__ bool last_block_of_column_group() {
  __shared__ bool s_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0 && threadIdx.y == 0) {
    const unsigned int t = atomicAdd(&g_red_ticket[blockIdx.x], 1u);
    s_last = (t == gridDim.y * gridDim.z - 1);
    if (s_last) g_red_ticket[blockIdx.x] = 0u;  // ready for the next launch (stream order)
  }
  __syncthreads();
  if (s_last) __threadfence();
  return s_last;
}

// Bandwidth-oriented variant for C % 4 == 0 (the BatchNorm layers: 64-512 channels, up to 245 760 rows): 256 threads,
// each owns 4 adjacent channels (one 128-bit load per row) of every RL-th row of the CTA's row range; the CTA covers
// Qb = 256 / RL channel quads (blockIdx.x selects which), so a warp reads >= 128 contiguous bytes per row and the
// loads of 4 rows are in flight per thread. (The one-channel-per-thread kernel above ran at 0.6 TB/s on the
// full-resolution layers, profiles/r02_launches_train_summary_s10.txt.) Partials land in the same g_red_part slots.
template <typename F>
__device__ __forceinline__ void column_partial2_v4(int M, int C, int Qb, F load4) {
  __shared__ double sh[8][256];
  const int tid = threadIdx.x;
  const int RL = 256 / Qb;
  const int q = tid % Qb, rl = tid / Qb;
  const int c0 = (blockIdx.x * Qb + q) * 4;
  const int rows_per = (M + gridDim.y - 1) / gridDim.y;
  const int m_base = blockIdx.z * M;
  const int m_lo = m_base + blockIdx.y * rows_per;
  const int m_hi = min(m_base + M, m_lo + rows_per);
  double acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (c0 < C) {
    int m = m_lo + rl;
    for (; m + 3 * RL < m_hi; m += 4 * RL) {
      float4 a[4], b[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) load4(m + u * RL, c0, a[u], b[u]);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        acc[0] += a[u].x; acc[1] += a[u].y; acc[2] += a[u].z; acc[3] += a[u].w;
        acc[4] += b[u].x; acc[5] += b[u].y; acc[6] += b[u].z; acc[7] += b[u].w;
      }
    }
    for (; m < m_hi; m += RL) {
      float4 a, b;
      load4(m, c0, a, b);
      acc[0] += a.x; acc[1] += a.y; acc[2] += a.z; acc[3] += a.w;
      acc[4] += b.x; acc[5] += b.y; acc[6] += b.z; acc[7] += b.w;
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) sh[j][tid] = acc[j];
  __syncthreads();
  if (rl == 0 && c0 < C) {
    const int slot = blockIdx.z * gridDim.y + blockIdx.y;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      double t = 0.0;
      for (int r = 0; r < RL; ++r) t += sh[j][r * Qb + q];  // fixed order
      g_red_part[((j >> 2) * kRedSlots + slot) * 2048 + c0 + (j & 3)] = t;
    }
  }
}
// channel quads per CTA of the v4 kernels: as few channel blocks as still give ~2 CTAs per SM
inline int red_quads_per_cta(int C, int ctas_rows) {
  int Qb = std::min(C / 4, 256);
  while (Qb > 8 && (C / 4 / Qb) * ctas_rows < 296) Qb /= 2;
  return Qb;
}
// the "last block folds" ticket of the v4 kernels (1-D thread blocks)
__device__ __forceinline__ bool last_block_of_column_group_v4() {
  __shared__ bool s_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int t = atomicAdd(&g_red_ticket[blockIdx.x], 1u);
    s_last = (t == gridDim.y * gridDim.z - 1);
    if (s_last) g_red_ticket[blockIdx.x] = 0u;
  }
  __syncthreads();
  if (s_last) __threadfence();
  return s_last;
}

__global__ void __launch_bounds__(1024)
bn_stats_kernel(const float* __restrict__ raw, int M, int C, float* __restrict__ mean, float* __restrict__ rstd,
                float* __restrict__ rmean, float* __restrict__ rvar, int updates) {
  pdl_entry();
  column_partial2(M, C, [&](int m, int c, float& v1, float& v2) {
    const float x = raw[static_cast<size_t>(m) * C + c];
    v1 = x;
    v2 = x * x;
  });
  if (!last_block_of_column_group()) return;
  const int c = blockIdx.x * 32 + threadIdx.x;
  if (threadIdx.y != 0 || c >= C) return;
  const int nsplit = gridDim.y, groups = gridDim.z;
  float rm = rmean[c], rv = rvar[c];
  for (int grp = 0; grp < groups; ++grp) {  // groups = time steps, in order: the running statistics are a recurrence
    double s1, s2;
    column_fold2(c, nsplit, s1, s2, grp);
    const double mu = s1 / M;
    double var = s2 / M - mu * mu;
    if (var < 0.0) var = 0.0;
    mean[grp * C + c] = static_cast<float>(mu);
    rstd[grp * C + c] = static_cast<float>(1.0 / sqrt(var + 1e-5));
    const double unbiased = var * M / (M - 1);
    for (int u = 0; u < updates; ++u) {  // reference: momentum 0.1, the encoder runs twice per step (dynamics.py:619)
      rm = 0.9f * rm + 0.1f * static_cast<float>(mu);
      rv = 0.9f * rv + 0.1f * static_cast<float>(unbiased);
    }
  }
  rmean[c] = rm;
  rvar[c] = rv;
}
// Collective fold of the last CTA (256 threads): the sums over the `nsplit` row-split partials of group `grp` for the
// channels [cb0, cb0 + cn) (cn <= 256, a power of two). 256 / cn threads share a channel (each takes every SL-th slot,
// 8 independent L2 loads in flight), lanes are combined in a fixed order: deterministic, and ~30x shorter than one
// thread walking all slots with dependent loads (that serial fold was ~100 us of the full-resolution layers' kernels).
// Returns the sums to the threads with tid < cn (channel cb0 + tid); every thread of the CTA must call it.
__device__ __forceinline__ void column_fold2_cta(int cb0, int cn, int grp, int nsplit, double& s1, double& s2) {
  __shared__ double shf[2][256];
  const int tid = threadIdx.x;
  const int SL = 256 / cn;
  const int cl = tid % cn, sl = tid / cn;
  double a = 0.0, b = 0.0;
  const int r0 = grp * nsplit;
  for (int r = sl; r < nsplit; r += 8 * SL) {
    double va[8], vb[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int rr = r + j * SL;
      const bool ok = rr < nsplit;
      va[j] = ok ? __ldcg(&g_red_part[(0 * kRedSlots + r0 + rr) * 2048 + cb0 + cl]) : 0.0;
      vb[j] = ok ? __ldcg(&g_red_part[(1 * kRedSlots + r0 + rr) * 2048 + cb0 + cl]) : 0.0;
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) { a += va[j]; b += vb[j]; }
  }
  __syncthreads();  // (previous call's readers are done)
  shf[0][tid] = a;
  shf[1][tid] = b;
  __syncthreads();
  s1 = s2 = 0.0;
  if (tid < cn)
    for (int l = 0; l < SL; ++l) { s1 += shf[0][l * cn + tid]; s2 += shf[1][l * cn + tid]; }
}

__global__ void __launch_bounds__(256)
bn_stats_v4_kernel(const float* __restrict__ raw, int M, int C, int Qb, float* __restrict__ mean, float* __restrict__ rstd,
                   float* __restrict__ rmean, float* __restrict__ rvar, int updates) {
  pdl_entry();
  column_partial2_v4(M, C, Qb, [&](int m, int c0, float4& v1, float4& v2) {
    const float4 x = *reinterpret_cast<const float4*>(raw + static_cast<size_t>(m) * C + c0);
    v1 = x;
    v2 = make_float4(x.x * x.x, x.y * x.y, x.z * x.z, x.w * x.w);
  });
  if (!last_block_of_column_group_v4()) return;
  const int nsplit = gridDim.y, groups = gridDim.z;
  const int cb = blockIdx.x * Qb * 4, cend = min(C, cb + Qb * 4);
  for (int cb0 = cb; cb0 < cend; cb0 += 256) {
    const int cn = min(256, cend - cb0);
    const int c = cb0 + threadIdx.x;
    const bool mine = static_cast<int>(threadIdx.x) < cn;
    float rm = mine ? rmean[c] : 0.f, rv = mine ? rvar[c] : 0.f;
    for (int grp = 0; grp < groups; ++grp) {  // groups = time steps, in order: the running statistics are a recurrence
      double s1, s2;
      column_fold2_cta(cb0, cn, grp, nsplit, s1, s2);
      if (!mine) continue;
      const double mu = s1 / M;
      double var = s2 / M - mu * mu;
      if (var < 0.0) var = 0.0;
      mean[grp * C + c] = static_cast<float>(mu);
      rstd[grp * C + c] = static_cast<float>(1.0 / sqrt(var + 1e-5));
      const double unbiased = var * M / (M - 1);
      for (int u = 0; u < updates; ++u) {
        rm = 0.9f * rm + 0.1f * static_cast<float>(mu);
        rv = 0.9f * rv + 0.1f * static_cast<float>(unbiased);
      }
    }
    if (mine) { rmean[c] = rm; rvar[c] = rv; }
  }
}
cudaError_t launch_bn_stats(const float* raw, int M, int C, float* mean, float* rstd, float* running_mean,
                            float* running_var, int updates, cudaStream_t s, int groups) {
  if (C > 2048 || groups < 1) return cudaErrorInvalidValue;
  int R = red_split(M);
  while (R * groups > kRedSlots) R /= 2;
  if (R < 1) return cudaErrorInvalidValue;
  if ((C & (C - 1)) == 0 && C >= 32 && C <= 1024) {
    const int Qb = red_quads_per_cta(C, R * groups);
    if (cudaError_t e_ = launch_pdl_small(bn_stats_v4_kernel, dim3(C / 4 / Qb, R, groups), dim3(256), 0, s, raw, M, C, Qb, mean, rstd, running_mean, running_var, updates); e_ != cudaSuccess) return e_;
    return cudaGetLastError();
  }
  if (cudaError_t e_ = launch_pdl_small(bn_stats_kernel, dim3((C + 31) / 32, R, groups), dim3(32, 32), 0, s, raw, M, C, mean, rstd, running_mean, running_var, updates); e_ != cudaSuccess) return e_;
  return cudaGetLastError();
}

__global__ void __launch_bounds__(256)
bn_act_kernel(const float* __restrict__ raw, const float* __restrict__ mean, const float* __restrict__ rstd,
              const float* __restrict__ gamma, const float* __restrict__ beta, int B, int H, int W, int C,
              __nv_bfloat16* __restrict__ out, int cstride, int coff, int upsample, int rows_per_group) {
  pdl_entry();
  const int C8 = C / 8;
  const size_t total = static_cast<size_t>(B) * H * W * C8;
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int c0 = static_cast<int>(i % C8) * 8;
  const size_t m = i / C8;
  const size_t so = (m / rows_per_group) * C;  // statistics of this row's group (time step)
  mean += so;
  rstd += so;
  const int x = static_cast<int>(m % W), y = static_cast<int>((m / W) % H);
  const size_t b = m / (static_cast<size_t>(W) * H);
  float v[8];
  const float4* rp = reinterpret_cast<const float4*>(raw + m * C + c0);  // (C % 8 == 0: 32-byte aligned)
  const float4 r0 = __ldcs(rp), r1 = __ldcs(rp + 1);                     // last use of the fp32 copy in the forward pass
  const float rw[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int c = c0 + j;
    float t = (rw[j] - mean[c]) * rstd[c] * gamma[c] + beta[c];
    v[j] = t > 0.f ? t : 0.2f * t;
  }
  const uint4 pk = make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]),
                              pack_bf16x2(v[6], v[7]));
  if (!upsample) {
    *reinterpret_cast<uint4*>(out + m * cstride + coff + c0) = pk;
  } else {
    for (int dy = 0; dy < 2; ++dy)
      for (int dx = 0; dx < 2; ++dx)
        *reinterpret_cast<uint4*>(out + ((b * 2 * H + 2 * y + dy) * (2 * W) + 2 * x + dx) * cstride + coff + c0) = pk;
  }
}
cudaError_t launch_bn_act(const float* raw, const float* mean, const float* rstd, const float* gamma,
                          const float* beta, int B, int H, int W, int C, __nv_bfloat16* out, int cstride, int coff,
                          int upsample, cudaStream_t s, int groups) {
  const size_t total = static_cast<size_t>(B) * H * W * (C / 8);
  if (groups < 1 || B % groups) return cudaErrorInvalidValue;
  if (cudaError_t e_ = launch_pdl_small(bn_act_kernel, dim3(static_cast<unsigned>((total + 255) / 256)), dim3(256), 0, s, raw, mean, rstd, gamma, beta, B, H, W, C,
                                                                           out, cstride, coff, upsample, B / groups * H * W); e_ != cudaSuccess) return e_;
  return cudaGetLastError();
}

struct BnBwdArgs {
  const float* dy;
  int dy_cstride, dy_coff, upsample;
  const float* raw;
  const float* mean;   // [groups][C]
  const float* rstd;
  const float* gamma;
  const float* beta;
  int H, W, C;
  int rows_per_group;  // rows of one time step (the unit of the BatchNorm statistics)
};
// gradient w.r.t. the BN output after the LeakyReLU derivative, and the normalised activation
__device__ __forceinline__ void bn_bwd_point(const BnBwdArgs& a, int m, int c, float& dz, float& xhat) {
  const int x = m % a.W, y = (m / a.W) % a.H;
  const size_t b = static_cast<size_t>(m) / (a.W * a.H);
  float g;
  if (!a.upsample) {
    g = a.dy[static_cast<size_t>(m) * a.dy_cstride + a.dy_coff + c];
  } else {
    g = 0.f;
    for (int dy = 0; dy < 2; ++dy)
      for (int dx = 0; dx < 2; ++dx)
        g += a.dy[((b * 2 * a.H + 2 * y + dy) * (2 * a.W) + 2 * x + dx) * a.dy_cstride + a.dy_coff + c];
  }
  const int so = (m / a.rows_per_group) * a.C;
  xhat = (a.raw[static_cast<size_t>(m) * a.C + c] - a.mean[so + c]) * a.rstd[so + c];
  const float bn = xhat * a.gamma[c] + a.beta[c];
  dz = bn > 0.f ? g : 0.2f * g;
}
__global__ void __launch_bounds__(1024)
bn_bwd_sums_kernel(BnBwdArgs a, int M, float* __restrict__ scratch, float* __restrict__ dgamma, float* __restrict__ dbeta) {
  pdl_entry();
  column_partial2(M, a.C, [&](int m, int c, float& v1, float& v2) {
    float dz, xh;
    bn_bwd_point(a, m, c, dz, xh);
    v1 = dz;
    v2 = dz * xh;
  });
  if (!last_block_of_column_group()) return;
  const int C = a.C;
  const int c = blockIdx.x * 32 + threadIdx.x;
  if (threadIdx.y != 0 || c >= C) return;
  double t1 = 0.0, t2 = 0.0;
  for (int grp = 0; grp < static_cast<int>(gridDim.z); ++grp) {
    double s1, s2;
    column_fold2(c, gridDim.y, s1, s2, grp);
    scratch[grp * 2 * C + c] = static_cast<float>(s1);
    scratch[grp * 2 * C + C + c] = static_cast<float>(s2);
    t1 += s1;
    t2 += s2;
  }
  dbeta[c] += static_cast<float>(t1);
  dgamma[c] += static_cast<float>(t2);
}
__global__ void __launch_bounds__(256)
bn_bwd_sums_v4_kernel(BnBwdArgs a, int M, int Qb, float* __restrict__ scratch, float* __restrict__ dgamma,
                      float* __restrict__ dbeta) {
  pdl_entry();
  column_partial2_v4(M, a.C, Qb, [&](int m, int c0, float4& v1, float4& v2) {
    float4 g;
    if (!a.upsample) {
      g = *reinterpret_cast<const float4*>(a.dy + static_cast<size_t>(m) * a.dy_cstride + a.dy_coff + c0);
    } else {
      const int x = m % a.W, y = (m / a.W) % a.H;
      const size_t b = static_cast<size_t>(m) / (a.W * a.H);
      g = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int dy = 0; dy < 2; ++dy)
#pragma unroll
        for (int dx = 0; dx < 2; ++dx) {
          const float4 t = *reinterpret_cast<const float4*>(
              a.dy + ((b * 2 * a.H + 2 * y + dy) * (2 * a.W) + 2 * x + dx) * a.dy_cstride + a.dy_coff + c0);
          g.x += t.x; g.y += t.y; g.z += t.z; g.w += t.w;
        }
    }
    const int so = (m / a.rows_per_group) * a.C + c0;
    const float4 r = *reinterpret_cast<const float4*>(a.raw + static_cast<size_t>(m) * a.C + c0);
    const float4 mu = *reinterpret_cast<const float4*>(a.mean + so), rs = *reinterpret_cast<const float4*>(a.rstd + so);
    const float4 ga = *reinterpret_cast<const float4*>(a.gamma + c0), be = *reinterpret_cast<const float4*>(a.beta + c0);
    const float gg[4] = {g.x, g.y, g.z, g.w}, rr[4] = {r.x, r.y, r.z, r.w}, mm[4] = {mu.x, mu.y, mu.z, mu.w};
    const float ss[4] = {rs.x, rs.y, rs.z, rs.w}, gm[4] = {ga.x, ga.y, ga.z, ga.w}, bt[4] = {be.x, be.y, be.z, be.w};
    float dz[4], dzx[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float xh = (rr[j] - mm[j]) * ss[j];
      const float bn = xh * gm[j] + bt[j];
      dz[j] = bn > 0.f ? gg[j] : 0.2f * gg[j];
      dzx[j] = dz[j] * xh;
    }
    v1 = make_float4(dz[0], dz[1], dz[2], dz[3]);
    v2 = make_float4(dzx[0], dzx[1], dzx[2], dzx[3]);
  });
  if (!last_block_of_column_group_v4()) return;
  const int C = a.C;
  const int cb = blockIdx.x * Qb * 4, cend = min(C, cb + Qb * 4);
  for (int cb0 = cb; cb0 < cend; cb0 += 256) {
    const int cn = min(256, cend - cb0);
    const int c = cb0 + threadIdx.x;
    const bool mine = static_cast<int>(threadIdx.x) < cn;
    double t1 = 0.0, t2 = 0.0;
    for (int grp = 0; grp < static_cast<int>(gridDim.z); ++grp) {
      double s1, s2;
      column_fold2_cta(cb0, cn, grp, gridDim.y, s1, s2);
      if (!mine) continue;
      scratch[grp * 2 * C + c] = static_cast<float>(s1);
      scratch[grp * 2 * C + C + c] = static_cast<float>(s2);
      t1 += s1;
      t2 += s2;
    }
    if (mine) {
      dbeta[c] += static_cast<float>(t1);
      dgamma[c] += static_cast<float>(t2);
    }
  }
}
// 8 channels per thread (C % 8 == 0): 128-bit loads of dy / raw, one 128-bit bf16 store
__global__ void __launch_bounds__(256)
bn_bwd_apply_kernel(BnBwdArgs a, int M, const float* __restrict__ scratch, __nv_bfloat16* __restrict__ draw,
                    float* __restrict__ draw32) {
  pdl_entry();
  const int C8 = a.C >> 3;
  const size_t total = static_cast<size_t>(M) * C8;
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int c0 = static_cast<int>(i % C8) * 8;
  const int m = static_cast<int>(i / C8);
  const int x = m % a.W, y = (m / a.W) % a.H;
  const size_t b = static_cast<size_t>(m) / (a.W * a.H);
  float g[8];
  if (!a.upsample) {
    const float4* p = reinterpret_cast<const float4*>(a.dy + static_cast<size_t>(m) * a.dy_cstride + a.dy_coff + c0);
    const float4 u = p[0], v = p[1];
    g[0] = u.x; g[1] = u.y; g[2] = u.z; g[3] = u.w; g[4] = v.x; g[5] = v.y; g[6] = v.z; g[7] = v.w;
  } else {
#pragma unroll
    for (int j = 0; j < 8; ++j) g[j] = 0.f;
    for (int dy = 0; dy < 2; ++dy)
      for (int dx = 0; dx < 2; ++dx) {
        const float4* p = reinterpret_cast<const float4*>(
            a.dy + ((b * 2 * a.H + 2 * y + dy) * (2 * a.W) + 2 * x + dx) * a.dy_cstride + a.dy_coff + c0);
        const float4 u = p[0], v = p[1];
        g[0] += u.x; g[1] += u.y; g[2] += u.z; g[3] += u.w; g[4] += v.x; g[5] += v.y; g[6] += v.z; g[7] += v.w;
      }
  }
  const float4* rp = reinterpret_cast<const float4*>(a.raw + static_cast<size_t>(m) * a.C + c0);
  const float4 r0 = rp[0], r1 = rp[1];
  const float raw[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
  const int grp = m / a.rows_per_group;
  const float inv = 1.f / a.rows_per_group;
  const float* sc = scratch + grp * 2 * a.C;
  const float* mean = a.mean + grp * a.C;
  const float* rstd = a.rstd + grp * a.C;
  float dx[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int c = c0 + j;
    const float xh = (raw[j] - mean[c]) * rstd[c];
    const float bn = xh * a.gamma[c] + a.beta[c];
    const float dz = bn > 0.f ? g[j] : 0.2f * g[j];
    dx[j] = a.gamma[c] * rstd[c] * (dz - sc[c] * inv - xh * sc[a.C + c] * inv);
  }
  if (draw)
    *reinterpret_cast<uint4*>(draw + static_cast<size_t>(m) * a.C + c0) =
        make_uint4(pack_bf16x2(dx[0], dx[1]), pack_bf16x2(dx[2], dx[3]), pack_bf16x2(dx[4], dx[5]), pack_bf16x2(dx[6], dx[7]));
  if (draw32) {
    float4* o = reinterpret_cast<float4*>(draw32 + static_cast<size_t>(m) * a.C + c0);
    o[0] = make_float4(dx[0], dx[1], dx[2], dx[3]);
    o[1] = make_float4(dx[4], dx[5], dx[6], dx[7]);
  }
}
cudaError_t launch_bn_bwd(const float* dy, int dy_cstride, int dy_coff, int upsample, const float* raw,
                          const float* mean, const float* rstd, const float* gamma, const float* beta, int B, int H,
                          int W, int C, float* scratch, __nv_bfloat16* draw, float* draw_f32_or_null, float* dgamma,
                          float* dbeta, cudaStream_t s, int groups) {
  if (C > 2048 || groups < 1 || B % groups) return cudaErrorInvalidValue;
  const int Mg = B / groups * H * W;
  BnBwdArgs a{dy, dy_cstride, dy_coff, upsample, raw, mean, rstd, gamma, beta, H, W, C, Mg};
  const int M = B * H * W;
  int R = red_split(Mg);
  while (R * groups > kRedSlots) R /= 2;
  if (R < 1) return cudaErrorInvalidValue;
  if ((C & (C - 1)) == 0 && C >= 32 && C <= 1024 && dy_cstride % 4 == 0 && dy_coff % 4 == 0) {
    const int Qb = red_quads_per_cta(C, R * groups);
    if (cudaError_t e_ = launch_pdl_small(bn_bwd_sums_v4_kernel, dim3(C / 4 / Qb, R, groups), dim3(256), 0, s, a, Mg, Qb, scratch, dgamma, dbeta); e_ != cudaSuccess) return e_;
  } else {
    if (cudaError_t e_ = launch_pdl_small(bn_bwd_sums_kernel, dim3((C + 31) / 32, R, groups), dim3(32, 32), 0, s, a, Mg, scratch, dgamma, dbeta); e_ != cudaSuccess) return e_;
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  if (C % 8 || dy_cstride % 4 || dy_coff % 4) return cudaErrorInvalidValue;
  const size_t total = static_cast<size_t>(M) * (C / 8);
  if (cudaError_t e_ = launch_pdl_small(bn_bwd_apply_kernel, dim3(static_cast<unsigned>((total + 255) / 256)), dim3(256), 0, s, a, M, scratch, draw, draw_f32_or_null); e_ != cudaSuccess) return e_;
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------ LSTM cell
// Forward cell of the training step when the gate convolution ran as split-K work items (and, for a teacher-forced
// clip, with the input half of K computed for all time steps by one earlier GEMM): pre-activation = gx + sum of the
// slices (fixed order) + bias; then the ConvLSTMCell update (lstm.py:135-149) with the same libm math as EPI_LSTM_TRAIN.
// Columns are (channel, gate) interleaved, gate order in / remember / out / cell: one thread per (row, channel).
__global__ void __launch_bounds__(256)
lstm_cell_fwd_kernel(const float* __restrict__ gx, const float* __restrict__ part, int nsplit, long long split_stride,
                     const float* __restrict__ bias, const float* __restrict__ c_prev, float* __restrict__ c_out,
                     __nv_bfloat16* __restrict__ h_out, float* __restrict__ gates_out, size_t total, int hid) {
  pdl_entry();
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;  // (m, channel)
  if (i >= total) return;
  const int ch = static_cast<int>(i % hid);
  float4 v = __ldg(reinterpret_cast<const float4*>(bias) + ch);
  if (gx) {
    const float4 o = reinterpret_cast<const float4*>(gx)[i];
    v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w;
  }
  for (int k = 0; k < nsplit; ++k) {
    const float4 o = reinterpret_cast<const float4*>(part + k * split_stride)[i];
    v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w;
  }
  const float ig = 1.0f / (1.0f + expf(-v.x)), fg = 1.0f / (1.0f + expf(-v.y)), og = 1.0f / (1.0f + expf(-v.z));
  const float cg = tanhf(v.w);
  const float cn = fg * (c_prev ? c_prev[i] : 0.f) + ig * cg;
  c_out[i] = cn;
  h_out[i] = __float2bfloat16_rn(og * tanhf(cn));
  reinterpret_cast<float4*>(gates_out)[i] = make_float4(ig, fg, og, cg);
}
cudaError_t launch_lstm_cell_fwd(const float* gx, const float* part, int nsplit, long long split_stride,
                                 const float* bias, const float* c_prev_or_null, float* c_out, __nv_bfloat16* h_out,
                                 float* gates_out, int M, int hid, cudaStream_t s) {
  const size_t total = static_cast<size_t>(M) * hid;
  if (cudaError_t e_ = launch_pdl_small(lstm_cell_fwd_kernel, dim3(static_cast<unsigned>((total + 255) / 256)), dim3(256), 0, s, gx, part, nsplit, split_stride, bias,
                                                                                  c_prev_or_null, c_out, h_out, gates_out,
                                                                                  total, hid); e_ != cudaSuccess) return e_;
  return cudaGetLastError();
}

__global__ void __launch_bounds__(256)
lstm_bwd_kernel(const float* __restrict__ dh, float* __restrict__ dc, const float* __restrict__ gates,
                const float* __restrict__ c_prev, const float* __restrict__ c_new, size_t total,
                __nv_bfloat16* __restrict__ dgates) {
  pdl_entry();
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;  // (m, channel)
  if (i >= total) return;
  const float4 gt = reinterpret_cast<const float4*>(gates)[i];  // i, f, o, g (post-activation)
  const float cp = c_prev ? c_prev[i] : 0.f;
  const float tc = tanhf(c_new[i]);  // the training forward uses libm tanh as well (EpiParams::exact_math)
  const float dhv = dh[i];
  const float dcv = dc[i] + dhv * gt.z * (1.f - tc * tc);
  const float d_o = dhv * tc;
  const float di = dcv * gt.w, dg = dcv * gt.x, df = dcv * cp;
  dc[i] = dcv * gt.y;
  const float pi = di * gt.x * (1.f - gt.x), pf = df * gt.y * (1.f - gt.y), po = d_o * gt.z * (1.f - gt.z);
  const float pg = dg * (1.f - gt.w * gt.w);
  reinterpret_cast<uint2*>(dgates)[i] = make_uint2(pack_bf16x2(pi, pf), pack_bf16x2(po, pg));
}
cudaError_t launch_lstm_bwd(const float* dh, float* dc, const float* gates, const float* c_prev_or_null,
                            const float* c_new, int M, int hid, __nv_bfloat16* dgates, cudaStream_t s) {
  const size_t total = static_cast<size_t>(M) * hid;
  if (cudaError_t e_ = launch_pdl_small(lstm_bwd_kernel, dim3(static_cast<unsigned>((total + 255) / 256)), dim3(256), 0, s, dh, dc, gates, c_prev_or_null, c_new,
                                                                             total, dgates); e_ != cudaSuccess) return e_;
  return cudaGetLastError();
}

__global__ void __launch_bounds__(1024) bias_grad_partial_kernel(const __nv_bfloat16* __restrict__ dy, int M, int ncols,
                                                                  int nvalid) {
  pdl_entry();
  column_partial2(M, nvalid, [&](int m, int c, float& v1, float& v2) {
    v1 = __bfloat162float(dy[static_cast<size_t>(m) * ncols + c]);
    v2 = 0.f;
  });
}
__global__ void bias_grad_final_kernel(int nvalid, int nsplit, const long long* __restrict__ bias_off,
                                       float* __restrict__ grads) {
  pdl_entry();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= nvalid || bias_off[c] < 0) return;
  double s1, s2;
  column_fold2(c, nsplit, s1, s2);
  grads[bias_off[c]] += static_cast<float>(s1);
}
cudaError_t launch_bias_grad(const __nv_bfloat16* dy, int M, int ncols, int nvalid, const long long* bias_off,
                             float* grads, cudaStream_t s) {
  if (nvalid > 2048) return cudaErrorInvalidValue;
  const int R = red_split(M);
  if (cudaError_t e_ = launch_pdl_small(bias_grad_partial_kernel, dim3((nvalid + 31) / 32, R), dim3(32, 32), 0, s, dy, M, ncols, nvalid); e_ != cudaSuccess) return e_;
  if (cudaError_t e_ = launch_pdl_small(bias_grad_final_kernel, dim3((nvalid + 127) / 128), dim3(128), 0, s, nvalid, R, bias_off, grads); e_ != cudaSuccess) return e_;
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------ z / KL
__global__ void __launch_bounds__(256)
gauss_bwd_kernel(const float* __restrict__ dz, const float* __restrict__ mu, const float* __restrict__ lv,
                 const float* __restrict__ eps, const float* __restrict__ mu_p, const float* __restrict__ lv_p, int B,
                 int z_dim, int hw, float klw, int bs, __nv_bfloat16* __restrict__ dpost,
                 __nv_bfloat16* __restrict__ dprior) {
  pdl_entry();
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;  // (m, zc in 0..63)
  const size_t total = static_cast<size_t>(B) * hw * 64;
  if (i >= total) return;
  const int zc = static_cast<int>(i & 63);
  const size_t m = i >> 6;
  float a0 = 0.f, a1 = 0.f, p0 = 0.f, p1 = 0.f;
  if (zc < z_dim) {
    const size_t b = m / hw;
    const int pos = static_cast<int>(m % hw);
    const size_t q = (b * z_dim + zc) * hw + pos;
    const float m1 = mu[q], l1 = lv[q], m2 = mu_p[q], l2 = lv_p[q], e = eps[q];
    const float g = dz ? dz[m * 64 + zc] : 0.f;
    const float inv2 = expf(-l2), e1 = expf(l1), d = m1 - m2;
    const float s = klw / bs;
    a0 = g + s * d * inv2;                                   // d/d mu (posterior)
    a1 = g * e * 0.5f * expf(0.5f * l1) + s * (-0.5f + 0.5f * e1 * inv2);  // d/d logvar (posterior)
    p0 = -s * d * inv2;                                      // d/d mu_p (prior)
    p1 = s * (0.5f - 0.5f * (e1 + d * d) * inv2);            // d/d logvar_p (prior)
  }
  reinterpret_cast<uint32_t*>(dpost)[i] = pack_bf16x2(a0, a1);
  reinterpret_cast<uint32_t*>(dprior)[i] = pack_bf16x2(p0, p1);
}
cudaError_t launch_gauss_bwd(const float* dz, const float* mu, const float* lv, const float* eps, const float* mu_p,
                             const float* lv_p, int B, int z_dim, int hw, float kl_weight, int bs,
                             __nv_bfloat16* dpost, __nv_bfloat16* dprior, cudaStream_t s) {
  const size_t total = static_cast<size_t>(B) * hw * 64;
  if (cudaError_t e_ = launch_pdl_small(gauss_bwd_kernel, dim3(static_cast<unsigned>((total + 255) / 256)), dim3(256), 0, s, dz, mu, lv, eps, mu_p, lv_p, B, z_dim, hw,
                                                                              kl_weight, bs, dpost, dprior); e_ != cudaSuccess) return e_;
  return cudaGetLastError();
}

// Step API (train-mode SVGConvModel.forward under torch autograd): the KL term is part of the caller's graph, so its
// gradients w.r.t. (mu, logvar, mu_p, logvar_p) arrive as tensors (B, z, hw) and are added to the z-sample path
__global__ void __launch_bounds__(256)
gauss_bwd_ext_kernel(const float* __restrict__ dz, const float* __restrict__ lv, const float* __restrict__ eps,
                     const float* __restrict__ dmu, const float* __restrict__ dlv, const float* __restrict__ dmu_p,
                     const float* __restrict__ dlv_p, int B, int z_dim, int hw, __nv_bfloat16* __restrict__ dpost,
                     __nv_bfloat16* __restrict__ dprior) {
  pdl_entry();
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;  // (m, zc in 0..63)
  const size_t total = static_cast<size_t>(B) * hw * 64;
  if (i >= total) return;
  const int zc = static_cast<int>(i & 63);
  const size_t m = i >> 6;
  float a0 = 0.f, a1 = 0.f, p0 = 0.f, p1 = 0.f;
  if (zc < z_dim) {
    const size_t b = m / hw;
    const int pos = static_cast<int>(m % hw);
    const size_t q = (b * z_dim + zc) * hw + pos;
    const float g = dz ? dz[m * 64 + zc] : 0.f;
    a0 = g + (dmu ? dmu[q] : 0.f);
    a1 = g * eps[q] * 0.5f * expf(0.5f * lv[q]) + (dlv ? dlv[q] : 0.f);
    p0 = dmu_p ? dmu_p[q] : 0.f;
    p1 = dlv_p ? dlv_p[q] : 0.f;
  }
  reinterpret_cast<uint32_t*>(dpost)[i] = pack_bf16x2(a0, a1);
  reinterpret_cast<uint32_t*>(dprior)[i] = pack_bf16x2(p0, p1);
}
cudaError_t launch_gauss_bwd_ext(const float* dz, const float* lv, const float* eps, const float* dmu, const float* dlv,
                                 const float* dmu_p, const float* dlv_p, int B, int z_dim, int hw, __nv_bfloat16* dpost,
                                 __nv_bfloat16* dprior, cudaStream_t s) {
  const size_t total = static_cast<size_t>(B) * hw * 64;
  if (cudaError_t e_ = launch_pdl_small(gauss_bwd_ext_kernel, dim3(static_cast<unsigned>((total + 255) / 256)), dim3(256), 0, s, dz, lv, eps, dmu, dlv, dmu_p, dlv_p, B,
                                                                                  z_dim, hw, dpost, dprior); e_ != cudaSuccess) return e_;
  return cudaGetLastError();
}

// Step API: dL/d(x_pred) (B, 4, HW; x_pred = sigmoid(logits), dynamics.py:640-642) -> the frame head's bf16 gradient
// operand [B * HW][64] (4 real columns)
__global__ void __launch_bounds__(256)
sigmoid_bwd_kernel(const float* __restrict__ x4, const float* __restrict__ dx4, int B, int HW,
                   __nv_bfloat16* __restrict__ dlogit) {
  pdl_entry();
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;  // (b, p)
  if (i >= static_cast<size_t>(B) * HW) return;
  const size_t b = i / HW;
  const int p = static_cast<int>(i % HW);
  float dl[4];
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const size_t q = (b * 4 + c) * HW + p;
    const float y = x4[q];
    dl[c] = (dx4 ? dx4[q] : 0.f) * y * (1.f - y);
  }
  uint4* d = reinterpret_cast<uint4*>(dlogit + i * 64);
  d[0] = make_uint4(pack_bf16x2(dl[0], dl[1]), pack_bf16x2(dl[2], dl[3]), 0u, 0u);
#pragma unroll
  for (int q = 1; q < 8; ++q) d[q] = make_uint4(0u, 0u, 0u, 0u);
}
cudaError_t launch_sigmoid_bwd(const float* x4, const float* dx4, __nv_bfloat16* dlogit, int B, int HW, cudaStream_t s) {
  const size_t total = static_cast<size_t>(B) * HW;
  if (cudaError_t e_ = launch_pdl_small(sigmoid_bwd_kernel, dim3(static_cast<unsigned>((total + 255) / 256)), dim3(256), 0, s, x4, dx4, B, HW, dlogit); e_ != cudaSuccess) return e_;
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------ frame loss
__global__ void __launch_bounds__(1024)
frame_loss_kernel(const float* __restrict__ x4, const float* __restrict__ xj, const float* __restrict__ xi,
                  const float* __restrict__ mask, int kind, float rw, int B, int HW, float* __restrict__ loss_out,
                  __nv_bfloat16* __restrict__ dlogit, const float* __restrict__ gp_in, float* __restrict__ gxj_out,
                  const float* __restrict__ batch_weight, int Bdiv) {
  pdl_entry();
  __shared__ double sh[32];
  __shared__ float s_scale;
  const int b = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  auto block_sum = [&](double v) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if (lane == 0) sh[warp] = v;
    __syncthreads();
    double t = 0.0;
    for (int i = 0; i < 32; ++i) t += sh[i];
    return t;
  };
  // kind: 0 l1, 1 dontcare_l1, 2 mse (nn.MSELoss), 3 dontcare_mse (trainer.py:149-161; losses.py:11-50). batch_weight
  // (movement weighting, trainer.py:426-429) multiplies the per-sample term of the two l1 kinds only, as the reference
  const bool dontcare = (kind & 1) != 0, squared = kind >= 2;
  // B = samples of this launch (one CTA each; several time steps when batched), Bdiv = batch size of ONE step: every
  // step's loss is a mean over its own batch
  const float bw = (batch_weight && !squared) ? batch_weight[b % Bdiv] : 1.f;
  B = Bdiv;
  if (dontcare) {
    double cnt = 0.0;
    for (int p = tid; p < HW; p += blockDim.x) cnt += mask[static_cast<size_t>(b) * HW + p] != 0.f ? 0.0 : 3.0;
    cnt = block_sum(cnt);
    if (tid == 0) s_scale = static_cast<float>(static_cast<double>(bw) / ((cnt + 1.0) * B));
  } else {
    if (tid == 0) s_scale = bw / (static_cast<float>(B) * 3.f * HW);
  }
  __syncthreads();
  const float scale = s_scale;
  double acc = 0.0;
  for (int p = tid; p < HW; p += blockDim.x) {
    const float mh = x4[(static_cast<size_t>(b) * 4 + 3) * HW + p];
    const bool robot = dontcare && mask[static_cast<size_t>(b) * HW + p] != 0.f;
    const float w = robot ? rw : 1.f;
    float dm = 0.f, dl[4];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float xh = x4[(static_cast<size_t>(b) * 4 + c) * HW + p];
      const float j = xj[(static_cast<size_t>(b) * 3 + c) * HW + p];
      const float t = xi[(static_cast<size_t>(b) * 3 + c) * HW + p];
      const float pr = (1.f - mh) * j + mh * xh;  // blends with the un-blacked x_j (trainer.py:406-407)
      const float diff = (t - pr) * w;
      acc += squared ? diff * diff : fabsf(diff);
      const float sg = squared ? 2.f * diff : (diff > 0.f ? 1.f : (diff < 0.f ? -1.f : 0.f));
      float dp = -sg * w * scale;                 // dL / d pred
      if (gp_in) dp += gp_in[(static_cast<size_t>(b) * 3 + c) * HW + p];
      if (gxj_out) gxj_out[(static_cast<size_t>(b) * 3 + c) * HW + p] = dp * (1.f - mh);
      dm += dp * (xh - j);
      dl[c] = dp * mh * xh * (1.f - xh);
    }
    dl[3] = dm * mh * (1.f - mh);
    uint4* d = reinterpret_cast<uint4*>(dlogit + (static_cast<size_t>(b) * HW + p) * 64);
    d[0] = make_uint4(pack_bf16x2(dl[0], dl[1]), pack_bf16x2(dl[2], dl[3]), 0u, 0u);
#pragma unroll
    for (int q = 1; q < 8; ++q) d[q] = make_uint4(0u, 0u, 0u, 0u);  // the operand is 64 channels wide, 4 are real
  }
  acc = block_sum(acc);
  if (tid == 0) loss_out[b] = static_cast<float>(acc * scale);
}
cudaError_t launch_frame_loss(const float* x4, const float* xj, const float* xi, const float* mask, int kind,
                              float robot_weight, int B, int HW, float* loss_out, __nv_bfloat16* dlogit,
                              const float* gp_in, float* gxj_out, cudaStream_t s, const float* batch_weight, int Bdiv) {
  if ((kind & 1) && !mask) return cudaErrorInvalidValue;
  if (cudaError_t e_ = launch_pdl_small(frame_loss_kernel, dim3(B), dim3(1024), 0, s, x4, xj, xi, mask, kind, robot_weight, B, HW, loss_out, dlogit, gp_in, gxj_out,
                                       batch_weight, Bdiv > 0 ? Bdiv : B); e_ != cudaSuccess) return e_;
  return cudaGetLastError();
}

__global__ void __launch_bounds__(256)
composite_kernel(const float* __restrict__ x4, const float* __restrict__ xj, float* __restrict__ xp, int B, int HW) {
  pdl_entry();
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;  // (b, c, p)
  if (i >= static_cast<size_t>(B) * 3 * HW) return;
  const int p = static_cast<int>(i % HW);
  const int c = static_cast<int>((i / HW) % 3);
  const size_t b = i / (static_cast<size_t>(HW) * 3);
  const float mh = x4[(b * 4 + 3) * HW + p];
  xp[i] = (1.f - mh) * xj[i] + mh * x4[(b * 4 + c) * HW + p];
}
cudaError_t launch_composite(const float* x4, const float* xj, float* xp, int B, int HW, cudaStream_t s) {
  const size_t total = static_cast<size_t>(B) * 3 * HW;
  if (cudaError_t e_ = launch_pdl_small(composite_kernel, dim3(static_cast<unsigned>((total + 255) / 256)), dim3(256), 0, s, x4, xj, xp, B, HW); e_ != cudaSuccess) return e_;
  return cudaGetLastError();
}

// out[p] = sum_tap in[p + off(tap)] * w[tap]  =>  d in[q] = sum_tap d out[q - off(tap)] * w[tap]
__global__ void __launch_bounds__(256)
first_dgrad_kernel(const float* __restrict__ draw, const float* __restrict__ wf, int cin,
                   const float* __restrict__ zero_mask, float* __restrict__ gimg, int B, int H, int W) {
  pdl_entry();
  __shared__ float sw[9 * 3 * 64];
  for (int i = threadIdx.x; i < 9 * 3 * 64; i += blockDim.x) {
    const int o = i & 63, c = (i >> 6) % 3, tap = (i >> 6) / 3;
    sw[i] = wf[(tap * cin + c) * 64 + o];
  }
  __syncthreads();
  const size_t pix = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const size_t total = static_cast<size_t>(B) * H * W;
  if (pix >= total) return;
  const int x = static_cast<int>(pix % W), y = static_cast<int>((pix / W) % H);
  const size_t b = pix / (static_cast<size_t>(W) * H);
  if (zero_mask && zero_mask[pix] != 0.f) return;
  float acc[3] = {0.f, 0.f, 0.f};
  for (int tap = 0; tap < 9; ++tap) {
    const int yy = y - (tap / 3 - 1), xx = x - (tap % 3 - 1);
    if (yy < 0 || yy >= H || xx < 0 || xx >= W) continue;
    const float* d = draw + ((b * H + yy) * W + xx) * 64;
    const float* w0 = &sw[(tap * 3) * 64];
    for (int o = 0; o < 64; ++o) {
      const float dv = d[o];
      acc[0] += dv * w0[o];
      acc[1] += dv * w0[64 + o];
      acc[2] += dv * w0[128 + o];
    }
  }
  const size_t hw = static_cast<size_t>(H) * W, pos = static_cast<size_t>(y) * W + x;
  for (int c = 0; c < 3; ++c) gimg[(b * 3 + c) * hw + pos] += acc[c];
}
cudaError_t launch_first_dgrad(const float* draw, const float* wf, int cin, const float* zero_mask, float* gimg, int B,
                               int H, int W, cudaStream_t s) {
  const size_t total = static_cast<size_t>(B) * H * W;
  if (cudaError_t e_ = launch_pdl_small(first_dgrad_kernel, dim3(static_cast<unsigned>((total + 255) / 256)), dim3(256), 0, s, draw, wf, cin, zero_mask, gimg, B, H, W); e_ != cudaSuccess) return e_;
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------ max pool backward
__global__ void __launch_bounds__(256)
pool_bwd_kernel(const __nv_bfloat16* __restrict__ in, int in_cstride, int in_coff, const float* __restrict__ dout,
                int B, int H, int W, int C, float* __restrict__ din, int din_cstride, int din_coff, int accumulate) {
  pdl_entry();
  const int Ho = H / 2, Wo = W / 2;
  const size_t total = static_cast<size_t>(B) * Ho * Wo * C;
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int c = static_cast<int>(i % C);
  const int xo = static_cast<int>((i / C) % Wo);
  const int yo = static_cast<int>((i / (static_cast<size_t>(C) * Wo)) % Ho);
  const size_t b = i / (static_cast<size_t>(C) * Wo * Ho);
  size_t pos[4];
  float v[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    pos[k] = (b * H + 2 * yo + (k >> 1)) * W + 2 * xo + (k & 1);
    v[k] = __bfloat162float(in[pos[k] * in_cstride + in_coff + c]);
  }
  int best = 0;
#pragma unroll
  for (int k = 1; k < 4; ++k)
    if (v[k] > v[best]) best = k;  // first maximum in scan order
  const float g = dout[i];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    float* d = din + pos[k] * din_cstride + din_coff + c;
    if (accumulate) {
      if (k == best) *d += g;
    } else {
      *d = (k == best) ? g : 0.f;
    }
  }
}
cudaError_t launch_pool_bwd(const __nv_bfloat16* in, int in_cstride, int in_coff, const float* dout, int B, int H,
                            int W, int C, float* din, int din_cstride, int din_coff, int accumulate, cudaStream_t s) {
  const size_t total = static_cast<size_t>(B) * (H / 2) * (W / 2) * C;
  if (cudaError_t e_ = launch_pdl_small(pool_bwd_kernel, dim3(static_cast<unsigned>((total + 255) / 256)), dim3(256), 0, s, in, in_cstride, in_coff, dout, B, H, W, C,
                                                                             din, din_cstride, din_coff, accumulate); e_ != cudaSuccess) return e_;
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------ layout helpers
__global__ void __launch_bounds__(256) cast_bf16_kernel(const float* __restrict__ src, long long n,
                                                        __nv_bfloat16* __restrict__ dst) {
  pdl_entry();
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = __float2bfloat16(src[i]);
}
cudaError_t launch_cast_bf16(const float* src, long long n, __nv_bfloat16* dst, cudaStream_t s) {
  if (cudaError_t e_ = launch_pdl_small(cast_bf16_kernel, dim3(static_cast<unsigned>((n + 255) / 256)), dim3(256), 0, s, src, n, dst); e_ != cudaSuccess) return e_;
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------ optimiser, noise
// torch.optim.Adam (no weight decay / amsgrad): 7 fp32 streams (p, g, m, v read; p, m, v written), 128-bit accesses,
// grid-stride over a grid of a few CTAs per SM, streaming cache hints (nothing is re-read before the next step)
__global__ void __launch_bounds__(512)
adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
            long long n, float lr, float b1, float b2, float eps, float bc1, float bc2_sqrt, float gscale, int head) {
  pdl_entry();
  const float step = lr / bc1;
  // gscale: 1 / world size after a SUM all-reduce (x * 1.0f is exact, so the single-GPU update is unchanged)
  auto upd = [&](float& pi, float gi, float& mi, float& vi) {
    adam_update(pi, gi * gscale, mi, vi, b1, b2, step, bc2_sqrt, eps);
  };
  // `head` scalar elements bring the four streams to a 16-byte boundary (a range of the flat buffers may start anywhere)
  if (blockIdx.x == 0 && static_cast<int>(threadIdx.x) < head) {
    const int i = threadIdx.x;
    float pi = p[i], mi = m[i], vi = v[i];
    upd(pi, g[i], mi, vi);
    m[i] = mi; v[i] = vi; p[i] = pi;
  }
  p += head; g += head; m += head; v += head; n -= head;
  const long long n4 = n >> 2;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 pi = __ldcs(reinterpret_cast<const float4*>(p) + i);
    const float4 gi = __ldcs(reinterpret_cast<const float4*>(g) + i);
    float4 mi = __ldcs(reinterpret_cast<const float4*>(m) + i);
    float4 vi = __ldcs(reinterpret_cast<const float4*>(v) + i);
    upd(pi.x, gi.x, mi.x, vi.x);
    upd(pi.y, gi.y, mi.y, vi.y);
    upd(pi.z, gi.z, mi.z, vi.z);
    upd(pi.w, gi.w, mi.w, vi.w);
    __stcs(reinterpret_cast<float4*>(m) + i, mi);
    __stcs(reinterpret_cast<float4*>(v) + i, vi);
    __stcs(reinterpret_cast<float4*>(p) + i, pi);
  }
  if (blockIdx.x == 0 && threadIdx.x >= 32 && threadIdx.x - 32 < (n & 3)) {  // tail (another warp than the head)
    const long long i = (n4 << 2) + (threadIdx.x - 32);
    float pi = p[i], mi = m[i], vi = v[i];
    upd(pi, g[i], mi, vi);
    m[i] = mi; v[i] = vi; p[i] = pi;
  }
}
cudaError_t launch_adam(float* p, const float* g, float* m, float* v, long long n, float lr, float b1, float b2,
                        float eps, int t, cudaStream_t s, float grad_scale) {
  const float bc1 = 1.f - powf(b1, static_cast<float>(t));
  const float bc2 = sqrtf(1.f - powf(b2, static_cast<float>(t)));
  if (n <= 0) return cudaSuccess;
  // the four streams must be aligned alike (ranges of buffers that are themselves 16-byte aligned are)
  const uintptr_t a = reinterpret_cast<uintptr_t>(p) & 15;
  if ((reinterpret_cast<uintptr_t>(g) & 15) != a || (reinterpret_cast<uintptr_t>(m) & 15) != a ||
      (reinterpret_cast<uintptr_t>(v) & 15) != a || (a & 3))
    return cudaErrorInvalidValue;
  int head = static_cast<int>(((16 - a) & 15) >> 2);
  if (head > n) head = static_cast<int>(n);
  const long long body = (n - head) >> 2;
  const int blocks = static_cast<int>(std::min<long long>(148 * 8, std::max<long long>(1, (body + 511) / 512)));
  if (cudaError_t e_ = launch_pdl_small(adam_kernel, dim3(blocks), dim3(512), 0, s, p, g, m, v, n, lr, b1, b2, eps, bc1, bc2, grad_scale, head); e_ != cudaSuccess) return e_;
  return cudaGetLastError();
}

__global__ void __launch_bounds__(256)
normal_fill_kernel(float* __restrict__ dst, long long n, unsigned long long seed, unsigned int ctr) {
  pdl_entry();
  const long long q = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;  // 4 values per thread
  if (q * 4 >= n) return;
  const Philox4 r = philox4x32_10(static_cast<uint32_t>(q), static_cast<uint32_t>(q >> 32), ctr, 0x7a1u,
                                  static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32));
  float z[4];
  box_muller(r.v[0], r.v[1], z[0], z[1]);
  box_muller(r.v[2], r.v[3], z[2], z[3]);
  for (int k = 0; k < 4 && q * 4 + k < n; ++k) dst[q * 4 + k] = z[k];
}
cudaError_t launch_normal_fill(float* dst, long long n, unsigned long long seed, unsigned int ctr, cudaStream_t s) {
  const long long quads = (n + 3) / 4;
  if (cudaError_t e_ = launch_pdl_small(normal_fill_kernel, dim3(static_cast<unsigned>((quads + 255) / 256)), dim3(256), 0, s, dst, n, seed, ctr); e_ != cudaSuccess) return e_;
  return cudaGetLastError();
}

__global__ void __launch_bounds__(256) sum_f32_kernel(const float* __restrict__ src, int n, float* __restrict__ dst) {
  pdl_entry();
  __shared__ double sh[8];
  double a = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) a += src[i];
  for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = a;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int i = 0; i < 8; ++i) t += sh[i];
    dst[0] += static_cast<float>(t);
  }
}
cudaError_t launch_sum_f32(const float* src, int n, float* dst_accum, cudaStream_t s) {
  if (cudaError_t e_ = launch_pdl_small(sum_f32_kernel, dim3(1), dim3(256), 0, s, src, n, dst_accum); e_ != cudaSuccess) return e_;
  return cudaGetLastError();
}

__global__ void gather_f32_kernel(const float* __restrict__ params, const long long* __restrict__ off, int n,
                                  float* __restrict__ dst) {
  pdl_entry();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = off[i] >= 0 ? params[off[i]] : 0.f;
}
cudaError_t launch_gather_f32(const float* params, const long long* off, int n, float* dst, cudaStream_t s) {
  if (cudaError_t e_ = launch_pdl_small(gather_f32_kernel, dim3((n + 255) / 256), dim3(256), 0, s, params, off, n, dst); e_ != cudaSuccess) return e_;
  return cudaGetLastError();
}

__global__ void __launch_bounds__(256)
img_prep_train_kernel(const float* __restrict__ img, const float* __restrict__ mask, float* __restrict__ img4, int B,
                      int HW) {
  pdl_entry();
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= static_cast<size_t>(B) * HW) return;
  const size_t b = i / HW;
  const int pos = static_cast<int>(i % HW);
  const float* p = img + b * 3 * HW + pos;
  float r = p[0], g = p[HW], bl = p[2 * HW];
  if (mask && mask[i] != 0.f) r = g = bl = 0.f;
  *reinterpret_cast<float4*>(img4 + i * 4) = make_float4(r, g, bl, 0.f);
}
cudaError_t launch_img_prep_train(const float* img_nchw, const float* mask, float* img4, int B, int HW, cudaStream_t s) {
  const size_t total = static_cast<size_t>(B) * HW;
  if (cudaError_t e_ = launch_pdl_small(img_prep_train_kernel, dim3(static_cast<unsigned>((total + 255) / 256)), dim3(256), 0, s, img_nchw, mask, img4, B, HW); e_ != cudaSuccess) return e_;
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------ logged metrics
// robot_mse_criterion / world_mse_criterion (losses.py:52-78) for n samples (several time steps at once): one CTA per
// sample, then a single thread adds the per-sample terms in index order (deterministic). out2[0] += sum_b robot_b / Bdiv.
__global__ void __launch_bounds__(256)
robot_world_mse_part_kernel(const float* __restrict__ p, const float* __restrict__ t, const float* __restrict__ mask,
                            float* __restrict__ part, int HW) {
  pdl_entry();
  __shared__ double sh[3][8];
  const int b = blockIdx.x;
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
  double v[3] = {sr, sw, cr};
  for (int k = 0; k < 3; ++k) {
    for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], o);
    if ((threadIdx.x & 31) == 0) sh[k][threadIdx.x >> 5] = v[k];
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a[3] = {0.0, 0.0, 0.0};
    for (int k = 0; k < 3; ++k)
      for (int w = 0; w < 8; ++w) a[k] += sh[k][w];
    part[2 * b] = static_cast<float>(a[0] / (a[2] + 1.0));
    part[2 * b + 1] = static_cast<float>(a[1] / (3.0 * HW - a[2] + 1.0));
  }
}
__global__ void metric_fold_kernel(const float* __restrict__ part, int n, int width, double scale, float* __restrict__ out) {
  pdl_entry();
  if (threadIdx.x < width && blockIdx.x == 0) {
    double s = 0.0;
    for (int i = 0; i < n; ++i) s += static_cast<double>(part[i * width + threadIdx.x]);
    out[threadIdx.x] += static_cast<float>(s * scale);
  }
}
cudaError_t launch_robot_world_mse_batched(const float* pred, const float* target, const float* mask, float* part,
                                           float* out2, int n, int Bdiv, int HW, cudaStream_t s) {
  if (cudaError_t e_ = launch_pdl_small(robot_world_mse_part_kernel, dim3(n), dim3(256), 0, s, pred, target, mask, part, HW); e_ != cudaSuccess) return e_;
  if (cudaError_t e_ = launch_pdl_small(metric_fold_kernel, dim3(1), dim3(32), 0, s, part, n, 2, 1.0 / Bdiv, out2); e_ != cudaSuccess) return e_;
  return cudaGetLastError();
}

// KL(N(mu1, s1) || N(mu2, s2)) summed over everything / bs (losses.py:97-106), n elements (several steps at once):
// 64 CTAs of partial sums, folded in order; out[0] += sum / bs
__global__ void __launch_bounds__(256)
kl_part_kernel(const float* __restrict__ mu1, const float* __restrict__ lv1, const float* __restrict__ mu2,
               const float* __restrict__ lv2, float* __restrict__ part, long long n) {
  pdl_entry();
  __shared__ double sh[8];
  double acc = 0.0;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float s1 = expf(0.5f * lv1[i]), s2 = expf(0.5f * lv2[i]);
    const float dm = mu1[i] - mu2[i];
    acc += static_cast<double>(logf(s2 / s1) + (expf(lv1[i]) + dm * dm) / (2.f * expf(lv2[i])) - 0.5f);
  }
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += sh[w];
    part[blockIdx.x] = static_cast<float>(t);
  }
}
cudaError_t launch_kl_loss_batched(const float* mu1, const float* lv1, const float* mu2, const float* lv2, float* part,
                                   float* out_accum, long long n, int bs, cudaStream_t s) {
  if (cudaError_t e_ = launch_pdl_small(kl_part_kernel, dim3(64), dim3(256), 0, s, mu1, lv1, mu2, lv2, part, n); e_ != cudaSuccess) return e_;
  if (cudaError_t e_ = launch_pdl_small(metric_fold_kernel, dim3(1), dim3(32), 0, s, part, 64, 1, 1.0 / bs, out_accum); e_ != cudaSuccess) return e_;
  return cudaGetLastError();
}

}  // namespace rac
