// encoder.c1.0 on the tensor cores (inference): 3x3 conv over [rgb | mask_t | mask_t+1] (K = 27 / 36 / 45) -> 64
// channels, folded eval BatchNorm, LeakyReLU(0.2) (reference vgg_64.py:99-101 via vgg_layer :8-18).
//
// K is far too small for a TMA-fed implicit GEMM (a k-block is 64 channels per tap), and the CUDA-core version
// (misc_kernels.cu::first_conv_kernel) is FMA-bound: 10.6 / 17.7 GFMA per 2000 candidates = 0.85 / 1.32 ms next to a
// 0.13 ms HBM write stream. Here the im2col row of a pixel (<= 45 values, zero-padded to K = 64) is gathered by its
// own thread straight into the 128B-swizzled K-major layout that tcgen05.mma reads -- 6 x 16-byte stores per pixel
// instead of 27..45 x 64 FMAs -- and the product with the resident [64 x 64] bf16 weight tile is 6 MMAs
// (M128 x N64 x K16, 2 sub-tiles x 3 k-steps) per 256-pixel tile. The same 8 warps then run the activation epilogue
// of the previous tile while the MMAs of the current one execute (two shared-memory stages, two TMEM stages; all
// hazards are covered by program order + the two mbarriers, see the loop). Inputs are rounded to bf16 here (pixel
// values in [0, 1]: <= 2^-9 absolute), like every other activation of the network.
#include "conv.cuh"
#include "epilogue.cuh"
#include "ptx.cuh"
#include "misc_kernels.cuh"

namespace rac {

namespace {

constexpr int kFcTile = 256;                    // pixels per tile: 4 image rows x 64 columns
constexpr int kFcABytes = kFcTile * 128;        // 32 KB per stage
constexpr int kFcWBytes = 64 * 128;             // 8 KB
constexpr int kFcThreads = 256 + 32;            // 8 worker warps + 1 MMA / TMEM warp
constexpr int kFcSmem = 1024 + 2 * kFcABytes + kFcWBytes + 256;
constexpr int kFcTmemCols = 256;                // 2 stages x 2 sub-tiles x 64 columns

template <int CIN>
__global__ void __launch_bounds__(kFcThreads, 2)
first_conv_tc_kernel(const float* __restrict__ img4, const float* __restrict__ mask_a, const float* __restrict__ mask_b,
                     long long mask_bstride, const float* __restrict__ w, const ConvGeom g, const EpiParams e, int B) {
  constexpr int K = 9 * CIN;                    // 27 / 36 / 45
  constexpr int kChunksK = (K + 7) / 8;         // 16-byte chunks of a row that carry data (4 / 5 / 6)
  constexpr int kSteps = (K + 15) / 16;         // k16 MMA steps (2 / 3 / 3)
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* s_a = smem;                          // 2 stages
  uint8_t* s_w = smem + 2 * kFcABytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(s_w + kFcWBytes);
  uint64_t* tmem_full = full_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int H = g.H, W = g.W;
  const int tiles_per_img = H / 4;
  const int num_tiles = B * tiles_per_img;

  // ---- one-time setup: barriers, TMEM, weights (fp32 [K][64] -> bf16 [64 n][64 k] swizzled), zero padding chunks
  if (threadIdx.x == 256) {
    mbar_init(&full_bar[0], 256); mbar_init(&full_bar[1], 256);
    mbar_init(&tmem_full[0], 1); mbar_init(&tmem_full[1], 1);
    fence_barrier_init();
  }
  if (warp == 8) {
    tmem_alloc(tmem_slot, kFcTmemCols);
    tmem_relinquish();
  }
  for (int i = threadIdx.x; i < (2 * kFcABytes + kFcWBytes) / 16; i += blockDim.x)
    reinterpret_cast<uint4*>(smem)[i] = make_uint4(0u, 0u, 0u, 0u);
  __syncthreads();
  for (int i = threadIdx.x; i < K * 64; i += blockDim.x) {
    const int k = i >> 6, n = i & 63;           // w[k][n]
    const uint32_t off = n * 128 + (((k >> 3) ^ (n & 7)) << 4) + (k & 7) * 2;
    *reinterpret_cast<__nv_bfloat16*>(s_w + off) = __float2bfloat16(w[i]);
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 8) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(64);
      const uint64_t bdesc = umma_desc_sw128(smem_u32(s_w));
      int it = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
        const int s = it & 1;
        mbar_wait(&full_bar[s], (it >> 1) & 1);
        tc_fence_after();
        const uint64_t adesc = umma_desc_sw128(smem_u32(s_a + s * kFcABytes));
#pragma unroll
        for (int k = 0; k < kSteps; ++k)
#pragma unroll
          for (int sub = 0; sub < 2; ++sub)
            umma_bf16_ss(tmem_base + s * 128 + sub * 64, adesc + 2 * k + sub * (128 * 128 / 16), bdesc + 2 * k, idesc,
                         k != 0 ? 1u : 0u);
        umma_commit(&tmem_full[s]);
      }
    }
  } else {
    // ===================== workers: im2col build of tile i+1, epilogue of tile i =====================
    const int r = threadIdx.x;                  // pixel inside the tile = MMA row
    const int wq = warp & 3, sub = warp >> 2;
    auto build = [&](int tile, int s) {
      const int b = tile / tiles_per_img;
      const int y = (tile - b * tiles_per_img) * 4 + (r >> 6);
      const int x = r & 63;
      float v[kChunksK * 8];
#pragma unroll
      for (int i = 0; i < kChunksK * 8; ++i) v[i] = 0.f;
#pragma unroll
      for (int kh = 0; kh < 3; ++kh)
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) {
          const int yy = y + kh - 1, xx = x + kw - 1;
          if (yy < 0 || yy >= H || xx < 0 || xx >= W) continue;
          const size_t q = (static_cast<size_t>(b) * H + yy) * W + xx;
          const float4 px = __ldg(reinterpret_cast<const float4*>(img4) + q);
          const int k0 = (kh * 3 + kw) * CIN;
          v[k0] = px.x; v[k0 + 1] = px.y; v[k0 + 2] = px.z;
          if constexpr (CIN > 3) v[k0 + 3] = __ldg(mask_a + static_cast<size_t>(b) * mask_bstride + static_cast<size_t>(yy) * W + xx);
          if constexpr (CIN > 4) v[k0 + 4] = __ldg(mask_b + static_cast<size_t>(b) * mask_bstride + static_cast<size_t>(yy) * W + xx);
        }
      uint8_t* row = s_a + s * kFcABytes + r * 128;
#pragma unroll
      for (int j = 0; j < kChunksK; ++j)
        *reinterpret_cast<uint4*>(row + ((j ^ (r & 7)) << 4)) =
            make_uint4(pack_bf16x2(v[8 * j], v[8 * j + 1]), pack_bf16x2(v[8 * j + 2], v[8 * j + 3]),
                       pack_bf16x2(v[8 * j + 4], v[8 * j + 5]), pack_bf16x2(v[8 * j + 6], v[8 * j + 7]));
      fence_proxy_async();                      // generic-proxy stores -> visible to the tensor core (async proxy)
      mbar_arrive(&full_bar[s]);
    };
    int it = 0;
    int tile = blockIdx.x;
    if (tile < num_tiles) build(tile, 0);
    for (; tile < num_tiles; tile += gridDim.x, ++it) {
      const int s = it & 1;
      const int next = tile + gridDim.x;
      // stage s^1 was read by the MMAs of tile it-1, whose completion this thread observed in the previous
      // iteration (tmem_full wait); TMEM stage s^1 was drained by this thread's epilogue of tile it-1
      if (next < num_tiles) build(next, s ^ 1);
      mbar_wait(&tmem_full[s], (it >> 1) & 1);
      tc_fence_after();
      const int b = tile / tiles_per_img;
      const int y = (tile - b * tiles_per_img) * 4 + (r >> 6);
      const int x = r & 63;
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(wq * 32) << 16) + s * 128 + sub * 64;
      float acc[2][32];
      tmem_ld32(t_row, acc[0]);
      tmem_ld32(t_row + 32, acc[1]);
      tmem_ld_wait();
      // this kernel is bound by its output stores: transpose inside 8-lane groups so that every store instruction
      // writes 4 complete 128-byte rows instead of 16 bytes of 32 different rows
      const int ty0 = (tile - b * tiles_per_img) * 4;
      auto row_of = [&](int rr, int& ob, int& oy, int& ox, bool& ov) {
        ob = b; oy = ty0 + (rr >> 6); ox = rr & 63; ov = true;
      };
      uint4 packed[8];
      act_pack32(e, 0, acc[0], packed);
      act_pack32(e, 32, acc[1], packed + 4);
      epi_act_store64(g, e, r, 0, packed, row_of);
      (void)y; (void)x;
      tc_fence_before();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) tmem_dealloc(tmem_base, kFcTmemCols);
}

}  // namespace

cudaError_t first_conv_tc_set_attributes() {
  cudaError_t err;
  if ((err = cudaFuncSetAttribute(first_conv_tc_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, kFcSmem)) != cudaSuccess) return err;
  if ((err = cudaFuncSetAttribute(first_conv_tc_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, kFcSmem)) != cudaSuccess) return err;
  return cudaFuncSetAttribute(first_conv_tc_kernel<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, kFcSmem);
}

cudaError_t launch_first_conv_tc(const float* img4, const float* mask_a, const float* mask_b, long long mask_bstride,
                                 const float* w, const float* bias, __nv_bfloat16* out, int B, int H, int W, int cin,
                                 int num_sms, cudaStream_t s) {
  if (cin < 3 || cin > 5 || W != 64 || H % 4 != 0) return cudaErrorInvalidValue;
  if ((cin > 3 && !mask_a) || (cin > 4 && !mask_b)) return cudaErrorInvalidValue;
  ConvGeom g{};
  g.B = B; g.H = H; g.W = W;
  EpiParams e{};
  e.bias = bias; e.cout = 64; e.out = out; e.out_cstride = 64; e.out_coff = 0; e.upsample = 0; e.lrelu = 1;
  const int tiles = B * (H / 4);
  int grid = 2 * num_sms;
  if (grid > tiles) grid = tiles;
  if (grid < 1) return cudaSuccess;
  if (cin == 3)
    first_conv_tc_kernel<3><<<grid, kFcThreads, kFcSmem, s>>>(img4, mask_a, mask_b, mask_bstride, w, g, e, B);
  else if (cin == 4)
    first_conv_tc_kernel<4><<<grid, kFcThreads, kFcSmem, s>>>(img4, mask_a, mask_b, mask_bstride, w, g, e, B);
  else
    first_conv_tc_kernel<5><<<grid, kFcThreads, kFcSmem, s>>>(img4, mask_a, mask_b, mask_bstride, w, g, e, B);
  return cudaGetLastError();
}

}  // namespace rac
