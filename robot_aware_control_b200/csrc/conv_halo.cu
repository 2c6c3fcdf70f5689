// Halo-tile 3x3 convolution for the full-resolution 64-wide layers (encoder.c1.1, decoder.upc4.1 / upc5.0 / upc5.1:
// N = 64 or 4 output channels, 64 or 128 input channels, 48x64 or 24x32 maps). With N <= 64 the generic kernel
// (conv_tc.cu) re-reads every activation row once per filter tap from L2 -- 9 x 128 B per 64 x 64 MACs -- and runs at
// the L2 -> SM bandwidth (~9 TB/s chip-wide, 340-550 TFLOP/s). Here one TMA load brings a halo tile into shared memory
// ONCE and all 9 taps are MMAs on shifted views of it, which is possible because of the row order chosen for the tile:
//
//   tile   = 6 output rows x 32 output columns of one candidate; input halo = 8 rows x 34 columns
//   smem   = the halo tile, COLUMN-major: row index (x' + 1) * 8 + yy, 128 B (64 channels) per row, 128B-swizzled by
//            TMA -> every column is exactly one 1024-byte swizzle atom (8 rows)
//   M rows = 256 = 32 columns x 8 halo rows: MMA row m = xo * 8 + yo stands for output pixel (x0 + xo, y0 - 1 + yo);
//            the A row it needs for tap (kh, kw) is  m + kw * 8 + (kh - 1):  a UNIFORM shift of the whole operand, i.e.
//            the same UMMA descriptor with its start address moved by kw atoms and kh - 1 rows. The start address is
//            then not 1024-byte aligned; measured on B200, the 128B swizzle XOR follows the ABSOLUTE shared-memory
//            address bits [7:9] (as TMA's does), so the descriptor's base-offset field must stay 0 (setting it to the
//            row phase gives wrong results). yo = 0 and 7 are halo-only rows: computed and discarded (75 % useful).
//
// Weights (9 taps x Cin x N, <= 144 KB) stay resident in shared memory for the whole persistent CTA. Warp roles as in
// conv_tc.cu: warp 0 TMA producer, warp 1 MMA issuer, warp 2 TMEM allocator, warps 4..11 epilogue (one thread per
// M row), two TMEM accumulator stages so the epilogue overlaps the next tile.
#include "conv.cuh"
#include "epilogue.cuh"
#include "ptx.cuh"

namespace rac {

template <int CIN_KB, int BLOCK_N>
struct HaloCfg {
  static constexpr int kXT = 32;                              // output columns per tile
  static constexpr int kYT = 6;                               // output rows per tile
  static constexpr int kM = kXT * 8;                          // 256 MMA rows (2 sub-tiles of 128)
  static constexpr int kABytes = (kXT + 2) * 8 * 128;         // 34 atoms = 34816 B per 64-channel k-block
  static constexpr int kWTile = BLOCK_N * 128;                // one (tap, k-block) weight tile
  static constexpr int kWBytes = 9 * CIN_KB * kWTile;
  static constexpr int kWBytesPad = (kWBytes + 1023) / 1024 * 1024;
  static constexpr int kBarBytes = 1024;
  static constexpr int kBudget = 227 * 1024 - 1024 /*alignment slack*/ - kBarBytes - kWBytesPad - 1024 /*tail guard*/;
  static constexpr int kStagesFit = kBudget / kABytes;
  static constexpr int kStages = kStagesFit > 4 ? 4 : kStagesFit;
  static constexpr int kAccCols = 2 * BLOCK_N;                // two 128-row sub-tiles
  static constexpr int kTmemNeed = 2 * kAccCols;
  static constexpr int kTmemCols = kTmemNeed <= 32 ? 32 : (kTmemNeed <= 64 ? 64 : (kTmemNeed <= 128 ? 128 : 256));
  static constexpr int kEpiThreads = kM;
  static constexpr int kThreads = 128 + kEpiThreads;
  static constexpr int kSmemBytes = 1024 + kWBytesPad + kStages * kABytes + 1024 + kBarBytes;
  static_assert(kStages >= 2, "need at least two activation stages");
};

struct HaloGeom {
  int B, H, W;
  int xtiles, ytiles;     // W / 32, H / 6
  int num_tiles;          // B * ytiles * xtiles
  int column_loads;       // 1: one TMA per halo column (fallback when the permuted-stride tensor map is rejected)
  int use_base_offset;    // bring-up switch (RAC_HALO_BASE_OFFSET=1): sets the descriptor base-offset field -- wrong on B200
};

// UMMA descriptor of the halo operand: start address may be any multiple of 128 B inside the stage
__device__ __forceinline__ uint64_t umma_desc_sw128_shifted(uint32_t smem_addr, int use_base_offset) {
  uint64_t d = umma_desc_sw128(smem_addr);
  if (use_base_offset) d |= static_cast<uint64_t>((smem_addr >> 7) & 7) << 49;
  return d;
}

template <int CIN_KB, int BLOCK_N, int EPI>
__global__ void __launch_bounds__(HaloCfg<CIN_KB, BLOCK_N>::kThreads, 1)
conv_halo_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_w, const HaloGeom hg,
                 const ConvGeom g, const EpiParams e) {
  using Cfg = HaloCfg<CIN_KB, BLOCK_N>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* s_w = smem;
  uint8_t* s_a = smem + Cfg::kWBytesPad;
  uint8_t* bar_base = s_a + Cfg::kStages * Cfg::kABytes + 1024;  // 1 KB guard: the last tap reads one row past a stage
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(bar_base);
  uint64_t* empty_bar = full_bar + Cfg::kStages;
  uint64_t* tmem_full = empty_bar + Cfg::kStages;
  uint64_t* tmem_empty = tmem_full + 2;
  uint64_t* w_full = tmem_empty + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(w_full + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_a);
    tma_prefetch_desc(&tm_w);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < Cfg::kStages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], Cfg::kEpiThreads);
    }
    mbar_init(w_full, 1);
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, Cfg::kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  griddep_launch_dependents();  // (PDL, ptx.cuh) the next kernel's prologue may start
  griddep_wait();               // everything below reads / writes global memory of earlier kernels
  const int tiles_per_img = hg.ytiles * hg.xtiles;

  if (warp == 0 && lane == 0) {
    // ===================== TMA producer =====================
    mbar_arrive_expect_tx(w_full, Cfg::kWBytes);
    for (int i = 0; i < 9 * CIN_KB; ++i) tma_load_2d(&tm_w, w_full, s_w + i * Cfg::kWTile, i * kBlockK, 0);
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < hg.num_tiles; tile += gridDim.x) {
      const int b = tile / tiles_per_img;
      const int t = tile - b * tiles_per_img;
      const int y0 = (t / hg.xtiles) * Cfg::kYT;
      const int x0 = (t % hg.xtiles) * Cfg::kXT;
      for (int kb = 0; kb < CIN_KB; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        uint8_t* sa = s_a + stage * Cfg::kABytes;
        mbar_arrive_expect_tx(&full_bar[stage], Cfg::kABytes);
        if (!hg.column_loads) {
          // tensor map dims (C, H, W, B): box {64, 8, 34, 1} lands column-major, zero fill outside the image
          tma_load_4d(&tm_a, &full_bar[stage], sa, kb * kBlockK, y0 - 1, x0 - 1, b);
        } else {
          // tensor map dims (C, W, H, B): one box {64, 1, 8, 1} = one swizzle atom per halo column
          for (int cx = 0; cx < Cfg::kXT + 2; ++cx)
            tma_load_4d(&tm_a, &full_bar[stage], sa + cx * 1024, kb * kBlockK, x0 - 1 + cx, y0 - 1, b);
        }
        if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1 && lane == 0) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc = umma_idesc_bf16(BLOCK_N);
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    mbar_wait(w_full, 0);
    tc_fence_after();
    const uint32_t w_addr = smem_u32(s_w);
    for (int tile = blockIdx.x; tile < hg.num_tiles; tile += gridDim.x) {
      mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * Cfg::kAccCols;
      for (int kb = 0; kb < CIN_KB; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint32_t a_addr = smem_u32(s_a + stage * Cfg::kABytes);
#pragma unroll 1
        for (int tap = 0; tap < 9; ++tap) {
          const int kh = tap / 3, kw = tap - kh * 3;
          const uint64_t bdesc = umma_desc_sw128(w_addr + (tap * CIN_KB + kb) * Cfg::kWTile);
#pragma unroll
          for (int sub = 0; sub < 2; ++sub) {
            // A row of MMA row m for this tap: m + kw * 8 + (kh - 1); sub-tile s starts 128 rows further
            const uint32_t start = a_addr + static_cast<uint32_t>((sub * 128 + kw * 8 + kh - 1) * 128);
            const uint64_t adesc = umma_desc_sw128_shifted(start, hg.use_base_offset);
#pragma unroll
            for (int k = 0; k < kBlockK / 16; ++k)
              umma_bf16_ss(d_tmem + sub * BLOCK_N, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | tap | k) != 0 ? 1u : 0u);
          }
        }
        umma_commit(&empty_bar[stage]);
        if (kb == CIN_KB - 1) umma_commit(&tmem_full[acc]);
        if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  } else if (warp >= 4) {
    // ===================== epilogue =====================
    const int we = warp - 4;
    const int wq = we & 3;
    const int sub = we >> 2;
    const int r = we * 32 + lane;  // MMA row m
    const int xo = r >> 3, yo = r & 7;
    int acc = 0;
    uint32_t acc_phase = 0;
    constexpr int CH = (BLOCK_N >= 32) ? 32 : 16;
    constexpr int kChunks = BLOCK_N / CH;
    for (int tile = blockIdx.x; tile < hg.num_tiles; tile += gridDim.x) {
      const int b = tile / tiles_per_img;
      const int t = tile - b * tiles_per_img;
      const int y = (t / hg.xtiles) * Cfg::kYT - 1 + yo;
      const int x = (t % hg.xtiles) * Cfg::kXT + xo;
      const bool valid = yo >= 1 && yo <= Cfg::kYT;  // y, x are inside the image by construction (H % 6 == W % 32 == 0)
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(wq * 32) << 16) + acc * Cfg::kAccCols + sub * BLOCK_N;
      float v[kChunks][CH];
#pragma unroll
      for (int c = 0; c < kChunks; ++c) {
        if constexpr (CH == 32) tmem_ld32(t_row + c * CH, v[c]); else tmem_ld16(t_row + c * CH, v[c]);
      }
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(&tmem_empty[acc]);  // accumulators are in registers: the next tile's MMAs may overwrite this stage
#pragma unroll
      for (int c = 0; c < kChunks; ++c) {
        if constexpr (EPI == EPI_ACT) epi_act<CH>(g, e, b, y, x, valid, c * CH, v[c]);
        if constexpr (EPI == EPI_FRAME) epi_frame(g, e, b, y, x, valid, t * (Cfg::kM / 32) + we, v[c]);
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, Cfg::kTmemCols);
}

// ---------------------------------------------------------------------------------------------------------------
template <int CIN_KB, int BLOCK_N, int EPI>
static cudaError_t launch_halo_t(const ConvOp& op, const HaloGeom& hg, const CUtensorMap& tm_a, int num_sms,
                                 cudaStream_t stream) {
  using Cfg = HaloCfg<CIN_KB, BLOCK_N>;
  const int grid = hg.num_tiles < num_sms ? hg.num_tiles : num_sms;
  return launch_pdl(conv_halo_kernel<CIN_KB, BLOCK_N, EPI>, dim3(grid), dim3(Cfg::kThreads), Cfg::kSmemBytes, stream, tm_a, op.tm.w, hg, op.g, op.e);
  return cudaGetLastError();
}

#define RAC_HALO_CASES(X) \
  X(1, 64, EPI_ACT)       \
  X(2, 64, EPI_ACT)       \
  X(1, 16, EPI_FRAME)

bool conv_halo_supported(const ConvOp& op) {
  const ConvGeom& g = op.g;
  if (g.ks != 3 || g.nsrc != 1 || g.W % 32 != 0 || g.H % 6 != 0 || op.e.upsample && false) return false;
#define X(K, N, E) if (g.src_kb[0] == K && op.block_n == N && op.epi == E && g.num_n_tiles == 1) return true;
  RAC_HALO_CASES(X)
#undef X
  return false;
}

cudaError_t launch_conv_halo(const ConvOp& op, const CUtensorMap& tm_a, int column_loads, int use_base_offset,
                             int num_sms, cudaStream_t stream) {
  HaloGeom hg;
  hg.B = op.g.B; hg.H = op.g.H; hg.W = op.g.W;
  hg.xtiles = op.g.W / 32; hg.ytiles = op.g.H / 6;
  hg.num_tiles = op.g.B * hg.xtiles * hg.ytiles;
  hg.column_loads = column_loads;
  hg.use_base_offset = use_base_offset;
#define X(K, N, E) \
  if (op.g.src_kb[0] == K && op.block_n == N && op.epi == E) return launch_halo_t<K, N, E>(op, hg, tm_a, num_sms, stream);
  RAC_HALO_CASES(X)
#undef X
  return cudaErrorInvalidValue;
}

cudaError_t conv_halo_set_attributes() {
  cudaError_t err;
#define X(K, N, E)                                                                                          \
  if ((err = cudaFuncSetAttribute(conv_halo_kernel<K, N, E>, cudaFuncAttributeMaxDynamicSharedMemorySize,   \
                                  HaloCfg<K, N>::kSmemBytes)) != cudaSuccess)                               \
    return err;
  RAC_HALO_CASES(X)
#undef X
  return cudaSuccess;
}

}  // namespace rac
