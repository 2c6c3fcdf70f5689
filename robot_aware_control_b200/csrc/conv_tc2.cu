// CTA-pair (cta_group::2) variant of the implicit-GEMM convolution for the 256 x 256 tiles of the ConvLSTM gate
// convolutions. In conv_tc.cu one CTA owns the whole 256 x 256 fp32 accumulator = all 512 TMEM columns, so TMEM is
// single-buffered and the LSTM epilogue is exposed (tensor pipe active 85-91 % of the kernel). Here the SAME tile is
// computed by two CTAs of a cluster (one TPC): each holds 128 rows of A, 128 of the 256 weight rows and its 128 x 256
// half of the accumulator = 256 TMEM columns, i.e. TWO accumulator stages fit and the epilogue of tile i overlaps the
// main loop of tile i + 1, at the same L2 -> SMEM bytes per FLOP (32 KB per CTA and k-block) as the 256 x 256 tile.
//   warp 0 (both CTAs) : TMA producer -- loads its own halves; the transaction bytes are counted on the LEADER's barrier
//   warp 1 (leader)    : MMA issuer   -- tcgen05.mma.cta_group::2 (M 256 x N 256 x K 16), commit multicast to both CTAs
//   warp 2 (both)      : TMEM allocator (cta_group::2)
//   warps 4..7 (both)  : epilogue of this CTA's 128 rows; accumulator stages are handed back on the leader's barrier
#include "conv.cuh"
#include "epilogue.cuh"
#include "ptx.cuh"

namespace rac {

namespace {

struct Tc2Cfg {
  static constexpr int kBlockN = 256;
  static constexpr int kRowsPerCta = 128;
  static constexpr int kABytes = kRowsPerCta * kBlockK * 2;        // 16 KB
  static constexpr int kBBytes = (kBlockN / 2) * kBlockK * 2;      // 16 KB: this CTA's half of the weight rows
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStages = 6;
  static constexpr int kAccCols = kBlockN;                         // per CTA: 128 lanes x 256 columns
  static constexpr int kTmemCols = 512;                            // two accumulator stages
  static constexpr int kEpiThreads = kRowsPerCta;
  static constexpr int kThreads = 128 + kEpiThreads;
  static constexpr int kBarBytes = 2048;                           // barriers (first 512 B) + bias tile at +1024
  static constexpr int kSmemBytes = kStages * kStageBytes + kBarBytes + 1024;
};

__device__ __forceinline__ bool tap_row_live2(const ConvGeom& g, int y0, int kh) {
  const int ylo = y0 + kh - g.pad;
  return ylo + g.BH > 0 && ylo < g.H;
}

}  // namespace

// g describes the PAIR tile (256 rows: box {64, W, BH, NB}); tm.a[] boxes hold 128 rows ({64, W, BH, NB / 2}),
// tm.w boxes 128 weight rows.
template <int EPI>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(Tc2Cfg::kThreads, 1)
conv_tc2_kernel(const __grid_constant__ ConvTmaps tm, const ConvGeom g, const EpiParams e) {
  using Cfg = Tc2Cfg;
  constexpr int BLOCK_N = Cfg::kBlockN;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* bar_base = smem + Cfg::kStages * Cfg::kStageBytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(bar_base);
  uint64_t* empty_bar = full_bar + Cfg::kStages;
  uint64_t* tmem_full = empty_bar + Cfg::kStages;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int pair = blockIdx.x >> 1;
  const int num_pairs = gridDim.x >> 1;
  const int num_tiles = g.num_m_tiles * g.num_n_tiles;
  int kb_per_tap = 0, live_kb_per_tap = 0;
  for (int s = 0; s < g.nsrc; ++s) {
    kb_per_tap += g.src_kb[s];
    if (!g.src_dead[s]) live_kb_per_tap += g.src_kb[s];
  }

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < g.nsrc; ++s) tma_prefetch_desc(&tm.a[s]);
    tma_prefetch_desc(&tm.w);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < Cfg::kStages; ++i) {
      mbar_init(&full_bar[i], 2);   // one arrive per CTA's producer (+ the transaction bytes of both)
      mbar_init(&empty_bar[i], 1);  // multicast commit of the leader's MMA thread
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], 2 * Cfg::kEpiThreads);  // epilogue threads of both CTAs (used on the leader only)
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc_2sm(tmem_slot, Cfg::kTmemCols);
    tmem_relinquish_2sm();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // the peer's barriers are initialised before anything is signalled on them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  griddep_launch_dependents();  // (PDL, ptx.cuh) the next kernel's prologue may start
  griddep_wait();               // everything below reads / writes global memory of earlier kernels

  if (warp == 0 && lane == 0) {
    // ===================== TMA producer (both CTAs) =====================
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = pair; tile < num_tiles; tile += num_pairs) {
      const int n_tile = tile / g.num_m_tiles;
      const int m_tile = tile - n_tile * g.num_m_tiles;
      const int grp = m_tile / g.tiles_per_img;
      const int b0 = grp * g.NB + static_cast<int>(rank) * (g.NB >> 1);  // this CTA's 128 rows = NB / 2 candidates
      const int y0 = (m_tile - grp * g.tiles_per_img) * g.BH;
      for (int kh = 0; kh < g.ks; ++kh) {
        if (!tap_row_live2(g, y0, kh)) continue;
        for (int kw = 0; kw < g.ks; ++kw) {
          int kidx = (kh * g.ks + kw) * kb_per_tap;
          for (int s = 0; s < g.nsrc; ++s) {
            if (g.src_dead[s]) { kidx += g.src_kb[s]; continue; }
            for (int kb = 0; kb < g.src_kb[s]; ++kb, ++kidx) {
              mbar_wait(&empty_bar[stage], phase ^ 1);
              uint8_t* sa = smem + stage * Cfg::kStageBytes;
              uint8_t* sb = sa + Cfg::kABytes;
              if (leader) mbar_arrive_expect_tx(&full_bar[stage], 2 * Cfg::kStageBytes);
              else mbar_arrive_leader(&full_bar[stage]);
              tma_load_4d_2sm(&tm.a[s], &full_bar[stage], sa, kb * kBlockK, kw - g.pad, y0 + kh - g.pad, b0);
              tma_load_2d_2sm(&tm.w, &full_bar[stage], sb, kidx * kBlockK, n_tile * BLOCK_N + static_cast<int>(rank) * (BLOCK_N / 2));
              if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
            }
          }
        }
      }
    }
  } else if (warp == 1 && lane == 0 && leader) {
    // ===================== MMA issuer (leader CTA) =====================
    constexpr uint32_t idesc = umma_idesc_bf16_m256(BLOCK_N);
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = pair; tile < num_tiles; tile += num_pairs) {
      const int n_tile = tile / g.num_m_tiles;
      const int m_tile = tile - n_tile * g.num_m_tiles;
      const int y0 = (m_tile % g.tiles_per_img) * g.BH;
      int live = 0;
      for (int kh = 0; kh < g.ks; ++kh) live += tap_row_live2(g, y0, kh) ? 1 : 0;
      const int num_kb = live * g.ks * live_kb_per_tap;
      mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * Cfg::kAccCols;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + stage * Cfg::kStageBytes);
        const uint64_t adesc = umma_desc_sw128(sa);
        const uint64_t bdesc = umma_desc_sw128(sa + Cfg::kABytes);
#pragma unroll
        for (int k = 0; k < kBlockK / 16; ++k)
          umma_bf16_ss_2sm(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
        umma_commit_2sm(&empty_bar[stage]);  // frees this stage in BOTH CTAs
        if (kb == num_kb - 1) umma_commit_2sm(&tmem_full[acc]);
        if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  } else if (warp >= 4) {
    // ===================== epilogue (both CTAs: 128 rows each) =====================
    const int we = warp - 4;                                   // TMEM lane quarter
    const int r = static_cast<int>(rank) * 128 + we * 32 + lane;  // row inside the 256-row pair tile
    int acc = 0;
    uint32_t acc_phase = 0;
    constexpr int CH = 32;
    constexpr int kChunks = BLOCK_N / CH;
    constexpr bool kLstm = (EPI == EPI_LSTM);
    float* s_bias = reinterpret_cast<float*>(bar_base + 1024);
    int bias_tile = -1;
    for (int tile = pair; tile < num_tiles; tile += num_pairs) {
      const int n_tile = tile / g.num_m_tiles;
      const int m_tile = tile - n_tile * g.num_m_tiles;
      const int grp = m_tile / g.tiles_per_img;
      const int yb = m_tile - grp * g.tiles_per_img;
      const int b = grp * g.NB + (r >> g.bhw_shift);
      const int y = yb * g.BH + ((r >> g.w_shift) & (g.BH - 1));
      const int x = r & (g.W - 1);
      const bool valid = b < g.B;
      const size_t ctile = (static_cast<size_t>(m_tile) * (e.hid >> 3) * 2 * 256 + r) * 4;
      if constexpr (kLstm) {
        if (n_tile != bias_tile) {
          epi_bar_sync(Cfg::kEpiThreads);
          for (int i = we * 32 + lane; i < BLOCK_N; i += Cfg::kEpiThreads) s_bias[i] = __ldg(e.bias + n_tile * BLOCK_N + i);
          epi_bar_sync(Cfg::kEpiThreads);
          bias_tile = n_tile;
        }
      }
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(we * 32) << 16) + acc * Cfg::kAccCols;
      float v[2][CH];
      float cprev[2][8];
      auto issue = [&](int c, float* dst) { tmem_ld32(t_row + c * CH, dst); };
      auto load_c = [&](int c, float* dst) {
        if constexpr (kLstm) lstm_load_c<false>(g, e, b, y, x, valid, n_tile * BLOCK_N + c * CH, dst, ctile, 256);
      };
      auto process = [&](int c, const float* acc_v, const float* cp) {
        const int n0 = n_tile * BLOCK_N + c * CH;
        if constexpr (kLstm) epi_lstm<false>(g, e, b, y, x, valid, n0, acc_v, cp, ctile, 256, s_bias + c * CH);
        if constexpr (EPI == EPI_ACT) epi_act<CH>(g, e, b, y, x, valid, n0, acc_v);
      };
      issue(0, v[0]);
      load_c(0, cprev[0]);
#pragma unroll 1
      for (int c = 0; c < kChunks; c += 2) {
        tmem_ld_wait();
        issue(c + 1, v[1]);
        load_c(c + 1, cprev[1]);
        process(c, v[0], cprev[0]);
        tmem_ld_wait();
        if (c + 2 < kChunks) {
          issue(c + 2, v[0]);
          load_c(c + 2, cprev[0]);
        }
        process(c + 1, v[1], cprev[1]);
      }
      tc_fence_before();
      if (leader) mbar_arrive(&tmem_empty[acc]); else mbar_arrive_leader(&tmem_empty[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // both CTAs are done with the shared accumulator / barriers before TMEM is released
  if (warp == 2) tmem_dealloc_2sm(tmem_base, Cfg::kTmemCols);
}

// ---------------------------------------------------------------------------------------------------------------
bool conv_tc2_supported(const ConvOp& op) {
  return op.block_m == 256 && op.block_n == 256 && (op.epi == EPI_LSTM || op.epi == EPI_ACT) && (op.g.NB % 2) == 0 &&
         op.g.NB >= 2;
}

cudaError_t launch_conv_tc2(const ConvOp& op, const ConvTmaps& tm2, int num_sms, cudaStream_t stream) {
  const int num_tiles = op.g.num_m_tiles * op.g.num_n_tiles;
  int pairs = num_sms / 2;
  if (pairs > num_tiles) pairs = num_tiles;
  if (pairs < 1) return cudaErrorInvalidValue;
  if (op.epi == EPI_LSTM)
    return launch_pdl(conv_tc2_kernel<EPI_LSTM>, dim3(2 * pairs), dim3(Tc2Cfg::kThreads), Tc2Cfg::kSmemBytes, stream, tm2, op.g, op.e);
  else if (op.epi == EPI_ACT)
    return launch_pdl(conv_tc2_kernel<EPI_ACT>, dim3(2 * pairs), dim3(Tc2Cfg::kThreads), Tc2Cfg::kSmemBytes, stream, tm2, op.g, op.e);
  else
    return cudaErrorInvalidValue;
  return cudaGetLastError();
}

cudaError_t conv_tc2_set_attributes() {
  cudaError_t err = cudaFuncSetAttribute(conv_tc2_kernel<EPI_LSTM>, cudaFuncAttributeMaxDynamicSharedMemorySize, Tc2Cfg::kSmemBytes);
  if (err != cudaSuccess) return err;
  return cudaFuncSetAttribute(conv_tc2_kernel<EPI_ACT>, cudaFuncAttributeMaxDynamicSharedMemorySize, Tc2Cfg::kSmemBytes);
}

}  // namespace rac
