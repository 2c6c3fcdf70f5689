// ConvLSTM gate convolution with the activation tile shared by a 2-CTA cluster through TMA multicast.
//
// conv_tc_kernel<256, 256, EPI_LSTM> pulls 64 KB (32 KB activations + 32 KB weights) per 64-channel k-block into every
// SM; at the MMA rate of the 256 x 256 tile (1024 cycles per k-block) that is 64 B/clk/SM x 148 SMs = 9.5 KB/clk of L2
// output, above what the L2 sustains (~6.3 KB/clk full chip, B300_MICROARCH.md): the main loop measured 1230 cycles
// per k-block -- the kernel is L2-feed bound, which is also why hiding its epilogue (CTA pairs) did not help.
// Here the two CTAs of a cluster work on the SAME 256 rows and two neighbouring 256-column tiles: each loads one
// 128-row half of the activation tile (with y-major rows: one row of the 6x8 map) and multicasts it into both CTAs'
// shared memory, plus its own weight tile: 48 KB per CTA and k-block (-25 %).
//   * full barrier (per CTA): own arrive.expect_tx(64 KB); bytes come from both CTAs' activation halves + own weights
//   * empty barrier (per CTA): count 2 -- the stage is rewritten by BOTH producers, so both consumers release it
//     (tcgen05.commit ... multicast::cluster to both CTAs)
// Everything else (single 512-column TMEM stage, rolled LSTM epilogue, tiled cell-state layout, zero-h skip, skipping of
// padding-only filter rows per sub-tile) is as in conv_tc.cu.
#include "conv.cuh"
#include "epilogue.cuh"
#include "ptx.cuh"

namespace rac {

namespace {

struct McCfg {
  static constexpr int kBlockM = 256, kBlockN = 256;
  static constexpr int kABytes = kBlockM * kBlockK * 2;   // 32 KB (two 16 KB halves, one per CTA of the cluster)
  static constexpr int kBBytes = kBlockN * kBlockK * 2;   // 32 KB
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStages = 3;
  static constexpr int kEpiThreads = kBlockM;
  static constexpr int kThreads = 128 + kEpiThreads;
  static constexpr int kBarBytes = 2048;
  static constexpr int kSmemBytes = kStages * kStageBytes + kBarBytes + 1024;
  static constexpr int kTmemCols = 512;
};

}  // namespace

// g: pair-invariant geometry of the 256-row tile (y_major, BH == 2, NB * W == 128); tm.a[]: 128-row boxes {64, W, NB, 1}
// on the (C, W, B, H) view; tm.w: box {64, 256}.
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(McCfg::kThreads, 1)
conv_tc_mc_kernel(const __grid_constant__ ConvTmaps tm, const ConvGeom g, const EpiParams e) {
  using Cfg = McCfg;
  constexpr int BLOCK_N = Cfg::kBlockN;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* bar_base = smem + Cfg::kStages * Cfg::kStageBytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(bar_base);
  uint64_t* empty_bar = full_bar + Cfg::kStages;
  uint64_t* tmem_full = empty_bar + Cfg::kStages;
  uint64_t* tmem_empty = tmem_full + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int rank = static_cast<int>(cluster_ctarank());
  const int pair = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;
  const int n_pairs = g.num_n_tiles >> 1;
  const int num_tiles = g.num_m_tiles * n_pairs;  // pair tiles: (m_tile, two neighbouring column tiles)
  int kb_per_tap = 0, live_kb_per_tap = 0;
  for (int s = 0; s < g.nsrc; ++s) {
    kb_per_tap += g.src_kb[s];
    if (!g.src_dead[s]) live_kb_per_tap += g.src_kb[s];
  }
  // Tail splitting (as in conv_tc.cu): the pair tiles left over after the full rounds become two work items each, one
  // per map row (= MMA sub-tile = activation half), when they then still fit into one round.
  const int full_items = (num_tiles / num_pairs) * num_pairs;
  const int rem = num_tiles - full_items;
  const bool split_tail = !g.no_split_tail && rem > 0 && 2 * rem <= num_pairs;
  const int num_items = split_tail ? full_items + 2 * rem : num_tiles;
  auto item_tile = [&](int item, int& half) {  // half: -1 = both map rows
    if (item < full_items || !split_tail) { half = -1; return item; }
    half = (item - full_items) & 1;
    return full_items + ((item - full_items) >> 1);
  };
  // filter row kh contributes to this work item: some (half < 0) / the (half >= 0) map row reads a real input row
  auto row_live = [&](int y0, int kh, int half) {
    if (half >= 0) {
      const int yy = y0 + half + kh - g.pad;
      return yy >= 0 && yy < g.H;
    }
    const int ylo = y0 + kh - g.pad;
    return ylo + 2 > 0 && ylo < g.H;
  };

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < g.nsrc; ++s) tma_prefetch_desc(&tm.a[s]);
    tma_prefetch_desc(&tm.w);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < Cfg::kStages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 2);  // released by the MMA threads of both CTAs
    }
    mbar_init(tmem_full, 1);
    mbar_init(tmem_empty, Cfg::kEpiThreads);
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, Cfg::kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  griddep_launch_dependents();  // (PDL, ptx.cuh) the next kernel's prologue may start
  griddep_wait();               // everything below reads / writes global memory of earlier kernels

  if (warp == 0 && lane == 0) {
    // ===================== TMA producer =====================
    int stage = 0;
    uint32_t phase = 0;
    for (int item = pair; item < num_items; item += num_pairs) {
      int half;
      const int tile = item_tile(item, half);
      const int np = tile / g.num_m_tiles;
      const int m_tile = tile - np * g.num_m_tiles;
      const int n_tile = 2 * np + rank;
      const int grp = m_tile / g.tiles_per_img;
      const int b0 = grp * g.NB;
      const int y0 = (m_tile - grp * g.tiles_per_img) * 2;
      // a half item needs only activation half `half`: the CTA of that rank loads (and multicasts) it
      const bool load_a = half < 0 || half == rank;
      const uint32_t tx = (half < 0 ? Cfg::kABytes : Cfg::kABytes / 2) + Cfg::kBBytes;
      for (int kh = 0; kh < g.ks; ++kh) {
        if (!row_live(y0, kh, half)) continue;
        for (int kw = 0; kw < g.ks; ++kw) {
          int kidx = (kh * g.ks + kw) * kb_per_tap;
          for (int s = 0; s < g.nsrc; ++s) {
            if (g.src_dead[s]) { kidx += g.src_kb[s]; continue; }
            for (int kb = 0; kb < g.src_kb[s]; ++kb, ++kidx) {
              mbar_wait(&empty_bar[stage], phase ^ 1);  // both CTAs have consumed the previous contents
              uint8_t* sa = smem + stage * Cfg::kStageBytes;
              uint8_t* sb = sa + Cfg::kABytes;
              mbar_arrive_expect_tx(&full_bar[stage], tx);
              // my half of the activation tile (map row y0 + rank) -> both CTAs
              if (load_a)
                tma_load_4d_mc(&tm.a[s], &full_bar[stage], sa + rank * (Cfg::kABytes / 2), kb * kBlockK, kw - g.pad, b0,
                               y0 + rank + kh - g.pad, 3);
              tma_load_2d(&tm.w, &full_bar[stage], sb, kidx * kBlockK, n_tile * BLOCK_N);
              if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
            }
          }
        }
      }
    }
  } else if (warp == 1 && lane == 0) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc = umma_idesc_bf16(BLOCK_N);
    int stage = 0;
    uint32_t phase = 0;
    uint32_t acc_phase = 0;
    for (int item = pair; item < num_items; item += num_pairs) {
      int half;
      const int tile = item_tile(item, half);
      const int np = tile / g.num_m_tiles;
      const int m_tile = tile - np * g.num_m_tiles;
      const int y0 = (m_tile % g.tiles_per_img) * 2;
      int live = 0;
      for (int kh = 0; kh < g.ks; ++kh) live += row_live(y0, kh, half) ? 1 : 0;
      const int num_kb = live * g.ks * live_kb_per_tap;
      mbar_wait(tmem_empty, acc_phase ^ 1);
      tc_fence_after();
      uint32_t started = 0;
      int kb = 0;
      for (int kh = 0; kh < g.ks; ++kh) {
        if (!row_live(y0, kh, half)) continue;
        uint32_t sub_live = 0;
        for (int sb = 0; sb < 2; ++sb) {
          const int yy = y0 + sb + kh - g.pad;
          if (yy >= 0 && yy < g.H && (half < 0 || half == sb)) sub_live |= 1u << sb;
        }
        for (int rest = g.ks * live_kb_per_tap; rest > 0; --rest, ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * Cfg::kStageBytes);
          const uint64_t adesc = umma_desc_sw128(sa);
          const uint64_t bdesc = umma_desc_sw128(sa + Cfg::kABytes);
#pragma unroll
          for (int k = 0; k < kBlockK / 16; ++k) {
#pragma unroll
            for (int sub = 0; sub < 2; ++sub) {
              if (!((sub_live >> sub) & 1u)) continue;
              umma_bf16_ss(tmem_base + sub * BLOCK_N, adesc + 2 * k + sub * (kTileM * 128 / 16), bdesc + 2 * k, idesc,
                           (started >> sub) & 1u);
              started |= 1u << sub;
            }
          }
          umma_commit_mc(&empty_bar[stage], 3);  // this stage may be rewritten once BOTH CTAs have read it
          if (kb == num_kb - 1) umma_commit(tmem_full);
          if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
        }
      }
      acc_phase ^= 1;
    }
  } else if (warp >= 4) {
    // ===================== epilogue =====================
    const int we = warp - 4;
    const int wq = we & 3;
    const int sub = we >> 2;
    const int r = we * 32 + lane;
    uint32_t acc_phase = 0;
    constexpr int CH = 32;
    constexpr int kChunks = BLOCK_N / CH;
    float* s_bias = reinterpret_cast<float*>(bar_base + 1024);
    int bias_tile = -1;
    for (int item = pair; item < num_items; item += num_pairs) {
      int half;
      const int tile = item_tile(item, half);
      const bool mine = half < 0 || half == sub;  // half item: the other map row belongs to another cluster
      const int np = tile / g.num_m_tiles;
      const int m_tile = tile - np * g.num_m_tiles;
      const int n_tile = 2 * np + rank;
      const int grp = m_tile / g.tiles_per_img;
      const int yb = m_tile - grp * g.tiles_per_img;
      int b, y, x;
      tile_row_to_pixel(g, grp, yb, r, b, y, x);
      const bool valid = b < g.B;
      const size_t ctile = (static_cast<size_t>(m_tile) * (e.hid >> 3) * 2 * Cfg::kBlockM + r) * 4;
      if (n_tile != bias_tile) {
        epi_bar_sync(Cfg::kEpiThreads);
        for (int i = r; i < BLOCK_N; i += Cfg::kEpiThreads) s_bias[i] = __ldg(e.bias + n_tile * BLOCK_N + i);
        epi_bar_sync(Cfg::kEpiThreads);
        bias_tile = n_tile;
      }
      mbar_wait(tmem_full, acc_phase);
      tc_fence_after();
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(wq * 32) << 16) + sub * BLOCK_N;
      float v[2][CH];
      float cprev[2][8];
      auto issue = [&](int c, float* dst) { tmem_ld32(t_row + c * CH, dst); };
      auto load_c = [&](int c, float* dst) {
        lstm_load_c<false>(g, e, b, y, x, valid, n_tile * BLOCK_N + c * CH, dst, ctile, Cfg::kBlockM);
      };
      auto process = [&](int c, const float* acc_v, const float* cp) {
        epi_lstm<false>(g, e, b, y, x, valid, n_tile * BLOCK_N + c * CH, acc_v, cp, ctile, Cfg::kBlockM, s_bias + c * CH);
      };
      if (mine) {
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
      }
      tc_fence_before();
      mbar_arrive(tmem_empty);
      acc_phase ^= 1;
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // the peer may still multicast into this CTA's shared memory / signal its barriers
  if (warp == 2) tmem_dealloc(tmem_base, Cfg::kTmemCols);
}

bool conv_tc_mc_supported(const ConvOp& op) {
  return op.epi == EPI_LSTM && op.block_m == 256 && op.block_n == 256 && op.g.y_major && op.g.BH == 2 &&
         op.g.NB * op.g.W == 128 && (op.g.num_n_tiles % 2) == 0;
}

cudaError_t launch_conv_tc_mc(const ConvOp& op, const ConvTmaps& tm_mc, int num_sms, cudaStream_t stream) {
  const int num_tiles = op.g.num_m_tiles * (op.g.num_n_tiles / 2);
  int pairs = num_sms / 2;
  if (pairs > num_tiles) pairs = num_tiles;
  if (pairs < 1) return cudaErrorInvalidValue;
  return launch_pdl(conv_tc_mc_kernel, dim3(2 * pairs), dim3(McCfg::kThreads), McCfg::kSmemBytes, stream, tm_mc, op.g, op.e);
  return cudaGetLastError();
}

cudaError_t conv_tc_mc_set_attributes() {
  return cudaFuncSetAttribute(conv_tc_mc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, McCfg::kSmemBytes);
}

}  // namespace rac
