// Implicit-GEMM convolution on the 5th-generation tensor cores (tcgen05.mma, accumulators in TMEM), operands staged
// by TMA into 128B-swizzled shared memory. Persistent, warp-specialised:
//   warp 0   : TMA producer  (one thread)   -- A: 4-D box {64 ch, W, BH, NB} per filter tap, B: 2-D weight box
//   warp 1   : MMA issuer    (one thread)   -- 4 x (BLOCK_M/128) tcgen05.mma (M128 x BLOCK_N x K16) per 64-channel k-block
//   warp 2   : TMEM allocator
//   warps 4.. : epilogue     (one thread per output row: 128 or 256) -- tcgen05.ld -> fused epilogue -> global
// A CTA tile is BLOCK_M (128 or 256 rows = 1 or 2 MMA sub-tiles sharing the weight tile) x BLOCK_N. Two TMEM accumulator
// stages let the epilogue of tile i overlap the main loop of tile i+1 whenever 2 x (BLOCK_M/128) x BLOCK_N <= 512 columns;
// the 256x256 tile of the LSTM gate convolutions uses all 512 columns single-buffered (its main loop is 150-400
// k-blocks long, the exposed epilogue is a few per cent) to halve the L2->SMEM bytes per FLOP.
//
// Replaces the cuDNN/ATen calls behind nn.Conv2d / BatchNorm2d / LeakyReLU / ConvTranspose2d / Sigmoid and the
// LSTM / reparameterisation / compositing / cost pointwise ops of the reference:
//   src/prediction/models/vgg_64.py:8-18,122-129,223-241; lstm.py:129-149,276-286; dynamics.py:594-643;
//   src/cem/trajectory_sampler.py:148-168.
#include "conv.cuh"
#include "epilogue.cuh"
#include "ptx.cuh"

namespace rac {

template <int BLOCK_M, int BLOCK_N>
struct TcCfg {
  static constexpr int kSub = BLOCK_M / kTileM;              // 128-row MMA sub-tiles per CTA tile
  static constexpr int kABytes = BLOCK_M * kBlockK * 2;      // BLOCK_M rows of 128 B
  static constexpr int kBBytes = BLOCK_N * kBlockK * 2;      // BLOCK_N rows of 128 B
  static constexpr int kBBytesPad = (kBBytes + 1023) / 1024 * 1024;
  static constexpr int kStageBytes = kABytes + kBBytesPad;
  static constexpr int kStagesFit = (200 * 1024) / kStageBytes;
  static constexpr int kStages = kStagesFit > 8 ? 8 : kStagesFit;
  static constexpr int kAccCols = kSub * BLOCK_N;            // TMEM columns of one accumulator stage
  static constexpr int kNumAcc = (2 * kAccCols <= 512) ? 2 : 1;
  static constexpr int kTmemNeed = kNumAcc * kAccCols;
  static constexpr int kTmemCols = kTmemNeed <= 32 ? 32 : (kTmemNeed <= 64 ? 64 : (kTmemNeed <= 128 ? 128 : (kTmemNeed <= 256 ? 256 : 512)));
  static constexpr int kEpiThreads = BLOCK_M;                // one thread per output row
  static constexpr int kThreads = 128 + kEpiThreads;
  static constexpr int kBarBytes = 2048;  // mbarriers + TMEM slot (first 512 B) + LSTM bias tile (BLOCK_N floats at +1024)
  static constexpr int kEpiStageBytes = (kEpiThreads / 32) * kStageWarpBytes;  // fp32 epilogues: per-warp transpose tiles
  static constexpr int kSmemBytes = kStages * kStageBytes + kBarBytes + kEpiStageBytes + 1024;  // +1024 alignment slack
  static_assert(kSmemBytes <= 227 * 1024, "shared memory budget");
  static_assert(kTmemNeed <= 512, "accumulators do not fit TMEM");
  static_assert(kStages >= 2, "pipeline needs at least two stages");
};

// A filter row kh contributes nothing to a tile whose input rows y0+kh-pad .. +BH-1 all fall into the zero padding
// (top / bottom tiles of the 6x8 latent maps with the 5x5 filter): producer and MMA issuer both skip it.
__device__ __forceinline__ bool tap_row_live(const ConvGeom& g, int y0, int kh) {
  const int ylo = y0 + kh - g.pad;
  return ylo + g.BH > 0 && ylo < g.H;
}

template <int BLOCK_M, int BLOCK_N, int EPI>
__global__ void __launch_bounds__(TcCfg<BLOCK_M, BLOCK_N>::kThreads, 1)
conv_tc_kernel(const __grid_constant__ ConvTmaps tm, const ConvGeom g, const EpiParams e) {
  using Cfg = TcCfg<BLOCK_M, BLOCK_N>;
  // dgrad with the weights in their forward packing Wp[n][tap][c]: tm.w is a 3-D map (c, tap, n), a 64-channel box
  // {64 c, 1 tap, 64 n} is one MN-major swizzle group of the B operand (K = n is the slow dimension), taps are flipped
  constexpr bool kBmn = (EPI == EPI_F32_BT);
  constexpr bool kSplitK = (EPI == EPI_F32 || EPI == EPI_F32_BT);
  static_assert(!kBmn || BLOCK_N % 64 == 0, "MN-major weight boxes are 64 channels wide");
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
  const int num_tiles = g.num_m_tiles * g.num_n_tiles;
  int kb_per_tap = 0, live_kb_per_tap = 0;
  for (int s = 0; s < g.nsrc; ++s) {
    kb_per_tap += g.src_kb[s];
    if (!g.src_dead[s]) live_kb_per_tap += g.src_kb[s];
  }
  // Tail splitting against wave quantisation: the persistent grid processes floor(tiles / grid) full rounds; if the
  // remaining `rem` tiles would occupy at most half of the CTAs for one more full tile time (e.g. 3000 LSTM tiles =
  // 20 x 148 + 40), each of them becomes two work items, one per 128-row MMA sub-tile (same operand loads, half the
  // MMAs, half the epilogue rows), so the last round takes about half a tile time on twice as many SMs.
  const int full_items = (num_tiles / static_cast<int>(gridDim.x)) * static_cast<int>(gridDim.x);
  const int rem = num_tiles - full_items;
  const bool split_tail = Cfg::kSub == 2 && !g.no_split_tail && rem > 0 && 2 * rem <= static_cast<int>(gridDim.x);
  // Split-K (g.ksplit > 1, fp32 epilogues of the training step only): the recurrent GEMMs of a batch-16 training step have
  // 768 rows = 48 tiles of 128 x 128 for 148 SMs but 288-800 k-blocks; every tile becomes ksplit work items over
  // contiguous k-block ranges, each writes its raw accumulator to its own slice of e.split_part, and
  // splitk_reduce_kernel sums the slices in a fixed order (deterministic) and applies the real epilogue.
  const int ksplit = (kSplitK && g.ksplit > 1) ? g.ksplit : 1;
  const bool to_slices = kSplitK && e.split_part != nullptr;  // (one slice when ksplit == 1)
  const int num_items = to_slices ? num_tiles * ksplit : (split_tail ? full_items + 2 * rem : num_tiles);
  auto item_tile = [&](int item, int& half, int& split) {  // half: -1 = both sub-tiles
    split = 0;
    if (to_slices) { half = -1; split = item / num_tiles; return item - split * num_tiles; }
    if (item < full_items || !split_tail) { half = -1; return item; }
    half = (item - full_items) & 1;
    return full_items + ((item - full_items) >> 1);
  };
  // live k-blocks of a tile whose first map row is y0, and the range [lo, hi) of them that split `split` handles
  auto kb_range = [&](int y0, int split, int& lo, int& hi) {
    int live = 0;
    for (int kh = 0; kh < g.ks; ++kh) live += tap_row_live(g, y0, kh) ? 1 : 0;
    const int num_kb = live * g.ks * live_kb_per_tap;
    lo = static_cast<int>(static_cast<long long>(split) * num_kb / ksplit);
    hi = static_cast<int>(static_cast<long long>(split + 1) * num_kb / ksplit);
  };

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < g.nsrc; ++s) tma_prefetch_desc(&tm.a[s]);
    tma_prefetch_desc(&tm.w);
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
    tmem_slot[1] = 0;  // (kb_issued)
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
  auto stamp = [&](int i) {
    if (g.timeline) {
      unsigned long long t;
      asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
      g.timeline[blockIdx.x * 8 + i] = t;
    }
  };
  if (threadIdx.x == 0) stamp(0);

  // the live k-blocks of one work item, in issue order: fn(n_tile, b0, y0, kh, kw, s, kb, kidx). A split-K item starts
  // at its first k-block directly (walking there one block at a time cost the single thread ~30 ns per skipped block:
  // up to 25 us before the first load of the last slice of an 800-block tile)
  auto for_each_kb = [&](int item, auto&& fn) {
    int half, split;
    const int tile = item_tile(item, half, split);
    const int n_tile = tile / g.num_m_tiles;
    const int m_tile = tile - n_tile * g.num_m_tiles;
    const int grp = m_tile / g.tiles_per_img;
    const int b0 = grp * g.NB;
    const int y0 = (m_tile - grp * g.tiles_per_img) * g.BH;
    int lo, hi;
    kb_range(y0, split, lo, hi);
    if (lo >= hi) return;
    const int per_row = g.ks * live_kb_per_tap;
    int khl = lo / per_row;                 // index among the LIVE filter rows
    int rem = lo - khl * per_row;
    int kw = rem / live_kb_per_tap;
    rem -= kw * live_kb_per_tap;            // live k-block inside the tap
    int kh = -1;
    for (int c = -1; c < khl;) { ++kh; if (tap_row_live(g, y0, kh)) ++c; }
    int s = 0, soff = 0;                    // source and its first k-block inside a tap (dead sources included)
    while (g.src_dead[s] || rem >= g.src_kb[s]) {
      if (!g.src_dead[s]) rem -= g.src_kb[s];
      soff += g.src_kb[s];
      ++s;
    }
    int kb = rem;
    for (int kbi = lo; kbi < hi; ++kbi) {
      fn(n_tile, b0, y0, kh, kw, s, kb, (kh * g.ks + kw) * kb_per_tap + soff + kb);
      if (++kb == g.src_kb[s]) {
        kb = 0;
        do { soff += g.src_kb[s]; ++s; } while (s < g.nsrc && g.src_dead[s]);
        if (s == g.nsrc) {
          s = 0; soff = 0;
          while (g.src_dead[s]) { soff += g.src_kb[s]; ++s; }
          if (++kw == g.ks) {
            kw = 0;
            do { ++kh; } while (kh < g.ks && !tap_row_live(g, y0, kh));
          }
        }
      }
    }
  };
  // weight-streaming GEMMs (g.w_prefetch > 0: few M-tiles, every weight box comes from DRAM and is used by 1-3 CTAs):
  // the producer's count of issued k-blocks, read by the L2 prefetch thread (warp 3) to stay w_prefetch blocks ahead
  volatile int* kb_issued = reinterpret_cast<volatile int*>(tmem_slot + 1);

  if (warp == 0 && lane == 0) {
    // ===================== TMA producer =====================
    int stage = 0;
    uint32_t phase = 0;
    int issued = 0;
    for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
      for_each_kb(item, [&](int n_tile, int b0, int y0, int kh, int kw, int s, int kb, int kidx) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        uint8_t* sa = smem + stage * Cfg::kStageBytes;
        uint8_t* sb = sa + Cfg::kABytes;
        mbar_arrive_expect_tx(&full_bar[stage], Cfg::kABytes + Cfg::kBBytes);
        if (g.y_major) tma_load_4d(&tm.a[s], &full_bar[stage], sa, kb * kBlockK, kw - g.pad, b0, y0 + kh - g.pad);
        else tma_load_4d(&tm.a[s], &full_bar[stage], sa, kb * kBlockK, kw - g.pad, y0 + kh - g.pad, b0);
        if constexpr (kBmn) {
          const int taps = g.ks * g.ks;
          const int tapf = taps - 1 - (kh * g.ks + kw);
#pragma unroll
          for (int j = 0; j < BLOCK_N / 64; ++j) {
            if (g.w_tiled)  // panel (flipped tap, 64-channel block of N), rows kb * 64 .. + 63 of it
              tma_load_3d(&tm.w, &full_bar[stage], sb + j * 8192, 0, kb * kBlockK,
                          tapf * (g.num_n_tiles * (BLOCK_N / 64)) + n_tile * (BLOCK_N / 64) + j);
            else
              tma_load_3d(&tm.w, &full_bar[stage], sb + j * 8192, n_tile * BLOCK_N + j * 64, tapf, kb * kBlockK);
          }
        } else if (g.w_tiled) {
          tma_load_3d(&tm.w, &full_bar[stage], sb, 0, n_tile * BLOCK_N, kidx);
        } else {
          tma_load_2d(&tm.w, &full_bar[stage], sb, kidx * kBlockK, n_tile * BLOCK_N);
        }
        if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
        if (g.w_prefetch > 0) *kb_issued = ++issued;
      });
    }
  } else if (warp == 3 && lane == 0 && g.w_prefetch > 0) {
    // ===================== L2 prefetch of the weight boxes =====================
    // The DRAM latency of a weight box (~2 us under load) is longer than the 2-5 stages of shared memory can cover
    // (profiles/r02_dgrad5x5_ncu_s7.txt: 0.8 TB/s of DRAM reads, tensor pipe 37 % busy, nothing saturated): this thread
    // walks the same k-blocks w_prefetch ahead of the producer and pulls their weight boxes into L2, so that the
    // producer's loads are L2 hits.
    int ahead = 0;
    for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
      for_each_kb(item, [&](int n_tile, int, int, int kh, int kw, int, int kb, int kidx) {
        while (ahead - *kb_issued >= g.w_prefetch) __nanosleep(64);
        if constexpr (kBmn) {
          const int taps = g.ks * g.ks;
          const int tapf = taps - 1 - (kh * g.ks + kw);
#pragma unroll
          for (int j = 0; j < BLOCK_N / 64; ++j) {
            if (g.w_tiled)
              tma_prefetch_l2_3d(&tm.w, 0, kb * kBlockK, tapf * (g.num_n_tiles * (BLOCK_N / 64)) + n_tile * (BLOCK_N / 64) + j);
            else
              tma_prefetch_l2_3d(&tm.w, n_tile * BLOCK_N + j * 64, tapf, kb * kBlockK);
          }
        } else if (g.w_tiled) {
          tma_prefetch_l2_3d(&tm.w, 0, n_tile * BLOCK_N, kidx);
        } else {
          tma_prefetch_l2_2d(&tm.w, kidx * kBlockK, n_tile * BLOCK_N);
        }
        ++ahead;
      });
    }
  } else if (warp == 1 && lane == 0) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc = umma_idesc_bf16(BLOCK_N) | (kBmn ? kIdescBMn : 0u);
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
      int half, split;
      const int tile = item_tile(item, half, split);
      const int n_tile = tile / g.num_m_tiles;
      const int m_tile = tile - n_tile * g.num_m_tiles;
      const int y0 = (m_tile % g.tiles_per_img) * g.BH;
      int lo, hi;
      kb_range(y0, split, lo, hi);  // (ksplit == 1: the whole range)
      mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * Cfg::kAccCols;
      // y_major tiles: sub-tile s is the single map row y0 + s, so for filter row kh its MMAs are dead when that
      // input row is zero padding (the operand rows TMA wrote are all zero)
      const bool per_sub = g.y_major != 0 && Cfg::kSub == 2;
      uint32_t started = 0;  // bit s: sub-tile s has received its first (non-accumulating) MMA
      const int per_row = g.ks * live_kb_per_tap;
      const int khl = lo / per_row;            // first LIVE filter row of this item and the k-blocks already behind it
      int in_row = lo - khl * per_row;
      int kh = -1;
      for (int c = -1; c < khl;) { ++kh; if (tap_row_live(g, y0, kh)) ++c; }
      int kb = lo;
      while (kb < hi) {
        uint32_t sub_live = 3;
        if (per_sub) {
          sub_live = 0;
          for (int sb = 0; sb < 2; ++sb) {
            const int yy = y0 + sb + kh - g.pad;
            if (yy >= 0 && yy < g.H) sub_live |= 1u << sb;
          }
        }
        const int row_end = min(hi, kb + per_row - in_row);
        for (; kb < row_end; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          if (kb == lo && item == static_cast<int>(blockIdx.x)) stamp(1);
          const uint32_t sa = smem_u32(smem + stage * Cfg::kStageBytes);
          const uint64_t adesc = umma_desc_sw128(sa);
          const uint64_t bdesc = kBmn ? umma_desc_sw128_mn(sa + Cfg::kABytes, 8192u) : umma_desc_sw128(sa + Cfg::kABytes);
          constexpr uint32_t kBStep = kBmn ? 128u : 2u;  // 16 K-rows of an MN-major operand = 2048 B; 16 K-elements of a K-major row = 32 B
#pragma unroll
          for (int k = 0; k < kBlockK / 16; ++k) {
#pragma unroll
            for (int sub = 0; sub < Cfg::kSub; ++sub) {
              if (half >= 0 && sub != half) continue;  // split tail item: only this sub-tile
              if (!((sub_live >> sub) & 1u)) continue;
              // +2 in the (addr >> 4) field = 32 B = 16 bf16 along K inside the 128B swizzle row;
              // sub-tile s starts 128 rows * 128 B = 16 KB further (1024-B aligned, swizzle phase preserved)
              umma_bf16_ss(d_tmem + sub * BLOCK_N, adesc + 2 * k + sub * (kTileM * 128 / 16), bdesc + kBStep * k, idesc,
                           (started >> sub) & 1u);
              started |= 1u << sub;
            }
          }
          umma_commit(&empty_bar[stage]);  // frees the smem stage once these MMAs have read it
          if (kb == hi - 1) { umma_commit(&tmem_full[acc]); if (item == static_cast<int>(blockIdx.x)) stamp(2); }
          if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
        }
        in_row = 0;
        do { ++kh; } while (kh < g.ks && !tap_row_live(g, y0, kh));
      }
      if (++acc == Cfg::kNumAcc) { acc = 0; acc_phase ^= 1; }
    }
  } else if (warp >= 4) {
    // ===================== epilogue =====================
    const int we = warp - 4;       // epilogue warp index
    const int wq = we & 3;         // TMEM lane quarter == warp_id % 4
    const int sub = we >> 2;       // 128-row sub-tile
    const int r = we * 32 + lane;  // row inside the BLOCK_M tile
    int acc = 0;
    uint32_t acc_phase = 0;
    constexpr int CH = (BLOCK_N >= 32) ? 32 : 16;
    float* s_bias = reinterpret_cast<float*>(bar_base + 1024);
    float* s_stage = reinterpret_cast<float*>(bar_base + Cfg::kBarBytes + we * kStageWarpBytes);  // this warp's transpose tile
    int bias_tile = -1;
    for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
      int half, split;
      const int tile = item_tile(item, half, split);
      const bool mine = half < 0 || sub == half;  // split tail item: the other sub-tile's rows belong to another CTA
      const int n_tile = tile / g.num_m_tiles;
      const int m_tile = tile - n_tile * g.num_m_tiles;
      const int grp = m_tile / g.tiles_per_img;
      const int yb = m_tile - grp * g.tiles_per_img;
      int b, y, x;
      tile_row_to_pixel(g, grp, yb, r, b, y, x);
      const bool valid = b < g.B;
      constexpr bool kLstm = (EPI == EPI_LSTM || EPI == EPI_LSTM_TRAIN);
      constexpr bool kTrain = (EPI == EPI_LSTM_TRAIN);
      const size_t ctile = (static_cast<size_t>(m_tile) * (e.hid >> 3) * 2 * BLOCK_M + r) * 4;
      if constexpr (kLstm) {
        // the BLOCK_N gate biases of this column tile go to shared memory while the main loop is still running
        // (every epilogue thread needs all of them: 8 broadcast LDS.128 per chunk instead of 8 global loads)
        if (n_tile != bias_tile) {
          epi_bar_sync(Cfg::kEpiThreads);  // nobody still reads the previous tile's values
          for (int i = r; i < BLOCK_N; i += Cfg::kEpiThreads) s_bias[i] = __ldg(e.bias + n_tile * BLOCK_N + i);
          epi_bar_sync(Cfg::kEpiThreads);
          bias_tile = n_tile;
        }
      }
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      if (r == 0 && item == static_cast<int>(blockIdx.x)) stamp(3);
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(wq * 32) << 16) + acc * Cfg::kAccCols + sub * BLOCK_N;
      // Software pipeline over column chunks, two chunks per (rolled) iteration: the TMEM load and, for the LSTM, the
      // cell-state load of the next chunk are in flight while a chunk is processed. The loop is NOT fully unrolled:
      // the straight-line epilogue of the 256-column tile was 37 KB of code and lost a third of its time to
      // instruction fetch (ncu stall_no_inst); the epilogue of this tile is exposed (single TMEM stage).
      constexpr int kChunks = BLOCK_N / CH;
      float v[2][CH];
      float cprev[2][8];
      auto issue = [&](int c, float* dst) {
        if constexpr (CH == 32) tmem_ld32(t_row + c * CH, dst); else tmem_ld16(t_row + c * CH, dst);
      };
      auto load_c = [&](int c, float* dst) {
        if constexpr (kLstm) lstm_load_c<kTrain>(g, e, b, y, x, valid, n_tile * BLOCK_N + c * CH, dst, ctile, BLOCK_M);
      };
      // row index inside the tile -> output pixel (the coalesced activation store addresses 8 neighbouring rows)
      auto row_of = [&](int rr, int& ob, int& oy, int& ox, bool& ov) {
        tile_row_to_pixel(g, grp, yb, rr, ob, oy, ox);
        ov = ob < g.B;
      };
      // Coalesced (warp-transposed) activation stores pay off where the epilogue is the bottleneck because the K loop
      // is short and TMEM is double-buffered (BLOCK_N <= 128: encoder.c2.x, decoder.upc2.2 / upc3.2 measured 10-20 %
      // faster); in the exposed epilogue of the 256-column tile the extra ~200 shuffle / select instructions per row
      // cost more issue slots than the stores save (measured 8-10 % slower), so it keeps the per-row stores.
      constexpr bool kCoalesced = (EPI == EPI_ACT && BLOCK_N <= 128);
      uint4 packed[kCoalesced ? 8 : 1];
      float gs[4] = {0.f, 0.f, 0.f, 0.f}, gq[4] = {0.f, 0.f, 0.f, 0.f};  // EPI_GATES: GroupNorm partial sums of this row
      auto process = [&](int c, const float* acc_v, const float* cp) {
        const int n0 = n_tile * BLOCK_N + c * CH;
        if constexpr (kCoalesced) {
          act_pack32(e, n0, acc_v, packed + (c & 1) * 4);
          if (c & 1) epi_act_store64(g, e, r, n0 - CH, packed, row_of);
        } else if constexpr (EPI == EPI_ACT) {
          epi_act<CH>(g, e, b, y, x, valid, n0, acc_v);
        }
        if constexpr (kLstm) epi_lstm<kTrain>(g, e, b, y, x, valid, n0, acc_v, cp, ctile, BLOCK_M, s_bias + c * CH);
        if constexpr (EPI == EPI_GAUSS) epi_gauss(g, e, b, y, x, valid, n0, acc_v);
        if constexpr (EPI == EPI_FRAME) epi_frame(g, e, b, y, x, valid, yb * (BLOCK_M / 32) + we, acc_v);
        if constexpr (EPI == EPI_F32 || EPI == EPI_F32_BT) {
          if constexpr (CH == 32) {
            if (g.epi_staged) {  // (uniform) coalesced stores through the warp's shared-memory tile
              epi_f32_staged<CH>(g, e, b, y, x, valid, n0, g.num_n_tiles * BLOCK_N, split, acc_v, s_stage, lane);
              return;
            }
          }
          if (kSplitK && e.split_part != nullptr) epi_split<CH>(g, e, b, y, x, valid, n0, g.num_n_tiles * BLOCK_N, split, acc_v);
          else epi_f32<CH>(g, e, b, y, x, valid, n0, acc_v);
        }
        if constexpr (EPI == EPI_GATES) epi_gates(g, e, b, y, x, valid, n0, acc_v, gs, gq);
      };
      if (!mine) {
        // nothing to read: still hand the accumulator stage back (the barrier counts every epilogue thread)
      } else if constexpr (kChunks == 1) {
        issue(0, v[0]);
        load_c(0, cprev[0]);
        tmem_ld_wait();
        process(0, v[0], cprev[0]);
      } else {
        issue(0, v[0]);
        load_c(0, cprev[0]);
        static_assert(kChunks == 1 || kChunks % 2 == 0, "column chunks are processed in pairs");
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
        if constexpr (EPI == EPI_GATES) epi_gates_finish(e, b, valid, n_tile, yb, gs, gq);
      }
      tc_fence_before();
      mbar_arrive(&tmem_empty[acc]);
      if (r == 0 && item == static_cast<int>(blockIdx.x)) stamp(4);
      if (++acc == Cfg::kNumAcc) { acc = 0; acc_phase ^= 1; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) stamp(5);
  if (warp == 2) tmem_dealloc(tmem_base, Cfg::kTmemCols);
}

// ---------------------------------------------------------------------------------------------------------------
// SIMT cross-check kernel: same geometry, same packed weights, same epilogues, no tensor cores / TMA. It exists so
// that the tcgen05 path can be validated tile by tile on the GPU (tests/, RAC_CONV_IMPL=simt); it is not a product
// path and nothing dispatches to it by default.
template <int BLOCK_M, int BLOCK_N, int EPI>
__global__ void __launch_bounds__(BLOCK_M)
conv_simt_kernel(const ConvRaw raw, const ConvGeom g, const EpiParams e) {
  constexpr int CH = (BLOCK_N >= 32) ? 32 : 16;
  constexpr int chunks = BLOCK_N / CH;
  const int c = blockIdx.x % chunks;
  const int tile = blockIdx.x / chunks;
  const int n_tile = tile / g.num_m_tiles;
  const int m_tile = tile - n_tile * g.num_m_tiles;
  const int grp = m_tile / g.tiles_per_img;
  const int yb = m_tile - grp * g.tiles_per_img;
  const int r = threadIdx.x;
  int b, y, x;
  tile_row_to_pixel(g, grp, yb, r, b, y, x);
  const bool valid = b < g.B;
  const int n0 = n_tile * BLOCK_N + c * CH;
  const int ktot = g.ks * g.ks * g.ctot;
  float acc[CH];
#pragma unroll
  for (int j = 0; j < CH; ++j) acc[j] = 0.f;
  if (valid) {
    for (int kh = 0; kh < g.ks; ++kh) {
      const int yy = y + kh - g.pad;
      if (yy < 0 || yy >= g.H) continue;
      for (int kw = 0; kw < g.ks; ++kw) {
        const int xx = x + kw - g.pad;
        if (xx < 0 || xx >= g.W) continue;
        int coff = 0;
        for (int s = 0; s < g.nsrc; ++s) {
          const int cs = g.src_kb[s] * kBlockK;
          if (g.src_dead[s]) { coff += cs; continue; }
          const __nv_bfloat16* ap = raw.src[s] + (static_cast<size_t>(b * g.H + yy) * g.W + xx) * cs;
          const __nv_bfloat16* wp = raw.w + static_cast<size_t>(n0) * ktot + (kh * g.ks + kw) * g.ctot + coff;
          for (int ch = 0; ch < cs; ++ch) {
            const float a = __bfloat162float(ap[ch]);
#pragma unroll
            for (int j = 0; j < CH; ++j) acc[j] += a * __bfloat162float(wp[static_cast<size_t>(j) * ktot + ch]);
          }
          coff += cs;
        }
      }
    }
  }
  if constexpr (EPI == EPI_ACT) epi_act<CH>(g, e, b, y, x, valid, n0, acc);
  if constexpr (EPI == EPI_LSTM || EPI == EPI_LSTM_TRAIN) {
    constexpr bool kTrain = (EPI == EPI_LSTM_TRAIN);
    float cprev[8];
    const size_t ctile = (static_cast<size_t>(m_tile) * (e.hid >> 3) * 2 * BLOCK_M + r) * 4;
    lstm_load_c<kTrain>(g, e, b, y, x, valid, n0, cprev, ctile, BLOCK_M);
    epi_lstm<kTrain>(g, e, b, y, x, valid, n0, acc, cprev, ctile, BLOCK_M, e.bias + n0);
  }
  if constexpr (EPI == EPI_GAUSS) epi_gauss(g, e, b, y, x, valid, n0, acc);
  if constexpr (EPI == EPI_FRAME) epi_frame(g, e, b, y, x, valid, yb * (BLOCK_M / 32) + (r >> 5), acc);
  if constexpr (EPI == EPI_F32) epi_f32<CH>(g, e, b, y, x, valid, n0, acc);
  if constexpr (EPI == EPI_GATES) {  // cross-check kernel: raw gates only (the 3-pass cell kernel computes its own statistics)
    float gs[4] = {}, gq[4] = {};
    epi_gates(g, e, b, y, x, valid, n0, acc, gs, gq);
  }
}

// ---------------------------------------------------------------------------------------------------------------
template <int BLOCK_M, int BLOCK_N, int EPI>
static cudaError_t launch_tc_t(const ConvOp& op, int num_sms, cudaStream_t stream) {
  using Cfg = TcCfg<BLOCK_M, BLOCK_N>;
  const int num_tiles = op.g.num_m_tiles * op.g.num_n_tiles * (op.g.ksplit > 1 ? op.g.ksplit : 1);
  const int grid = num_tiles < num_sms ? num_tiles : num_sms;
  return launch_pdl(conv_tc_kernel<BLOCK_M, BLOCK_N, EPI>, dim3(grid), dim3(Cfg::kThreads), Cfg::kSmemBytes, stream, op.tm, op.g, op.e);
}
template <int BLOCK_M, int BLOCK_N, int EPI>
static cudaError_t launch_simt_t(const ConvOp& op, cudaStream_t stream) {
  constexpr int CH = (BLOCK_N >= 32) ? 32 : 16;
  const int grid = op.g.num_m_tiles * op.g.num_n_tiles * (BLOCK_N / CH);
  conv_simt_kernel<BLOCK_M, BLOCK_N, EPI><<<grid, BLOCK_M, 0, stream>>>(op.raw, op.g, op.e);
  return cudaGetLastError();
}

// the (BLOCK_M, BLOCK_N, epilogue) instantiations the layer table uses (rac_api.cu::fill_specs / make_conv)
#define RAC_CONV_CASES(X)   \
  X(256, 256, EPI_ACT)      \
  X(256, 128, EPI_ACT)      \
  X(256, 64, EPI_ACT)       \
  X(256, 256, EPI_LSTM)     \
  X(256, 128, EPI_GAUSS)    \
  X(256, 16, EPI_FRAME)     \
  X(256, 256, EPI_F32)      \
  X(256, 256, EPI_GATES)    \
  X(256, 128, EPI_F32)      \
  X(256, 64, EPI_F32)       \
  X(128, 128, EPI_F32)      \
  X(128, 64, EPI_F32)       \
  X(128, 128, EPI_ACT)      \
  X(128, 64, EPI_ACT)       \
  X(128, 128, EPI_LSTM)     \
  X(256, 256, EPI_LSTM_TRAIN) \
  X(128, 128, EPI_LSTM_TRAIN) \
  X(128, 128, EPI_GAUSS)    \
  X(128, 16, EPI_FRAME)     \
  X(256, 256, EPI_F32_BT)   \
  X(256, 128, EPI_F32_BT)   \
  X(256, 64, EPI_F32_BT)    \
  X(128, 128, EPI_F32_BT)   \
  X(128, 64, EPI_F32_BT)

cudaError_t launch_conv_tc(const ConvOp& op, int num_sms, cudaStream_t stream) {
#define X(M, N, E) if (op.block_m == M && op.block_n == N && op.epi == E) return launch_tc_t<M, N, E>(op, num_sms, stream);
  RAC_CONV_CASES(X)
#undef X
  return cudaErrorInvalidValue;
}
cudaError_t launch_conv_simt(const ConvOp& op, cudaStream_t stream) {
  if (op.epi == EPI_F32_BT) return cudaErrorNotSupported;  // (the training step keeps the transposed copy with conv_impl = 1)
#define X(M, N, E) if (op.block_m == M && op.block_n == N && op.epi == E) return launch_simt_t<M, N, E>(op, stream);
  RAC_CONV_CASES(X)
#undef X
  return cudaErrorInvalidValue;
}
// ---------------------------------------------------------------------------------------------------------------
// Split-K reduction: out = sum over the ksplit slices (fixed order: bit-reproducible) + bias, routed to the fp32
// destination segments exactly as epi_f32 would have done. One thread per 4 columns of one row.
__global__ void __launch_bounds__(256)
splitk_reduce_kernel(const EpiParams e, int ksplit, long long rows, int ncols) {
  pdl_entry();
  const int q4 = ncols >> 2;
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= rows * q4) return;
  const long long row = i / q4;
  const int n = static_cast<int>(i - row * q4) * 4;
  if (n + 4 > e.cout) return;
  int s = 0;
  for (; s < e.nseg; ++s)
    if (n >= e.seg[s].n_begin && n < e.seg[s].n_end) break;
  if (s == e.nseg || e.seg[s].dst == nullptr) return;
  const F32Seg& sg = e.seg[s];
  const float* p = e.split_part + row * ncols + n;
  float4 v = *reinterpret_cast<const float4*>(p);
  for (int k = 1; k < ksplit; ++k) {
    const float4 o = *reinterpret_cast<const float4*>(p + k * e.split_stride);
    v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w;
  }
  if (e.bias) {
    const float4 b = __ldg(reinterpret_cast<const float4*>(e.bias + n));
    v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w;
  }
  float4* d = reinterpret_cast<float4*>(sg.dst + row * sg.cstride + sg.coff + (n - sg.n_begin));
  if (sg.accumulate) {
    const float4 o = *d;
    v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w;
  }
  *d = v;
}
cudaError_t launch_splitk_reduce(const EpiParams& e, int ksplit, long long rows, int ncols, cudaStream_t stream) {
  const long long total = rows * (ncols >> 2);
  if (cudaError_t e_ = launch_pdl_small(splitk_reduce_kernel, dim3(static_cast<unsigned>((total + 255) / 256)), dim3(256), 0, stream, e, ksplit, rows, ncols); e_ != cudaSuccess) return e_;
  return cudaGetLastError();
}

cudaError_t conv_tc_set_attributes() {
  cudaError_t err;
#define X(M, N, E)                                                                                         \
  if ((err = cudaFuncSetAttribute(conv_tc_kernel<M, N, E>, cudaFuncAttributeMaxDynamicSharedMemorySize,    \
                                  TcCfg<M, N>::kSmemBytes)) != cudaSuccess)                                \
    return err;
  RAC_CONV_CASES(X)
#undef X
  return cudaSuccess;
}

}  // namespace rac
