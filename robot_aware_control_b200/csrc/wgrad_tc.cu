// Weight gradient of a convolution as an IMPLICIT GEMM on tcgen05, straight from the NHWC tensors of the tape:
//
//   dWp[n][tap * ctot + c] = sum over (t, b, y, x) of  dY_t[(b, y, x)][n] * X_t[(b, y + kh - pad, x + kw - pad)][c]
//
// (reference: the `.backward()` of every nn.Conv2d inside PredictionTrainer._train_step, src/prediction/trainer.py:459-460;
// layers of vgg_64.py:8-18,87-121,196-221, lstm.py:121-127, dynamics.py:496-516). The contraction runs over the rows of
// the activation maps, which are the SLOW dimension of an NHWC tensor, so both MMA operands are "MN-major": a TMA box
// {64 channels, W, BH, NB, 1 step} lands in shared memory as rows of 128 B (64 channels of one position), and with the
// 128-byte swizzle that is exactly the canonical MN-major SWIZZLE_128B layout of tcgen05 (8 K-rows x 64 MN-elements per
// 1024-byte atom; SBO = 1024 B between groups of 8 positions, LBO = distance between two 64-channel boxes). Nothing is
// transposed or replicated in memory: the filter tap is a shift of the TMA box coordinates with out-of-bounds zero fill
// (the zero padding of the convolution), the sum over the time steps of BPTT is a fifth tensor-map dimension whose stride
// is the size of one tape step, and `h_{t-1}` of a ConvLSTM is the `h` tensor read at step coordinate t - 1 (t = 0 is out
// of bounds = the zero initial state). Round 1 materialised im2col(X)^T and dY^T in HBM instead (26 % of the step).
//
// One CTA = one output tile [256 packed output channels = two M128 accumulators] x [<= 256 input channels of one source]
// of one filter tap, for one slice of the contraction (split-K partials are summed in a fixed order by
// wgrad_reduce_kernel: deterministic). The tile is as large as TMEM allows (2 x 256 columns): the ConvLSTM weight
// gradients (K = 3840 rows only, 2048 x 1024 x 25 outputs) are bound by the L2 -> SM operand traffic, which a 256 x 256
// tile cuts by a third against 128 x 256.
//   warp 0: TMA producer, warp 1: MMA issuer (one thread), warp 2: TMEM allocator, warps 4-11: epilogue (TMEM -> fp32 global)
//
// MC = true (RAC_WGRAD_MC=1, off by default -- measured slower, see the pipeline-ring note below; layers with an even
// number of output-channel tiles AND of input-channel tiles): four CTAs of a cluster
// compute a 2 x 2 block of tiles of one filter tap -- rank bit 0 picks the output-channel tile, bit 1 the input-channel
// tile -- and share both operands through TMA multicast: a CTA loads every second box of its dY tile into itself and
// the CTA with the same output-channel tile (rank ^ 2), and every second box of its X tile into itself and the CTA with
// the same input-channel tile (rank ^ 1). L2 -> SM traffic per CTA and k-block drops from 8 boxes to 4. ncu of the
// plain kernel on the 5x5 gate layers (profiles/r02_train_top_ncu_s18.txt): 436 us for 403 GFLOP, sm__throughput 45 %,
// 800 CTAs x 80 k-blocks x 49 KB = 3.1 GB of operand reads in that time = 7.2 TB/s into the SMs.
//   * full barrier (per CTA): own arrive.expect_tx(all boxes); the bytes come from this CTA and its two peers
//   * empty barrier (per CTA): count 3 -- a stage is rewritten by this CTA and both peers, so the MMA threads of all
//     three release it (tcgen05.commit multicast to {self, rank ^ 1, rank ^ 2})
#include "ptx.cuh"
#include "wgrad_tc.cuh"

namespace rac {

namespace {

// Pipeline ring: 192 KB cut into stages of (A boxes + B boxes) x rows x 128 bytes, A = up to 4 boxes (256 output
// channels), B = up to 4 boxes (256 input channels), both the largest of the launch: 3 stages for a 256 x 256 tile with
// 64-row k-blocks, 4 with the 48-row k-blocks of the 6 x 8 ConvLSTM maps, 8 for the 64 x 128 tiles of the 48 x 64 maps.
// What bounds the main loop (0.85 us per 48-row k-block of a 256 x 256 tile against 0.42 us of nominal MMA time),
// measured on a B200 (profiles/r02_train_top_ncu_s18.txt, _s19.txt, r02_train_ab_s19/_s20/_s21.txt):
//   * no unit is saturated: L2 26 %, L2 -> SM fabric 21 %, shared-memory operand pipe 31 %, tensor pipe 29 % of peak;
//   * NOT the depth of the ring: 4 / 8 stages instead of a fixed 3 changed nothing (13.31 vs 13.31 ms per step);
//   * NOT the L2 reads: sharing both operands of a 2 x 2 block of tiles by TMA multicast (RAC_WGRAD_MC=1) halves them
//     and is SLOWER (436 -> 466 us on the 5x5 gate layers; the stage hand-back then waits for three CTAs): off;
//   * the up-to-8 TMA instructions per k-block cost a little when ONE thread issues them: two producer warps (dY boxes /
//     X boxes) gain 1 % of the training step (13.51 -> 13.38 ms), one lane per box nothing more: two warps it is;
//   * with the MMAs left out (RAC_WGRAD_EXP=1) the step is only 0.35 ms shorter, with the loads left out (=2) not
//     shorter at all: the load side and the MMA side (both operands MN-major) each need about the time the kernel takes.
constexpr int kMaxStages = 12;
constexpr int kMaxRows = 64;                       // positions per k-block (K of one pipeline stage)
constexpr int kRingBytes = 3 * 8 * kMaxRows * 128; // 192 KB
constexpr int kEpiWarps = 8;
constexpr int kSmemBytes = kRingBytes + 1024 + kEpiWarps * kStageWarpBytes + 1024;  // ring, barriers, transpose tiles, slack
static_assert(kSmemBytes <= 227 * 1024, "shared memory budget");
constexpr int kThreads = 128 + 32 * kEpiWarps;

// bf16 x bf16 -> fp32, A and B both MN-major, M = 128
__device__ __forceinline__ uint32_t umma_idesc_bf16_mn(int n) { return umma_idesc_bf16(n) | kIdescAMn | kIdescBMn; }

template <bool MC>
__global__ void __launch_bounds__(kThreads, 1)
wgrad_tc_kernel(const __grid_constant__ WgradTmaps tm, const WgradGeom g) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* bar_base = smem + kRingBytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(bar_base);
  uint64_t* empty_bar = full_bar + kMaxStages;
  uint64_t* tmem_full = empty_bar + kMaxStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // tile decode: blockIdx.x = (tap, c tile, n tile), blockIdx.y = K split; MC: blockIdx.x = (tap, c pair, n pair, rank)
  int n_tile, ct, tap;
  uint32_t rank = 0;
  if constexpr (MC) {
    rank = cluster_ctarank();
    int cid = blockIdx.x >> 2;
    const int np = g.n_tiles >> 1, cp = g.num_ctiles >> 1;
    n_tile = 2 * (cid % np) + static_cast<int>(rank & 1u);
    cid /= np;
    ct = 2 * (cid % cp) + static_cast<int>(rank >> 1);
    tap = cid / cp;
  } else {
    int t_id = blockIdx.x;
    n_tile = t_id % g.n_tiles;
    t_id /= g.n_tiles;
    ct = t_id % g.num_ctiles;
    tap = t_id / g.num_ctiles;
  }
  const uint16_t mask_a = static_cast<uint16_t>((1u << rank) | (1u << (rank ^ 2u)));  // CTAs that read my dY boxes
  const uint16_t mask_b = static_cast<uint16_t>((1u << rank) | (1u << (rank ^ 1u)));  // CTAs that read my X boxes
  const int kh = tap / g.ks, kw = tap - kh * g.ks;
  const int src = g.ct_src[ct], c0 = g.ct_c0[ct], cw = g.ct_w[ct];
  const int nb_boxes = cw >> 6;
  const int n0 = n_tile * 256;
  const int a_boxes = min(4, (g.kpad - n0) >> 6);  // output channels beyond kpad do not exist: their rows are never written
  const int subs = a_boxes > 2 ? 2 : 1;            // M128 accumulators in use
  const int kb_begin = blockIdx.y * g.kb_per_split;
  const int kb_end = min(kb_begin + g.kb_per_split, g.kb_total);
  const uint32_t box_bytes = static_cast<uint32_t>(g.rows) * 128u;
  const uint32_t stage_bytes = static_cast<uint32_t>(g.stage_a_boxes + g.stage_b_boxes) * box_bytes;
  const int kStages = min(g.max_stages > 0 ? g.max_stages : kMaxStages, static_cast<int>(kRingBytes / stage_bytes));

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm.dy);
    tma_prefetch_desc(&tm.x[src]);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], MC ? 3 : 1);
    }
    mbar_init(tmem_full, 1);
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (MC) cluster_sync_all();  // the peers' barriers are initialised before anything is signalled on them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  griddep_launch_dependents();  // (PDL, ptx.cuh) the next kernel's prologue may start
  griddep_wait();               // everything below reads / writes global memory of earlier kernels

  const int pmode = MC ? 1 : g.producers;
  const int plane = pmode == 3 ? 4 : 1;  // producer lanes per producer warp
  if ((warp == 0 || (warp == 3 && pmode >= 2)) && lane < plane) {
    // ===================== TMA producer(s) =====================
    // one cp.async.bulk.tensor per 64-channel box (the 128-byte swizzle caps the inner box dimension): up to 8 per
    // k-block; pmode says how many threads share them
    const bool do_a = warp == 0, do_b = pmode == 1 || warp == 3;
    int stage = 0;
    uint32_t phase = 0;
    for (int kb = kb_begin; kb < kb_end; ++kb) {
      // k-block -> (step, candidate group, row group)
      const int hg = kb % g.hgroups;
      const int rest = kb / g.hgroups;
      const int bg = rest % g.bgroups;
      const int t = rest / g.bgroups;
      const int y0 = hg * g.BH, b0 = bg * g.NB;
      mbar_wait(&empty_bar[stage], phase ^ 1);
      uint8_t* sa = smem + stage * stage_bytes;
      uint8_t* sb = sa + g.stage_a_boxes * box_bytes;
      if (g.experiment == 2) {  // (timing experiment: no loads, the MMAs run on whatever the ring holds)
        if (warp == 0 && lane == 0) mbar_arrive(&full_bar[stage]);
      } else {
        if (warp == 0 && lane == 0) mbar_arrive_expect_tx(&full_bar[stage], (a_boxes + nb_boxes) * box_bytes);
        if constexpr (MC) {
          // every second box of each operand, into this CTA and the peer that shares the operand
          for (int j = static_cast<int>(rank >> 1); j < a_boxes; j += 2)
            tma_load_5d_mc(&tm.dy, &full_bar[stage], sa + j * box_bytes, n0 + j * 64, 0, y0, b0, t, mask_a);
          for (int j = static_cast<int>(rank & 1u); j < nb_boxes; j += 2)
            tma_load_5d_mc(&tm.x[src], &full_bar[stage], sb + j * box_bytes, c0 + j * 64, kw - g.pad, y0 + kh - g.pad, b0,
                           t - g.src_tshift[src], mask_b);
        } else {
          if (do_a)
            for (int j = lane; j < a_boxes; j += plane)
              tma_load_5d(&tm.dy, &full_bar[stage], sa + j * box_bytes, n0 + j * 64, 0, y0, b0, t);
          if (do_b)
            for (int j = lane; j < nb_boxes; j += plane)
              tma_load_5d(&tm.x[src], &full_bar[stage], sb + j * box_bytes, c0 + j * 64, kw - g.pad, y0 + kh - g.pad, b0,
                          t - g.src_tshift[src]);
        }
      }
      if (++stage == kStages) { stage = 0; phase ^= 1; }
    }
  } else if (warp == 1 && lane == 0) {
    // ===================== MMA issuer =====================
    const uint32_t idesc = umma_idesc_bf16_mn(cw);
    int stage = 0;
    uint32_t phase = 0;
    uint32_t accumulate = 0;
    for (int kb = kb_begin; kb < kb_end; ++kb) {
      mbar_wait(&full_bar[stage], phase);
      tc_fence_after();
      const uint32_t sa = smem_u32(smem + stage * stage_bytes);
      const uint32_t sb = sa + g.stage_a_boxes * box_bytes;
      const uint64_t adesc = umma_desc_sw128_mn(sa, box_bytes);
      const uint64_t bdesc = umma_desc_sw128_mn(sb, box_bytes);
      for (int k = 0; k < g.rows / 16; ++k) {
        // 16 positions = two 1024-byte atoms = 2048 B further along K: + 128 in the (addr >> 4) field; the second
        // accumulator's 128 output channels are the A boxes 2 and 3 (2 * box_bytes further)
        for (int sub = 0; sub < subs && g.experiment != 1; ++sub)
          umma_bf16_ss(tmem_base + sub * 256, adesc + 128u * k + sub * ((2u * box_bytes) >> 4), bdesc + 128u * k, idesc,
                       accumulate);
        accumulate = 1;
      }
      if constexpr (MC) umma_commit_mc(&empty_bar[stage], static_cast<uint16_t>(mask_a | mask_b));
      else umma_commit(&empty_bar[stage]);
      if (kb == kb_end - 1) umma_commit(tmem_full);
      if (++stage == kStages) { stage = 0; phase ^= 1; }
    }
  } else if (warp >= 4) {
    // ===================== epilogue: TMEM lane = output channel n, column = input channel c =====================
    const int q = (warp - 4) & 3;   // TMEM lanes 32 q .. 32 q + 31 (a warp may only touch the lane quarter warp % 4)
    const int sub = (warp - 4) >> 2;  // accumulator: output channels n0 + 128 sub ..
    const int n = n0 + sub * 128 + q * 32 + lane;
    const long long ld = static_cast<long long>(g.taps) * g.ctot;
    float* out = g.out + static_cast<long long>(blockIdx.y) * g.out_split_stride + static_cast<long long>(n) * ld +
                 static_cast<long long>(tap) * g.ctot + g.src_coff[src] + c0;
    if (kb_end > kb_begin) {
      mbar_wait(tmem_full, 0);
      tc_fence_after();
    }
    // g.epi_staged: the 32 x 32 chunk goes through this warp's shared-memory transpose tile and leaves as 64 contiguous
    // bytes of eight rows per store instruction (ptx.cuh, warp_store_rows_f32_v4) instead of 16 bytes of 32 rows
    float* stage = reinterpret_cast<float*>(bar_base + 1024 + (warp - 4) * kStageWarpBytes);
    const long long row_off = n < g.kpad ? static_cast<long long>(n) * ld : -1;
    float* base = out - static_cast<long long>(n) * ld;  // row 0 of this (split, tap, source, c0) block
    for (int cc = 0; cc < cw && sub < subs; cc += 32) {
      float v[32];
      if (kb_end > kb_begin) {
        tmem_ld32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + sub * 256 + cc, v);
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = 0.f;
      }
      if (g.epi_staged) {  // (uniform)
        warp_store_rows_f32_v4(stage, base + cc, row_off, v, nullptr, false, lane);
      } else if (n < g.kpad) {
#pragma unroll
        for (int i = 0; i < 32; i += 4)
          *reinterpret_cast<float4*>(out + cc + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (MC) cluster_sync_all();  // the peers may still multicast into this CTA's shared memory / signal its barriers
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// out[i] = sum over splits (in order) of part[s][i]
__global__ void __launch_bounds__(256)
wgrad_reduce_kernel(const float4* __restrict__ part, int splits, long long n4, long long stride4, float4* __restrict__ out) {
  pdl_entry();
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  float4 a = part[i];
  for (int s = 1; s < splits; ++s) {
    const float4 b = part[s * stride4 + i];
    a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
  }
  out[i] = a;
}

}  // namespace

cudaError_t wgrad_tc_set_attributes() {
  cudaError_t err = cudaFuncSetAttribute(wgrad_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
  if (err != cudaSuccess) return err;
  return cudaFuncSetAttribute(wgrad_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
}

// RAC_WGRAD_MC=1: 2 x 2 multicast clusters where the tile counts allow (off by default: measured slower, see above)
static bool wgrad_mc_enabled() {
  static const int on = [] { const char* v = getenv("RAC_WGRAD_MC"); return v ? atoi(v) : 0; }();
  return on != 0;
}

cudaError_t launch_wgrad_tc(const WgradTmaps& tm, const WgradGeom& g_in, cudaStream_t s) {
  // RAC_WGRAD_MAX_STAGES=3: the fixed three stages of the first version (A/B switch)
  static const int max_stages = [] { const char* v = getenv("RAC_WGRAD_MAX_STAGES"); return v ? atoi(v) : 0; }();
  WgradGeom g = g_in;
  g.max_stages = max_stages;
  static const int producers = [] { const char* v = getenv("RAC_WGRAD_PRODUCERS"); return v ? atoi(v) : 2; }();
  static const int experiment = [] { const char* v = getenv("RAC_WGRAD_EXP"); return v ? atoi(v) : 0; }();
  // transposed 128-bit epilogue stores (see the kernel's epilogue): 13.22 -> 12.80 ms per training step
  // (profiles/r02_train_ab_s25.txt); RAC_WGRAD_EPI_STAGED=0: the direct per-row stores (A/B switch)
  static const int epi_staged = [] { const char* v = getenv("RAC_WGRAD_EPI_STAGED"); return v ? atoi(v) : 1; }();
  g.epi_staged = epi_staged != 0;
  g.producers = producers < 1 || producers > 3 ? 1 : producers;
  g.experiment = experiment;
  // a stage holds the largest tile pair of the launch, not always 4 + 4 boxes: the 64- / 128-channel layers of the
  // 48 x 64 maps (3 boxes of 8 KB per k-block) get 8 stages in flight instead of 3
  g.stage_a_boxes = g.kpad >= 256 ? 4 : g.kpad / 64;
  g.stage_b_boxes = 1;
  for (int i = 0; i < g.num_ctiles && i < kWgMaxCTiles; ++i) g.stage_b_boxes = g.ct_w[i] / 64 > g.stage_b_boxes ? g.ct_w[i] / 64 : g.stage_b_boxes;
  if (max_stages > 0) { g.stage_a_boxes = 4; g.stage_b_boxes = 4; }  // (A/B: the first version's fixed 8-box stages)
  if (g.stage_a_boxes < 1 || g.stage_b_boxes > 4) return cudaErrorInvalidValue;
  if (g.rows % 16 || g.rows > kMaxRows || g.num_ctiles < 1 || g.num_ctiles > kWgMaxCTiles) return cudaErrorInvalidValue;
  dim3 grid(static_cast<unsigned>(g.n_tiles * g.num_ctiles * g.taps), static_cast<unsigned>(g.splits));
  if (wgrad_mc_enabled() && g.n_tiles % 2 == 0 && g.num_ctiles % 2 == 0) {
    // (same CTA count; blockIdx.x is decoded as (tap, c pair, n pair, rank) instead)
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = dim3(kThreads); cfg.dynamicSmemBytes = kSmemBytes; cfg.stream = s;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 4; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 2 : 1;
    return cudaLaunchKernelEx(&cfg, wgrad_tc_kernel<true>, tm, g);
  }
  return launch_pdl(wgrad_tc_kernel<false>, grid, dim3(kThreads), kSmemBytes, s, tm, g);
}

cudaError_t launch_wgrad_reduce(const float* part, int splits, long long n, long long split_stride, float* out,
                                cudaStream_t s) {
  if ((n & 3) || (split_stride & 3)) return cudaErrorInvalidValue;
  const long long n4 = n / 4;
  if (cudaError_t e_ = launch_pdl_small(wgrad_reduce_kernel, dim3(static_cast<unsigned>((n4 + 255) / 256)), dim3(256), 0, s, 
      reinterpret_cast<const float4*>(part), splits, n4, split_stride / 4, reinterpret_cast<float4*>(out)); e_ != cudaSuccess) return e_;
  return cudaGetLastError();
}

}  // namespace rac
