// Thin inline-PTX wrappers for sm_100a: mbarrier, TMA (cp.async.bulk.tensor), tcgen05 (MMA / TMEM).
// Everything here is hand-written for B200; nothing is shared with other architectures.
#pragma once
#include <stdlib.h>
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>

namespace rac {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ---------------------------------------------------------------- mbarrier
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug must become a launch failure, never a hung GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000ll) { __trap(); }  // ~2 s at 1.9 GHz: far beyond any legitimate wait
  }
}

// named barrier 1 among the epilogue warps only (the TMA / MMA warps never join it)
__device__ __forceinline__ void epi_bar_sync(int nthreads) {
  asm volatile("bar.sync 1, %0;" ::"r"(nthreads) : "memory");
}

// ---------------------------------------------------------------- TMA
// Programmatic dependent launch (PDL): a kernel launched with the programmatic-stream-serialization attribute may
// start while its predecessor in the stream is still running, as soon as every CTA of the predecessor has executed
// launch_dependents (or exited); it must execute griddep_wait() before it touches anything the predecessor (or any
// earlier kernel) reads or writes -- the wait returns when the predecessor has completed and its memory is visible.
// The tcgen05 kernels put their whole prologue (tensor-map prefetch, mbarrier init, TMEM allocation, cluster sync)
// in front of the wait: a few microseconds per launch that now overlap the tail of the previous kernel.
__device__ __forceinline__ void griddep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void griddep_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
// The small (non-tcgen05) kernels have no prologue worth overlapping; for them the point is the launch latency itself:
// trigger + wait as the first statement lets the NEXT kernel's CTAs become resident (and, for a tcgen05 kernel, run
// their prologue) during this kernel's last wave, and lets this kernel be resident before its predecessor has drained.
// (The trigger only takes effect once EVERY CTA of the grid has executed it, i.e. when the last wave has started, so
// early dependents never take SM slots from CTAs of this grid that have not started yet. No TMEM is held here.)
__device__ __forceinline__ void pdl_entry() {
  griddep_launch_dependents();
  griddep_wait();
}

__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* m, uint64_t* bar, void* dst, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(const CUtensorMap* m, uint64_t* bar, void* dst, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
// L2 prefetch of a tensor box (no shared-memory destination, no completion tracking)
__device__ __forceinline__ void tma_prefetch_l2_2d(const CUtensorMap* m, int c0, int c1) {
  asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];"
               ::"l"(reinterpret_cast<uint64_t>(m)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_prefetch_l2_3d(const CUtensorMap* m, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global.tile [%0, {%1, %2, %3}];"
               ::"l"(reinterpret_cast<uint64_t>(m)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tma_load_5d(const CUtensorMap* m, uint64_t* bar, void* dst, int c0, int c1, int c2,
                                            int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], "
      "[%2];" ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d(const CUtensorMap* m, uint64_t* bar, void* dst, int c0, int c1, int c2,
                                            int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], "
      "[%2];" ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

// ---------------------------------------------------------------- tcgen05 / TMEM
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem] * B[smem]; bf16 operands, fp32 accumulate, issued by ONE thread for the CTA.
__device__ __forceinline__ void umma_bf16_ss(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Arrive on an mbarrier when all previously issued tcgen05.mma of this thread have completed.
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

// TMEM -> registers: 32 lanes x 32 consecutive fp32 columns (one row per thread of the warp).
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---------------------------------------------------------------- CTA pair (cta_group::2) variants
// Two CTAs of a cluster (same TPC) execute one M = 256 MMA: each holds 128 rows of A, half of the B rows and its half
// of the accumulator in its own SMEM / TMEM at IDENTICAL offsets; the leader (cluster rank 0) issues the MMA.
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;  // shared::cluster address of "my" offset in the even (leader) CTA
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive (count 1) on the LEADER CTA's barrier that sits at the same offset as `bar`
__device__ __forceinline__ void mbar_arrive_leader(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(smem_u32(bar) & kPeerBitMask) : "memory");
}
// TMA loads of a CTA pair: the bytes land in THIS CTA's shared memory, the transaction count on the leader's barrier
__device__ __forceinline__ void tma_load_2d_2sm(const CUtensorMap* m, uint64_t* bar, void* dst, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar) & kPeerBitMask), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d_2sm(const CUtensorMap* m, uint64_t* bar, void* dst, int c0, int c1, int c2,
                                                int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], "
      "[%2];" ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar) & kPeerBitMask), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish_2sm() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem of both CTAs] (+)= A[smem of both] * B[smem of both]; M = 256. Issued by ONE thread of the leader CTA.
__device__ __forceinline__ void umma_bf16_ss_2sm(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                                 uint32_t accumulate) {
  const uint32_t z = 0;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate), "r"(z)
      : "memory");
}
// arrive on the barrier at this offset in BOTH CTAs once all previously issued MMAs of this thread have completed
__device__ __forceinline__ void umma_commit_2sm(uint64_t* bar) {
  const uint16_t mask = 3;
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   smem_u32(bar)),
               "h"(mask)
               : "memory");
}
__host__ __device__ constexpr uint32_t umma_idesc_bf16_m256(int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(n >> 3) << 17) |
         (static_cast<uint32_t>(256 >> 4) << 24);
}

// ---------------------------------------------------------------- cluster multicast (cta_group::1 MMAs)
// TMA load whose box lands at the same shared-memory offset in every CTA of `mask` and completes transaction bytes on
// the barrier at the same offset in each of them
__device__ __forceinline__ void tma_load_4d_mc(const CUtensorMap* m, uint64_t* bar, void* dst, int c0, int c1, int c2,
                                               int c3, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster"
      " [%0], [%1, {%4, %5, %6, %7}], [%2], %3;" ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "h"(mask), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_5d_mc(const CUtensorMap* m, uint64_t* bar, void* dst, int c0, int c1, int c2,
                                               int c3, int c4, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster"
      " [%0], [%1, {%4, %5, %6, %7, %8}], [%2], %3;" ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "h"(mask), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
// arrive on the barrier at this offset in every CTA of `mask` once the MMAs issued so far by this thread completed
__device__ __forceinline__ void umma_commit_mc(uint64_t* bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   smem_u32(bar)),
               "h"(mask)
               : "memory");
}

// UMMA shared-memory matrix descriptor for a K-major bf16 tile stored as rows of 128 B (64 channels) with the
// 128-byte swizzle TMA writes: 8-row groups are 1024 B apart (SBO), LBO is unused for swizzled K-major layouts.
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr >> 4) & 0x3FFF);  // start address, bits [0,14)
  d |= static_cast<uint64_t>(1) << 16;                    // leading byte offset (ignored), bits [16,30)
  d |= static_cast<uint64_t>(1024 >> 4) << 32;            // stride byte offset, bits [32,46)
  d |= static_cast<uint64_t>(1) << 46;                    // descriptor version (Blackwell)
  d |= static_cast<uint64_t>(2) << 61;                    // layout type: SWIZZLE_128B
  return d;
}
// MN-major SWIZZLE_128B operand (the contraction index is the SLOW dimension in memory): a TMA box {64 elements of M/N,
// rows of K} lands as rows of 128 B = the canonical layout ((8, n), (8, k)) : ((1, LBO), (8, SBO)) in 16-byte units --
// 8 K-rows x 64 MN-elements per 1024-byte atom, SBO = 1024 B between groups of 8 K-rows, LBO = distance between two
// 64-element MN groups (one box each). 16 K-rows (one MMA) = 2048 B: +128 in the start-address field.
__device__ __forceinline__ uint64_t umma_desc_sw128_mn(uint32_t smem_addr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr >> 4) & 0x3FFF);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= static_cast<uint64_t>(1024 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}
constexpr uint32_t kIdescAMn = 1u << 15;  // instruction descriptor: A operand MN-major
constexpr uint32_t kIdescBMn = 1u << 16;  // instruction descriptor: B operand MN-major
// Instruction descriptor: bf16 x bf16 -> fp32, both operands K-major, M = 128.
__host__ __device__ constexpr uint32_t umma_idesc_bf16(int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(n >> 3) << 17) |
         (static_cast<uint32_t>(128 >> 4) << 24);
}

// Host side: kernel launch with (pdl) or without the programmatic-stream-serialization attribute. RAC_PDL=0 turns it off.
inline bool pdl_enabled() {
  static const int on = [] { const char* v = getenv("RAC_PDL"); return v ? atoi(v) : 1; }();
  return on != 0;
}
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
// The same for the small kernels that begin with pdl_entry(); RAC_PDL_SMALL=0 launches those the plain way (A/B switch).
inline bool pdl_small_enabled() {
  static const int on = [] { const char* v = getenv("RAC_PDL_SMALL"); return v ? atoi(v) : 1; }();
  return on != 0 && pdl_enabled();
}
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl_small(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                                    Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_small_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// ---------------------------------------------------------------- transposed fp32 stores (GEMM epilogues)
// Warp-collective store of 32 rows x 32 fp32 columns through a per-warp shared-memory tile (conv_tc_kernel's fp32
// epilogues). TMEM hands every lane one ROW of the accumulator, so the direct store (epi_f32 / epi_split) writes 16 bytes
// per lane to 32 different rows per instruction -- 32 memory requests of half a sector each; the epilogue of a 256 x 256
// tile (256 KB) took 9.4 us that way, 22 % of the time of the training GEMMs (profiles/r02_train_timeline_s10.txt).
// Here a 16-column half of the chunk goes to shared memory row by row (row stride 20 floats: the 128-bit stores of a
// quarter warp hit 8 disjoint bank groups) and comes back TRANSPOSED: every store instruction writes 64 contiguous
// bytes of two rows. Row offsets (elements from `base`, < 0 = do not write) travel through the tile's spare columns.
constexpr int kStageRowFloats = 20;
constexpr int kStageWarpBytes = 32 * kStageRowFloats * 4;  // 2560 B per epilogue warp
__device__ __forceinline__ void warp_store_rows_f32(float* stage, float* base, long long row_off, const float* acc,
                                                    const float* bias, bool accumulate, int lane) {
  float* mine = stage + lane * kStageRowFloats;
  const int rsel = lane >> 4, col = lane & 15;
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    __syncwarp();  // the previous half has been read
#pragma unroll
    for (int q = 0; q < 4; ++q)
      *reinterpret_cast<float4*>(mine + 4 * q) =
          make_float4(acc[16 * h + 4 * q], acc[16 * h + 4 * q + 1], acc[16 * h + 4 * q + 2], acc[16 * h + 4 * q + 3]);
    if (h == 0) *reinterpret_cast<long long*>(mine + 16) = row_off;
    __syncwarp();
    const float bv = bias ? __ldg(bias + 16 * h + col) : 0.f;
#pragma unroll 4
    for (int i = 0; i < 16; ++i) {
      const float* row = stage + (2 * i + rsel) * kStageRowFloats;
      const long long off = *reinterpret_cast<const long long*>(row + 16);
      if (off < 0) continue;
      float* p = base + off + 16 * h + col;
      float v = row[col] + bv;
      if (accumulate) v += *p;
      *p = v;
    }
  }
}
// Same tile, 128-bit on the way back as well (g.epi_staged == 2): four lanes read one row's 16 columns as float4, so a
// store instruction writes the 64 contiguous bytes of EIGHT rows -- 8 LDS.128 + 8 STG.128 per 32-column chunk and lane
// instead of 32 scalar pairs. A quarter warp reads rows r and r + 4 of the tile: with the 80-byte row stride their
// 16-byte bank groups {5r .. 5r+3} and {5r+20 .. 5r+23} (mod 8) are disjoint.
__device__ __forceinline__ void warp_store_rows_f32_v4(float* stage, float* base, long long row_off, const float* acc,
                                                       const float* bias, bool accumulate, int lane) {
  float* mine = stage + lane * kStageRowFloats;
  const int rsel = (lane >> 3) + 4 * ((lane >> 2) & 1), c4 = (lane & 3) * 4;
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    __syncwarp();  // the previous half has been read
#pragma unroll
    for (int q = 0; q < 4; ++q)
      *reinterpret_cast<float4*>(mine + 4 * q) =
          make_float4(acc[16 * h + 4 * q], acc[16 * h + 4 * q + 1], acc[16 * h + 4 * q + 2], acc[16 * h + 4 * q + 3]);
    if (h == 0) *reinterpret_cast<long long*>(mine + 16) = row_off;
    __syncwarp();
    float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
    if (bias) {
      const float* bp = bias + 16 * h + c4;
      bv = make_float4(__ldg(bp), __ldg(bp + 1), __ldg(bp + 2), __ldg(bp + 3));
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float* row = stage + (8 * i + rsel) * kStageRowFloats;
      const long long off = *reinterpret_cast<const long long*>(row + 16);
      if (off < 0) continue;
      float4* p = reinterpret_cast<float4*>(base + off + 16 * h + c4);
      float4 v = *reinterpret_cast<const float4*>(row + c4);
      v.x += bv.x; v.y += bv.y; v.z += bv.z; v.w += bv.w;
      if (accumulate) {
        const float4 o = *p;
        v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w;
      }
      *p = v;
    }
  }
}

}  // namespace rac
