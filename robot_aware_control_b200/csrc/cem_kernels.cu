// CEM sample / elite top-k / Gaussian refit, all on device (reference src/cem/cem.py:76-104).
#include "misc_kernels.cuh"
#include "epilogue.cuh"

namespace rac {

// ---------------------------------------------------------------- sample
// act_seq = clamp(mean + std * n(0,1), +-clamp); at iteration 0 the LAST candidate is the do-nothing sequence
// (cem.py:80-86). act2: every rank keeps all N candidates' 2-D actions (needed by the replicated refit);
// act5: zero-padded model actions (cem.py:86) for this rank's shard only.
__global__ void __launch_bounds__(256)
cem_sample_kernel(const float* __restrict__ mean, const float* __restrict__ stdv, const float* __restrict__ noise,
                  unsigned long long seed, int iter, int n_total, int L, int adim, int cand_offset, int n_local,
                  float clampv, float* __restrict__ act2, float* __restrict__ act5) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;  // (candidate, step)
  if (i >= n_total * L) return;
  const int n = i / L, t = i - n * L;
  float z0, z1;
  if (noise) {
    z0 = noise[static_cast<size_t>(i) * 2];
    z1 = noise[static_cast<size_t>(i) * 2 + 1];
  } else {
    const Philox4 r = philox4x32_10(static_cast<uint32_t>(n), static_cast<uint32_t>(t), static_cast<uint32_t>(iter),
                                    0xce3u, static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32));
    box_muller(r.v[0], r.v[1], z0, z1);
  }
  // separate multiply and add (no FMA contraction): bit-identical to torch's `mean + std * noise` on the host
  float a0 = __fadd_rn(mean[t * 2], __fmul_rn(stdv[t * 2], z0));
  float a1 = __fadd_rn(mean[t * 2 + 1], __fmul_rn(stdv[t * 2 + 1], z1));
  if (iter == 0 && n == n_total - 1) a0 = a1 = 0.f;
  a0 = fminf(fmaxf(a0, -clampv), clampv);
  a1 = fminf(fmaxf(a1, -clampv), clampv);
  act2[static_cast<size_t>(i) * 2] = a0;
  act2[static_cast<size_t>(i) * 2 + 1] = a1;
  const int nl = n - cand_offset;
  if (nl >= 0 && nl < n_local) {
    float* d = act5 + (static_cast<size_t>(nl) * L + t) * adim;
    for (int c = 0; c < adim; ++c) d[c] = (c == 0) ? a0 : (c == 1 ? a1 : 0.f);
  }
}
cudaError_t launch_cem_sample(const float* mean, const float* stdv, const float* noise, unsigned long long seed,
                              int iter, int n_total, int L, int adim_model, int cand_offset, int n_local,
                              float clampv, float* act2, float* act5, cudaStream_t s) {
  const int total = n_total * L;
  cem_sample_kernel<<<(total + 255) / 256, 256, 0, s>>>(mean, stdv, noise, seed, iter, n_total, L, adim_model,
                                                        cand_offset, n_local, clampv, act2, act5);
  return cudaGetLastError();
}

// ---------------------------------------------------------------- top-k (K largest, ties -> lowest index)
// Matches torch.topk(costs, K) (cem.py:97) as a set whenever the K-th and (K+1)-th values differ and equals
// torch.sort(descending, stable)[:K] on ties; output order is (value descending, index ascending).
// Single CTA: 8-pass MSB radix select on order-preserving 64-bit keys, index-ordered compaction, bitonic sort.
constexpr int kTopkThreads = 1024;
constexpr int kTopkMaxK = 4096;

__device__ __forceinline__ unsigned long long cost_key(double v) {
  v = v + 0.0;  // -0.0 -> +0.0 so that equal values give equal keys
  unsigned long long u = static_cast<unsigned long long>(__double_as_longlong(v));
  return (u >> 63) ? ~u : (u | 0x8000000000000000ull);
}

__global__ void __launch_bounds__(kTopkThreads)
topk_kernel(const double* __restrict__ costs, int n, int k, long long* __restrict__ idx_out,
            double* __restrict__ val_out) {
  extern __shared__ unsigned char topk_smem[];
  unsigned long long* skey = reinterpret_cast<unsigned long long*>(topk_smem);  // [kpad]
  int* sidx = reinterpret_cast<int*>(skey + kTopkMaxK);                         // [kpad]
  __shared__ unsigned int hist[256];
  __shared__ unsigned long long s_prefix;
  __shared__ int s_krem;
  __shared__ int warp_gt[32], warp_eq[32];
  __shared__ int run_gt, run_eq;
  const int tid = threadIdx.x;

  if (tid == 0) { s_prefix = 0ull; s_krem = k; }
  __syncthreads();
  // ---- radix select of the K-th largest key
  for (int pass = 0; pass < 8; ++pass) {
    const int shift = 56 - 8 * pass;
    if (tid < 256) hist[tid] = 0u;
    __syncthreads();
    const unsigned long long prefix = s_prefix;
    const unsigned long long himask = (pass == 0) ? 0ull : (~0ull << (shift + 8));
    for (int i = tid; i < n; i += kTopkThreads) {
      const unsigned long long key = cost_key(costs[i]);
      if (((key ^ prefix) & himask) == 0ull) atomicAdd(&hist[(key >> shift) & 255u], 1u);
    }
    __syncthreads();
    if (tid == 0) {
      int krem = s_krem;
      int above = 0;
      int d = 255;
      for (; d > 0; --d) {
        const int h = static_cast<int>(hist[d]);
        if (above + h >= krem) break;
        above += h;
      }
      s_krem = krem - above;
      s_prefix = prefix | (static_cast<unsigned long long>(d) << shift);
    }
    __syncthreads();
  }
  const unsigned long long T = s_prefix;
  const int need_eq = s_krem;  // how many keys == T to take (lowest indices first)
  const int count_gt = k - need_eq;

  // ---- compaction in index order
  int kpad = 1;
  while (kpad < k) kpad <<= 1;
  for (int i = tid; i < kpad; i += kTopkThreads) { skey[i] = 0ull; sidx[i] = 0x7fffffff; }
  if (tid == 0) { run_gt = 0; run_eq = 0; }
  __syncthreads();
  const int lane = tid & 31, warp = tid >> 5;
  for (int base = 0; base < n; base += kTopkThreads) {
    const int i = base + tid;
    unsigned long long key = 0ull;
    bool gt = false, eq = false;
    if (i < n) {
      key = cost_key(costs[i]);
      gt = key > T;
      eq = key == T;
    }
    const unsigned bg = __ballot_sync(0xffffffffu, gt), be = __ballot_sync(0xffffffffu, eq);
    const unsigned lt_mask = (1u << lane) - 1u;
    const int pg = __popc(bg & lt_mask), pe = __popc(be & lt_mask);
    if (lane == 0) { warp_gt[warp] = __popc(bg); warp_eq[warp] = __popc(be); }
    __syncthreads();
    int og = run_gt, oe = run_eq;
    for (int w = 0; w < warp; ++w) { og += warp_gt[w]; oe += warp_eq[w]; }
    if (gt) {
      skey[og + pg] = key;
      sidx[og + pg] = i;
    } else if (eq) {
      const int rank = oe + pe;
      if (rank < need_eq) {
        skey[count_gt + rank] = key;
        sidx[count_gt + rank] = i;
      }
    }
    __syncthreads();
    if (tid == 0) {
      int tg = 0, te = 0;
      for (int w = 0; w < kTopkThreads / 32; ++w) { tg += warp_gt[w]; te += warp_eq[w]; }
      run_gt += tg;
      run_eq += te;
    }
    __syncthreads();
  }
  // ---- bitonic sort: key descending, index ascending (padding entries have key 0 / idx INT_MAX -> sink to the end)
  for (int size = 2; size <= kpad; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int i = tid; i < kpad; i += kTopkThreads) {
        const int j = i ^ stride;
        if (j > i) {
          const unsigned long long ki = skey[i], kj = skey[j];
          const int ii = sidx[i], ij = sidx[j];
          const bool i_first = (ki > kj) || (ki == kj && ii < ij);  // i belongs before j in the final order
          const bool up = (i & size) == 0;
          if (up ? !i_first : i_first) {
            skey[i] = kj; skey[j] = ki;
            sidx[i] = ij; sidx[j] = ii;
          }
        }
      }
      __syncthreads();
    }
  }
  for (int i = tid; i < k; i += kTopkThreads) {
    idx_out[i] = sidx[i];
    if (val_out) val_out[i] = costs[sidx[i]];
  }
}
cudaError_t cem_set_attributes() {
  return cudaFuncSetAttribute(topk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kTopkMaxK * 12);
}
cudaError_t launch_topk(const double* costs, int n, int k, int64_t* idx_out, double* val_out, cudaStream_t s) {
  if (k < 1 || k > n || k > kTopkMaxK) return cudaErrorInvalidValue;
  topk_kernel<<<1, kTopkThreads, kTopkMaxK * 12, s>>>(costs, n, k, reinterpret_cast<long long*>(idx_out), val_out);
  return cudaGetLastError();
}

// ---------------------------------------------------------------- refit
// std, mean = torch.std_mean(top_act_seq, dim=0) (unbiased); std = max(std, floor) (cem.py:101-104).
// One warp per (step, action-dim) column; fp64 accumulation.
__global__ void __launch_bounds__(1024)
refit_kernel(const float* __restrict__ act2, int L2, const long long* __restrict__ idx, int k, float std_floor,
             float* __restrict__ mean_out, float* __restrict__ std_out) {
  const int col = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (col >= L2) return;
  double s = 0.0;
  for (int i = lane; i < k; i += 32) s += static_cast<double>(act2[static_cast<size_t>(idx[i]) * L2 + col]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const double m = s / k;
  double v = 0.0;
  for (int i = lane; i < k; i += 32) {
    const double d = static_cast<double>(act2[static_cast<size_t>(idx[i]) * L2 + col]) - m;
    v += d * d;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if (lane == 0) {
    mean_out[col] = static_cast<float>(m);
    const float sd = static_cast<float>(sqrt(v / (k - 1)));
    std_out[col] = fmaxf(std_floor, sd);
  }
}
cudaError_t launch_refit(const float* act2, int L2, const int64_t* idx, int k, float std_floor, float* mean_out,
                         float* std_out, cudaStream_t s) {
  if (L2 < 1 || L2 > 32) return cudaErrorInvalidValue;
  refit_kernel<<<1, 32 * L2, 0, s>>>(act2, L2, reinterpret_cast<const long long*>(idx), k, std_floor, mean_out,
                                     std_out);
  return cudaGetLastError();
}

}  // namespace rac
