// Device-side producer of the robot states and masks that feed a robot-aware plan every CEM iteration
// (SURVEY.md 8(f) rank 1). Replaces the host round trip `robot_model.predict_batch(data)` of
// TrajectorySampler.generate_model_rollouts (reference src/cem/trajectory_sampler.py:86-109).
//
// * robot_states_kernel -- the STATE half, pinned to the reference: the planar end-effector integration of
//   WX250sAnalyticalModel.predict_batch / predict_trajectory / predict_next_state_qpos
//   (src/dataset/wx250s/wx250s_model.py:57-66,98-117,149-167) and of FrankaAnalyticalModel.predict_batch
//   (src/dataset/franka/franka_model.py:48-80), with the reference's exact mix of float32 / float64 arithmetic
//   (numpy promotes `float32 -= float64 array` and `float64 + float32` to double; torch's normalize stays fp32), so the
//   result is bit-equal to the reference's (tests/golden/robot_states.npz).
// * robot_masks_kernel -- the MASK half: a capsule model of the arm (column, upper arm, forearm, wrist + gripper)
//   posed by a closed-form planar IK and rasterised through a pinhole camera, one thread per pixel. The reference
//   obtains masks from a MuJoCo segmentation render of the arm meshes after an Interbotix IK call (wx250s_model.py:
//   38-118, base_mask_env.py:41-83); neither exists in this environment, so this half is NOT reference-pinned -- its
//   oracle is oracle/robot_oracle.py (float64 numpy restatement of the same geometry).
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/racb200.h"

namespace rac {

struct RobotParams {
  int kind;
  float low[5], high[5];
  double diff[2];
  double push_height;
  // rasteriser
  float cam_center[3];
  float minv[9];  // pixel (u, v, 1) -> ray direction in the robot base frame
  float shoulder_z, l_upper, l_fore, l_wrist, pitch;
  float radius[4];
};

// No FMA contraction anywhere in here: every operation is a separately rounded IEEE op, as numpy / torch CPU do.
__global__ void __launch_bounds__(128)
robot_states_kernel(RobotParams P, const float* __restrict__ start_state /* (5) normalised */,
                    const float* __restrict__ actions /* (n, L, adim) */, int n, int L, int adim,
                    float* __restrict__ states /* (L+1, n, 5) */, long long t_stride) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float s0[5];
#pragma unroll
  for (int d = 0; d < 5; ++d) {
    // denormalize (robonet_dataset.py:470-473): states * (high - low) + low, float32
    const float span = __fsub_rn(P.high[d], P.low[d]);
    s0[d] = __fadd_rn(__fmul_rn(start_state[d], span), P.low[d]);
  }
  // start_state[:2] -= DIFF (float32 array -= float64 array: computed in double, stored as float32)
  s0[0] = static_cast<float>(__dsub_rn(static_cast<double>(s0[0]), P.diff[0]));
  s0[1] = static_cast<float>(__dsub_rn(static_cast<double>(s0[1]), P.diff[1]));
  auto store = [&](int t, const float raw[5]) {
    float* out = states + static_cast<size_t>(t) * t_stride + static_cast<size_t>(i) * 5;
#pragma unroll
    for (int d = 0; d < 5; ++d) {
      float v = raw[d];
      // raw[:, :2] += DIFF (float32 tensor += float64: computed in double, stored as float32)
      if (d < 2) v = static_cast<float>(__dadd_rn(static_cast<double>(v), P.diff[d]));
      // normalize (robonet_dataset.py:476-479): (states - low) / (high - low), float32
      out[d] = __fdiv_rn(__fsub_rn(v, P.low[d]), __fsub_rn(P.high[d], P.low[d]));
    }
  };
  const float* a = actions + static_cast<size_t>(i) * L * adim;
  if (P.kind == RAC_ROBOT_WX250S) {
    store(0, s0);
    // predict_trajectory (wx250s_model.py:98-117): the first sum is float32 + float32 (start state and action are both
    // float32) stored into a float64 array; from then on the eef is float64 and float64 + float32 promotes to double
    double ex = 0.0, ey = 0.0;
    for (int t = 0; t < L; ++t) {
      const float ax = a[t * adim + 0], ay = a[t * adim + 1];
      if (t == 0) {
        ex = static_cast<double>(__fadd_rn(s0[0], ax));
        ey = static_cast<double>(__fadd_rn(s0[1], ay));
      } else {
        ex = __dadd_rn(ex, static_cast<double>(ax));
        ey = __dadd_rn(ey, static_cast<double>(ay));
      }
      const float raw[5] = {static_cast<float>(ex), static_cast<float>(ey), static_cast<float>(P.push_height), 0.f, 0.f};
      store(t + 1, raw);
    }
  } else {
    // FrankaAnalyticalModel.predict_batch (franka_model.py:48-79): float32 waypoints; rows t >= 1 start as the
    // denormalised ZERO state (= low, with the frame shift applied to x, y), then xyz[t+1] = xyz[t] + act[:3] in float32
    float w[5];
#pragma unroll
    for (int d = 0; d < 5; ++d) w[d] = s0[d];
    store(0, w);
    float rest[2];
#pragma unroll
    for (int d = 0; d < 2; ++d)
      rest[d] = __fadd_rn(__fmul_rn(0.f, __fsub_rn(P.high[3 + d], P.low[3 + d])), P.low[3 + d]);
    for (int t = 0; t < L; ++t) {
#pragma unroll
      for (int d = 0; d < 3; ++d) w[d] = __fadd_rn(w[d], d < adim ? a[t * adim + d] : 0.f);
      w[3] = rest[0];
      w[4] = rest[1];
      store(t + 1, w);
    }
  }
}

__device__ __forceinline__ float dot3(const float* a, const float* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }

// squared distance between the ray c + s d (s >= 0) and the segment [a, b]
__device__ float ray_segment_dist2(const float* c, const float* d, const float* a, const float* b) {
  float u[3], w[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) { u[k] = b[k] - a[k]; w[k] = c[k] - a[k]; }
  const float dd = dot3(d, d), du = dot3(d, u), uu = dot3(u, u), dw = dot3(d, w), uw = dot3(u, w);
  const float den = dd * uu - du * du;
  float tseg = den > 1e-12f ? (dd * uw - du * dw) / den : 0.f;  // parameter on the segment of the closest line points
  tseg = fminf(fmaxf(tseg, 0.f), 1.f);
  float s = (tseg * du - dw) / dd;                               // closest ray parameter for that segment point
  if (s < 0.f) {  // behind the camera: closest ray point is the camera centre, re-project it on the segment
    s = 0.f;
    tseg = uu > 0.f ? fminf(fmaxf(uw / uu, 0.f), 1.f) : 0.f;
  }
  float r2 = 0.f;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const float e = (c[k] + s * d[k]) - (a[k] + tseg * u[k]);
    r2 += e * e;
  }
  return r2;
}

__global__ void __launch_bounds__(256)
robot_masks_kernel(RobotParams P, const float* __restrict__ states /* (L+1, n, 5) normalised */, long long st_stride,
                   int n, int T1, int H, int W, float extra_radius, float* __restrict__ masks /* (L+1, n, 1, H, W) */,
                   long long m_stride) {
  __shared__ float seg[5][3];  // joint chain: base, shoulder, elbow, wrist, finger tip
  const int i = blockIdx.x, t = blockIdx.y;
  if (threadIdx.x == 0) {
    const float* s = states + static_cast<size_t>(t) * st_stride + static_cast<size_t>(i) * 5;
    // normalised loco frame -> metres in the robot base frame (the raw eef of predict_trajectory)
    float p[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) p[d] = s[d] * (P.high[d] - P.low[d]) + P.low[d];
    p[0] -= static_cast<float>(P.diff[0]);
    p[1] -= static_cast<float>(P.diff[1]);
    const float yaw = atan2f(p[1], p[0]);
    const float r = sqrtf(p[0] * p[0] + p[1] * p[1]);
    // wrist point in the vertical plane through the waist axis; pitch > 0 tilts the gripper down
    const float wr = r - P.l_wrist * cosf(P.pitch), wz = p[2] + P.l_wrist * sinf(P.pitch);
    const float sz = P.shoulder_z;
    float dx = wr, dz = wz - sz;
    float dist = sqrtf(dx * dx + dz * dz);
    const float a = P.l_upper, b = P.l_fore;
    const float dmin = fabsf(a - b) + 1e-4f, dmax = a + b - 1e-4f;
    const float dc = fminf(fmaxf(dist, dmin), dmax);  // unreachable targets: the arm points at them, fully folded / stretched
    const float base = atan2f(dz, dx);
    const float ca = fminf(fmaxf((a * a + dc * dc - b * b) / (2.f * a * dc), -1.f), 1.f);
    const float alpha = base + acosf(ca);  // elbow-up
    const float er = a * cosf(alpha), ez = sz + a * sinf(alpha);
    // forearm: from the elbow towards the wrist, length b (coincides with the wrist when the target is reachable)
    float fx = wr - er, fz = wz - ez;
    const float fl = fmaxf(sqrtf(fx * fx + fz * fz), 1e-6f);
    const float wr2 = er + b * fx / fl, wz2 = ez + b * fz / fl;
    const float tr = wr2 + P.l_wrist * cosf(P.pitch), tz = wz2 - P.l_wrist * sinf(P.pitch);
    const float cy = cosf(yaw), sy = sinf(yaw);
    const float pr[5] = {0.f, 0.f, er, wr2, tr}, pz[5] = {0.f, sz, ez, wz2, tz};
    for (int j = 0; j < 5; ++j) {
      seg[j][0] = pr[j] * cy;
      seg[j][1] = pr[j] * sy;
      seg[j][2] = pz[j];
    }
  }
  __syncthreads();
  float* out = masks + static_cast<size_t>(t) * m_stride + static_cast<size_t>(i) * H * W;
  for (int pix = threadIdx.x; pix < H * W; pix += blockDim.x) {
    const float u = (pix % W) + 0.5f, v = (pix / W) + 0.5f;
    float d[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) d[k] = P.minv[3 * k] * u + P.minv[3 * k + 1] * v + P.minv[3 * k + 2];
    bool hit = false;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float rad = P.radius[j] + extra_radius;
      hit = hit || ray_segment_dist2(P.cam_center, d, seg[j], seg[j + 1]) <= rad * rad;
    }
    out[pix] = hit ? 1.f : 0.f;
  }
}

static RobotParams to_params(const rac_robot_model* m) {
  RobotParams P{};
  P.kind = m->kind;
  for (int d = 0; d < 5; ++d) { P.low[d] = m->low[d]; P.high[d] = m->high[d]; }
  P.diff[0] = m->frame_diff[0]; P.diff[1] = m->frame_diff[1];
  P.push_height = m->push_height;
  for (int k = 0; k < 3; ++k) P.cam_center[k] = m->cam_center[k];
  for (int k = 0; k < 9; ++k) P.minv[k] = m->cam_minv[k];
  P.shoulder_z = m->shoulder_z; P.l_upper = m->l_upper; P.l_fore = m->l_fore; P.l_wrist = m->l_wrist; P.pitch = m->pitch;
  for (int k = 0; k < 4; ++k) P.radius[k] = m->radius[k];
  return P;
}

cudaError_t launch_robot_states(const rac_robot_model* m, const float* start_state, const float* actions, int n, int L,
                                int adim, float* states, long long t_stride, cudaStream_t s) {
  if (n <= 0) return cudaSuccess;
  robot_states_kernel<<<(n + 127) / 128, 128, 0, s>>>(to_params(m), start_state, actions, n, L, adim, states, t_stride);
  return cudaGetLastError();
}

cudaError_t launch_robot_masks(const rac_robot_model* m, const float* states, long long st_stride, int n, int T1, int H,
                               int W, float extra_radius, float* masks, long long m_stride, cudaStream_t s) {
  if (n <= 0) return cudaSuccess;
  robot_masks_kernel<<<dim3(n, T1), 256, 0, s>>>(to_params(m), states, st_stride, n, T1, H, W, extra_radius, masks,
                                                 m_stride);
  return cudaGetLastError();
}

}  // namespace rac
