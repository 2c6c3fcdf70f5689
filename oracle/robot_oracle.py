"""TEST INFRASTRUCTURE -- CPU oracle of the robot state / mask producer (robot_aware_control_b200/csrc/robot_kernels.cu).

* `start_state_norm`, `predict_states`: numpy / torch restatement of the reference arithmetic, dtype for dtype
  (src/cem/trajectory_sampler.py:93-99; src/dataset/wx250s/wx250s_model.py:57-66,98-117,149-167;
  src/dataset/franka/franka_model.py:48-79; robonet_dataset.py:470-479). PINNED: tests/test_oracle_golden.py compares it
  bit for bit with tests/golden/robot_states.npz, which oracle/make_golden_robot.py recorded from the unmodified reference.
* `render_masks`: float64 restatement of the capsule rasteriser. PARITY UNPINNED against the reference -- the reference
  renders MuJoCo meshes after an Interbotix IK call, neither of which exists here; this function only states what the
  CUDA kernel is supposed to compute.
"""
import numpy as np
import torch

LOCO_FRANKA_DIFF = np.array([-0.365, -0.06103333])
LOCO_WX250S_DIFF = np.array([-0.13, -0.01])
LOW = np.array([0.015, -0.3, 0.1, 0, 0], dtype=np.float32)
HIGH = np.array([0.55, 0.3, 0.4, 1, 1], dtype=np.float32)


def start_state_norm(start_state, kind):
    """trajectory_sampler.py:93-99: torch float32 start state, `[:2] + DIFF` (float64 numpy array -> computed in double,
    stored back as float32), normalize in float32."""
    s = torch.tensor(np.asarray(start_state, dtype=np.float32))
    diff = LOCO_WX250S_DIFF if kind == "wx250s" else LOCO_FRANKA_DIFF
    s[:2] = s[:2] + torch.from_numpy(diff)
    low, high = torch.from_numpy(LOW), torch.from_numpy(HIGH)
    return ((s - low) / (high - low)).numpy()


def predict_states(start_norm, actions, kind, push_height=0.0):
    """start_norm (5,) float32 normalised; actions (N, L, A) float32 -> (L+1, N, 5) float32 normalised states."""
    actions = np.asarray(actions, dtype=np.float32)
    N, L, _ = actions.shape
    out = np.zeros((L + 1, N, 5), dtype=np.float32)
    low, high = LOW, HIGH
    diff = LOCO_WX250S_DIFF if kind == "wx250s" else LOCO_FRANKA_DIFF
    for i in range(N):
        if kind == "wx250s":
            s = np.asarray(start_norm, dtype=np.float32) * (high - low) + low          # denormalize, float32
            s[:2] -= diff                                                             # float32 -= float64: double, stored f32
            states = [s]
            eef = s
            for t in range(L):
                nxt = np.zeros(3)                                                     # float64
                nxt[0:2] = eef[0:2] + actions[i, t, :2]                               # f32+f32 at t = 0, f64+f32 afterwards
                nxt[2] = push_height
                eef = np.concatenate([nxt, [0, 0]])
                states.append(eef)
            raw = torch.from_numpy(np.stack(states).astype(np.float32))
            raw[:, :2] += torch.from_numpy(diff)                                      # f32 tensor += f64: double, stored f32
            out[:, i] = ((raw - torch.from_numpy(low)) / torch.from_numpy(high - low)).numpy()
        else:
            st = np.zeros((L + 1, 5), dtype=np.float32)
            st[0] = start_norm
            w = st * (high - low) + low                                               # float32
            w[:, :2] -= diff
            for t in range(L):
                w[t + 1, :3] = w[t, :3] + actions[i, t, :3]                           # float32
            w[:, :2] += diff
            wt = torch.from_numpy(w)
            out[:, i] = ((wt - torch.from_numpy(low)) / torch.from_numpy(high - low)).numpy()
    return out


def _ray_segment_dist2(c, d, a, b):
    u, w = b - a, c - a
    dd, du, uu, dw, uw = d @ d, d @ u, u @ u, d @ w, u @ w
    den = dd * uu - du * du
    ts = (dd * uw - du * dw) / den if den > 1e-12 else 0.0
    ts = min(max(ts, 0.0), 1.0)
    s = (ts * du - dw) / dd
    if s < 0:
        s = 0.0
        ts = min(max(uw / uu, 0.0), 1.0) if uu > 0 else 0.0
    e = (c + s * d) - (a + ts * u)
    return e @ e


def arm_chain(p, shoulder_z, l_upper, l_fore, l_wrist, pitch):
    """Joint chain (base, shoulder, elbow, wrist, finger tip) of the capsule arm reaching for eef position p (robot frame)."""
    yaw = np.arctan2(p[1], p[0])
    r = np.hypot(p[0], p[1])
    wr, wz = r - l_wrist * np.cos(pitch), p[2] + l_wrist * np.sin(pitch)
    dx, dz = wr, wz - shoulder_z
    dist = np.hypot(dx, dz)
    a, b = l_upper, l_fore
    dc = min(max(dist, abs(a - b) + 1e-4), a + b - 1e-4)
    ca = min(max((a * a + dc * dc - b * b) / (2 * a * dc), -1.0), 1.0)
    alpha = np.arctan2(dz, dx) + np.arccos(ca)
    er, ez = a * np.cos(alpha), shoulder_z + a * np.sin(alpha)
    fx, fz = wr - er, wz - ez
    fl = max(np.hypot(fx, fz), 1e-6)
    wr2, wz2 = er + b * fx / fl, ez + b * fz / fl
    tr, tz = wr2 + l_wrist * np.cos(pitch), wz2 - l_wrist * np.sin(pitch)
    pr, pz = [0.0, 0.0, er, wr2, tr], [0.0, shoulder_z, ez, wz2, tz]
    return np.array([[pr[j] * np.cos(yaw), pr[j] * np.sin(yaw), pz[j]] for j in range(5)])


def render_masks(states, kind, cam_center, cam_minv, shoulder_z, l_upper, l_fore, l_wrist, pitch, radius,
                 extra_radius=0.0, H=48, W=64, margin=None):
    """states (T, N, 5) normalised -> masks (T, N, 1, H, W) float32 {0,1}. With `margin` also returns a boolean array
    of pixels whose ray passes within `margin` metres of a capsule surface (rounding-sensitive pixels)."""
    diff = LOCO_WX250S_DIFF if kind == "wx250s" else LOCO_FRANKA_DIFF
    T, N, _ = states.shape
    masks = np.zeros((T, N, 1, H, W), dtype=np.float32)
    edge = np.zeros((T, N, 1, H, W), dtype=bool)
    c = np.asarray(cam_center, dtype=np.float64)
    minv = np.asarray(cam_minv, dtype=np.float64).reshape(3, 3)
    for t in range(T):
        for n in range(N):
            p = states[t, n, :3].astype(np.float64) * (HIGH[:3] - LOW[:3]).astype(np.float64) + LOW[:3]
            p[:2] -= diff
            chain = arm_chain(p, shoulder_z, l_upper, l_fore, l_wrist, pitch)
            for v in range(H):
                for u in range(W):
                    d = minv @ np.array([u + 0.5, v + 0.5, 1.0])
                    for j in range(4):
                        rad = radius[j] + extra_radius
                        dist = np.sqrt(_ray_segment_dist2(c, d, chain[j], chain[j + 1]))
                        if dist <= rad:
                            masks[t, n, 0, v, u] = 1.0
                        if margin is not None and abs(dist - rad) < margin:
                            edge[t, n, 0, v, u] = True
    return (masks, edge) if margin is not None else masks
