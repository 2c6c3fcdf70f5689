"""Reads the subset of the reference's argparse Namespace (src/config/__init__.py:165-357) that the SVG / CEM path
uses. Any object with these attributes works (the reference `cfg`, a SimpleNamespace, ...); missing attributes take
the reference defaults."""
from types import SimpleNamespace

_DEFAULTS = dict(
    image_width=64, image_height=48, channels=3, g_dim=128, z_dim=10, action_dim=2, robot_dim=6,
    model_use_mask=False, model_use_future_mask=False, model_use_robot_state=True,
    model_use_future_robot_state=False, model_use_heatmap=False, model_use_future_heatmap=False,
    reconstruction_loss="mse", reward_type="weighted", last_frame_skip=False, sample_mean=False, sparse_cost=False,
    robot_cost_weight=0.0, world_cost_weight=1.0, black_robot_input=False, candidates_batch_size=200, topk=5,
    lstm_group_norm=False, debug_cem=False, img_cost_threshold=None, img_cost_world_norm=True, experiment="",
    robot_joint_dim=6, log_dir="logs",
)


def svg_config_from(cfg):
    """Namespace with every attribute the path reads, taken from `cfg` (reference defaults otherwise)."""
    out = SimpleNamespace()
    for k, v in _DEFAULTS.items():
        setattr(out, k, getattr(cfg, k, v))
    out.device = getattr(cfg, "device", None)
    return out


def validate_model_config(c):
    """Raises the exception types the reference raises (ValueError for image size, dynamics.py:470-473)."""
    if c.image_width not in (64,) or c.image_height != 48:
        raise ValueError(f"unsupported image size {c.image_height}x{c.image_width}: the B200 path handles 48x64")
    if c.model_use_heatmap:
        raise NotImplementedError("model_use_heatmap is outside the B200 hot path (SURVEY.md section 8)")
    if c.lstm_group_norm and (c.g_dim > 512 or c.g_dim % 64):
        raise NotImplementedError("lstm_group_norm (NormConvLSTMCell) is implemented for g_dim <= 512")
    if c.g_dim % 64 != 0 or c.g_dim < 64:
        raise ValueError("g_dim must be a multiple of 64 for the tcgen05 tiles")
    if not 1 <= c.z_dim <= 64:
        raise ValueError("z_dim must be in [1, 64]")
    if c.model_use_future_mask and not c.model_use_mask:
        raise ValueError("model_use_future_mask requires model_use_mask")
