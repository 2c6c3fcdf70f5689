"""TEST INFRASTRUCTURE -- not a product path.

Imports the UNMODIFIED reference (penn-pal-lab/robot_aware_control at /root/reference) in this container so that
`oracle/make_golden.py` can run the reference's own PyTorch code and dump golden vectors. The reference imports 14
third-party modules that are absent here and take no part in the hot-path arithmetic (matplotlib, skimage, imageio,
ipdb, h5py, gym, mujoco_py, colorlog, rospy, ...); they are replaced by permissive stub modules.

Nothing under robot_aware_control_b200/ imports this file; /root/reference does not exist on the GPU box.
"""
import importlib
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("RAC_REFERENCE_ROOT", "/root/reference")


class _Stub(types.ModuleType):
    __path__ = []  # looks like a package, so `import a.b.c` works

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        if name[:1].isupper():
            cls = type(name, (), {"__init__": lambda self, *a, **k: None})
            setattr(self, name, cls)
            return cls
        full = f"{self.__name__}.{name}"
        mod = sys.modules.get(full)
        if mod is None:
            mod = _Stub(full)
            sys.modules[full] = mod
        setattr(self, name, mod)
        return mod

    def __call__(self, *a, **k):
        return None


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "src", "prediction"))


def import_reference():
    """Returns a dict of the reference modules on the hot path."""
    if not reference_available():
        raise RuntimeError(f"reference checkout not found at {REFERENCE_ROOT}")
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    names = [
        "src.config",
        "src.prediction.models.dynamics",
        "src.prediction.losses",
        "src.utils.state",
        "src.utils.image",
        "src.cem.trajectory_sampler",
        "src.cem.cem",
        "src.prediction.trainer",
    ]
    mods = {}
    for n in names:
        for _ in range(64):  # discover missing third-party modules one by one
            try:
                mods[n] = importlib.import_module(n)
                break
            except ModuleNotFoundError as e:
                missing = e.name
                if missing is None or missing.startswith("src"):
                    raise
                parts = missing.split(".")
                for i in range(1, len(parts) + 1):
                    key = ".".join(parts[:i])
                    if key not in sys.modules:
                        sys.modules[key] = _Stub(key)
        else:
            raise RuntimeError(f"could not import {n}")
    return mods


def make_cfg(g_dim=512, z_dim=64, action_dim=5, robot_dim=5, robot_aware=False, future_mask=False,
             future_robot_state=False, extra=()):
    """argparse Namespace exactly as the reference builds it (src/config/__init__.py), CPU device."""
    import torch

    mods = import_reference()
    argv = [
        "x", "--g_dim", str(g_dim), "--z_dim", str(z_dim), "--action_dim", str(action_dim),
        "--robot_dim", str(robot_dim), "--n_past", "1", "--n_future", "5", "--batch_size", "16",
        "--model", "svg", "--last_frame_skip", "True",
    ]
    if robot_aware:
        argv += ["--model_use_robot_state", "True", "--model_use_mask", "True",
                 "--reconstruction_loss", "dontcare_l1", "--reward_type", "dontcare",
                 "--robot_joint_dim", "6", "--experiment", "control_wx250s"]
        if future_mask:
            argv += ["--model_use_future_mask", "True"]
        if future_robot_state:
            argv += ["--model_use_future_robot_state", "True"]
    else:
        argv += ["--model_use_robot_state", "False", "--model_use_mask", "False", "--reconstruction_loss", "l1"]
    argv += list(extra)
    old = sys.argv
    sys.argv = argv
    try:
        cfg, _ = mods["src.config"].argparser()
    finally:
        sys.argv = old
    cfg.device = torch.device("cpu")
    return cfg
