"""GPU tests of the training-data path (rac_process_batch, robot_aware_control_b200/data.py) against the reference
loader goldens (tests/golden/data_path.npz) and the CPU oracle."""
import os
import random

import numpy as np
import pytest
import torch

from oracle import data_oracle as do

pytestmark = pytest.mark.gpu
ATOL = 2e-5  # [0, 1] pixels, float32: fused multiply-adds and the order of the grey-mean reduction differ from torch's


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, "data_path.npz"))


def test_plain_path_bit_exact_vs_reference(gold):
    from robot_aware_control_b200 import preprocess_clips

    x, m = preprocess_clips(gold["frames"], gold["masks"])
    assert x.is_cuda and x.shape == (2, 6, 3, 48, 64) and m.shape == (2, 6, 1, 48, 64)
    assert np.array_equal(x.cpu().numpy(), gold["images_plain"])
    assert np.array_equal(m.cpu().numpy(), gold["masks_plain"])
    x2, none = preprocess_clips(gold["frames"])  # masks are optional
    assert none is None and torch.equal(x2, x)


def test_augmented_path_vs_reference(gold):
    from robot_aware_control_b200 import preprocess_clips

    augs = [do.params_to_aug(r) for r in gold["params"]]
    x, m = preprocess_clips(gold["frames"], gold["masks"], augs)
    np.testing.assert_allclose(x.cpu().numpy(), gold["images_aug"], rtol=0, atol=ATOL)
    assert np.array_equal(m.cpu().numpy(), gold["masks_aug"])
    # every single transform on its own, and the crop alone, against the oracle
    for order in ([0], [1], [2], [3], []):
        a1 = [(a[0], a[1], a[2], a[3], a[4], order) for a in augs]
        x, m = preprocess_clips(gold["frames"], gold["masks"], a1)
        xo, mo = do.process_batch(gold["frames"], gold["masks"], a1)
        np.testing.assert_allclose(x.cpu().numpy(), xo.numpy(), rtol=0, atol=ATOL, err_msg=str(order))
        assert np.array_equal(m.cpu().numpy(), mo.numpy())


def test_training_batch_vs_oracle_and_dict_api():
    from robot_aware_control_b200 import process_batch, sample_augment

    rs = np.random.RandomState(3)
    B, T = 16, 6
    frames = rs.randint(0, 256, (B, T, 48, 64, 3)).astype(np.uint8)
    frames[0, 0] = 0
    frames[1, 0] = 255
    frames[2, 0] = frames[2, 0, :, :, :1]          # grey frame: max == min for every pixel
    masks = rs.rand(B, T, 48, 64) > 0.8
    random.seed(11)
    torch.manual_seed(11)
    augs = [sample_augment() for _ in range(B)]
    data = {"images": torch.from_numpy(frames), "masks": torch.from_numpy(masks),
            "states": torch.rand(B, T, 5), "actions": torch.rand(B, T - 1, 5), "robot": ["sawyer"] * B}
    out = process_batch(data, "cuda", augment=augs)
    xo, mo = do.process_batch(frames, masks.astype(np.float32), augs)
    assert out["images"].shape == (T, B, 3, 48, 64) and out["masks"].shape == (T, B, 1, 48, 64)
    np.testing.assert_allclose(out["images"].cpu().numpy(), xo.numpy(), rtol=0, atol=ATOL)
    assert np.array_equal(out["masks"].cpu().numpy(), mo.numpy())
    assert torch.equal(out["states"].cpu(), data["states"].transpose(1, 0)) and out["actions"].shape == (T - 1, B, 5)
    # float32 and uint8 masks take different loads in the kernel: same result
    for mk in (masks.astype(np.float32), masks.astype(np.uint8)):
        o2 = process_batch({"images": frames, "masks": mk}, "cuda", augment=augs)
        assert torch.equal(o2["masks"], out["masks"]) and torch.equal(o2["images"], out["images"])
    # deterministic, and an empty batch is fine
    o3 = process_batch(data, "cuda", augment=augs)
    assert torch.equal(o3["images"], out["images"])
    e = process_batch({"images": torch.zeros(0, T, 48, 64, 3, dtype=torch.uint8)}, "cuda")
    assert e["images"].shape == (T, 0, 3, 48, 64)
    # another stored size goes through the dataset's Resize first (test_stored_size_frames_vs_reference)
    o4 = process_batch({"images": torch.zeros(1, 1, 64, 85, 3, dtype=torch.uint8)}, "cuda")
    assert o4["images"].shape == (1, 1, 3, 48, 64) and float(o4["images"].abs().max()) == 0.0


def test_feeds_the_training_step():
    """The tensors come out in the layout SVGTrainer.train_step takes (time-first, float32, device)."""
    from oracle import svg_oracle as so
    from robot_aware_control_b200 import SVGConvModel, SVGTrainer, process_batch, sample_augment

    cfg = so.make_cfg(g_dim=128, z_dim=10, model_use_mask=True, model_use_robot_state=True,
                      reconstruction_loss="dontcare_l1", lr=1e-3, beta=1e-4, beta1=0.9, n_future=2, n_past=1)
    model = SVGConvModel(cfg)
    model.load_state_dict(so.make_state_dict(cfg, 3))
    model.train()
    trainer = SVGTrainer(cfg, model)
    rs = np.random.RandomState(0)
    B, T = 4, 3
    data = {"images": rs.randint(0, 256, (B, T, 48, 64, 3)).astype(np.uint8), "masks": rs.rand(B, T, 48, 64) > 0.8,
            "states": torch.rand(B, T, 5), "actions": torch.rand(B, T - 1, 5) * 0.1 - 0.05}
    random.seed(0)
    batch = process_batch(data, "cuda", augment=[sample_augment() for _ in range(B)])
    out = trainer.train_step(batch)
    assert np.isfinite(out["recon_loss"]) and np.isfinite(out["kld"])


@pytest.fixture(scope="module")
def glue(golden_dir):
    return np.load(os.path.join(golden_dir, "dataset_glue.npz"))


@pytest.mark.parametrize("tag", ["s96", "s240"])
def test_stored_size_frames_vs_reference(glue, tag):
    """Clips stored at 96 x 128 / 240 x 320 (RoboNet's size): tf.Resize((48, 64)) of the reference dataset (bilinear, no
    antialiasing: the pinned torchvision) + the rest of _preprocess_images_masks, in the same launch."""
    from robot_aware_control_b200 import preprocess_clips

    frames, masks = glue[f"{tag}_frames"], glue[f"{tag}_masks"]
    x, m = preprocess_clips(frames, masks)
    assert x.shape == (2, frames.shape[0], 3, 48, 64)
    np.testing.assert_allclose(x.cpu().numpy(), glue[f"{tag}_images_plain"], rtol=0, atol=2e-6)
    assert np.array_equal(m.cpu().numpy(), glue[f"{tag}_masks_plain"])
    augs = [do.params_to_aug(r) for r in glue[f"{tag}_params"]]
    x, m = preprocess_clips(frames, masks, augs)
    np.testing.assert_allclose(x.cpu().numpy(), glue[f"{tag}_images_aug"], rtol=0, atol=3e-5)
    assert (m.cpu().numpy() != glue[f"{tag}_masks_aug"]).mean() < 1e-3
    xo, mo = do.process_batch(frames, masks, augs)
    np.testing.assert_allclose(x.cpu().numpy(), xo.numpy(), rtol=0, atol=3e-5)
    # uint8 masks: same result
    x2, m2 = preprocess_clips(frames, masks.astype(np.uint8), augs)
    assert torch.equal(x2, x) and torch.equal(m2, m)


def test_states_actions_vs_reference(glue):
    """RoboNetDataset._load_states / _load_actions / _preprocess_bounds / _preprocess_states / _preprocess_actions +
    process_batch for all three robot kinds, world and camera frame, against the reference's own outputs."""
    from robot_aware_control_b200.data import clip_calibration, preprocess_bounds, preprocess_states_actions

    for tag, robot, mode, adim in zip(glue["case_tags"], glue["case_robots"], glue["case_modes"], glue["case_action_dims"]):
        tag, robot, mode = str(tag), str(robot), str(mode)
        g = lambda k: glue[f"{tag}_{k}"]
        fs, fa = g("file_states"), g("file_actions")
        low, high = preprocess_bounds(g("raw_low"), g("raw_high"), mode, g("world2cam"))
        assert np.array_equal(low, g("low")) and np.array_equal(high, g("high"))
        cal = clip_calibration(robot, mode, g("raw_low"), g("raw_high"), g("world2cam"), stored_state_dim=fs.shape[1])
        # a batch of two clips (the second one time-reversed) exercises the batch-first -> time-first layout
        states = np.stack([fs, fs[::-1]])
        actions = np.stack([fa, fa[::-1]])
        s, a = preprocess_states_actions(states, actions, [cal, cal], robot_dim=5, action_dim=int(adim), preprocess_action=mode)
        assert s.shape == (fs.shape[0], 2, 5) and a.shape == (fa.shape[0], 2, int(adim)) and s.is_cuda
        np.testing.assert_allclose(s[:, 0].cpu().numpy(), g("states"), rtol=0, atol=1e-6, err_msg=tag)
        np.testing.assert_allclose(s[:, 1].cpu().numpy()[::-1], g("states"), rtol=0, atol=1e-6, err_msg=tag)
        np.testing.assert_allclose(a[:, 0].cpu().numpy(), g("actions"), rtol=0, atol=1e-7, err_msg=tag)
    with pytest.raises(ValueError):  # the reference raises for an action width it cannot reconcile (:193-194)
        preprocess_states_actions(states, actions, [cal, cal], action_dim=actions.shape[-1] + 2)
    with pytest.raises(NotImplementedError):  # state_infer / camera_state_infer (:341-352)
        preprocess_states_actions(states, actions, [cal, cal], preprocess_action="state_infer")
