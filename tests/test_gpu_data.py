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
    with pytest.raises(NotImplementedError):
        process_batch({"images": torch.zeros(1, 1, 64, 85, 3, dtype=torch.uint8)}, "cuda")


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
