"""The data-path oracle (oracle/data_oracle.py) and the host-side augmentation sampling against the UNMODIFIED
reference loader code (tests/golden/data_path.npz from oracle/make_golden_data.py). CPU only."""
import os
import random

import numpy as np
import pytest
import torch

from oracle import data_oracle as do


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, "data_path.npz"))


def test_plain_path_is_bit_exact(gold):
    x, m = do.process_batch(gold["frames"], gold["masks"])
    assert np.array_equal(x.numpy(), gold["images_plain"]) and np.array_equal(m.numpy(), gold["masks_plain"])


def test_augmented_path_matches_reference(gold):
    augs = [do.params_to_aug(r) for r in gold["params"]]
    assert any(a[2:4] == (48, 64) for a in augs) and any(a[2:4] != (48, 64) for a in augs)  # both branches covered
    x, m = do.process_batch(gold["frames"], gold["masks"], augs)
    np.testing.assert_allclose(x.numpy(), gold["images_aug"], rtol=0, atol=5e-6)  # float re-association only
    assert np.array_equal(m.numpy(), gold["masks_aug"])
    assert np.abs(gold["images_aug"] - gold["images_plain"]).max() > 0.05         # the augmentation did something


def test_sample_augment_consumes_the_generators_like_the_reference(gold):
    from robot_aware_control_b200.data import pack_augment, sample_augment, AUGMENT_DTYPE

    got = []
    for seed, row in zip(gold["seeds"], gold["params"]):
        random.seed(int(seed))
        torch.manual_seed(int(seed))
        i, j, th, tw, factors, order = sample_augment(48, 64)
        got.append((i, j, th, tw, factors, order))
        assert [i, j, th, tw] == [int(v) for v in row[:4]]
        assert factors == [float(v) for v in row[4:8]]           # same draws, bit for bit
        assert order == [int(v) for v in row[8:12]]
    packed = pack_augment(got)
    assert packed.dtype == AUGMENT_DTYPE and packed.itemsize == 64 and packed.shape == (len(got),)
    assert packed[1]["crop"].tolist() == [int(v) for v in gold["params"][1][:4]]
    with pytest.raises(ValueError):
        pack_augment([(40, 0, 20, 64, [1, 1, 1, 0], [0, 1, 2, 3])])   # window leaves the frame
    with pytest.raises(ValueError):
        pack_augment([(0, 0, 48, 64, [1, 1, 1, 0.7], [0, 1, 2, 3])])  # torchvision's hue range


def test_process_batch_transposes_like_the_reference():
    from robot_aware_control_b200.data import process_batch

    g = torch.Generator().manual_seed(0)
    data = {"images": torch.rand(3, 4, 3, 48, 64, generator=g), "masks": torch.rand(3, 4, 1, 48, 64, generator=g),
            "states": torch.rand(3, 4, 5, generator=g), "actions": torch.rand(3, 3, 5, generator=g), "robot": ["a"] * 3}
    out = process_batch(dict(data), "cpu")
    for k in ("images", "masks", "states", "actions"):
        assert torch.equal(out[k], data[k].transpose(1, 0))
    assert out["robot"] == data["robot"]
    with pytest.raises(ValueError):
        process_batch(dict(data), "cpu", augment=[None] * 3)      # float images cannot be augmented on the device
    if not torch.cuda.is_available():
        with pytest.raises((RuntimeError, AssertionError)):       # raw frames need the CUDA path: no CPU fallback
            process_batch({"images": torch.zeros(1, 2, 48, 64, 3, dtype=torch.uint8)}, "cuda")
