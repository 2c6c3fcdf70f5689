"""CPU: the dataset-glue oracle (oracle/dataset_oracle.py, oracle/data_oracle.py for stored-size frames) against the
golden vectors the UNMODIFIED reference RoboNetDataset produced (oracle/make_golden_dataset.py)."""
import os

import numpy as np
import pytest

from oracle import data_oracle as do
from oracle import dataset_oracle as dso


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, "dataset_glue.npz"))


@pytest.mark.parametrize("tag", ["s96", "s240"])
def test_stored_size_frames_match_reference(gold, tag):
    frames, masks = gold[f"{tag}_frames"], gold[f"{tag}_masks"]
    img, m = do.process_batch(frames, masks)
    np.testing.assert_allclose(img.numpy(), gold[f"{tag}_images_plain"], rtol=0, atol=2e-6)
    assert np.array_equal(m.numpy(), gold[f"{tag}_masks_plain"])
    augs = [do.params_to_aug(r) for r in gold[f"{tag}_params"]]
    img, m = do.process_batch(frames, masks, augs)
    np.testing.assert_allclose(img.numpy(), gold[f"{tag}_images_aug"], rtol=0, atol=3e-5)
    assert (m.numpy() != gold[f"{tag}_masks_aug"]).mean() < 1e-3  # (a mask edge sample that interpolates to +-1e-8)


def test_states_actions_bounds_match_reference(gold):
    for tag, robot, mode, adim in zip(gold["case_tags"], gold["case_robots"], gold["case_modes"], gold["case_action_dims"]):
        tag, robot, mode = str(tag), str(robot), str(mode)
        g = lambda k: gold[f"{tag}_{k}"]
        raw_low, raw_high = dso.load_bounds(robot, g("raw_low"), g("raw_high"))
        assert np.array_equal(raw_low, g("raw_low")) and raw_low.dtype == g("raw_low").dtype
        states = dso.load_states(g("file_states"), 5)
        assert np.array_equal(states, g("loaded_states"))
        actions = dso.load_actions(g("file_actions"), g("file_states"), raw_low[4], raw_high[4], int(adim), True)
        assert np.array_equal(actions, g("loaded_actions"))
        low, high = dso.preprocess_bounds(raw_low, raw_high, mode, g("world2cam"))
        assert np.array_equal(low, g("low")) and np.array_equal(high, g("high"))
        ps = dso.preprocess_states(states, low, high, robot, mode, g("world2cam"))
        assert ps.dtype == np.float32 and np.array_equal(ps, g("states")), tag
        assert np.array_equal(dso.preprocess_actions(ps, actions, mode), g("actions"))
