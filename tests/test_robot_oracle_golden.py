"""Pins oracle/robot_oracle.py (state half of the robot-model producer) bit for bit against the states the UNMODIFIED
reference returned (tests/golden/robot_states.npz from oracle/make_golden_robot.py: TrajectorySampler's start-state
normalisation + WX250sAnalyticalModel / FrankaAnalyticalModel.predict_batch). CPU only."""
import os

import numpy as np
import pytest

from oracle import robot_oracle as ro
from oracle.make_golden_robot import PUSH_HEIGHT, inputs


@pytest.mark.parametrize("kind", ["wx250s", "franka"])
def test_predict_states_bit_equal_to_reference(golden_dir, kind):
    gold = np.load(os.path.join(golden_dir, "robot_states.npz"))
    for seed in (1, 2):
        actions, start = inputs(seed)
        sn = ro.start_state_norm(start, kind)
        np.testing.assert_array_equal(sn, gold[f"{kind}_start_norm_{seed}"])
        st = ro.predict_states(sn, actions.numpy(), kind, PUSH_HEIGHT)
        assert st.dtype == np.float32 and st.shape == (int(gold["L"]) + 1, int(gold["N"]), 5)
        np.testing.assert_array_equal(st, gold[f"{kind}_states_{seed}"])


def test_capsule_arm_geometry():
    """The mask half has no reference pin; check the stated geometry itself: link lengths are preserved for reachable
    targets, the finger tip lands on the requested end-effector position, unreachable targets stretch the arm."""
    kw = dict(shoulder_z=0.11, l_upper=0.255, l_fore=0.25, l_wrist=0.17, pitch=1.5)
    p = np.array([0.30, 0.05, 0.12])
    ch = ro.arm_chain(p, **kw)
    np.testing.assert_allclose(np.linalg.norm(ch[2] - ch[1]), kw["l_upper"], rtol=1e-9)
    np.testing.assert_allclose(np.linalg.norm(ch[3] - ch[2]), kw["l_fore"], rtol=1e-9)
    np.testing.assert_allclose(np.linalg.norm(ch[4] - ch[3]), kw["l_wrist"], rtol=1e-9)
    np.testing.assert_allclose(ch[4], p, atol=1e-9)
    far = ro.arm_chain(np.array([0.9, 0.0, 0.1]), **kw)
    np.testing.assert_allclose(np.linalg.norm(far[3] - far[1]), kw["l_upper"] + kw["l_fore"], atol=2e-4)
