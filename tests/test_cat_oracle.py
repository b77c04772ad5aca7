"""The CaT tail's CPU restatement (oracle/cat_oracle.py) against a sequence produced by the reference's own constraint functions and
CaT class (tests/golden/cat_sequence.npz, made by tests/golden/make_cat_goldens.py).  SURVEY.md 8(f) rank 3: the oracle comes first;
the CUDA side of this variant does not exist yet."""
import os

import numpy as np

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "cat_sequence.npz")
ILLEGAL, FEET, STEP_DT, DEADZONE = [2, 3, 4, 5], [0, 1], 0.02, 0.2


def test_cat_oracle_reproduces_the_reference_sequence():
    from oracle import cat_oracle as O
    g = np.load(GOLD)
    order, max_p = [str(k) for k in g["order"]], dict(zip([str(k) for k in g["order"]], g["max_p"]))
    cat, clearance = O.CaT(0.95, 0.0), O.FootClearance()
    T = g["joint_pos"].shape[0]
    seen_empty_gather = False
    for t in range(T):
        fh, cmd = g["force_hist"][t], g["cmd"][t]
        raw = {
            "contact": O.contact(fh, ILLEGAL), "joint_position_limits": O.joint_position_limits(g["joint_pos"][t], g["soft_limits"]),
            "joint_velocity_limits": O.joint_velocity_limits(g["joint_vel"][t], g["vel_limits"]),
            "joint_torque_limits": O.joint_torque_limits(g["torque"][t], g["effort_limits"]),
            "foot_contact_force": O.foot_contact_force(fh, FEET, 750.0), "no_move": O.no_move(cmd, g["joint_vel"][t], DEADZONE, 6.0),
            "base_orientation": O.base_orientation(g["gravity"][t], 0.1), "base_height": O.base_height(g["root_z"][t], 1.0, 0.05),
            "foot_contact": O.foot_contact(fh, FEET),
            "foot_clearance": clearance(g["foot_z"][t], O.first_contact(g["contact_time"][t], STEP_DT), cmd, 0.1, DEADZONE)}
        seen_empty_gather |= not (np.abs(cmd) < DEADZONE).all(axis=1).any()
        for k in order:
            r = np.asarray(raw[k], np.float32); r = r[:, None] if r.ndim == 1 else r
            np.testing.assert_allclose(r, g[f"raw_{k}"][t], rtol=1e-5, atol=1e-5, err_msg=f"{k} raw, step {t}")  # torch vs numpy float32 norms
            p = cat.add(k, raw[k], max_p[k])
            np.testing.assert_allclose(cat.running_maxes[k], g[f"rmax_{k}"][t], rtol=1e-5, atol=1e-9, err_msg=f"{k} running max, step {t}")
            np.testing.assert_allclose(p, g[f"prob_{k}"][t], rtol=1e-5, atol=1e-7, err_msg=f"{k} probabilities, step {t}")
        np.testing.assert_allclose(cat.get_probs(), g["cstr_prob"][t], rtol=1e-5, atol=1e-7, err_msg=f"combined probability, step {t}")
    assert seen_empty_gather  # the K == 0 branch of no_move was exercised
    # the fixture is not degenerate: every constraint fires somewhere, none everywhere
    for k in order:
        frac = float((g[f"prob_{k}"] > 0).mean())
        assert 0.0 < frac < 1.0, (k, frac)
    # no_move really judges env i on another env's joints (the reference's gather + tile)
    t = 0
    inactive = np.nonzero((np.abs(g["cmd"][t]) < DEADZONE).all(axis=1))[0]
    i = 5
    np.testing.assert_allclose(g["raw_no_move"][t][i], np.abs(g["joint_vel"][t][inactive[i % len(inactive)]]) - 6.0, rtol=1e-6)


def test_constraint_probability_ramp_and_reward_scaling():
    from oracle import cat_oracle as O
    g = np.load(GOLD)
    for c, p in zip(g["ramp_counter"], g["ramp_max_p"]):
        assert abs(O.modify_constraint_p(int(c), 24 * 5000, 0.25) - p) < 1e-12
    assert abs(g["ramp_max_p"][0] - 0.05) < 1e-12 and abs(g["ramp_max_p"][-1] - 0.25) < 1e-12
    r, d = O.constrained_reward(np.array([1.0, 2.0, -1.0], np.float32), np.array([0.0, 0.25, 1.0], np.float32), np.array([False, False, True]))
    np.testing.assert_allclose(r, [1.0, 1.5, 0.0]); np.testing.assert_allclose(d, [0.0, 0.25, 1.0])
