"""CPU tests: pin the oracle against every known answer the reference tree holds for the path."""
import json
import os
import re

import numpy as np
import pytest

from oracle import oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEY = np.array([0, 0, 1.05, 1, 0, 0, 0] + [0, -0.16, 0, 0.36, -0.2, 0] * 2, float)


def test_philox_known_answers():
    # Random123 kat_vectors for philox4x32-10
    assert [hex(x) for x in O.philox([0, 0, 0, 0], [0, 0])] == ["0x6627e8d5", "0xe169c58d", "0xbc57ac4c", "0x9b00dbd8"]
    assert [hex(x) for x in O.philox([0xFFFFFFFF] * 4, [0xFFFFFFFF] * 2)] == ["0x408f276d", "0x41c83b0e", "0xa20bc7c6", "0x6d5451fd"]
    assert [hex(x) for x in O.philox([0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344], [0xA4093822, 0x299F31D0])] == [
        "0xd16cfe09", "0x94fdcceb", "0x5001e420", "0x24126ea1"]
    u = O.rng4(42, 7, 3, 1, 0)
    assert np.all((u >= 0) & (u < 1))


def test_total_mass_and_weight(cfg):
    # D/utils/mj_logger.py:64 total_mass_force = sum(m) * |g| ; SURVEY KAT (1)
    model = json.load(open(os.path.join(ROOT, "h1v2_isaac_b200", "model", "h12_12dof.json")))
    assert abs(model["total_mass"] - 67.3675873) < 1e-6
    b = O.bias(cfg, KEY, np.zeros(18))
    assert abs(b[2] - 67.3675873 * 9.81) < 1e-3  # gravity load on the vertical root dof
    assert abs(b[0]) < 1e-9 and abs(b[1]) < 1e-9


def test_keyframe_fk(cfg):
    # SURVEY KAT (3): left ankle_roll_link origin at the keyframe; R = I because -0.16 + 0.36 - 0.2 = 0
    R, x = O.fk(KEY)
    np.testing.assert_allclose(x[6], [-0.015740, 0.163, 0.079882], atol=2e-6)
    np.testing.assert_allclose(x[12], [-0.015740, -0.163, 0.079882], atol=2e-6)
    np.testing.assert_allclose(R[6], np.eye(3), atol=1e-12)
    # pelvis height with flat soles on the ground = 1.05 - (0.079882 - 0.045)
    assert abs((1.05 - (x[6][2] - 0.045)) - 1.015118) < 2e-6


def test_mass_matrix_against_independent_numpy(cfg):
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import compile_model as cm
    model = json.load(open(os.path.join(ROOT, "h1v2_isaac_b200", "model", "h12_12dof.json")))
    rng = np.random.default_rng(0)
    for _ in range(4):
        q = np.concatenate([rng.normal(size=3), rng.normal(size=4), rng.uniform(-0.5, 0.5, 12)])
        q[3:7] /= np.linalg.norm(q[3:7])
        M = O.mass_matrix(cfg, q)
        M2 = cm.mass_matrix(model, q)[0]
        np.testing.assert_allclose(M, M2, atol=1e-8)
        assert np.linalg.eigvalsh(M).min() > 0


def test_bias_matches_finite_difference_of_energy(cfg):
    """Free flight without dissipation: d/dt(KE+PE) = 0 to first order in dt (Coriolis/gravity consistent with M)."""
    c = cfg.copy()
    for d in range(18):
        c.dof_damping[d] = 0
        c.dof_frictionloss[d] = 0
    for j in range(12):
        c.joint_range[j][0], c.joint_range[j][1] = -100, 100
    drift = []
    for dt in (1e-3, 5e-4):
        c.sim_dt = dt
        q = np.array([0, 0, 10.0, 1, 0, 0, 0] + [0, -0.16, 0, 0.36, -0.2, 0] * 2, float)
        v = np.concatenate([[0.3, -0.2, 1.0], [1.0, -2.0, 0.5], [1, -2, 1.5, -3, 2, -1, -1, 2, -1.5, 3, -2, 1.0]])
        e0 = O.total_energy(c, q, v)
        for _ in range(int(0.1 / dt)):
            q, v, _, _, _ = O.physics_step(c, q, v, np.zeros(12))
        drift.append(O.total_energy(c, q, v) - e0)
    assert abs(drift[0]) < 1.0  # ~30 J of kinetic energy in play
    assert 1.6 < drift[0] / drift[1] < 2.4  # first-order integrator: halving dt halves the drift


def test_momentum_conservation_in_flight(cfg):
    """Internal torques cannot change the horizontal momentum; the semi-implicit integrator leaves an O(dt) drift."""
    c = cfg.copy()
    for d in range(18):
        c.dof_damping[d] = 0
        c.dof_frictionloss[d] = 0
    tau = np.array([5, -10, 3, 20, -2, 1, -5, 10, -3, -20, 2, -1.0])
    drift = []
    for dt in (1e-3, 5e-4):
        c.sim_dt = dt
        q = np.array([0, 0, 10.0, 1, 0, 0, 0] + [0.1, -0.3, 0.05, 0.6, -0.2, 0.1] * 2, float)
        v = np.concatenate([[0.3, -0.2, 0.0], [1.0, -2.0, 0.5], [1, -2, 1.5, -3, 2, -1, -1, 2, -1.5, 3, -2, 1.0]])
        p0 = (O.mass_matrix(c, q) @ v)[:3]
        for _ in range(int(0.05 / dt)):
            q, v, _, _, _ = O.physics_step(c, q, v, tau)
        drift.append((O.mass_matrix(c, q) @ v)[:3] - p0)
    assert np.abs(drift[1][:2]).max() < 0.1  # of ~30 kg m/s
    r = drift[0][:2] / drift[1][:2]
    assert np.all((r > 1.7) & (r < 2.3))
    assert abs(drift[1][2] + 67.3675873 * 9.81 * 0.05) < 0.5  # gravity impulse on the vertical momentum


def test_static_stance_supports_the_weight(cfg):
    """SURVEY section 7 gate: double support, PD holding the default pose: sum of normal forces ~ 660.88 N."""
    q0 = np.array([0, -0.16, 0, 0.36, -0.2, 0] * 2, float)
    q = np.array([0, 0, 1.0151, 1, 0, 0, 0] + list(q0), float)
    v = np.zeros(18)
    kp, kd, ef = np.array(cfg.kp[:]), np.array(cfg.kd[:]), np.array(cfg.effort_limit[:])
    fz = []
    for i in range(60):
        tau = np.clip(kp * (q0 - q[7:]) - kd * v[6:], -ef, ef)
        q, v, sf, it, res = O.physics_step(cfg, q, v, tau)
        assert res < 1e-9 and it < 30
        fz.append(sf[:, 2].sum())
    assert abs(np.mean(fz[30:]) - 660.876) < 40.0
    assert sf[2:, :].max() == 0.0  # only the feet touch


def test_soft_joint_limits(cfg):
    # SURVEY KAT (8): soft limits = mid +- 0.45*range (A/robots/h12.py:56)
    lo, hi = cfg.joint_range[4][0], cfg.joint_range[4][1]
    mid, half = 0.5 * (lo + hi), 0.5 * (hi - lo) * cfg.soft_limit_factor
    assert abs((mid - half) - (-0.826287)) < 1e-5 and abs((mid + half) - 0.452551) < 1e-5
    assert abs(0.5 * (cfg.joint_range[5][1] - cfg.joint_range[5][0]) * 0.9 - 0.235619) < 1e-5


def test_projected_gravity_closed_form(cfg):
    """SURVEY KAT (4): D/controllers/rl.py:86-95 closed form == the oracle's projected-gravity observation."""
    c = cfg.copy()
    c.enable_corruption = 0
    n = 64
    orc = O.Oracle(c, n, seed=1)
    rng = np.random.default_rng(2)
    quat = rng.normal(size=(n, 4))
    quat /= np.linalg.norm(quat, axis=1, keepdims=True)
    orc.set_state({"root_quat": quat.astype(np.float32), "fresh": np.full((n, 1), 3, np.int32)})
    obs = orc.observe().reshape(n, -1)
    H = c.history_length
    g_obs = obs[:, 3 * H + 3 * (H - 1): 3 * H + 3 * H]  # newest sample of the projected_gravity block
    qw, qx, qy, qz = quat.astype(np.float32).astype(np.float64).T
    g = np.stack([2 * (-qz * qx + qw * qy), -2 * (qz * qy + qw * qx), 1 - 2 * (qw * qw + qz * qz)], axis=1)
    np.testing.assert_allclose(g_obs, g, atol=2e-6)
    ident = np.zeros((n, 4), np.float32); ident[:, 0] = 1
    orc.set_state({"root_quat": ident, "fresh": np.full((n, 1), 3, np.int32)})
    o2 = orc.observe()
    np.testing.assert_allclose(o2[:, 3 * H: 3 * H + 3], np.tile([0, 0, -1], (n, 1)), atol=1e-7)


def test_history_fill_and_layout(cfg):
    """SURVEY KAT (5): first observation after a reset is repeated H times; blocks are term-major, oldest->newest
    (T/utils/history/circular_buffer.py:131-135,79-87 ; D/controllers/rl.py:68-81)."""
    c = cfg.copy()
    c.enable_corruption = 0
    n, H = 8, c.history_length
    orc = O.Oracle(c, n, seed=3)
    o0 = orc.observe()
    off = [0, 3, 6, 9, 21, 33, 45]
    for t in range(6):
        d = off[t + 1] - off[t]
        blk = o0[:, off[t] * H: off[t + 1] * H].reshape(n, H, d)
        assert np.all(blk == blk[:, :1, :])  # back-filled with the first sample
    a = np.random.default_rng(0).normal(size=(n, 12)).astype(np.float32) * 0.1
    o1, _, term, trunc = orc.step(a)
    assert not term.any() and not trunc.any()
    for t in range(6):
        d = off[t + 1] - off[t]
        b0 = o0[:, off[t] * H: off[t + 1] * H].reshape(n, H, d)
        b1 = o1[:, off[t] * H: off[t + 1] * H].reshape(n, H, d)
        assert np.array_equal(b1[:, :-1], b0[:, 1:])  # shifted by one, newest at the end
    # last_action block: newest entry is the action just applied (external joint order)
    assert np.array_equal(o1[:, 33 * H:].reshape(n, H, 12)[:, -1], a)


def test_wrap_to_pi():
    for a, w in ((0.0, 0.0), (3.0, 3.0), (-3.0, -3.0), (4.0, 4.0 - 2 * np.pi), (-4.0, -4.0 + 2 * np.pi)):
        assert abs(O.wrap_to_pi(a) - w) < 1e-6


def test_episode_timeout_and_reset_semantics(cfg):
    c = cfg.copy()
    n = 16
    orc = O.Oracle(c, n, seed=5)
    assert orc.max_episode_length == 1000  # 20 s / (4 * 0.005 s)  V/velocity_env_cfg.py:302-305
    orc.observe()
    ep = np.full(n, 998, np.int64); ep[0] = 10
    orc.episode_length = ep
    _, _, term, trunc = orc.step(np.zeros((n, 12), np.float32))
    assert not trunc.any()
    _, _, term, trunc = orc.step(np.zeros((n, 12), np.float32))
    assert trunc[1:].all() and not trunc[0]
    assert (orc.episode_length[1:] == 0).all() and orc.episode_length[0] == 12
    st = orc.get_state(["last_action", "fresh", "joint_pos"])
    assert (st["fresh"][1:, 0] & 1).all()  # delay line empty after the reset; the observation history was refilled
    np.testing.assert_allclose(st["joint_pos"][1], np.array(c.default_joint_pos[:]), atol=1e-6)


def test_manager_velocities_refer_to_the_link_centre_of_mass(cfg):
    """isaaclab 2.1.0 ArticulationData: poses are of the link frame, velocities of the link's centre of mass.
    A pelvis spinning about its own origin (origin velocity 0) therefore has root_lin_vel_b = w x r_com, with r_com the inertial
    origin of the pelvis LINK (h12_12dof.urdf:16).  Checked through lin_vel_z_l2 and the base / yaw-frame tracking terms."""
    c = cfg.copy()
    for t in (1, 12, 14):  # track_lin_vel_xy_yaw_frame_exp, lin_vel_z_l2, track_lin_vel_xy_exp
        c.rew_weight[t] = 1.0
    n = 4
    o = O.Oracle(c, n, seed=1)
    o.observe()
    s = o.get_state(["root_pos", "root_quat", "joint_pos"])
    yaw = 0.7
    quat = np.tile(np.array([np.cos(yaw / 2), 0, 0, np.sin(yaw / 2)], np.float32), (n, 1))
    w = np.array([[1.0, 0, 0], [0, 2.0, 0], [0, 0, 3.0], [1.0, -2.0, 0.5]])  # body-frame angular velocity
    qvel = np.zeros((n, 18)); qvel[:, 3:6] = w
    cmd = np.tile(np.array([0.3, -0.2, 0.1], np.float32), (n, 1))
    o.set_state({"command": cmd})
    post = {"pre_reset_qpos": np.concatenate([s["root_pos"], quat, s["joint_pos"]], axis=1), "pre_reset_qvel": qvel,
            "pre_reset_timers": np.zeros((n, 8)), "slot_force_hist": np.zeros((n, 18)), "applied_torque": np.zeros((n, 12)),
            "joint_acc": np.zeros((n, 12)), "foot_vel": np.zeros((n, 6))}
    o.step_injected(np.zeros((n, 12), np.float32), post)
    r = o.get_state(["reward_terms"])["reward_terms"] / (c.sim_dt * c.decimation)
    rc = np.array(c.root_link_com[:], np.float64)
    assert np.allclose(rc, [-0.0004, 3.7e-05, -0.046864], atol=1e-7)
    vb = np.cross(w, rc)  # base frame; the yaw frame differs from it by the roll/pitch only, which are 0 here
    np.testing.assert_allclose(r[:, 12], vb[:, 2] ** 2, rtol=1e-5, atol=1e-9)
    want = np.exp(-((cmd[:, 0] - vb[:, 0]) ** 2 + (cmd[:, 1] - vb[:, 1]) ** 2) / c.track_std ** 2)
    np.testing.assert_allclose(r[:, 14], want, rtol=1e-5)
    np.testing.assert_allclose(r[:, 1], want, rtol=1e-5)
    assert np.abs(vb[:, :2]).max() > 0.04  # the convention matters at the 5 cm/s level
