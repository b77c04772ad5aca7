"""Golden vectors produced by the reference's OWN code (tests/golden/make_ref_goldens.py, run where /root/reference exists):
  * feet_air_time / feet_air_time_positive_biped   (velocity/mdp/rewards.py:13-62)
  * CircularBuffer history + term-major flatten    (utils/history/circular_buffer.py, observation_manager.py:335-355)
  * deployment ObservationHandler                  (biped_deploy/controllers/rl.py:34-121)
  * dead-zone command class, dead zone 0           (utils/mdp/commands.py:41-96; survival statistics of a command)
  * contact / limit idioms                         (utils/cat/constraints.py:22-31,86-99,161-168: in-tree bodies of the
                                                    expressions inside illegal_contact, contact_forces, joint_pos_limits)
The CPU oracle is pinned against them here; the CUDA path is checked against the same files in the gpu-marked tests."""
import os

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
BFS = [0, 6, 1, 7, 2, 8, 3, 9, 4, 10, 5, 11]  # external (PhysX breadth-first) joint i -> MJCF joint


class _OracleBackend:
    def __init__(self, cfg, n, seed=1):
        from oracle.oracle import Oracle
        self.o = Oracle(cfg, n, seed=seed, threads=4)

    def set_state(self, d):
        self.o.set_state(d)

    def reset(self, ids):
        self.o.reset(np.asarray(ids, np.int64))

    def observe(self):
        return self.o.observe()

    def step(self, a):
        _, _, t, u = self.o.step(a)
        return t | u

    def get_state(self, names):
        return self.o.get_state(names)

    def close(self):
        pass


class _GpuBackend:
    def __init__(self, cfg, n, seed=1):
        import torch
        from h1v2_isaac_b200.backend import H1v2Sim
        self.torch, self.s = torch, H1v2Sim(n, cfg, device="cuda:0", seed=seed)

    def set_state(self, d):
        self.s.set_state(d)

    def reset(self, ids):
        self.s.reset(self.torch.as_tensor(np.asarray(ids, np.int64)).cuda())

    def observe(self):
        return self.s.observe().cpu().numpy()

    def step(self, a):
        _, _, t, u = self.s.step(self.torch.from_numpy(a).cuda())
        return (t | u).cpu().numpy()

    def get_state(self, names):
        return {k: v.cpu().numpy() for k, v in self.s.get_state(names).items()}

    def close(self):
        self.s.close()


def _history_case(backend_cls, cfg):
    g = np.load(os.path.join(GOLD, "obs_history.npz"))
    T, NE = g["reset"].shape
    c = cfg.copy(); c.enable_corruption = 0
    q0 = np.array(list(c.default_joint_pos), np.float32)
    b = backend_cls(c, NE)
    worst = 0.0
    for t in range(T):
        ids = np.nonzero(g["reset"][t])[0]
        if t > 0 and len(ids):
            b.reset(ids)
        jp, jv = np.zeros((NE, 12), np.float32), np.zeros((NE, 12), np.float32)
        jp[:, BFS] = g["qrel"][t] + q0[BFS]
        jv[:, BFS] = g["qvel"][t]
        b.set_state({"root_quat": g["quat"][t], "root_ang_vel": g["ang"][t], "command": g["cmd"][t], "joint_pos": jp, "joint_vel": jv,
                     "last_action": g["act"][t]})
        obs = b.observe()
        worst = max(worst, float(np.abs(obs - g["obs"][t]).max()))
        np.testing.assert_allclose(obs, g["obs"][t], rtol=1e-5, atol=2e-6, err_msg=f"step {t}")
    b.close()
    return worst


def _deploy_case(backend_cls, cfg, fixture="deploy_obs.npz"):
    g = np.load(os.path.join(GOLD, fixture))
    if fixture == "deploy_obs_rsl.npz":  # the Rsl id's format == the env.yaml the reference ships (history 6, scaled gyro / joint velocity)
        from h1v2_isaac_b200._capi import rsl_config
        cfg = rsl_config()
    c = cfg.copy(); c.enable_corruption = 0
    for i in range(12):
        c.joint_perm[i] = i  # deployment / Rsl order: MJCF leg-major (robots/h12.py:40-53 "preserved order for sim2sim")
    H = c.history_length
    assert g["obs"].shape[1] == 45 * H
    b = backend_cls(c, 1)
    for t in range(g["obs"].shape[0]):
        cmd = g["obs"][t][6 * H + 3 * (H - 1):9 * H]  # newest command of the golden row (ObservationHandler.generated_commands)
        expect = (g["cmd_unit"][t] + 1) / 2 * (g["cmd_upper"] - g["cmd_lower"]) + g["cmd_lower"]
        np.testing.assert_allclose(cmd, expect, atol=1e-6)
        b.set_state({"root_quat": g["quat"][t][None], "root_ang_vel": g["ang"][t][None], "command": cmd[None], "joint_pos": g["q"][t][None],
                     "joint_vel": g["qd"][t][None], "last_action": g["act"][t][None]})
        np.testing.assert_allclose(b.observe()[0], g["obs"][t], rtol=1e-5, atol=2e-6, err_msg=f"step {t}")
    b.close()


def test_oracle_history_against_reference_circular_buffer(cfg):
    assert _history_case(_OracleBackend, cfg) < 2e-6


@pytest.mark.parametrize("fixture", ["deploy_obs.npz", "deploy_obs_rsl.npz"])
def test_oracle_against_reference_deploy_observation_handler(cfg, fixture):
    _deploy_case(_OracleBackend, cfg, fixture)


@pytest.mark.parametrize("thr", [0.4, 0.5])
def test_oracle_feet_air_time_against_reference_functions(cfg, thr):
    from oracle.oracle import Oracle
    g = np.load(os.path.join(GOLD, "feet_air_time.npz"))
    n = g["cmd"].shape[0]
    c = cfg.copy()
    c.feet_air_threshold = thr
    c.rew_weight[3], c.rew_weight[16] = 0.75, 1.0  # feet_air_time_positive_biped, feet_air_time (L2)
    o = Oracle(c, n, seed=2, threads=8)
    o.observe()
    o.set_state({"command": g["cmd"]})
    s = o.get_state(["root_pos", "root_quat", "joint_pos"])
    timers = np.stack([g["cur_air"], g["last_air"], g["cur_con"], g["last_con"]], axis=-1)  # [N,2,4]
    post = {"pre_reset_qpos": np.concatenate([s["root_pos"], s["root_quat"], s["joint_pos"]], axis=1), "pre_reset_qvel": np.zeros((n, 18)),
            "pre_reset_timers": timers.reshape(n, 8), "slot_force_hist": np.zeros((n, 18)), "applied_torque": np.zeros((n, 12)),
            "joint_acc": np.zeros((n, 12)), "foot_vel": np.zeros((n, 6))}
    o.step_injected(np.zeros((n, 12), np.float32), post)
    r = o.get_state(["reward_terms"])["reward_terms"]
    dt = c.sim_dt * c.decimation
    np.testing.assert_allclose(r[:, 3] / (0.75 * dt), g[f"biped_thr{thr}"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(r[:, 16] / (1.0 * dt), g[f"l2_thr{thr}"], rtol=1e-5, atol=1e-6)
    assert (g[f"biped_thr{thr}"] > 0).sum() > 100 and (g[f"l2_thr{thr}"] != 0).sum() > 50  # the fixture exercises both branches


def test_oracle_contact_and_limit_terms_against_reference_constraint_functions(cfg):
    """illegal_contact (termination mask, bit-exact), contact_forces and joint_pos_limits (1e-5) against the in-tree functions
    that hold the same expressions: any_b max_h |F| > 1.0, max_h |F| - limit, max(lo - q, q - hi) (clipped at 0 and summed by
    the upstream reward terms)."""
    from oracle.oracle import Oracle
    g = np.load(os.path.join(GOLD, "contact_limit_idioms.npz"))
    n = g["joint_pos"].shape[0]
    c = cfg.copy()
    c.mask_illegal_slots = 0b111100  # knee links, torso, pelvis: the bodies the fixture's `illegal` was evaluated on
    c.rew_weight[19], c.mask_contact_forces_slots, c.contact_forces_threshold = -1.0e-3, 0b11, 800.0
    c.rew_weight[5], c.mask_pos_limits = -1.0, 0xFFF
    o = Oracle(c, n, seed=2, threads=8)
    o.observe()
    s = o.get_state(["root_pos", "root_quat"])
    norms = np.linalg.norm(g["force_hist"].astype(np.float64), axis=-1).astype(np.float32)  # [N, H, B]
    post = {"pre_reset_qpos": np.concatenate([s["root_pos"], s["root_quat"], g["joint_pos"]], axis=1), "pre_reset_qvel": np.zeros((n, 18)),
            "pre_reset_timers": np.zeros((n, 8)), "slot_force_hist": norms.transpose(0, 2, 1).reshape(n, 18), "applied_torque": np.zeros((n, 12)),
            "joint_acc": np.zeros((n, 12)), "foot_vel": np.zeros((n, 6))}
    _, _, term, trunc = o.step_injected(np.zeros((n, 12), np.float32), post)
    assert np.array_equal(term.astype(bool), g["illegal"]) and not trunc.any()
    assert 0.3 < g["illegal"].mean() < 0.8
    r = o.get_state(["reward_terms"])["reward_terms"]
    dt = c.sim_dt * c.decimation
    np.testing.assert_allclose(r[:, 19] / (-1.0e-3 * dt), np.clip(g["foot_force_800"], 0, None).sum(1), rtol=1e-5, atol=1e-4)
    np.testing.assert_allclose(r[:, 5] / (-1.0 * dt), np.clip(g["pos_limit"], 0, None).sum(1), rtol=1e-5, atol=1e-6)
    assert (g["foot_force_800"] > 0).sum() > 50 and (g["pos_limit"] > 0).sum() > 1000


def _deadzone_case(backend_cls, steps):
    """The Rsl id's command class against statistics of the reference's own _update_command (tests/golden/deadzone_command.json):
    with velocity_deadzone = 0 every call zeroes the xy command of half of ALL envs (reference: exactly n // 2 by randperm; here
    an independent draw per env with the same probability), never touches |yaw| (standing envs included), and flips the sign
    of the yaw-rate command with probability physics_dt / max_episode_length_s."""
    import json
    from h1v2_isaac_b200._capi import rsl_config
    g = json.load(open(os.path.join(GOLD, "deadzone_command.json")))
    n, calls = g["n_envs"], g["calls"]
    assert g["zero_xy_after_call"][0] == n // 2 and g["standing_envs_yaw_untouched_by_zeroing"]  # what the reference does
    b = backend_cls(rsl_config(), n, seed=3)
    b.observe()
    rng = np.random.default_rng(4)
    cmd = rng.uniform(0.2, 1.0, (n, 3)).astype(np.float32) * rng.choice([-1.0, 1.0], (n, 3)).astype(np.float32)
    standing = np.zeros((n, 1), np.int32); standing[:64] = 1
    b.set_state({"command": cmd, "is_standing": standing, "time_left": np.full((n, 1), 100.0, np.float32)})
    a = np.zeros((n, 12), np.float32)
    prev, prev_tl, flips, trials = cmd.copy(), np.full(n, 100.0, np.float32), 0, 0
    alive = np.ones(n, bool)  # envs whose command has not been redrawn since the start (no reset, no resample)
    for k in range(steps):
        done = b.step(a)
        st = b.get_state(["command", "time_left"])
        c, tl = st["command"], st["time_left"][:, 0]
        same = ~done & (tl < prev_tl)  # this step neither reset nor resampled the env's command
        alive &= same
        if k < calls:
            zero = ((c[:, 0] == 0) & (c[:, 1] == 0))[alive]
            q = 0.5 ** (k + 1)
            sigma = np.sqrt(alive.sum() * q * (1 - q))
            want = g["zero_xy_after_call"][k] * alive.sum() / n
            assert abs(zero.sum() - want) <= 5 * sigma + 2, (k, int(zero.sum()), want)
            assert alive[:64].sum() > 32  # standing envs are part of the next check: the override does not zero them
            assert np.array_equal(np.abs(c[alive, 2]), np.abs(cmd[alive, 2])), "|yaw-rate command| never changes between resamples"
        assert np.array_equal(np.abs(c[same, 2]), np.abs(prev[same, 2]))
        flips += int(((c[same, 2] == -prev[same, 2]) & (prev[same, 2] != 0)).sum()); trials += int(same.sum())
        prev, prev_tl = c, tl
    b.close()
    p_ref, p = g["yaw_flips"] / g["flip_trials"], flips / trials
    assert abs(p_ref - g["physics_dt"] / g["max_episode_length_s"]) < 2e-5
    assert abs(p - p_ref) <= 5 * np.sqrt(p_ref / trials) + 1e-6, (flips, trials, p_ref)
    return flips, trials


def _balanced_case(backend_cls, deadzone):
    """Positive dead zone: the reference's balancing keeps exactly n // 2 envs inside (golden: 2048 of 4096 after every call, from 30 / 126
    before the first); the per-env restatement with the census of the previous step must settle at the same half, within the binomial
    spread, from the first step on, and keep the envs it does not move untouched."""
    import json
    from h1v2_isaac_b200._capi import rsl_config
    g = json.load(open(os.path.join(GOLD, "deadzone_command.json")))
    ref = g["in_deadzone_count_before_and_after_each_call"][str(deadzone)]
    n = g["n_envs"]
    assert all(c == n // 2 for c in ref[1:]) and ref[0] < 0.05 * n
    c = rsl_config(); c.velocity_deadzone = deadzone
    b = backend_cls(c, n, seed=6)
    b.observe()
    rng = np.random.default_rng(7)
    cmd = rng.uniform(-1, 1, (n, 3)).astype(np.float32)
    b.set_state({"command": cmd, "time_left": np.full((n, 1), 100.0, np.float32)})
    a = np.zeros((n, 12), np.float32)
    fracs, prev = [], cmd
    for k in range(12):
        done = b.step(a)
        st = b.get_state(["command", "time_left"])
        cur = st["command"]
        inside = np.hypot(cur[:, 0], cur[:, 1]) < deadzone
        fracs.append(float(inside[~done].mean()))
        # an env is either untouched, zeroed (was outside), or redrawn with a fresh time_left (was inside)
        was_inside = np.hypot(prev[:, 0], prev[:, 1]) < deadzone
        same = (cur[:, :2] == prev[:, :2]).all(axis=1)
        zeroed = ~same & (cur[:, 0] == 0) & (cur[:, 1] == 0)
        redrawn = ~same & ~zeroed
        ok = ~done
        assert not (zeroed & was_inside & ok).any() or deadzone == 0
        assert (was_inside | ~redrawn | ~ok).all() and (st["time_left"][redrawn & ok, 0] <= 8.0).all()
        prev = cur
    b.close()
    sigma = 0.5 / np.sqrt(n)
    assert abs(fracs[0] - (ref[0] / n + 0.5 * (1 - ref[0] / n))) < 5 * sigma, fracs  # half of the envs outside go in at once
    assert all(abs(f - 0.5) < 6 * sigma for f in fracs[1:]), fracs


@pytest.mark.parametrize("deadzone", [0.1, 0.2])
def test_oracle_deadzone_balancing_against_reference_counts(deadzone):
    _balanced_case(_OracleBackend, deadzone)


@pytest.mark.gpu
@pytest.mark.parametrize("deadzone", [0.1, 0.2])
def test_cuda_deadzone_balancing_against_reference_counts(deadzone):
    _balanced_case(_GpuBackend, deadzone)


def test_oracle_deadzone_command_against_reference_statistics():
    _deadzone_case(_OracleBackend, 14)


@pytest.mark.gpu
def test_cuda_deadzone_command_against_reference_statistics():
    flips, trials = _deadzone_case(_GpuBackend, 600)
    assert flips > 300 and trials > 1_500_000


@pytest.mark.gpu
def test_cuda_history_against_reference_circular_buffer(cfg):
    assert _history_case(_GpuBackend, cfg) < 2e-6


@pytest.mark.gpu
@pytest.mark.parametrize("fixture", ["deploy_obs.npz", "deploy_obs_rsl.npz"])
def test_cuda_against_reference_deploy_observation_handler(cfg, fixture):
    _deploy_case(_GpuBackend, cfg, fixture)
