"""Golden vectors produced by the reference's OWN code (tests/golden/make_ref_goldens.py, run where /root/reference exists):
  * feet_air_time / feet_air_time_positive_biped   (velocity/mdp/rewards.py:13-62)
  * CircularBuffer history + term-major flatten    (utils/history/circular_buffer.py, observation_manager.py:335-355)
  * deployment ObservationHandler                  (biped_deploy/controllers/rl.py:34-121)
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


def _deploy_case(backend_cls, cfg):
    g = np.load(os.path.join(GOLD, "deploy_obs.npz"))
    c = cfg.copy(); c.enable_corruption = 0
    for i in range(12):
        c.joint_perm[i] = i  # deployment / Rsl order: MJCF leg-major (robots/h12.py:40-53 "preserved order for sim2sim")
    b = backend_cls(c, 1)
    for t in range(g["obs"].shape[0]):
        cmd = g["obs"][t][60 + 27:90]  # newest command of the golden row (ObservationHandler.generated_commands)
        expect = (g["cmd_unit"][t] + 1) / 2 * (g["cmd_upper"] - g["cmd_lower"]) + g["cmd_lower"]
        np.testing.assert_allclose(cmd, expect, atol=1e-6)
        b.set_state({"root_quat": g["quat"][t][None], "root_ang_vel": g["ang"][t][None], "command": cmd[None], "joint_pos": g["q"][t][None],
                     "joint_vel": g["qd"][t][None], "last_action": g["act"][t][None]})
        np.testing.assert_allclose(b.observe()[0], g["obs"][t], rtol=1e-5, atol=2e-6, err_msg=f"step {t}")
    b.close()


def test_oracle_history_against_reference_circular_buffer(cfg):
    assert _history_case(_OracleBackend, cfg) < 2e-6


def test_oracle_against_reference_deploy_observation_handler(cfg):
    _deploy_case(_OracleBackend, cfg)


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


@pytest.mark.gpu
def test_cuda_history_against_reference_circular_buffer(cfg):
    assert _history_case(_GpuBackend, cfg) < 2e-6


@pytest.mark.gpu
def test_cuda_against_reference_deploy_observation_handler(cfg):
    _deploy_case(_GpuBackend, cfg)
