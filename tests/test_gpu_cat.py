"""GPU: the Constraints-as-Terminations tail (h1v2_cat_step) against the numpy restatement oracle/cat_oracle.py, which is pinned to the
reference's own constraint functions and CaT class by tests/test_cat_oracle.py.  The tail is evaluated on the values the step kernel
left (pre-reset state, contact history, torques, the command before its update), so the comparison is on identical inputs."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ILLEGAL, FEET = [2, 3, 4, 5], [0, 1]


@pytest.mark.parametrize("epw", [16, 8, 4])  # 16: the plain instantiation of the CaT step; 8 / 4: the mirror-lane ones (two mirrors per lane: what 4096 envs run; four: <= 2368 envs)
def test_cat_tail_matches_the_pinned_oracle(epw):
    import torch
    from h1v2_isaac_b200._capi import CSTR_COL0, CSTR_NAMES, rsl_config
    from h1v2_isaac_b200.backend import H1v2Sim
    from oracle import cat_oracle as O
    from oracle import oracle as PO
    c = rsl_config()
    c.cat_enable = 1
    c.velocity_deadzone = 0.2  # the CaT cfg's command dead zone (cat_env_cfg.py:48,114)
    c.reserved[2] = epw
    n = 1024
    cat, twin = H1v2Sim(n, c, device="cuda:0", seed=21), H1v2Sim(n, c, device="cuda:0", seed=21, diagnostics=True)
    cat.observe(); twin.observe()
    ep = np.random.default_rng(1).integers(0, 1000, n)  # as rsl_rl's init_at_random_ep_len: some envs time out during the test
    ep[:8] = 995
    for s_ in (cat, twin):
        s_.episode_length_buf.copy_(torch.from_numpy(ep).cuda())
    ref, clearance = O.CaT(c.cat_tau, c.cat_min_p), O.FootClearance()
    max_p = dict(zip(CSTR_NAMES, list(c.cat_max_p)))
    soft = None
    gen = torch.Generator(device="cuda").manual_seed(3)
    sums_v = {k: np.zeros(n, np.float32) for k in CSTR_NAMES}; sums_p = {k: np.zeros(n, np.float32) for k in CSTR_NAMES}
    seen = {k: 0 for k in CSTR_NAMES}
    n_reset = 0
    for step in range(40):
        act = torch.randn((n, 12), device="cuda", generator=gen) * (8.0 if step % 4 == 0 else 3.0)  # action scale 0.25: large enough to fall
        cmd = twin.get_state(["command"])["command"].cpu().numpy()  # what the constraints of this step read
        if step == 12:
            cat.set_constraint_max_p([1.0] + [0.1] * 9); max_p = dict(zip(CSTR_NAMES, [1.0] + [0.1] * 9))
        obs, rew, dones, trunc = cat.cat_step(act)
        obs2, rew2, term2, trunc2 = twin.step(act)  # the plain step on a twin: same physics, raw reward, diagnostics
        assert torch.equal(obs, obs2) and torch.equal(trunc, trunc2)
        g = {k: v.cpu().numpy() for k, v in twin.get_state(["pre_reset_qpos", "pre_reset_qvel", "pre_reset_timers", "slot_force_hist", "applied_torque"]).items()}
        qpos, qvel = g["pre_reset_qpos"], g["pre_reset_qvel"]
        fh = np.zeros((n, 3, 6, 3), np.float32); fh[..., 0] = g["slot_force_hist"].reshape(n, 6, 3).transpose(0, 2, 1)
        if soft is None:
            lo, hi = np.array([r[0] for r in c.joint_range]), np.array([r[1] for r in c.joint_range])
            mid, half = 0.5 * (lo + hi), 0.5 * (hi - lo) * c.soft_limit_factor
            soft = np.stack([mid - half, mid + half], -1).astype(np.float32)
        quat = qpos[:, 3:7].astype(np.float64)
        w, x, y, z = quat.T
        grav = np.stack([-2 * (x * z - w * y), -2 * (y * z + w * x), -(1 - 2 * (x * x + y * y))], 1).astype(np.float32)  # R^T (0,0,-1)
        foot_z = np.stack([[PO.fk(qpos[i].astype(np.float64))[1][b][2] for b in (6, 12)] for i in range(n)]).astype(np.float32)
        touchdown = O.first_contact(g["pre_reset_timers"].reshape(n, 2, 4)[:, :, 2], c.sim_dt * c.decimation)
        raw = {
            "contact": O.contact(fh, ILLEGAL), "joint_position_limits": O.joint_position_limits(qpos[:, 7:], soft),
            "joint_velocity_limits": O.joint_velocity_limits(qvel[:, 6:], np.float32(c.joint_vel_limit)),
            "joint_torque_limits": O.joint_torque_limits(g["applied_torque"], np.array(list(c.effort_limit), np.float32)),
            "foot_contact_force": O.foot_contact_force(fh, FEET, c.cat_foot_force_limit),
            "no_move": O.no_move(cmd, qvel[:, 6:], c.cat_no_move_deadzone, c.cat_no_move_vel_limit),
            "base_orientation": O.base_orientation(grav, c.cat_orientation_limit), "base_height": O.base_height(qpos[:, 2], c.cat_height, c.cat_height_std),
            "foot_contact": O.foot_contact(fh, FEET),
            "foot_clearance": clearance(foot_z, touchdown, cmd, c.cat_clearance_min_height, c.cat_clearance_deadzone)}
        kraw, kprob, krm = cat.cat_debug()
        for t, k in enumerate(CSTR_NAMES):
            r = np.asarray(raw[k], np.float32); r = r[:, None] if r.ndim == 1 else r
            np.testing.assert_allclose(kraw[CSTR_COL0[t]:CSTR_COL0[t + 1]].T, r, rtol=2e-5, atol=2e-5, err_msg=f"{k} raw, step {step}")
            p = ref.add(k, kraw[CSTR_COL0[t]:CSTR_COL0[t + 1]].T, max_p[k])  # fed with the kernel's own raw values: isolates the CaT arithmetic
            np.testing.assert_allclose(krm[CSTR_COL0[t]:CSTR_COL0[t + 1]], ref.running_maxes[k][0], rtol=1e-6, err_msg=f"{k} running max, step {step}")
            np.testing.assert_allclose(kprob[CSTR_COL0[t]:CSTR_COL0[t + 1]].T, p, rtol=1e-5, atol=1e-7, err_msg=f"{k} probabilities, step {step}")
            seen[k] += int((p > 0).sum())
            sums_v[k] += (p.max(1) > 0); sums_p[k] += p.max(1)
        prob = ref.get_probs()
        reset = (term2 | trunc2).cpu().numpy()
        want_rew, want_dones = O.constrained_reward(rew2.cpu().numpy(), prob, reset)
        np.testing.assert_allclose(rew.cpu().numpy(), want_rew, rtol=1e-5, atol=1e-7)
        np.testing.assert_allclose(dones.cpu().numpy(), want_dones, rtol=1e-5, atol=1e-7)
        ep = ep + 1  # episode_length_buf += 1 precedes the constraints (cat_env.py:139)
        if reset.any():  # Episode_Constraint_* of the envs that reset (constraint_manager.py:185-203)
            n_reset += int(reset.sum())
            lg = cat.cat_log_host()
            assert lg[20] == reset.sum()
            for t, k in enumerate(CSTR_NAMES):
                np.testing.assert_allclose(lg[t], (sums_v[k][reset] / ep[reset]).mean() * 100, rtol=1e-4, atol=1e-5, err_msg=f"violation log {k}")
                np.testing.assert_allclose(lg[10 + t], (sums_p[k][reset] / ep[reset]).mean(), rtol=1e-4, atol=1e-6, err_msg=f"probability log {k}")
                sums_v[k][reset] = 0; sums_p[k][reset] = 0
            ep[reset] = 0
        assert np.array_equal(cat.episode_length_buf.cpu().numpy(), ep)
    assert n_reset >= 8
    # every constraint fired somewhere -- except the two that cannot in this backend: joint velocities are clamped to the velocity limit and
    # applied torques to the effort limit, so `|x| - limit` never exceeds 0 (it does in the oracle's own golden, tests/test_cat_oracle.py)
    never = {"joint_velocity_limits", "joint_torque_limits"}
    for k in CSTR_NAMES:
        assert (seen[k] == 0) == (k in never), f"{k}: fired {seen[k]} times"
    cat.close(); twin.close()


def test_cat_env_contract_and_constraint_curriculum():
    """gym.make of the CaT id (self-contained tree, pinned to the reference's cfg class by tests/test_boundary.py): CaTEnv's return
    signature (cat_env.py:193) -- float dones = termination probability, 1 on reset --, the constraint log keys, the reward scaled by
    1 - p against a plain twin, and modify_constraint_p (curriculums.py:20-42) driving the kernel's max_p from 1/20 upwards."""
    import torch
    from h1v2_isaac_b200 import tasks
    from h1v2_isaac_b200._capi import CSTR_NAMES
    from h1v2_isaac_b200.backend import H1v2Sim
    tasks.register()
    import gymnasium as gym
    n = 512
    env = gym.make(tasks.CAT_TASK_ID, cfg=tasks.cat_env_cfg(n))
    plain_cfg = env.kernel_cfg.copy()
    twin = H1v2Sim(n, plain_cfg, device="cuda:0", seed=42)
    obs, _ = env.reset(); twin.reset(None); twin.observe()
    assert obs["policy"].shape == (n, 270)
    assert abs(env.sim.cfg.cat_max_p[1] - 0.05) < 1e-7 and env.sim.cfg.cat_max_p[0] == 1.0  # the curriculum's starting point
    g = torch.Generator(device="cuda").manual_seed(0)
    scaled = 0
    for k in range(40):
        a = torch.randn((n, 12), device="cuda", generator=g) * (6.0 if k % 5 == 0 else 1.0)
        o, r, d, tr, ex = env.step(a)
        o2, r2, t2, u2 = twin.step(a)
        assert d.dtype == torch.float32 and tr.dtype == torch.bool and torch.equal(tr, u2) and torch.equal(o["policy"], o2)
        reset = t2 | u2
        assert (d[reset] == 1).all() and (d[~reset] < 1).all() and (d >= 0).all()
        p = torch.where(reset, torch.zeros_like(d), d)
        assert torch.allclose(r[~reset], r2[~reset] * (1 - p[~reset]), rtol=1e-5, atol=1e-7)
        scaled += int((p > 0).sum())
    assert scaled > 100
    keys = set(ex["log"])
    for nme in CSTR_NAMES:
        assert f"Episode_Constraint_violation/{nme}" in keys and f"Episode_Constraint_probability/{nme}" in keys
    assert abs(env.sim.cfg.cat_max_p[1] - 1.0 / (20 + (40 / 120000) * (4 - 20))) < 1e-6
    env.close(); twin.close()


@pytest.mark.parametrize("mode", ["rows", "assemble"])
def test_cat_step_host_and_graph_replay_match_the_device_path(mode, monkeypatch):
    """h1v2_cat_step_host (HOST buffers, both observation paths) and a CUDA-graph replay of h1v2_cat_step give exactly what eager
    device-path calls give: every piece of step-to-step CaT state (running maxima, their parity, the first-step flag, the
    dead-zone list) lives on the device.  An API reset of some envs restarts their per-term episode statistics
    (_reset_idx -> constraint_manager.reset, constraint_manager.py:185-214)."""
    import torch
    from h1v2_isaac_b200 import tasks
    from h1v2_isaac_b200.backend import H1v2Sim
    monkeypatch.setenv("H1V2_HOST_PATH", mode)
    c = tasks.cat_config()
    n = 768
    a_sim, b_sim, g_sim = (H1v2Sim(n, c, device="cuda:0", seed=9) for _ in range(3))
    for s_ in (a_sim, b_sim, g_sim):
        s_.observe()
    hobs = torch.empty((n, a_sim.obs_dim)).pin_memory(); hrew = torch.empty(n).pin_memory()
    hd = torch.empty(n).pin_memory(); hu = torch.empty(n, dtype=torch.uint8).pin_memory()
    # graph: static input / output tensors, two kernel launches captured
    act_s = torch.zeros((n, 12), device="cuda")
    gout = [torch.empty((n, g_sim.obs_dim), device="cuda"), torch.empty(n, device="cuda"), torch.empty(n, device="cuda"), torch.empty(n, dtype=torch.uint8, device="cuda")]
    gen = torch.Generator(device="cuda").manual_seed(5)
    graph = None
    ids = torch.tensor([0, 3, 700], device="cuda")
    for step in range(12):
        act = torch.randn((n, 12), device="cuda", generator=gen) * (6.0 if step % 3 == 0 else 2.0)
        o, r, d, u = a_sim.cat_step(act)
        b_sim.cat_step_host(act.cpu().pin_memory(), hobs, hrew, hd, hu)
        assert torch.equal(o.cpu(), hobs) and torch.equal(r.cpu(), hrew) and torch.equal(d.cpu(), hd) and torch.equal(u.cpu().to(torch.uint8), hu), step
        act_s.copy_(act)
        if step < 2:  # eager warm-up (function attributes), then capture once and replay
            g_sim.cat_step_into(act_s, *gout)
        else:
            if graph is None:
                torch.cuda.synchronize()
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph):
                    g_sim.cat_step_into(act_s, *gout)
            graph.replay()
        assert torch.equal(o, gout[0]) and torch.equal(r, gout[1]) and torch.equal(d, gout[2]) and torch.equal(u.to(torch.uint8), gout[3]), step
        if step == 6:
            for s_ in (a_sim, b_sim, g_sim):
                s_.reset(ids)
    la, lb = a_sim.cat_log_host(), b_sim.cat_log_host()
    assert np.array_equal(la, lb)
    for s_ in (a_sim, b_sim, g_sim):
        s_.close()
