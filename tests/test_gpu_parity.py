"""GPU parity tests proper: the CUDA path, called through the C-ABI, against the CPU oracle on identical seeded
states and actions (BASELINE.json north_star (a)(b)(c)):
  (a) reward and observation terms within 1e-5 relative on IDENTICAL post-physics states,
  (b) termination / reset masks and command-resample masks bit-exact,
  (c) physics within 1e-4 rad / 1e-3 rad/s for one physics step, bounded divergence over 1000 steps,
plus size-independent properties at BASELINE's full sizes (4096 / 32768 envs).
"""
import numpy as np
from h1v2_isaac_b200._capi import LOG_NAN_RESETS
import pytest

pytestmark = pytest.mark.gpu

PHYS = ["root_pos", "root_quat", "root_lin_vel", "root_ang_vel", "joint_pos", "joint_vel"]
SYNC = PHYS + ["last_action", "target_hist", "lag", "fresh", "command", "heading_target", "time_left", "is_standing",
               "is_heading", "cmd_metrics", "feet_timers", "episode_sums", "obs_history", "friction", "mass_add",
               "push_time_left", "terrain_level"]
POST = ["pre_reset_qpos", "pre_reset_qvel", "pre_reset_timers", "slot_force_hist", "applied_torque", "joint_acc", "foot_vel"]


def _mk(cfg, n, seed):
    import torch
    from h1v2_isaac_b200.backend import H1v2Sim
    from oracle.oracle import Oracle
    return torch, H1v2Sim(n, cfg, device="cuda:0", seed=seed, diagnostics=True), Oracle(cfg, n, seed=seed, threads=16)


def _np(d):
    return {k: v.detach().cpu().numpy() for k, v in d.items()}


def _resync(sim, orc, g):
    orc.set_state({k: g[k] for k in SYNC})
    orc.episode_length = sim.episode_length_buf.cpu().numpy()


def test_reset_state_matches_oracle(cfg):
    """create/reset: Philox draws, root pose, joint pose, lag, command resample -- all from the same counter-based stream."""
    torch, sim, orc = _mk(cfg, 512, 11)
    g, o = _np(sim.get_state(SYNC)), orc.get_state(SYNC)
    for k in SYNC:
        if g[k].dtype.kind == "i":
            assert np.array_equal(g[k], o[k]), k
        else:
            np.testing.assert_allclose(g[k], o[k], rtol=0, atol=2e-6, err_msg=k)
    np.testing.assert_allclose(sim.observe().cpu().numpy(), orc.observe(), rtol=0, atol=3e-6)
    # partial reset through the API
    ids = np.array([3, 77, 500], np.int64)
    sim.reset(torch.from_numpy(ids).cuda()); orc.reset(ids)
    g, o = _np(sim.get_state(SYNC)), orc.get_state(SYNC)
    for k in ("root_pos", "root_quat", "joint_pos", "command", "lag", "fresh"):
        np.testing.assert_allclose(g[k], o[k], rtol=0, atol=2e-6, err_msg=k)


@pytest.mark.parametrize("decimation,action_scale,epw", [(1, 1.0, 16), (1, 0.3, 16), (4, 1.0, 16), (1, 1.0, 8), (4, 1.0, 8), (1, 1.0, 4), (4, 1.0, 4)])
def test_physics_parity_from_identical_states(cfg, decimation, action_scale, epw):
    """(c): every control step starts from the SAME state on both sides (oracle re-synchronised to the GPU state).
    decimation=1 is the single physics step of the north star and is asserted LITERALLY: positions within 1e-4 rad / m and
    velocities within 1e-3 rad/s (m/s) on EVERY kept env-step.  Measured on B200 over 785 k env-steps per seed
    (tools/diag_fp64.py, profiles/r2_notes.md): max 8.5e-4 rad/s at action scale 1.0, 3.4e-4 at 0.3, q99.9 1.7e-4.  What is
    left is fp32 rounding inside hard landings (2-3 kN on one foot): the SAME kernel source built in double agrees with the
    oracle to 1.6e-4 (tests/test_gpu_fp64.py), and storing any single intermediate in float does not bring the error back.
    decimation=4 chains four such substeps without re-synchronising, so the single-step budget compounds: bounded at 4e-3.
    Envs within 2e-6 m (rad) of a contact (joint-limit) activation boundary at a substep start are skipped and counted:
    the soft contact switches on discontinuously at dist = 0, so float-vs-double rounding of the distance decides those.
    epw = 16: the plain instantiation with every lane pair on its own env (what >= 5625 envs, e.g. the 32768 of the north star, run).
    epw = 8: the mirror-lane instantiation 2369..5624 envs run (BASELINE configs[1], 4096 envs): lanes 16..31 mirror lanes 0..15
    and take every other row / sole point of the gradient evaluation, the line search and the step.  epw = 4: four mirrors per
    lane (lanes l, l+8, l+16, l+24), what <= 2368 envs run."""
    c = cfg.copy()
    c.reserved[2] = epw
    c.decimation = decimation
    c.max_delay = min(c.max_delay, 2 * decimation)
    n = 2048
    torch, sim, orc = _mk(c, n, 3)
    sim.observe(); orc.observe()
    rng = np.random.default_rng(0)
    errs = {k: [] for k in PHYS}
    steps = 24 * (4 // decimation)
    n_alive = n_boundary = 0
    for step in range(steps):
        a = (action_scale * rng.normal(size=(n, 12))).astype(np.float32)
        _, _, tg, ug = sim.step(torch.from_numpy(a).cuda())
        _, _, to, uo = orc.step(a)
        g, o = _np(sim.get_state(SYNC)), orc.get_state(PHYS)
        mc, ml = orc.activation_margin()
        alive = ~(to | uo | tg.cpu().numpy())
        keep = alive & (mc > 2e-6) & (ml > 2e-6)
        n_alive += int(alive.sum()); n_boundary += int((alive & ~keep).sum())
        for k in PHYS:
            errs[k].append(np.abs(g[k][keep] - o[k][keep]).max(axis=1))
        _resync(sim, orc, g)
    e = {k: np.concatenate(v) for k, v in errs.items()}
    print({k: (float(v.max()), float(np.quantile(v, 0.999)), float(np.quantile(v, 0.99))) for k, v in e.items()}, "env-steps", len(e["joint_pos"]),
          f"excluded on an activation boundary: {n_boundary} of {n_alive} ({100.0 * n_boundary / max(n_alive, 1):.3f} %)")
    assert len(e["joint_pos"]) > 0.7 * n * steps and n_boundary < 0.01 * n_alive
    for k in ("joint_pos", "root_pos", "root_quat"):
        assert e[k].max() < 1e-4, k
    vmax = 1e-3 if decimation == 1 else 4e-3
    for k in ("joint_vel", "root_lin_vel", "root_ang_vel"):
        assert e[k].max() < vmax, (k, float(e[k].max()))
        assert np.quantile(e[k], 0.99) < 5e-4 and np.quantile(e[k], 0.999) < 1e-3, k
    assert np.median(e["joint_vel"]) < 1e-4


def test_fall_contacts_and_termination_decisions_from_identical_states(cfg):
    """The colliders that only matter when the robot falls (shin capsules, torso box, pelvis sphere; h12_12dof.urdf:116-121,
    387-392) and the termination they trigger: from identical states the kernel's own physics must give the oracle's contact
    forces on those bodies (1 % + 0.5 N, envs on an activation boundary excluded) and, but for forces right at the 1 N
    threshold, the same termination decision."""
    n = 2048
    torch, sim, orc = _mk(cfg, n, 41)
    sim.observe(); orc.observe()
    rng = np.random.default_rng(8)
    n_fall_contacts = n_term = n_mismatch = n_steps = n_overflow = n_overflow_alive = 0
    worst = 0.0
    for step in range(48):
        a = rng.normal(size=(n, 12)).astype(np.float32)
        _, _, tg, ug = sim.step(torch.from_numpy(a).cuda())
        _, _, to, uo = orc.step(a)
        g = _np(sim.get_state(SYNC + ["slot_force", "slot_force_hist", "solver_iters"]))
        # a leg's active contact list holds 5 points (4 sole corners + 1): a dropped point is only acceptable in a step that
        # terminates anyway (shin / torso / pelvis on the ground), where the state is thrown away
        ovf = g["solver_iters"][:, 2] > 0
        n_overflow += int(ovf.sum()); n_overflow_alive += int((ovf & ~tg.cpu().numpy()).sum())
        o = orc.get_state(["slot_force", "slot_force_hist"])
        mc, ml = orc.activation_margin()
        keep = (mc > 2e-6) & (ml > 2e-6)
        tg = tg.cpu().numpy()
        # the force history of the last three substeps on the fall bodies (slots 2..5), |F| per substep
        hg, ho = g["slot_force_hist"].reshape(n, 6, 3)[keep][:, 2:], o["slot_force_hist"].reshape(n, 6, 3)[keep][:, 2:]
        active = ho > 5.0
        n_fall_contacts += int(active.sum())
        if active.any():
            rel = np.abs(hg - ho)[active] / (0.01 * ho[active] + 0.5)
            worst = max(worst, float(rel.max()))
        cmax = ho.max(axis=(1, 2))
        decided = (cmax < 0.9) | (cmax > 1.1)  # not within 10 % of the 1 N threshold
        n_mismatch += int((tg[keep] != to[keep])[decided].sum()); n_term += int(to.sum()); n_steps += int(keep.sum())
        _resync(sim, orc, g)
    print(f"fall-body force samples {n_fall_contacts}, worst error in units of (1 % + 0.5 N) {worst:.2f}, terminations {n_term}, decision mismatches {n_mismatch} of {n_steps}")
    print(f"contact-list overflows {n_overflow}, of which in envs that did not terminate in that step: {n_overflow_alive}")
    assert n_overflow_alive == 0
    assert n_fall_contacts > 300 and n_term > 100
    assert worst < 0.25  # measured 0.02: the fall-body forces agree to ~2e-4 relative
    assert n_mismatch == 0


def test_tail_parity_on_identical_states(cfg):
    """(a)+(b): the oracle's physics loop is replaced by the kernel's own post-physics values, so everything after the
    physics -- contact-sensor logic, terminations, the reward terms, reset, command resampling, noisy observation with
    history -- is compared on identical inputs: masks bit-exact, values to 1e-5."""
    _tail_parity(cfg, 4096, 60, 100)


def test_tail_parity_ideal_pd_and_union_rewards(cfg):
    """Same check on the H12_12DOF_IDEAL actuator variant (IdealPD: no delay line, A/robots/h12.py:117-206) with every
    one of the 22 reward slots switched on (the union of the H1-2 cfgs), history 6 and the Rsl observation
    scales (C12/rsl_env_cfg.py) -- the configuration space flatten_cfg can produce beyond the Flat id's defaults."""
    c = cfg.copy()
    c.min_delay = c.max_delay = 0
    for t in range(len(c.rew_weight)):
        if c.rew_weight[t] == 0.0:
            c.rew_weight[t] = -0.01 * (t + 1)
    c.history_length = 6
    c.scale_ang_vel, c.scale_joint_vel, c.action_scale = 0.25, 0.05, 0.25
    for i in range(12):
        c.joint_perm[i] = i
    c.reserved[2] = 16  # the plain instantiation (16 envs per warp, what >= 5625 envs run); the other tail tests run the mirror-lane one
    _tail_parity(c, 2048, 30, 30, min_term=0)  # action scale 0.25: nobody falls within 30 steps


@pytest.mark.parametrize("H", [1, 3])
def test_tail_parity_short_histories(cfg, H):
    """History 1 is the deployment format without stacking (D/controllers/rl.py with history_length 1, a 45-float observation: less
    than two 32-column passes of the flatten), 3 an odd length: ring indexing, first-push back-fill and flatten must not depend on
    the Flat id's 10."""
    c = cfg.copy()
    c.history_length = H
    _tail_parity(c, 1024, 16, 0, min_term=0)


@pytest.mark.parametrize("task", ["flat", "rsl"])
def test_natural_command_resample_parity(cfg, task):
    """CommandTerm.compute's own resample (time_left <= 0, not a reset): the draws come from Philox blocks 2-3 of the command
    stream, the new time_left from the cfg's range; masks and values must match the oracle bit for bit."""
    from h1v2_isaac_b200._capi import rsl_config
    c = rsl_config() if task == "rsl" else cfg
    n = 1024
    torch, sim, orc = _mk(c, n, 31)
    sim.observe(); orc.observe()
    tl = np.full((n, 1), 50.0, np.float32)
    tl[::3] = 0.03; tl[1::7] = 0.05  # these run out in the 2nd and 3rd step
    sim.set_state({"time_left": tl}); orc.set_state({"time_left": tl})
    n_nat = 0
    for step in range(4):
        a = np.zeros((n, 12), np.float32)
        _, _, tg, ug = sim.step(torch.from_numpy(a).cuda())
        g = _np(sim.get_state(SYNC + POST))
        _, _, to, uo = orc.step_injected(a, g)
        o = orc.get_state(SYNC)
        assert not to.any() and not uo.any() and not tg.any()
        assert np.array_equal(g["time_left"], o["time_left"]) and np.array_equal(g["command"][:, :2], o["command"][:, :2])
        if c.heading_command:  # the yaw-rate command is 0.5 * wrap(target - atan2(...)): float vs double atan2, not a draw
            np.testing.assert_allclose(g["command"][:, 2], o["command"][:, 2], atol=2e-6)
        else:
            assert np.array_equal(g["command"][:, 2], o["command"][:, 2])
        assert np.array_equal(g["is_standing"], o["is_standing"]) and np.array_equal(g["heading_target"], o["heading_target"])
        n_nat += int((g["time_left"][:, 0] > tl[:, 0]).sum())
        tl = g["time_left"].copy()
        _resync(sim, orc, g)
    lo = float(c.cmd_resample_time[0])
    assert n_nat >= n // 3 and (tl[::3, 0] >= lo - 0.1).all()
    sim.close()


@pytest.mark.parametrize("deadzone", [0.0, 0.2])
def test_tail_parity_rsl_task(deadzone):
    """the resolved cfg of Isaac-Velocity-Rsl-H12_12dof-v0 (C12/rsl_env_cfg.py:44-540; tests/golden/
    rsl_cfg_resolved.json pins it to the reference's own cfg classes): IdealPD, action scale 0.25, history 6, scaled gyro / joint
    velocity, 16 reward terms incl. the second joint-set instances and contact_forces at 800 N, per-env friction 0.1..1.25, and
    the dead-zone command class (T/utils/mdp/commands.py:41-96) whose zeroing / sign-flip masks must be bit-exact.  Half way
    through, the reward weights are replaced through h1v2_set_reward_weights (the curriculum's modify_reward_weight path)."""
    from h1v2_isaac_b200._capi import rsl_config
    c = rsl_config()
    assert c.command_class == 1 and c.history_length == 6
    c.velocity_deadzone = deadzone  # 0.2 = the CaT cfg's value: both balancing branches (zeroing and re-drawing) run
    w2 = [w * (1.5 if i % 2 else 0.5) for i, w in enumerate(c.rew_weight)]
    stats = _tail_parity(c, 2048, 40, 20, min_term=0, reweight=(20, w2))
    if deadzone == 0.0:
        # C12/rsl_env_cfg.py:98 velocity_deadzone = 0.0: every step half of ALL envs lose their xy command, so a command survives
        # k steps with probability 2^-k -- the task as configured is in-place stepping and turning
        assert stats["xy_zero_frac"][0] > 0.4 and stats["xy_zero_frac"][5] > 0.95 and stats["xy_zero_frac"][-1] > 0.99
    else:  # balanced: about half of the envs keep a live xy command
        assert all(0.4 < f < 0.6 for f in stats["xy_zero_frac"][1:])
    assert stats["yaw_flips"] > 0


def _tail_parity(cfg, n, steps, min_events, min_term=None, reweight=None):
    from h1v2_isaac_b200._capi import LOG_ERR_XY, LOG_REW0, LOG_TERM_CONTACT, LOG_TERM_TIMEOUT, NUM_REW
    torch, sim, orc = _mk(cfg, n, 21)
    sim.observe(); orc.observe()
    ep = np.random.default_rng(1).integers(0, 1000, n)  # rsl_rl's init_at_random_ep_len: exercises time-outs
    sim.episode_length_buf.copy_(torch.from_numpy(ep).cuda()); orc.episode_length = ep
    rng = np.random.default_rng(2)
    n_term = n_trunc = n_resample = 0
    stats = {"xy_zero_frac": [], "yaw_flips": 0}
    prev = _np(sim.get_state(["time_left", "command"]))
    prev_level = _np(sim.get_state(["terrain_level"]))["terrain_level"]
    for step in range(steps):
        if reweight is not None and step == reweight[0]:
            sim.set_reward_weights(reweight[1]); orc.set_reward_weights(reweight[1])
        a = (rng.normal(size=(n, 12)) * (0.3 if step % 3 else 1.0)).astype(np.float32)
        og, rg, tg, ug = sim.step(torch.from_numpy(a).cuda())
        g = _np(sim.get_state(SYNC + POST + ["reward_terms"]))
        # foot_vel is NOT injected: the oracle derives body_lin_vel_w of the feet from the injected state with its own kinematics
        oo, ro, to, uo = orc.step_injected(a, {k: v for k, v in g.items() if k != "foot_vel"})
        o = orc.get_state(SYNC + ["reward_terms", "foot_vel"])
        np.testing.assert_allclose(g["foot_vel"], o["foot_vel"], rtol=1e-4, atol=2e-5, err_msg=f"foot velocity, step {step}")
        tg, ug, og, rg = tg.cpu().numpy(), ug.cpu().numpy(), og.cpu().numpy(), rg.cpu().numpy()
        # (b) masks bit-exact
        assert np.array_equal(tg, to), f"terminated mask, step {step}"
        assert np.array_equal(ug, uo), f"truncated mask, step {step}"
        for k in ("is_standing", "is_heading", "lag", "fresh"):
            assert np.array_equal(g[k], o[k]), f"{k}, step {step}"
        assert np.array_equal(g["time_left"], o["time_left"]), "command resample mask (time_left is float32 on both sides)"
        assert np.array_equal(sim.episode_length_buf.cpu().numpy(), orc.episode_length)
        assert np.array_equal(g["feet_timers"], o["feet_timers"]), "contact-sensor air/contact timers"
        assert np.array_equal(g["terrain_level"], o["terrain_level"]), f"terrain level after the resets of step {step}"
        if cfg.terrain_enable:
            stats.setdefault("level_moves", 0)
            stats["level_moves"] += int((g["terrain_level"] != prev_level).sum()); prev_level = g["terrain_level"].copy()
            assert abs(float(sim.terrain_log_buf[1]) - orc.terrain_level_mean()) < 1e-4
        # (a) values
        np.testing.assert_allclose(g["reward_terms"], o["reward_terms"], rtol=1e-5, atol=2e-7, err_msg=f"reward terms, step {step}")
        np.testing.assert_allclose(rg, ro, rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(og, oo, rtol=1e-5, atol=2e-6, err_msg=f"observation, step {step}")
        for k in ("command", "heading_target", "cmd_metrics", "episode_sums", "root_pos", "root_quat", "joint_pos", "target_hist"):
            np.testing.assert_allclose(g[k], o[k], rtol=1e-5, atol=2e-6, err_msg=f"{k}, step {step}")
        resampled = g["time_left"][:, 0] > prev["time_left"][:, 0]  # natural resamples and those of resets
        n_term += int(to.sum()); n_trunc += int(uo.sum()); n_resample += int(resampled.sum())
        if cfg.command_class == 1:  # dead-zone class: the zeroing and sign-flip draws are masks, compare them exactly
            assert np.array_equal(g["command"], o["command"]), f"dead-zone command, step {step}"
            stats["xy_zero_frac"].append(float(((g["command"][:, 0] == 0) & (g["command"][:, 1] == 0)).mean()))
            redrawn = (g["command"][:, :2] != prev["command"][:, :2]).any(axis=1) & ~(g["command"][:, :2] == 0).all(axis=1)  # dead-zone re-activation
            same = ~(to | uo | resampled | redrawn) & (prev["command"][:, 2] != 0)
            stats["yaw_flips"] += int((g["command"][same, 2] == -prev["command"][same, 2]).sum())
            assert np.all(np.abs(g["command"][same, 2]) == np.abs(prev["command"][same, 2]))
        lg, lo = sim.log_host(), orc.log()
        assert lg[0] == lo[0] and lg[LOG_TERM_TIMEOUT] == lo[LOG_TERM_TIMEOUT] and lg[LOG_TERM_CONTACT] == lo[LOG_TERM_CONTACT]  # reset counts
        if lo[0] > 0:
            np.testing.assert_allclose(lg[LOG_REW0:LOG_REW0 + NUM_REW], lo[LOG_REW0:LOG_REW0 + NUM_REW], rtol=2e-4, atol=1e-6)  # Episode_Reward/* (float atomics)
            np.testing.assert_allclose(lg[LOG_ERR_XY:LOG_ERR_XY + 2], lo[LOG_ERR_XY:LOG_ERR_XY + 2], rtol=2e-4, atol=1e-6)
        prev = {"time_left": g["time_left"], "command": g["command"]}
        _resync(sim, orc, g)
    print(f"terminated {n_term}, truncated {n_trunc}, command resamples {n_resample}", {k: (v if not isinstance(v, list) else v[:8]) for k, v in stats.items()})
    assert n_term >= (min_events if min_term is None else min_term) and n_trunc > min_events and n_resample > min_events
    return stats


@pytest.mark.parametrize("epw", [16, 8, 4])  # 16: the plain instantiation (>= 5625 envs); 8 / 4: the mirror-lane ones (two / four mirrors per lane)
def test_bounded_divergence_over_1000_steps(cfg, epw):
    """(c) free-running: 1000 control steps of both implementations under a stabilising (zero) action stay statistically
    together and finite; individual trajectories are allowed to separate (contact dynamics are chaotic)."""
    n = 256
    c = cfg.copy()
    c.reserved[2] = epw
    c.enable_corruption = 0
    torch, sim, orc = _mk(c, n, 9)
    sim.observe(); orc.observe()
    a = np.zeros((n, 12), np.float32)
    at = torch.from_numpy(a).cuda()
    rg_sum, ro_sum, dq = 0.0, 0.0, []
    for step in range(1000):
        _, rg, tg, _ = sim.step(at)
        _, ro, to, _ = orc.step(a)
        rg_sum += float(rg.mean()); ro_sum += float(ro.mean())
        if step in (0, 4, 24, 99):
            g, o = _np(sim.get_state(["joint_pos"])), orc.get_state(["joint_pos"])
            dq.append(float(np.median(np.abs(g["joint_pos"] - o["joint_pos"]).max(1))))
    g = _np(sim.get_state(PHYS))
    assert all(np.isfinite(v).all() for v in g.values())
    assert sim.log_host()[LOG_NAN_RESETS] == 0  # no non-finite force-resets
    print("median max-joint divergence after 1/5/25/100 steps:", dq, " mean reward/step gpu", rg_sum / 1000, "oracle", ro_sum / 1000)
    assert dq[0] < 1e-5 and dq[1] < 1e-4
    assert abs(rg_sum - ro_sum) / 1000 < 0.05 * max(abs(ro_sum) / 1000, 0.01) + 0.02


def test_flight_conserves_momentum_and_matches_oracle(cfg):
    """No contacts, no dissipation: the horizontal momentum is conserved and float tracks double to 1e-6."""
    c = cfg.copy()
    for d in range(18):
        c.dof_damping[d] = 0.0
        c.dof_frictionloss[d] = 0.0
    c.init_root_height = 20.0
    c.decimation = 1
    c.max_delay = 0
    n = 256
    torch, sim, orc = _mk(c, n, 5)
    sim.observe(); orc.observe()
    rng = np.random.default_rng(0)
    st = _np(sim.get_state(PHYS))
    st["root_lin_vel"] = rng.normal(size=(n, 3)).astype(np.float32)
    st["root_ang_vel"] = rng.normal(size=(n, 3)).astype(np.float32)
    st["joint_vel"] = rng.normal(size=(n, 12)).astype(np.float32) * 2
    sim.set_state(st); orc.set_state(st)
    for step in range(10):
        a = (rng.normal(size=(n, 12)) * 0.2).astype(np.float32)
        sim.step(torch.from_numpy(a).cuda()); orc.step(a)
        g, o = _np(sim.get_state(SYNC)), orc.get_state(PHYS)
        for k, tol in (("joint_pos", 2e-6), ("joint_vel", 2e-4), ("root_lin_vel", 2e-5), ("root_ang_vel", 1e-4)):
            assert np.abs(g[k] - o[k]).max() < tol, (k, step, np.abs(g[k] - o[k]).max())
        _resync(sim, orc, g)


@pytest.mark.parametrize("task,n", [("flat", 4096), ("flat", 32768), ("rsl", 32768)])
def test_full_size_properties(cfg, task, n):
    """Size-independent properties at BASELINE's env counts: determinism (same seed -> bit-identical run), finite
    outputs, history shift property of the observation, episode counters, sharding invariance of the Philox key."""
    import torch
    from h1v2_isaac_b200._capi import rsl_config
    from h1v2_isaac_b200.backend import H1v2Sim
    if task == "rsl":  # history 6, per-env friction, pushes, dead-zone command class
        cfg = rsl_config()
    H = cfg.history_length

    def run(offset, count, steps=12, plain=0):
        c = cfg.copy()
        c.env_id_offset = offset
        c.reserved[3] = plain
        sim = H1v2Sim(count, c, device="cuda:0", seed=123)
        obs = [sim.observe().clone()]
        rews = []
        for i in range(steps):
            act = sim.random_actions(i)
            o, r, t, u = sim.step(act)
            obs.append(o.clone()); rews.append(r.clone())
        ep = sim.episode_length_buf.clone()
        lg = sim.log_host()
        sim.close()
        return obs, rews, ep, lg

    obs_a, rew_a, ep_a, lg = run(0, n)
    obs_b, rew_b, ep_b, _ = run(0, n)
    assert all(torch.equal(x, y) for x, y in zip(obs_a, obs_b)) and all(torch.equal(x, y) for x, y in zip(rew_a, rew_b))
    assert all(torch.isfinite(x).all() for x in obs_a) and all(torch.isfinite(x).all() for x in rew_a)
    assert lg[LOG_NAN_RESETS] == 0
    # history shift: for envs that did not reset, block[h] at step t+1 equals block[h+1] at step t
    off = [0, 3, 6, 9, 21, 33, 45]
    alive = ep_a > 1
    o0, o1 = obs_a[-2][alive], obs_a[-1][alive]
    for t in range(6):
        d = off[t + 1] - off[t]
        b0 = o0[:, off[t] * H: off[t + 1] * H].reshape(-1, H, d)
        b1 = o1[:, off[t] * H: off[t + 1] * H].reshape(-1, H, d)
        assert torch.equal(b1[:, :-1], b0[:, 1:])
    assert int(ep_a.max()) <= 12
    # sharding: the second half of the envs simulated on its own (rank 1 of 2) reproduces the same trajectories
    # (bit for bit between handles on the same kernel instantiation: <= 5624 envs run the mirror-lane ones, whose row sums are
    # associated differently -- for that size the whole job is re-run on the plain instantiation, cfg.reserved[3] = 1)
    half = n // 2
    obs_w = obs_a if n != 4096 else run(0, n, steps=3, plain=1)[0]
    obs_h, rew_h, _, _ = run(half, half, steps=3, plain=1)
    assert torch.equal(obs_h[0], obs_w[0][half:])
    # actions are drawn from the global env id as well, so the whole rollout is identical
    for k in range(1, 4):
        assert torch.equal(obs_h[k], obs_w[k][half:]), k


@pytest.mark.parametrize("mode", ["rows", "assemble", "hybrid", "auto"])
@pytest.mark.parametrize("pinned", [True, False])
def test_step_host_matches_device_path(cfg, pinned, mode, monkeypatch):
    """The HOST-buffer entry point (what a non-torch caller binds) gives exactly the device path's results in both of its
    modes (and while it switches between them to pick one, "auto") -- "rows": the kernel writes whole observation rows into the caller's buffer (zero-copy over PCIe when pinned, staged
    when pageable); "assemble": only the new 45-float sample crosses PCIe and host threads assemble the term-major rows from
    a host mirror of the history ring -- and when it is mixed with the stream-taking entry points (a device-path step, an API
    reset on the caller's stream, a state write): the host path orders itself after them and re-fetches the ring."""
    import torch
    from h1v2_isaac_b200.backend import H1v2Sim
    if mode == "auto":  # the handle times five candidates over its first forty calls and keeps the fastest: results must not show it
        monkeypatch.delenv("H1V2_HOST_PATH", raising=False)
    elif mode == "hybrid":  # the kernel writes the rows of the first 3/8 of the envs, host threads assemble the rest
        monkeypatch.setenv("H1V2_HOST_PATH", "hybrid"); monkeypatch.setenv("H1V2_HOST_ROWS_FRAC", "0.375")
    else:
        monkeypatch.setenv("H1V2_HOST_PATH", mode)
    n = 1024
    cfg = cfg.copy()
    # 8 envs per warp = the mirror-lane instantiation 4096 envs run (lanes 16..31 store nothing, raise no flag twice); 16 = the plain one
    cfg.reserved[2] = 8 if mode in ("assemble", "hybrid") else 16
    s1, s2 = H1v2Sim(n, cfg, seed=4), H1v2Sim(n, cfg, seed=4)
    s1.observe(); s2.observe()
    pin = (lambda x: x.pin_memory()) if pinned else (lambda x: x)
    hobs = pin(torch.empty((n, s1.obs_dim))); hrew = pin(torch.empty(n))
    ht = pin(torch.empty(n, dtype=torch.uint8)); hu = pin(torch.empty(n, dtype=torch.uint8))
    ids = torch.tensor([1, 17, 500, 1023], device="cuda")
    for i in range(46):
        a = s1.random_actions(i)
        o, r, t, u = s1.step(a)
        if i == 5:  # a device-path step in between: the host mirror of the ring is stale afterwards
            o2, r2, t2, u2 = s2.step(a)
            assert torch.equal(o, o2)
            continue
        s2.step_host(pin(a.cpu()), hobs, hrew, ht, hu)
        assert torch.equal(o.cpu(), hobs), i
        assert torch.equal(r.cpu(), hrew) and torch.equal(t.cpu().to(torch.uint8), ht) and torch.equal(u.cpu().to(torch.uint8), hu)
        if i == 8:  # API reset of a few envs, queued on the caller's stream and NOT synchronised before the next host step
            s1.reset(ids); s2.reset(ids)
        if i == 10:  # state write (incl. the observation history) on the caller's stream
            st = {k: v.clone() for k, v in s1.get_state(["obs_history", "joint_pos"]).items()}
            st["obs_history"][::3] += 0.125
            s1.set_state(st); s2.set_state(st)
    if mode == "hybrid" and pinned:
        assert s2.host_path_info()[0] == 1 and s2.host_path_rows() == 384
    s1.close(); s2.close()


def test_step_host_large_batch_matches_device_path(cfg, monkeypatch):
    """16384 envs: the warp-by-warp hand-over (per-warp flags in mapped host memory) at 1024 warps, rows larger than the host's caches.
    Bit-identical to the device path."""
    import torch
    from h1v2_isaac_b200.backend import H1v2Sim
    monkeypatch.setenv("H1V2_HOST_PATH", "assemble")
    n = 16384
    s1, s2 = H1v2Sim(n, cfg, seed=6), H1v2Sim(n, cfg, seed=6)
    s1.observe(); s2.observe()
    hobs = torch.empty((n, s1.obs_dim)).pin_memory(); hrew = torch.empty(n).pin_memory()
    ht = torch.empty(n, dtype=torch.uint8).pin_memory(); hu = torch.empty(n, dtype=torch.uint8).pin_memory()
    for i in range(12):
        a = s1.random_actions(i)
        o, r, t, u = s1.step(a)
        s2.step_host(a.cpu().pin_memory(), hobs, hrew, ht, hu)
        assert torch.equal(o.cpu(), hobs), i
        assert torch.equal(r.cpu(), hrew) and torch.equal(t.cpu().to(torch.uint8), ht) and torch.equal(u.cpu().to(torch.uint8), hu)
    # the device-side bookkeeping of the last host step is complete too (log vector, counters)
    assert torch.equal(torch.from_numpy(s1.log_host()), torch.from_numpy(s2.log_host()))
    s1.close(); s2.close()


def _randomised(cfg, mu_lo=0.3):
    """BASELINE configs[4]: push events, base-mass and friction randomisation on (V/velocity_env_cfg.py:153-173,212-217;
    friction range of C12/rsl_env_cfg.py:213-223); the push interval is shortened so that pushes fire inside the test.
    The strict physics bounds are asserted for mu >= 0.3; below that the fp32 primal Newton loses convergence in a small
    fraction of solves (MuJoCo's pyramidal regulariser makes the normal contact stiffness grow as 1/mu^2: 64x at mu = 0.1,
    which puts stiffness/inertia past 1/eps of fp32) -- measured and bounded in test_low_friction_is_bounded."""
    c = cfg.copy()
    c.push_enable = 1
    c.push_interval_s[0], c.push_interval_s[1] = 0.06, 0.3
    c.push_vel_xy[0], c.push_vel_xy[1] = -0.5, 0.5
    c.mass_add_range[0], c.mass_add_range[1] = -5.0, 5.0
    c.friction_range[0], c.friction_range[1] = mu_lo, 1.25
    if mu_lo < 0.3:
        c.solver_iterations = 30  # what env.flatten_cfg selects for such a friction range
    return c


def test_randomised_events_parity(cfg):
    """Startup randomisation (per-env friction and base mass), interval pushes: draws bit-identical to the oracle's,
    physics from identical states within the same bounds as the nominal task, push velocities / timers on identical
    post-physics states to 1e-6."""
    c = _randomised(cfg)
    n = 2048
    torch, sim, orc = _mk(c, n, 17)
    g, o = _np(sim.get_state(SYNC)), orc.get_state(SYNC)
    for k in ("friction", "mass_add", "push_time_left"):
        np.testing.assert_allclose(g[k], o[k], rtol=0, atol=1e-6, err_msg=k)
    assert g["friction"].min() < 0.35 and g["friction"].max() > 1.0 and g["mass_add"].min() < -4 and g["mass_add"].max() > 4
    sim.observe(); orc.observe()
    rng = np.random.default_rng(5)
    errs = {k: [] for k in PHYS}
    n_push = 0
    for step in range(16):
        a = rng.normal(size=(n, 12)).astype(np.float32)
        # (1) physics with randomised mass / friction from identical states
        _, _, tg, ug = sim.step(torch.from_numpy(a).cuda())
        g = _np(sim.get_state(SYNC + POST))
        before = orc.get_state(["push_time_left"])["push_time_left"][:, 0]
        # (2) the tail on identical post-physics states: push draw, push timer, reset
        _, _, to, uo = orc.step_injected(a, g)
        o = orc.get_state(SYNC)
        assert np.array_equal(tg.cpu().numpy(), to) and np.array_equal(ug.cpu().numpy(), uo)
        for k in ("root_lin_vel", "push_time_left", "friction", "mass_add", "root_pos", "joint_pos"):
            np.testing.assert_allclose(g[k], o[k], rtol=1e-5, atol=2e-6, err_msg=f"{k}, step {step}")
        fired = (o["push_time_left"][:, 0] > before) & ~(to | uo)
        n_push += int(fired.sum())
        _resync(sim, orc, g)
    assert n_push > 500, n_push
    # free physics comparison (oracle runs its own substeps) on the randomised model
    for step in range(12):
        a = rng.normal(size=(n, 12)).astype(np.float32)
        _, _, tg, ug = sim.step(torch.from_numpy(a).cuda())
        _, _, to, uo = orc.step(a)
        g, o = _np(sim.get_state(SYNC)), orc.get_state(PHYS + ["push_time_left"])
        mc, ml = orc.activation_margin()
        pushed = np.abs(g["push_time_left"][:, 0] - o["push_time_left"][:, 0]) > 1e-6
        keep = ~(to | uo | tg.cpu().numpy()) & (mc > 2e-6) & (ml > 2e-6) & ~pushed
        for k in PHYS:
            errs[k].append(np.abs(g[k][keep] - o[k][keep]).max(axis=1))
        _resync(sim, orc, g)
    e = {k: np.concatenate(v) for k, v in errs.items()}
    print({k: (float(v.max()), float(np.quantile(v, 0.999))) for k, v in e.items()})
    # same quantile bounds as the nominal task; the absolute worst case is not asserted at 5e-3 here because the low end
    # of the friction range (0.3..0.5) already shows ~1e-5 of solves that stop at the Newton cap (tools/diag_rand2.py)
    for k in ("joint_pos", "root_pos", "root_quat"):
        assert np.quantile(e[k], 0.9995) < 1e-4 and e[k].max() < 1e-2, k
    for k in ("joint_vel", "root_lin_vel", "root_ang_vel"):
        assert np.quantile(e[k], 0.99) < 5e-4 and np.quantile(e[k], 0.999) < 1.5e-3 and float((e[k] > 5e-3).mean()) < 3e-4, k


def test_low_friction_is_bounded(cfg):
    """Known fp32 limit: at mu in [0.1, 0.3) a small fraction of contact solves does not converge
    within the Newton cap.  This pins how small: >= 99.5 % of env-steps stay within 5e-3 rad/s of the oracle at mu ~ 0.1,
    every state stays finite, and nothing is force-reset by the runaway guard."""
    c = _randomised(cfg, mu_lo=0.1)
    c.friction_range[1] = 0.12
    c.push_enable = 0
    n = 4096
    torch, sim, orc = _mk(c, n, 17)
    sim.observe(); orc.observe()
    rng = np.random.default_rng(5)
    errs = []
    for step in range(20):
        a = rng.normal(size=(n, 12)).astype(np.float32)
        _, _, tg, ug = sim.step(torch.from_numpy(a).cuda())
        _, _, to, uo = orc.step(a)
        g, o = _np(sim.get_state(SYNC)), orc.get_state(PHYS)
        mc, ml = orc.activation_margin()
        keep = ~(to | uo | tg.cpu().numpy()) & (mc > 2e-6) & (ml > 2e-6)
        errs.append(np.abs(g["joint_vel"][keep] - o["joint_vel"][keep]).max(axis=1))
        assert all(np.isfinite(v).all() for v in g.values() if v.dtype.kind == "f")
        _resync(sim, orc, g)
    e = np.concatenate(errs)
    frac = float((e > 5e-3).mean())
    print(f"mu~0.1: {len(e)} env-steps, {100 * frac:.3f} % beyond 5e-3 rad/s, q99 {np.quantile(e, 0.99):.2e}")
    assert frac < 1e-3 and np.quantile(e, 0.99) < 5e-3  # measured 2e-4 with the 30-iteration cap (1.8e-3 with 12)
    assert sim.log_host()[LOG_NAN_RESETS] == 0
