"""Speed / accuracy of a FIXED Newton-iteration budget (solver_iterations = 2, 3, 4, 6 instead of the default 12): batched-GPU
MuJoCo ports are commonly run that way.  For every budget: throughput at 32768 envs, share of solves cut off by the budget, and
the one-control-step deviation from the float64 oracle (converged solve) on identical states, 2048 envs x 24 steps."""
import sys
sys.path.insert(0, '/root/repo')
import numpy as np, torch
from h1v2_isaac_b200.backend import H1v2Sim
from h1v2_isaac_b200._capi import default_config, LOG_CAP_HITS
from oracle.oracle import Oracle
PHYS = ["root_pos", "root_quat", "root_lin_vel", "root_ang_vel", "joint_pos", "joint_vel"]
SYNC = PHYS + ["last_action", "target_hist", "lag", "fresh", "command", "heading_target", "time_left", "is_standing", "is_heading", "cmd_metrics",
               "feet_timers", "episode_sums", "obs_history", "friction", "mass_add", "push_time_left"]
for cap in (12, 6, 4, 3, 2):
    cfg = default_config(); cfg.solver_iterations = cap
    n = 32768
    sim = H1v2Sim(n, cfg, seed=1); sim.observe()
    acts = [sim.random_actions(i) for i in range(8)]
    obs = torch.empty((n, sim.obs_dim), device='cuda'); rew = torch.empty(n, device='cuda')
    term = torch.empty(n, dtype=torch.uint8, device='cuda'); trunc = torch.empty(n, dtype=torch.uint8, device='cuda')
    for i in range(30): sim.step_into(acts[i % 8], obs, rew, term, trunc)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    caps = torch.zeros((), device='cuda')
    e0.record()
    for i in range(200): sim.step_into(acts[i % 8], obs, rew, term, trunc)
    e1.record(); torch.cuda.synchronize()
    for i in range(50):
        sim.step_into(acts[i % 8], obs, rew, term, trunc); caps += sim.log_buf[LOG_CAP_HITS]
    ms = e0.elapsed_time(e1) / 200
    cut = float(caps) / (50 * n * 4)
    sim.close()
    # deviation from the converged float64 solve, one control step from identical states
    m = 2048
    ocfg = default_config()  # the oracle always converges (100 iterations, 1e-10)
    sim = H1v2Sim(m, cfg, seed=3, diagnostics=True); orc = Oracle(ocfg, m, seed=3, threads=16)
    sim.observe(); orc.observe()
    rng = np.random.default_rng(0); errs = {k: [] for k in PHYS}
    for step in range(24):
        a = rng.normal(size=(m, 12)).astype(np.float32)
        _, _, tg, ug = sim.step(torch.from_numpy(a).cuda())
        _, _, to, uo = orc.step(a)
        g = {k: v.cpu().numpy() for k, v in sim.get_state(SYNC).items()}; o = orc.get_state(PHYS)
        mc, ml = orc.activation_margin()
        keep = ~(to | uo | tg.cpu().numpy().astype(bool)) & (mc > 2e-6) & (ml > 2e-6)
        for k in PHYS: errs[k].append(np.abs(g[k][keep] - o[k][keep]).max(axis=1))
        orc.set_state({k: g[k] for k in SYNC}); orc.episode_length = sim.episode_length_buf.cpu().numpy()
    sim.close()
    e = {k: np.concatenate(v) for k, v in errs.items()}
    q = lambda k: "%.1e / %.1e / %.1e" % (np.quantile(e[k], 0.99), np.quantile(e[k], 0.999), e[k].max())
    print(f"budget {cap:2d}: {ms:.4f} ms/step = {n / ms / 1e3:.1f} M env-steps/s at 32768 envs; solves cut off {100 * cut:.2f} %; "
          f"|d joint_vel| 99% / 99.9% / max {q('joint_vel')} rad/s; |d joint_pos| {q('joint_pos')} rad", flush=True)
