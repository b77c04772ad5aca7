"""Developer diagnostic (GPU box): single-step parity error distribution vs solver settings."""
import sys, time
sys.path.insert(0, '/root/repo')
import numpy as np, torch
from h1v2_isaac_b200.backend import H1v2Sim
from h1v2_isaac_b200._capi import default_config
from oracle.oracle import Oracle

PHYS = ["root_pos", "root_quat", "root_lin_vel", "root_ang_vel", "joint_pos", "joint_vel"]
SYNC = PHYS + ["last_action", "target_hist", "lag", "fresh", "command", "heading_target", "time_left", "is_standing",
               "is_heading", "cmd_metrics", "feet_timers", "episode_sums", "obs_history", "friction", "mass_add", "push_time_left"]
n = 1024
settings = [(12, 1e-5), (12, 1e-6), (20, 1e-5)] if len(sys.argv) < 2 else [tuple(map(float, a.split(','))) for a in sys.argv[1:]]
for iters, tol in settings:
    cfg = default_config(); cfg.solver_iterations = int(iters); cfg.solver_tolerance = tol
    sim = H1v2Sim(n, cfg, seed=3, diagnostics=True); orc = Oracle(cfg, n, seed=3, threads=16)
    sim.observe(); orc.observe()
    rng = np.random.default_rng(0)
    errs = {k: [] for k in PHYS}; caps = 0; mism = 0; its = []; sumit = 0
    for step in range(30):
        a = rng.normal(size=(n, 12)).astype(np.float32)
        _, rg, tg, ug = sim.step(torch.from_numpy(a).cuda())
        _, ro, to, uo = orc.step(a)
        g = {k: v.cpu().numpy() for k, v in sim.get_state(SYNC + ["slot_force_hist"]).items()}
        o = orc.get_state(SYNC + ["slot_force_hist"])
        tg = tg.cpu().numpy()
        lg = sim.log_host(); caps += lg[27]; its.append(lg[26]); sumit += lg[28]
        mism += int((tg != to).sum())
        keep = ~(to | uo | tg)
        for k in PHYS:
            errs[k].append(np.abs(g[k][keep] - o[k][keep]).max(axis=1))
        orc.set_state(g); orc.episode_length = sim.episode_length_buf.cpu().numpy()
    print(f"iters={iters} tol={tol:g}: cap_hits/solve={caps/(30*4*n):.4f} mean_it={sumit/(30*4*n):.2f} max_it={max(its)} term mismatches={mism}")
    for k in PHYS:
        e = np.concatenate(errs[k])
        print(f"   {k:13s} max {e.max():.2e}  p99.9 {np.quantile(e,0.999):.2e}  p99 {np.quantile(e,0.99):.2e}  median {np.median(e):.2e}")
    sim.close()
