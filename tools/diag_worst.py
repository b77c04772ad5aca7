import sys
sys.path.insert(0, '/root/repo')
import numpy as np, torch
from h1v2_isaac_b200.backend import H1v2Sim
from h1v2_isaac_b200._capi import default_config
from oracle.oracle import Oracle
PHYS = ["root_pos", "root_quat", "root_lin_vel", "root_ang_vel", "joint_pos", "joint_vel"]
SYNC = PHYS + ["last_action", "target_hist", "lag", "fresh", "command", "heading_target", "time_left", "is_standing",
               "is_heading", "cmd_metrics", "feet_timers", "episode_sums", "obs_history", "friction", "mass_add", "push_time_left"]
n = 2048
for tol, stol, ascale, dec in ((1e-6, 1e-3, 1.0, 1), (1e-6, 1e-3, 0.3, 1), (1e-6, 1e-3, 1.0, 4), (1e-6, 1e-3, 0.3, 4), (1e-5, 1e-2, 0.3, 4)):
    cfg = default_config(); cfg.decimation = dec; cfg.max_delay = min(5, 2 * dec); cfg.solver_tolerance = tol; cfg.solver_step_tolerance = stol; cfg.solver_iterations = 20
    sim = H1v2Sim(n, cfg, seed=3, diagnostics=True); orc = Oracle(cfg, n, seed=3, threads=16)
    sim.observe(); orc.observe()
    rng = np.random.default_rng(0)
    E = []; info = []
    for step in range(96 // dec):
        a = (ascale * rng.normal(size=(n, 12))).astype(np.float32)
        _, _, tg, ug = sim.step(torch.from_numpy(a).cuda()); _, _, to, uo = orc.step(a)
        g = {k: v.cpu().numpy() for k, v in sim.get_state(SYNC + ["solver_iters", "slot_force_hist"]).items()}
        o = orc.get_state(PHYS)
        mc, ml = orc.activation_margin(); oit, ores = orc.solver_stats()
        keep = ~(to | uo | tg.cpu().numpy()) & (mc > 2e-6) & (ml > 2e-6)
        e = np.abs(g["joint_vel"] - o["joint_vel"]); e[~keep] = 0
        E.append(e.max(1)[keep])
        w = np.argmax(e.max(1))
        info.append((e.max(1)[w], int(np.argmax(e[w])), g["solver_iters"][w, 0], oit[w], mc[w], ml[w], g["slot_force_hist"][w].reshape(6, 3)[:, 2].round(1)))
        orc.set_state({k: g[k] for k in SYNC}); orc.episode_length = sim.episode_length_buf.cpu().numpy()
    E = np.concatenate(E)
    lg = sim.log_host()
    print(f"tol={tol:g} step_tol={stol:g} action_scale={ascale} dec={dec}: max {E.max():.2e} p99.99 {np.quantile(E,.9999):.2e} p99.9 {np.quantile(E,.999):.2e} p99 {np.quantile(E,.99):.2e}  mean iters {lg[28]/n/dec:.2f}")
    for r in sorted(info, key=lambda r: -r[0])[:2]:
        print("    err %.2e joint %d gpu_it %d orc_it %d margin %.1e lim %.1e |F| %s" % r)
    sim.close()
