import sys
sys.path.insert(0, '/root/repo')
import numpy as np, torch
from h1v2_isaac_b200.backend import H1v2Sim
from h1v2_isaac_b200._capi import default_config
from oracle.oracle import Oracle
PHYS = ["root_pos", "root_quat", "root_lin_vel", "root_ang_vel", "joint_pos", "joint_vel"]
SYNC = PHYS + ["last_action", "target_hist", "lag", "fresh", "command", "heading_target", "time_left", "is_standing",
               "is_heading", "cmd_metrics", "feet_timers", "episode_sums", "obs_history", "friction", "mass_add", "push_time_left"]
n = 1024
for dec in (4, 1):
    cfg = default_config(); cfg.decimation = dec; cfg.solver_tolerance = 1e-6
    sim = H1v2Sim(n, cfg, seed=3, diagnostics=True); orc = Oracle(cfg, n, seed=3, threads=16)
    sim.observe(); orc.observe()
    rng = np.random.default_rng(0)
    E = []; Q = []; A = []
    for step in range(30 * (4 // dec)):
        a = rng.normal(size=(n, 12)).astype(np.float32)
        _, rg, tg, ug = sim.step(torch.from_numpy(a).cuda()); _, ro, to, uo = orc.step(a)
        g = {k: v.cpu().numpy() for k, v in sim.get_state(SYNC + ["joint_acc"]).items()}
        o = orc.get_state(SYNC + ["joint_acc"])
        keep = ~(to | uo | tg.cpu().numpy())
        E.append(np.abs(g["joint_vel"][keep] - o["joint_vel"][keep])); Q.append(np.abs(o["joint_vel"][keep])); A.append(np.abs(o["joint_acc"][keep]))
        orc.set_state(g); orc.episode_length = sim.episode_length_buf.cpu().numpy()
    E = np.concatenate(E); Q = np.concatenate(Q); A = np.concatenate(A)
    print(f"decimation={dec}: per-joint p99 / p99.9 / max of |dqd| (MJCF order), median |qd|, median |qacc|")
    for j in range(12):
        print(f"  j{j:2d}: {np.quantile(E[:, j], .99):.2e} {np.quantile(E[:, j], .999):.2e} {E[:, j].max():.2e}   |qd| med {np.median(Q[:, j]):.2f} p99 {np.quantile(Q[:, j], .99):.1f}  |qacc| med {np.median(A[:, j]):.0f} p99 {np.quantile(A[:, j], .99):.0f}")
    rel = E / (1e-3 + 0.005 * A)
    print("  error relative to h*|qacc|: p99", np.quantile(rel, .99), "max", rel.max())
    sim.close()
