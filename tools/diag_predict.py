"""How well can a warp's Newton work be predicted at step start?  Expected cost of a 16-env warp = sum over substeps of the
max iteration count; proxy here: per-step iteration SUM per env (diag), grouped by 16 after sorting by a predictor."""
import sys
sys.path.insert(0, '/root/repo')
import numpy as np, torch
from h1v2_isaac_b200.backend import H1v2Sim
from h1v2_isaac_b200._capi import default_config
n = 32768
sim = H1v2Sim(n, default_config(), seed=1, diagnostics=True); sim.observe()
acts = [sim.random_actions(i) for i in range(8)]
for i in range(80): sim.step(acts[i % 8])
prev = None; rows = []
for i in range(20):
    s0 = {k: v.cpu().numpy() for k, v in sim.get_state(["feet_timers", "joint_vel", "root_lin_vel", "root_pos", "root_ang_vel"]).items()}
    sim.step(acts[i % 8])
    it = sim.get_state(["solver_iters"])["solver_iters"].cpu().numpy()[:, 1]
    if prev is not None:
        ncon = (s0["feet_timers"].reshape(n, 2, 4)[:, :, 2] > 0).sum(1)
        feats = {"prev_sum_iters": prev, "feet_in_contact": ncon, "joint_speed": np.linalg.norm(s0["joint_vel"], axis=1),
                 "root_height": -s0["root_pos"][:, 2], "root_speed": np.linalg.norm(s0["root_lin_vel"], axis=1),
                 "combo": prev + 2.0 * ncon}
        def cost(order): return it[order].reshape(-1, 16).max(1).mean()
        r = {"unsorted": cost(np.arange(n)), "ideal": cost(np.argsort(it))}
        for k, f in feats.items(): r[k] = cost(np.argsort(f, kind="stable"))
        r["corr_prev"] = np.corrcoef(prev, it)[0, 1]; r["mean"] = it.mean()
        rows.append(r)
    prev = it.copy()
keys = rows[0].keys()
print({k: round(float(np.mean([r[k] for r in rows])), 3) for k in keys})
