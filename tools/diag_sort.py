import sys
sys.path.insert(0, '/root/repo')
import numpy as np, torch
from h1v2_isaac_b200.backend import H1v2Sim
from h1v2_isaac_b200._capi import default_config
n = 16384
sim = H1v2Sim(n, default_config(), seed=1, diagnostics=True); sim.observe()
acts = [sim.random_actions(i) for i in range(8)]
for i in range(60): sim.step(acts[i % 8])
prev = None; res = []
for i in range(30):
    sim.step(acts[i % 8])
    it = sim.get_state(["solver_iters"])["solver_iters"].cpu().numpy()
    mx, sm = it[:, 0], it[:, 1]
    if prev is not None:
        rho = np.corrcoef(prev, sm)[0, 1]
        # warp cost proxy: per-step max over groups of 16 of the per-env max (upper bound proxy: use sum/4 as per-substep mean)
        def cost(order):
            g = sm[order].reshape(-1, 16)
            return g.max(1).mean()
        unsorted = cost(np.arange(n)); pred = cost(np.argsort(prev, kind="stable")); ideal = cost(np.argsort(sm))
        res.append((rho, sm.mean(), unsorted, pred, ideal))
    prev = sm.copy()
r = np.array(res).mean(0)
print(f"corr(prev,cur)={r[0]:.3f} mean sum-iters={r[1]:.2f}  group-max: unsorted {r[2]:.2f}  sorted-by-prev {r[3]:.2f}  ideal {r[4]:.2f}")
