import sys
sys.path.insert(0, '/root/repo')
import numpy as np, torch
from h1v2_isaac_b200.backend import H1v2Sim
from h1v2_isaac_b200._capi import default_config
n = 8192
for name in ("default", "no_frictionloss", "tol1e-4", "steptol0.1"):
    cfg = default_config()
    if name == "no_frictionloss":
        for d in range(18): cfg.dof_frictionloss[d] = 0.0
    if name == "tol1e-4": cfg.solver_tolerance = 1e-4
    if name == "steptol0.1": cfg.solver_step_tolerance = 0.1
    sim = H1v2Sim(n, cfg, seed=1, diagnostics=True); sim.observe()
    acts = [sim.random_actions(i) for i in range(8)]
    prev = None; cors = []
    for i in range(60):
        sim.step(acts[i % 8])
    h0 = sim.iter_hist().copy()
    for i in range(40):
        sim.step(acts[i % 8])
        # per-env max iterations of this step is in the diag buffer slot 86 -> exposed through reward_terms? use slot_force_hist hack: not exposed; skip
    h = sim.iter_hist() - h0
    frac = h / h.sum()
    mean = (frac * np.arange(32)).sum()
    # expected max of 16 iid draws
    cdf = np.cumsum(frac); emax = ((cdf ** 16)[1:] - (cdf ** 16)[:-1]) @ np.arange(1, 32) + (cdf[0] ** 16) * 0
    print(f"{name:16s} mean {mean:.2f}  E[max16 iid] {emax:.2f}  hist {frac[:10].round(3)}")
    sim.close()
