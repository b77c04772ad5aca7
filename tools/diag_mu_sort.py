"""Rsl id (per-env friction 0.1..1.25, fixed at startup): would grouping a warp's 16 envs by friction shorten the warps?
Proxy as in diag_predict.py: per-step iteration SUM per env (diag), grouped by 16; cost of a group = its max."""
import sys
sys.path.insert(0, '/root/repo')
import numpy as np, torch
from h1v2_isaac_b200.backend import H1v2Sim
from h1v2_isaac_b200._capi import rsl_config
n = 32768
sim = H1v2Sim(n, rsl_config(), seed=1, diagnostics=True); sim.observe()
mu = sim.get_state(["friction"])["friction"].cpu().numpy()[:, 0]
acts = [sim.random_actions(i) for i in range(8)]
for i in range(60): sim.step(acts[i % 8])
its = []
for i in range(40):
    sim.step(acts[i % 8])
    its.append(sim.get_state(["solver_iters"])["solver_iters"].cpu().numpy()[:, 1])
its = np.stack(its)  # [T, n]
def cost(order): return its[:, order].reshape(its.shape[0], -1, 16).max(2).mean()
order_mu = np.argsort(mu, kind="stable")
print("mean iteration sum per env-step", its.mean(), " unsorted group max", cost(np.arange(n)), " sorted by friction", cost(order_mu),
      " ideal (sorted by the step's own count)", np.mean([np.sort(r).reshape(-1, 16).max(1).mean() for r in its]))
dec = np.quantile(mu, np.linspace(0, 1, 11))
for a, b in zip(dec[:-1], dec[1:]):
    m = (mu >= a) & (mu <= b)
    print(f"mu {a:.2f}..{b:.2f}: mean iteration sum {its[:, m].mean():.2f}  p99 {np.quantile(its[:, m], 0.99):.0f}")
