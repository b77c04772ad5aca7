import sys
sys.path.insert(0, '/root/repo')
import numpy as np, torch
from h1v2_isaac_b200.backend import H1v2Sim
from h1v2_isaac_b200._capi import default_config
n = 32768
for cap in (12, 6, 4, 3, 2):
    cfg = default_config(); cfg.reserved[1] = cap
    sim = H1v2Sim(n, cfg, seed=1); sim.observe()
    acts = [sim.random_actions(i) for i in range(8)]
    obs = torch.empty((n, sim.obs_dim), device='cuda'); rew = torch.empty(n, device='cuda')
    term = torch.empty(n, dtype=torch.uint8, device='cuda'); trunc = torch.empty(n, dtype=torch.uint8, device='cuda')
    for i in range(60): sim.step_into(acts[i % 8], obs, rew, term, trunc)
    h0 = sim.iter_hist().copy()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(100): sim.step_into(acts[i % 8], obs, rew, term, trunc)
    e1.record(); torch.cuda.synchronize()
    h = sim.iter_hist() - h0; frac = h / h.sum(); mean = (frac * np.arange(32)).sum()
    cdf = np.cumsum(frac); emax = ((cdf ** 16)[1:] - (cdf ** 16)[:-1]) @ np.arange(1, 32)
    print(f"ls cap {cap:2d}: {e0.elapsed_time(e1)/100:.4f} ms/step  mean iters {mean:.3f}  E[max16] {emax:.2f}  cap hits {sim.log_host()[29]:.0f}  hist {frac[:9].round(3)}")
    sim.close()
