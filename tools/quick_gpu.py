import sys, time
sys.path.insert(0, '/root/repo')
import numpy as np, torch
from h1v2_isaac_b200.backend import H1v2Sim
from h1v2_isaac_b200._capi import default_config
cfg = default_config()
for n in (4096, 32768):
    sim = H1v2Sim(n, cfg, seed=1)
    sim.observe()
    acts = [sim.random_actions(i) for i in range(8)]
    obs = torch.empty((n, sim.obs_dim), device='cuda'); rew = torch.empty(n, device='cuda')
    term = torch.empty(n, dtype=torch.uint8, device='cuda'); trunc = torch.empty(n, dtype=torch.uint8, device='cuda')
    for i in range(20): sim.step_into(acts[i % 8], obs, rew, term, trunc)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    K = 200
    e0.record()
    for i in range(K): sim.step_into(acts[i % 8], obs, rew, term, trunc)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / K
    print(f"n={n}: {ms:.4f} ms/step -> {n / ms * 1e3 / 1e6:.2f} M env-steps/s", "log", sim.log_host()[[0, 24, 27, 28, 29]], "rew mean", rew.mean().item())
    hst = sim.iter_hist(); print("  iteration histogram (fraction):", (hst / hst.sum()).round(4)[:14])
    sim.close()
