"""Rough id vs Flat id: step time and Newton statistics at 4096 / 32768 envs (warm L2, back-to-back steps)."""
import sys
sys.path.insert(0, '/root/repo')
import torch
from h1v2_isaac_b200.backend import H1v2Sim
from h1v2_isaac_b200._capi import default_config, rough_config, LOG_SUM_ITERS, LOG_MAX_ITERS, LOG_CAP_HITS, LOG_CONTACT_OVERFLOW
variants = {"flat": default_config(), "rough": rough_config()}
r0 = rough_config(); r0.obs_height_scan = 0
variants["rough, no height scan"] = r0
r1 = rough_config(); r1.terrain_enable = 0
variants["rough instantiation on the plane"] = r1
for name, cfg in variants.items():
    for n in (4096, 32768):
        sim = H1v2Sim(n, cfg, seed=1)
        sim.observe()
        acts = [sim.random_actions(i) for i in range(8)]
        obs = torch.empty((n, sim.obs_dim), device='cuda'); rew = torch.empty(n, device='cuda')
        term = torch.empty(n, dtype=torch.uint8, device='cuda'); trunc = torch.empty(n, dtype=torch.uint8, device='cuda')
        for i in range(30): sim.step_into(acts[i % 8], obs, rew, term, trunc)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        K = 200
        e0.record()
        for i in range(K): sim.step_into(acts[i % 8], obs, rew, term, trunc)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / K
        lg = sim.log_host()
        print(f"{name:34s} n={n:6d}: {ms:.4f} ms/step -> {n / ms * 1e3 / 1e6:6.2f} M env-steps/s | iters/substep {lg[LOG_SUM_ITERS] / (4 * n):.2f} max {lg[LOG_MAX_ITERS]:.0f} "
              f"cap hits {lg[LOG_CAP_HITS]:.0f} overflow {lg[LOG_CONTACT_OVERFLOW]:.0f} resets/step {lg[0]:.0f}")
        sim.close()
