import sys
sys.path.insert(0, '/root/repo')
import torch
from h1v2_isaac_b200.backend import H1v2Sim
from h1v2_isaac_b200._capi import default_config, rsl_config
"""Step time vs envs per warp (reserved[2] override).  usage: python tools/diag_epw.py [flat|rsl]"""
task = sys.argv[1] if len(sys.argv) > 1 else "flat"
default_config = rsl_config if task == "rsl" else default_config
print("task", task)
for n in (1024, 2048, 4096, 8192, 16384, 32768):
    row = []
    for epw in (16, 8, 4, 2):
        cfg = default_config(); cfg.reserved[2] = epw
        sim = H1v2Sim(n, cfg, seed=1); sim.observe()
        acts = [sim.random_actions(i) for i in range(8)]
        obs = torch.empty((n, sim.obs_dim), device='cuda'); rew = torch.empty(n, device='cuda')
        term = torch.empty(n, dtype=torch.uint8, device='cuda'); trunc = torch.empty(n, dtype=torch.uint8, device='cuda')
        for i in range(30): sim.step_into(acts[i % 8], obs, rew, term, trunc)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(200): sim.step_into(acts[i % 8], obs, rew, term, trunc)
        e1.record(); torch.cuda.synchronize()
        row.append(f"epw {epw}: {e0.elapsed_time(e1)/200:.4f} ms")
        sim.close()
    print(n, " | ".join(row))
