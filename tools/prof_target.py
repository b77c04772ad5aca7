"""Profiling target: a handful of control steps at a fixed env count (used under ncu on the GPU box)."""
import sys
sys.path.insert(0, '/root/repo')
import torch
from h1v2_isaac_b200.backend import H1v2Sim
from h1v2_isaac_b200._capi import default_config, rough_config, rsl_config
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 12
task = sys.argv[3] if len(sys.argv) > 3 else "flat"
sim = H1v2Sim(n, {"rsl": rsl_config, "rough": rough_config}.get(task, default_config)(), seed=1)
sim.observe()
obs = torch.empty((n, sim.obs_dim), device='cuda'); rew = torch.empty(n, device='cuda')
term = torch.empty(n, dtype=torch.uint8, device='cuda'); trunc = torch.empty(n, dtype=torch.uint8, device='cuda')
acts = [sim.random_actions(i) for i in range(4)]
flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda') if len(sys.argv) > 4 and sys.argv[4] == "flush" else None  # > 126 MB L2, like bench.py
for i in range(steps):
    if flush is not None:
        flush.zero_()
    sim.step_into(acts[i % 4], obs, rew, term, trunc)
torch.cuda.synchronize()
print("done", float(rew.mean()))
