import sys
sys.path.insert(0, '/root/repo')
import torch
from h1v2_isaac_b200.backend import H1v2Sim
from h1v2_isaac_b200._capi import default_config
n = 4096
sim = H1v2Sim(n, default_config(), seed=42); sim.observe()
g = torch.Generator(device='cuda').manual_seed(0)
mx = 0.0; bad = 0; rmin = 0.0
for i in range(600):
    scale = 1.0 if i < 300 else 5.0
    a = scale * torch.randn((n, 12), device='cuda', generator=g)
    if i % 50 == 49: a[:64] = float('1e6')
    o, r, t, u = sim.step(a)
    bad += int((~torch.isfinite(o)).sum()) + int((~torch.isfinite(r)).sum())
    mx = max(mx, float(o[torch.isfinite(o)].abs().max())); rmin = min(rmin, float(r.min()))
    if i % 100 == 99: print(i, "non-finite so far", bad, "max |obs|", mx, "min rew", rmin, "nan resets", sim.log_host()[27])
