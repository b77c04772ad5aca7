"""Step time vs the solver's velocity tolerance (solver_vel_tolerance) and gradient tolerance; no L2 flush, relative numbers."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from h1v2_isaac_b200._capi import default_config
from h1v2_isaac_b200.backend import H1v2Sim
def t(n, **kw):
    cfg = default_config()
    for k, v in kw.items(): setattr(cfg, k, v)
    sim = H1v2Sim(n, cfg, seed=1); sim.observe()
    acts = [sim.random_actions(i) for i in range(8)]
    obs = torch.empty((n, sim.obs_dim), device='cuda'); rew = torch.empty(n, device='cuda')
    term = torch.empty(n, dtype=torch.uint8, device='cuda'); trunc = torch.empty(n, dtype=torch.uint8, device='cuda')
    for i in range(40): sim.step_into(acts[i % 8], obs, rew, term, trunc)
    torch.cuda.synchronize()
    best = 1e9
    for rep in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(100): sim.step_into(acts[i % 8], obs, rew, term, trunc)
        e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1) / 100)
    lg = sim.log_host(); sim.close()
    return f"{best:.4f} ms ({n / best / 1e3:.1f} M/s, iters/substep {lg[30] / (4 * n):.2f}, max {lg[28]:.0f}, cap hits {lg[29]:.0f})"
for kw in ({}, {"solver_vel_tolerance": 5e-4}, {"solver_vel_tolerance": 3e-4}, {"solver_vel_tolerance": 2e-4}, {"solver_vel_tolerance": 1e-4}, {"solver_vel_tolerance": 5e-5},
           {"solver_tolerance": 3e-6}, {"solver_tolerance": 1e-6}):
    print(kw, "| 4096:", t(4096, **kw), "| 32768:", t(32768, **kw), flush=True)
