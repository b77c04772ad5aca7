"""Long-run stability: 60 000 control steps (20 simulated minutes per env) of 4096 envs under N(0,1) actions and, for the second
half, a 3x larger action scale; every output checked for finiteness on the device, solver statistics from the log vector."""
import sys, time
sys.path.insert(0, '/root/repo')
import torch
from h1v2_isaac_b200.backend import H1v2Sim
from h1v2_isaac_b200._capi import default_config, rough_config, rsl_config
n, steps = 4096, 60000
task = sys.argv[1] if len(sys.argv) > 1 else "flat"  # "rsl": the Rsl id (friction 0.1..1.25, pushes, dead-zone commands, scale 0.25)
sim = H1v2Sim(n, {"rsl": rsl_config, "rough": rough_config}.get(task, default_config)(), seed=123); sim.observe()  # "rough": height field, scan, terrain curriculum
print("task", task)
pool = [sim.random_actions(i) for i in range(64)]
obs = torch.empty((n, sim.obs_dim), device='cuda'); rew = torch.empty(n, device='cuda')
term = torch.empty(n, dtype=torch.uint8, device='cuda'); trunc = torch.empty(n, dtype=torch.uint8, device='cuda')
bad = torch.zeros((), device='cuda'); maxabs = torch.zeros((), device='cuda'); nterm = torch.zeros((), device='cuda'); ntrunc = torch.zeros((), device='cuda')
capsum = torch.zeros((), device='cuda'); ovf = torch.zeros((), device='cuda')
t0 = time.time()
half = {}
for i in range(steps):
    if i == steps // 2:
        torch.cuda.synchronize()
        half = dict(bad=int(bad), nterm=int(nterm), resets=float(sim.log_host()[27]), cap=int(capsum), ovf=int(ovf))
    a = pool[i % 64] * (3.0 if i >= steps // 2 else 1.0)
    sim.step_into(a, obs, rew, term, trunc)
    bad += (~torch.isfinite(obs)).sum() + (~torch.isfinite(rew)).sum()
    maxabs = torch.maximum(maxabs, obs.abs().max())
    nterm += term.sum(); ntrunc += trunc.sum()
    capsum += sim.log_buf[29]; ovf += sim.log_buf[31]
torch.cuda.synchronize()
lg = sim.log_host()
print(f"{steps} steps x {n} envs = {steps*n/1e6:.0f} M env-steps in {time.time()-t0:.1f} s (with per-step finiteness checks)")
print(f"first half (N(0,1) actions): non-finite {half['bad']}, terminations {half['nterm']}, runaway resets {half['resets']:.0f}, Newton-cap hits {half['cap']} of {steps*n*2} solves, overflows {half['ovf']}")
print(f"whole run (second half: 3x action scale) --")
print(f"non-finite outputs: {int(bad)}  max |obs|: {float(maxabs):.1f}  terminations: {int(nterm)}  time-outs: {int(ntrunc)}")
if task == "rough":
    lv = sim.get_state(["terrain_level"])["terrain_level"].float()
    print(f"terrain levels after the run: mean {float(lv.mean()):.2f} (log {float(sim.terrain_log_buf[1]):.2f}), min {int(lv.min())}, max {int(lv.max())}; guard bytes overwritten: {sim.check_guards()}")
print(f"runaway/non-finite force-resets (cumulative): {lg[27]:.0f}  Newton-cap hits: {int(capsum)} of {steps*n*4} solves  contact-list overflows: {int(ovf)}")
