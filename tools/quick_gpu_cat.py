"""Throughput of the Constraints-as-Terminations step (h1v2_cat_step: fused step + 3 tail launches) at 4096 / 32768 envs."""
import sys
sys.path.insert(0, '/root/repo')
import torch
from h1v2_isaac_b200 import tasks
from h1v2_isaac_b200.backend import H1v2Sim
for n in (4096, 32768):
    sim = H1v2Sim(n, tasks.cat_config(), seed=1); sim.observe()
    acts = [sim.random_actions(i) for i in range(8)]
    for i in range(20): sim.cat_step(acts[i % 8])
    torch.cuda.synchronize()
    l0 = sim.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    K = 200
    e0.record()
    for i in range(K): sim.cat_step(acts[i % 8])
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / K
    print(f"CaT n={n}: {ms:.4f} ms/step -> {n / ms * 1e3 / 1e6:.2f} M env-steps/s, {(sim.launch_count - l0) / K:.0f} launches per step")
    sim.close()
