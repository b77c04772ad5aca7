"""Throughput of the Constraints-as-Terminations step (h1v2_cat_step: fused step + the apply kernel) at 4096 / 32768 envs, beside the
plain step of the same config (cfg.cat_enable kept, h1v2_step called): the difference is the whole cost of the constraint tail."""
import sys
sys.path.insert(0, '/root/repo')
import torch
from h1v2_isaac_b200 import tasks
from h1v2_isaac_b200.backend import H1v2Sim
for n in (4096, 32768):
    for label in ("cat_step", "plain step"):
        sim = H1v2Sim(n, tasks.cat_config(), seed=1); sim.observe()
        acts = [sim.random_actions(i) for i in range(8)]
        obs = torch.empty((n, sim.obs_dim), device='cuda'); rew = torch.empty(n, device='cuda'); d = torch.empty(n, device='cuda')
        t = torch.empty(n, dtype=torch.uint8, device='cuda'); u = torch.empty(n, dtype=torch.uint8, device='cuda')
        step = (lambda a: sim.cat_step_into(a, obs, rew, d, u)) if label == "cat_step" else (lambda a: sim.step_into(a, obs, rew, t, u))
        for i in range(20): step(acts[i % 8])
        torch.cuda.synchronize()
        l0 = sim.launch_count
        best = 1e9
        for rep in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(100): step(acts[i % 8])
            e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1) / 100)
        print(f"CaT cfg n={n} {label}: {best:.4f} ms/step -> {n / best * 1e3 / 1e6:.2f} M env-steps/s, {(sim.launch_count - l0) / 300:.0f} launches per step")
        sim.close()
