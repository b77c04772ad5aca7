"""A/B of the mirror-lane (QUAD) instantiation at 8 envs per warp: cfg.reserved[3] = 1 selects the plain kernel.
Prints ms per step (warm L2, best of 3 x 100 steps) and the difference of the two after one and after 50 identical steps."""
import sys
sys.path.insert(0, '/root/repo')
import torch
from h1v2_isaac_b200.backend import H1v2Sim
from h1v2_isaac_b200._capi import default_config
ns = [int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "4096").split(",")]
for n in ns:
    res = {}
    for plain in (1, 0):
        cfg = default_config(); cfg.reserved[3] = plain
        if len(sys.argv) > 2: cfg.reserved[2] = int(sys.argv[2])
        sim = H1v2Sim(n, cfg, seed=1); sim.observe()
        acts = [sim.random_actions(i) for i in range(8)]
        obs = torch.empty((n, sim.obs_dim), device='cuda'); rew = torch.empty(n, device='cuda')
        term = torch.empty(n, dtype=torch.uint8, device='cuda'); trunc = torch.empty(n, dtype=torch.uint8, device='cuda')
        sim.step_into(acts[0], obs, rew, term, trunc); o1 = obs.clone(); r1 = rew.clone()
        for i in range(1, 50): sim.step_into(acts[i % 8], obs, rew, term, trunc)
        o50 = obs.clone(); t50 = term.clone()
        torch.cuda.synchronize()
        best = 1e9
        for rep in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(100): sim.step_into(acts[i % 8], obs, rew, term, trunc)
            e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1) / 100)
        lg = sim.log_host()
        res[plain] = (best, o1, r1, o50, t50, float(rew.mean()), lg[30] / (4 * n))
        sim.close()
    a, b = res[1], res[0]
    print(f"n={n}: plain {a[0]:.4f} ms ({n / a[0] / 1e3:.2f} M/s) | mirror-lane {b[0]:.4f} ms ({n / b[0] / 1e3:.2f} M/s) | "
          f"step 1: max |d obs| {float((a[1] - b[1]).abs().max()):.2e}, max |d rew| {float((a[2] - b[2]).abs().max()):.2e}; "
          f"step 50: q99.9 |d obs| {float(torch.quantile((a[3] - b[3]).abs().max(dim=1).values, 0.999)):.2e}, term differs in {int((a[4] != b[4]).sum())} envs; "
          f"rew mean {a[5]:.5f} / {b[5]:.5f}, iters/substep {a[6]:.3f} / {b[6]:.3f}", flush=True)
