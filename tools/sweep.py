"""BASELINE configs[4]: env-count scaling sweep 1k-64k envs on one GPU, nominal Flat task and with push events +
friction / base-mass randomisation enabled (V/velocity_env_cfg.py:153-173,212-217; friction range of C12/rsl_env_cfg.py:213-223).
Same timing rule as bench.py: per-step CUDA events, L2 flushed (untimed) between steps, random N(0,1) actions resident in HBM.
usage (GPU box): python tools/sweep.py > gpurun_out/sweep.json"""
import json, sys
sys.path.insert(0, '/root/repo')
import torch
from h1v2_isaac_b200.backend import H1v2Sim
from h1v2_isaac_b200._capi import default_config

flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
rows = []
for randomize in (False, True):
    for n in (1024, 2048, 4096, 8192, 16384, 32768, 65536):
        cfg = default_config()
        if randomize:
            cfg.push_enable = 1
            cfg.mass_add_range[0], cfg.mass_add_range[1] = -5.0, 5.0
            cfg.friction_range[0], cfg.friction_range[1] = 0.1, 1.25
        sim = H1v2Sim(n, cfg, seed=42); sim.observe()
        acts = [sim.random_actions(i) for i in range(16)]
        obs = torch.empty((n, sim.obs_dim), device='cuda'); rew = torch.empty(n, device='cuda')
        term = torch.empty(n, dtype=torch.uint8, device='cuda'); trunc = torch.empty(n, dtype=torch.uint8, device='cuda')
        for i in range(40): sim.step_into(acts[i % 16], obs, rew, term, trunc)
        torch.cuda.synchronize()
        K = 100
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
        for i in range(K):
            flush.zero_(); ev[i][0].record(); sim.step_into(acts[i % 16], obs, rew, term, trunc); ev[i][1].record()
        torch.cuda.synchronize()
        ms = sum(a.elapsed_time(b) for a, b in ev) / K
        lg = sim.log_host()
        rows.append({"envs": n, "randomised": randomize, "ms_per_step": round(ms, 4), "env_steps_per_s": round(n / ms * 1e3),
                     "mean_newton_iters": round(float(lg[30]) / (4 * n), 3), "nan_resets_total": float(lg[27])})
        print(rows[-1], file=sys.stderr)
        sim.close()
print(json.dumps({"gpu": torch.cuda.get_device_name(0), "rows": rows}, indent=1))
