"""BASELINE configs[4]: env-count scaling sweep 1k-64k envs PER GPU, nominal Flat task and with push events +
friction / base-mass randomisation enabled (V/velocity_env_cfg.py:153-173,212-217; friction range of C12/rsl_env_cfg.py:213-223).
Same timing rule as bench.py: per-step CUDA events, L2 flushed (untimed) between steps, random N(0,1) actions resident in HBM.
One GPU:   python tools/sweep.py > gpurun_out/sweep.json
N GPUs:    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/sweep.py > gpurun_out/sweep_N.json
(one rank per GPU, envs sharded by rank with the global env id in the Philox key, no data-path collective; every row is timed between
two barriers, the time is the MAX over ranks and the rate the sum of all ranks' envs over it)"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from h1v2_isaac_b200.backend import H1v2Sim
from h1v2_isaac_b200._capi import default_config

rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
sizes = [int(x) for x in os.environ.get("SWEEP_ENVS", "1024,2048,4096,8192,16384,32768,65536").split(",")]
rows = []
for randomize in (False, True):
    for n in sizes:
        cfg = default_config()
        cfg.env_id_offset = rank * n
        if randomize:
            cfg.push_enable = 1
            cfg.mass_add_range[0], cfg.mass_add_range[1] = -5.0, 5.0
            cfg.friction_range[0], cfg.friction_range[1] = 0.1, 1.25
        sim = H1v2Sim(n, cfg, device=dev, seed=42); sim.observe()
        acts = [sim.random_actions(i) for i in range(16)]
        obs = torch.empty((n, sim.obs_dim), device=dev); rew = torch.empty(n, device=dev)
        term = torch.empty(n, dtype=torch.uint8, device=dev); trunc = torch.empty(n, dtype=torch.uint8, device=dev)
        for i in range(40): sim.step_into(acts[i % 16], obs, rew, term, trunc)
        torch.cuda.synchronize()
        if world > 1: dist.barrier()
        K = 100
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
        for i in range(K):
            flush.zero_(); ev[i][0].record(); sim.step_into(acts[i % 16], obs, rew, term, trunc); ev[i][1].record()
        torch.cuda.synchronize()
        ms = sum(a.elapsed_time(b) for a, b in ev) / K
        lg = sim.log_host()
        stat = torch.tensor([ms, float(lg[30]) / (4 * n), float(lg[27])], dtype=torch.float64, device=dev)
        if world > 1:
            dist.barrier()
            mx = stat.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
            sm = stat.clone(); dist.all_reduce(sm, op=dist.ReduceOp.SUM)
            ms, iters, nan = float(mx[0]), float(sm[1]) / world, float(sm[2])
        else:
            ms, iters, nan = float(stat[0]), float(stat[1]), float(stat[2])
        rows.append({"envs_per_gpu": n, "n_gpus": world, "randomised": randomize, "ms_per_step": round(ms, 4), "env_steps_per_s": round(world * n / ms * 1e3),
                     "mean_newton_iters": round(iters, 3), "nan_resets_total": nan})
        if rank == 0: print(rows[-1], file=sys.stderr, flush=True)
        sim.close()
if rank == 0:
    print(json.dumps({"gpu": torch.cuda.get_device_name(0), "n_gpus": world, "timer": "per-step CUDA events, max over ranks", "rows": rows}, indent=1))
if world > 1:
    dist.destroy_process_group()
