#!/usr/bin/env python
"""bench.py -- env-steps/sec of the fused control step (physics + obs + reward + termination + reset), B200.

Contract (driver): `python bench.py --gpus N --steps K --warmup W [--impl reference]` prints ONE JSON line on rank 0.
  * ours:      one H1v2Sim per rank (envs shard with no data-path collective -> "scaling": "weak"), K timed control
               steps on synthetic N(0,1) actions resident in HBM, per-step CUDA events, L2 flushed between steps,
               max over ranks.  `e2e` is the same metric through h1v2_step_host (HOST buffers, copies inside).
  * reference: the reference's CPU path restated (oracle/h1v2_oracle.c, float64, MuJoCo-semantics physics at the Isaac
               timing 5 ms x 4) on all host cores, on the SAME config as the GPU arm (env count, warm-up); it loads the
               oracle library only (its own copy of the resolved configs), never libh1v2_b200.so.
Workload = BASELINE.json configs[1]: Isaac-Velocity-Flat-H12_12dof-v0 step, 4096 envs, random actions (--envs overrides).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "env-steps/sec"
# algorithmic HBM bytes per env-step of the fused kernel: action 48, root r+w 128, leg 192,
# actuator line 320, command 64, timers 64, warm start 192, episode sums 192, episode length 16, reward+flags 6 = 1222,
# plus per history slot H: ring read (H-1)x180 + ring write 180 + observation write 180 H = 360 H  (H = 10: 4822)
def algo_bytes_per_env_step(history: int, obs_dim: int | None = None) -> int:
    if obs_dim is not None and obs_dim != 45 * history:  # Rough id: no history ring to read; one 192-byte slot and the 235-float row are written
        return 1222 + 192 + 4 * obs_dim
    return 1222 + 360 * history


TASKS = {"flat": "Isaac-Velocity-Flat-H12_12dof-v0", "rsl": "Isaac-Velocity-Rsl-H12_12dof-v0", "cat": "Isaac-Velocity-CaT-Flat-H12_12dof-v0",
         "rough": "Isaac-Velocity-Rough-H12_12dof-v0"}


def task_config(task: str):
    from h1v2_isaac_b200 import _capi, tasks
    return {"flat": _capi.default_config, "rsl": _capi.rsl_config, "cat": tasks.cat_config, "rough": _capi.rough_config}[task]()


def oracle_task_config(task: str):
    """The same resolved configs from the ORACLE library (oracle/Makefile links its own h1v2_config.cpp): the reference arm and the
    cpu_baseline leg must not load the CUDA library."""
    from oracle.oracle import task_config as otc
    return otc(task)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--warmup", type=int, default=30)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--envs", type=int, default=4096, help="envs per GPU (BASELINE configs[1] = 4096; configs[3] = 32768)")
    ap.add_argument("--task", default="flat", choices=sorted(TASKS), help="flat = BASELINE's metric config (default); rsl / cat / rough = the SURVEY 8(f) variants (cat: fused step + constraint tail; rough: height-field terrain, height scan)")
    ap.add_argument("--seed", type=int, default=42)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-big", action="store_true", help="skip the extra 32768-envs/GPU measurement")
    ap.add_argument("--no-ppo", action="store_true", help="skip the PPO-loop measurement (BASELINE configs[2]/[3])")
    return ap.parse_args()


class ClockSampler:
    """nvidia-smi sampling DURING the timed region (B200_PROFILING.md clocks line)."""

    def __init__(self, index: int):
        self.rows, self.proc = [], None
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={index}", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = sorted(float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit())
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 3 + i and r[3 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


def cpu_baseline(cfg, seed: int, n: int, target_s: float = 12.0, steps: int | None = None, warmup: int = 3, max_s: float = 150.0):
    """Time the CPU restatement of the reference path on all host cores, on the workload's own env count; the number of timed
    steps is bounded (steps=None: about target_s seconds; else `steps`, cut so that the run ends within max_s)."""
    import numpy as np
    from oracle.oracle import Oracle

    cores = os.cpu_count() or 1
    orc = Oracle(cfg, n, seed=seed, threads=cores)
    orc.observe()
    rng = np.random.default_rng(0)
    acts = [rng.normal(size=(n, 12)).astype(np.float32) for _ in range(4)]
    t0 = time.perf_counter()
    orc.step(acts[0])
    one = time.perf_counter() - t0
    warmup = max(0, min(warmup - 1, int(0.2 * max_s / max(one, 1e-6))))
    for i in range(warmup):
        orc.step(acts[(i + 1) % 4])
    k = max(3, int(target_s / max(one, 1e-6))) if steps is None else max(1, min(steps, int(max_s / max(one, 1e-6))))
    t0 = time.perf_counter()
    for i in range(k):
        orc.step(acts[i % 4])
    dt = time.perf_counter() - t0
    return {"value": n * k / dt, "unit": METRIC, "cores": cores, "kind": "port",
            "sample": f"{n} envs x {k} control steps (4 x 5 ms MuJoCo-semantics substeps + managers), float64 C oracle, {cores} pthreads",
            "ms_per_step": dt / k * 1e3, "n_envs": n, "steps": k, "warmup": warmup + 1}


def cpu_baseline_sim2sim(cfg, seed: int, target_s: float = 4.0):
    """BASELINE.md B0 -- the reference's literal CPU configuration (scripts/deploy/sim2sim.py:46-54, deploy/config.yaml:7-8):
    ONE env on ONE core, MuJoCo-semantics step at 1 ms x 20 substeps per 50 Hz control step, PD in the loop, random-init
    obs_dim->512->256->128->12 ELU policy on torch-CPU (obs_dim 450 on the Flat id).  Engine: the float64 C oracle (mujoco is not installable here)."""
    import numpy as np
    import torch
    from oracle.oracle import Oracle

    c = cfg.copy()
    c.sim_dt, c.decimation = 0.001, 20
    orc = Oracle(c, 1, seed=seed, threads=1)
    torch.manual_seed(0)
    torch.set_num_threads(1)
    pi = torch.nn.Sequential(torch.nn.Linear(orc.obs_dim, 512), torch.nn.ELU(), torch.nn.Linear(512, 256), torch.nn.ELU(),
                             torch.nn.Linear(256, 128), torch.nn.ELU(), torch.nn.Linear(128, 12))
    obs = orc.observe()
    k, t0 = 0, time.perf_counter()
    with torch.inference_mode():
        while time.perf_counter() - t0 < target_s:
            a = pi(torch.from_numpy(obs)).numpy().astype(np.float32)
            obs, _, _, _ = orc.step(a)
            k += 1
    dt = time.perf_counter() - t0
    return {"value": k / dt, "unit": METRIC, "cores": 1, "kind": "port", "physics_steps_per_s": 20 * k / dt,
            "sample": f"1 env x {k} control steps (20 x 1 ms MuJoCo-semantics substeps + PD + torch-CPU MLP policy), float64 C oracle, 1 thread"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg = oracle_task_config(args.task)
    cfg.cat_enable = 0  # the CPU restatement of the constraint tail is numpy test infrastructure (oracle/cat_oracle.py), not timed here
    cb = cpu_baseline(cfg, args.seed, args.envs, steps=max(1, args.steps), warmup=max(args.warmup, 3))
    line = {
        "impl": "reference", "metric": METRIC, "value": cb["value"], "unit": METRIC, "n_gpus": args.gpus, "steps": cb["steps"],
        "warmup": cb["warmup"], "ms_per_step": cb["ms_per_step"], "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"{TASKS[args.task]} physics+obs+reward step, {args.envs} envs, random N(0,1) actions" + (" (BASELINE configs[1] when 4096)" if args.task == "flat" else ""),
                   "envs_per_gpu": args.envs, "decimation": 4, "sim_dt": 0.005, "history": int(cfg.history_length),
                   "engine": "CPU restatement of the MuJoCo sim2sim path (oracle/h1v2_oracle.c); mujoco itself is not installable here"},
        "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": cb["value"], "unit": METRIC, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    from h1v2_isaac_b200 import _capi
    from h1v2_isaac_b200._capi import default_config, load_library
    from h1v2_isaac_b200.backend import H1v2Sim

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # NCCL prints its version banner on STDOUT at communicator creation when NCCL_DEBUG is set on the box; the contract
        # is ONE JSON line on stdout, so the banner is sent to stderr (fd-level: it comes from the C library).
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    n = args.envs
    cfg = task_config(args.task)
    is_cat = args.task == "cat"
    ALGO_BYTES_PER_ENV_STEP = algo_bytes_per_env_step(cfg.history_length, _capi.obs_dim_of(cfg))
    cfg.env_id_offset = rank * n  # envs shard across ranks; the Philox key uses the global env id
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2
    W, K = max(args.warmup, 3), args.steps

    def timed(n_envs, steps):
        """K timed control steps of one H1v2Sim: per-step CUDA events, L2 flushed (untimed) between steps, max over ranks."""
        c = cfg.copy()
        c.env_id_offset = rank * n_envs
        sim = H1v2Sim(n_envs, c, device=dev, seed=args.seed)
        sim.observe()
        pool = [sim.random_actions(i) for i in range(16)]  # synthetic N(0,1) actions, resident in HBM
        obs = torch.empty((n_envs, sim.obs_dim), device=dev)
        rew = torch.empty(n_envs, device=dev)
        term = torch.empty(n_envs, dtype=torch.float32 if is_cat else torch.uint8, device=dev)  # CaT: float dones
        trunc = torch.empty(n_envs, dtype=torch.uint8, device=dev)
        step_into = sim.cat_step_into if is_cat else sim.step_into
        for i in range(W):
            step_into(pool[i % 16], obs, rew, term, trunc)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        l0 = sim.launch_count
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        for i in range(steps):
            flush.zero_()  # L2 flush between timed iterations (not timed)
            ev[i][0].record()
            step_into(pool[i % 16], obs, rew, term, trunc)
            ev[i][1].record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t = torch.tensor([sum(a.elapsed_time(b) for a, b in ev)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return sim, pool, float(t.item()), sim.launch_count - l0

    sampler = ClockSampler(local) if rank == 0 else None
    sim, pool, total_ms, launches = timed(n, K)
    clocks = sampler.stop() if sampler else None
    ms_per_step = total_ms / K
    value = world * n * K / (total_ms * 1e-3)
    logv = sim.log_host()

    # ---- e2e: the same metric through the C-ABI with HOST buffers (pinned), copies inside the timed region ----
    def e2e_run(sim_, pool_, n_envs):
        ha = [p.cpu().pin_memory() for p in pool_[:4]]
        hobs = torch.empty((n_envs, sim_.obs_dim), dtype=torch.float32).pin_memory()
        hrew = torch.empty(n_envs, dtype=torch.float32).pin_memory()
        hterm = torch.empty(n_envs, dtype=torch.uint8).pin_memory()
        htrunc = torch.empty(n_envs, dtype=torch.uint8).pin_memory()
        Ke = max(10, min(K, 100))
        if is_cat:  # h1v2_cat_step_host: float dones instead of the terminated flags
            hterm = torch.empty(n_envs, dtype=torch.float32).pin_memory()
        host_step = sim_.cat_step_host if is_cat else sim_.step_host
        for i in range(160):  # the first forty calls also time the host paths against each other (h1v2_host_path_info); a watchdog may repeat that once
            host_step(ha[i % 4], hobs, hrew, hterm, htrunc)
        # three blocks of Ke synchronous calls, each between barriers and reduced with MAX over the ranks; the MEDIAN block is reported: a block
        # is ~25 ms of wall clock at 4096 envs, so one descheduled host thread on one of N ranks would otherwise be the whole result
        blocks = []
        for _ in range(3):
            if world > 1:
                dist.barrier()
            t0 = time.perf_counter()
            for i in range(Ke):
                host_step(ha[i % 4], hobs, hrew, hterm, htrunc)  # synchronises inside
            tb = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
            if world > 1:
                dist.all_reduce(tb, op=dist.ReduceOp.MAX)
            blocks.append(float(tb.item()))
        td = torch.tensor([sorted(blocks)[1]], device=dev, dtype=torch.float64)
        mode, threads = sim_.host_path_info()
        n_rows = sim_.host_path_rows()
        # "assemble": only the new 45-float sample (a 192-byte slot) of an env crosses PCIe, host threads assemble its [obs_dim] row in the
        # caller's buffer; "rows": the kernel writes the rows themselves (zero-copy); "hybrid": rows for the first n_rows envs, samples for the rest
        d2h = n_rows * sim_.obs_dim * 4 + (n_envs - n_rows) * 48 * 4 + n_envs * (4 + 1 + 1)
        return {"value": world * n_envs * Ke / float(td.item()), "unit": METRIC, "h2d_bytes_per_step": n_envs * 12 * 4, "d2h_bytes_per_step": d2h,
                "steps": Ke, "ms_per_step": float(td.item()) / Ke * 1e3, "timer": "host wall clock around synchronous h1v2_step_host calls, max over ranks; median of three blocks of `steps` calls",
                "blocks_ms_per_step": [round(x / Ke * 1e3, 5) for x in blocks],
                "host_path": {"mode": ("rows" if mode == 0 else ("assemble" if n_rows == 0 else "hybrid")), "host_threads": threads, "envs_with_rows_over_pcie": n_rows,
                              "rows_bytes_written_by_host_threads_per_step": (n_envs - n_rows) * sim_.obs_dim * 4 if mode == 1 else 0}}

    e2e = None
    if not args.no_e2e:
        e2e = e2e_run(sim, pool, n)

    big = None
    if n != 32768 and not args.no_big:  # the north-star target is quoted at 32768 envs/GPU: report it beside configs[1]
        sim.close()
        sampler_b = ClockSampler(local) if rank == 0 else None  # clocks / throttle reasons of THIS timed region too
        sim_b, pool_b, ms_b, _ = timed(32768, max(20, K // 3))
        clocks_b = sampler_b.stop() if sampler_b else None
        kb = max(20, K // 3)
        big = {"envs_per_gpu": 32768, "value": world * 32768 * kb / (ms_b * 1e-3), "unit": METRIC, "ms_per_step": ms_b / kb, "steps": kb,
               "clocks": clocks_b, "e2e": None if args.no_e2e else e2e_run(sim_b, pool_b, 32768)}
        sim_b.close()
    # ---- BASELINE configs[2] / [3]: the PPO loop of scripts/rsl_rl/train.py:120-141 on this backend (RslRlVecEnvWrapper ->
    #      OnPolicyRunner.learn, random-init ActorCritic [512, 256, 128], 24 steps per env, 5 epochs x 4 mini-batches), with the
    #      learner's NCCL collectives inside when N > 1: 20 gradient all-reduces (791 961 parameters, 3.17 MB) + 20 KL reductions
    #      + the rollout-statistics reductions per iteration.  Reported beside the step metric, not instead of it.
    def ppo_leg(n_envs, iters):
        import contextlib
        import tempfile
        from h1v2_isaac_b200 import shims, tasks
        shims.install()
        tasks.register()
        import gymnasium as gym
        from isaaclab_rl.rsl_rl import RslRlVecEnvWrapper
        from rsl_rl.runners import OnPolicyRunner
        with contextlib.redirect_stdout(sys.stderr):  # the env and the runner print progress lines; stdout carries ONE JSON line
            torch.manual_seed(args.seed + rank)
            torch.backends.cuda.matmul.allow_tf32 = True  # as scripts/rsl_rl/train.py:70-73 sets them
            torch.backends.cudnn.allow_tf32 = True
            env = gym.make(TASKS["flat"], cfg=tasks.default_env_cfg(n_envs, device=f"cuda:{local}"))
            agent = tasks.default_agent_cfg()
            runner = OnPolicyRunner(RslRlVecEnvWrapper(env), agent.to_dict(), log_dir=tempfile.mkdtemp(prefix="h1v2_bench_ppo_"), device=f"cuda:{local}")
            runner.learn(num_learning_iterations=3, init_at_random_ep_len=True)  # warm-up: eager body, graph capture, first plain replay (uploads the graphs)
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            t0 = time.perf_counter()
            runner.learn(num_learning_iterations=iters)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            td = torch.tensor([dt], device=dev, dtype=torch.float64)
            if world > 1:
                dist.all_reduce(td, op=dist.ReduceOp.MAX)
            T = runner.num_steps_per_env
            col = sum(c for c, _ in runner.iteration_times) / len(runner.iteration_times)
            lrn = sum(l for _, l in runner.iteration_times) / len(runner.iteration_times)
            out = {"value": world * n_envs * T * iters / float(td.item()), "unit": METRIC, "envs_per_gpu": n_envs, "iterations": iters, "steps_per_env": T,
                   "s_per_iteration": float(td.item()) / iters, "collection_s_per_iteration": col, "learn_s_per_iteration": lrn,
                   "collection_ms_per_control_step": col / T * 1e3, "graph_rollout": bool(runner.graph_rollout),
                   "policy": "random-init ActorCritic [512, 256, 128] ELU, fp32 parameters, TF32 matmul allowed (scripts/rsl_rl/train.py:70-73)", "mean_episode_length": runner.stats.get("mean_episode_length"),
                   "collectives_per_iteration": (f"NCCL x{world}: 20 gradient all-reduces of {sum(p.numel() for p in runner.alg.policy.parameters())} parameters, 20 KL all-reduces, 2 rollout-statistics all-reduces" if world > 1 else "none (1 GPU)"),
                   "timer": "host wall clock around OnPolicyRunner.learn, synchronised, max over ranks"}
            runner.release_graphs()  # captured graphs (with NCCL work inside when H1V2_GRAPH_LEARNER=1) go before the process group does
            env.close()
            del runner
        return out

    ppo = None
    if not args.no_ppo and args.task == "flat":
        ppo = ppo_leg(n, 4)
        if big is not None:
            big["ppo"] = ppo_leg(32768, 3)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        hbm_peak, peak_src = json.load(open(peaks_path))["hbm_gbs"], "MEASURED_PEAKS.json hbm_gbs (of measured)"
    else:
        hbm_peak, peak_src = 6650.0, "B200_PROFILING.md fallback (of fallback)"
    hbm_achieved = ALGO_BYTES_PER_ENV_STEP * n / (ms_per_step * 1e-3) / 1e9
    import ctypes as C
    fp = C.c_float(0.0)
    load_library().h1v2_measure_fp32_peak(local, C.byref(fp))  # FFMA micro-benchmark on this GPU, outside every timed region
    fp32_peak = float(fp.value)
    rf_path = os.path.join(ROOT, "profiles", "roofline.json")
    rf = json.load(open(rf_path)) if os.path.exists(rf_path) else {}
    # FP32 work per env-step: the instrumented oracle's operation count on this workload, frozen in profiles/roofline.json
    # (oracle/flopcount/count.py); the Flat id's count is used for every task
    flops = rf.get("fp32_flops_per_env_step")
    traffic = rf.get("dram_bytes_per_launch_l2_flushed", {}).get(str(n))  # ncu dram read+write of one launch after an L2 flush, else null
    fp32_achieved = (flops * n / (ms_per_step * 1e-3) / 1e12) if flops else None
    wi = rf.get("warp_instructions_per_env_step", {}).get(str(n)) or rf.get("warp_instructions_per_env_step", {}).get("32768")
    clk = (clocks or {}).get("sm_mhz") or 1965.0

    def fracs(ms, nn):
        return {"roofline_fp32_frac": (flops * nn / (ms * 1e-3) / 1e12 / fp32_peak) if flops and fp32_peak > 0 else None,
                "roofline_hbm_frac": ALGO_BYTES_PER_ENV_STEP * nn / (ms * 1e-3) / 1e9 / hbm_peak}

    line = {
        "metric": METRIC, "value": value, "unit": METRIC, "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms_per_step,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{TASKS[args.task]} physics+obs+reward step, {n} envs/GPU, random N(0,1) actions" + (" (BASELINE configs[1] when 4096)" if args.task == "flat" else " (SURVEY 8(f) variant; not BASELINE's metric config)"),
                   "envs_per_gpu": n, "decimation": 4, "sim_dt": 0.005, "history": int(cfg.history_length), "obs_dim": sim.obs_dim, "parallelism": f"env-shard x{world}",
                   "l2": "flushed between timed steps (256 MiB write, untimed); per-step CUDA events summed"},
        "clocks": clocks,
        "e2e": e2e,
        "gpu_launches": int(launches),
        # the step is bounded by the FP32 (non-tensor) pipe first, HBM second
        "roofline": {"bound": "fp32", "achieved": fp32_achieved, "peak": fp32_peak, "unit": "TFLOP/s",
                     "frac": (fp32_achieved / fp32_peak) if fp32_achieved and fp32_peak > 0 else None, "traffic": traffic,
                     "peak_source": "h1v2_measure_fp32_peak: FFMA micro-benchmark on this GPU in this run (MEASURED_PEAKS.json has no FP32 entry)",
                     "flops_per_env_step": flops,
                     "flops_source": "instrumented oracle (oracle/flopcount): exact add+mul+div+sqrt+transcendental count of the float64 restatement on this workload, profiles/roofline.json",
                     "traffic_note": "dram bytes of one launch, ncu --set full capture taken after an L2 flush like the timed steps (null: no such capture for this env count)"},
        "roofline_hbm": {"bound": "hbm", "achieved": hbm_achieved, "peak": hbm_peak, "unit": "GB/s", "frac": hbm_achieved / hbm_peak,
                         "peak_source": peak_src, "algorithmic_bytes_per_env_step": ALGO_BYTES_PER_ENV_STEP},
        "roofline_issue": ({"bound": "warp-instruction issue", "unit": "G warp-instr/s", "peak": 148 * 4 * clk * 1e-3,
                            "achieved": wi * n / (ms_per_step * 1e-3) / 1e9, "frac": wi * n / (ms_per_step * 1e-3) / 1e9 / (148 * 4 * clk * 1e-3),
                            "warp_instructions_per_env_step": wi,
                            "note": "instruction count per env-step from the ncu capture in profiles/roofline.json; peak = 148 SMs x 4 schedulers x SM clock"} if wi else None),
        "at_32768_envs_per_gpu": (dict(big, **fracs(big["ms_per_step"], 32768)) if big else None),
        "solver": {"mean_newton_iters_per_substep": float(logv[_capi.LOG_SUM_ITERS]) / (4.0 * n), "max_iters_last_step": float(logv[_capi.LOG_MAX_ITERS]),
                   "cap_hits_last_step": float(logv[_capi.LOG_CAP_HITS]), "nan_resets": float(logv[_capi.LOG_NAN_RESETS])},
    }
    if ppo is not None:
        line["ppo"] = ppo
    if not args.no_cpu_baseline and world == 1:
        try:
            ocfg = oracle_task_config(args.task)
            ocfg.cat_enable = 0
            cb = cpu_baseline(ocfg, args.seed, n)
            line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
            line["cpu_baseline_sim2sim_1core"] = cpu_baseline_sim2sim(ocfg, args.seed)
        except Exception as e:  # the checker library missing must not hide the GPU number
            line["cpu_baseline"] = {"value": None, "unit": METRIC, "cores": os.cpu_count(), "kind": "port", "sample": f"unavailable: {e}"}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
