"""End-to-end step time through h1v2_step_host (pinned HOST buffers, copies inside) in both host modes, beside the device-timed step.
usage: python tools/diag_e2e.py [n,n,...]"""
import os, sys, time, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
if len(sys.argv) > 2:  # child: mode n threads
    mode, n, nt = sys.argv[2], int(sys.argv[3]), sys.argv[4]
    if mode.startswith("hybrid"):
        os.environ["H1V2_HOST_PATH"], os.environ["H1V2_HOST_ROWS_FRAC"] = "hybrid", mode[6:] or "0.5"
    elif mode != "auto":
        os.environ["H1V2_HOST_PATH"] = mode
    if nt != "0": os.environ["H1V2_HOST_THREADS"] = nt
    import torch
    from h1v2_isaac_b200._capi import default_config
    from h1v2_isaac_b200.backend import H1v2Sim
    sim = H1v2Sim(n, default_config(), seed=1); sim.observe()
    pool = [sim.random_actions(i) for i in range(8)]
    ha = [p.cpu().pin_memory() for p in pool]
    hobs = torch.empty((n, sim.obs_dim)).pin_memory(); hrew = torch.empty(n).pin_memory()
    ht = torch.empty(n, dtype=torch.uint8).pin_memory(); hu = torch.empty(n, dtype=torch.uint8).pin_memory()
    for i in range(48): sim.step_host(ha[i % 8], hobs, hrew, ht, hu)
    best = 1e9
    for rep in range(3):
        t0 = time.perf_counter()
        for i in range(100): sim.step_host(ha[i % 8], hobs, hrew, ht, hu)
        best = min(best, (time.perf_counter() - t0) / 100)
    # device-timed step for reference
    obs = torch.empty((n, sim.obs_dim), device='cuda'); rew = torch.empty(n, device='cuda')
    term = torch.empty(n, dtype=torch.uint8, device='cuda'); trunc = torch.empty(n, dtype=torch.uint8, device='cuda')
    for i in range(20): sim.step_into(pool[i % 8], obs, rew, term, trunc)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(100): sim.step_into(pool[i % 8], obs, rew, term, trunc)
    e1.record(); torch.cuda.synchronize()
    dev = e0.elapsed_time(e1) / 100
    mode = f"{mode} -> chose mode {sim.host_path_info()[0]}, rows for {sim.host_path_rows()} envs" if mode == "auto" else mode
    print(f"n={n} mode={mode} threads={nt}: e2e {best * 1e3:.4f} ms ({n / best / 1e6:.2f} M/s)  device {dev:.4f} ms  e2e/device rate ratio {dev / (best * 1e3):.3f}", flush=True)
    sys.exit(0)
ns = sys.argv[1] if len(sys.argv) > 1 else "4096,32768"
for n in ns.split(","):
    for mode, nt in (("rows", "0"), ("assemble", "0"), ("hybrid0.25", "0"), ("hybrid0.5", "0"), ("hybrid0.75", "0"), ("auto", "0"), ("hybrid0.5", "8"), ("hybrid0.5", "4"), ("auto", "4")):
        subprocess.run([sys.executable, __file__, "x", mode, n, nt])
