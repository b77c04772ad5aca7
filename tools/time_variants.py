"""Time library variants (build/variants/lib_*.so) on the GPU box: python tools/time_variants.py [n,n,...] [epw,epw,...] [name-prefix]
Each variant runs in its own process (H1V2_LIB selects the library).  L2 is not flushed: relative numbers only."""
import glob, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ns = sys.argv[1] if len(sys.argv) > 1 else "4096,32768"
epws = sys.argv[2] if len(sys.argv) > 2 else "0"
prefix = sys.argv[3] if len(sys.argv) > 3 else ""
code = r"""
import sys; sys.path.insert(0, %r)
import torch
from h1v2_isaac_b200.backend import H1v2Sim
from h1v2_isaac_b200._capi import default_config, rsl_config
out = []
for n in [int(x) for x in %r.split(',')]:
    for epw in [int(x) for x in %r.split(',')]:
        cfg = default_config(); cfg.reserved[2] = epw
        sim = H1v2Sim(n, cfg, seed=1); sim.observe()
        acts = [sim.random_actions(i) for i in range(8)]
        obs = torch.empty((n, sim.obs_dim), device='cuda'); rew = torch.empty(n, device='cuda')
        term = torch.empty(n, dtype=torch.uint8, device='cuda'); trunc = torch.empty(n, dtype=torch.uint8, device='cuda')
        for i in range(40): sim.step_into(acts[i %% 8], obs, rew, term, trunc)
        torch.cuda.synchronize()
        best = 1e9
        for rep in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(100): sim.step_into(acts[i %% 8], obs, rew, term, trunc)
            e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1) / 100)
        lg = sim.log_host()
        out.append(f"n={n} epw={epw}: {best:.4f} ms ({n / best / 1e3:.1f} M/s) rew {float(rew.mean()):.5f} iters/substep {lg[30] / (4 * n):.3f} max {lg[28]:.0f}")
        sim.close()
print(' | '.join(out))
""" % (ROOT, ns, epws)
libs = sorted(glob.glob(os.path.join(ROOT, "build", "variants", f"lib_{prefix}*.so")))
for path in libs:
    env = dict(os.environ, H1V2_LIB=path)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env)
    print(os.path.basename(path), r.stdout.strip() or r.stderr.strip()[-400:], flush=True)
