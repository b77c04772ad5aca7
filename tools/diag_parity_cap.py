import sys
sys.path.insert(0, '/root/repo')
import numpy as np, torch
from h1v2_isaac_b200.backend import H1v2Sim
from h1v2_isaac_b200._capi import default_config
from oracle.oracle import Oracle
sys.path.insert(0, '/root/repo/tests')
from test_gpu_parity import PHYS, SYNC, _np, _resync
for cap, scale in [(int(x), sc) for x in (sys.argv[1:] or [4, 12]) for sc in (1.0, 0.3)]:
    c = default_config(); c.decimation = 1; c.max_delay = min(c.max_delay, 2); c.reserved[1] = cap
    n = 2048
    sim = H1v2Sim(n, c, device="cuda:0", seed=3, diagnostics=True); orc = Oracle(c, n, seed=3, threads=16)
    sim.observe(); orc.observe()
    rng = np.random.default_rng(0)
    errs = {k: [] for k in PHYS}; its = []
    for step in range(96):
        a = (scale * rng.normal(size=(n, 12))).astype(np.float32)
        _, _, tg, ug = sim.step(torch.from_numpy(a).cuda())
        _, _, to, uo = orc.step(a)
        g, o = _np(sim.get_state(SYNC + ["solver_iters"])), orc.get_state(PHYS)
        mc, ml = orc.activation_margin()
        keep = ~(to | uo | tg.cpu().numpy()) & (mc > 2e-6) & (ml > 2e-6)
        for k in PHYS:
            errs[k].append(np.abs(g[k][keep] - o[k][keep]).max(axis=1))
        its.append(g["solver_iters"][keep, 0])
        _resync(sim, orc, g)
    e = {k: np.concatenate(v) for k, v in errs.items()}; its = np.concatenate(its)
    bad = e["joint_vel"] > 5e-3
    print(f"cap {cap} scale {scale}: joint_vel max {e['joint_vel'].max():.2e} q999 {np.quantile(e['joint_vel'], 0.999):.2e}  n>5e-3: {bad.sum()}  iters of offenders {its[bad][:10]}  ang_vel max {e['root_ang_vel'].max():.2e}")
    sim.close()
