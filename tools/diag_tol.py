import sys
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np, torch
from h1v2_isaac_b200.backend import H1v2Sim
from h1v2_isaac_b200._capi import default_config
from oracle.oracle import Oracle
from test_gpu_parity import PHYS, SYNC, _np, _resync
for tol in (1e-6, 3e-6, 1e-5, 3e-5):
    for dec, scale in ((1, 1.0), (4, 1.0)):
        c = default_config(); c.decimation = dec; c.max_delay = min(c.max_delay, 2 * dec); c.solver_tolerance = tol
        n = 2048
        sim = H1v2Sim(n, c, device="cuda:0", seed=3, diagnostics=True); orc = Oracle(default_config() if False else c, n, seed=3, threads=16)
        sim.observe(); orc.observe()
        rng = np.random.default_rng(0)
        errs = {k: [] for k in PHYS}
        for step in range(24 * (4 // dec)):
            a = (scale * rng.normal(size=(n, 12))).astype(np.float32)
            _, _, tg, ug = sim.step(torch.from_numpy(a).cuda())
            _, _, to, uo = orc.step(a)
            g, o = _np(sim.get_state(SYNC)), orc.get_state(PHYS)
            mc, ml = orc.activation_margin()
            keep = ~(to | uo | tg.cpu().numpy()) & (mc > 2e-6) & (ml > 2e-6)
            for k in PHYS:
                errs[k].append(np.abs(g[k][keep] - o[k][keep]).max(axis=1))
            _resync(sim, orc, g)
        e = {k: np.concatenate(v) for k, v in errs.items()}
        h = sim.iter_hist(); frac = h / h.sum(); mean = (frac * np.arange(32)).sum(); cdf = np.cumsum(frac); emax = ((cdf ** 16)[1:] - (cdf ** 16)[:-1]) @ np.arange(1, 32)
        jv = e["joint_vel"]
        print(f"tol {tol:.0e} dec {dec}: iters mean {mean:.2f} E[max16] {emax:.2f} | joint_vel max {jv.max():.2e} q999 {np.quantile(jv,0.999):.2e} q99 {np.quantile(jv,0.99):.2e} med {np.median(jv):.2e} | joint_pos max {e['joint_pos'].max():.2e} | ang_vel q999 {np.quantile(e['root_ang_vel'],0.999):.2e}")
        sim.close()
