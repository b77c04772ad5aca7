import sys
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np, torch
from h1v2_isaac_b200.backend import H1v2Sim
from h1v2_isaac_b200._capi import default_config
from oracle.oracle import Oracle
from test_gpu_parity import PHYS, SYNC, _np, _resync
for lstol in (0.3, 0.5, 0.7, 1.0, 2.0):
    cfg = default_config(); cfg.solver_ls_tolerance = lstol
    n = 32768
    sim = H1v2Sim(n, cfg, seed=1); sim.observe()
    acts = [sim.random_actions(i) for i in range(8)]
    obs = torch.empty((n, sim.obs_dim), device='cuda'); rew = torch.empty(n, device='cuda')
    term = torch.empty(n, dtype=torch.uint8, device='cuda'); trunc = torch.empty(n, dtype=torch.uint8, device='cuda')
    for i in range(60): sim.step_into(acts[i % 8], obs, rew, term, trunc)
    h0 = sim.iter_hist().copy(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(100): sim.step_into(acts[i % 8], obs, rew, term, trunc)
    e1.record(); torch.cuda.synchronize()
    h = sim.iter_hist() - h0; frac = h / h.sum(); mean = (frac * np.arange(32)).sum(); cdf = np.cumsum(frac); emax = ((cdf ** 16)[1:] - (cdf ** 16)[:-1]) @ np.arange(1, 32)
    ms = e0.elapsed_time(e1) / 100
    sim.close()
    res = []
    for scale in (1.0, 0.3):
        c = cfg.copy(); c.decimation = 1; c.max_delay = 2
        m = 2048
        s2 = H1v2Sim(m, c, device="cuda:0", seed=3, diagnostics=True); orc = Oracle(c, m, seed=3, threads=16)
        s2.observe(); orc.observe(); rng = np.random.default_rng(0); ev = []
        for step in range(96):
            a = (scale * rng.normal(size=(m, 12))).astype(np.float32)
            _, _, tg, ug = s2.step(torch.from_numpy(a).cuda()); _, _, to, uo = orc.step(a)
            g, o = _np(s2.get_state(SYNC)), orc.get_state(PHYS)
            mc, ml = orc.activation_margin()
            keep = ~(to | uo | tg.cpu().numpy()) & (mc > 2e-6) & (ml > 2e-6)
            ev.append(np.abs(g["joint_vel"][keep] - o["joint_vel"][keep]).max(axis=1)); _resync(s2, orc, g)
        ev = np.concatenate(ev); res.append(f"scale {scale}: max {ev.max():.1e} q999 {np.quantile(ev, .999):.1e}")
        s2.close()
    print(f"ls_tol {lstol}: {ms:.4f} ms  iters mean {mean:.2f} E[max16] {emax:.2f} | " + " | ".join(res))
