import sys
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np, torch
from h1v2_isaac_b200._capi import default_config
from test_gpu_parity import PHYS, SYNC, _np, _resync, _mk, _randomised
for iters in (12, 30, 60):
  for lo, hi in ((0.10, 0.12), (0.2, 0.22)):
    c = _randomised(default_config()); c.friction_range[0], c.friction_range[1] = lo, hi; c.push_enable = 0; c.solver_iterations = iters
    n = 4096
    torch, sim, orc = _mk(c, n, 17)
    sim.observe(); orc.observe()
    rng = np.random.default_rng(5); nbad = 0; ncap = 0; tot = 0; allerr = []
    for step in range(30):
        a = rng.normal(size=(n, 12)).astype(np.float32)
        _, _, tg, ug = sim.step(torch.from_numpy(a).cuda())
        _, _, to, uo = orc.step(a)
        g, o = _np(sim.get_state(SYNC + ["solver_iters"])), orc.get_state(PHYS)
        mc, ml = orc.activation_margin()
        keep = ~(to | uo | tg.cpu().numpy()) & (mc > 2e-6) & (ml > 2e-6)
        err = np.abs(g["joint_vel"] - o["joint_vel"]).max(axis=1)
        nbad += int((keep & (err > 5e-3)).sum()); ncap += int((g["solver_iters"][:, 0] >= iters).sum()); tot += int(keep.sum()); allerr.append(err[keep])
        _resync(sim, orc, g)
    e = np.concatenate(allerr)
    print(f"max_iters {iters} mu [{lo},{hi}]: env-steps {tot} offenders(>5e-3) {nbad} cap hits {ncap} q999 {np.quantile(e,0.999):.2e}")
    sim.close()
