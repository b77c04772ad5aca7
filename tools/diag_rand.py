import sys
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np, torch
from h1v2_isaac_b200._capi import default_config
from test_gpu_parity import PHYS, SYNC, _np, _resync, _mk, _randomised
lst = float(sys.argv[1]) if len(sys.argv) > 1 else 0.3
c = _randomised(default_config()); c.solver_ls_tolerance = lst
if len(sys.argv) > 2: c.reserved[1] = int(sys.argv[2])
n = 2048
torch, sim, orc = _mk(c, n, 17)
sim.observe(); orc.observe()
rng = np.random.default_rng(5)
for step in range(40):
    a = rng.normal(size=(n, 12)).astype(np.float32)
    _, _, tg, ug = sim.step(torch.from_numpy(a).cuda())
    _, _, to, uo = orc.step(a)
    g, o = _np(sim.get_state(SYNC + ["solver_iters"])), orc.get_state(PHYS + ["push_time_left"])
    mc, ml = orc.activation_margin()
    pushed = np.abs(g["push_time_left"][:, 0] - o["push_time_left"][:, 0]) > 1e-6
    keep = ~(to | uo | tg.cpu().numpy()) & (mc > 2e-6) & (ml > 2e-6) & ~pushed
    err = np.abs(g["joint_vel"] - o["joint_vel"]).max(axis=1)
    bad = keep & (err > 5e-3)
    oit, ores = orc.solver_stats()
    for i in np.nonzero(bad)[0][:5]:
        print(f"step {step} env {i}: err {err[i]:.3e} friction {g['friction'][i,0]:.3f} mass_add {g['mass_add'][i,0]:.2f} gpu iters max/sum {g['solver_iters'][i]} oracle iters {oit[i]} res {ores[i]:.2e} margins {mc[i]:.2e} {ml[i]:.2e}")
    _resync(sim, orc, g)
print("done")
