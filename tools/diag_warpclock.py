"""Where the 4096-env step goes, warp by warp (diagnostic build -DH1V2_WARPCLOCK, build/variants/lib_wc.so):
every warp records its entry time, the cycles of its physics loop and of the whole kernel, and how many warp-synchronous
Newton / line-search trips it ran.  Prints, per configuration, the launch span against the mean / slowest warp and the
trip statistics: python tools/diag_warpclock.py [n] (H1V2_LIB must point at the variant)."""
import ctypes as C, os, sys
sys.path.insert(0, '/root/repo')
import numpy as np, torch
from h1v2_isaac_b200.backend import H1v2Sim
from h1v2_isaac_b200 import _capi
from h1v2_isaac_b200._capi import default_config

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
lib = _capi.load_library()
lib.h1v2_debug_warpclock.argtypes = [C.c_void_p, C.c_int]
for name, cap, epw in (("cap12", 12, 0), ("cap6", 6, 0)):
    cfg = default_config(); cfg.solver_iterations = cap; cfg.reserved[2] = epw
    sim = H1v2Sim(n, cfg, seed=1); sim.observe()
    acts = [sim.random_actions(i) for i in range(8)]
    obs = torch.empty((n, sim.obs_dim), device='cuda'); rew = torch.empty(n, device='cuda')
    term = torch.empty(n, dtype=torch.uint8, device='cuda'); trunc = torch.empty(n, dtype=torch.uint8, device='cuda')
    for i in range(60): sim.step_into(acts[i % 8], obs, rew, term, trunc)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(100): sim.step_into(acts[i % 8], obs, rew, term, trunc)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 100
    e = epw if epw else (8 if n <= 4096 else 16)
    nw = min((n + e - 1) // e, 16384)
    rows = []
    for i in range(40):
        sim.step_into(acts[i % 8], obs, rew, term, trunc); torch.cuda.synchronize()
        buf = np.zeros(4 * nw, dtype=np.uint64)
        assert lib.h1v2_debug_warpclock(buf.ctypes.data, nw) == 0
        b = buf.reshape(nw, 4)
        t0 = b[:, 0].astype(np.float64); cyc = b[:, 2].astype(np.float64); phys = b[:, 1].astype(np.float64)
        trips = (b[:, 3] & np.uint64(0xffffffff)).astype(np.float64); ls = (b[:, 3] >> np.uint64(32)).astype(np.float64)
        A = np.stack([trips, ls, np.ones(nw)], 1); coef = np.linalg.lstsq(A, phys, rcond=None)[0]
        rows.append([cyc.mean(), cyc.max(), np.quantile(cyc, 0.9), phys.mean(), phys.max(), (t0.max() - t0.min()) / 1e3, trips.mean(), trips.max(), ls.mean(), ls.max(), *coef,
                     np.corrcoef(phys, A @ coef)[0, 1]])
    r = np.array(rows).mean(0)
    f = 1.965e3  # cycles per us
    print(f"{name}: n={n} epw={e} warps={nw} step {ms * 1e3:.1f} us | warp kernel mean {r[0] / f:.1f} us, q90 {r[2] / f:.1f}, max {r[1] / f:.1f}; physics loop mean {r[3] / f:.1f}, max {r[4] / f:.1f}; "
          f"entry spread {r[5]:.1f} us | trips mean {r[6]:.1f} max {r[7]:.1f}, ls trips mean {r[8]:.1f} max {r[9]:.1f} | fit: {r[10] / f:.2f} us/trip + {r[11] / f:.2f} us/ls-trip + {r[12] / f:.1f} us (corr {r[13]:.3f})", flush=True)
    sim.close()
