"""Target of tools/sanitize.sh: a few control steps of every kernel path at a small env count (flat step, observe, API reset,
state io, host path in both modes, CaT step), so that compute-sanitizer sees each kernel at least once."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from h1v2_isaac_b200 import tasks
from h1v2_isaac_b200._capi import default_config, rsl_config
from h1v2_isaac_b200.backend import H1v2Sim
n = int(sys.argv[1]) if len(sys.argv) > 1 else 192
for name, cfg in (("flat", default_config()), ("rsl", rsl_config())):
    sim = H1v2Sim(n, cfg, seed=3, diagnostics=(name == "flat")); sim.observe()
    for i in range(3):
        sim.step(sim.random_actions(i))
    sim.reset(torch.tensor([0, 5, n - 1], device="cuda"))
    st = sim.get_state(["joint_pos", "obs_history"]); sim.set_state(st)
    for mode in ("rows", "assemble"):
        os.environ["H1V2_HOST_PATH"] = mode
        s2 = H1v2Sim(n, cfg, seed=3); s2.observe()
        hobs = torch.empty((n, s2.obs_dim)).pin_memory(); hrew = torch.empty(n).pin_memory()
        ht = torch.empty(n, dtype=torch.uint8).pin_memory(); hu = torch.empty(n, dtype=torch.uint8).pin_memory()
        for i in range(3):
            s2.step_host(s2.random_actions(i).cpu().pin_memory(), hobs, hrew, ht, hu)
        s2.close()
    sim.close()
cat = H1v2Sim(n, tasks.cat_config(), seed=3); cat.observe()
for i in range(3):
    cat.cat_step(cat.random_actions(i))
torch.cuda.synchronize()
cat.close()
print("sanitize target done")
