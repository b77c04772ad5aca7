import sys
sys.path.insert(0, '/root/repo')
import torch
from h1v2_isaac_b200 import tasks
from h1v2_isaac_b200.backend import H1v2Sim
n = 32768
for label, kw, cat in (("step, no diagnostics row", {}, False), ("step + diagnostics row", {"diagnostics": True}, False), ("cat_step", {}, True)):
    c = tasks.cat_config()
    if not cat: c.cat_enable = 0
    sim = H1v2Sim(n, c, seed=1, **kw); sim.observe()
    acts = [sim.random_actions(i) for i in range(8)]
    f = sim.cat_step if cat else sim.step
    for i in range(20): f(acts[i % 8])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(100): f(acts[i % 8])
    e1.record(); torch.cuda.synchronize()
    print(f"{label}: {e0.elapsed_time(e1)/100:.4f} ms/step")
    sim.close()
