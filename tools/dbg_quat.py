import sys, os
sys.path.insert(0, '/root/repo')
import numpy as np, torch
from h1v2_isaac_b200 import _capi
if len(sys.argv) > 1:
    _capi.LIB_PATH = sys.argv[1]
from h1v2_isaac_b200.backend import H1v2Sim
from h1v2_isaac_b200._capi import default_config
cfg = default_config(); cfg.decimation = int(sys.argv[2]) if len(sys.argv) > 2 else 1; cfg.max_delay = min(cfg.max_delay, 2 * cfg.decimation)
n = 64
sim = H1v2Sim(n, cfg, seed=3, diagnostics=True); sim.observe()
a = torch.zeros((n, 12), device='cuda')
for i in range(3):
    s0 = {k: v.cpu().numpy() for k, v in sim.get_state(["root_quat", "root_ang_vel", "root_pos", "joint_pos"]).items()}
    _, _, t, u = sim.step(a)
    s1 = {k: v.cpu().numpy() for k, v in sim.get_state(["root_quat", "root_ang_vel", "root_pos", "pre_reset_qpos"]).items()}
    print(i, "resets", int((t | u).sum()), "quat norm", np.linalg.norm(s1["root_quat"], axis=1)[:4], "\n q0", s0["root_quat"][0], "\n q1", s1["root_quat"][0], "\n pre", s1["pre_reset_qpos"][0][:7], "w", s1["root_ang_vel"][0])
