"""ncu target: 30 CaT steps then 30 plain steps of the same CaT config at a given env count (kernel durations side by side)."""
import sys
sys.path.insert(0, '/root/repo')
import torch
from h1v2_isaac_b200 import tasks
from h1v2_isaac_b200.backend import H1v2Sim
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
for label in ("cat", "plain"):
    sim = H1v2Sim(n, tasks.cat_config(), seed=1); sim.observe()
    acts = [sim.random_actions(i) for i in range(8)]
    obs = torch.empty((n, sim.obs_dim), device='cuda'); rew = torch.empty(n, device='cuda'); d = torch.empty(n, device='cuda')
    t = torch.empty(n, dtype=torch.uint8, device='cuda'); u = torch.empty(n, dtype=torch.uint8, device='cuda')
    for i in range(30):
        if label == "cat": sim.cat_step_into(acts[i % 8], obs, rew, d, u)
        else: sim.step_into(acts[i % 8], obs, rew, t, u)
    torch.cuda.synchronize(); sim.close()
print("done")
