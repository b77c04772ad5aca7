import sys, time
sys.path.insert(0, '/root/repo')
import torch
from h1v2_isaac_b200 import tasks
tasks.register()
import gymnasium as gym
for n in (4096,):
    env = gym.make(tasks.TASK_ID, cfg=tasks.default_env_cfg(n))
    a = torch.randn((n, 12), device='cuda')
    for _ in range(50): env.step(a)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(500): env.step(a)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    sim = env.sim
    obs = torch.empty((n, 450), device='cuda'); rew = torch.empty(n, device='cuda'); te = torch.empty(n, dtype=torch.uint8, device='cuda'); tr = torch.empty(n, dtype=torch.uint8, device='cuda')
    torch.cuda.synchronize(); t2 = time.perf_counter()
    for _ in range(500): sim.step_into(a, obs, rew, te, tr)
    torch.cuda.synchronize(); t3 = time.perf_counter()
    print(f"n={n}: env.step {(t1-t0)/500*1e3:.3f} ms/step, raw C-ABI step {(t3-t2)/500*1e3:.3f} ms/step")
