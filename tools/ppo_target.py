"""PPO-loop profiling target: a few OnPolicyRunner iterations of the Flat id on the backend (used under ncu on the GPU box)."""
import sys, tempfile
sys.path.insert(0, '/root/repo')
import torch
from h1v2_isaac_b200 import shims, tasks
shims.install(); tasks.register()
import gymnasium as gym
from isaaclab_rl.rsl_rl import RslRlVecEnvWrapper
from rsl_rl.runners import OnPolicyRunner
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
torch.backends.cuda.matmul.allow_tf32 = True
torch.backends.cudnn.allow_tf32 = True
env = gym.make(tasks.TASK_ID, cfg=tasks.default_env_cfg(n))
runner = OnPolicyRunner(RslRlVecEnvWrapper(env), tasks.default_agent_cfg().to_dict(), log_dir=tempfile.mkdtemp(), device="cuda:0")
runner.learn(num_learning_iterations=iters, init_at_random_ep_len=True)
torch.cuda.synchronize()
print("done")
