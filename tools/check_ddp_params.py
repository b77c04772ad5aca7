"""Under torchrun (N >= 2): a few PPO iterations of the runner shim with the learner graph (NCCL all-reduces inside the capture), then the
spread of every parameter across the ranks -- it must be exactly zero (same broadcast start, same averaged gradients, same Adam)."""
import contextlib, os, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
from h1v2_isaac_b200 import shims, tasks
shims.install(); tasks.register()
import gymnasium as gym
from isaaclab_rl.rsl_rl import RslRlVecEnvWrapper
from rsl_rl.runners import OnPolicyRunner
with contextlib.redirect_stdout(sys.stderr):
    torch.manual_seed(42 + rank)
    env = gym.make("Isaac-Velocity-Flat-H12_12dof-v0", cfg=tasks.default_env_cfg(1024, device=f"cuda:{local}"))
    runner = OnPolicyRunner(RslRlVecEnvWrapper(env), tasks.default_agent_cfg().to_dict(), log_dir=tempfile.mkdtemp(), device=f"cuda:{local}")
    runner.learn(num_learning_iterations=5, init_at_random_ep_len=True)
flat = torch.cat([p.detach().flatten() for p in runner.alg.policy.parameters()])
hi, lo = flat.clone(), flat.clone()
dist.all_reduce(hi, op=dist.ReduceOp.MAX); dist.all_reduce(lo, op=dist.ReduceOp.MIN)
moved = float((flat - torch.cat([p.detach().flatten() for p in runner.alg.policy.parameters()])).abs().max())
if rank == 0:
    print(f"ranks {world}: graph learner {runner.alg._graph_update_ok()}, {flat.numel()} parameters, max spread across ranks {float((hi - lo).abs().max()):.3e}, |params| mean {float(flat.abs().mean()):.4f}, lr {runner.alg.lr:.2e}", flush=True)
runner.release_graphs(); env.close()
dist.destroy_process_group()
