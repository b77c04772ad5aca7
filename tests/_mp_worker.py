"""Worker of tests/test_multiproc.py: one rank of a world-size-2 gloo job on CPU.  Exercises the host-side N>1 logic:
rank-offset env sharding of the Philox key (oracle as the stand-in device) and the runner's gradient all-reduce."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from h1v2_isaac_b200 import shims  # noqa: E402

shims.install()
from rsl_rl.runners import OnPolicyRunner  # noqa: E402

from h1v2_isaac_b200._capi import default_config  # noqa: E402
from h1v2_isaac_b200.env import H1v2ManagerBasedRLEnv  # noqa: E402
from oracle.oracle import Oracle  # noqa: E402


class OracleVecEnv:
    """rsl_rl VecEnv view over the CPU oracle (test stand-in for the CUDA backend; same sharding rule as env.py)."""

    def __init__(self, n):
        rank, world = H1v2ManagerBasedRLEnv._dist_info()
        cfg = default_config()
        cfg.env_id_offset = rank * n  # env.py: envs shard by rank, the Philox key uses the global env id
        self.o = Oracle(cfg, n, seed=42, threads=2)
        self.num_envs, self.num_actions, self.device, self.max_episode_length = n, 12, "cpu", 1000
        self.episode_length_buf = torch.zeros(n, dtype=torch.long)
        self._obs = torch.from_numpy(self.o.observe())

    def get_observations(self):
        return self._obs, {"observations": {"policy": self._obs}}

    def step(self, a):
        obs, rew, term, trunc = self.o.step(a.numpy().astype(np.float32))
        self._obs = torch.from_numpy(obs)
        rank, _ = H1v2ManagerBasedRLEnv._dist_info()
        self.k = getattr(self, "k", 0) + 1
        if rank == 0 and self.k == 2:
            # an episode "finishes" on rank 0 only: the runner's reward / length reduction must not depend on every rank having
            # a finished episode (a collective inside `if len(rewbuffer) > 0` would hang or pair with the wrong collective)
            term = term.copy(); term[0] = True
        return self._obs, torch.from_numpy(rew), torch.from_numpy(term | trunc).long(), {
            "time_outs": torch.from_numpy(trunc), "observations": {"policy": self._obs},
            "log": {"Probe/rank_plus_one": torch.tensor(float(rank + 1))}}  # rank-dependent: the runner must log the mean over ranks


def main():
    out_path = sys.argv[1]
    rank = int(os.environ["RANK"])
    torch.manual_seed(100 + rank)  # different init per rank: the runner must broadcast rank 0's parameters
    env = OracleVecEnv(16)
    first = env.o.get_state(["root_pos", "root_quat", "command"])
    cfg = {"num_steps_per_env": 4, "save_interval": 1000, "empirical_normalization": False,
           "policy": {"init_noise_std": 1.0, "actor_hidden_dims": [32, 32], "critic_hidden_dims": [32, 32], "activation": "elu"},
           "algorithm": {"value_loss_coef": 1.0, "use_clipped_value_loss": True, "clip_param": 0.2, "entropy_coef": 0.0081, "num_learning_epochs": 2,
                         "num_mini_batches": 2, "learning_rate": 1e-3, "schedule": "adaptive", "gamma": 0.99, "lam": 0.95, "desired_kl": 0.01,
                         "max_grad_norm": 1.0}}
    runner = OnPolicyRunner(env, cfg, log_dir=os.path.join(os.path.dirname(out_path), "logs"), device="cpu")  # book-keeping needs a log dir
    runner.learn(2, init_at_random_ep_len=False)
    flat = torch.cat([p.detach().reshape(-1) for p in runner.alg.policy.parameters()])
    json.dump({"rank": rank, "world": runner.gpu_world_size, "param_sum": float(flat.double().sum()), "param_abs": float(flat.double().abs().sum()),
               "lr": float(runner.alg.lr), "mean_len": runner.stats.get("mean_episode_length"), "probe": runner.stats.get("episode/Probe/rank_plus_one"), "root_pos": first["root_pos"].tolist(), "command": first["command"].tolist()}, open(out_path, "w"))


if __name__ == "__main__":
    main()
