"""isaaclab_rl.rsl_rl: cfg classes consumed by C12/agents/rsl_rl_ppo_cfg.py:11-47 and the VecEnv wrapper that
scripts/rsl_rl/train.py:120 puts around the env."""
from __future__ import annotations

import os

import torch

from h1v2_isaac_b200.shims._configclass import MISSING, configclass


@configclass
class RslRlPpoActorCriticCfg:
    class_name: str = "ActorCritic"
    init_noise_std: float = MISSING
    noise_std_type: str = "scalar"
    actor_hidden_dims: list = MISSING
    critic_hidden_dims: list = MISSING
    activation: str = MISSING


@configclass
class RslRlPpoAlgorithmCfg:
    class_name: str = "PPO"
    value_loss_coef: float = MISSING
    use_clipped_value_loss: bool = MISSING
    clip_param: float = MISSING
    entropy_coef: float = MISSING
    num_learning_epochs: int = MISSING
    num_mini_batches: int = MISSING
    learning_rate: float = MISSING
    schedule: str = MISSING
    gamma: float = MISSING
    lam: float = MISSING
    desired_kl: float = MISSING
    max_grad_norm: float = MISSING
    normalize_advantage_per_mini_batch: bool = False
    symmetry_cfg = None
    rnd_cfg = None


@configclass
class RslRlOnPolicyRunnerCfg:
    seed: int = 42
    device: str = "cuda:0"
    num_steps_per_env: int = MISSING
    max_iterations: int = MISSING
    empirical_normalization: bool = MISSING
    policy = MISSING
    algorithm = MISSING
    clip_actions = None
    save_interval: int = MISSING
    experiment_name: str = MISSING
    run_name: str = ""
    logger: str = "tensorboard"
    neptune_project: str = "isaaclab"
    wandb_project: str = "isaaclab"
    resume: bool = False
    load_run: str = ".*"
    load_checkpoint: str = "model_.*.pt"


class RslRlVecEnvWrapper:
    """rsl_rl VecEnv view of a ManagerBasedRLEnv: obs = obs_dict["policy"], dones = terminated | truncated (long),
    extras["observations"] = obs_dict, extras["time_outs"] = truncated (only for infinite-horizon tasks)."""

    def __init__(self, env, clip_actions: float | None = None):
        from isaaclab.envs import DirectRLEnv, ManagerBasedRLEnv
        if not isinstance(env.unwrapped, (ManagerBasedRLEnv, DirectRLEnv)):
            raise ValueError(f"The environment must be inherited from ManagerBasedRLEnv or DirectRLEnv. Environment type: {type(env)}")
        self.env = env
        self.clip_actions = clip_actions
        self.num_envs = self.unwrapped.num_envs
        self.device = self.unwrapped.device
        self.max_episode_length = self.unwrapped.max_episode_length
        self.num_actions = self.unwrapped.action_manager.total_action_dim
        self.num_obs = self.unwrapped.observation_manager.group_obs_dim["policy"][0]
        dims = self.unwrapped.observation_manager.group_obs_dim
        self.num_privileged_obs = dims["critic"][0] if "critic" in dims else 0
        self.env.reset()

    def __str__(self):
        return f"<{type(self).__name__}{self.env}>"

    @property
    def cfg(self):
        return self.unwrapped.cfg

    @property
    def render_mode(self):
        return self.env.render_mode

    @property
    def observation_space(self):
        return self.env.observation_space

    @property
    def action_space(self):
        return self.env.action_space

    @classmethod
    def class_name(cls) -> str:
        return cls.__name__

    @property
    def unwrapped(self):
        return self.env.unwrapped

    def get_observations(self):
        obs_dict = self.unwrapped.observation_manager.compute()
        return obs_dict["policy"], {"observations": obs_dict}

    @property
    def episode_length_buf(self) -> torch.Tensor:
        return self.unwrapped.episode_length_buf

    @episode_length_buf.setter
    def episode_length_buf(self, value: torch.Tensor):
        self.unwrapped.episode_length_buf = value

    def seed(self, seed: int = -1) -> int:
        return self.unwrapped.seed(seed)

    def reset(self):
        obs_dict, _ = self.env.reset()
        return obs_dict["policy"], {"observations": obs_dict}

    def step(self, actions: torch.Tensor):
        if self.clip_actions is not None:
            actions = torch.clamp(actions, -self.clip_actions, self.clip_actions)
        obs_dict, rew, terminated, truncated, extras = self.env.step(actions)
        dones = (terminated | truncated).to(dtype=torch.long)
        obs = obs_dict["policy"]
        extras["observations"] = obs_dict
        if not self.unwrapped.cfg_is_finite_horizon:
            extras["time_outs"] = truncated
        return obs, rew, dones, extras

    def close(self):
        return self.env.close()


def export_policy_as_jit(actor_critic, normalizer, path: str, filename: str = "policy.pt"):
    os.makedirs(path, exist_ok=True)

    class _Exporter(torch.nn.Module):
        def __init__(self, actor, normalizer):
            super().__init__()
            import copy
            self.actor = copy.deepcopy(actor)
            self.normalizer = copy.deepcopy(normalizer) if normalizer is not None else torch.nn.Identity()

        def forward(self, x):
            return self.actor(self.normalizer(x))

    m = _Exporter(actor_critic.actor, normalizer).to("cpu")
    torch.jit.script(m).save(os.path.join(path, filename))


def export_policy_as_onnx(actor_critic, path: str, normalizer=None, filename: str = "policy.onnx", verbose: bool = False):
    os.makedirs(path, exist_ok=True)
    import copy
    actor = copy.deepcopy(actor_critic.actor).to("cpu")
    obs = torch.zeros(1, actor[0].in_features)
    try:
        torch.onnx.export(actor, obs, os.path.join(path, filename), export_params=True, opset_version=11, verbose=verbose,
                          input_names=["obs"], output_names=["actions"], dynamic_axes={})
    except Exception as e:  # onnx may be absent
        print(f"[isaaclab_rl shim] onnx export skipped: {e}")
