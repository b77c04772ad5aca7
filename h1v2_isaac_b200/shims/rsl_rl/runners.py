"""OnPolicyRunner / PPO / ActorCritic with rsl-rl-lib 2.3.3 semantics:
24-step rollouts, GAE(gamma, lam) with time-out bootstrapping, clipped surrogate + clipped value loss, entropy bonus,
adaptive-KL learning rate, num_learning_epochs x num_mini_batches updates, grad-norm clipping.  Multi-GPU (WORLD_SIZE>1):
parameters broadcast from rank 0 once, gradients all-reduced (mean) every mini-batch, KL all-reduced for the LR schedule,
rollout statistics reduced to rank 0 once per iteration (north star)."""
from __future__ import annotations

import os
import statistics
import time
from collections import deque

import torch
import torch.distributed as dist
import torch.nn as nn


def _act(name: str):
    return {"elu": nn.ELU, "selu": nn.SELU, "relu": nn.ReLU, "lrelu": nn.LeakyReLU, "tanh": nn.Tanh, "sigmoid": nn.Sigmoid}[name]()


class _AlignedLinear(nn.Linear):
    """nn.Linear whose matmul CAN see an input width that is a multiple of 8 on CUDA (H1V2_KPAD=1): same parameters and state_dict as
    nn.Linear, the input and the weight are zero-padded on the fly (the extra products are exact zeros).  The policy's first layer has
    450 inputs (45 x 10 history): with K % 4 != 0 cuBLAS falls back to the unaligned TF32 kernels (cutlass_80 ... align1), which were
    22 % of the PPO loop's GPU time at 4096 envs (profiles/r3_ppo_launches_summary.txt); padded to 456 the aligned sm_100 kernels run
    (learn 42 -> 34 ms per iteration at 4096 envs).  OFF by default: with it the reference's unmodified train.py learns the Flat id three
    times more slowly (mean episode length 116 instead of 939 after 300 iterations, same kernel, same seed; profiles/r4_notes.md section 8)
    -- the aligned TF32 kernels evidently do not give the first layer the gradient the unaligned ones do; not understood yet."""

    def forward(self, x):
        pad = (-self.in_features) % 8
        if pad and x.is_cuda and os.environ.get("H1V2_KPAD") == "1":
            return nn.functional.linear(nn.functional.pad(x, (0, pad)), nn.functional.pad(self.weight, (0, pad)), self.bias)
        return nn.functional.linear(x, self.weight, self.bias)


def _mlp(i, hidden, o, act):
    layers, d = [], i
    for h in hidden:
        layers += [(_AlignedLinear if d % 8 else nn.Linear)(d, h), _act(act)]
        d = h
    layers.append(nn.Linear(d, o))
    return nn.Sequential(*layers)


class ActorCritic(nn.Module):
    is_recurrent = False

    def __init__(self, num_actor_obs, num_critic_obs, num_actions, actor_hidden_dims=(256, 256, 256), critic_hidden_dims=(256, 256, 256),
                 activation="elu", init_noise_std=1.0, noise_std_type="scalar", **kwargs):
        super().__init__()
        self.actor = _mlp(num_actor_obs, actor_hidden_dims, num_actions, activation)
        self.critic = _mlp(num_critic_obs, critic_hidden_dims, 1, activation)
        self.std = nn.Parameter(init_noise_std * torch.ones(num_actions))
        self.distribution = None
        torch.distributions.Normal.set_default_validate_args(False)

    def reset(self, dones=None):
        pass

    def update_distribution(self, obs):
        mean = self.actor(obs)
        self.distribution = torch.distributions.Normal(mean, self.std.expand_as(mean))

    @property
    def action_mean(self):
        return self.distribution.mean

    @property
    def action_std(self):
        return self.distribution.stddev

    @property
    def entropy(self):
        return self.distribution.entropy().sum(dim=-1)

    def act(self, obs, **kw):
        self.update_distribution(obs)
        return self.distribution.sample()

    def get_actions_log_prob(self, actions):
        return self.distribution.log_prob(actions).sum(dim=-1)

    def act_inference(self, obs):
        return self.actor(obs)

    def evaluate(self, critic_obs, **kw):
        return self.critic(critic_obs)


class EmpiricalNormalization(nn.Module):
    def __init__(self, shape, eps=1e-2, until=None):
        super().__init__()
        self.eps, self.until = eps, until
        self.register_buffer("_mean", torch.zeros(shape).unsqueeze(0))
        self.register_buffer("_var", torch.ones(shape).unsqueeze(0))
        self.register_buffer("_std", torch.ones(shape).unsqueeze(0))
        self.register_buffer("count", torch.tensor(0, dtype=torch.long))

    def forward(self, x):
        if self.training:
            self.update(x)
        return (x - self._mean) / (self._std + self.eps)

    @torch.jit.unused
    def update(self, x):
        if self.until is not None and self.count >= self.until:
            return
        n = x.shape[0]
        self.count += n
        rate = n / self.count
        var_x, mean_x = torch.var(x, dim=0, unbiased=False, keepdim=True), torch.mean(x, dim=0, keepdim=True)
        delta = mean_x - self._mean
        self._mean += rate * delta
        self._var += rate * (var_x - self._var + delta * (mean_x - self._mean))
        self._std = torch.sqrt(self._var)


class RolloutStorage:
    def __init__(self, num_envs, T, obs_shape, critic_shape, act_shape, device):
        z = lambda *s: torch.zeros(T, num_envs, *s, device=device)  # noqa: E731
        self.obs, self.critic_obs, self.actions = z(*obs_shape), z(*critic_shape), z(*act_shape)
        self.rewards, self.dones, self.values, self.logp = z(1), z(1).byte(), z(1), z(1)
        self.mu, self.sigma = z(*act_shape), z(*act_shape)
        self.returns, self.adv = z(1), z(1)
        self.T, self.num_envs, self.step = T, num_envs, 0

    def add(self, obs, critic_obs, actions, rewards, dones, values, logp, mu, sigma):
        t = self.step
        self.obs[t].copy_(obs); self.critic_obs[t].copy_(critic_obs); self.actions[t].copy_(actions)
        self.rewards[t].copy_(rewards.view(-1, 1)); self.dones[t].copy_(dones.view(-1, 1)); self.values[t].copy_(values)
        self.logp[t].copy_(logp.view(-1, 1)); self.mu[t].copy_(mu); self.sigma[t].copy_(sigma)
        self.step += 1

    def clear(self):
        self.step = 0

    def compute_returns(self, last_values, gamma, lam, normalize=True):
        adv = 0
        for t in reversed(range(self.T)):
            nv = last_values if t == self.T - 1 else self.values[t + 1]
            nt = 1.0 - self.dones[t].float()
            delta = self.rewards[t] + nt * gamma * nv - self.values[t]
            adv = delta + nt * gamma * lam * adv
            self.returns[t] = adv + self.values[t]
        self.adv.copy_(self.returns - self.values)  # in place: the captured learner graph reads this tensor's storage
        if normalize:
            self.adv.copy_((self.adv - self.adv.mean()) / (self.adv.std() + 1e-8))

    def mini_batches(self, num_mini_batches, num_epochs):
        B = self.T * self.num_envs
        mb = B // num_mini_batches
        flat = lambda x: x.flatten(0, 1)  # noqa: E731
        obs, cobs, act, val, ret, logp, adv, mu, sig = map(flat, (self.obs, self.critic_obs, self.actions, self.values, self.returns, self.logp,
                                                                  self.adv, self.mu, self.sigma))
        for _ in range(num_epochs):
            idx = torch.randperm(num_mini_batches * mb, device=obs.device)
            for i in range(num_mini_batches):
                b = idx[i * mb:(i + 1) * mb]
                yield obs[b], cobs[b], act[b], val[b], adv[b], ret[b], logp[b], mu[b], sig[b]


class PPO:
    def __init__(self, policy, num_learning_epochs=1, num_mini_batches=1, clip_param=0.2, gamma=0.998, lam=0.95, value_loss_coef=1.0,
                 entropy_coef=0.0, learning_rate=1e-3, max_grad_norm=1.0, use_clipped_value_loss=True, schedule="fixed", desired_kl=0.01,
                 device="cpu", normalize_advantage_per_mini_batch=False, multi_gpu_cfg=None, **kwargs):
        self.policy, self.device = policy.to(device), device
        # On CUDA the adaptive learning rate lives in a 0-d device tensor (Adam(capturable=True) reads it on the device), so the
        # KL test of every mini-batch costs no host synchronisation; self.lr (float, for logs / checkpoints) is refreshed once per update().
        self._lr_on_device = str(device).startswith("cuda")
        if self._lr_on_device:
            self.lr_t = torch.tensor(float(learning_rate), device=device)
            self.opt = torch.optim.Adam(self.policy.parameters(), lr=self.lr_t, capturable=True)
        else:
            self.opt = torch.optim.Adam(self.policy.parameters(), lr=learning_rate)
        self.epochs, self.nmb, self.clip, self.gamma, self.lam = num_learning_epochs, num_mini_batches, clip_param, gamma, lam
        self.vcoef, self.ecoef, self.lr, self.max_grad_norm = value_loss_coef, entropy_coef, learning_rate, max_grad_norm
        self.clip_v, self.schedule, self.desired_kl = use_clipped_value_loss, schedule, desired_kl
        self.multi_gpu = multi_gpu_cfg
        self.storage = None
        self._tr = {}

    def init_storage(self, num_envs, T, obs_shape, critic_shape, act_shape):
        self.storage = RolloutStorage(num_envs, T, obs_shape, critic_shape, act_shape, self.device)

    def act(self, obs, critic_obs):
        a = self.policy.act(obs).detach()
        self._tr = dict(obs=obs, critic_obs=critic_obs, actions=a, values=self.policy.evaluate(critic_obs).detach(),
                        logp=self.policy.get_actions_log_prob(a).detach(), mu=self.policy.action_mean.detach(), sigma=self.policy.action_std.detach())
        return a

    def process_env_step(self, rewards, dones, infos):
        r = rewards.clone()
        if "time_outs" in infos:  # bootstrap on time-outs
            r += self.gamma * (self._tr["values"].squeeze(1) * infos["time_outs"].to(self.device).float())
        t = self._tr
        self.storage.add(t["obs"], t["critic_obs"], t["actions"], r, dones, t["values"], t["logp"], t["mu"], t["sigma"])
        self._tr = {}

    def compute_returns(self, last_critic_obs):
        self.storage.compute_returns(self.policy.evaluate(last_critic_obs).detach(), self.gamma, self.lam)

    def broadcast_parameters(self):
        for p in self.policy.parameters():
            dist.broadcast(p.data, src=0)

    def reduce_parameters(self):
        grads = [p.grad.view(-1) for p in self.policy.parameters() if p.grad is not None]
        flat = torch.cat(grads)
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
        flat /= self.multi_gpu["world_size"]
        off = 0
        for p in self.policy.parameters():
            if p.grad is not None:
                n = p.numel()
                p.grad.data.copy_(flat[off:off + n].view_as(p.grad.data))
                off += n

    def _minibatch_step(self, obs, cobs, act, val, adv, ret, old_logp, old_mu, old_sigma):
        """One PPO mini-batch update (rsl-rl-lib 2.3.3 PPO.update body); returns the three detached loss statistics."""
        self.policy.update_distribution(obs)  # upstream calls policy.act(obs) and discards the sample; torch.normal's host-side check of std cannot be captured
        logp = self.policy.get_actions_log_prob(act)
        value = self.policy.evaluate(cobs)
        mu, sigma, ent = self.policy.action_mean, self.policy.action_std, self.policy.entropy
        if self.desired_kl is not None and self.schedule == "adaptive":
            with torch.inference_mode():
                kl = torch.sum(torch.log(sigma / old_sigma + 1e-5) + (old_sigma.square() + (old_mu - mu).square()) / (2.0 * sigma.square()) - 0.5, dim=-1)
                kl_mean = kl.mean()
                if self.multi_gpu:
                    dist.all_reduce(kl_mean, op=dist.ReduceOp.SUM)
                    kl_mean /= self.multi_gpu["world_size"]
                if self._lr_on_device:  # same rule as below, evaluated on the device
                    down = (self.lr_t / 1.5).clamp(min=1e-5)
                    up = (self.lr_t * 1.5).clamp(max=1e-2)
                    self.lr_t.copy_(torch.where(kl_mean > self.desired_kl * 2.0, down,
                                                torch.where((kl_mean > 0.0) & (kl_mean < self.desired_kl / 2.0), up, self.lr_t)))
                else:
                    if kl_mean > self.desired_kl * 2.0:
                        self.lr = max(1e-5, self.lr / 1.5)
                    elif 0.0 < kl_mean < self.desired_kl / 2.0:
                        self.lr = min(1e-2, self.lr * 1.5)
                    for g in self.opt.param_groups:
                        g["lr"] = self.lr
        ratio = torch.exp(logp - old_logp.squeeze(1))
        a = adv.squeeze(1)
        surrogate = torch.max(-a * ratio, -a * torch.clamp(ratio, 1.0 - self.clip, 1.0 + self.clip)).mean()
        if self.clip_v:
            vc = val + (value - val).clamp(-self.clip, self.clip)
            vloss = torch.max((value - ret).square(), (vc - ret).square()).mean()
        else:
            vloss = (ret - value).square().mean()
        loss = surrogate + self.vcoef * vloss - self.ecoef * ent.mean()
        self.opt.zero_grad()
        loss.backward()
        if self.multi_gpu:
            self.reduce_parameters()
        nn.utils.clip_grad_norm_(self.policy.parameters(), self.max_grad_norm)
        self.opt.step()
        return vloss.detach(), surrogate.detach(), ent.mean().detach()

    def _graph_update_ok(self) -> bool:
        """The whole update (epochs x mini-batches of forward / backward / gradient clip / Adam, adaptive learning rate on the device) as
        one CUDA graph, on CUDA by default (H1V2_GRAPH_LEARNER=0 turns it off).  On several GPUs the NCCL gradient / KL all-reduces are
        inside the capture (2 GPUs, 4096 envs each: learn 76 -> 38 ms per iteration, profiles/r4_notes.md); the runner drops its graphs at
        interpreter exit, before the process group goes away (OnPolicyRunner.release_graphs)."""
        flag = os.environ.get("H1V2_GRAPH_LEARNER", "")
        return self._lr_on_device and flag != "0"

    def update(self):
        if self._graph_update_ok():
            return self._update_graphed()
        mv = ms = me = torch.zeros((), device=self.device)
        n = 0
        for batch in self.storage.mini_batches(self.nmb, self.epochs):
            v, s_, e = self._minibatch_step(*batch)
            mv = mv + v; ms = ms + s_; me = me + e; n += 1
        self.storage.clear()
        if self._lr_on_device:
            self.lr = float(self.lr_t)  # the one host read of the update
        return {"value_function": float(mv) / n, "surrogate": float(ms) / n, "entropy": float(me) / n}

    def _update_graphed(self):
        st = self.storage
        B = st.T * st.num_envs
        mb = B // self.nmb
        nb = self.nmb * self.epochs
        if getattr(self, "_gu", None) is None:
            self._gu = {"idx": torch.zeros((nb, mb), dtype=torch.long, device=self.device), "stats": torch.zeros(3, device=self.device),
                        "graph": None, "calls": 0, "stream": torch.cuda.Stream(device=self.device)}
        G = self._gu
        for ep in range(self.epochs):  # the same index stream as RolloutStorage.mini_batches: one permutation per epoch
            G["idx"][ep * self.nmb:(ep + 1) * self.nmb].copy_(torch.randperm(self.nmb * mb, device=self.device).view(self.nmb, mb))

        def body():
            flat = lambda x: x.flatten(0, 1)  # noqa: E731
            obs, cobs, act, val, ret, logp, adv, mu, sig = map(flat, (st.obs, st.critic_obs, st.actions, st.values, st.returns, st.logp, st.adv, st.mu, st.sigma))
            G["stats"].zero_()
            for i in range(nb):
                b = G["idx"][i]
                v, s_, e = self._minibatch_step(obs[b], cobs[b], act[b], val[b], adv[b], ret[b], logp[b], mu[b], sig[b])
                G["stats"] += torch.stack((v, s_, e))

        cur = torch.cuda.current_stream(self.device)
        if G["calls"] == 0:  # first update: eagerly on a side stream (the warm-up torch asks for before capturing a backward pass)
            G["stream"].wait_stream(cur)
            with torch.cuda.stream(G["stream"]):
                body()
            cur.wait_stream(G["stream"])
        else:
            if G["graph"] is None:
                torch.cuda.synchronize(self.device)
                self.opt.zero_grad(set_to_none=True)
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, stream=G["stream"]):
                    body()
                G["graph"] = g
            G["graph"].replay()
        G["calls"] += 1
        st.clear()
        out = (G["stats"] / nb).tolist()  # the one host read of the update
        self.lr = float(self.lr_t)
        return {"value_function": out[0], "surrogate": out[1], "entropy": out[2]}


class _GraphRollout:
    """The whole num_steps_per_env rollout of one iteration as ONE CUDA graph: per step the actor / critic
    forward, the action sample, its log-probability, the fused env step (h1v2_step through H1v2Sim.step_into, which writes the
    next observation straight into the rollout storage), the time-out bootstrap and the storage writes -- the same arithmetic as
    PPO.act / env.step / PPO.process_env_step above, with the per-step Python, the wrapper dictionaries and the logging
    synchronisations (nonzero / tolist) replaced by device-side accumulators read once per iteration.  Used when the env is this
    package's CUDA backend without privileged observations or empirical normalisation (a reward-weight curriculum is applied between
    iterations and triggers a re-capture when it changes a weight); otherwise (and
    with H1V2_GRAPH_ROLLOUT=0) OnPolicyRunner.learn runs the eager loop.  The first iteration runs the same body eagerly (warm-up
    of cuBLAS workspaces and kernel attributes), the graph is captured after it and replayed from the second iteration on."""

    def __init__(self, runner):
        self.r = runner
        env, alg = runner.env, runner.alg
        base = env.unwrapped
        self.base, self.sim, self.alg, self.T = base, base.sim, alg, runner.num_steps_per_env
        self.clip = getattr(env, "clip_actions", None)
        self.bootstrap = not getattr(base, "cfg_is_finite_horizon", False)
        dev, n = self.sim.device, self.sim.num_envs
        self.obs_carry = torch.zeros((n, self.sim.obs_dim), device=dev)
        self.rew = torch.zeros(n, device=dev)
        self.term = torch.zeros(n, dtype=torch.uint8, device=dev)
        self.trunc = torch.zeros(n, dtype=torch.uint8, device=dev)
        self.cur_rew, self.cur_len = torch.zeros(n, device=dev), torch.zeros(n, device=dev)
        self.ep_stats = torch.zeros(3, device=dev)  # sum of returns, sum of lengths, count of the episodes that ended in this iteration
        base._log_dict()  # builds the key list / gather index of extras["log"]
        self.log_keys, self.log_index = list(base._log_keys), base._log_index
        self.log_sum = torch.zeros(len(self.log_keys), device=dev)
        # Rough id: Curriculum/terrain_levels (mdp.terrain_levels_vel returns the mean level) comes from the kernel's own two-float log
        self.terrain_log = self.sim.terrain_log_buf if getattr(base.kernel_cfg, "terrain_curriculum", 0) else None
        self.terrain_sum = torch.zeros((), device=dev)
        self.graph = None
        self.iterations = 0

    @staticmethod
    def supported(runner) -> bool:
        if os.environ.get("H1V2_GRAPH_ROLLOUT", "1") == "0" or not str(runner.device).startswith("cuda"):
            return False
        base = getattr(runner.env, "unwrapped", None)
        sim = getattr(base, "sim", None)
        return (sim is not None and type(sim).__name__ == "H1v2Sim" and hasattr(sim, "step_into") and not runner.empirical_normalization
                and runner.privileged is None and not getattr(sim.cfg, "cat_enable", 0)
                and torch.device(runner.device) == sim.device and type(runner.alg.policy).__name__ == "ActorCritic")

    def _body(self):
        st, pol, alg, sim = self.alg.storage, self.alg.policy, self.alg, self.sim
        self.ep_stats.zero_(); self.log_sum.zero_(); self.terrain_sum.zero_()
        st.obs[0].copy_(self.obs_carry)
        for t in range(self.T):
            obs = st.obs[t]
            mean = pol.actor(obs)
            value = pol.critic(obs)
            std = pol.std.expand_as(mean)
            act = mean + std * torch.randn_like(mean)  # = Normal(mean, std).sample() without torch.normal's host-side check of std (a sync: illegal in capture)
            logp = (-((act - mean) ** 2) / (2 * std ** 2) - std.log() - 0.9189385332046727).sum(dim=-1)  # Normal(mean, std).log_prob(act).sum(-1)
            st.critic_obs[t].copy_(obs); st.actions[t].copy_(act); st.values[t].copy_(value); st.logp[t].copy_(logp.view(-1, 1))
            st.mu[t].copy_(mean); st.sigma[t].copy_(std)
            a_env = act if self.clip is None else torch.clamp(act, -self.clip, self.clip)
            nxt = st.obs[t + 1] if t + 1 < self.T else self.obs_carry
            sim.step_into(a_env.contiguous(), nxt, self.rew, self.term, self.trunc)
            done = (self.term | self.trunc)
            r = self.rew + alg.gamma * value.squeeze(1) * self.trunc.float() if self.bootstrap else self.rew
            st.rewards[t].copy_(r.view(-1, 1)); st.dones[t].copy_(done.view(-1, 1))
            # logging accumulators (the eager loop's cur_reward_sum / cur_episode_length / ep_infos)
            d = done.float()
            self.cur_rew += self.rew; self.cur_len += 1.0
            self.ep_stats[0] += (self.cur_rew * d).sum(); self.ep_stats[1] += (self.cur_len * d).sum(); self.ep_stats[2] += d.sum()
            self.cur_rew *= 1.0 - d; self.cur_len *= 1.0 - d
            self.log_sum += sim.log_buf[self.log_index]
            if self.terrain_log is not None:
                self.terrain_sum += self.terrain_log[1]
        st.step = self.T

    def run(self, obs):
        """One rollout starting from `obs` (only read in the first iteration; afterwards the carried observation is the env's own)."""
        if self.iterations == 0:
            self._body()
        else:
            if self.graph is None:
                torch.cuda.synchronize()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    self._body()
                self.graph = g
            self.graph.replay()
            self.alg.storage.step = self.T
        self.iterations += 1
        base = self.base
        base.common_step_counter += self.T
        if getattr(base, "_curriculum", None):
            # mdp.modify_reward_weight terms that became due during this rollout (env.step applies them step by step; here once per
            # iteration): the captured step kernel carries the reward weights of capture time, so a changed weight means a new capture
            before = list(self.sim.cfg.rew_weight)
            base._apply_curriculum()
            if list(self.sim.cfg.rew_weight) != before:
                self.graph = None
        base.obs_buf = {"policy": self.obs_carry}
        base.reward_buf, base.reset_terminated, base.reset_time_outs = self.rew, self.term.bool(), self.trunc.bool()
        base.reset_buf = base.reset_terminated | base.reset_time_outs
        return self.obs_carry

    def stats(self):
        """(episode statistics [sum_return, sum_length, count], {log key: mean over the steps}) -- the one host read of the rollout."""
        ep = self.ep_stats.tolist()
        lg = (self.log_sum / self.T).tolist()
        out = dict(zip(self.log_keys, lg))
        if self.terrain_log is not None:
            out["Curriculum/terrain_levels"] = float(self.terrain_sum) / self.T
        return ep, out


class OnPolicyRunner:
    def __init__(self, env, train_cfg: dict, log_dir: str | None = None, device="cpu"):
        self.cfg, self.alg_cfg, self.policy_cfg = train_cfg, dict(train_cfg["algorithm"]), dict(train_cfg["policy"])
        self.device, self.env, self.log_dir = device, env, log_dir
        self._configure_multi_gpu()
        obs, extras = self.env.get_observations()
        num_obs = obs.shape[1]
        num_critic = extras["observations"]["critic"].shape[1] if "critic" in extras["observations"] else num_obs
        self.privileged = "critic" if "critic" in extras["observations"] else None
        self.policy_cfg.pop("class_name", None); self.alg_cfg.pop("class_name", None)
        for k in ("symmetry_cfg", "rnd_cfg"):
            self.alg_cfg.pop(k, None)
        policy = ActorCritic(num_obs, num_critic, self.env.num_actions, **self.policy_cfg).to(self.device)
        self.alg = PPO(policy, device=self.device, multi_gpu_cfg=self.multi_gpu_cfg, **self.alg_cfg)
        if self.is_distributed:  # captured graphs hold NCCL work: they must be gone before the communicator is torn down at exit
            import atexit
            import weakref
            ref = weakref.ref(self)
            atexit.register(lambda: ref() is not None and ref().release_graphs())
        self.num_steps_per_env, self.save_interval = self.cfg["num_steps_per_env"], self.cfg["save_interval"]
        self.empirical_normalization = self.cfg.get("empirical_normalization", False)
        if self.empirical_normalization:
            self.obs_normalizer = EmpiricalNormalization([num_obs], until=1.0e8).to(self.device)
            self.critic_obs_normalizer = EmpiricalNormalization([num_critic], until=1.0e8).to(self.device)
        else:
            self.obs_normalizer = self.critic_obs_normalizer = nn.Identity().to(self.device)
        self.alg.init_storage(self.env.num_envs, self.num_steps_per_env, [num_obs], [num_critic], [self.env.num_actions])
        self.disable_logs = self.is_distributed and self.gpu_global_rank != 0
        self.writer, self.tot_timesteps, self.tot_time, self.current_learning_iteration, self.git_status_repos = None, 0, 0.0, 0, []
        self.stats = {}

    def _configure_multi_gpu(self):
        self.gpu_world_size = int(os.getenv("WORLD_SIZE", "1"))
        self.is_distributed = self.gpu_world_size > 1
        if not self.is_distributed:
            self.gpu_local_rank = self.gpu_global_rank = 0
            self.multi_gpu_cfg = None
            return
        self.gpu_local_rank, self.gpu_global_rank = int(os.getenv("LOCAL_RANK", "0")), int(os.getenv("RANK", "0"))
        self.multi_gpu_cfg = {"global_rank": self.gpu_global_rank, "local_rank": self.gpu_local_rank, "world_size": self.gpu_world_size}
        if str(self.device).startswith("cuda") and self.device != f"cuda:{self.gpu_local_rank}":
            raise ValueError(f"Device '{self.device}' does not match expected device for local rank '{self.gpu_local_rank}'.")
        if not dist.is_initialized():
            dist.init_process_group(backend="nccl" if str(self.device).startswith("cuda") else "gloo", rank=self.gpu_global_rank, world_size=self.gpu_world_size)
        if str(self.device).startswith("cuda"):
            torch.cuda.set_device(self.gpu_local_rank)

    def add_git_repo_to_log(self, repo_file_path):
        self.git_status_repos.append(repo_file_path)

    def _init_writer(self):
        if self.log_dir is None or self.disable_logs or self.writer is not None:
            return
        os.makedirs(self.log_dir, exist_ok=True)
        try:
            from torch.utils.tensorboard import SummaryWriter
            self.writer = SummaryWriter(log_dir=self.log_dir, flush_secs=10)
        except Exception:
            self.writer = None

    def learn(self, num_learning_iterations: int, init_at_random_ep_len: bool = False):
        self._init_writer()
        if init_at_random_ep_len:
            self.env.episode_length_buf = torch.randint_like(self.env.episode_length_buf, high=int(self.env.max_episode_length))
        obs, extras = self.env.get_observations()
        critic_obs = extras["observations"].get(self.privileged, obs) if self.privileged else obs
        obs, critic_obs = obs.to(self.device), critic_obs.to(self.device)
        self.train_mode()
        ep_infos, rewbuffer, lenbuffer = [], deque(maxlen=100), deque(maxlen=100)
        cur_rew = torch.zeros(self.env.num_envs, device=self.device)
        cur_len = torch.zeros(self.env.num_envs, device=self.device)
        if self.is_distributed:
            self.alg.broadcast_parameters()
        start = self.current_learning_iteration
        tot = start + num_learning_iterations
        fast = getattr(self, "_fast", None)  # the captured rollout graph survives across learn() calls
        if fast is None and _GraphRollout.supported(self):
            fast = self._fast = _GraphRollout(self)
        if fast is not None:
            fast.obs_carry.copy_(obs)  # learn() starts from a freshly computed observation (env.get_observations above)
        self.graph_rollout = fast is not None
        self.iteration_times = []  # (collection_s, learn_s) of every iteration of this call
        for it in range(start, tot):
            t0 = time.time()
            if fast is not None:
                with torch.inference_mode():
                    obs = critic_obs = fast.run(obs)
                    ep, log_means = fast.stats()  # synchronises: the collection time below is that of the finished rollout
                    if self.log_dir is not None:
                        ep_infos.append(log_means)
                        if ep[2] > 0:  # mean return / length of the episodes that ended in this iteration
                            rewbuffer.clear(); lenbuffer.clear()
                            rewbuffer.append(ep[0] / ep[2]); lenbuffer.append(ep[1] / ep[2])
                    t1 = time.time()
                    self.alg.compute_returns(critic_obs)
            else:
              with torch.inference_mode():
                for _ in range(self.num_steps_per_env):
                    actions = self.alg.act(obs, critic_obs)
                    obs, rewards, dones, infos = self.env.step(actions.to(self.env.device))
                    obs, rewards, dones = obs.to(self.device), rewards.to(self.device), dones.to(self.device)
                    obs = self.obs_normalizer(obs)
                    critic_obs = self.critic_obs_normalizer(infos["observations"][self.privileged].to(self.device)) if self.privileged else obs
                    self.alg.process_env_step(rewards, dones, infos)
                    if self.log_dir is not None:
                        if "episode" in infos:
                            ep_infos.append(infos["episode"])
                        elif "log" in infos:
                            ep_infos.append(infos["log"])
                        cur_rew += rewards
                        cur_len += 1
                        new_ids = (dones > 0).nonzero(as_tuple=False)
                        rewbuffer.extend(cur_rew[new_ids][:, 0].cpu().numpy().tolist())
                        lenbuffer.extend(cur_len[new_ids][:, 0].cpu().numpy().tolist())
                        cur_rew[new_ids] = 0
                        cur_len[new_ids] = 0
                t1 = time.time()
                self.alg.compute_returns(critic_obs)
            loss = self.alg.update()
            t2 = time.time()
            self.current_learning_iteration = it
            self.iteration_times.append((t1 - t0, t2 - t1))
            self._log(it, tot, t1 - t0, t2 - t1, loss, ep_infos, rewbuffer, lenbuffer)
            ep_infos.clear()
            if self.log_dir is not None and not self.disable_logs and it % self.save_interval == 0:
                self.save(os.path.join(self.log_dir, f"model_{it}.pt"))
        if self.log_dir is not None and not self.disable_logs:
            self.save(os.path.join(self.log_dir, f"model_{self.current_learning_iteration}.pt"))

    def release_graphs(self):
        """Drop the captured rollout / update graphs.  With H1V2_GRAPH_LEARNER=1 on several GPUs the update graph carries the NCCL gradient
        all-reduces: it has to be gone before the process group is destroyed (a communicator with live captured work does not tear down)."""
        import gc
        self._fast = None
        if getattr(self.alg, "_gu", None) is not None:
            self.alg._gu = None
        gc.collect()
        if torch.cuda.is_available():
            torch.cuda.synchronize()

    def _log(self, it, tot, collection_time, learn_time, loss, ep_infos, rewbuffer, lenbuffer):
        steps = self.num_steps_per_env * self.env.num_envs * self.gpu_world_size
        self.tot_timesteps += steps
        self.tot_time += collection_time + learn_time
        fps = int(steps / (collection_time + learn_time))
        stats = {"iteration": it, "fps": fps, "collection_time": collection_time, "learn_time": learn_time, **{f"loss/{k}": v for k, v in loss.items()},
                 "mean_noise_std": self.alg.policy.action_std.mean().item() if self.alg.policy.distribution is not None else float("nan"),
                 "lr": self.alg.lr}
        have = len(rewbuffer) > 0
        mr, ml = (statistics.mean(rewbuffer), statistics.mean(lenbuffer)) if have else (0.0, 0.0)
        if self.is_distributed:
            # rollout statistics reduced over ranks once per iteration (north star).  EVERY rank takes part, whether or not an
            # episode has finished on it yet: a rank with an empty buffer contributes zeros with weight 0 (a collective inside
            # `if len(rewbuffer) > 0` would pair with the next collective of a rank that skipped it)
            t = torch.tensor([mr, ml, 1.0] if have else [0.0, 0.0, 0.0], device=self.device)
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
            have = float(t[2]) > 0
            if have:
                mr, ml = (t[0] / t[2]).item(), (t[1] / t[2]).item()
        if have:
            stats["mean_reward"], stats["mean_episode_length"] = mr, ml
        ep = {}
        if ep_infos:
            for key in ep_infos[0]:
                vals = [torch.as_tensor(e[key], dtype=torch.float32, device=self.device).reshape(-1) for e in ep_infos if key in e]
                if vals:
                    ep[key] = torch.cat(vals).mean().item()
        if self.is_distributed:  # the packed extras["log"] means, one small all-reduce per iteration (every rank logs the job's mean)
            keys = sorted(ep)
            n_keys = torch.tensor([len(keys), -len(keys)], device=self.device)
            dist.all_reduce(n_keys, op=dist.ReduceOp.MAX)
            if len(keys) > 0 and int(n_keys[0]) == len(keys) == -int(n_keys[1]):  # same key set on every rank (always, with this env)
                t = torch.tensor([ep[k] for k in keys], dtype=torch.float32, device=self.device)
                dist.all_reduce(t, op=dist.ReduceOp.SUM)
                ep = dict(zip(keys, (t / self.gpu_world_size).tolist()))
        self.stats = {**stats, **{f"episode/{k}": v for k, v in ep.items()}}
        if self.disable_logs:
            return
        if self.writer is not None:
            for k, v in self.stats.items():
                if isinstance(v, (int, float)):
                    self.writer.add_scalar(k, v, it)
        terrain = self.stats.get("episode/Curriculum/terrain_levels")
        print(f"[rsl_rl shim] it {it + 1}/{tot}  steps/s {fps}  collect {collection_time:.3f}s learn {learn_time:.3f}s  "
              f"reward {stats.get('mean_reward', float('nan')):.3f}  len {stats.get('mean_episode_length', float('nan')):.1f}  "
              f"vloss {loss['value_function']:.4f}  lr {self.alg.lr:.2e}" + (f"  terrain level {terrain:.2f}" if terrain is not None else ""), flush=True)

    def save(self, path: str, infos=None):
        os.makedirs(os.path.dirname(path) or ".", exist_ok=True)
        torch.save({"model_state_dict": self.alg.policy.state_dict(), "optimizer_state_dict": self.alg.opt.state_dict(),
                    "iter": self.current_learning_iteration, "infos": infos,
                    **({"obs_norm_state_dict": self.obs_normalizer.state_dict(), "critic_obs_norm_state_dict": self.critic_obs_normalizer.state_dict()}
                       if self.empirical_normalization else {})}, path)

    def load(self, path: str, load_optimizer: bool = True):
        d = torch.load(path, weights_only=False, map_location=self.device)
        self.alg.policy.load_state_dict(d["model_state_dict"])
        if self.empirical_normalization and "obs_norm_state_dict" in d:
            self.obs_normalizer.load_state_dict(d["obs_norm_state_dict"])
            self.critic_obs_normalizer.load_state_dict(d["critic_obs_norm_state_dict"])
        if load_optimizer:
            self.alg.opt.load_state_dict(d["optimizer_state_dict"])
        self.current_learning_iteration = d["iter"]
        return d.get("infos")

    def get_inference_policy(self, device=None):
        self.eval_mode()
        if device is not None:
            self.alg.policy.to(device)
        if self.empirical_normalization:
            norm = self.obs_normalizer.to(device) if device is not None else self.obs_normalizer
            return lambda x: self.alg.policy.act_inference(norm(x))
        return self.alg.policy.act_inference

    def train_mode(self):
        self.alg.policy.train()
        if self.empirical_normalization:
            self.obs_normalizer.train(); self.critic_obs_normalizer.train()

    def eval_mode(self):
        self.alg.policy.eval()
        if self.empirical_normalization:
            self.obs_normalizer.eval(); self.critic_obs_normalizer.eval()
