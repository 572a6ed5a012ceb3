"""PPO training loop on device tensors -- what ``main.py`` does through Stable-Baselines3
(/root/reference/main.py:39-48, :114-118: ``PPO(MlpPolicy, lr 2e-3, batch_size 128, n_epochs 40, clip_range 0.25,
activation Tanh)`` with SB3's defaults ``gamma 0.99, gae_lambda 0.95, ent_coef 0, vf_coef 0.5, max_grad_norm 0.5,
normalize_advantage``), for the case where SB3 is not installed (it is not part of this image) or its host-side
numpy rollout buffer is the bottleneck.  With SB3 present, ``make_vec_env`` + ``stable_baselines3.PPO`` works as
in the reference; this module keeps every rollout tensor on the GPU instead:

* collection, fused (default): ONE ``rdv_rollout`` launch per iteration -- the actor runs on the tensor cores inside
  the kernel, the action is drawn from the Gaussian head with Philox noise, the env steps, finished envs restart;
  observations / unclipped actions / rewards / dones come back as [n_steps, N, ...] tensors and the critic values
  and log-probabilities are one batched torch forward afterwards
* collection, stepwise (``fused=False``): obs -> actor/critic (torch) -> sample -> clip -> ``env.step`` per step
* GAE(lambda) and the clipped-surrogate update in torch autograd (the optimiser is library code, not the hot path)
* ``evaluate`` = the callback's deterministic evaluation (custom/custom_callbacks.py:427-475) as ONE fused
  policy rollout (``rollout(policy=...)``), best model kept like ``CustomCallback`` (:477-493).

The network has SB3's ``MlpPolicy`` layout and state-dict keys, so a trained model loads into
:class:`~reinforcement_learning_rendezvous_b200.policy.MlpPolicy` (and into SB3) unchanged.
"""
from __future__ import annotations

import time
from dataclasses import dataclass, field
from typing import Optional

import numpy as np
import torch
from torch import nn

from .batched_env import BatchedRendezvousEnv
from .policy import MlpPolicy


class _Extractor(nn.Module):
    def __init__(self, obs_dim, hidden):
        super().__init__()
        self.policy_net = nn.Sequential(nn.Linear(obs_dim, hidden), nn.Tanh(), nn.Linear(hidden, hidden), nn.Tanh())
        self.value_net = nn.Sequential(nn.Linear(obs_dim, hidden), nn.Tanh(), nn.Linear(hidden, hidden), nn.Tanh())


class ActorCritic(nn.Module):
    """SB3 ``ActorCriticPolicy`` with ``net_arch=[64, 64]`` for pi and vf, Tanh, orthogonal init, state-independent
    ``log_std`` -- same parameter names as the SB3 checkpoint (``mlp_extractor.policy_net.0.weight`` ...)."""

    def __init__(self, obs_dim=17, act_dim=6, hidden=64, log_std_init=0.0):
        super().__init__()
        self.mlp_extractor = _Extractor(obs_dim, hidden)
        self.action_net = nn.Linear(hidden, act_dim)
        self.value_net = nn.Linear(hidden, 1)
        self.log_std = nn.Parameter(torch.full((act_dim,), float(log_std_init)))
        for module, gain in ((self.mlp_extractor, np.sqrt(2)), (self.action_net, 0.01), (self.value_net, 1.0)):
            for m in module.modules():
                if isinstance(m, nn.Linear):
                    nn.init.orthogonal_(m.weight, gain=gain)
                    nn.init.zeros_(m.bias)

    def forward(self, obs):
        mean = self.action_net(self.mlp_extractor.policy_net(obs))
        value = self.value_net(self.mlp_extractor.value_net(obs)).squeeze(-1)
        return mean, value

    def distribution(self, mean):
        # validate_args=False: the argument checks of torch.distributions synchronise with the device on every call
        return torch.distributions.Normal(mean, self.log_std.exp().expand_as(mean), validate_args=False)

    def to_mlp_policy(self, device) -> MlpPolicy:
        sd = {k: v.detach().float().cpu().numpy() for k, v in self.state_dict().items()}
        return MlpPolicy(sd, device=device)


def gae(rewards: torch.Tensor, values: torch.Tensor, dones: torch.Tensor, last_value: torch.Tensor, gamma: float,
        gae_lambda: float):
    """GAE(lambda) of SB3's ``RolloutBuffer.compute_returns_and_advantage`` on [T, N] tensors: ``dones[t]`` is the done
    flag returned by step t (SB3's ``episode_starts[t + 1]``; the last row is its ``dones`` argument).  No time-limit
    bootstrapping, like the reference (it never sets ``TimeLimit.truncated``).  Returns (advantages, returns)."""
    T = rewards.shape[0]
    adv = torch.zeros_like(rewards)
    last = torch.zeros_like(last_value)
    for t in reversed(range(T)):
        next_value = last_value if t == T - 1 else values[t + 1]
        not_done = 1.0 - dones[t].to(rewards.dtype)
        delta = rewards[t] + gamma * next_value * not_done - values[t]
        last = delta + gamma * gae_lambda * not_done * last
        adv[t] = last
    return adv, adv + values


@dataclass
class PPOConfig:
    """Defaults = /root/reference/main.py:39-48 (lr 2e-3, batch_size 128, n_epochs 40, clip_range 0.25) + the SB3
    defaults the shipped model was trained with (gamma 0.99, gae_lambda 0.95, ent_coef 0, vf_coef 0.5,
    max_grad_norm 0.5).  ``n_steps`` is per env: the reference collects 2048 steps from ONE env per iteration; with
    thousands of envs a few steps per env give a much larger buffer, and callers normally raise ``batch_size`` with it
    (128-sample minibatches over a 262,144-sample buffer are 2,048 optimiser steps per epoch)."""
    n_steps: int = 16                 # per env and iteration (the reference: 2048 with ONE env)
    batch_size: int = 128             # main.py:43
    n_epochs: int = 40                # main.py:44
    learning_rate: float = 2e-3       # main.py:42
    clip_range: float = 0.25          # main.py:45
    gamma: float = 0.99
    gae_lambda: float = 0.95
    ent_coef: float = 0.0
    vf_coef: float = 0.5
    max_grad_norm: float = 0.5
    normalize_advantage: bool = True
    n_evals: int = 50                 # custom_callbacks.py: n_evals
    fused: bool = True                # collect with the policy-fused rollout kernel (device env only)
    cuda_graph: bool = True           # replay one captured minibatch step (forward, loss, backward, clip, Adam)
    seed: int = 0
    log: list = field(default_factory=list)


class PPO:
    """``env``: a :class:`BatchedRendezvousEnv` (device tensors end to end: fused or stepwise collection) or a
    :class:`RendezvousVecEnv` -- the SB3 drop-in of BASELINE.json configs[2]: collection then goes through
    ``VecEnv.step`` exactly like SB3's ``collect_rollouts`` (torch policy on the GPU, numpy actions in, numpy
    observations / rewards / dones out)."""

    def __init__(self, env, config: Optional[PPOConfig] = None, policy: Optional[ActorCritic] = None):
        self.venv = None
        if not isinstance(env, BatchedRendezvousEnv):          # RendezvousVecEnv (or anything wrapping one as .env)
            self.venv, env = env, env.env
        if not env.auto_reset:
            raise ValueError("training needs an auto-resetting env")
        self.env, self.cfg = env, config or PPOConfig()
        self.device = env.device
        torch.manual_seed(self.cfg.seed)
        self.policy = (policy or ActorCritic()).to(self.device)
        # capturable: the step counters live on the device, so an optimiser step can be part of a CUDA graph
        self.optimizer = torch.optim.Adam(self.policy.parameters(), lr=self.cfg.learning_rate, eps=1e-5,
                                          capturable=bool(self.cfg.cuda_graph))
        self._graph = None                # (CUDAGraph, static minibatch tensors, static loss tensors)
        n, T = env.num_envs, self.cfg.n_steps
        dev = self.device
        self.buf_obs = torch.empty((T, n, 17), dtype=torch.float32, device=dev)
        self.buf_act = torch.empty((T, n, 6), dtype=torch.float32, device=dev)
        self.buf_logp = torch.empty((T, n), dtype=torch.float32, device=dev)
        self.buf_val = torch.empty((T, n), dtype=torch.float32, device=dev)
        self.buf_rew = torch.empty((T, n), dtype=torch.float32, device=dev)
        self.buf_done = torch.empty((T, n), dtype=torch.bool, device=dev)
        self._np_obs = self.venv.reset() if self.venv is not None else None
        self.obs = env.obs.clone() if self.venv is not None else env.reset().clone()
        self.episode_start = torch.ones(n, dtype=torch.bool, device=dev)
        self.num_timesteps = 0
        self._noise_step = 0
        self.best_eval = -float("inf")
        self.best_state = None
        kw = {k: v for k, v in env.ctor_kwargs.items()}
        self.eval_env = BatchedRendezvousEnv(self.cfg.n_evals, device=dev, seed=env.seed + 12345, auto_reset=False, **kw)

    # -- rollout collection (OnPolicyAlgorithm.collect_rollouts) ------------------------------------------------
    @torch.no_grad()
    def collect(self):
        if self.venv is not None:
            return self._collect_vecenv()
        return self._collect_fused() if self.cfg.fused else self._collect_stepwise()

    def _gae(self, last_value):
        self.num_timesteps += self.cfg.n_steps * self.env.num_envs
        return gae(self.buf_rew, self.buf_val, self.buf_done, last_value, self.cfg.gamma, self.cfg.gae_lambda)

    def _collect_vecenv(self):
        """SB3 ``OnPolicyAlgorithm.collect_rollouts`` over the VecEnv drop-in: per step one policy forward on the GPU,
        the clipped numpy action into ``VecEnv.step``, numpy results back into the device rollout buffer."""
        venv, T, dev = self.venv, self.cfg.n_steps, self.device
        obs_np = self._np_obs
        for t in range(T):
            obs_t = torch.from_numpy(obs_np).to(dev, non_blocking=True)
            mean, value = self.policy(obs_t)
            dist = self.policy.distribution(mean)
            action = dist.sample()
            self.buf_obs[t], self.buf_act[t] = obs_t, action
            self.buf_logp[t], self.buf_val[t] = dist.log_prob(action).sum(-1), value
            a_np = action.clamp(-1.0, 1.0).cpu().numpy()                    # SB3 clips to the Box before step
            obs_np, rew, done, _ = venv.step_arrays(a_np)
            self.buf_rew[t] = torch.from_numpy(rew).to(dev, non_blocking=True)
            self.buf_done[t] = torch.from_numpy(done).to(dev, non_blocking=True)
        self._np_obs = obs_np
        self.obs = torch.from_numpy(obs_np).to(dev)
        self.episode_start = self.buf_done[-1]
        _, last_value = self.policy(self.obs)
        return self._gae(last_value)

    def _collect_fused(self):
        env, T, n = self.env, self.cfg.n_steps, self.env.num_envs
        actor = self.policy.to_mlp_policy(self.device)
        out = env.rollout(T, policy=actor, stochastic=True, action_seed=self.cfg.seed + 977,
                          step_base=self._noise_step, record_obs=True, record_actions=True, record_rewards=True,
                          record_dones=True)
        self._noise_step += T
        self.buf_obs[0] = self.obs
        self.buf_obs[1:] = out["obs_steps"][:-1]
        self.buf_act.copy_(out["actions"])
        self.buf_rew.copy_(out["rewards"])
        self.buf_done.copy_(out["dones"].bool())
        mean, value = self.policy(self.buf_obs.reshape(T * n, 17))
        self.buf_val.copy_(value.reshape(T, n))
        self.buf_logp.copy_(self.policy.distribution(mean).log_prob(self.buf_act.reshape(T * n, 6)).sum(-1).reshape(T, n))
        self.obs = out["obs_steps"][-1].clone()
        self.episode_start = self.buf_done[-1]
        _, last_value = self.policy(self.obs)
        return self._gae(last_value)

    def _collect_stepwise(self):
        env, T = self.env, self.cfg.n_steps
        starts = torch.empty((T, env.num_envs), dtype=torch.bool, device=self.device)
        for t in range(T):
            mean, value = self.policy(self.obs)
            dist = self.policy.distribution(mean)
            action = dist.sample()
            self.buf_obs[t], self.buf_act[t] = self.obs, action
            self.buf_logp[t], self.buf_val[t] = dist.log_prob(action).sum(-1), value
            starts[t] = self.episode_start
            obs, rew, done = env.step(action.clamp(-1.0, 1.0).contiguous())       # SB3 clips to the Box before step
            self.buf_rew[t], self.buf_done[t] = rew.float(), done.bool()
            self.obs = obs.clone()
            self.episode_start = done.bool()
        _, last_value = self.policy(self.obs)
        # GAE(lambda) (RolloutBuffer.compute_returns_and_advantage); no time-limit bootstrapping, like the reference
        return self._gae(last_value)

    # -- clipped-surrogate update (PPO.train) --------------------------------------------------------------------
    def _minibatch_step(self, obs, act, old_logp, adv, ret):
        """One optimiser step of SB3's ``PPO.train`` on one minibatch; returns (pg_loss, value_loss, entropy)."""
        cfg = self.cfg
        mean, value = self.policy(obs)
        dist = self.policy.distribution(mean)
        logp = dist.log_prob(act).sum(-1)
        a = adv
        if cfg.normalize_advantage and a.numel() > 1:
            a = (a - a.mean()) / (a.std() + 1e-8)
        ratio = (logp - old_logp).exp()
        pg_loss = -torch.min(a * ratio, a * ratio.clamp(1 - cfg.clip_range, 1 + cfg.clip_range)).mean()
        v_loss = nn.functional.mse_loss(ret, value)
        ent_loss = -dist.entropy().sum(-1).mean()
        loss = pg_loss + cfg.ent_coef * ent_loss + cfg.vf_coef * v_loss
        self.optimizer.zero_grad(set_to_none=True)
        loss.backward()
        nn.utils.clip_grad_norm_(self.policy.parameters(), cfg.max_grad_norm)
        self.optimizer.step()
        return pg_loss.detach(), v_loss.detach(), -ent_loss.detach()

    def _captured_step(self, B):
        """The minibatch step as a CUDA graph over static [B, ...] input tensors: the ~90 small kernels of forward,
        loss, backward, gradient clipping and Adam are launch-bound at these sizes, so replaying them as one graph is
        what sets the update rate.  Captured once per batch size, after three eager steps on a side stream (PyTorch's
        whole-network capture recipe); those warm-up steps run on a copy of the parameters and optimiser state."""
        if self._graph is not None and self._graph[1][0].shape[0] == B:
            return self._graph
        dev = self.device
        static = (torch.zeros((B, 17), device=dev), torch.zeros((B, 6), device=dev), torch.zeros(B, device=dev),
                  torch.linspace(-1.0, 1.0, B, device=dev), torch.zeros(B, device=dev))   # (no draw from the RNG)
        params = list(self.policy.parameters())
        saved_p = {k: v.detach().clone() for k, v in self.policy.state_dict().items()}
        saved_o = {q: {k: v.clone() for k, v in self.optimizer.state[q].items() if torch.is_tensor(v)}
                   for q in params if q in self.optimizer.state}
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(3):
                self._minibatch_step(*static)
        torch.cuda.current_stream(dev).wait_stream(side)
        graph = torch.cuda.CUDAGraph()
        self.optimizer.zero_grad(set_to_none=True)
        with torch.cuda.graph(graph):
            losses = self._minibatch_step(*static)
        # the warm-up and the capture pass must not count as training: parameters and optimiser state go back -- IN
        # PLACE, the graph refers to these very tensors (moments and step counters that did not exist before are zeroed)
        self.policy.load_state_dict(saved_p)
        with torch.no_grad():
            for q in params:
                for k, v in self.optimizer.state[q].items():
                    if torch.is_tensor(v):
                        v.copy_(saved_o[q][k]) if q in saved_o else v.zero_()
        self._graph = (graph, static, losses)
        return self._graph

    def update(self, adv, ret):
        cfg = self.cfg
        N = adv.numel()
        obs, act = self.buf_obs.reshape(N, 17), self.buf_act.reshape(N, 6)
        old_logp, adv, ret = self.buf_logp.reshape(N), adv.reshape(N), ret.reshape(N)
        B = min(cfg.batch_size, N)
        use_graph = bool(cfg.cuda_graph) and self.device.type == "cuda"
        stats = ()
        for _ in range(cfg.n_epochs):
            perm = torch.randperm(N, device=self.device)
            for s in range(0, N, B):
                idx = perm[s:s + B]
                if use_graph and idx.numel() == B:
                    graph, static, losses = self._captured_step(B)
                    for dst, src in zip(static, (obs, act, old_logp, adv, ret)):
                        torch.index_select(src, 0, idx, out=dst)
                    graph.replay()
                    stats = losses
                else:                                       # a ragged last minibatch, or graphs switched off
                    stats = self._minibatch_step(obs[idx], act[idx], old_logp[idx], adv[idx], ret[idx])
        # the losses of the last minibatch, read back once (a read-back per minibatch would serialise host and device)
        return {k: float(v) for k, v in zip(("pg_loss", "value_loss", "entropy"), stats)} if stats else {}

    # -- deterministic evaluation (CustomCallback._on_rollout_start / evaluate_policy) --------------------------
    @torch.no_grad()
    def evaluate(self) -> dict:
        """One deterministic episode per evaluation env, scored on the device by the rollout's evaluator mode: an env
        stops at its first done, so the success / collided flags are those of the episode's end (what the reference
        callback reads), not of whatever happens afterwards."""
        from . import _native as N
        env = self.eval_env
        env.reset()
        steps = int(env.params.done_steps)
        mc = env.rollout(steps, policy=self.policy.to_mlp_policy(self.device), monte_carlo=True)["mc"]
        return {"mean_return": float(mc[:, N.MC_TOTAL_REWARD].mean()), "mean_length": float(mc[:, N.MC_EP_LEN].mean()),
                "success_rate": float((env.success > 0).float().mean()),
                "collision_rate": float((env.collided > 0).float().mean())}

    def learn(self, total_timesteps: int, eval_every: int = 1, verbose: bool = False):
        it = 0
        while self.num_timesteps < total_timesteps:
            if eval_every and it % eval_every == 0:
                ev = self.evaluate()
                if ev["mean_return"] > self.best_eval:                      # CustomCallback: keep the best model
                    self.best_eval = ev["mean_return"]
                    self.best_state = {k: v.detach().clone() for k, v in self.policy.state_dict().items()}
            else:
                ev = {}
            t0 = time.perf_counter()
            adv, ret = self.collect()
            torch.cuda.synchronize(self.device)
            t1 = time.perf_counter()
            st = self.update(adv, ret)
            torch.cuda.synchronize(self.device)
            t2 = time.perf_counter()
            row = dict(iteration=it, timesteps=self.num_timesteps, collect_s=t1 - t0, update_s=t2 - t1,
                       collect_steps_per_s=self.cfg.n_steps * self.env.num_envs / (t1 - t0),
                       mean_step_reward=float(self.buf_rew.mean()), **st, **{"eval_" + k: v for k, v in ev.items()})
            self.cfg.log.append(row)
            if verbose:
                print(row, flush=True)
            it += 1
        return self

    def save(self, path: str):
        """npz with the SB3 state-dict keys ('.' -> '__'), loadable by MlpPolicy.load."""
        sd = self.best_state or self.policy.state_dict()
        np.savez(path, **{k.replace(".", "__"): v.detach().float().cpu().numpy() for k, v in sd.items()})
