"""RendezvousVecEnv -- Stable-Baselines3 ``VecEnv`` surface over BatchedRendezvousEnv.

Replaces ``DummyVecEnv([lambda: Monitor(RendezvousEnv(...))])`` (/root/reference/main.py:33-34,
utils/general.py:55-56) with N environments stepped by one kernel launch.  Contract kept
(SB3 1.6.2 ``DummyVecEnv.step_wait`` + ``Monitor.step``):

* ``step_wait() -> (obs f32[N,17], rewards f32[N], dones bool[N], infos list[dict])``
* a finished env is reset in the same call; the returned obs is the post-reset observation and
  ``infos[i]["terminal_observation"]`` holds the episode's last observation
* ``infos[i]["episode"] = {"r": return, "l": length, "t": wall seconds}`` for finished envs (Monitor)
* no ``TimeLimit.truncated`` key is ever set (the reference treats time-outs as terminations)

Host side per step: one H2D copy of the actions from pinned memory, one kernel launch (step + fused
auto-reset), D2H copies of obs/reward/done/terminal_obs/episode_record into pinned buffers, one stream
synchronise, then one small dict per FINISHED env (all other entries share one empty dict).  The returned arrays are views of those pinned buffers and are
overwritten by the next ``step_wait``; SB3 copies them into its rollout buffer immediately.
"""
from __future__ import annotations

import gc
import time
from typing import Any, List, Optional, Sequence

import numpy as np
import torch

from . import _native as N
from .batched_env import BatchedRendezvousEnv
from .spaces import action_space, observation_space

try:                                                   # pragma: no cover - SB3 is not part of this image
    from stable_baselines3.common.vec_env import VecEnv as _SB3VecEnv
except Exception:                                      # noqa: BLE001
    _SB3VecEnv = None


class _VecEnvBase:
    """The subset of SB3's VecEnv base class that callers rely on."""

    def __init__(self, num_envs, observation_space, action_space):
        self.num_envs = num_envs
        self.observation_space = observation_space
        self.action_space = action_space
        self.render_mode = None

    def step(self, actions):
        self.step_async(actions)
        return self.step_wait()

    def _get_indices(self, indices):
        if indices is None:
            return range(self.num_envs)
        if isinstance(indices, int):
            return [indices]
        return indices

    @property
    def unwrapped(self):
        return self


_Base = _SB3VecEnv if _SB3VecEnv is not None else _VecEnvBase
_EMPTY_INFO: dict = {}


class RendezvousVecEnv(_Base):
    def __init__(self, num_envs: int, device="cuda", seed: int = 0, env_offset: int = 0, copy_outputs: bool = False,
                 rich_infos: bool = False, **ctor_kwargs):
        self.env = BatchedRendezvousEnv(num_envs, device=device, seed=seed, env_offset=env_offset, auto_reset=True,
                                        **ctor_kwargs)
        _Base.__init__(self, num_envs, observation_space(), action_space())
        n = num_envs
        self.copy_outputs = copy_outputs
        # False: infos carry only the keys SB3 reads (terminal_observation, episode); True adds is_success,
        # collided, total_delta_v, total_delta_w, end_reason (about 1 us of host time more per finished env)
        self.rich_infos = rich_infos
        self._h_act32 = torch.zeros((n, N.ACT_DIM), dtype=torch.float32).pin_memory()
        self._h_act64 = torch.zeros((n, N.ACT_DIM), dtype=torch.float64).pin_memory()
        self._d_act32 = torch.zeros((n, N.ACT_DIM), dtype=torch.float32, device=self.env.device)
        self._d_act64 = torch.zeros((n, N.ACT_DIM), dtype=torch.float64, device=self.env.device)
        self._d_rew32 = torch.zeros(n, dtype=torch.float32, device=self.env.device)
        self._h_obs = torch.zeros((n, N.OBS_DIM), dtype=torch.float32).pin_memory()
        self._h_rew = torch.zeros(n, dtype=torch.float32).pin_memory()
        self._h_done = torch.zeros(n, dtype=torch.uint8).pin_memory()
        self._h_term = torch.zeros((n, N.OBS_DIM), dtype=torch.float32).pin_memory()
        self._h_rec = torch.zeros((n, N.EP_NCOL), dtype=torch.float64).pin_memory()
        self._h_reason = torch.zeros(n, dtype=torch.int8).pin_memory()
        self._h_idx = torch.zeros(n, dtype=torch.int64).pin_memory()
        self._pending = None
        self._t_start = time.time()
        self.h2d_bytes_per_step = n * N.ACT_DIM * 4
        # device -> host per step: obs, reward, done for every env, plus the compacted rows of the finished ones
        self._d2h_fixed = n * (N.OBS_DIM * 4 + 4 + 1)
        self._d2h_row = N.OBS_DIM * 4 + N.EP_NCOL * 8 + 1 + 8
        self.d2h_bytes_per_step = self._d2h_fixed          # + _d2h_row per finished env; see d2h_bytes_last_step
        self.d2h_bytes_last_step = self._d2h_fixed
        self.d2h_bytes_total = 0

    # ------------------------------------------------------------------ VecEnv API
    def reset(self):
        self.env.reset()
        self._h_obs.copy_(self.env.obs, non_blocking=True)
        torch.cuda.current_stream(self.env.device).synchronize()
        obs = self._h_obs.numpy()
        return obs.copy() if self.copy_outputs else obs

    def step_async(self, actions):
        a = np.asarray(actions)
        if a.shape != (self.num_envs, N.ACT_DIM):
            raise ValueError(f"actions must have shape ({self.num_envs}, {N.ACT_DIM})")
        if a.dtype == np.float64:
            self._h_act64.numpy()[...] = a
            self._d_act64.copy_(self._h_act64, non_blocking=True)
            self._pending = self._d_act64
        else:
            self._h_act32.numpy()[...] = a          # float32 (what SB3 policies emit); other dtypes are cast
            self._d_act32.copy_(self._h_act32, non_blocking=True)
            self._pending = self._d_act32

    def _launch_and_fetch(self):
        if self._pending is None:
            raise RuntimeError("step_wait() called without step_async()")
        env = self.env
        env.step(self._pending)
        self._pending = None
        self._d_rew32.copy_(env.reward)
        self._h_obs.copy_(env.obs, non_blocking=True)
        self._h_rew.copy_(self._d_rew32, non_blocking=True)
        self._h_done.copy_(env.done, non_blocking=True)
        # finished envs: compact their rows on the device (ascending env index) and copy only those
        idx_d = torch.nonzero(env.done).squeeze(1)
        m = int(idx_d.numel())
        if m:
            self._h_idx[:m].copy_(idx_d, non_blocking=True)
            self._h_term[:m].copy_(env.terminal_obs.index_select(0, idx_d), non_blocking=True)
            self._h_rec[:m].copy_(env.episode_record.index_select(0, idx_d), non_blocking=True)
            self._h_reason[:m].copy_(env.end_reason.index_select(0, idx_d), non_blocking=True)
        torch.cuda.current_stream(env.device).synchronize()
        self.d2h_bytes_last_step = self._d2h_fixed + m * self._d2h_row
        self.d2h_bytes_total += self.d2h_bytes_last_step
        obs, rew = self._h_obs.numpy(), self._h_rew.numpy()
        done = self._h_done.numpy().view(np.bool_)
        return obs, rew, done, self._h_idx.numpy()[:m].copy()

    def step_arrays(self, actions):
        """``step`` without per-env Python objects: returns ``(obs, rewards, dones, finished)`` where ``finished`` is
        a dict of arrays over the envs whose episode just ended -- ``index``, ``terminal_observation`` [m,17],
        ``episode_return``, ``episode_length``, ``is_success``, ``collided``, ``total_delta_v``, ``total_delta_w``,
        ``end_reason`` (0 obs, 1 time, 2 bubble, 3 attitude).  Same data as the ``infos`` of ``step``."""
        self.step_async(actions)
        obs, rew, done, idx = self._launch_and_fetch()
        m = idx.size
        rec = self._h_rec.numpy()[:m].copy()
        finished = {
            "index": idx, "terminal_observation": self._h_term.numpy()[:m].copy(),
            "episode_return": rec[:, N.EP_RETURN], "episode_length": rec[:, N.EP_LENGTH].astype(np.int64),
            "is_success": rec[:, N.EP_SUCCESS] > 0, "collided": rec[:, N.EP_COLLIDED] > 0,
            "total_delta_v": rec[:, N.EP_DELTA_V], "total_delta_w": rec[:, N.EP_DELTA_W],
            "end_reason": self._h_reason.numpy()[:m].copy(),
        }
        if self.copy_outputs:
            return obs.copy(), rew.copy(), done.copy(), finished
        return obs, rew, done, finished

    def step_wait(self):
        obs, rew, done, idx = self._launch_and_fetch()
        # The cyclic garbage collector is paused while the per-env objects are created: a step at 65,536 envs
        # allocates a 65,536-slot list and ~6,500 small dicts, none of them cyclic, and the generation-0 passes
        # they trigger (each one walking the new list) cost more than building them (3.9 -> 1.6 ms per step).
        paused = gc.isenabled()
        if paused:
            gc.disable()
        try:
            infos = self._build_infos(idx)
        finally:
            if paused:
                gc.enable()
        if self.copy_outputs:
            return obs.copy(), rew.copy(), done.copy(), infos
        return obs, rew, done, infos

    def _build_infos(self, idx) -> List[dict]:
        infos: List[dict] = [_EMPTY_INFO] * self.num_envs
        if idx.size:
            # bulk numpy work first, then one small dict per finished env, built by comprehensions over zipped
            # columns (row views of ONE gathered array) -- about half the cost of an indexed Python loop
            m = idx.size
            rec = self._h_rec.numpy()[:m]                          # rows of the finished envs, compacted on the device
            term = list(self._h_term.numpy()[:m].copy())           # one copy; a list of row views of it
            elapsed = round(time.time() - self._t_start, 6)
            rets = np.round(rec[:, N.EP_RETURN], 6).tolist()
            lens = rec[:, N.EP_LENGTH].astype(np.int64).tolist()
            episodes = [{"r": r, "l": l, "t": elapsed} for r, l in zip(rets, lens)]
            if self.rich_infos:
                reason = [N.END_REASONS[j] for j in self._h_reason.numpy()[:m].tolist()]
                succ = (rec[:, N.EP_SUCCESS] > 0).tolist()
                coll = (rec[:, N.EP_COLLIDED] > 0).tolist()
                dv, dw = rec[:, N.EP_DELTA_V].tolist(), rec[:, N.EP_DELTA_W].tolist()
                new = [{"terminal_observation": t, "episode": e, "is_success": s, "collided": c, "total_delta_v": v,
                        "total_delta_w": w, "end_reason": r}
                       for t, e, s, c, v, w, r in zip(term, episodes, succ, coll, dv, dw, reason)]
            else:
                new = [{"terminal_observation": t, "episode": e} for t, e in zip(term, episodes)]
            for i, d in zip(idx.tolist(), new):
                infos[i] = d
        return infos

    def close(self):
        return None

    def seed(self, seed: Optional[int] = None):
        if seed is not None:
            self.env.seed = int(seed)
        return [self.env.seed + i for i in range(self.num_envs)]

    def get_attr(self, attr_name: str, indices=None) -> List[Any]:
        idx = list(self._get_indices(indices))
        env = self.env
        if attr_name in ("rc", "vc", "qc", "wc", "qt", "wt"):
            v = env.state_view(attr_name).cpu().numpy()
            return [v[i].copy() for i in idx]
        per_env = {"collided": env.collided, "success": env.success, "total_delta_v": env.total_delta_v,
                   "total_delta_w": env.total_delta_w}
        if attr_name in per_env:
            v = per_env[attr_name].cpu().numpy()
            return [v[i].item() for i in idx]
        if attr_name == "t":
            steps = env.step_count.cpu().numpy()
            return [round(float(steps[i]) * self._group_of(i).params.dt, 3) for i in idx]
        if attr_name in ("observation_space", "action_space", "render_mode"):
            return [getattr(self, attr_name)] * len(idx)
        out = []
        for i in idx:
            p = self._group_of(i).params
            if not hasattr(p, attr_name):
                raise AttributeError(attr_name)
            v = getattr(p, attr_name)
            out.append(np.array(v[:]) if hasattr(v, "__len__") else v)
        return out

    def set_attr(self, attr_name: str, value: Any, indices=None) -> None:
        idx = list(self._get_indices(indices))
        if attr_name in ("rc", "vc", "qc", "wc", "qt", "wt"):
            view = self.env.state_view(attr_name)
            view[idx] = torch.as_tensor(np.asarray(value, dtype=np.float64), device=self.env.device)
            return
        raise AttributeError(f"cannot set {attr_name!r} on RendezvousVecEnv (environment constants are fixed at "
                             "construction; build a new env or use param_batches)")

    def env_method(self, method_name: str, *method_args, indices=None, **method_kwargs) -> List[Any]:
        idx = list(self._get_indices(indices))
        if method_name in ("get_errors", "check_collision", "check_success", "dist_from_koz", "get_attitude_error"):
            err, col, suc, koz = (t.cpu().numpy() for t in self.env.errors())
            table = {"get_errors": lambda i: err[i].copy(), "check_collision": lambda i: bool(col[i]),
                     "check_success": lambda i: int(suc[i]), "dist_from_koz": lambda i: float(koz[i]),
                     "get_attitude_error": lambda i: float(err[i, 2])}
            return [table[method_name](i) for i in idx]
        if method_name == "get_observation":
            obs = self.env.observe().cpu().numpy()
            return [obs[i].copy() for i in idx]
        raise AttributeError(f"env_method {method_name!r} is not available on the batched environment")

    def env_is_wrapped(self, wrapper_class, indices=None) -> List[bool]:
        return [False] * len(list(self._get_indices(indices)))

    def get_images(self) -> Sequence[np.ndarray]:
        return []

    def render(self, mode: str = "human"):
        return None

    def _group_of(self, i: int):
        for g in self.env.groups:
            if g.lo <= i < g.hi:
                return g
        raise IndexError(i)

    # rollout statistics gathered on the device (optionally all-reduced over ranks, see distributed.py)
    def read_stats(self, reset: bool = False) -> dict:
        return self.env.read_stats(reset=reset)
