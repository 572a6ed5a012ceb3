"""RendezvousVecEnv -- Stable-Baselines3 ``VecEnv`` surface over BatchedRendezvousEnv.

Replaces ``DummyVecEnv([lambda: Monitor(RendezvousEnv(...))])`` (/root/reference/main.py:33-34,
utils/general.py:55-56) with N environments stepped by one kernel launch.  Contract kept
(SB3 1.6.2 ``DummyVecEnv.step_wait`` + ``Monitor.step``):

* ``step_wait() -> (obs f32[N,17], rewards f32[N], dones bool[N], infos list[dict])``
* a finished env is reset in the same call; the returned obs is the post-reset observation and
  ``infos[i]["terminal_observation"]`` holds the episode's last observation
* ``infos[i]["episode"] = {"r": return, "l": length, "t": wall seconds}`` for finished envs (Monitor)
* no ``TimeLimit.truncated`` key is ever set (the reference treats time-outs as terminations)
* the returned arrays are never overwritten by a later step (``DummyVecEnv`` hands out copies; SB3's
  ``collect_rollouts`` keeps ``_last_obs`` / ``_last_episode_starts`` across the next ``env.step``)

Host side per step: ONE host-to-device copy of the actions from pinned memory, ONE kernel launch (step +
fused auto-reset; it writes float32 rewards and appends the finished envs' rows -- index, end reason, terminal
observation, episode record -- through one atomic counter), ONE device-to-host copy of a contiguous block
``[count | rewards | dones | obs | finished rows]`` into a pinned host block, ONE stream synchronise, then one
small dict per FINISHED env.  No torch kernel runs on the step path.

Zero-copy without aliasing: the pinned host blocks come from a small pool, and a block is reused only when
no numpy view of it is alive any more (``sys.getrefcount`` of the block's base array), so the arrays a step
returns stay valid for as long as the caller keeps them -- exactly like copies -- at the cost of none.  A
caller that hoards arrays beyond the pool's capacity gets real copies from then on.
"""
from __future__ import annotations

import gc
import sys
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

# numpy mirror of RdvFinishedRow (include/rdv_b200.h), 128 bytes
_NO_TERM = np.zeros((0, N.OBS_DIM), dtype=np.float32)
FINISHED_ROW = np.dtype([("env", "<i4"), ("end_reason", "<i4"), ("terminal_obs", "<f4", (N.OBS_DIM,)), ("pad", "<f4"),
                         ("record", "<f8", (N.EP_NCOL,))])
_NO_ROWS = np.zeros(0, dtype=FINISHED_ROW)
assert FINISHED_ROW.itemsize == 128


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


class _PinnedBlockPool:
    """Pinned host blocks of ``nbytes``; ``acquire`` returns one no live numpy view refers to."""

    def __init__(self, nbytes: int, initial: int = 3, capacity: int = 8):
        self.nbytes, self.capacity = int(nbytes), int(capacity)
        self._tensors: List[torch.Tensor] = []
        self._arrays: List[np.ndarray] = []
        for _ in range(initial):
            self._grow()
        self._next = 0

    def _grow(self):
        t = torch.zeros(self.nbytes, dtype=torch.uint8).pin_memory()     # zeros: every page is touched here
        self._tensors.append(t)
        self._arrays.append(t.numpy())

    def acquire(self):
        """(tensor, base array, shared) -- ``shared`` is True when every block is still referenced by the caller
        and the pool is at capacity: the block returned then is a scratch whose views must be copied."""
        k = len(self._arrays)
        for j in range(k):
            i = (self._next + j) % k
            if sys.getrefcount(self._arrays[i]) == 2:          # the list's reference + getrefcount's argument
                self._next = (i + 1) % k
                return self._tensors[i], self._arrays[i], False
        if k < self.capacity:
            self._grow()
            self._next = 0
            return self._tensors[k], self._arrays[k], False
        if not hasattr(self, "_scratch"):
            self._scratch = torch.zeros(self.nbytes, dtype=torch.uint8).pin_memory()
        return self._scratch, self._scratch.numpy(), True


class RendezvousVecEnv(_Base):
    def __init__(self, num_envs: int, device="cuda", seed: int = 0, env_offset: int = 0, copy_outputs: bool = False,
                 rich_infos: bool = False, gc_freeze: bool = True, **ctor_kwargs):
        self.env = BatchedRendezvousEnv(num_envs, device=device, seed=seed, env_offset=env_offset, auto_reset=True,
                                        **ctor_kwargs)
        _Base.__init__(self, num_envs, observation_space(), action_space())
        n = num_envs
        # True: obs / rewards / dones are np.copy'd out of the pinned block (plain pageable arrays).  False
        # (default): views of a pinned block that is not reused while any view of it is alive -- equally safe.
        self.copy_outputs = copy_outputs
        # False: infos carry only the keys SB3 reads (terminal_observation, episode); True adds is_success,
        # collided, total_delta_v, total_delta_w, end_reason (about 1 us of host time more per finished env)
        self.rich_infos = rich_infos
        self.layout = self.env.enable_host_block()
        self._pool = _PinnedBlockPool(self.layout["total"])
        self._h_act32 = torch.zeros((n, N.ACT_DIM), dtype=torch.float32).pin_memory()
        self._h_act64 = torch.zeros((n, N.ACT_DIM), dtype=torch.float64).pin_memory()
        self._a_act32, self._a_act64 = self._h_act32.numpy(), self._h_act64.numpy()
        self._d_act32 = torch.zeros((n, N.ACT_DIM), dtype=torch.float32, device=self.env.device)
        self._d_act64 = torch.zeros((n, N.ACT_DIM), dtype=torch.float64, device=self.env.device)
        self._pending = None
        self._host = N.host()
        self._t_start = time.time()
        # one independent dict per env (DummyVecEnv semantics); finished envs get a fresh one per episode end
        self._infos: List[dict] = [{} for _ in range(n)]
        self._dirty = b""                               # packed int32: envs whose slot holds last step's episode-end dict
        self._rows_guess = max(64, n // 8)              # finished rows fetched with the fixed part of the block
        self.h2d_bytes_per_step = n * N.ACT_DIM * 4
        # device -> host per step: count, obs, reward, done of every env, plus the rows of the finished ones
        self._d2h_fixed = self.layout["rows"]
        self._d2h_row = self.layout["row_bytes"]
        self.d2h_bytes_per_step = self._d2h_fixed          # + _d2h_row per fetched row; see d2h_bytes_last_step
        self.d2h_bytes_last_step = self._d2h_fixed
        self.d2h_bytes_total = 0
        self.extra_fetches = 0                             # steps that needed a second copy for their finished rows
        if gc_freeze:
            # The long-lived objects above never need another walk by the cyclic GC.  An empty dict is not tracked by
            # the collector yet -- it would start being tracked, as a YOUNG object, when its env's first episode-end
            # dict goes in -- so each per-env dict is made to hold a container once; then all 65,536 of them are frozen
            # into the permanent generation, and the generation-1 / -2 passes of a long run (6 ms each otherwise: they
            # walk every per-env dict) have nothing left to walk.
            for d in self._infos:
                d[0] = d
                del d[0]
            gc.freeze()

    # ------------------------------------------------------------------ VecEnv API
    def _views(self, base: np.ndarray, shared: bool):
        n, L = self.num_envs, self.layout
        obs = base[L["obs"]:L["obs"] + 4 * N.OBS_DIM * n].view(np.float32).reshape(n, N.OBS_DIM)
        rew = base[L["reward"]:L["reward"] + 4 * n].view(np.float32)
        done = base[L["done"]:L["done"] + n].view(np.bool_)
        if shared or self.copy_outputs:
            return obs.copy(), rew.copy(), done.copy()
        return obs, rew, done

    def reset(self):
        env = self.env
        env.reset()
        t, base, shared = self._pool.acquire()
        lo, hi = self.layout["obs"], self.layout["rows"]
        t[lo:hi].copy_(env.host_block[lo:hi], non_blocking=True)
        torch.cuda.current_stream(env.device).synchronize()
        return self._views(base, shared)[0]

    def step_async(self, actions):
        a = np.asarray(actions)
        if a.shape != (self.num_envs, N.ACT_DIM):
            raise ValueError(f"actions must have shape ({self.num_envs}, {N.ACT_DIM})")
        if a.dtype == np.float64:
            self._a_act64[...] = a
            self._d_act64.copy_(self._h_act64, non_blocking=True)
            self._pending = self._d_act64
        else:
            self._a_act32[...] = a                  # float32 (what SB3 policies emit); other dtypes are cast
            self._d_act32.copy_(self._h_act32, non_blocking=True)
            self._pending = self._d_act32

    def _launch_and_fetch(self, renew_infos: bool = False):
        """One launch, one device-to-host copy, one synchronise.  Returns (obs, rew, done, rows) with ``rows`` the
        finished envs' records (structured view of the pinned block, in the order the kernel appended them).
        ``renew_infos``: the dicts of the envs that finished LAST step are replaced by fresh ones while the GPU works
        (host work that does not depend on this step's results, done in the shadow of the kernel and the copy)."""
        if self._pending is None:
            raise RuntimeError("step_wait() called without step_async()")
        env, L = self.env, self.layout
        env.step(self._pending)
        self._pending = None
        t, base, shared = self._pool.acquire()
        guess = self._rows_guess
        nbytes = L["rows"] + self._d2h_row * guess
        stream = torch.cuda.current_stream(env.device)
        t[:nbytes].copy_(env.host_block[:nbytes], non_blocking=True)
        if renew_infos and self._dirty:
            self._dirty = self._host.build_infos(self._infos, self._dirty, _NO_ROWS, _NO_TERM, 0.0, False, N.END_REASONS)
        stream.synchronize()
        m = int(base[0:4].view(np.int32)[0])
        fetched = guess
        if m > guess:                                   # rare: more episodes ended than the running bound allowed for
            lo, hi = nbytes, L["rows"] + self._d2h_row * m
            t[lo:hi].copy_(env.host_block[lo:hi], non_blocking=True)
            stream.synchronize()
            fetched = m
            self.extra_fetches += 1
        # running bound for the next step: 25 % + 4 sigma (binomial) above what this step saw
        self._rows_guess = min(self.num_envs, int(1.25 * m + 4.0 * (m ** 0.5)) + 32)
        self.d2h_bytes_last_step = self._d2h_fixed + fetched * self._d2h_row
        self.d2h_bytes_total += self.d2h_bytes_last_step
        obs, rew, done = self._views(base, shared)
        rows = base[L["rows"]:L["rows"] + self._d2h_row * m].view(FINISHED_ROW)
        return obs, rew, done, rows

    def step_arrays(self, actions):
        """``step`` without per-env Python objects: returns ``(obs, rewards, dones, finished)`` where ``finished`` is
        a dict of arrays over the envs whose episode just ended -- ``index``, ``terminal_observation`` [m,17],
        ``episode_return``, ``episode_length``, ``is_success``, ``collided``, ``total_delta_v``, ``total_delta_w``,
        ``end_reason`` (0 obs, 1 time, 2 bubble, 3 attitude).  Same data as the ``infos`` of ``step``."""
        self.step_async(actions)
        obs, rew, done, rows = self._launch_and_fetch()
        rows = np.take(rows, np.argsort(rows["env"]))   # ascending env index; ONE gather of whole 128-byte rows into
        rec = rows["record"]                            # own memory, the fields below are views of it (no view of the
        finished = {                                    # pinned block outlives this call)
            "index": rows["env"].astype(np.int64), "terminal_observation": rows["terminal_obs"],
            "episode_return": rec[:, N.EP_RETURN], "episode_length": rec[:, N.EP_LENGTH].astype(np.int64),
            "is_success": rec[:, N.EP_SUCCESS] > 0, "collided": rec[:, N.EP_COLLIDED] > 0,
            "total_delta_v": rec[:, N.EP_DELTA_V], "total_delta_w": rec[:, N.EP_DELTA_W],
            "end_reason": rows["end_reason"].astype(np.int8),
        }
        return obs, rew, done, finished

    def step_wait(self):
        # The cyclic garbage collector is paused while the per-env objects are created: a step at 65,536 envs
        # allocates ~6,500 small dicts, none of them cyclic, and the generation-0 passes they trigger cost more
        # than building them.
        paused = gc.isenabled()
        if paused:
            gc.disable()
        try:
            obs, rew, done, rows = self._launch_and_fetch(renew_infos=True)
            infos = self._build_infos(rows)
        finally:
            if paused:
                gc.enable()
        del rows                                        # the last view of the pinned block besides obs / rew / done
        return obs, rew, done, infos

    def _build_infos(self, rows) -> List[dict]:
        """One fresh dict per finished env into this VecEnv's list of per-env dicts (built by csrc/rdv_host.c: the
        interpreter would spend longer on these ~2 objects per episode end than the GPU on the whole step).  The
        LIST object is the same every step -- like DummyVecEnv's buf_infos -- and an env that is still running keeps
        its own (normally empty) dict; consume or copy the infos before the next step."""
        m = rows.shape[0]
        term = np.ascontiguousarray(rows["terminal_obs"]) if m else _NO_TERM      # own memory: frees the pinned block
        elapsed = round(time.time() - self._t_start, 6)
        self._dirty = self._host.build_infos(self._infos, self._dirty, rows, term, elapsed, self.rich_infos,
                                             N.END_REASONS)
        return self._infos

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
