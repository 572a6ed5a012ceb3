"""Environment factory with the reference's signature (/root/reference/utils/environment_utils.py:9-73)."""
from __future__ import annotations

from copy import deepcopy

import numpy as np

from .rendezvous_env import RendezvousEnv
from .vec_env import RendezvousVecEnv

RANGE_KEYS = ("rc0_range", "vc0_range", "qc0_range", "wc0_range", "qt0_range", "wt0_range")
CONFIG_KEYS = ("rc0", "vc0", "qc0", "wc0", "qt0", "wt0") + RANGE_KEYS + \
              ("koz_radius", "corridor_half_angle", "h", "dt", "t_max")


def config_to_kwargs(config=None, stochastic=True) -> dict:
    """The config-dict handling of make_env (environment_utils.py:22-59): ``stochastic=False`` zeroes the six
    ranges, a scalar ``rc0`` becomes [0, -rc0, 0] and a scalar ``wt0`` becomes [0, 0, wt0]."""
    config = {} if config is None else dict(config)
    if stochastic is False:
        for k in RANGE_KEYS:
            config[k] = 0
    kw = {k: config.get(k) for k in CONFIG_KEYS}
    rc0, wt0 = kw["rc0"], kw["wt0"]
    if rc0 is not None and not isinstance(rc0, np.ndarray):
        kw["rc0"] = np.array([0., -rc0, 0.])
    if wt0 is not None and not isinstance(wt0, np.ndarray):
        kw["wt0"] = np.array([0., 0., wt0])
    return kw


def make_env(reward_kwargs, quiet=True, config=None, stochastic=True, **extra) -> RendezvousEnv:
    """Single-env Gym object, same arguments as the reference ``make_env``."""
    if reward_kwargs is None and not quiet:
        print("Note: reward_kwargs have not been defined. Using default values.")
    kw = config_to_kwargs(config, stochastic)
    return RendezvousEnv(reward_kwargs=reward_kwargs, quiet=quiet, **kw, **extra)


def make_vec_env(num_envs, reward_kwargs=None, config=None, stochastic=True, **extra) -> RendezvousVecEnv:
    """N-env SB3 VecEnv on the GPU; replaces ``DummyVecEnv([lambda: Monitor(make_env(...))])`` (main.py:33-34)."""
    kw = config_to_kwargs(config, stochastic)
    return RendezvousVecEnv(num_envs, reward_kwargs=reward_kwargs, **kw, **extra)


def copy_env(env):
    """environment_utils.py:66-73"""
    return deepcopy(env)
