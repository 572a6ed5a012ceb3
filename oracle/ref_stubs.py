"""TEST INFRASTRUCTURE ONLY -- never imported by the product package.

Makes the UNMODIFIED reference tree under /root/reference importable in the
build container, where gym / stable_baselines3 / pickle5 / matplotlib are not
installed.  Only `oracle/make_golden.py` and the CPU
tests that pin the oracle use this; nothing that runs on the GPU box does,
because /root/reference does not exist there.

What is stubbed and why (reference file:line):
  * rendezvous_env.py:1-2      imports gym, gym.spaces  -> minimal Env and Box
    (Box.contains follows gym 0.21.0: castable dtype, shape, x>=low, x<=high)
  * utils/general.py:8,12-16   imports pickle5, stable_baselines3, sb3_contrib
  * utils/general.py:163       uses the removed alias np.float
  * utils/environment_utils.py:3 imports the git-ignored other.new_env
"""
from __future__ import annotations

import importlib
import io
import os
import pickle
import sys
import types
import zipfile

import numpy as np

REFERENCE_ROOT = os.environ.get("RDV_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "rendezvous_env.py"))


class _Box:
    """gym 0.21.0 spaces.Box, reduced to what rendezvous_env.py touches."""

    def __init__(self, low, high, shape=None, dtype=np.float32):
        self.dtype = np.dtype(dtype)
        self.shape = tuple(shape)
        self.low = np.full(self.shape, low, dtype=self.dtype)
        self.high = np.full(self.shape, high, dtype=self.dtype)

    def contains(self, x):
        if not isinstance(x, np.ndarray):
            x = np.asarray(x)
        return bool(
            np.can_cast(x.dtype, self.dtype)
            and x.shape == self.shape
            and np.all(x >= self.low)
            and np.all(x <= self.high)
        )

    def sample(self):
        return np.random.uniform(self.low, self.high).astype(self.dtype)


def _module(name, **attrs):
    mod = types.ModuleType(name)
    mod.__dict__.update(attrs)
    sys.modules[name] = mod
    return mod


def install() -> None:
    """Register the stub modules and put the reference tree on sys.path."""
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    if not hasattr(np, "float"):
        np.float = float  # utils/general.py:163 annotation

    class Env:  # gym.Env
        pass

    spaces = _module("gym.spaces", Box=_Box)
    _module("gym", Env=Env, spaces=spaces)
    sys.modules.setdefault("pickle5", pickle)

    class _Passthrough:
        def __init__(self, env, *a, **k):
            self.env = env

        def __getattr__(self, item):
            return getattr(self.__dict__["env"], item)

    class _NotInstalled:
        def __init__(self, *a, **k):
            raise RuntimeError("stable_baselines3 is not installed in this image")

        @classmethod
        def load(cls, *a, **k):
            raise RuntimeError("stable_baselines3 is not installed in this image")

    sb3 = _module("stable_baselines3", PPO=_NotInstalled)
    common = _module("stable_baselines3.common")
    _module("stable_baselines3.common.monitor", Monitor=_Passthrough)
    _module("stable_baselines3.common.vec_env", DummyVecEnv=_Passthrough)
    _module("stable_baselines3.common.utils", get_schedule_fn=lambda v: (lambda _p: v))
    sb3.common = common
    _module("sb3_contrib", RecurrentPPO=_NotInstalled)
    _module("other")
    _module("other.new_env", NewEnv=object)
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)


def import_reference():
    """Return (rendezvous_env, monte_carlo, environment_utils) reference modules."""
    install()
    env_mod = importlib.import_module("rendezvous_env")
    mc_mod = importlib.import_module("monte_carlo")
    eu_mod = importlib.import_module("utils.environment_utils")
    return env_mod, mc_mod, eu_mod


class ReferencePolicy:
    """`model.predict` of models/mlp_model_best.zip without SB3.

    SB3 1.6.2 MlpPolicy with Tanh, deterministic=True: the action is the mean
    of the Gaussian, `action_net(policy_net(obs))`, clipped to the action Box
    (monte_carlo.py:128-133).  fp32 torch on one thread, like the reference.
    """

    def __init__(self, zip_path=None):
        import torch

        zip_path = zip_path or os.path.join(REFERENCE_ROOT, "models", "mlp_model_best.zip")
        with zipfile.ZipFile(zip_path) as z:
            sd = torch.load(io.BytesIO(z.read("policy.pth")), weights_only=True, map_location="cpu")
        self.sd = {k: v.clone() for k, v in sd.items()}
        self._torch = torch
        torch.set_num_threads(1)

    def predict(self, observation, state=None, episode_start=None, deterministic=True):
        torch = self._torch
        sd = self.sd
        with torch.no_grad():
            x = torch.as_tensor(np.asarray(observation, dtype=np.float32)).reshape(1, -1)
            x = torch.tanh(torch.nn.functional.linear(
                x, sd["mlp_extractor.policy_net.0.weight"], sd["mlp_extractor.policy_net.0.bias"]))
            x = torch.tanh(torch.nn.functional.linear(
                x, sd["mlp_extractor.policy_net.2.weight"], sd["mlp_extractor.policy_net.2.bias"]))
            a = torch.nn.functional.linear(x, sd["action_net.weight"], sd["action_net.bias"])
        a = np.clip(a.numpy()[0], -1.0, 1.0)
        return a, state
