"""TEST / BENCH INFRASTRUCTURE ONLY -- the CPU baseline harness.

A stand-in for Stable-Baselines3's ``SubprocVecEnv`` (SB3 is not part of this image) with the identical
protocol: one environment per worker process, a ``multiprocessing.Pipe`` per worker, commands
``step`` / ``reset`` / ``close``, auto-reset and ``terminal_observation`` inside the worker.  The
environment is the numpy restatement of the reference (oracle/rdv_oracle.py): the reference's numpy
3-/4-vector math plus the restated adaptive RK45 (``integrator="restated"``, bit-identical to
``scipy.integrate.solve_ivp`` but without SciPy's object overhead).  Measured in the build container on one
core: the unmodified reference env 195-253 steps/s, this restatement 247 steps/s, the restatement calling
SciPy itself (``integrator="scipy"``) 122 steps/s -- so the default never under-states the reference.  The
reference tree itself cannot travel to the GPU box; the restatement is bit-exact against it
(tests/test_oracle_golden.py).

Only bench.py (``cpu_baseline`` leg and ``--impl reference``) and tests import this module.
"""
from __future__ import annotations

import multiprocessing as mp
import os
import time

import numpy as np


def _worker(remote, parent_remote, seed, env_kwargs):
    parent_remote.close()
    from oracle.rdv_oracle import OracleEnv
    env = OracleEnv(integrator=env_kwargs.pop("integrator", "restated"), rng=np.random.RandomState(seed), **env_kwargs)
    try:
        while True:
            cmd, data = remote.recv()
            if cmd == "step":
                obs, rew, done, info = env.step(data)
                info = {}
                if done:
                    info["terminal_observation"] = obs
                    obs = env.reset()
                remote.send((obs, rew, done, info))
            elif cmd == "reset":
                remote.send(env.reset())
            elif cmd == "close":
                remote.close()
                break
            else:
                raise NotImplementedError(cmd)
    except (EOFError, KeyboardInterrupt):
        pass


class OracleSubprocVecEnv:
    def __init__(self, n_procs=None, seed=0, start_method="spawn", **env_kwargs):
        self.num_envs = n_procs = int(n_procs or os.cpu_count() or 1)
        ctx = mp.get_context(start_method)
        self.remotes, self.work_remotes = zip(*[ctx.Pipe() for _ in range(n_procs)])
        self.processes = []
        # one BLAS/OpenMP thread per worker (the spawned interpreters inherit the environment at start-up)
        saved = {k: os.environ.get(k) for k in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS")}
        for k in saved:
            os.environ[k] = "1"
        for i, (work, remote) in enumerate(zip(self.work_remotes, self.remotes)):
            proc = ctx.Process(target=_worker, args=(work, remote, seed + i, dict(env_kwargs)), daemon=True)
            proc.start()
            self.processes.append(proc)
            work.close()
        for k, v in saved.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v

    def reset(self):
        for r in self.remotes:
            r.send(("reset", None))
        return np.stack([r.recv() for r in self.remotes])

    def step(self, actions):
        for r, a in zip(self.remotes, actions):
            r.send(("step", a))
        results = [r.recv() for r in self.remotes]
        obs, rews, dones, infos = zip(*results)
        return np.stack(obs), np.array(rews), np.array(dones), list(infos)

    def close(self):
        for r in self.remotes:
            try:
                r.send(("close", None))
            except (BrokenPipeError, OSError):
                pass
        for p in self.processes:
            p.join(timeout=10)


def time_subproc_baseline(steps, warmup=1, n_procs=None, seed=0, max_seconds=None, min_seconds=None, **env_kwargs):
    """Random fp64 actions through the SubprocVecEnv protocol.  Returns dict(value=env-steps/s, cores, steps,
    seconds).  ``max_seconds`` bounds the timed part (the loop stops early and reports the steps it did);
    ``min_seconds`` is a floor: the loop keeps stepping past ``steps`` until that much time has been sampled."""
    venv = OracleSubprocVecEnv(n_procs=n_procs, seed=seed, **env_kwargs)
    try:
        rng = np.random.default_rng(seed)
        venv.reset()
        for _ in range(warmup):
            venv.step(rng.uniform(-1, 1, (venv.num_envs, 6)))
        done_steps = 0
        t0 = time.perf_counter()
        while True:
            venv.step(rng.uniform(-1, 1, (venv.num_envs, 6)))
            done_steps += 1
            el = time.perf_counter() - t0
            if max_seconds is not None and el > max_seconds:
                break
            if done_steps >= steps and (min_seconds is None or el >= min_seconds):
                break
        dt = time.perf_counter() - t0
    finally:
        venv.close()
    return dict(value=done_steps * venv.num_envs / dt, cores=venv.num_envs, steps=done_steps, seconds=dt)


def time_single_process(seconds=3.0, seed=0, **env_kwargs):
    """One env stepped in this process (what the reference's ``DummyVecEnv`` of main.py:33-34 does): random fp64
    actions, reset on done, no IPC.  Returns dict(value=env-steps/s, steps, seconds)."""
    from oracle.rdv_oracle import OracleEnv
    env = OracleEnv(integrator=env_kwargs.pop("integrator", "restated"), rng=np.random.RandomState(seed), **env_kwargs)
    rng = np.random.default_rng(seed)
    env.reset()
    steps = 0
    t0 = time.perf_counter()
    while time.perf_counter() - t0 < seconds:
        _, _, done, _ = env.step(rng.uniform(-1, 1, 6))
        if done:
            env.reset()
        steps += 1
    dt = time.perf_counter() - t0
    return dict(value=steps / dt, steps=steps, seconds=dt)


def time_c_port(n_envs=4096, steps=20, threads=None, seed=0):
    """The plain-C oracle (oracle/rdv_oracle.c) on all host threads -- what a compiled CPU port achieves."""
    from oracle import c_oracle as CO
    threads = int(threads or os.cpu_count() or 1)
    b = CO.COracleBatch(CO.make_params(), n_envs)
    rng = np.random.default_rng(seed)
    b.reset_from_uniforms(rng.random((n_envs, 24)))
    b.step(rng.uniform(-1, 1, (n_envs, 6)), threads=threads)
    t0 = time.perf_counter()
    for _ in range(steps):
        a = rng.uniform(-1, 1, (n_envs, 6))
        _, _, done = b.step(a, threads=threads)
        if done.any():
            b.reset_from_uniforms(rng.random((n_envs, 24)), mask=done)
    dt = time.perf_counter() - t0
    return dict(value=n_envs * steps / dt, cores=threads, steps=steps, seconds=dt, envs=n_envs)
