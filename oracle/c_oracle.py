"""TEST INFRASTRUCTURE ONLY -- ctypes front-end of the plain-C oracle
(oracle/rdv_oracle.c -> oracle/_build/librdv_oracle.so).

Used by tests/ (as the checker the CUDA kernels are compared with), by
__graft_entry__.smoke() and by bench.py's cpu_baseline leg.  Never imported by
the product package.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import threading

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "_build", "librdv_oracle.so")


class OrcParams(C.Structure):
    _fields_ = [
        ("rc0", C.c_double * 3), ("vc0", C.c_double * 3), ("qc0", C.c_double * 4),
        ("wc0", C.c_double * 3), ("qt0", C.c_double * 4), ("wt0", C.c_double * 3),
        ("rc0_range", C.c_double), ("vc0_range", C.c_double), ("qc0_range", C.c_double),
        ("wc0_range", C.c_double), ("qt0_range", C.c_double), ("wt0_range", C.c_double),
        ("koz_radius", C.c_double), ("corridor_half_angle", C.c_double), ("h", C.c_double),
        ("dt", C.c_double), ("t_max", C.c_double),
        ("collision_coef", C.c_double), ("bonus_coef", C.c_double), ("fuel_coef", C.c_double),
        ("att_coef", C.c_double),
        ("inertia_c", C.c_double * 9), ("inv_inertia_c", C.c_double * 9),
        ("inertia_t", C.c_double * 9), ("inv_inertia_t", C.c_double * 9),
        ("torque_c", C.c_double * 3),
        ("max_delta_v", C.c_double), ("max_delta_w", C.c_double), ("max_axial_distance", C.c_double),
        ("max_axial_speed", C.c_double), ("max_wc", C.c_double),
        ("max_attitude_error", C.c_double), ("max_rd_error", C.c_double), ("max_vd_error", C.c_double),
        ("max_qd_error", C.c_double), ("max_wd_error", C.c_double),
        ("rd", C.c_double * 3), ("capture_axis", C.c_double * 3), ("corridor_axis", C.c_double * 3),
        ("bubble0", C.c_double), ("bubble_rate", C.c_double), ("bubble_min", C.c_double), ("n", C.c_double),
        ("dt_is_integer", C.c_int),
    ]


_lib = None


def build(force=False):
    if force or not os.path.exists(LIB_PATH) or \
            os.path.getmtime(LIB_PATH) < os.path.getmtime(os.path.join(HERE, "rdv_oracle.c")):
        subprocess.run(["make", "-C", HERE, "-B" if force else "-s"], check=True,
                       stdout=subprocess.DEVNULL)
    return LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(LIB_PATH)
        assert _lib.orc_sizeof_params() == C.sizeof(OrcParams), "OrcParams layout mismatch"
        _lib.orc_philox_uniforms.argtypes = [C.c_uint64, C.c_int64, C.c_int32, C.c_void_p]
        _lib.orc_philox_raw.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32]
        _lib.orc_philox_uniforms_batch.argtypes = [C.c_uint64, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]
        _lib.orc_philox_actions.argtypes = [C.c_uint64, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p]
    return _lib


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def make_params(reward_kwargs=None, inertia=None, inertia_target=None, chaser_torque=None, **cfg):
    """Constructor kwargs of the reference env (None = default) -> OrcParams."""
    L = lib()
    p = OrcParams()
    L.orc_params_default(C.byref(p))
    for k, v in cfg.items():
        if v is None:
            continue
        if k in ("rc0", "vc0", "qc0", "wc0", "qt0", "wt0"):
            arr = np.asarray(v, dtype=float)
            getattr(p, k)[:] = arr.tolist()
        elif k in ("quiet",):
            continue
        else:
            setattr(p, k, float(v))
    for k, v in (reward_kwargs or {}).items():
        setattr(p, k, float(v))
    if inertia is not None:
        m = np.asarray(inertia, dtype=float)
        p.inertia_c[:] = m.ravel().tolist()
        p.inv_inertia_c[:] = np.linalg.inv(m).ravel().tolist()
    if inertia_target is not None:
        m = np.asarray(inertia_target, dtype=float)
        p.inertia_t[:] = m.ravel().tolist()
        p.inv_inertia_t[:] = np.linalg.inv(m).ravel().tolist()
    if chaser_torque is not None:
        p.torque_c[:] = np.asarray(chaser_torque, dtype=float).tolist()
    L.orc_params_derive(C.byref(p))
    return p


class COracleBatch:
    """n independent environments stepped by the C oracle (AoS numpy buffers)."""

    def __init__(self, params: OrcParams, n: int):
        self.p = params
        self.n = int(n)
        self.state = np.zeros((n, 20))
        self.aux = np.zeros((n, 4))            # total_delta_v, total_delta_w, t, bubble
        self.flags = np.zeros((n, 2), dtype=np.int32)   # collided, success
        self.aux[:, 3] = params.bubble0
        self.obs = np.zeros((n, 17), dtype=np.float32)
        self.rew = np.zeros(n)
        self.done = np.zeros(n, dtype=np.uint8)
        self.reason = np.zeros(n, dtype=np.int8)
        self.rk = np.zeros((n, 3), dtype=np.int32)

    def set_state(self, state20, recompute_flags=False):
        """Inject states the way monte_carlo.evaluate does (after a reset(); the
        sticky flags are NOT recomputed unless asked)."""
        self.state[:] = np.asarray(state20, dtype=float).reshape(self.n, 20)
        self.aux[:, 0:3] = 0
        self.aux[:, 3] = self.p.bubble0
        self.flags[:] = 0
        if recompute_flags:
            e, c, s, k = self.errors()
            self.flags[:, 0] = c
            self.flags[:, 1] = s

    def reset_from_uniforms(self, u, mask=None):
        u = np.ascontiguousarray(u, dtype=float).reshape(self.n, 24)
        m = None if mask is None else np.ascontiguousarray(mask, dtype=np.uint8)
        lib().orc_reset(C.byref(self.p), C.c_int64(self.n), _ptr(u), _ptr(m), _ptr(self.state),
                        _ptr(self.aux), _ptr(self.flags), _ptr(self.obs))
        return self.obs

    def _step_range(self, lo, hi, actions, f32):
        asz = 4 if f32 else 8
        L = lib()
        L.orc_step(C.byref(self.p), C.c_int64(hi - lo),
                   C.c_void_p(self.state.ctypes.data + lo * 160), C.c_void_p(self.aux.ctypes.data + lo * 32),
                   C.c_void_p(self.flags.ctypes.data + lo * 8), C.c_void_p(actions.ctypes.data + lo * 6 * asz),
                   C.c_int(int(f32)), C.c_void_p(self.obs.ctypes.data + lo * 68),
                   C.c_void_p(self.rew.ctypes.data + lo * 8), C.c_void_p(self.done.ctypes.data + lo),
                   C.c_void_p(self.reason.ctypes.data + lo), C.c_void_p(self.rk.ctypes.data + lo * 12),
                   C.c_int(0))

    def step(self, actions, threads=1):
        actions = np.ascontiguousarray(actions)
        if actions.dtype == np.float32:
            f32 = True
        else:
            actions = actions.astype(np.float64, copy=False)
            f32 = False
        assert actions.shape == (self.n, 6)
        if threads <= 1 or self.n < 2 * threads:
            self._step_range(0, self.n, actions, f32)
        else:
            bounds = np.linspace(0, self.n, threads + 1).astype(int)
            ts = [threading.Thread(target=self._step_range, args=(int(bounds[i]), int(bounds[i + 1]), actions, f32))
                  for i in range(threads)]
            for t in ts:
                t.start()
            for t in ts:
                t.join()
        return self.obs, self.rew, self.done

    def observe(self):
        out = np.zeros((self.n, 17), dtype=np.float32)
        lib().orc_observe(C.byref(self.p), C.c_int64(self.n), _ptr(self.state), _ptr(out))
        return out

    def errors(self):
        e = np.zeros((self.n, 4))
        c = np.zeros(self.n, dtype=np.uint8)
        s = np.zeros(self.n, dtype=np.uint8)
        k = np.zeros(self.n)
        lib().orc_errors(C.byref(self.p), C.c_int64(self.n), _ptr(self.state), _ptr(self.flags), _ptr(e),
                         _ptr(c), _ptr(s), _ptr(k))
        return e, c, s, k


def philox_uniforms(seed, env_ids, episode_idx):
    """The 24 reset() draws of the device's Philox stream for (seed; env id, episode index)."""
    env_ids = np.ascontiguousarray(np.asarray(env_ids, dtype=np.int64).ravel())
    episode_idx = np.ascontiguousarray(np.broadcast_to(np.asarray(episode_idx, dtype=np.int32), env_ids.shape))
    out = np.zeros((env_ids.size, 24))
    lib().orc_philox_uniforms_batch(C.c_uint64(int(seed)), C.c_int64(env_ids.size), _ptr(env_ids), _ptr(episode_idx),
                                    _ptr(out))
    return out


def philox_actions(seed, env_ids, step_index):
    """The U(-1,1) fp64 actions the fused rollout draws for (action seed; env id, step index)."""
    env_ids = np.ascontiguousarray(np.asarray(env_ids, dtype=np.int64).ravel())
    out = np.zeros((env_ids.size, 6))
    lib().orc_philox_actions(C.c_uint64(int(seed)), C.c_int64(env_ids.size), _ptr(env_ids), C.c_int64(int(step_index)),
                             _ptr(out))
    return out


def philox_raw(ctr, k0, k1):
    c = np.array(ctr, dtype=np.uint32)
    lib().orc_philox_raw(_ptr(c), C.c_uint32(k0), C.c_uint32(k1))
    return c
