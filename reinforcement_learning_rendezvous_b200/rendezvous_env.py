"""RendezvousEnv -- the reference's single-environment Gym surface over the CUDA kernels.

Drop-in for ``rendezvous_env.RendezvousEnv`` (/root/reference/rendezvous_env.py:10-604):
same constructor keywords, same ``reset()`` / ``step(action)`` contract, same public
attributes (``rc vc qc wc qt wt t collided success total_delta_v ...``, readable AND
writable the way monte_carlo.evaluate (monte_carlo.py:106-112) and the callbacks
(custom/custom_callbacks.py:211-267) use them) and the same helper methods.

All environment arithmetic runs on the GPU through librdv_b200.so on a one-env
batch; the host keeps a mirror of the 20 state numbers + counters, uploads it before
and downloads it after each device call so attribute writes and in-place edits behave
like they do on the reference's plain Python object.  ``copy.deepcopy(env)`` works.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _native as N
from .batched_env import BatchedRendezvousEnv, _stream_ptr
from .spaces import action_space, observation_space

_STATE = (("rc", 0, 3), ("vc", 3, 6), ("qc", 6, 10), ("wc", 10, 13), ("qt", 13, 17), ("wt", 17, 20))


class RendezvousEnv:
    metadata = {"render.modes": []}
    reward_range = (-float("inf"), float("inf"))
    spec = None

    def __init__(self, rc0=None, vc0=None, qc0=None, wc0=None, qt0=None, wt0=None,
                 rc0_range=None, vc0_range=None, qc0_range=None, wc0_range=None, qt0_range=None, wt0_range=None,
                 reward_kwargs=None, koz_radius=None, corridor_half_angle=None, h=None, dt=None, t_max=None,
                 quiet=False, *, device="cuda", seed=None, integrator="rk45", inertia=None, inertia_target=None,
                 chaser_torque=None):
        self._ctor = dict(rc0=rc0, vc0=vc0, qc0=qc0, wc0=wc0, qt0=qt0, wt0=wt0, rc0_range=rc0_range,
                          vc0_range=vc0_range, qc0_range=qc0_range, wc0_range=wc0_range, qt0_range=qt0_range,
                          wt0_range=wt0_range, reward_kwargs=reward_kwargs, koz_radius=koz_radius,
                          corridor_half_angle=corridor_half_angle, h=h, dt=dt, t_max=t_max,
                          integrator=integrator, inertia=inertia, inertia_target=inertia_target,
                          chaser_torque=chaser_torque)
        if seed is None:        # the reference draws from the global np.random stream (rendezvous_env.py:229-250)
            seed = int(np.random.randint(0, 2 ** 31 - 1))
        self._seed = int(seed)
        self._b = BatchedRendezvousEnv(1, device=device, seed=self._seed, auto_reset=False, track_stats=False, ld=1,
                                       **self._ctor)
        p = self._b.params
        # ---- attributes of the reference constructor (rendezvous_env.py:47-126) ----
        self.nominal_rc0, self.nominal_vc0 = np.array(p.rc0[:]), np.array(p.vc0[:])
        self.nominal_qc0, self.nominal_wc0 = np.array(p.qc0[:]), np.array(p.wc0[:])
        self.nominal_qt0, self.nominal_wt0 = np.array(p.qt0[:]), np.array(p.wt0[:])
        for k in ("rc0_range", "vc0_range", "qc0_range", "wc0_range", "qt0_range", "wt0_range", "koz_radius",
                  "corridor_half_angle", "h", "max_delta_v", "max_delta_w", "max_axial_distance", "max_wc",
                  "max_attitude_error", "max_rd_error", "max_vd_error", "max_qd_error", "max_wd_error", "bubble_min",
                  "n"):
            setattr(self, k, getattr(p, k))
        self.dt = 1 if dt is None else dt           # keeps the caller's int/float type, like the reference (:69)
        self.t_max = 120 if t_max is None else t_max
        self.capture_axis = np.array([0, 1, 0])
        self.corridor_axis = np.array([0, -1, 0])
        self.rd = np.array([0, -2, 0])
        self.m = 100
        self.inertia = np.array(p.inertia_c[:]).reshape(3, 3)
        self.inv_inertia = np.array(p.inv_inertia_c[:]).reshape(3, 3)
        self.inertia_target = np.array(p.inertia_t[:]).reshape(3, 3)
        self.inv_inertia_target = np.array(p.inv_inertia_t[:]).reshape(3, 3)
        self.max_axial_speed = 5
        self.max_wt = np.radians(10)
        self.bubble_radius0 = p.bubble0
        self.bubble_decrease_rate = p.bubble_rate
        self.reward_kwargs = {} if reward_kwargs is None else reward_kwargs
        self.mu, self.Re = 3.986004418e14, 6371e3
        self.ro = self.Re + self.h
        self.viewer = None
        self.quiet = quiet
        self.observation_space = observation_space()
        self.action_space = action_space()
        # ---- state (None until reset(), like the reference) ----
        self.rc = self.vc = self.qc = self.wc = self.qt = self.wt = None
        self.t = None
        self.collided = None
        self.success = None
        self.bubble_radius = None
        self.total_delta_v = None
        self.total_delta_w = None
        self._steps = 0
        self._episode_return = 0.0
        self._eval_cache = None
        self._alloc_staging()

    # ------------------------------------------------------------------ host <-> device mirror
    def _alloc_staging(self):
        dev = self._b.device
        self._h_f64 = torch.zeros(N.NF64, dtype=torch.float64).pin_memory()
        self._h_i32 = torch.zeros(N.NI32, dtype=torch.int32).pin_memory()
        self._h_act64 = torch.zeros((1, 6), dtype=torch.float64).pin_memory()
        self._h_act32 = torch.zeros((1, 6), dtype=torch.float32).pin_memory()
        self._d_act64 = torch.zeros((1, 6), dtype=torch.float64, device=dev)
        self._d_act32 = torch.zeros((1, 6), dtype=torch.float32, device=dev)
        self._h_obs = torch.zeros((1, 17), dtype=torch.float32).pin_memory()
        self._h_rew = torch.zeros(1, dtype=torch.float64).pin_memory()
        self._h_done = torch.zeros(1, dtype=torch.uint8).pin_memory()
        self._h_reason = torch.zeros(1, dtype=torch.int8).pin_memory()
        self._h_misc = torch.zeros(6, dtype=torch.float64).pin_memory()      # errors[4], koz, (unused)
        self._d_vec = torch.zeros(10, dtype=torch.float64, device=dev)        # q[4] v[3] out[3]
        self._h_vec = torch.zeros(10, dtype=torch.float64).pin_memory()

    def _push(self):
        """host mirror -> device state (picks up attribute assignment and in-place edits)."""
        if self.rc is None:
            raise RuntimeError("call reset() before using the environment")
        f = self._h_f64.numpy()
        for name, lo, hi in _STATE:
            f[lo:hi] = np.asarray(getattr(self, name), dtype=np.float64).reshape(hi - lo)
        f[N.TDV], f[N.TDW], f[N.EPRET] = float(self.total_delta_v), float(self.total_delta_w), self._episode_return
        i = self._h_i32.numpy()
        self._steps = int(round(float(self.t) / float(self.dt)))
        i[N.I_STEP], i[N.I_SUCCESS], i[N.I_COLLIDED] = self._steps, int(self.success), int(bool(self.collided))
        self._b.f64.view(-1).copy_(self._h_f64, non_blocking=True)
        self._b.i32.view(-1)[:N.I_EPISODE].copy_(self._h_i32[:N.I_EPISODE], non_blocking=True)

    def _pull(self):
        """device state -> host mirror (synchronises the stream)."""
        self._h_f64.copy_(self._b.f64.view(-1), non_blocking=True)
        self._h_i32.copy_(self._b.i32.view(-1), non_blocking=True)
        torch.cuda.current_stream(self._b.device).synchronize()
        f, i = self._h_f64.numpy(), self._h_i32.numpy()
        for name, lo, hi in _STATE:
            setattr(self, name, f[lo:hi].copy())
        self.total_delta_v, self.total_delta_w = float(f[N.TDV]), float(f[N.TDW])
        self._episode_return = float(f[N.EPRET])
        self._steps = int(i[N.I_STEP])
        self.success = int(i[N.I_SUCCESS])
        self.collided = bool(i[N.I_COLLIDED])
        self.t = self._time_of(self._steps)
        p = self._b.params
        self.bubble_radius = max(p.bubble0 - self._steps * p.bubble_rate, p.bubble_min)

    def _time_of(self, steps):
        # t = round(t + dt, 3) per step (rendezvous_env.py:193); stays an int when dt is an int
        if isinstance(self.dt, (int, np.integer)):
            return int(steps) * int(self.dt)
        return round(steps * float(self.dt), 3)

    # ------------------------------------------------------------------ Gym API
    def reset(self):
        """rendezvous_env.py:223-270.  Initial-state draws come from the device Philox stream keyed by
        (seed, episode index)."""
        self._b.reset()
        self._h_obs.copy_(self._b.obs, non_blocking=True)
        self._pull()
        return self._h_obs.numpy()[0].copy()

    def step(self, action):
        """rendezvous_env.py:160-221."""
        action_in = action
        a = np.asarray(action)
        assert a.shape == (6,)
        if a.dtype == np.float32:
            self._h_act32.numpy()[0] = a
            self._d_act32.copy_(self._h_act32, non_blocking=True)
            d_act = self._d_act32
        else:
            self._h_act64.numpy()[0] = a.astype(np.float64)
            self._d_act64.copy_(self._h_act64, non_blocking=True)
            d_act = self._d_act64
        self._push()
        b = self._b
        b.step(d_act)
        self._h_obs.copy_(b.obs, non_blocking=True)
        self._h_rew.copy_(b.reward, non_blocking=True)
        self._h_done.copy_(b.done, non_blocking=True)
        self._h_reason.copy_(b.end_reason, non_blocking=True)
        self._pull()
        obs = self._h_obs.numpy()[0].copy()
        rew = float(self._h_rew[0])
        done = bool(self._h_done[0])
        if done and not self.quiet:
            dist = float(np.linalg.norm(self.rc))
            print("Episode end" + " | r = " + str(round(dist, 2)).rjust(5) + " | t = " + str(self.t).rjust(4) +
                  " | " + N.END_REASONS[int(self._h_reason[0])].center(8) + " | " +
                  ("Collided" if self.collided else " "))
        info = {"observation": obs, "reward": rew, "done": done, "action": action_in}
        return obs, rew, done, info

    def render(self, mode="human"):
        return None

    def close(self):
        return None

    def seed(self, seed=None):
        if seed is not None:
            self._seed = self._b.seed = int(seed)
        return [self._seed]

    # ------------------------------------------------------------------ helper methods callers use
    def get_observation(self):
        """rendezvous_env.py:294-311"""
        self._push()
        self._b.observe(out=self._b.obs)
        self._h_obs.copy_(self._b.obs, non_blocking=True)
        torch.cuda.current_stream(self._b.device).synchronize()
        return self._h_obs.numpy()[0].copy()

    def _evaluate(self):
        """errors / collision / success / dist_from_koz of the CURRENT host-side state, one kernel + one read-back;
        callers like monte_carlo.evaluate ask for all four after every step, so the result is kept until the
        state (including in-place edits of the arrays) or the sticky collided flag changes."""
        key = (np.concatenate([np.asarray(getattr(self, n), dtype=np.float64).ravel() for n, _, _ in _STATE]).tobytes(),
               bool(self.collided))
        if self._eval_cache is not None and self._eval_cache[0] == key:
            return self._eval_cache[1]
        self._push()
        err, col, suc, koz = self._b.errors()
        out = torch.cat([err.view(-1), koz.view(-1), col.to(torch.float64), suc.to(torch.float64)]).cpu().numpy()
        res = (out[0:4].copy(), bool(out[5]), int(out[6]), float(out[4]))
        self._eval_cache = (key, res)
        return res

    def get_errors(self):
        """rendezvous_env.py:451-468 -> array [pos, vel, att, rot]"""
        return self._evaluate()[0].copy()

    def get_pos_error(self, goal_pos=None):
        """rendezvous_env.py:443-449"""
        if goal_pos is None:
            return float(self._evaluate()[0][0])
        return float(np.linalg.norm(np.asarray(self.rc) - np.asarray(goal_pos)))

    def get_goal_pos(self):
        """rendezvous_env.py:436-441"""
        return self.target2lvlh(self.rd)

    def get_attitude_error(self):
        """rendezvous_env.py:424-434"""
        return float(self._evaluate()[0][2])

    def check_collision(self):
        """rendezvous_env.py:388-404"""
        return self._evaluate()[1]

    def check_success(self):
        """rendezvous_env.py:406-422"""
        return self._evaluate()[2]

    def dist_from_koz(self):
        """rendezvous_env.py:510-537"""
        return self._evaluate()[3]

    def _frame(self, q, v, transpose):
        h = self._h_vec.numpy()
        h[0:4] = np.asarray(q, dtype=np.float64)
        h[4:7] = np.asarray(v, dtype=np.float64)
        self._d_vec.copy_(self._h_vec, non_blocking=True)
        d = self._d_vec
        with torch.cuda.device(self._b.device):
            N.check(self._b.lib.rdv_frame_transform(d.data_ptr(), d.data_ptr() + 32, d.data_ptr() + 56, 1,
                                                    int(transpose), _stream_ptr(self._b.device)),
                    "rdv_frame_transform")
        return d[7:10].cpu().numpy()

    def chaser2lvlh(self, vec):
        return self._frame(self.qc, vec, 0)

    def target2lvlh(self, vec):
        return self._frame(self.qt, vec, 0)

    def lvlh2chaser(self, vec):
        return self._frame(self.qc, vec, 1)

    def lvlh2target(self, vec):
        return self._frame(self.qt, vec, 1)

    # ------------------------------------------------------------------ copy support (environment_utils.copy_env)
    def __deepcopy__(self, memo):
        new = RendezvousEnv.__new__(RendezvousEnv)
        memo[id(self)] = new
        skip = {"_b", "_h_f64", "_h_i32", "_h_act64", "_h_act32", "_d_act64", "_d_act32", "_h_obs", "_h_rew",
                "_h_done", "_h_reason", "_h_misc", "_d_vec", "_h_vec"}
        import copy
        for k, v in self.__dict__.items():
            if k not in skip:
                new.__dict__[k] = copy.deepcopy(v, memo)
        new._b = self._b.clone()
        new._alloc_staging()
        return new
