"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the RendezvousEnv step/reset hot path.

This file is a from-scratch numpy restatement of the reference algorithm.  It
is the *checker*: only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / `--impl reference` leg may import it.  The product package
(reinforcement_learning_rendezvous_b200/) never does, and fails loudly when its
CUDA library is missing instead of falling back to this.

Parity status: PINNED.  tests/test_oracle_vs_reference.py drives the unmodified
reference (imported through oracle/ref_stubs.py) and this oracle with the same
states and actions and requires bit-for-bit equal trajectories; the frozen
outputs of the reference live in tests/golden/*.npz (made by
oracle/make_golden.py) so the pin also holds on the GPU box, where
/root/reference does not exist.

Every function cites the reference lines it restates (paths relative to
/root/reference).  The adaptive integrator lives in a third-party dependency
that is NOT vendored by the reference: scipy.integrate.solve_ivp(method="RK45")
(SciPy version unpinned by the reference; restated from SciPy 1.18.1,
scipy/integrate/_ivp/{rk.py,common.py,ivp.py,base.py}).  `rk45_restated` below
follows that published algorithm; `integrator="scipy"` calls SciPy itself, as
the reference does, and the tests require both to agree bit-for-bit.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np

# ----------------------------------------------------------------------------
# Dormand-Prince 5(4) tableau: scipy/integrate/_ivp/rk.py, class RK45 body.
# ----------------------------------------------------------------------------
RK_C = np.array([0, 1 / 5, 3 / 10, 4 / 5, 8 / 9, 1])
RK_A = np.array([
    [0, 0, 0, 0, 0],
    [1 / 5, 0, 0, 0, 0],
    [3 / 40, 9 / 40, 0, 0, 0],
    [44 / 45, -56 / 15, 32 / 9, 0, 0],
    [19372 / 6561, -25360 / 2187, 64448 / 6561, -212 / 729, 0],
    [9017 / 3168, -355 / 33, 46732 / 5247, 49 / 176, -5103 / 18656],
])
RK_B = np.array([35 / 384, 0, 500 / 1113, 125 / 192, -2187 / 6784, 11 / 84])
RK_E = np.array([-71 / 57600, 0, 71 / 16695, -71 / 1920, 17253 / 339200, -22 / 525, 1 / 40])
RK_P = np.array([
    [1, -8048581381 / 2820520608, 8663915743 / 2820520608, -12715105075 / 11282082432],
    [0, 0, 0, 0],
    [0, 131558114200 / 32700410799, -68118460800 / 10900136933, 87487479700 / 32700410799],
    [0, -1754552775 / 470086768, 14199869525 / 1410260304, -10690763975 / 1880347072],
    [0, 127303824393 / 49829197408, -318862633887 / 49829197408, 701980252875 / 199316789632],
    [0, -282668133 / 205662961, 2019193451 / 616988883, -1453857185 / 822651844],
    [0, 40617522 / 29380423, -110615467 / 29380423, 69997945 / 29380423],
])
RK_SAFETY, RK_MIN_FACTOR, RK_MAX_FACTOR = 0.9, 0.2, 10
RK_RTOL, RK_ATOL = 1e-7, 1e-6          # rendezvous_env.py:567-568, :594-595


def _rms(x):
    """scipy common.py:63-65 `norm`."""
    return np.linalg.norm(x) / x.size ** 0.5


class RK45Failure(RuntimeError):
    pass


def rk45_restated(fun, y0, t_bound, rtol=RK_RTOL, atol=RK_ATOL, counters=None):
    """solve_ivp(fun, (0, t_bound), y0, 'RK45', t_eval=[t_bound]).y[:, 0], restated.

    Control flow: RungeKutta.__init__ (rk.py:85-103), select_initial_step
    (common.py:68-134), _step_impl (rk.py:111-179), rk_step (rk.py:14-69), and
    the t_eval branch of solve_ivp (ivp.py:710-728) which evaluates the dense
    output polynomial (rk.py:715-737) at t_bound on the step that reaches it.
    `counters`, if given, receives nfev / accepted / rejected tallies.
    """
    y = np.asarray(y0, dtype=float)
    n = y.size
    nfev = 0
    t = 0.0
    f = fun(t, y)
    nfev += 1

    # -- select_initial_step (order = error_estimator_order = 4) --------------
    interval = abs(t_bound - t)
    scale = atol + np.abs(y) * rtol
    d0 = _rms(y / scale)
    d1 = _rms(f / scale)
    h0 = 1e-6 if (d0 < 1e-5 or d1 < 1e-5) else 0.01 * d0 / d1
    h0 = min(h0, interval)
    y1 = y + h0 * 1.0 * f
    f1 = fun(t + h0 * 1.0, y1)
    nfev += 1
    d2 = _rms((f1 - f) / scale) / h0
    if d1 <= 1e-15 and d2 <= 1e-15:
        h1 = max(1e-6, h0 * 1e-3)
    else:
        h1 = (0.01 / max(d1, d2)) ** (1 / (4 + 1))
    h_abs = min(100 * h0, h1, interval, np.inf)

    err_exp = -1 / (4 + 1)
    K = np.empty((7, n), dtype=float)
    accepted = rejected = 0
    while True:
        # -- one solver.step(): _step_impl --------------------------------------
        min_step = 10 * np.abs(np.nextafter(t, np.inf) - t)
        if h_abs < min_step:
            h_abs = min_step
        was_rejected = False
        while True:
            if h_abs < min_step:
                raise RK45Failure("TOO_SMALL_STEP")
            h = h_abs
            t_new = t + h
            if t_new - t_bound > 0:
                t_new = t_bound
            h = t_new - t
            h_abs = np.abs(h)
            # rk_step
            K[0] = f
            for s in range(1, 6):
                dy = np.dot(K[:s].T, RK_A[s, :s]) * h
                K[s] = fun(t + RK_C[s] * h, y + dy)
            y_new = y + h * np.dot(K[:-1].T, RK_B)
            f_new = fun(t + h, y_new)
            nfev += 6
            K[-1] = f_new
            sc = atol + np.maximum(np.abs(y), np.abs(y_new)) * rtol
            err = _rms(np.dot(K.T, RK_E) * h / sc)
            if err < 1:
                factor = RK_MAX_FACTOR if err == 0 else min(RK_MAX_FACTOR, RK_SAFETY * err ** err_exp)
                if was_rejected:
                    factor = min(1, factor)
                h_abs *= factor
                accepted += 1
                break
            h_abs *= max(RK_MIN_FACTOR, RK_SAFETY * err ** err_exp)
            was_rejected = True
            rejected += 1
        t_old, y_old = t, y
        t, y, f = t_new, y_new, f_new
        if t - t_bound >= 0:
            # dense output at t_eval = [t_bound]  (x == 1 on this step)
            Q = K.T.dot(RK_P)
            hh = t - t_old
            x = (t_bound - t_old) / hh
            p = np.cumprod(np.tile(x, 4))
            out = hh * np.dot(Q, p) + y_old
            if counters is not None:
                counters["nfev"] = counters.get("nfev", 0) + nfev
                counters["accepted"] = counters.get("accepted", 0) + accepted
                counters["rejected"] = counters.get("rejected", 0) + rejected
                counters["calls"] = counters.get("calls", 0) + 1
            return out


# ----------------------------------------------------------------------------
# Quaternion / vector helpers
# ----------------------------------------------------------------------------
def rotation_matrix(q):
    """utils/quaternions.py:48-68 quat2mat (re-normalises q first, :57)."""
    q = q / np.linalg.norm(q)
    w, x, y, z = q
    return np.array([
        [2 * (w ** 2 + x ** 2) - 1, 2 * (x * y - w * z), 2 * (x * z + w * y)],
        [2 * (x * y + w * z), 2 * (w ** 2 + y ** 2) - 1, 2 * (y * z - w * x)],
        [2 * (x * z - w * y), 2 * (y * z + w * x), 2 * (w ** 2 + z ** 2) - 1],
    ])


def axis_angle_quat(axis, theta):
    """utils/quaternions.py:11-27 rot2quat."""
    axis = axis / np.linalg.norm(axis)
    q = np.append(np.cos(theta / 2), axis * np.sin(theta / 2))
    return q / np.linalg.norm(q)


def hamilton(q1, q2):
    """utils/quaternions.py:149-170 quat_product (normalises inputs, not output)."""
    q1 = q1 / np.linalg.norm(q1)
    q2 = q2 / np.linalg.norm(q2)
    s1, v1, s2, v2 = q1[0], q1[1:], q2[0], q2[1:]
    return np.append(s1 * s2 - np.dot(v1, v2), s1 * v2 + s2 * v1 + np.cross(v1, v2))


def rounded_angle(v1, v2):
    """utils/general.py:163-181 angle_between_vectors: acos of the cosine rounded
    to 5 decimals (np.float64.__round__ == rint(x*1e5)/1e5)."""
    c = np.dot(v1, v2) / (np.linalg.norm(v1) * np.linalg.norm(v2))
    return np.arccos(round(c, 5))


def to_unit_range(val, low, high):
    """utils/general.py:230-245 normalize_value with custom_range [-1, 1]."""
    a, b = -1, 1
    return (b - a) * (val - low) / (high - low) + a


def attitude_rhs(_t, y, inertia, inv_inertia, torque):
    """utils/dynamics.py:93-119 (+ quat_derivative :122-153, angular_acceleration
    :156-175).  q is normalised twice (:108 and :134), as in the reference."""
    q = y[0:4]
    w = y[4:]
    q = q / np.linalg.norm(q)
    q = q / np.linalg.norm(q)
    w1, w2, w3 = w
    omega = np.array([
        [0, -w1, -w2, -w3],
        [w1, 0, w3, -w2],
        [w2, -w3, 0, w1],
        [w3, w2, -w1, 0],
    ])
    q_dot = 0.5 * np.matmul(omega, q)
    w_dot = np.matmul(inv_inertia, torque - np.cross(w, np.matmul(inertia, w)))
    return np.append(q_dot, w_dot)


def cw_propagate(r0, v0, n, t):
    """utils/dynamics.py:24-55 closed-form Clohessy-Wiltshire transition.
    Note `1 / n*np.sin(nt)` parses as (1/n)*sin(nt)."""
    nt = n * t
    s, c = np.sin(nt), np.cos(nt)
    stm = np.array([
        [4 - 3 * c, 0, 0, 1 / n * s, 2 / n * (1 - c), 0],
        [6 * (s - nt), 1, 0, -2 / n * (1 - c), 1 / n * (4 * s - 3 * nt), 0],
        [0, 0, c, 0, 0, 1 / n * s],
        [3 * n * s, 0, 0, c, 2 * s, 0],
        [-6 * n * (1 - c), 0, 0, -2 * s, 4 * c - 3, 0],
        [0, 0, -n * s, 0, 0, c],
    ])
    xt = np.matmul(stm, np.append(r0, v0))
    return xt[0:3], xt[3:]


# ----------------------------------------------------------------------------
# Environment constants: rendezvous_env.py:17-158
# ----------------------------------------------------------------------------
@dataclass
class OracleParams:
    rc0: np.ndarray = field(default_factory=lambda: np.array([0., -10., 0.]))
    vc0: np.ndarray = field(default_factory=lambda: np.array([0., 0., 0.]))
    qc0: np.ndarray = field(default_factory=lambda: np.array([1., 0., 0., 0.]))
    wc0: np.ndarray = field(default_factory=lambda: np.array([0., 0., 0.]))
    qt0: np.ndarray = field(default_factory=lambda: np.array([1., 0., 0., 0.]))
    wt0: np.ndarray = field(default_factory=lambda: np.array([0., 0., 0.]))
    rc0_range: float = 1
    vc0_range: float = 0.1
    qc0_range: float = float(np.radians(1))
    wc0_range: float = float(np.radians(0.1))
    qt0_range: float = float(np.radians(45))
    wt0_range: float = float(np.radians(3))
    koz_radius: float = 5
    corridor_half_angle: float = float(np.radians(30))
    h: float = 800e3
    dt: float = 1
    t_max: float = 120
    reward_kwargs: dict = field(default_factory=dict)
    inertia: np.ndarray = None            # chaser (overridable for the verify_attitude ports)
    inertia_target: np.ndarray = None
    chaser_torque: np.ndarray = field(default_factory=lambda: np.array([0, 0, 0]))

    @classmethod
    def from_kwargs(cls, **kw):
        """None means 'use the default', exactly like the reference ctor."""
        return cls(**{k: v for k, v in kw.items() if v is not None})

    def __post_init__(self):
        self.m = 100
        iso = np.array([[1, 0, 0], [0, 1, 0], [0, 0, 1]]) * 1 / 12 * self.m * (2 * 1 ** 2)   # :75-79
        if self.inertia is None:
            self.inertia = iso
        if self.inertia_target is None:
            self.inertia_target = iso.copy()
        self.inv_inertia = np.linalg.inv(self.inertia)
        self.inv_inertia_target = np.linalg.inv(self.inertia_target)
        self.max_delta_v = 10 / self.m * 0.5                     # :81
        self.max_delta_w = 0.2 / iso[0, 0] * 0.5                 # :82 (always from the nominal inertia)
        self.max_axial_distance = np.linalg.norm(self.rc0) + 10  # :85
        self.max_axial_speed = 5
        self.max_wc = np.radians(10)
        self.max_wt = np.radians(10)
        self.max_attitude_error = np.radians(30)
        self.capture_axis = np.array([0, 1, 0])
        self.corridor_axis = np.array([0, -1, 0])
        self.rd = np.array([0, -2, 0])
        self.max_rd_error = 0.5
        self.max_vd_error = 0.1
        self.max_qd_error = np.radians(5)
        self.max_wd_error = np.radians(1)
        self.bubble_radius0 = self.max_axial_distance            # :113
        self.bubble_decrease_rate = 0.5 * self.dt                # :114
        self.bubble_min = np.linalg.norm(self.rd) + 2 * self.max_rd_error
        self.mu = 3.986004418e14
        self.Re = 6371e3
        self.ro = self.Re + self.h
        self.n = np.sqrt(self.mu / self.ro ** 3)                 # :126
        assert np.linalg.norm(self.rd) < self.koz_radius         # :155
        assert np.linalg.norm(self.rd) - self.max_rd_error > 0   # :156


END_REASONS = ("obs", "time", "bubble", "attitude")    # rendezvous_env.py:377


class OracleEnv:
    """Single-environment oracle with the reference's attribute names."""

    def __init__(self, params: OracleParams = None, integrator="restated", rng=None, **kwargs):
        self.p = params if params is not None else OracleParams.from_kwargs(**kwargs)
        self.integrator = integrator
        self.rng = rng if rng is not None else np.random
        self.rc = self.vc = self.qc = self.wc = self.qt = self.wt = None
        self.t = None
        self.collided = None
        self.success = None
        self.bubble_radius = None
        self.total_delta_v = None
        self.total_delta_w = None
        self.end_reason = -1
        self.rk_counters = {}
        self.obs_low = np.full(17, -1, dtype=np.float32)
        self.obs_high = np.full(17, 1, dtype=np.float32)

    # -- convenience passthroughs used by evaluators ---------------------------
    def __getattr__(self, item):
        p = self.__dict__.get("p")
        if p is not None and hasattr(p, item):
            return getattr(p, item)
        raise AttributeError(item)

    # -- frame transforms: rendezvous_env.py:470-508 ----------------------------
    def chaser2lvlh(self, v):
        return np.matmul(rotation_matrix(self.qc), v)

    def target2lvlh(self, v):
        return np.matmul(rotation_matrix(self.qt), v)

    def lvlh2chaser(self, v):
        return np.matmul(rotation_matrix(self.qc).T, v)

    def lvlh2target(self, v):
        return np.matmul(rotation_matrix(self.qt).T, v)

    # -- reset: rendezvous_env.py:223-270 ----------------------------------------
    def _uniform(self, low, high, size=None):
        return self.rng.uniform(low=low, high=high, size=size)

    def _unit_vector(self):
        """utils/general.py:248-254: cube sample, then normalise."""
        v = self._uniform(-1, 1, (3,))
        return v / np.linalg.norm(v)

    def reset(self, uniforms=None):
        """`uniforms`: optional 24 numbers in [0,1) consumed in the reference's
        draw order (3+1, 3+1, 1, 3, 3+1, 1, 3, 3+1); each is mapped with
        numpy's own `low + (high-low)*u`.  Without it the draws come from
        self.rng (np.random by default, like the reference)."""
        if uniforms is not None:
            it = iter(np.asarray(uniforms, dtype=float).tolist())

            def uni(low, high, size=None):
                if size is None:
                    return low + (high - low) * next(it)
                return np.array([low + (high - low) * next(it) for _ in range(int(np.prod(size)))])
            self._uniform = uni
        p = self.p
        try:
            rc_dev = self._unit_vector() * self._uniform(0, p.rc0_range)
            vc_dev = self._unit_vector() * self._uniform(0, p.vc0_range)
            theta_c = self._uniform(0, p.qc0_range)
            qc_dev = axis_angle_quat(self._unit_vector(), theta_c)
            wc_dev = self._unit_vector() * self._uniform(0, p.wc0_range)
            theta_t = self._uniform(0, p.qt0_range)
            qt_dev = axis_angle_quat(self._unit_vector(), theta_t)
            wt_dev = self._unit_vector() * self._uniform(0, p.wt0_range)
        finally:
            self.__dict__.pop("_uniform", None)
        self.rc = p.rc0 + rc_dev
        self.vc = p.vc0 + vc_dev
        self.qc = hamilton(qc_dev, p.qc0)
        self.wc = self.lvlh2chaser(p.wc0 + wc_dev)
        self.qt = hamilton(qt_dev, p.qt0)
        self.wt = self.lvlh2target(p.wt0 + wt_dev)
        self.collided = self.check_collision()
        self.success = int(self.check_success())
        self.bubble_radius = p.bubble_radius0
        self.total_delta_v = 0
        self.total_delta_w = 0
        self.t = 0
        self.end_reason = -1
        return self.get_observation()

    # -- step: rendezvous_env.py:160-221 ------------------------------------------
    def _integrate(self, q, w, inertia, inv_inertia, torque):
        """rendezvous_env.py:552-604."""
        y0 = np.append(q, w)
        dt = self.p.dt
        if self.integrator == "scipy":
            from scipy.integrate import solve_ivp
            sol = solve_ivp(fun=attitude_rhs, t_span=(0, dt), y0=y0, method="RK45",
                            t_eval=np.array([dt]), rtol=RK_RTOL, atol=RK_ATOL,
                            args=(inertia, inv_inertia, torque))
            yf = sol.y.flatten()
        else:
            yf = rk45_restated(lambda t, y: attitude_rhs(t, y, inertia, inv_inertia, torque),
                               y0, dt, counters=self.rk_counters)
        qn = yf[0:4]
        return qn / np.linalg.norm(qn), yf[4:]

    def step(self, action):
        p = self.p
        a = action.copy()
        delta_v = self.chaser2lvlh(a[0:3] * p.max_delta_v)
        delta_w = a[3:] * p.max_delta_w
        self.rc, self.vc = cw_propagate(self.rc, self.vc + delta_v, p.n, p.dt)
        self.wc = self.wc + delta_w
        self.qc, self.wc = self._integrate(self.qc, self.wc, p.inertia, p.inv_inertia, p.chaser_torque)
        self.qt, self.wt = self._integrate(self.qt, self.wt, p.inertia_target, p.inv_inertia_target,
                                           np.array([0, 0, 0]))
        if not self.collided:
            self.collided = self.check_collision()
            if self.check_success():
                self.success += 1
        self.t = round(self.t + p.dt, 3)
        self.bubble_radius -= p.bubble_decrease_rate
        if self.bubble_radius < p.bubble_min:
            self.bubble_radius = p.bubble_min
        self.total_delta_v += np.abs(a[0:3]).sum() * p.max_delta_v
        self.total_delta_w += np.abs(a[3:]).sum() * p.max_delta_w
        obs = self.get_observation()
        done = self.get_done_condition(obs)
        rew = self.get_bubble_reward(a, **p.reward_kwargs)
        return obs, rew, done, {"observation": obs, "reward": rew, "done": done, "action": action}

    # -- observation: rendezvous_env.py:294-311 -------------------------------------
    def get_observation(self):
        p = self.p
        obs = np.hstack((
            to_unit_range(self.rc, -p.max_axial_distance, p.max_axial_distance),
            to_unit_range(self.vc, -p.max_axial_speed, p.max_axial_speed),
            self.qc,
            to_unit_range(self.wc, -p.max_wc, p.max_wc),
            self.qt,
        ))
        return obs.astype(np.float32)

    # -- reward: rendezvous_env.py:313-353 --------------------------------------------
    def get_bubble_reward(self, action, collision_coef=0.5, bonus_coef=8, fuel_coef=0.2, att_coef=1):
        p = self.p
        rew = 0
        rew += p.dt * att_coef * (1 - self.get_attitude_error() / p.max_attitude_error)
        rew += p.dt * fuel_coef * np.abs(action[0:3]).sum() / (3 * p.max_delta_v)
        if self.check_collision():
            rew -= p.dt * collision_coef
        if np.linalg.norm(self.rc) < p.koz_radius and not self.collided:
            pos_error, _vel, att_error, _rot = self.get_errors()
            if pos_error < p.max_rd_error:
                rew += p.dt * bonus_coef * (2 - pos_error / p.max_rd_error)
                if att_error < p.max_qd_error:
                    rew += p.dt * bonus_coef * (2 - att_error / p.max_qd_error)
        return rew

    # -- termination: rendezvous_env.py:355-386 -----------------------------------------
    def get_done_condition(self, obs):
        p = self.p
        inside = bool(np.can_cast(obs.dtype, np.float32) and obs.shape == (17,)
                      and np.all(obs >= self.obs_low) and np.all(obs <= self.obs_high))
        conds = [
            not inside,
            self.t >= p.t_max,
            np.linalg.norm(self.rc) > self.bubble_radius,
            self.get_attitude_error() > p.max_attitude_error,
        ]
        self.end_reason = conds.index(True) if any(conds) else -1
        return bool(any(conds))

    # -- collision / success / errors: rendezvous_env.py:388-468 ---------------------------
    def check_collision(self):
        p = self.p
        if np.linalg.norm(self.rc) < p.koz_radius:
            if rounded_angle(self.rc, self.target2lvlh(p.corridor_axis)) > p.corridor_half_angle:
                return True
        return False

    def check_success(self):
        p = self.p
        if self.collided:
            return 0
        limits = np.array([p.max_rd_error, p.max_vd_error, p.max_qd_error, p.max_wd_error])
        return 1 if np.all(self.get_errors() <= limits) else 0

    def get_attitude_error(self):
        return rounded_angle(-self.rc, self.chaser2lvlh(self.p.capture_axis))

    def get_errors(self):
        p = self.p
        wc_l = self.chaser2lvlh(self.wc)
        wt_l = self.target2lvlh(self.wt)
        rd_l = self.target2lvlh(p.rd)
        vd_l = np.cross(wt_l, rd_l)
        return np.array([
            np.linalg.norm(self.rc - rd_l),
            np.linalg.norm(self.vc - vd_l),
            self.get_attitude_error(),
            np.linalg.norm(wc_l - wt_l),
        ])

    # -- evaluator helper: rendezvous_env.py:510-537 -------------------------------------------
    def dist_from_koz(self):
        p = self.p
        r = np.linalg.norm(self.rc)
        rk, th_c = p.koz_radius, p.corridor_half_angle
        th = rounded_angle(self.rc, self.target2lvlh(p.corridor_axis))
        if r < rk:
            if th >= th_c:
                return -1 * min(rk - r, r * np.sin(min(th - th_c, np.pi / 2)))
            return r * np.sin(th_c - th)
        if th >= th_c:
            return r - rk
        d_rad = r - rk * np.cos(th_c - th)
        d_tan = rk * np.sin(th_c - th)
        return np.sqrt(d_rad ** 2 + d_tan ** 2)

    # -- state I/O for batch drivers ---------------------------------------------------------------
    def set_state(self, row20):
        row20 = np.asarray(row20, dtype=float)
        self.rc, self.vc = row20[0:3].copy(), row20[3:6].copy()
        self.qc, self.wc = row20[6:10].copy(), row20[10:13].copy()
        self.qt, self.wt = row20[13:17].copy(), row20[17:20].copy()

    def get_state(self):
        return np.hstack((self.rc, self.vc, self.qc, self.wc, self.qt, self.wt)).astype(float)


def make_oracle_env(reward_kwargs=None, config=None, stochastic=True, integrator="restated", rng=None):
    """utils/environment_utils.py:9-63 make_env, without the printing."""
    config = dict(config or {})
    if stochastic is False:
        for k in ("rc0_range", "vc0_range", "qc0_range", "wc0_range", "qt0_range", "wt0_range"):
            config[k] = 0
    rc0 = config.get("rc0")
    if rc0 is not None and not isinstance(rc0, np.ndarray):
        rc0 = np.array([0., -rc0, 0.])
    wt0 = config.get("wt0")
    if wt0 is not None and not isinstance(wt0, np.ndarray):
        wt0 = np.array([0., 0., wt0])
    keys = ("vc0", "qc0", "wc0", "qt0", "rc0_range", "vc0_range", "qc0_range", "wc0_range",
            "qt0_range", "wt0_range", "koz_radius", "corridor_half_angle", "h", "dt", "t_max")
    kw = {k: config.get(k) for k in keys}
    kw.update(rc0=rc0, wt0=wt0, reward_kwargs=reward_kwargs)
    return OracleEnv(OracleParams.from_kwargs(**kw), integrator=integrator, rng=rng)


def evaluate_episode(policy, env, initial_state):
    """monte_carlo.py:94-207 `evaluate`, restated for the oracle env."""
    p = env.p
    n_slots = int(p.t_max / p.dt) + 1
    errors = np.full((4, n_slots), np.nan)
    times = np.full(n_slots, np.nan)
    num_collisions = num_successes = 0
    total_reward = 0
    env.reset()
    for k in ("rc", "vc", "qc", "wc", "qt", "wt"):
        setattr(env, k, initial_state[k])
    obs = env.get_observation()
    errors[:, 0] = env.get_errors()
    times[0] = env.t
    num_collisions += int(env.check_collision())
    if not env.collided:
        num_successes += int(env.check_success())
    min_dist = env.dist_from_koz()
    k = 1
    done = False
    while not done:
        action, _ = policy.predict(observation=obs, state=None, episode_start=None, deterministic=True)
        obs, reward, done, _info = env.step(action)
        errors[:, k] = env.get_errors()
        times[k] = env.t
        num_collisions += int(env.check_collision())
        if not env.collided:
            num_successes += int(env.check_success())
        min_dist = min(min_dist, env.dist_from_koz())
        total_reward += reward
        k += 1
    if env.t < p.t_max:
        times = times[~np.isnan(times)]
        errors = errors[:, 0:times.size]
    pos, vel, att, rot = errors
    m_pos, m_vel = pos < p.max_rd_error, vel < p.max_vd_error
    m_att, m_rot = att < p.max_qd_error, rot < p.max_wd_error
    for mask in (m_pos * m_vel * m_att * m_rot,
                 m_pos * m_vel * m_att + m_pos * m_vel * m_rot,
                 m_pos * m_vel,
                 m_pos):
        if np.any(mask):
            index = int(np.argmax(mask))
            break
    else:
        index = -1
    return dict(
        ep_len=times[-1],
        num_collisions=num_collisions,
        collided=int(num_collisions > 0),
        total_reward=total_reward,
        total_delta_v=env.total_delta_v,
        num_successes=num_successes,
        succeeded=int(num_successes > 0),
        min_dist_from_koz=min_dist,
        pos_error=pos[index:].mean(),
        vel_error=vel[index:].mean(),
        att_error=np.degrees(att[index:].mean()),
        rot_error=np.degrees(rot[index:].mean()),
    )
