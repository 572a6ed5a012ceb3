"""Headless restatements of the reference's manual verification scripts, with numeric results instead of
plots, run against the CUDA kernels through the single-env facade:

* ``verify_cw``               verification/verify_cw.py:12-74, :179-207
* ``verify_cw2``              verification/verify_cw2.py:13-64
* ``verify_attitude``         verification/verify_attitude.py:13-113 (env vs an independent rigid-body propagation)
* ``verify_attitude_torque``  verification/verify_attitude_torque.py:34-57, :141-146
* ``verify_attitude_racket``  verification/verify_attitude_racket.py:34-76
* ``initial_state_distribution``  verification/initial_state_distribution.py:16-123

Each returns a dict of numbers; tests/test_gpu_verification.py asserts the pass criteria of SURVEY.md section 4.
The scripts ignore ``done`` and so do these.
"""
from __future__ import annotations

import numpy as np

from .batched_env import BatchedRendezvousEnv
from .rendezvous_env import RendezvousEnv

NO_RANGE = dict(rc0_range=0, vc0_range=0, qc0_range=0, wc0_range=0, qt0_range=0, wt0_range=0)


def cw_closed_form(r0, v0, n, t):
    """Analytic Clohessy-Wiltshire solution the script compares with (utils/dynamics.py:24-55) -- the
    independent check value, evaluated on the host in one shot from t = 0."""
    nt, s, c = n * t, np.sin(n * t), np.cos(n * t)
    rr = np.array([[4 - 3 * c, 0, 0], [6 * (s - nt), 1, 0], [0, 0, c]])
    rv = np.array([[s / n, 2 / n * (1 - c), 0], [-2 / n * (1 - c), (4 * s - 3 * nt) / n, 0], [0, 0, s / n]])
    vr = np.array([[3 * n * s, 0, 0], [-6 * n * (1 - c), 0, 0], [0, 0, -n * s]])
    vv = np.array([[c, 2 * s, 0], [-2 * s, 4 * c - 3, 0], [0, 0, c]])
    return rr @ r0 + rv @ v0, vr @ r0 + vv @ v0


def _state(env):
    return np.hstack([env.rc, env.vc, env.qc, env.wc, env.qt, env.wt])


def verify_cw(steps=755, device="cuda"):
    env = RendezvousEnv(rc0=np.array([0., -10., 1.]), vc0=np.array([-0.01, 0.01, 0.]), dt=1, quiet=True,
                        device=device, **NO_RANGE)
    env.reset()
    r0, v0 = env.rc.copy(), env.vc.copy()
    rs, vs = [], []
    for _ in range(steps):
        env.step(np.zeros(6))
        rs.append(env.rc.copy())
        vs.append(env.vc.copy())
    rs, vs = np.array(rs), np.array(vs)
    ana = [cw_closed_form(r0, v0, env.n, (k + 1) * env.dt) for k in range(steps)]
    ra, va = np.array([a[0] for a in ana]), np.array([a[1] for a in ana])
    return dict(r=rs, v=vs, r_analytic=ra, v_analytic=va,
                max_pos_diff=float(np.linalg.norm(rs - ra, axis=1).max()),
                max_vel_diff=float(np.linalg.norm(vs - va, axis=1).max()))


def verify_cw2(steps=120, device="cuda"):
    env = RendezvousEnv(rc0=np.array([0., -120., 0.]), vc0=np.array([0., 2., 0.]), dt=0.5, quiet=True,
                        device=device, **NO_RANGE)
    env.reset()
    # thrust that cancels the Coriolis term for vy0 = 2 m/s: f = 2 n m vy0, action = -f/10  (verify_cw2.py:22-26)
    f = 2 * env.n * 100 * 2
    action = np.array([-f / 10, 0, 0, 0, 0, 0])
    states = []
    for _ in range(steps):
        env.step(action)
        states.append(_state(env))
    states = np.array(states)
    return dict(state=states, action=action, max_vy_dev=float(np.abs(states[:, 4] - 2.0).max()))


def verify_attitude_torque(steps=65, device="cuda"):
    env = RendezvousEnv(dt=0.5, quiet=True, device=device, **NO_RANGE)
    env.reset()
    action = np.array([0, 0, 0, 0, 0, 0.5])
    states = []
    for _ in range(steps):
        env.step(action)
        states.append(_state(env))
    states = np.array(states)
    w_dot = action[5] * env.max_delta_w / env.dt            # equivalent constant angular acceleration
    t = steps * env.dt
    theta = 2 * np.arctan2(np.linalg.norm(states[-1, 7:10]), states[-1, 6])
    return dict(state=states, w_final=states[-1, 10:13], w_ideal=w_dot * t, theta=float(theta),
                theta_ideal=0.5 * w_dot * t ** 2, dt=env.dt)


def verify_attitude_racket(steps=760, device="cuda"):
    env = RendezvousEnv(wc0=np.radians(np.array([0, 5, 0.01])), dt=1, quiet=True, device=device, **NO_RANGE)
    env.reset()
    w0 = env.wc.copy()
    states = []
    for _ in range(steps):
        env.step(np.zeros(6))
        states.append(_state(env))
    states = np.array(states)
    return dict(state=states, wc_drift=float(np.abs(states[:, 10:13] - w0).max()))


def rigid_body_reference(q0, w0, inertia, torque, dt, steps, substeps=200):
    """Independent propagation of q' = 0.5 Omega(w) q, I w' = tau - w x I w with classical RK4 at dt/substeps
    (the role RigidBody.integrate plays in verify_attitude.py:13-66).  Host fp64, check value only."""
    inertia = np.asarray(inertia, dtype=float)
    inv = np.linalg.inv(inertia)
    tau = np.asarray(torque, dtype=float)

    def f(y):
        q, w = y[:4] / np.linalg.norm(y[:4]), y[4:]
        om = np.array([[0, -w[0], -w[1], -w[2]], [w[0], 0, w[2], -w[1]], [w[1], -w[2], 0, w[0]],
                       [w[2], w[1], -w[0], 0]])
        return np.hstack([0.5 * om @ q, inv @ (tau - np.cross(w, inertia @ w))])

    y = np.hstack([q0, w0]).astype(float)
    out = []
    h = dt / substeps
    for _ in range(steps):
        for _ in range(substeps):
            k1 = f(y)
            k2 = f(y + 0.5 * h * k1)
            k3 = f(y + 0.5 * h * k2)
            k4 = f(y + h * k3)
            y = y + h / 6 * (k1 + 2 * k2 + 2 * k3 + k4)
        y[:4] /= np.linalg.norm(y[:4])
        out.append(y.copy())
    return np.array(out)


def verify_attitude(steps=200, inertia=None, torque=None, wc0=None, device="cuda"):
    """Env attitude propagation (RK45 replica on the GPU) vs the independent rigid-body integration above, for
    a general diagonal inertia and a held body torque -- the cases the reference's isotropic env cannot
    exercise but its RigidBody check was written for."""
    inertia = np.diag([10.0, 16.0, 22.0]) if inertia is None else np.asarray(inertia, dtype=float)
    torque = np.zeros(3) if torque is None else np.asarray(torque, dtype=float)
    wc0 = np.radians(np.array([1.0, 5.0, 0.5])) if wc0 is None else np.asarray(wc0, dtype=float)
    env = RendezvousEnv(wc0=wc0, dt=1, quiet=True, device=device, inertia=inertia, chaser_torque=torque, **NO_RANGE)
    env.reset()
    q0, w0 = env.qc.copy(), env.wc.copy()
    states = []
    for _ in range(steps):
        env.step(np.zeros(6))
        states.append(np.hstack([env.qc, env.wc]))
    states = np.array(states)
    ref = rigid_body_reference(q0, w0, inertia, torque, 1.0, steps)
    sign = np.sign(np.sum(states[:, :4] * ref[:, :4], axis=1, keepdims=True))
    return dict(state=states, reference=ref,
                max_q_diff=float(np.abs(states[:, :4] - sign * ref[:, :4]).max()),
                max_w_diff=float(np.abs(states[:, 4:] - ref[:, 4:]).max()))


def initial_state_distribution(num_samples=100_000, seed=0, device="cuda"):
    """verification/initial_state_distribution.py:16-123: magnitudes of the six initial-state deviations over
    ``num_samples`` resets -> min / max / mean / variance and the uniform expectations range/2, range^2/12."""
    env = BatchedRendezvousEnv(num_samples, device=device, seed=seed, track_stats=False)
    env.reset()
    s = env.get_state().cpu().numpy()
    p = env.params
    mags = dict(
        rc=(np.linalg.norm(s[:, 0:3] - np.array(p.rc0[:]), axis=1), p.rc0_range),
        vc=(np.linalg.norm(s[:, 3:6] - np.array(p.vc0[:]), axis=1), p.vc0_range),
        qc=(2 * np.arccos(np.clip(np.abs(s[:, 6]), 0, 1)), p.qc0_range),
        wc=(np.linalg.norm(s[:, 10:13], axis=1), p.wc0_range),
        qt=(2 * np.arccos(np.clip(np.abs(s[:, 13]), 0, 1)), p.qt0_range),
        wt=(np.linalg.norm(s[:, 17:20], axis=1), p.wt0_range),
    )
    return {k: dict(min=float(m.min()), max=float(m.max()), mean=float(m.mean()), var=float(m.var()),
                    expected_mean=r / 2, expected_var=r * r / 12, range=r) for k, (m, r) in mags.items()}
