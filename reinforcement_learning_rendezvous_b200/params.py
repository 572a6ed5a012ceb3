"""Constructor arguments of the reference ``RendezvousEnv`` -> the ``RdvParams`` block the kernels read.

The keyword names, defaults and derived constants are those of
``RendezvousEnv.__init__`` (/root/reference/rendezvous_env.py:17-158); the
defaults and the derivation themselves live in the C ABI
(``rdv_params_default`` / ``rdv_params_derive``), so Python and any other host
language get identical constants.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _native as N

VECTOR_KEYS = {"rc0": 3, "vc0": 3, "qc0": 4, "wc0": 3, "qt0": 4, "wt0": 3}
RANGE_KEYS = ("rc0_range", "vc0_range", "qc0_range", "wc0_range", "qt0_range", "wt0_range")
SCALAR_KEYS = RANGE_KEYS + ("koz_radius", "corridor_half_angle", "h", "dt", "t_max")
REWARD_KEYS = ("collision_coef", "bonus_coef", "fuel_coef", "att_coef")     # rendezvous_env.py:313
INTEGRATORS = {"rk45": N.INTEGRATOR_RK45, "closed_form": N.INTEGRATOR_CLOSED_FORM}


def make_params(reward_kwargs=None, inertia=None, inertia_target=None, chaser_torque=None, integrator="rk45",
                **ctor_kwargs) -> N.RdvParams:
    """Build and derive an RdvParams from RendezvousEnv constructor kwargs (``None`` = default).

    ``inertia`` / ``inertia_target`` / ``chaser_torque`` expose what the reference hard-codes
    (rendezvous_env.py:75-79, :96-100, :558) so the verify_attitude*-style checks can use a
    general rigid body.  Raises ValueError where the reference constructor's asserts
    (rendezvous_env.py:155-156) would fail.
    """
    L = N.lib()
    p = N.RdvParams()
    L.rdv_params_default(C.byref(p))
    for key, value in ctor_kwargs.items():
        if key == "quiet" or value is None:
            continue
        if key in VECTOR_KEYS:
            arr = np.asarray(value, dtype=np.float64).ravel()
            if arr.size != VECTOR_KEYS[key]:
                raise ValueError(f"{key} must have {VECTOR_KEYS[key]} components")
            getattr(p, key)[:] = arr.tolist()
        elif key in SCALAR_KEYS:
            setattr(p, key, float(value))
        else:
            raise TypeError(f"unexpected RendezvousEnv argument {key!r}")
    for key, value in (reward_kwargs or {}).items():
        if key not in REWARD_KEYS:
            raise TypeError(f"unexpected reward keyword {key!r}")
        setattr(p, key, float(value))
    if inertia is not None:
        p.inertia_c[:] = np.asarray(inertia, dtype=np.float64).reshape(9).tolist()
    if inertia_target is not None:
        p.inertia_t[:] = np.asarray(inertia_target, dtype=np.float64).reshape(9).tolist()
    if chaser_torque is not None:
        p.torque_c[:] = np.asarray(chaser_torque, dtype=np.float64).reshape(3).tolist()
    if integrator not in INTEGRATORS:
        raise ValueError(f"integrator must be one of {sorted(INTEGRATORS)}")
    p.integrator = INTEGRATORS[integrator]
    status = L.rdv_params_derive(C.byref(p))
    if status != 0:
        raise ValueError(f"invalid RendezvousEnv configuration: {N.strerror(status)}")
    return p


def params_to_dict(p: N.RdvParams) -> dict:
    out = {}
    for name, ctype in p._fields_:
        v = getattr(p, name)
        out[name] = np.array(v[:]) if hasattr(v, "__len__") else v
    return out


def copy_params(p: N.RdvParams) -> N.RdvParams:
    q = N.RdvParams()
    C.memmove(C.byref(q), C.byref(p), C.sizeof(N.RdvParams))
    return q
