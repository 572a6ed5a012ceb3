"""CPU: the C-ABI library loads and exports every symbol include/rdv_b200.h declares; the host-side pieces of
the ABI (defaults, derived constants, argument validation, error strings) behave like the reference
constructor.  No kernels are launched here."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from helpers import ROOT

from reinforcement_learning_rendezvous_b200 import _native as N
from reinforcement_learning_rendezvous_b200.params import make_params


@pytest.fixture(scope="module")
def lib():
    N.build()
    return N.lib()


def test_header_symbols_exported(lib):
    header = open(os.path.join(ROOT, "include", "rdv_b200.h")).read()
    declared = set(re.findall(r"\b(rdv_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 12
    assert declared == set(N.PROTOTYPES), declared ^ set(N.PROTOTYPES)
    raw = C.CDLL(N.LIB_PATH)
    for name in declared:
        assert hasattr(raw, name), name
    assert lib.rdv_abi_version() == int(re.search(r"#define RDV_ABI_VERSION (\d+)", header).group(1))
    assert lib.rdv_sizeof_params() == C.sizeof(N.RdvParams)
    for enum_name, value in (("RDV_NF64", N.NF64), ("RDV_NI32", N.NI32), ("RDV_NSTATS", N.NSTATS),
                             ("RDV_EP_NCOL", N.EP_NCOL)):
        assert enum_name in header and value > 0


def test_default_constants_match_reference_constructor(lib):
    """rendezvous_env.py:17-158 defaults, as printed by the reference."""
    p = make_params()
    assert p.max_delta_v == 0.05 and p.max_delta_w == 0.006000000000000001
    assert p.inertia_c[0] == 16.666666666666664 and p.inv_inertia_c[0] == 0.06000000000000001
    assert p.max_axial_distance == 20 and p.max_axial_speed == 5
    assert p.max_wc == np.radians(10) and p.max_attitude_error == np.radians(30)
    assert p.koz_radius == 5 and p.corridor_half_angle == np.radians(30)
    assert (p.max_rd_error, p.max_vd_error, p.max_qd_error, p.max_wd_error) == \
           (0.5, 0.1, np.radians(5), np.radians(1))
    assert p.bubble0 == 20 and p.bubble_rate == 0.5 and p.bubble_min == 3
    assert p.n == 0.001039679077003123 and p.dt == 1 and p.t_max == 120
    assert (p.rc0_range, p.vc0_range, p.qc0_range, p.wc0_range, p.qt0_range, p.wt0_range) == \
           (1, 0.1, np.radians(1), np.radians(0.1), np.radians(45), np.radians(3))
    assert (p.collision_coef, p.bonus_coef, p.fuel_coef, p.att_coef) == (0.5, 8, 0.2, 1)
    assert p.iso_c == 1 and p.iso_t == 1
    # CW transition matrix entries (utils/dynamics.py:40-47)
    n, nt = p.n, p.n * p.dt
    s, c = np.sin(nt), np.cos(nt)
    want = [4 - 3 * c, 1 / n * s, 2 / n * (1 - c), 6 * (s - nt), 1, -2 / n * (1 - c), 1 / n * (4 * s - 3 * nt), c,
            1 / n * s, 3 * n * s, c, 2 * s, -6 * n * (1 - c), -2 * s, 4 * c - 3, -n * s, c]
    np.testing.assert_allclose(p.cw[:], want, rtol=1e-15)


def test_sensitivity_parameters(lib):
    """orbit mean motion for the sensitivity altitudes."""
    want = {400e3: 0.0011331559073083758, 600e3: 0.001084741520136686, 800e3: 0.001039679077003123,
            1000e3: 0.0009976524445962423, 2000e3: 0.0008243333586326639}
    for h, n in want.items():
        assert abs(make_params(h=h).n - n) <= 1e-18
    p = make_params(rc0=np.array([0., -30., 0.]), dt=0.5)
    assert p.max_axial_distance == 40 and p.bubble0 == 40 and p.bubble_rate == 0.25
    assert make_params(reward_kwargs=dict(bonus_coef=4, fuel_coef=0.1)).bonus_coef == 4
    with pytest.raises(ValueError):
        make_params(koz_radius=2)                        # rendezvous_env.py:155 assert
    with pytest.raises(TypeError):
        make_params(not_an_argument=1)
    with pytest.raises(TypeError):
        make_params(reward_kwargs=dict(nope=1))
    aniso = make_params(inertia=np.diag([10.0, 16.0, 22.0]))
    assert aniso.iso_c == 0 and aniso.iso_t == 1
    np.testing.assert_allclose(np.array(aniso.inv_inertia_c[:]).reshape(3, 3), np.diag([1 / 10, 1 / 16, 1 / 22]))
    with pytest.raises(ValueError):
        make_params(inertia=np.diag([10.0, 16.0, 22.0]), integrator="closed_form")


def test_argument_validation_without_gpu(lib):
    """Bad arguments are rejected before any launch; with valid arguments and no GPU the call reports
    RDV_ERR_CUDA instead of silently computing on the host."""
    p = make_params()
    st = N.RdvState(None, None, 0)
    io = N.RdvStepIO()
    assert lib.rdv_step(C.byref(p), C.byref(st), C.byref(io), 4, 0, 0, None) == -1           # NULL
    assert lib.rdv_step(None, None, None, 4, 0, 0, None) == -1
    buf = (C.c_double * 4096)()
    base = C.addressof(buf)
    st = N.RdvState(base, base, 2)
    io = N.RdvStepIO(base, 1, 0, base, base, base, None, None, None, None)
    assert lib.rdv_step(C.byref(p), C.byref(st), C.byref(io), 4, 0, 0, None) == -2           # ld < n
    assert lib.rdv_step(C.byref(p), C.byref(st), C.byref(io), -1, 0, 0, None) == -2
    st = N.RdvState(base + 4, base, 8)
    assert lib.rdv_step(C.byref(p), C.byref(st), C.byref(io), 4, 0, 0, None) == -3           # misaligned
    st = N.RdvState(base, base, 8)
    assert lib.rdv_step(C.byref(p), C.byref(st), C.byref(io), 0, 0, 0, None) == 0            # n = 0: no-op
    io.auto_reset = 7
    assert lib.rdv_step(C.byref(p), C.byref(st), C.byref(io), 4, 0, 0, None) == -2           # bad enum value
    assert lib.rdv_policy_forward(None, None, None, 1, None) == -1
    assert lib.rdv_fp64_peak_probe(None, 1, 1, 1, None) == -1
    for code in range(0, -7, -1):
        assert len(N.strerror(code)) > 1
    assert "unknown" in N.strerror(-99)
    import torch
    if not torch.cuda.is_available():
        io.auto_reset = 0
        assert lib.rdv_step(C.byref(p), C.byref(st), C.byref(io), 4, 0, 0, None) == -5       # RDV_ERR_CUDA


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from reinforcement_learning_rendezvous_b200 import BatchedRendezvousEnv, MlpPolicy, RendezvousEnv
    with pytest.raises(RuntimeError):
        BatchedRendezvousEnv(4)
    with pytest.raises(RuntimeError):
        RendezvousEnv()
    with pytest.raises(RuntimeError):
        MlpPolicy.load(os.path.join(ROOT, "tests", "golden", "policy.npz"))


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "reinforcement_learning_rendezvous_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("the oracle", "").lower() or f == "verification.py", f
