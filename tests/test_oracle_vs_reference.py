"""Build-container only: the numpy oracle against the UNMODIFIED reference imported live from /root/reference
(through oracle/ref_stubs.py).  Skipped where the reference tree does not exist (the GPU box); the frozen
outputs of the same comparison are tests/golden/*.npz (tests/test_oracle_golden.py)."""
import numpy as np
import pytest

from helpers import NO_RANGE

from oracle import ref_stubs
from oracle import rdv_oracle as O

pytestmark = pytest.mark.skipif(not ref_stubs.reference_available(), reason="/root/reference is not present")


@pytest.fixture(scope="module")
def ref_env_mod():
    env_mod, _, _ = ref_stubs.import_reference()
    return env_mod


@pytest.mark.parametrize("cfg", [dict(), dict(dt=0.5, t_max=30), dict(koz_radius=10.0, h=400e3)])
def test_step_bit_exact_against_live_reference(ref_env_mod, cfg):
    rng = np.random.default_rng(3)
    ref = ref_env_mod.RendezvousEnv(quiet=True, **NO_RANGE, **cfg)
    orc = O.OracleEnv(**NO_RANGE, **cfg)
    ref.reset()
    orc.reset()
    state = np.hstack([ref.rc, ref.vc, ref.qc, ref.wc, ref.qt, ref.wt])
    state[0:3] += rng.uniform(-0.5, 0.5, 3)
    state[17:20] = rng.uniform(-0.03, 0.03, 3)
    for env in (ref, orc):
        env.rc, env.vc = state[0:3].copy(), state[3:6].copy()
        env.qc, env.wc = state[6:10].copy(), state[10:13].copy()
        env.qt, env.wt = state[13:17].copy(), state[17:20].copy()
    for k in range(40):
        a = rng.uniform(-1, 1, 6) * (0.3 if k % 3 else 1.0)
        o1, r1, d1, _ = ref.step(a)
        o2, r2, d2, _ = orc.step(a)
        assert np.array_equal(o1, o2) and r1 == r2 and bool(d1) == bool(d2)
        for name in ("rc", "vc", "qc", "wc", "qt", "wt"):
            assert np.array_equal(getattr(ref, name), getattr(orc, name)), (k, name)
        assert ref.t == orc.t and ref.bubble_radius == orc.bubble_radius
        assert bool(ref.collided) == bool(orc.collided) and ref.success == orc.success
        assert np.array_equal(ref.get_errors(), orc.get_errors()) and ref.dist_from_koz() == orc.dist_from_koz()
        if d1:
            break


def test_reset_draw_order_against_live_reference(ref_env_mod):
    """reset() consumes the global numpy stream in the same order and maps it to the same state."""
    ref = ref_env_mod.RendezvousEnv(quiet=True)
    orc = O.OracleEnv(rng=np.random)
    for seed in (0, 1, 2):
        np.random.seed(seed)
        o1 = ref.reset()
        np.random.seed(seed)
        o2 = orc.reset()
        assert np.array_equal(o1, o2)
        for name in ("rc", "vc", "qc", "wc", "qt", "wt"):
            assert np.array_equal(getattr(ref, name), getattr(orc, name)), name
        assert bool(ref.collided) == bool(orc.collided) and ref.success == orc.success
