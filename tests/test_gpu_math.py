"""GPU: accuracy of the device math helpers against numpy (fp64), element-wise through rdv_math_probe."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _probe(x, op):
    import torch
    from reinforcement_learning_rendezvous_b200 import _native as N
    xd = torch.as_tensor(np.ascontiguousarray(x, dtype=np.float64), device="cuda")
    yd = torch.empty_like(xd)
    N.check(N.lib().rdv_math_probe(xd.data_ptr(), yd.data_ptr(), xd.numel(), op,
                                   C.c_void_p(torch.cuda.current_stream().cuda_stream)), "rdv_math_probe")
    return yd.cpu().numpy()


def _ulps(got, want):
    return np.abs(got - want) / np.spacing(np.abs(want))


def test_rsqrt_rcp_sqrt_are_ulp_accurate():
    rng = np.random.default_rng(0)
    x = np.concatenate([10.0 ** rng.uniform(-12, 12, 200000), 1.0 + rng.uniform(-0.05, 0.05, 200000),
                        [1.0, 4.0, 0.25, 2.0, 1e-300, 1e300]])
    assert _ulps(_probe(x, 0), 1.0 / np.sqrt(x)).max() <= 2.0
    assert _ulps(_probe(x, 1), 1.0 / x).max() <= 1.5
    assert _ulps(_probe(x, 2), np.sqrt(x)).max() <= 2.5
    assert _probe(np.array([0.0]), 2)[0] == 0.0


def test_pow_neg_tenth():
    rng = np.random.default_rng(1)
    x = 10.0 ** rng.uniform(-29, 29, 200000)
    assert (np.abs(_probe(x, 3) / x ** -0.1 - 1)).max() < 5e-15          # fp64 controller path
    xs = 10.0 ** rng.uniform(-12, 8, 200000)
    assert (np.abs(_probe(xs, 4) / xs ** -0.1 - 1)).max() < 2e-6          # float32 controller path


def test_rounded_acos_matches_numpy_rounding():
    """acos(round(c, 5)) with numpy's round == rint(c * 1e5) / 1e5, including the exact +-1 end points."""
    rng = np.random.default_rng(2)
    c = np.concatenate([rng.uniform(-1, 1, 300000), [1.0, -1.0, 0.999995, 0.9999949999, -0.999995, 0.0, 0.5],
                        (np.arange(-100000, 100001, 7) + 0.5) / 1e5 * (1 - 1e-12)])
    c = np.clip(c, -1, 1)
    want = np.arccos(np.round(c, 5))
    got = _probe(c, 5)
    assert np.isfinite(got).all()
    assert np.abs(got - want).max() < 4e-15 * np.pi + 1e-7 * 0      # same rounded argument -> same angle
    assert got[300000] == 0.0 and abs(got[300001] - np.pi) < 1e-15
    # the angle is a 200,001-entry table look-up (filled by the same acos on the device): every entry, and the
    # arguments that have no entry (a cosine beyond +-1 by more than the rounding, NaN) take the library path
    k = np.arange(-100000, 100001, dtype=np.float64)
    got = _probe(k / 1e5, 5)
    assert np.abs(got - np.arccos(np.round(k / 1e5, 5))).max() < 4e-15 * np.pi
    odd = _probe(np.array([1.5, -1.00002, np.nan, 1.0 + 1e-9, -1.0 - 1e-9]), 5)
    assert np.isnan(odd[:3]).all() and odd[3] == 0.0 and abs(odd[4] - np.pi) < 1e-15
