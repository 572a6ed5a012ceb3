"""Shared helpers of the test-suite (golden fixtures, tolerances, comparison utilities)."""
import json
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
NO_RANGE = dict(rc0_range=0, vc0_range=0, qc0_range=0, wc0_range=0, qt0_range=0, wt0_range=0)

# BASELINE.json north_star: trajectories within 1e-9 relative fp64 error over full episodes.
REL_TOL = 1e-9


def golden(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


# Magnitude floors of the relative error, per column of a [.., 20] state (rc vc qc wc qt wt): O(1-10) m positions and
# unit quaternions get 1, velocities (~0.05 m/s) 1e-2, body rates (~5e-3 rad/s) 1e-3 -- so "1e-9 relative" means
# 1e-9 of the component's own scale, not of 1.
STATE_FLOORS = np.array([1.0] * 3 + [1e-2] * 3 + [1.0] * 4 + [1e-3] * 3 + [1.0] * 4 + [1e-3] * 3)
# get_errors(): position [m], velocity [m/s], attitude [rad], rate [rad/s]
ERROR_FLOORS = np.array([1.0, 1e-2, 1e-2, 1e-3])


def rel_err(a, b, floor=None):
    """max |a-b| / max(|b|, floor) -- relative to the magnitude of the quantity; used with REL_TOL.  ``floor``
    defaults to the per-column STATE_FLOORS for [.., 20] states and to 1 otherwise (rewards, O(1) quantities)."""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    if a.size == 0:
        return 0.0
    if floor is None:
        floor = STATE_FLOORS if (b.ndim >= 1 and b.shape[-1] == 20) else 1.0
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), floor)))


def cfg_kwargs(cfg_json):
    """constructor kwargs stored as JSON in the golden files -> (ctor kwargs, reward_kwargs)"""
    cfg = json.loads(str(cfg_json))
    reward = cfg.pop("reward_kwargs", None)
    kw = {k: (np.array(v, dtype=float) if isinstance(v, list) else v) for k, v in cfg.items()}
    return kw, reward


def near_threshold(values, thresholds, tol):
    """True where any value lies within tol of one of its thresholds (discrete flags may differ there)."""
    values, thresholds = np.asarray(values, dtype=float), np.asarray(thresholds, dtype=float)
    return np.abs(values - thresholds) <= tol
