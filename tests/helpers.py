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


def rel_err(a, b, floor=1.0):
    """max |a-b| / max(|b|, floor) -- relative to the magnitude of the quantity (floor 1: unit quaternions, O(1-10)
    positions); used with REL_TOL."""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    if a.size == 0:
        return 0.0
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
