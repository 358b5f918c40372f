"""Loader for the committed golden fixtures (tests/golden/*.npz, written by tests/golden/make_golden.py)."""

import json
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


class Case:
    def __init__(self, name):
        self.name = name
        self.inp = {}
        self.out = {}
        self.meta = {}

    def scalar_kwargs(self, names):
        return {k: float(self.inp[k]) if k != "m" else int(self.inp[k]) for k in names}


def load(fname):
    """-> dict case_name -> Case"""
    cases = {}
    with np.load(os.path.join(GOLDEN_DIR, fname)) as z:
        for key in z.files:
            parts = key.split("/")
            c = cases.setdefault(parts[0], Case(parts[0]))
            if parts[1] == "meta":
                c.meta = json.loads(str(z[key]))
            elif parts[1] == "in":
                c.inp[parts[2]] = z[key]
            else:
                c.out[parts[2]] = z[key]
    return cases


def assert_close(actual, desired, rtol=1e-10, atol_frac=1e-12, what=""):
    """Parity protocol: |a-b| <= rtol*|b| + atol_frac*max|b|."""
    actual = np.asarray(actual, dtype=float)
    desired = np.asarray(desired, dtype=float)
    assert actual.shape == desired.shape, "%s: shape %s vs %s" % (what, actual.shape, desired.shape)
    scale = np.max(np.abs(desired)) if desired.size else 0.0
    tol = rtol * np.abs(desired) + atol_frac * scale
    bad = ~(np.abs(actual - desired) <= tol)
    if np.any(bad):
        i = np.argmax(np.where(bad, np.abs(actual - desired) - tol, -np.inf))
        raise AssertionError(
            "%s: %d/%d outside tolerance; worst at flat index %d: actual %.17g desired %.17g (|diff| %.3g, tol %.3g)"
            % (what, int(bad.sum()), bad.size, i, actual.flat[i], desired.flat[i],
               abs(actual.flat[i] - desired.flat[i]), tol.flat[i]))
