"""Shared by the CPU (host-compiled kernel arithmetic) and GPU Hessian tests: fixtures, masks and tolerances.

Fixtures: tests/golden/hessians.npz, written by tests/golden/make_golden.py from the UNMODIFIED reference
(LensModel.hessian, lens_model.py:384-414).

Two facts about the reference shape the comparison:

* NIE / SIE: the reference's Hessian is a FORWARD difference of its own deflection with step 1e-10
  (Profiles/nie.py:298-312), whose rounding noise is ~eps |alpha| / 1e-10 ~ 1e-6 |alpha| (observed: up to 5e-3 of the
  Hessian's scale at e = 0.99).  The product returns the exact derivative; the tolerance for these two profiles is
  therefore the reference's own noise (NIE_ATOL of the fixture's scale), and the exact derivative is additionally
  checked against the oracle's 4th-order central difference at 1e-8.
* Core clamps: inside a profile's clamp radius (R < 1e-6 .. 1e-4 around its centre, e.g. sis.py:35-43) the reference's
  closed-form Hessian is no longer the derivative of its clamped deflection.  The product differentiates the
  deflection it computes; points closer than CLAMP_RADIUS to a profile centre are outside the stated parity scope.
"""

import numpy as np

from jaxtronomy_b200 import _model
from tests.golden_io import load

HESSIANS = load("hessians.npz")
PROFILE_CASES = [n for n, c in HESSIANS.items() if "profile" in c.meta]
MODEL_CASES = [n for n in HESSIANS if n.startswith("model_k")]
FLEXION_CASES = [n for n in HESSIANS if n.startswith("model_flexion")]
KEYS = ("f_xx", "f_xy", "f_yx", "f_yy")
CLAMP_RADIUS = 1e-3
NIE_ATOL = 2e-2  # of the fixture's scale; reference forward-difference noise, see above (worst case e = 0.99)


def centre(kw):
    return (kw.get("center_x", kw.get("ra_0", 0.0)), kw.get("center_y", kw.get("dec_0", 0.0)))


def keep_mask(x, y, kwargs_list):
    keep = np.ones(np.shape(x), dtype=bool)
    for kw in kwargs_list:
        cx, cy = centre(kw)
        keep &= np.hypot(x - cx, y - cy) > CLAMP_RADIUS
    return keep


def profile_case(name):
    c = HESSIANS[name]
    prof = c.meta["profile"]
    kw = c.scalar_kwargs(_model.LENS_PARAM_NAMES[prof])
    return c, prof, kw, c.meta["profile_kwargs"]


def check_profile(name, got, c, prof, kw):
    """1e-10 of the Hessian's scale (its four entries share one scale: single entries cross zero)."""
    keep = keep_mask(c.inp["x"], c.inp["y"], [kw])
    scale = max(np.max(np.abs(c.out[k][keep])) for k in KEYS)
    fd_noise = prof in ("NIE", "SIE") and not c.meta["profile_kwargs"].get("NIE") is False
    for g, k in zip(got, KEYS):
        want = c.out[k]
        tol = (NIE_ATOL * scale) if fd_noise else (1e-10 * np.abs(want) + 1e-10 * scale)
        bad = keep & ~(np.abs(np.asarray(g) - want) <= tol)
        assert not bad.any(), "%s %s: %d outside tolerance, worst |diff| %.3g (scale %.3g)" % (
            name, k, int(bad.sum()), np.max(np.abs(np.asarray(g) - want)[keep]), scale)
