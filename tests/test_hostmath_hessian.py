"""CPU: k_hessian's per-point arithmetic (lens_prepare_hess / lens_hessian_add of jaxtronomy_b200/csrc/lsim_grad.cuh,
compiled for the host by tests/hostmath) against the reference-generated fixtures and the oracle."""

import numpy as np
import pytest

from oracle import lens_model as olm
from tests import hostmath_util as hm
from tests.hessian_cases import HESSIANS, KEYS, PROFILE_CASES, check_profile, keep_mask, profile_case


@pytest.mark.parametrize("name", PROFILE_CASES)
def test_profile_hessian_matches_the_reference_fixture(name):
    c, prof, kw, pk = profile_case(name)
    got = hm.lens_hessian([prof], [kw], c.inp["x"], c.inp["y"], profile_kwargs_list=[pk])
    check_profile(name, got, c, prof, kw)


@pytest.mark.parametrize("name", PROFILE_CASES)
def test_profile_hessian_is_the_derivative_of_the_oracle_deflection(name):
    """Independent of the fixtures: 4th-order central difference of the oracle's alpha, 1e-8 of the scale.
    This is the check that carries NIE / SIE, whose reference Hessian is itself a noisy forward difference."""
    c, prof, kw, pk = profile_case(name)
    x, y = c.inp["x"], c.inp["y"]
    keep = keep_mask(x, y, [kw]) & (np.hypot(x - kw.get("center_x", kw.get("ra_0")), y - kw.get("center_y", kw.get("dec_0"))) > 0.3)
    x, y = x[keep], y[keep]
    got = hm.lens_hessian([prof], [kw], x, y, profile_kwargs_list=[pk])
    want = olm.hessian_differential([prof], x, y, [kw], profile_kwargs_list=[pk], h=1e-3)
    scale = max(np.max(np.abs(w)) for w in want)
    # sie_2 / nie_2: |e| clamps to 0.9999 (q = 5e-5), the deflection varies on the scale q R and the finite
    # difference's own truncation error grows accordingly
    tol = 1e-6 if name in ("sie_2", "nie_2") else 1e-8
    for g, w, k in zip(got, want, KEYS):
        assert np.max(np.abs(g - w)) <= tol * scale, "%s %s: %.3g of scale %.3g" % (name, k, np.max(np.abs(g - w)), scale)


def test_model_hessian_sums_the_profiles():
    c = HESSIANS["model_knull"]
    names, kw = c.meta["lens_model_list"], c.meta["kwargs"]
    x, y = c.inp["x"], c.inp["y"]
    keep = keep_mask(x, y, kw)
    got = hm.lens_hessian(names, kw, x, y)
    scale = max(np.max(np.abs(c.out[k][keep])) for k in KEYS)
    for g, k in zip(got, KEYS):
        assert np.max(np.abs(g - c.out[k])[keep]) <= 1e-10 * scale, k
