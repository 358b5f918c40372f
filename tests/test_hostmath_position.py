"""CPU: PositionLikelihood -- the oracle restatement and k_source_position's arithmetic (compiled for the host) against
fixtures produced by the unmodified reference (tests/golden/position_likelihood.npz)."""

import numpy as np
import pytest

from oracle import position_likelihood as opl
from tests import hostmath_util as hm
from tests.golden_io import load

POS = load("position_likelihood.npz")
SRC_CASES = [n for n, c in POS.items() if "logL_source_position" in c.out]


def conditioning(c):
    """The penalty is delta^T (A^T A)^-1 delta / (2 sigma^2), i.e. ~ 1 / det(A)^2: an image close to a critical curve
    (quad_m1m3m4 has one at det A = 2.9e-4) amplifies the 1e-13-level agreement of the Hessian by ~1/|det A|.  The
    parity bar is therefore 1e-10 * max(1, 1e-2 / min |det A|), relative."""
    worst = 1.0
    for k, kw in enumerate(c.meta["kwargs_ps"]):
        if c.meta["point_source_type_list"][k] == "LENSED_POSITION":
            h = hm.lens_hessian(c.meta["lens_model_list"], c.meta["kwargs_lens"], np.asarray(kw["ra_image"], dtype=float),
                                np.asarray(kw["dec_image"], dtype=float))
            det = (1 - h[0]) * (1 - h[3]) - h[1] * h[2]
            worst = max(worst, 1e-2 / np.min(np.abs(det)))
    return worst


def _opts(c):
    o = c.meta["options"]
    return o.get("source_position_sigma", 0.001), o.get("source_position_tolerance")


@pytest.mark.parametrize("name", SRC_CASES)
def test_host_kernel_arithmetic_matches_the_reference(name):
    c = POS[name]
    sigma, hb = _opts(c)
    got = 0.0
    for k, kw in enumerate(c.meta["kwargs_ps"]):
        if c.meta["point_source_type_list"][k] == "LENSED_POSITION":
            got += hm.source_position_logl(c.meta["lens_model_list"], c.meta["kwargs_lens"], kw["ra_image"],
                                           kw["dec_image"], sigma, hb)
    want = float(c.out["logL_source_position"])
    assert abs(got - want) <= 1e-10 * conditioning(c) * abs(want), (name, got, want)


@pytest.mark.parametrize("name", SRC_CASES)
def test_oracle_matches_the_reference(name):
    """Closed-form Hessians (EPL, SHEAR, SIS): the 1e-10 bar; quad_m1m3m4 goes through the oracle's numerical
    derivative of the multipole deflection (4th order): 1e-7."""
    c = POS[name]
    sigma, hb = _opts(c)
    got = opl.source_position_likelihood(c.meta["lens_model_list"], c.meta["kwargs_lens"],
                                         c.meta["point_source_type_list"], c.meta["kwargs_ps"], sigma, hb)
    want = float(c.out["logL_source_position"])
    tol = 1e-7 if name == "quad_m1m3m4" else 1e-10
    assert abs(got - want) <= tol * conditioning(c) * abs(want), (name, got, want)


def test_oracle_image_position_and_astrometric_terms():
    c = POS["quad_image_positions"]
    o = c.meta["options"]
    ip = opl.image_position_likelihood(c.meta["kwargs_ps"], o["ra_image_list"], o["dec_image_list"],
                                       o["image_position_uncertainty"])
    want = float(c.out["logL"]) - float(c.out["logL_source_position"])
    assert abs(ip - want) <= 1e-6 * abs(want)   # the fixture difference cancels 8 digits of a 2.7e8 logL
    np.testing.assert_allclose(ip, -1.8125, rtol=1e-12)  # closed form of the four offsets
    a = POS["astrometric"]
    quad = POS["quad"].meta["kwargs_ps"]
    assert opl.astrometric_likelihood(quad, a.meta["kwargs_special"], a.meta["sigma"]) == pytest.approx(
        float(a.out["logL"]), rel=1e-13)
    assert float(POS["star_only"].out["logL"]) == 0.0


def test_image_on_the_critical_curve_takes_the_singular_branch():
    """det(Sigma_beta) == 0 exactly (SIS image at r = theta_E: A = diag(1, 0)): penalty 1e15 (position_likelihood.py:338-342),
    kernel arithmetic (host build) against the oracle restatement."""
    kw = [dict(theta_E=1.0, center_x=0.0, center_y=0.0)]
    ra, dec = [1.0, -0.9], [0.0, 0.0]
    got = hm.source_position_logl(["SIS"], kw, ra, dec, 0.001)
    want = opl.source_position_likelihood(["SIS"], kw, ["LENSED_POSITION"], [dict(ra_image=ra, dec_image=dec)], 0.001)
    assert want < -0.9e15 and abs(got - want) <= 1e-12 * abs(want)
    # hard bound: + 1e3 per image beyond the tolerance, none within
    far = hm.source_position_logl(["SIS"], kw, [1.1, -0.7], dec, 0.01, 0.05)
    near = hm.source_position_logl(["SIS"], kw, [1.1, -0.7], dec, 0.01, 0.5)
    assert abs((near - far) - 2e3) <= 1e-9 * abs(far)
