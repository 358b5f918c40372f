"""CPU: the product's per-point profile arithmetic (jaxtronomy_b200/csrc/lsim_profiles.cuh compiled for the host by
tests/hostmath) against the golden fixtures from the reference source.  This is the same C++ the CUDA kernels
inline; the GPU parity tests (tests/test_gpu_*.py) repeat the checks through the C ABI on the device."""

import numpy as np
import pytest

from jaxtronomy_b200 import _model
from tests import hostmath_util as hm
from tests.golden_io import assert_close, load

PROFILES = load("profiles.npz")
LENS_MODELS = load("lens_models.npz")

# north-star parity bar: rtol 1e-10 with the atol companion of SURVEY.md fact 10
BAR = dict(rtol=1e-10, atol_frac=1e-12)


def conditioning_rtol(prof, kw, x, y):
    """1e-10 everywhere, except where the REFERENCE's own formula is ill-conditioned: NFW.g (nfw.py:347-365)
    evaluates log(X/2) + acosh(1/X)/sqrt(1-X^2), two terms of size ln(2/X) that cancel to O(X^2 ln(2/X)); in
    fp64 its result carries a relative error of about eps/X^2, so for X = R/Rs < 1e-2 no two libm
    implementations agree to 1e-10 (the reference's own test_nfw.py uses decimal=5..8)."""
    rtol = np.full(np.shape(x), 1e-10)
    if prof == "NFW":
        R = np.hypot(x - kw["center_x"], y - kw["center_y"])
        X = np.maximum(np.maximum(R, 1e-8) / max(kw["Rs"], 1e-7), 1e-6)
        rtol = np.maximum(rtol, 20 * 2.2e-16 / X**2)
    return rtol


def test_log2_exp2_accuracy():
    rng = np.random.default_rng(0)
    x = np.concatenate([10.0 ** rng.uniform(-300, 300, 20000), rng.uniform(0.5, 2.0, 20000),
                        [1.0, 0.5, 2.0, np.sqrt(0.5), np.sqrt(2.0), 1e-18, 1e-8, 4.0]])
    got, ref = hm.log2(x), np.log2(x)
    # log2 feeds exponents (t log2 b - t/2 log2 R^2, log2(R^2)/(2n)): absolute accuracy relative to max(1, |.|)
    err = np.abs(got - ref) / np.maximum(np.abs(ref), 1.0)
    assert err.max() < 4e-16, err.max()
    near1 = np.abs(x - 1) < 0.3
    assert np.max(np.abs(got[near1] - ref[near1])) < 3e-16
    assert abs(hm.log2(np.array([1.0]))[0]) < 1e-17
    e = np.concatenate([rng.uniform(-1022, 1022, 40000), rng.uniform(-2, 2, 20000),
                        [0.0, 1.0, -1.0, 0.5, -1021.5, -1022.0, 1022.0, 1021.999, -1021.999]])
    got, ref = hm.exp2(e), np.exp2(e)
    assert np.max(np.abs(got - ref) / ref) < 4e-16
    assert hm.exp2(np.array([0.0]))[0] == 1.0
    # branch-free saturation through ONE integer clamp of the table index: finite arguments below -1022 read
    # 2^-1022 (1 + eps) with eps < 1 (instead of denormals / 0), above 1022 at most 2^1023 (instead of +inf); the
    # reduced argument r = y - rint(y) is taken from the unclamped argument, so +-inf (inf - inf) and nan give nan --
    # every caller follows with its own nan_to_num, as the reference does
    lo, hi = 2.0 ** -1022, 2.0 ** 1023
    sat = hm.exp2(np.array([-1022.5, -1100.0, -5000.0, -1e300]))
    assert np.all((sat >= lo) & (sat < 2 * lo)), sat
    sat = hm.exp2(np.array([1023.5, 1025.0, 1e300]))
    assert np.all((sat >= hi / 2) & (sat <= np.finfo(float).max)), sat
    assert np.all(np.isnan(hm.exp2(np.array([np.nan, -np.nan, np.inf, -np.inf]))))
    neg = np.concatenate([rng.uniform(-1022, 1021, 20000), [-1022.0, 0.0, 1021.0]])
    got, ref = hm.exp2_neg(neg), np.exp2(neg)
    assert np.max(np.abs(got - ref) / ref) < 4e-16 and np.isnan(hm.exp2_neg(np.array([np.nan]))[0])
    # the pre-scaled form used by the hot profiles: bit-identical to the unscaled one
    np.testing.assert_array_equal(hm.exp2_s(2048.0 * e), hm.exp2(e))


@pytest.mark.parametrize("name", [n for n, c in PROFILES.items() if "f_x" in c.out and "profile" in c.meta])
def test_lens_profile_deflection(name):
    c = PROFILES[name]
    prof = c.meta["profile"]
    kw = c.scalar_kwargs(_model.LENS_PARAM_NAMES[prof])
    fx, fy = hm.ray_shoot([prof], [kw], c.inp["x"], c.inp["y"], what=1, profile_kwargs_list=[c.meta["profile_kwargs"]])
    scale = max(np.max(np.abs(c.out["f_x"])), np.max(np.abs(c.out["f_y"])))
    # the two components share one scale (|alpha|): a component crossing zero has no relative accuracy on its own
    rtol = conditioning_rtol(prof, kw, c.inp["x"], c.inp["y"])
    for got, want, lab in ((fx, c.out["f_x"], "f_x"), (fy, c.out["f_y"], "f_y")):
        tol = rtol * np.abs(want) + 1e-12 * scale
        bad = np.abs(got - want) > tol
        assert not bad.any(), "%s %s: worst %.3g at %d (want %.17g got %.17g)" % (
            name, lab, np.max(np.abs(got - want)), int(np.argmax(np.abs(got - want))),
            want[np.argmax(np.abs(got - want))], got[np.argmax(np.abs(got - want))])


def test_series_truncation_is_adaptive_and_bounded():
    import ctypes as C
    from jaxtronomy_b200 import _lib
    names = _model.LENS_PARAM_NAMES["EPL"]
    for emod, nmax in ((0.05, 10), (0.25, 20), (0.39, 20), (0.5, 35), (0.7, 50), (0.8, 100), (0.9, 200), (0.95, 500)):
        row = np.array([1.0, 2.0, emod, 0.0, 0.0, 0.0])
        d = _model.lens_desc("EPL", [("col", k) for k in range(len(names))])
        full = hm.lib().hm_epl_terms(C.byref(d), row.ctypes.data_as(C.POINTER(C.c_double)), 0.0)
        cut = hm.lib().hm_epl_terms(C.byref(d), row.ctypes.data_as(C.POINTER(C.c_double)), 1e-14)
        assert full == nmax
        assert 2 <= cut <= nmax


@pytest.mark.parametrize("name", [n for n, c in PROFILES.items() if "flux" in c.out])
def test_light_profile(name):
    c = PROFILES[name]
    prof = c.meta["profile"]
    kw = c.scalar_kwargs(_model.LIGHT_PARAM_NAMES[prof])
    flux = hm.surface_brightness([prof], [kw], c.inp["x"], c.inp["y"])[0]
    assert_close(flux, c.out["flux"], what=name, **BAR)
    split = hm.surface_brightness([prof], [kw], c.inp["x"], c.inp["y"], split=True)[0]
    assert_close(split * kw["amp"], c.out["flux"], what=name + " split", **BAR)


@pytest.mark.parametrize("name", [n for n in LENS_MODELS if n.startswith("single_plane_kn")])
def test_single_plane(name):
    c = LENS_MODELS[name]
    bx, by = hm.ray_shoot(c.meta["lens_model_list"], c.meta["kwargs"], c.inp["x"], c.inp["y"])
    assert_close(bx, c.out["beta_x"], what=name, **BAR)
    assert_close(by, c.out["beta_y"], what=name, **BAR)


def test_bulk_and_multi_plane():
    c = LENS_MODELS["bulk_single_plane"]
    bx, by = hm.ray_shoot(c.meta["lens_model_list"], c.meta["kwargs"], c.inp["x"], c.inp["y"])
    assert_close(bx, c.out["beta_x"], what="bulk", **BAR)
    assert_close(by, c.out["beta_y"], what="bulk", **BAR)
    c = LENS_MODELS["multi_plane"]
    order = [int(i) for i in c.inp["sorted_index"]]
    names = [c.meta["lens_model_list"][i] for i in order]
    kws = [c.meta["kwargs"][i] for i in order]
    mp = dict(delta_T=c.inp["T_ij_list"], T_z=c.inp["T_z_list"], red2phys=c.inp["reduced2physical"],
              T_ij_end=float(c.inp["T_ij_stop"]), T_z_source=float(c.inp["T_z_source"]))
    bx, by = hm.ray_shoot(names, kws, c.inp["x"], c.inp["y"], multi_plane=mp)
    assert_close(bx, c.out["beta_x"], what="multi-plane", **BAR)
    assert_close(by, c.out["beta_y"], what="multi-plane", **BAR)


def test_round_epl_reads_no_unwritten_series_coefficient():
    """e1 = e2 = 0 exactly: the reference returns 2F1 = 1 without summing (epl.py:308-319, nmax "0"), so the prepared
    series is F_0 alone -- but epl_core takes term 1 unconditionally and must find F_1 = 0 there, not whatever the
    coefficient pool held before (the host harness fills its pools with NaN; on the device the pool is stale shared
    memory, which compute-sanitizer exposed as run-to-run differences for exactly this parameter set).  Against the
    closed form of the round power law: alpha = theta_E^t R^(1-t) (x, y)/R, t = gamma - 1."""
    rng = np.random.default_rng(3)
    x, y = rng.uniform(-2, 2, 301), rng.uniform(-2, 2, 301)
    kw = dict(theta_E=1.2, gamma=2.1, e1=0.0, e2=0.0, center_x=0.05, center_y=-0.02)
    fx, fy = hm.ray_shoot(["EPL"], [kw], x, y, what=1)
    assert np.all(np.isfinite(fx)) and np.all(np.isfinite(fy))
    xs, ys = x - kw["center_x"], y - kw["center_y"]
    R = np.hypot(xs, ys)
    t = kw["gamma"] - 1
    a = kw["theta_E"] ** t * R ** (1 - t)
    np.testing.assert_allclose(fx, a * xs / R, rtol=1e-12, atol=1e-14)
    np.testing.assert_allclose(fy, a * ys / R, rtol=1e-12, atol=1e-14)
    bx, by = hm.ray_shoot_epl_shear([kw, dict(gamma1=0.0, gamma2=0.0, ra_0=0.0, dec_0=0.0)], x, y)
    np.testing.assert_allclose(bx, x - a * xs / R, rtol=1e-12, atol=1e-13)
    np.testing.assert_allclose(by, y - a * ys / R, rtol=1e-12, atol=1e-13)
