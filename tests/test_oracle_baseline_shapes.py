"""CPU pins at the BASELINE.json shapes themselves (tests/golden/baseline_shapes.npz, produced by the unmodified
reference through tests/golden/make_golden.py baseline_shapes):

* the NumPy oracle against the reference, per pixel, for the headline workload (180 x 180, ss 3, 21 x 21 PSF; one
  parameter set per nmax bracket of epl.py:308-319), C3, C4 (1 EPL + 1,000 TNFW) and C5 (multi-plane x 4 bands);
* the product's own per-point arithmetic (the kernels' C++ compiled for the host, tests/hostmath) against the
  reference on the same full-size grids -- the CUDA build of the same code is checked in tests/test_gpu_baseline_shapes.py.
"""

import numpy as np
import pytest

from oracle import imsim, lens_model
from oracle.lenstronomy_port import PSF
from tests.golden_io import assert_close, load

CASES = load("baseline_shapes.npz")
TOL = dict(rtol=1e-11, atol_frac=1e-13)


def static_kwargs_data():
    from bench import static_config

    return static_config()["kwargs_data"]


def headline_rows():
    return CASES["headline_static"].inp["rows"]


def oracle_headline():
    from bench import LENS_LIST, SOURCE_LIST

    s = CASES["headline_static"]
    data = imsim.ImageData(image_data=s.inp["image_data"], **static_kwargs_data())
    return imsim.ImageModel(data, PSF(psf_type="PIXEL", kernel_point_source=s.inp["kernel"]), LENS_LIST, SOURCE_LIST,
                            [], dict(supersampling_factor=3))


def test_fixture_covers_every_series_bracket():
    """|e| of the ten headline sets: 2 x nmax 10, 3 x 20, 2 x 35, 2 x 50, 1 x 100 (epl.py:308-319); the data image is
    the reference's own simulation of bench.TRUTH plus the seeded noise bench.py adds."""
    rows = headline_rows()
    f = np.hypot(rows[:, 2], rows[:, 3])
    brackets = np.digitize(f, [0.12, 0.4, 0.6, 0.73, 0.86])
    assert sorted(brackets.tolist()) == [0, 0, 1, 1, 1, 2, 2, 3, 3, 4]
    s = CASES["headline_static"]
    rng = np.random.default_rng(0)
    t = s.out["truth"]
    noisy = t + rng.normal(0, 0.05, t.shape) + rng.normal(0, 1, t.shape) * np.sqrt(np.maximum(t, 0) / 100.0)
    np.testing.assert_array_equal(noisy, s.inp["image_data"])


@pytest.mark.parametrize("b", range(10))
def test_oracle_headline_per_pixel(b):
    from bench import row_to_kwargs

    c = CASES["headline_%02d" % b]
    om = oracle_headline()
    kl, ks, kll = row_to_kwargs(c.inp["row"])
    assert_close(om.image(kl, ks, kll, unconvolved=True), c.out["image_unconvolved"], what="unconvolved", **TOL)
    assert_close(om.linear_response_matrix(kl, ks, kll), c.out["A"], what="A", **TOL)
    logL, param = om.likelihood_data_given_model_linear(kl, ks, kll)
    assert_close(param, c.out["param"], what="param", rtol=1e-10, atol_frac=0)
    assert_close(logL, c.out["logL_linear"], what="logL", rtol=1e-11, atol_frac=0)
    assert_close(om.likelihood_data_given_model(kl, ks, kll), c.out["logL"], what="logL (amp from kwargs)", rtol=1e-11,
                 atol_frac=0)
    if "wls_model" in c.out:
        assert_close(om.image_linear_solve(kl, ks, kll)[0], c.out["wls_model"], what="wls", rtol=1e-10, atol_frac=1e-12)
        assert_close(om.image(kl, ks, kll), c.out["image"], what="image", **TOL)


@pytest.mark.parametrize("b", range(3))
def test_oracle_c3_per_pixel(b):
    s, c = CASES["c3_static"], CASES["c3_%02d" % b]
    data = imsim.ImageData(image_data=s.inp["image_data"], **static_kwargs_data())
    om = imsim.ImageModel(data, PSF(psf_type="PIXEL", kernel_point_source=s.inp["kernel"]), s.meta["lens_model_list"],
                          s.meta["source_model_list"], s.meta["lens_light_model_list"], s.meta["kwargs_numerics"])
    kl, ks, kll = c.meta["kwargs_lens"], c.meta["kwargs_source"], c.meta["kwargs_lens_light"]
    assert_close(om.linear_response_matrix(kl, ks, kll), c.out["A"], what="A", **TOL)
    assert_close(om.image(kl, ks, kll, unconvolved=True), c.out["image_unconvolved"], what="unconvolved", **TOL)
    logL, param = om.likelihood_data_given_model_linear(kl, ks, kll)
    assert_close(param, c.out["param"], what="param", rtol=1e-9, atol_frac=0)
    assert_close(logL, c.out["logL_linear"], what="logL", rtol=1e-10, atol_frac=0)


def c4_kwargs():
    """-> (names, kwargs list) of the 1 EPL + 1,000 TNFW model from the fixture's SoA arrays (single_plane_bulk.py:47-106)."""
    c = CASES["c4_substructure"]
    idx = c.inp["index_list"]
    uniq = c.meta["unique_lens_model_list"]
    names, kws = [], []
    for d, i in enumerate(idx):
        if int(i) >= len(uniq):
            continue  # NULL padding
        name = uniq[int(i)]
        names.append(name)
        from oracle.lens_profiles import PARAM_NAMES

        kws.append({p: float(c.inp["kw_" + p][d]) for p in PARAM_NAMES[name]})
    return names, kws


def test_oracle_c4_substructure():
    c = CASES["c4_substructure"]
    names, kws = c4_kwargs()
    assert names.count("TNFW") == 1000 and names[0] == "EPL"
    bx, by = lens_model.ray_shooting(names, c.inp["x"], c.inp["y"], kws)
    assert_close(bx, c.out["beta_x"], what="beta_x", rtol=1e-11, atol_frac=1e-12)
    assert_close(by, c.out["beta_y"], what="beta_y", rtol=1e-11, atol_frac=1e-12)


def oracle_c5_band(b):
    s = CASES["c5_static"]
    km = s.meta["kwargs_model"]
    data = imsim.ImageData(image_data=s.inp["image_data_%d" % b], **static_kwargs_data())
    return imsim.ImageModel(data, PSF(psf_type="PIXEL", kernel_point_source=s.inp["kernel_%d" % b]),
                            km["lens_model_list"], km["source_light_model_list"], km["lens_light_model_list"],
                            s.meta["band_numerics"][b], multi_plane=s.inp)


@pytest.mark.parametrize("n", range(3))
def test_oracle_c5_multi_plane_multi_band(n):
    c = CASES["c5_%02d" % n]
    kl, ks, kll = c.meta["kwargs_lens"], c.meta["kwargs_source"], c.meta["kwargs_lens_light"]
    total = 0.0
    for b in range(4):
        om = oracle_c5_band(b)
        logL, param = om.likelihood_data_given_model_linear(kl, ks, kll)
        assert_close(param, c.out["param_%d" % b], what="param band %d" % b, rtol=1e-9, atol_frac=0)
        assert_close(logL, c.out["logL_%d" % b], what="logL band %d" % b, rtol=1e-10, atol_frac=0)
        if "A_%d" % b in c.out:
            assert_close(om.linear_response_matrix(kl, ks, kll), c.out["A_%d" % b], what="A band %d" % b, **TOL)
        total += logL
    assert_close(total, c.out["logL"], what="sum over bands", rtol=1e-10, atol_frac=0)


# ---------------------------------------------------------------------------------------------------------------
# the product's per-point arithmetic (host build of the kernels' C++) on the full-size grids
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("b", range(10))
def test_hostmath_headline_unconvolved_per_pixel(b):
    """lens_eval + light_eval of the product on all 291,600 sub-pixels, re-binned in NumPy, against the reference's
    unconvolved image: per pixel, 1e-10 |b| + 1e-12 max|b|."""
    from bench import LENS_LIST, SOURCE_LIST, row_to_kwargs, static_config
    from tests import hostmath_util as hm

    c = CASES["headline_%02d" % b]
    kl, ks, _ = row_to_kwargs(c.inp["row"])
    om = oracle_headline()
    x, y = om.ImageNumerics.coordinates_evaluate
    bx, by = hm.ray_shoot(LENS_LIST, kl, x, y)
    # the specialised kernels' compile-time profile list (constant terms, series, linear terms) sums in another order
    sx, sy = hm.ray_shoot_epl_shear(kl, x, y)
    assert np.max(np.abs(sx - bx)) <= 1e-14 and np.max(np.abs(sy - by)) <= 1e-14
    flux = hm.surface_brightness(SOURCE_LIST, ks, sx, sy)[0]
    st = static_config()
    # the reference's flattened sub-pixel order -> re-binned [ny, nx] (oracle grid bookkeeping, pinned bit-exact)
    low, _ = om.ImageNumerics._numerics._grid.flux_array2image_low_high(flux)
    assert_close(low * st["delta_pix"] ** 2, c.out["image_unconvolved"], what="unconvolved", rtol=1e-10, atol_frac=1e-12)


def test_hostmath_c4_substructure():
    from tests import hostmath_util as hm

    c = CASES["c4_substructure"]
    names, kws = c4_kwargs()
    bx, by = hm.ray_shoot(names, kws, c.inp["x"], c.inp["y"])
    assert_close(bx, c.out["beta_x"], what="beta_x", rtol=1e-10, atol_frac=1e-11)
    assert_close(by, c.out["beta_y"], what="beta_y", rtol=1e-10, atol_frac=1e-11)
