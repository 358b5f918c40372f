"""CPU: the NumPy oracle against (a) the golden fixtures produced by the unmodified reference source and
(b) scipy.special.hyp2f1, the reference's only oracle-independent numeric pin (test/test_Util/test_hyp2f1_util.py)."""

import numpy as np
import pytest
from scipy.special import hyp2f1

from oracle import imsim, lens_model, lens_profiles, light_profiles, util
from oracle.lenstronomy_port import PSF, PixelGrid
from tests.golden_io import assert_close, load

PROFILES = load("profiles.npz")
LENS_MODELS = load("lens_models.npz")
IMSIM = load("imsim.npz")

# the oracle restates the same NumPy arithmetic as the reference run through refshim: demand ~ulp agreement
TIGHT = dict(rtol=1e-13, atol_frac=1e-15)


@pytest.mark.parametrize("name", [n for n, c in PROFILES.items() if "f_x" in c.out and "profile" in c.meta])
def test_lens_profile_deflection(name):
    c = PROFILES[name]
    prof = c.meta["profile"]
    kw = c.scalar_kwargs(lens_profiles.PARAM_NAMES[prof])
    fx, fy = lens_profiles.derivatives(prof, c.inp["x"], c.inp["y"], c.meta.get("profile_kwargs"), **kw)
    assert_close(fx, c.out["f_x"], what=name + " f_x", **TIGHT)
    assert_close(fy, c.out["f_y"], what=name + " f_y", **TIGHT)


def test_epl_major_axis_sweep():
    c = PROFILES["epl_major_sweep"]
    for i, q in enumerate(c.inp["q"]):
        fx, fy = lens_profiles.epl_major_axis(c.inp["x"], c.inp["y"], float(c.inp["b"]), float(c.inp["t"]), float(q))
        assert_close(fx, c.out["f_x"][i], what="q=%g" % q, **TIGHT)
        assert_close(fy, c.out["f_y"][i], what="q=%g" % q, **TIGHT)


@pytest.mark.parametrize("name", [n for n, c in PROFILES.items() if "flux" in c.out])
def test_light_profile(name):
    c = PROFILES[name]
    prof = c.meta["profile"]
    kw = c.scalar_kwargs(light_profiles.PARAM_NAMES[prof])
    flux = light_profiles.surface_brightness([prof], c.inp["x"], c.inp["y"], [kw])
    assert_close(flux, c.out["flux"], what=name, **TIGHT)


@pytest.mark.parametrize("name", [n for n in PROFILES if n.startswith("hyp2f1")])
def test_hyp2f1_series_vs_reference_and_scipy(name):
    c = PROFILES[name]
    nmax = int(c.inp["nmax"])
    for i, t in enumerate(c.inp["t"]):
        h = util.hyp2f1_lopez_temme_8(1, t / 2, 2 - t / 2, c.inp["z"], nmax=nmax)
        assert np.max(np.abs(h - c.out["h"][i]) / np.abs(c.out["h"][i])) < 1e-14
        ref = hyp2f1(1, t / 2, 2 - t / 2, c.inp["z"])
        # the nmax schedule is the reference's accuracy spec: truncated series vs exact 2F1
        # (at f=0.99 the reference's 500-term series is itself only good to ~1e-5: matching the
        # truncated sum, not the exact function, is the parity target there)
        assert np.max(np.abs(h - ref) / np.abs(ref)) < (2e-10 if nmax < 500 else 1e-4), name


def test_epl_vs_scipy_over_reference_sweep():
    """test_epl.py:199-206 sweep, against the closed form with scipy's exact 2F1 (oracle-independent pin)."""
    y = np.linspace(-10, 10, 101)
    x = np.ones_like(y)
    b, t = 1.7, 0.9
    worst = 0.0
    for q in np.linspace(0.03, 1, 200)[::-1]:
        fx, fy = lens_profiles.epl_major_axis(x, y, b, t, q)
        Z = q * x + 1j * y
        f = (1 - q) / (1 + q)
        ref = 2 / (1 + q) * (b / np.abs(Z)) ** t * Z * hyp2f1(1, t / 2, 2 - t / 2, -f * Z / np.conj(Z))
        worst = max(worst, np.max(np.abs(fx + 1j * fy - ref) / np.abs(ref)))
    assert worst < 1e-10, worst


def test_nmax_schedule():
    assert [util.epl_nmax(f) for f in (0.0, 0.05, 0.12, 0.39, 0.4, 0.6, 0.73, 0.86, 0.92, 0.99)] == [
        0, 10, 20, 20, 35, 50, 100, 200, 500, 500]


@pytest.mark.parametrize("name", [n for n in LENS_MODELS if n.startswith("single_plane")])
def test_single_plane_ray_shooting(name):
    c = LENS_MODELS[name]
    k = c.meta["k"]
    if isinstance(k, list) and not isinstance(k[0], bool):
        k = tuple(k)
    bx, by = lens_model.ray_shooting(c.meta["lens_model_list"], c.inp["x"], c.inp["y"], c.meta["kwargs"], k=k)
    assert_close(bx, c.out["beta_x"], what=name, **TIGHT)
    assert_close(by, c.out["beta_y"], what=name, **TIGHT)


def test_bulk_single_plane():
    c = LENS_MODELS["bulk_single_plane"]
    all_kwargs = {k[3:]: v for k, v in c.inp.items() if k.startswith("kw_")}
    bx, by = lens_model.ray_shooting_bulk(c.meta["unique_lens_model_list"], c.inp["x"], c.inp["y"], all_kwargs,
                                          c.inp["index_list"])
    assert_close(bx, c.out["beta_x"], what="bulk x", rtol=1e-12, atol_frac=1e-14)
    assert_close(by, c.out["beta_y"], what="bulk y", rtol=1e-12, atol_frac=1e-14)
    # same answer through the plain profile-list path
    bx2, by2 = lens_model.ray_shooting(c.meta["lens_model_list"], c.inp["x"], c.inp["y"], c.meta["kwargs"])
    assert_close(bx2, c.out["beta_x"], what="bulk vs list", rtol=1e-12, atol_frac=1e-13)


def test_multi_plane():
    c = LENS_MODELS["multi_plane"]
    bx, by = lens_model.ray_shooting_multi_plane(
        c.meta["lens_model_list"], c.inp["x"], c.inp["y"], c.meta["kwargs"], [int(i) for i in c.inp["sorted_index"]],
        c.inp["T_ij_list"], c.inp["T_z_list"], c.inp["reduced2physical"], float(c.inp["T_ij_stop"]),
        float(c.inp["T_z_source"]))
    assert_close(bx, c.out["beta_x"], what="mp x", **TIGHT)
    assert_close(by, c.out["beta_y"], what="mp y", **TIGHT)


@pytest.mark.parametrize("ss", [1, 2, 3])
def test_grid_bookkeeping_bit_exact(ss):
    c = IMSIM["grid_ss%d" % ss]
    g = imsim.RegularGrid(6, 6, c.inp["transform_pix2angle"], float(c.inp["ra_at_xy_0"]), float(c.inp["dec_at_xy_0"]), ss)
    ra, dec = g.coordinates_evaluate
    np.testing.assert_array_equal(g._compute_indexes, c.out["compute_indexes"])
    np.testing.assert_array_equal(ra, c.out["ra"])
    np.testing.assert_array_equal(dec, c.out["dec"])
    low, high = g.flux_array2image_low_high(c.inp["flux"])
    np.testing.assert_array_equal(low, c.out["low"])
    if ss > 1:
        np.testing.assert_array_equal(high, c.out["high"])


def _psf(meta_psf, inp):
    kw = dict(meta_psf)
    if kw["psf_type"] == "PIXEL":
        kw["kernel_point_source"] = inp["kernel"]
    return PSF(**kw)


@pytest.mark.parametrize("name", [n for n in IMSIM if n.startswith("numerics_")])
def test_numerics_re_size_convolve(name):
    c = IMSIM[name]
    pg = PixelGrid(c.meta["nx"], c.meta["ny"], c.inp["transform_pix2angle"], float(c.inp["ra_at_xy_0"]),
                   float(c.inp["dec_at_xy_0"]))
    psf = _psf(c.meta["kwargs_psf"], c.inp)
    psf.set_pixel_size(pg.pixel_width)
    if c.meta.get("subframe"):
        num = imsim.NumericsSubFrame(pg, psf, flux_evaluate_indexes=c.inp["flux_evaluate_indexes"],
                                     **c.meta["kwargs_numerics"])
    else:
        num = imsim.Numerics(pg, psf, **c.meta["kwargs_numerics"])
    ra, dec = num.coordinates_evaluate
    np.testing.assert_array_equal(ra, c.out["ra"])
    np.testing.assert_array_equal(dec, c.out["dec"])
    assert_close(num.re_size_convolve(c.inp["flux"]), c.out["image"], what=name, **TIGHT)
    assert_close(num.re_size_convolve(c.inp["flux"], unconvolved=True), c.out["image_unconvolved"], what=name, **TIGHT)


def build_oracle_image_model(c):
    kd = dict(image_data=c.inp["image_data"], transform_pix2angle=c.inp["transform_pix2angle"],
              ra_at_xy_0=float(c.inp["ra_at_xy_0"]), dec_at_xy_0=float(c.inp["dec_at_xy_0"]),
              background_rms=float(c.inp["background_rms"]),
              exposure_time=c.inp["exposure_time"] if c.inp["exposure_time"].ndim else float(c.inp["exposure_time"]))
    if "noise_map" in c.inp:
        kd["noise_map"] = c.inp["noise_map"]
    data = imsim.ImageData(**kd)
    psf = _psf(c.meta["kwargs_psf"], c.inp)
    kwargs_numerics = dict(c.meta["kwargs_numerics"])
    if "flux_evaluate_indexes" in c.inp:
        kwargs_numerics["flux_evaluate_indexes"] = c.inp["flux_evaluate_indexes"]
    return imsim.ImageModel(data, psf, c.meta["lens_model_list"], c.meta["source_model_list"],
                            c.meta["lens_light_model_list"], kwargs_numerics,
                            likelihood_mask=c.inp.get("likelihood_mask"),
                            lens_profile_kwargs_list=c.meta.get("lens_profile_kwargs"))


IMAGE_CASES = [n for n, c in IMSIM.items() if "logL_linear" in c.out]


@pytest.mark.parametrize("name", IMAGE_CASES)
def test_image_model_and_linear_fit(name):
    c = IMSIM[name]
    im = build_oracle_image_model(c)
    kl, ks, kll = c.meta["kwargs_lens"], c.meta["kwargs_source"], c.meta["kwargs_lens_light"]
    tol = dict(rtol=1e-11, atol_frac=1e-13)
    assert_close(im.image(kl, ks, kll), c.out["image"], what=name + " image", **tol)
    assert_close(im.image(kl, ks, kll, unconvolved=True), c.out["image_unconvolved"], what=name + " unconv", **tol)
    assert_close(im.source_surface_brightness(ks, kl), c.out["source"], what=name + " source", **tol)
    if kll:
        assert_close(im.lens_surface_brightness(kll), c.out["lens_light"], what=name + " lens light", **tol)
    assert_close(im.likelihood_data_given_model(kl, ks, kll), c.out["logL"], what=name + " logL", rtol=1e-11, atol_frac=0)
    assert_close(im.linear_response_matrix(kl, ks, kll), c.out["A"], what=name + " A", **tol)
    model, _, _, param = im.image_linear_solve(kl, ks, kll)
    assert_close(param, c.out["param"], what=name + " param", rtol=1e-9, atol_frac=1e-12)
    assert_close(model, c.out["wls_model"], what=name + " wls model", rtol=1e-9, atol_frac=1e-11)
    logL, _ = im.likelihood_data_given_model_linear(kl, ks, kll)
    assert_close(logL, c.out["logL_linear"], what=name + " logL linear", rtol=1e-10, atol_frac=0)
    logLm, _ = im.likelihood_data_given_model_linear(kl, ks, kll, source_marg=True)
    assert_close(logLm, c.out["logL_marg"], what=name + " logL marg", rtol=1e-10, atol_frac=0)


def test_degenerate_system_returns_zero_amplitudes():
    c = IMSIM["linear_degenerate"]
    assert np.all(c.out["param"] == 0)
    im = build_oracle_image_model(c)
    _, _, _, param = im.image_linear_solve(c.meta["kwargs_lens"], c.meta["kwargs_source"], c.meta["kwargs_lens_light"])
    assert np.all(param == 0)


# ---------------------------------------------------------------------------------------------------------------
# point sources: oracle/point_source.py + the point-source branches of oracle/imsim.py
# ---------------------------------------------------------------------------------------------------------------
POINTS = load("point_sources.npz")


@pytest.mark.parametrize("name", [n for n in POINTS if n.startswith("rendering_")])
def test_point_source_rendering(name):
    from oracle.lenstronomy_port import PixelGrid
    from oracle.point_source import point_source_rendering

    c = POINTS[name]
    n, ss = c.meta["numPix"], c.meta["supersampling"]
    pg = PixelGrid(n, n, c.inp["transform_pix2angle"], float(c.inp["ra_at_xy_0"]), float(c.inp["dec_at_xy_0"]))
    kernel = c.inp["kernel"] / np.sum(c.inp["kernel"])  # PSF(kernel_point_source_normalisation=True)
    tol = dict(rtol=1e-11, atol_frac=1e-13)
    got = point_source_rendering(pg, kernel, ss, c.inp["ra"], c.inp["dec"], c.inp["amp"])
    assert_close(got, c.out["image"], what=name, **tol)
    gotu = point_source_rendering(pg, kernel, ss, c.inp["ra"], c.inp["dec"], c.inp["amp"], unconvolved=True)
    assert_close(gotu, c.out["image_unconvolved"], what=name + " unconvolved", **tol)
    # flux conservation for stamps fully inside the frame: the first image carries amp[0] in total
    one = point_source_rendering(pg, kernel, ss, c.inp["ra"][:1], c.inp["dec"][:1], c.inp["amp"][:1])
    assert abs(one.sum() - c.inp["amp"][0]) < 1e-9 * c.inp["amp"][0]


def _oracle_ps_model(c):
    kd = dict(image_data=c.inp["image_data"], transform_pix2angle=c.inp["transform_pix2angle"],
              ra_at_xy_0=float(c.inp["ra_at_xy_0"]), dec_at_xy_0=float(c.inp["dec_at_xy_0"]),
              background_rms=float(c.inp["background_rms"]), exposure_time=float(c.inp["exposure_time"]))
    return imsim.ImageModel(imsim.ImageData(**kd), _psf(c.meta["kwargs_psf"], c.inp), c.meta["lens_model_list"],
                            c.meta["source_model_list"], c.meta["lens_light_model_list"], c.meta["kwargs_numerics"],
                            point_source_type_list=c.meta["point_source_type_list"])


def _ps_kwargs(c):
    return [{k: np.asarray(v, dtype=float) for k, v in kw.items()} for kw in c.meta["kwargs_ps"]]


def test_image_model_with_point_sources():
    c = POINTS["image_quad_and_star"]
    im = _oracle_ps_model(c)
    kl, ks, kll, kps = c.meta["kwargs_lens"], c.meta["kwargs_source"], c.meta["kwargs_lens_light"], _ps_kwargs(c)
    tol = dict(rtol=1e-11, atol_frac=1e-13)
    assert_close(im.image(kl, ks, kll, kwargs_ps=kps), c.out["image"], what="image", **tol)
    assert_close(im.image(kl, ks, kll, kwargs_ps=kps, unconvolved=True), c.out["image_unconvolved"], what="unconv", **tol)
    assert_close(im.image(kl, ks, kll, kwargs_ps=kps, point_source_add=False), c.out["image_no_ps"], what="no ps", **tol)
    assert_close(im.point_source(kps), c.out["point_source"], what="point_source", **tol)
    assert_close(im.point_source(kps, k=1), c.out["point_source_k1"], what="point_source k=1", **tol)
    logL = im.likelihood_data_given_model(kl, ks, kll, kwargs_ps=kps)
    assert abs(logL - float(c.out["logL"])) <= 1e-11 * abs(float(c.out["logL"]))
    A = im.linear_response_matrix(kl, ks, kll, kwargs_ps=kps)
    assert A.shape == c.out["A"].shape == (7, 1600)  # source, lens light, 4 + 1 point-source images
    assert_close(A, c.out["A"], what="A", **tol)
    model, _, _, param = im.image_linear_solve(kl, ks, kll, kwargs_ps=kps)
    assert_close(param, c.out["param"], what="param", rtol=1e-9, atol_frac=1e-11)
    assert_close(model, c.out["wls_model"], what="wls model", rtol=1e-9, atol_frac=1e-11)
    logL_lin, _ = im.likelihood_data_given_model_linear(kl, ks, kll, kwargs_ps=kps)
    assert abs(logL_lin - float(c.out["logL_linear"])) <= 1e-9 * abs(float(c.out["logL_linear"]))


@pytest.mark.parametrize("name", ["error_map_all", "error_map_first_only"])
def test_psf_error_map(name):
    """ImageModel._error_map_psf + its use in the likelihood and the linear solve (C_D + model error)."""
    c = POINTS[name]
    kd = dict(image_data=c.inp["image_data"], transform_pix2angle=c.inp["transform_pix2angle"],
              ra_at_xy_0=float(c.inp["ra_at_xy_0"]), dec_at_xy_0=float(c.inp["dec_at_xy_0"]),
              background_rms=float(c.inp["background_rms"]), exposure_time=float(c.inp["exposure_time"]))
    im = imsim.ImageModel(imsim.ImageData(**kd), _psf(c.meta["kwargs_psf"], c.inp), c.meta["lens_model_list"],
                          c.meta["source_model_list"], c.meta["lens_light_model_list"], c.meta["kwargs_numerics"],
                          point_source_type_list=c.meta["point_source_type_list"],
                          psf_variance_map=c.inp["psf_variance_map"],
                          psf_error_map_bool_list=c.meta["psf_error_map_bool_list"])
    kl, ks, kll, kps = c.meta["kwargs_lens"], c.meta["kwargs_source"], c.meta["kwargs_lens_light"], _ps_kwargs(c)
    tol = dict(rtol=1e-11, atol_frac=1e-13)
    assert c.out["error_map"].max() > 0
    assert_close(im.error_map_psf(kps), c.out["error_map"], what="error map", **tol)
    logL = im.likelihood_data_given_model(kl, ks, kll, kwargs_ps=kps)
    assert abs(logL - float(c.out["logL"])) <= 1e-11 * abs(float(c.out["logL"]))
    model, model_error, _, param = im.image_linear_solve(kl, ks, kll, kwargs_ps=kps)
    assert_close(model_error, c.out["model_error"], what="model error", **tol)
    assert_close(param, c.out["param"], what="param", rtol=1e-9, atol_frac=1e-11)
    logL_lin, _ = im.likelihood_data_given_model_linear(kl, ks, kll, kwargs_ps=kps)
    assert abs(logL_lin - float(c.out["logL_linear"])) <= 1e-9 * abs(float(c.out["logL_linear"]))


# ---------------------------------------------------------------------------------------------------------------
# Hessians: the restated closed forms (EPL family, SIS, SHEAR, CONVERGENCE) against the reference's fixtures,
# INCLUDING the points inside the core clamps (the closed forms reproduce the reference there)
# ---------------------------------------------------------------------------------------------------------------
def _hessian_cases():
    from tests.hessian_cases import HESSIANS

    return [n for n, c in HESSIANS.items()
            if c.meta.get("profile") in ("EPL", "SIS", "SHEAR", "CONVERGENCE")
            or (c.meta.get("profile") == "SIE" and c.meta["profile_kwargs"].get("NIE") is False)]


@pytest.mark.parametrize("name", _hessian_cases())
def test_closed_form_hessians(name):
    from oracle import lens_profiles as olp
    from tests.hessian_cases import KEYS, profile_case

    c, prof, kw, pk = profile_case(name)
    got = olp.hessian(prof, c.inp["x"], c.inp["y"], pk, **kw)
    scale = max(np.max(np.abs(c.out[k][np.isfinite(c.out[k])])) for k in KEYS)
    for g, k in zip(got, KEYS):
        assert np.all(np.abs(g - c.out[k]) <= 1e-10 * np.abs(c.out[k]) + 1e-12 * scale), (name, k)


def test_model_hessian_sum_and_selection():
    from oracle import lens_model as olm
    from tests.hessian_cases import HESSIANS, KEYS, MODEL_CASES, keep_mask

    for name in MODEL_CASES:
        c = HESSIANS[name]
        k = c.meta["k"]
        if isinstance(k, list) and not isinstance(k[0], bool):
            k = tuple(k)
        got = olm.hessian(c.meta["lens_model_list"], c.inp["x"], c.inp["y"], c.meta["kwargs"], k=k)
        # NFW enters through the numerical derivative: 1e-8 of the scale, away from the centres
        keep = keep_mask(c.inp["x"], c.inp["y"], c.meta["kwargs"]) & (np.hypot(c.inp["x"] + 0.7, c.inp["y"] - 0.3) > 0.3)
        scale = max(np.max(np.abs(c.out[q][keep])) for q in KEYS)
        for g, q in zip(got, KEYS):
            assert np.max(np.abs(g - c.out[q])[keep]) <= 1e-8 * scale, (name, q)
