"""GPU parity tests: the CUDA path, called through the C ABI (ctypes -> liblsim.so), against
  (a) the committed golden fixtures produced by the unmodified reference source, and
  (b) the NumPy oracle on seeded inputs (batched and at the headline size).
Bar (north star): |a - b| <= 1e-10 |b| + 1e-12 max|b|  (fp64; SURVEY.md fact 10 for the atol companion);
index / grid bookkeeping bit-exact.
"""

import numpy as np
import pytest

from tests.golden_io import assert_close, load

pytestmark = pytest.mark.gpu

PROFILES = load("profiles.npz")
LENS_MODELS = load("lens_models.npz")
IMSIM = load("imsim.npz")
BAR = dict(rtol=1e-10, atol_frac=1e-12)


@pytest.fixture(scope="module", autouse=True)
def _native_loaded():
    """Fail loudly if the CUDA extension is missing: there is no fallback to fall back to."""
    from jaxtronomy_b200 import _lib

    lib = _lib.load()
    import torch

    assert torch.cuda.is_available()
    before = _lib.launch_count()
    yield
    assert _lib.launch_count() > before, "GPU tests ran without launching a single lsim kernel"


def _lens_models():
    from jaxtronomy_b200.LensModel.lens_model import LensModel

    return LensModel


def _conditioning_rtol(prof, kw, x, y):
    from tests.test_hostmath import conditioning_rtol

    return conditioning_rtol(prof, kw, x, y)


# ---------------------------------------------------------------------------------------------------------------
# profiles
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", [n for n, c in PROFILES.items() if "f_x" in c.out and "profile" in c.meta])
def test_lens_profile_deflection(name):
    from jaxtronomy_b200 import _model

    c = PROFILES[name]
    prof = c.meta["profile"]
    kw = c.scalar_kwargs(_model.LENS_PARAM_NAMES[prof])
    lm = _lens_models()([prof], profile_kwargs_list=[c.meta["profile_kwargs"]])
    fx, fy = lm.alpha(c.inp["x"], c.inp["y"], [kw])
    scale = max(np.max(np.abs(c.out["f_x"])), np.max(np.abs(c.out["f_y"])))
    rtol = _conditioning_rtol(prof, kw, c.inp["x"], c.inp["y"])
    for got, want, lab in ((fx, c.out["f_x"], "f_x"), (fy, c.out["f_y"], "f_y")):
        err = np.abs(got - want)
        tol = rtol * np.abs(want) + 1e-12 * scale
        i = int(np.argmax(err - tol))
        assert np.all(err <= tol), "%s %s: |diff| %.3g > tol %.3g at %d (want %.17g got %.17g)" % (
            name, lab, err[i], tol[i], i, want[i], got[i])
    # ray_shooting = theta - alpha
    bx, by = lm.ray_shooting(c.inp["x"], c.inp["y"], [kw])
    np.testing.assert_allclose(bx, c.inp["x"] - fx, rtol=0, atol=1e-15 * max(1.0, scale))


def test_epl_major_axis_reference_sweep():
    """test_epl.py:199-206: q in [0.03, 1] x y in [-10, 10], through EPL with phi = 0 (e2 = 0, e1 = (1-q)/(1+q))."""
    c = PROFILES["epl_major_sweep"]
    lm = _lens_models()(["EPL"])
    b, t = float(c.inp["b"]), float(c.inp["t"])
    for i, q in enumerate(c.inp["q"]):
        q = max(float(q), (1 - 0.9999) / (1 + 0.9999))
        e1 = (1 - q) / (1 + q)
        kw = dict(theta_E=b / np.sqrt(q), gamma=t + 1, e1=e1, e2=0.0, center_x=0.0, center_y=0.0)
        fx, fy = lm.alpha(c.inp["x"], c.inp["y"], [kw])
        if float(c.inp["q"][i]) < q:
            continue  # below the ellipticity clamp of param_util.ellipticity2phi_q the profile frame differs
        scale = max(np.max(np.abs(c.out["f_x"][i])), np.max(np.abs(c.out["f_y"][i])))
        assert np.max(np.abs(fx - c.out["f_x"][i])) <= 1e-10 * scale, (q, np.max(np.abs(fx - c.out["f_x"][i])))
        assert np.max(np.abs(fy - c.out["f_y"][i])) <= 1e-10 * scale, q


@pytest.mark.parametrize("name", [n for n, c in PROFILES.items() if "flux" in c.out])
def test_light_profile(name):
    from jaxtronomy_b200 import _model
    from jaxtronomy_b200.LightModel.light_model import LightModel

    c = PROFILES[name]
    prof = c.meta["profile"]
    kw = c.scalar_kwargs(_model.LIGHT_PARAM_NAMES[prof])
    lm = LightModel([prof])
    assert_close(lm.surface_brightness(c.inp["x"], c.inp["y"], [kw]), c.out["flux"], what=name, **BAR)
    resp, n = lm.functions_split(c.inp["x"], c.inp["y"], [kw])
    assert n == 1
    assert_close(resp[0] * kw["amp"], c.out["flux"], what=name + " split", **BAR)


# ---------------------------------------------------------------------------------------------------------------
# lens models
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", [n for n in LENS_MODELS if n.startswith("single_plane")])
def test_single_plane_ray_shooting(name):
    c = LENS_MODELS[name]
    k = c.meta["k"]
    if isinstance(k, list) and not isinstance(k[0], bool):
        k = tuple(k)
    lm = _lens_models()(c.meta["lens_model_list"])
    bx, by = lm.ray_shooting(c.inp["x"], c.inp["y"], c.meta["kwargs"], k=k)
    assert_close(bx, c.out["beta_x"], what=name, **BAR)
    assert_close(by, c.out["beta_y"], what=name, **BAR)


def test_batched_kwargs_and_torch_io():
    """[B] arrays in kwargs = vmap over parameter sets; 2-d coordinate arrays keep their shape; torch in -> torch out."""
    import torch

    from oracle import lens_model as olm

    rng = np.random.default_rng(5)
    x, y = rng.uniform(-2, 2, (7, 9)), rng.uniform(-2, 2, (7, 9))
    B = 5
    th, g1 = rng.uniform(0.8, 1.5, B), rng.uniform(-0.05, 0.05, B)
    kw = [dict(theta_E=th, gamma=2.0, e1=0.1, e2=-0.2, center_x=0.0, center_y=0.05),
          dict(gamma1=g1, gamma2=0.02, ra_0=0.0, dec_0=0.0)]
    lm = _lens_models()(["EPL", "SHEAR"])
    bx, by = lm.ray_shooting(x, y, kw)
    assert bx.shape == (B, 7, 9)
    for b in range(B):
        kwb = [dict(kw[0], theta_E=th[b]), dict(kw[1], gamma1=g1[b])]
        ox, oy = olm.ray_shooting(["EPL", "SHEAR"], x, y, kwb)
        assert_close(bx[b], ox, what="batched %d" % b, **BAR)
        assert_close(by[b], oy, what="batched %d" % b, **BAR)
    tx, ty = lm.ray_shooting(torch.as_tensor(x, device="cuda"), torch.as_tensor(y, device="cuda"), kw)
    assert isinstance(tx, torch.Tensor) and tx.is_cuda and tx.shape == (B, 7, 9)
    np.testing.assert_array_equal(tx.cpu().numpy(), bx)
    # empty input
    ex, ey = lm.ray_shooting(np.zeros(0), np.zeros(0), kw)
    assert ex.shape == (B, 0)


def test_bulk_single_plane():
    from jaxtronomy_b200.LensModel.lens_model_bulk import LensModelBulk

    c = LENS_MODELS["bulk_single_plane"]
    lmb = LensModelBulk(c.meta["unique_lens_model_list"], multi_plane=False)
    rs = lmb.prepare_ray_shooting_kwargs(c.meta["lens_model_list"], c.meta["kwargs"],
                                         num_deflectors=c.meta["num_deflectors"])
    np.testing.assert_array_equal(rs["index_list"], c.inp["index_list"])
    for k, v in rs["all_kwargs"].items():
        np.testing.assert_array_equal(v, c.inp["kw_" + k])
    bx, by = lmb.ray_shooting(c.inp["x"], c.inp["y"], rs)
    assert_close(bx, c.out["beta_x"], what="bulk", **BAR)
    assert_close(by, c.out["beta_y"], what="bulk", **BAR)


def test_bulk_many_deflectors_streams_in_chunks():
    """> kLensChunk deflectors: 1 EPL + 300 TNFW against the oracle (substructure-notebook shape)."""
    from jaxtronomy_b200.LensModel.lens_model_bulk import LensModelBulk
    from oracle import lens_model as olm

    rng = np.random.default_rng(8)
    nh = 300
    names = ["EPL"] + ["TNFW"] * nh
    kws = [dict(theta_E=1.0, gamma=2.0, e1=0.1, e2=0.1, center_x=0.0, center_y=0.0)]
    kws += [dict(Rs=float(rng.uniform(0.01, 0.1)), alpha_Rs=float(rng.uniform(1e-4, 5e-3)),
                 r_trunc=float(rng.uniform(0.05, 0.5)), center_x=float(rng.uniform(-2, 2)),
                 center_y=float(rng.uniform(-2, 2))) for _ in range(nh)]
    x, y = rng.uniform(-2, 2, 2000), rng.uniform(-2, 2, 2000)
    lmb = LensModelBulk(["EPL", "TNFW"], multi_plane=False)
    bx, by = lmb.ray_shooting(x, y, lmb.prepare_ray_shooting_kwargs(names, kws, num_deflectors=320))
    ox, oy = olm.ray_shooting(names, x, y, kws)
    assert_close(bx, ox, what="bulk 300", rtol=1e-10, atol_frac=1e-11)
    assert_close(by, oy, what="bulk 300", rtol=1e-10, atol_frac=1e-11)


def test_multi_plane():
    c = LENS_MODELS["multi_plane"]
    lm = _lens_models()(c.meta["lens_model_list"], z_source=c.meta["z_source"],
                        lens_redshift_list=c.meta["lens_redshift_list"], multi_plane=True)
    # same cosmological constants as the fixture (the reference takes them from astropy)
    assert lm._sorted_redshift_index == [int(i) for i in c.inp["sorted_index"]]
    lm.set_multi_plane_constants(c.inp["T_ij_list"], c.inp["T_z_list"], c.inp["reduced2physical"],
                                 float(c.inp["T_ij_stop"]), float(c.inp["T_z_source"]))
    bx, by = lm.ray_shooting(c.inp["x"], c.inp["y"], c.meta["kwargs"])
    assert_close(bx, c.out["beta_x"], what="multi-plane", **BAR)
    assert_close(by, c.out["beta_y"], what="multi-plane", **BAR)


# ---------------------------------------------------------------------------------------------------------------
# numerics: re-bin + convolution
# ---------------------------------------------------------------------------------------------------------------
def _psf(meta_psf, inp):
    from jaxtronomy_b200.Data.psf import PSF

    kw = dict(meta_psf)
    if kw["psf_type"] == "PIXEL":
        kw["kernel_point_source"] = inp["kernel"]
    return PSF(**kw)


@pytest.mark.parametrize("name", [n for n in IMSIM if n.startswith("numerics_")])
def test_numerics_re_size_convolve(name):
    from jaxtronomy_b200.Data.pixel_grid import PixelGrid
    from jaxtronomy_b200.ImSim.Numerics.numerics import Numerics

    c = IMSIM[name]
    pg = PixelGrid(c.meta["nx"], c.meta["ny"], c.inp["transform_pix2angle"], float(c.inp["ra_at_xy_0"]),
                   float(c.inp["dec_at_xy_0"]))
    psf = _psf(c.meta["kwargs_psf"], c.inp)
    psf.set_pixel_size(pg.pixel_width)
    if c.meta.get("subframe"):
        from jaxtronomy_b200.ImSim.Numerics.numerics import NumericsSubFrame

        num = NumericsSubFrame(pg, psf, flux_evaluate_indexes=c.inp["flux_evaluate_indexes"],
                               **c.meta["kwargs_numerics"])
    else:
        num = Numerics(pg, psf, **c.meta["kwargs_numerics"])
    assert_close(num.re_size_convolve(c.inp["flux"]), c.out["image"], what=name, **BAR)
    assert_close(num.re_size_convolve(c.inp["flux"], unconvolved=True), c.out["image_unconvolved"], what=name, **BAR)
    # batched flux: leading axes are kept
    stack = np.stack([c.inp["flux"], 2 * c.inp["flux"]])
    out = num.re_size_convolve(stack)
    assert out.shape == (2,) + c.out["image"].shape
    assert_close(out[1], 2 * c.out["image"], what=name + " batched", **BAR)


@pytest.mark.parametrize("K,n,ss", [(21, 180, 1), (21, 97, 1), (7, 33, 1), (7, 40, 3), (11, 64, 1), (15, 61, 1), (25, 130, 1), (31, 70, 1)])
def test_fast_fir_convolution_matches_the_direct_one(K, n, ss, monkeypatch):
    """k_conv_tiled_fxy (the default) against k_conv_tiled_u (LSIM_CONV_NO_FF=1) and against the oracle's direct sum
    on random non-negative images with a non-symmetric kernel: every kernel-size template, odd / even frame sizes,
    partial tiles, and the supersampled-convolution path (accumulate mode).  The two kernels order their additions
    differently (lsim_kernels.cuh): 1e-13 of the image peak; the oracle bar is the usual 1e-10."""
    from jaxtronomy_b200.Data.pixel_grid import PixelGrid
    from jaxtronomy_b200.Data.psf import PSF
    from jaxtronomy_b200.ImSim.Numerics.numerics import Numerics
    from oracle import imsim as oimsim
    from oracle.lenstronomy_port import PSF as OPSF, PixelGrid as OPixelGrid

    rng = np.random.default_rng(K * 1000 + n)
    kern = rng.uniform(0.0, 1.0, (K, K)) * np.exp(-((np.arange(K) - K // 2) ** 2)[:, None] / (0.1 * K * K)
                                                   - ((np.arange(K) - K // 2) ** 2)[None, :] / (0.07 * K * K))
    if ss > 1:  # the supersampled kernel must have an odd size after degrading
        kern = rng.uniform(0.0, 1.0, (K * ss if (K * ss) % 2 else K * ss + 1,) * 2)
    M = np.array([[0.05, 0.0], [0.0, 0.05]])
    kw_num = dict(supersampling_factor=ss, supersampling_convolution=ss > 1)
    kw_psf = dict(psf_type="PIXEL", kernel_point_source=kern, point_source_supersampling_factor=ss)

    def run():
        pg = PixelGrid(n, n, M, -1.0, -1.0)
        psf = PSF(**kw_psf)
        psf.set_pixel_size(pg.pixel_width)
        num = Numerics(pg, psf, **kw_num)
        flux = rng2.uniform(0.0, 1.0, (3, n * ss * n * ss)) ** 3
        return flux, np.asarray(num.re_size_convolve(flux))

    rng2 = np.random.default_rng(5)
    flux, fast = run()
    monkeypatch.setenv("LSIM_CONV_NO_FF", "1")
    rng2 = np.random.default_rng(5)
    _, direct = run()
    monkeypatch.delenv("LSIM_CONV_NO_FF")
    assert np.abs(fast - direct).max() <= 1e-13 * np.abs(direct).max()
    # the fast path really is another kernel (K >= 27 goes through cuFFT whatever the switch says)
    assert not np.array_equal(fast, direct) or K >= 27
    opg = OPixelGrid(n, n, M, -1.0, -1.0)
    opsf = OPSF(**kw_psf)
    opsf.set_pixel_size(opg.pixel_width)
    onum = oimsim.Numerics(opg, opsf, **kw_num)
    assert_close(fast[1], onum.re_size_convolve(flux[1]), what="fast FIR vs oracle", **BAR)


# ---------------------------------------------------------------------------------------------------------------
# ImageModel / ImageLinearFit against the reference fixtures
# ---------------------------------------------------------------------------------------------------------------
def build_image_model(c, linear=True):
    from jaxtronomy_b200.Data.imaging_data import ImageData
    from jaxtronomy_b200.ImSim.image_linear_solve import ImageLinearFit
    from jaxtronomy_b200.ImSim.image_model import ImageModel
    from jaxtronomy_b200.LensModel.lens_model import LensModel
    from jaxtronomy_b200.LightModel.light_model import LightModel

    kd = dict(image_data=c.inp["image_data"], transform_pix2angle=c.inp["transform_pix2angle"],
              ra_at_xy_0=float(c.inp["ra_at_xy_0"]), dec_at_xy_0=float(c.inp["dec_at_xy_0"]),
              background_rms=float(c.inp["background_rms"]),
              exposure_time=c.inp["exposure_time"] if c.inp["exposure_time"].ndim else float(c.inp["exposure_time"]))
    if "noise_map" in c.inp:
        kd["noise_map"] = c.inp["noise_map"]
    cls = ImageLinearFit if linear else ImageModel
    kwargs_numerics = dict(c.meta["kwargs_numerics"])
    if "flux_evaluate_indexes" in c.inp:
        kwargs_numerics["flux_evaluate_indexes"] = c.inp["flux_evaluate_indexes"]
    return cls(ImageData(**kd), _psf(c.meta["kwargs_psf"], c.inp),
               LensModel(c.meta["lens_model_list"], profile_kwargs_list=c.meta.get("lens_profile_kwargs")),
               LightModel(c.meta["source_model_list"]), LightModel(c.meta["lens_light_model_list"]),
               kwargs_numerics=kwargs_numerics, likelihood_mask=c.inp.get("likelihood_mask"))


IMAGE_CASES = [n for n, c in IMSIM.items() if "logL_linear" in c.out]


@pytest.mark.parametrize("name", IMAGE_CASES)
def test_image_model(name):
    c = IMSIM[name]
    im = build_image_model(c, linear=False)
    kl, ks, kll = c.meta["kwargs_lens"], c.meta["kwargs_source"], c.meta["kwargs_lens_light"]
    assert_close(im.image(kl, ks, kll), c.out["image"], what=name + " image", **BAR)
    assert_close(im.image(kl, ks, kll, unconvolved=True), c.out["image_unconvolved"], what=name + " unconvolved", **BAR)
    assert_close(im.source_surface_brightness(ks, kl), c.out["source"], what=name + " source", **BAR)
    assert_close(im.image(kl, ks, kll, lens_light_add=False), c.out["source"], what=name + " source_add only", **BAR)
    if kll:
        assert_close(im.lens_surface_brightness(kll), c.out["lens_light"], what=name + " lens light", **BAR)
    logL = im.likelihood_data_given_model(kl, ks, kll)
    assert abs(logL - float(c.out["logL"])) <= 1e-10 * abs(float(c.out["logL"])), (logL, float(c.out["logL"]))


@pytest.mark.parametrize("name", IMAGE_CASES)
def test_image_linear_fit(name):
    c = IMSIM[name]
    lin = build_image_model(c, linear=True)
    kl, ks, kll = c.meta["kwargs_lens"], c.meta["kwargs_source"], c.meta["kwargs_lens_light"]
    A = lin.linear_response_matrix(kl, ks, kll)
    assert A.shape == c.out["A"].shape
    assert_close(A, c.out["A"], what=name + " response matrix", **BAR)
    model, model_error, cov, param = lin.image_linear_solve(kl, ks, kll)
    # amplitudes inherit the conditioning of the normal equations: the reference's own tests use 2e-9 on param
    cond = np.linalg.cond((c.out["A"] / 1.0) @ c.out["A"].T) if c.out["A"].shape[0] > 1 else 1.0
    ptol = max(1e-10, 10 * 2.2e-16 * cond)
    if np.all(c.out["param"] == 0):
        assert np.all(param == 0)  # cond gate of de_lens.py:26-33
    else:
        assert_close(param, c.out["param"], what=name + " param", rtol=ptol, atol_frac=ptol)
    assert_close(model, c.out["wls_model"], what=name + " wls model", rtol=max(1e-10, ptol), atol_frac=max(1e-12, ptol))
    assert cov is None and np.all(model_error == 0)
    logL, param2 = lin.likelihood_data_given_model(kl, ks, kll)
    want = float(c.out["logL_linear"])
    assert abs(logL - want) <= max(1e-10, ptol) * abs(want), (logL, want)
    np.testing.assert_array_equal(param2, param)
    logLm, _ = lin.likelihood_data_given_model(kl, ks, kll, source_marg=True)
    wantm = float(c.out["logL_marg"])
    assert abs(logLm - wantm) <= max(1e-10, ptol) * abs(wantm), (logLm, wantm)
    # covariance on request
    _, _, cov, _ = lin.image_linear_solve(kl, ks, kll, inv_bool=True)
    assert cov.shape == (len(param), len(param))


def test_batched_likelihood_matches_oracle_per_set():
    """A [B, P] parameter matrix through the batch entry points vs the oracle, one set at a time."""
    import torch

    from tests.test_oracle_golden import build_oracle_image_model

    c = IMSIM["headline_epl_shear_pixel"]
    lin = build_image_model(c, linear=True)
    om = build_oracle_image_model(c)
    kl, ks, kll = c.meta["kwargs_lens"], c.meta["kwargs_source"], c.meta["kwargs_lens_light"]
    names = lin.param_names
    base = np.array([({"lens": kl, "source": ks, "lens_light": kll}[comp][i])[n] for comp, i, n in names], dtype=float)
    rng = np.random.default_rng(2)
    B = 33
    mat = np.tile(base, (B, 1))
    free = [j for j, (comp, i, n) in enumerate(names) if n in ("theta_E", "gamma", "e1", "e2", "gamma1", "R_sersic")]
    mat[:, free] += rng.normal(0, 0.03, (B, len(free)))
    logl, amps, model = lin.logL_linear_batch(torch.as_tensor(mat, device="cuda"), return_model=True)
    logl_j = lin.logL_batch(torch.as_tensor(mat, device="cuda"))
    logl, amps, model, logl_j = logl.cpu().numpy(), amps.cpu().numpy(), model.cpu().numpy(), logl_j.cpu().numpy()
    for b in (0, 7, 32):
        row = mat[b]
        kws = {"lens": [dict() for _ in kl], "source": [dict() for _ in ks], "lens_light": [dict() for _ in kll]}
        for j, (comp, i, n) in enumerate(names):
            kws[comp][i][n] = row[j]
        ol, op = om.likelihood_data_given_model_linear(kws["lens"], kws["source"], kws["lens_light"])
        assert abs(logl[b] - ol) <= 1e-9 * abs(ol), (b, logl[b], ol)
        np.testing.assert_allclose(amps[b], op, rtol=1e-8)
        oj = om.likelihood_data_given_model(kws["lens"], kws["source"], kws["lens_light"])
        assert abs(logl_j[b] - oj) <= 1e-10 * abs(oj), (b, logl_j[b], oj)


# ---------------------------------------------------------------------------------------------------------------
# headline size: 180 x 180, ss 3, 21 x 21 PSF -- oracle spot checks + size-independent properties
# ---------------------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def headline():
    from bench import build_headline

    return build_headline(n_sets=296, seed=77)


def test_headline_matches_oracle(headline):
    from bench import oracle_headline_logl

    logl, amps, _ = headline["model"].logL_linear_batch(headline["params"])
    logl, amps = logl.cpu().numpy(), amps.cpu().numpy()
    assert np.all(np.isfinite(logl))
    for b in (0, 295):
        ol, op = oracle_headline_logl(headline, b)
        assert abs(logl[b] - ol) <= 1e-10 * abs(ol), (b, logl[b], ol)
        np.testing.assert_allclose(amps[b], op, rtol=1e-10)


def test_headline_properties(headline):
    import torch

    model, params = headline["model"], headline["params"]
    sub = params[:24].clone()
    img1 = model.image_batch(sub)
    # linearity in the amplitude (column 10 = source amp)
    sub2 = sub.clone()
    sub2[:, 10] *= 3.0
    img3 = model.image_batch(sub2)
    assert torch.allclose(img3, 3.0 * img1, rtol=1e-13, atol=0)
    # basis image (amp = 1) x solved amplitude = WLS model returned by the likelihood kernel
    logl, amps, wls = model.logL_linear_batch(sub, return_model=True)
    basis = model.response_batch(sub)[:, 0]
    assert torch.allclose(wls, amps[:, :1, None] * basis, rtol=1e-12, atol=1e-300)
    # batch order / chunking independence: one big call == concatenation of small calls
    l_a, _, _ = model.logL_linear_batch(params[:200])
    l_b, _, _ = model.logL_linear_batch(params[200:])
    l_all, _, _ = model.logL_linear_batch(params)
    assert torch.equal(torch.cat([l_a, l_b]), l_all)
    # convolution with a normalised PSF conserves flux up to what leaks over the frame edge
    unconv = model._split_plan(True).image(sub)[:, 0]
    assert torch.all(basis.sum(dim=(1, 2)) <= unconv.sum(dim=(1, 2)) * (1 + 1e-12))
    assert torch.all(basis.sum(dim=(1, 2)) >= unconv.sum(dim=(1, 2)) * 0.9)
    # determinism
    l_again, _, _ = model.logL_linear_batch(params)
    assert torch.equal(l_again, l_all)


def test_fp64_peak_microbenchmark_runs():
    from jaxtronomy_b200 import _lib

    flops = _lib.measure_fp64_peak(20000)
    assert 1e12 < flops < 1e14


# ---------------------------------------------------------------------------------------------------------------
# opt-in fp32 image kernel (jaxtronomy_b200.config.update("enable_x64", False)): stated looser tolerance
# ---------------------------------------------------------------------------------------------------------------
FP32_IMAGE_RTOL, FP32_IMAGE_ATOL_FRAC, FP32_LOGL_RTOL = 2e-4, 2e-5, 2e-4


def test_fp32_variant(headline):
    """fp32 image kernel vs the fp64 path (itself at 1e-10 of the reference) on the headline workload:
    images |a-b| <= 2e-4 |b| + 2e-5 max|b|, log-likelihood and amplitudes within 2e-4 relative."""
    import jaxtronomy_b200
    from bench import build_headline
    from jaxtronomy_b200 import _lib
    from jaxtronomy_b200.LensModel.lens_model import LensModel

    params = headline["params"][:64]
    img64 = headline["model"].image_batch(params).cpu().numpy()
    l64, a64, _ = headline["model"].logL_linear_batch(params)
    jaxtronomy_b200.config.update("enable_x64", False)
    try:
        m32 = build_headline(n_sets=4, seed=77, data_image=headline["data_image"])["model"]
        img32 = m32.image_batch(params).cpu().numpy()
        l32, a32, _ = m32.logL_linear_batch(params)
        # a model outside the fp32 kernel's scope is refused loudly, not run in fp64 behind the caller's back
        with pytest.raises(_lib.LsimError):
            from jaxtronomy_b200.Data.imaging_data import ImageData
            from jaxtronomy_b200.Data.psf import PSF
            from jaxtronomy_b200.ImSim.image_model import ImageModel
            from jaxtronomy_b200.LightModel.light_model import LightModel

            st = headline["static"]
            data = ImageData(image_data=headline["data_image"], **st["kwargs_data"])
            im = ImageModel(data, PSF(psf_type="NONE"), LensModel(["NFW"]), LightModel(["SERSIC"]), LightModel([]),
                            kwargs_numerics=dict(supersampling_factor=1))
            im.image([dict(Rs=1.0, alpha_Rs=1.0, center_x=0.0, center_y=0.0)],
                     [dict(amp=1.0, R_sersic=0.3, n_sersic=2.0, center_x=0.0, center_y=0.0)], [])
    finally:
        jaxtronomy_b200.config.update("enable_x64", True)
    tol = FP32_IMAGE_RTOL * np.abs(img64) + FP32_IMAGE_ATOL_FRAC * np.abs(img64).max(axis=(1, 2), keepdims=True)
    err = np.abs(img32 - img64)
    assert np.all(err <= tol), float(np.max(err / tol))
    assert not np.array_equal(img32, img64)  # it really is a different arithmetic
    l64, l32, a64, a32 = l64.cpu().numpy(), l32.cpu().numpy(), a64.cpu().numpy(), a32.cpu().numpy()
    assert np.max(np.abs(l32 - l64) / np.abs(l64)) <= FP32_LOGL_RTOL
    assert np.max(np.abs(a32 - a64) / np.abs(a64)) <= FP32_LOGL_RTOL


# ---------------------------------------------------------------------------------------------------------------
# FFT path of the PSF convolution (kernel side >= 27, or forced with LSIM_FFT_MIN_K)
# ---------------------------------------------------------------------------------------------------------------
def _numerics_with_pixel_psf(nx, ny, kernel, ss=1, supersampling_convolution=False):
    from jaxtronomy_b200.Data.pixel_grid import PixelGrid
    from jaxtronomy_b200.Data.psf import PSF
    from jaxtronomy_b200.ImSim.Numerics.numerics import Numerics

    pg = PixelGrid(nx, ny, np.array([[0.05, 0.0], [0.0, 0.05]]), -1.0, -1.0)
    psf = PSF(psf_type="PIXEL", kernel_point_source=kernel, point_source_supersampling_factor=1)
    return Numerics(pg, psf, supersampling_factor=ss, supersampling_convolution=supersampling_convolution,
                    convolution_kernel_size=kernel.shape[0])


def test_fft_convolution_matches_direct_and_the_oracle(monkeypatch):
    """A 45 x 45 asymmetric PSF on a non-square 70 x 52 frame takes the cuFFT path; the same kernel with the FFT path
    disabled takes the direct kernels; both against the oracle's convolution (scipy), zero boundary."""
    from oracle.imsim import PixelKernelConvolution

    rng = np.random.default_rng(5)
    nx, ny, K = 70, 52, 45
    kernel = rng.uniform(0.0, 1.0, (K, K)) * np.exp(-((np.arange(K) - 20.0) ** 2)[:, None] / 60.0)
    kernel /= kernel.sum()
    flux = rng.uniform(0, 1, (5, ny * nx))  # 5 images: the transform batch is padded with idle slots
    want = np.stack([PixelKernelConvolution(kernel).convolution2d(f.reshape(ny, nx)) for f in flux]) * 0.05**2
    got_fft = _numerics_with_pixel_psf(nx, ny, kernel).re_size_convolve(flux)
    monkeypatch.setenv("LSIM_FFT_MIN_K", "0")
    got_direct = _numerics_with_pixel_psf(nx, ny, kernel).re_size_convolve(flux)
    for got, what in ((got_fft, "fft"), (got_direct, "direct")):
        assert got.shape == want.shape
        assert_close(got, want, what=what, **BAR)
    assert not np.array_equal(got_fft, got_direct)  # two different algorithms really ran
    # forcing the FFT path on the headline-size kernel agrees with the tiled direct kernel
    k21 = kernel[12:33, 12:33] / kernel[12:33, 12:33].sum()
    direct21 = _numerics_with_pixel_psf(nx, ny, k21).re_size_convolve(flux)
    monkeypatch.setenv("LSIM_FFT_MIN_K", "5")
    fft21 = _numerics_with_pixel_psf(nx, ny, k21).re_size_convolve(flux)
    assert_close(fft21, direct21, what="fft vs tiled", **BAR)


# ---------------------------------------------------------------------------------------------------------------
# point sources against the reference fixtures
# ---------------------------------------------------------------------------------------------------------------
POINTS = load("point_sources.npz")


@pytest.mark.parametrize("name", [n for n in POINTS if n.startswith("rendering_")])
def test_point_source_rendering(name):
    """PointSourceRendering.point_source_rendering through ImageModel.point_source (k_point_sources): stamps inside,
    across the edge of and outside the frame, supersampling 1 and 3, convolved and unconvolved."""
    from jaxtronomy_b200.Data.imaging_data import ImageData
    from jaxtronomy_b200.Data.psf import PSF
    from jaxtronomy_b200.ImSim.image_model import ImageModel
    from jaxtronomy_b200.PointSource.point_source import PointSource

    c = POINTS[name]
    n, ss = c.meta["numPix"], c.meta["supersampling"]
    data = ImageData(image_data=np.zeros((n, n)), transform_pix2angle=c.inp["transform_pix2angle"],
                     ra_at_xy_0=float(c.inp["ra_at_xy_0"]), dec_at_xy_0=float(c.inp["dec_at_xy_0"]),
                     background_rms=1.0, exposure_time=1.0)
    psf = PSF(psf_type="PIXEL", kernel_point_source=c.inp["kernel"], point_source_supersampling_factor=ss)
    im = ImageModel(data, psf, point_source_class=PointSource(["UNLENSED"]),
                    kwargs_numerics=dict(point_source_supersampling_factor=ss))
    kps = [dict(ra_image=c.inp["ra"], dec_image=c.inp["dec"], point_amp=c.inp["amp"])]
    assert_close(im.point_source(kps), c.out["image"], what=name, **BAR)
    assert_close(im.point_source(kps, unconvolved=True), c.out["image_unconvolved"], what=name + " unconvolved", **BAR)
    assert_close(im.image(kwargs_ps=kps), c.out["image"], what=name + " via image()", **BAR)


def _ps_model(c, linear):
    from jaxtronomy_b200.Data.imaging_data import ImageData
    from jaxtronomy_b200.ImSim.image_linear_solve import ImageLinearFit
    from jaxtronomy_b200.ImSim.image_model import ImageModel
    from jaxtronomy_b200.LensModel.lens_model import LensModel
    from jaxtronomy_b200.LightModel.light_model import LightModel
    from jaxtronomy_b200.PointSource.point_source import PointSource

    kd = dict(image_data=c.inp["image_data"], transform_pix2angle=c.inp["transform_pix2angle"],
              ra_at_xy_0=float(c.inp["ra_at_xy_0"]), dec_at_xy_0=float(c.inp["dec_at_xy_0"]),
              background_rms=float(c.inp["background_rms"]), exposure_time=float(c.inp["exposure_time"]))
    lens = LensModel(c.meta["lens_model_list"])
    cls = ImageLinearFit if linear else ImageModel
    return cls(ImageData(**kd), _psf(c.meta["kwargs_psf"], c.inp), lens, LightModel(c.meta["source_model_list"]),
               LightModel(c.meta["lens_light_model_list"]),
               point_source_class=PointSource(c.meta["point_source_type_list"], lens_model=lens),
               kwargs_numerics=c.meta["kwargs_numerics"])


def test_image_model_with_point_sources():
    c = POINTS["image_quad_and_star"]
    kl, ks, kll = c.meta["kwargs_lens"], c.meta["kwargs_source"], c.meta["kwargs_lens_light"]
    kps = [{k: np.asarray(v, dtype=float) for k, v in kw.items()} for kw in c.meta["kwargs_ps"]]
    im = _ps_model(c, linear=False)
    assert_close(im.image(kl, ks, kll, kps), c.out["image"], what="image", **BAR)
    assert_close(im.image(kl, ks, kll, kps, unconvolved=True), c.out["image_unconvolved"], what="unconvolved", **BAR)
    assert_close(im.image(kl, ks, kll, kps, point_source_add=False), c.out["image_no_ps"], what="no ps", **BAR)
    assert_close(im.point_source(kps, kl), c.out["point_source"], what="point_source", **BAR)
    assert_close(im.point_source(kps, kl, k=1), c.out["point_source_k1"], what="point_source k=1", **BAR)
    logL = im.likelihood_data_given_model(kl, ks, kll, kps)
    assert abs(logL - float(c.out["logL"])) <= 1e-10 * abs(float(c.out["logL"]))
    lin = _ps_model(c, linear=True)
    assert lin.num_param_linear(kl, ks, kll, kps) == 7
    A = lin.linear_response_matrix(kl, ks, kll, kps)
    assert_close(A, c.out["A"], what="response matrix", **BAR)
    model, _, _, param = lin.image_linear_solve(kl, ks, kll, kps)
    assert_close(param, c.out["param"], what="param", rtol=1e-8, atol_frac=1e-10)
    assert_close(model, c.out["wls_model"], what="wls model", rtol=1e-8, atol_frac=1e-10)
    logL_lin, param2 = lin.likelihood_data_given_model(kl, ks, kll, kps)
    assert abs(logL_lin - float(c.out["logL_linear"])) <= 1e-9 * abs(float(c.out["logL_linear"]))
    # the solved amplitudes land in the kwargs like in the reference (update_linear_kwargs)
    _, ks2, kll2, kps2 = lin.update_linear_kwargs(param2, kl, [dict(k) for k in ks], [dict(k) for k in kll],
                                                  [dict(k) for k in kps])
    assert np.allclose(kps2[0]["point_amp"], param2[2:6]) and np.allclose(kps2[1]["point_amp"], param2[6:7])
    assert abs(im.likelihood_data_given_model(kl, ks2, kll2, kps2) - logL_lin) <= 1e-9 * abs(logL_lin)
    # batch entry point: [B, P] matrix with the point-source columns
    import torch

    lin.num_point_source_images = [4, 1]
    mat, _, _ = lin._pack(kl, ks, kll, kps, [4, 1])
    l_b, a_b, _ = lin.logL_linear_batch(torch.as_tensor(np.repeat(mat, 3, axis=0), device="cuda"))
    assert np.allclose(l_b.cpu().numpy(), logL_lin, rtol=1e-12) and np.allclose(a_b.cpu().numpy()[2], param2, rtol=1e-12)


@pytest.mark.parametrize("name", ["error_map_all", "error_map_first_only"])
def test_psf_error_map(name):
    """PSF error map (k_ps_error_map) and its use in k_likelihood: weights 1/(C_D + error), chi^2 on C_D(model) + |error|."""
    from jaxtronomy_b200.Data.imaging_data import ImageData
    from jaxtronomy_b200.Data.psf import PSF
    from jaxtronomy_b200.ImSim.image_linear_solve import ImageLinearFit
    from jaxtronomy_b200.LensModel.lens_model import LensModel
    from jaxtronomy_b200.LightModel.light_model import LightModel
    from jaxtronomy_b200.PointSource.point_source import PointSource

    c = POINTS[name]
    kd = dict(image_data=c.inp["image_data"], transform_pix2angle=c.inp["transform_pix2angle"],
              ra_at_xy_0=float(c.inp["ra_at_xy_0"]), dec_at_xy_0=float(c.inp["dec_at_xy_0"]),
              background_rms=float(c.inp["background_rms"]), exposure_time=float(c.inp["exposure_time"]))
    lens = LensModel(c.meta["lens_model_list"])
    psf = PSF(psf_type="PIXEL", kernel_point_source=c.inp["kernel"], psf_variance_map=c.inp["psf_variance_map"])
    lin = ImageLinearFit(ImageData(**kd), psf, lens, LightModel(c.meta["source_model_list"]),
                         LightModel(c.meta["lens_light_model_list"]),
                         point_source_class=PointSource(c.meta["point_source_type_list"], lens_model=lens),
                         kwargs_numerics=c.meta["kwargs_numerics"],
                         psf_error_map_bool_list=c.meta["psf_error_map_bool_list"])
    kl, ks, kll = c.meta["kwargs_lens"], c.meta["kwargs_source"], c.meta["kwargs_lens_light"]
    kps = [{k: np.asarray(v, dtype=float) for k, v in kw.items()} for kw in c.meta["kwargs_ps"]]
    assert_close(lin._error_map_psf(kl, kps), c.out["error_map"], what="error map", **BAR)
    # ImageModel.likelihood_data_given_model (amplitudes from the kwargs) through the parent class
    from jaxtronomy_b200.ImSim.image_model import ImageModel

    logL = ImageModel.likelihood_data_given_model(lin, kl, ks, kll, kps)
    assert abs(logL - float(c.out["logL"])) <= 1e-10 * abs(float(c.out["logL"]))
    model, model_error, _, param = lin.image_linear_solve(kl, ks, kll, kps)
    assert_close(model_error, c.out["model_error"], what="model error", **BAR)
    assert_close(param, c.out["param"], what="param", rtol=1e-8, atol_frac=1e-10)
    assert_close(model, c.out["wls_model"], what="wls model", rtol=1e-8, atol_frac=1e-10)
    logL_lin, _ = lin.likelihood_data_given_model(kl, ks, kll, kps)
    assert abs(logL_lin - float(c.out["logL_linear"])) <= 1e-9 * abs(float(c.out["logL_linear"]))


# ---------------------------------------------------------------------------------------------------------------
# batch-shape and robustness edge cases of the [B, P] entry points
# ---------------------------------------------------------------------------------------------------------------
def test_batch_edge_cases(headline):
    """Empty batch, one set, a ragged batch that spans several chunks, and rows with wild / nan parameters: the
    wild rows get finite or nan results of their own and never disturb their neighbours."""
    import torch

    from bench import oracle_headline_logl

    model, params = headline["model"], headline["params"]
    empty = params[:0]
    l0, a0, _ = model.logL_linear_batch(empty)
    assert l0.shape == (0,) and a0.shape == (0, 1) and model.image_batch(empty).shape == (0, 180, 180)
    l1, a1, _ = model.logL_linear_batch(params[:1])
    ref, _ = oracle_headline_logl(headline, 0)
    assert abs(float(l1[0]) - ref) <= 1e-10 * abs(ref)
    # ragged: one full chunk (64 x SMs sets, lsim_api.cu chunk_sets) + an eighth of a chunk + 37 sets, built by tiling
    # the 296 fixture rows
    n_sm = torch.cuda.get_device_properties(0).multi_processor_count
    n_big = 64 * n_sm + 8 * n_sm + 37
    reps = -(-n_big // params.shape[0])
    big = params.repeat(reps, 1)[:n_big].contiguous()
    lb, _, _ = model.logL_linear_batch(big)
    base, _, _ = model.logL_linear_batch(params)
    assert torch.equal(lb, base.repeat(reps)[:n_big])
    # wild rows
    wild = params[:8].clone()
    wild[1, 0] = float("nan")      # theta_E
    wild[2, 11] = 1e-9             # R_sersic far below its clamp (sersic_utils.py: R_sersic >= 1e-4)
    wild[3, 12] = 0.2              # n_sersic below the usual range: steep profile, exponent clamps
    wild[4, 2:4] = 0.0             # e1 = e2 = 0: f = 0, the series collapses to its first term
    wild[5, 0] = 1e6               # absurd Einstein radius
    wild[6, 1] = 2.999             # gamma at the edge of the power-law range
    lw, aw, _ = model.logL_linear_batch(wild)
    clean, _, _ = model.logL_linear_batch(params[:8])
    for b in (0, 7):
        assert lw[b] == clean[b]
    assert all(np.isfinite(float(lw[b])) for b in (2, 3, 4, 5, 6)), lw
    assert not np.isfinite(float(lw[1])) or abs(float(lw[1])) > 0  # a nan row stays confined to itself
    # the e = 0 and gamma-edge rows against the oracle
    headline_w = dict(headline, params_host=wild.cpu().numpy())
    for b in (4, 6):
        ref, ramp = oracle_headline_logl(headline_w, b)
        assert abs(float(lw[b]) - ref) <= 1e-9 * abs(ref), (b, float(lw[b]), ref)
        np.testing.assert_allclose(aw[b].cpu().numpy(), ramp, rtol=1e-8)


def test_point_source_model_without_linear_flux_keeps_the_solve_intact():
    """flux_from_point_source_list=[False]: the model's images own no basis image (point_source.py:296,361,398) -- the
    linear system, its amplitudes and logL are those of the same model without point sources (round-1 advisor finding:
    the empty slots used to zero every amplitude through the conditioning gate)."""
    from jaxtronomy_b200.Data.imaging_data import ImageData
    from jaxtronomy_b200.ImSim.image_linear_solve import ImageLinearFit
    from jaxtronomy_b200.LensModel.lens_model import LensModel
    from jaxtronomy_b200.LightModel.light_model import LightModel
    from jaxtronomy_b200.PointSource.point_source import PointSource

    c = POINTS["image_quad_and_star"]
    kl, ks, kll = c.meta["kwargs_lens"], c.meta["kwargs_source"], c.meta["kwargs_lens_light"]
    kps = [{k: np.asarray(v, dtype=float) for k, v in kw.items()} for kw in c.meta["kwargs_ps"]]
    kd = dict(image_data=c.inp["image_data"], transform_pix2angle=c.inp["transform_pix2angle"],
              ra_at_xy_0=float(c.inp["ra_at_xy_0"]), dec_at_xy_0=float(c.inp["dec_at_xy_0"]),
              background_rms=float(c.inp["background_rms"]), exposure_time=float(c.inp["exposure_time"]))
    lens = LensModel(c.meta["lens_model_list"])

    def build(ps):
        return ImageLinearFit(ImageData(**kd), _psf(c.meta["kwargs_psf"], c.inp), lens,
                              LightModel(c.meta["source_model_list"]), LightModel(c.meta["lens_light_model_list"]),
                              point_source_class=ps, kwargs_numerics=c.meta["kwargs_numerics"])

    plain = build(None)
    l0, p0 = plain.likelihood_data_given_model(kl, ks, kll)
    off = build(PointSource(c.meta["point_source_type_list"], lens_model=lens, flux_from_point_source_list=[False, False]))
    assert off.num_param_linear(kl, ks, kll, kps) == 2
    l1, p1 = off.likelihood_data_given_model(kl, ks, kll, kps)
    assert p1.shape == p0.shape == (2,) and np.all(p1 != 0)
    np.testing.assert_allclose(p1, p0, rtol=1e-12)
    assert abs(l1 - l0) <= 1e-12 * abs(l0)
    # one model in, one out: the quad's four images are the only point-source basis images
    mixed = build(PointSource(c.meta["point_source_type_list"], lens_model=lens, flux_from_point_source_list=[True, False]))
    assert mixed.num_param_linear(kl, ks, kll, kps) == 6
    A = mixed.linear_response_matrix(kl, ks, kll, kps)
    assert A.shape[0] == 6 and np.all(np.abs(A).sum(axis=1) > 0)
    l2, p2 = mixed.likelihood_data_given_model(kl, ks, kll, kps)
    assert p2.shape == (6,) and np.all(p2 != 0) and np.isfinite(l2)
