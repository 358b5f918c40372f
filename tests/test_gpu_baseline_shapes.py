"""GPU parity at the BASELINE.json shapes themselves, against fixtures produced by the unmodified reference
(tests/golden/baseline_shapes.npz, `python tests/golden/make_golden.py baseline_shapes`), per pixel and on every set:

* headline: EPL+SHEAR, SERSIC_ELLIPSE, 180 x 180, ss 3, 21 x 21 pixel PSF, ImageLinearFit -- the `k_image<signature>`,
  `k_conv_tiled_u<21,...>` and `k_likelihood<1>` instantiations bench.py times -- for ten parameter sets spread over the
  nmax = 10 / 20 / 35 / 50 / 100 brackets of epl.py:308-319, through the kwargs API, the [B, P] batch entry points and
  the sampler boundary (prepare_logL_func);
* C3: EPL_MULTIPOLE_M1M3M4 + SHEAR, two SERSIC_ELLIPSE (n_lin 2);
* C4: 1 EPL + 1,000 TNFW through LensModelBulk on 10^4 rays of the 1000 x 1000 grid;
* C5: multi-plane (EPL, NFW_ELLIPSE_CSE, SIS) inside MultiLinear over 4 bands (multi_linear.py:112-161,
  multi_plane.py:236-275) -- the interpreted `k_image` path with the multi-plane recursion.
Bar: |a - b| <= 1e-10 |b| + 1e-12 max|b| per pixel; amplitudes and logL inherit the conditioning of the normal equations.
"""

import numpy as np
import pytest

from tests.golden_io import assert_close, load

pytestmark = pytest.mark.gpu

CASES = load("baseline_shapes.npz")
BAR = dict(rtol=1e-10, atol_frac=1e-12)


@pytest.fixture(scope="module", autouse=True)
def _native_loaded():
    from jaxtronomy_b200 import _lib

    _lib.load()
    import torch

    assert torch.cuda.is_available()
    before = _lib.launch_count()
    yield
    assert _lib.launch_count() > before, "GPU tests ran without launching a single lsim kernel"


def _static():
    from bench import static_config

    return static_config()


def _linear_fit(image_data, kernel, lens_list, source_list, lens_light_list, kwargs_numerics, linear=True, **lens_kw):
    from jaxtronomy_b200.Data.imaging_data import ImageData
    from jaxtronomy_b200.Data.psf import PSF
    from jaxtronomy_b200.ImSim.image_linear_solve import ImageLinearFit
    from jaxtronomy_b200.ImSim.image_model import ImageModel
    from jaxtronomy_b200.LensModel.lens_model import LensModel
    from jaxtronomy_b200.LightModel.light_model import LightModel

    cls = ImageLinearFit if linear else ImageModel
    return cls(ImageData(image_data=image_data, **_static()["kwargs_data"]),
               PSF(psf_type="PIXEL", kernel_point_source=kernel), LensModel(lens_list, **lens_kw),
               LightModel(source_list), LightModel(lens_light_list), kwargs_numerics=kwargs_numerics)


# ---------------------------------------------------------------------------------------------------------------
# headline
# ---------------------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def headline():
    from bench import LENS_LIST, SOURCE_LIST

    s = CASES["headline_static"]
    args = (s.inp["image_data"], s.inp["kernel"], LENS_LIST, SOURCE_LIST, [], s.meta["kwargs_numerics"])
    return dict(lin=_linear_fit(*args), im=_linear_fit(*args, linear=False), rows=s.inp["rows"])


def test_headline_truth_image_is_the_reference_simulation(headline):
    """bench.py builds its data image from `image_batch(TRUTH)`: that image against the reference's."""
    import torch

    s = CASES["headline_static"]
    got = headline["im"].image_batch(torch.as_tensor(s.inp["truth_row"][None, :], device="cuda"))[0].cpu().numpy()
    assert_close(got, s.out["truth"], what="truth image", **BAR)


@pytest.mark.parametrize("b", range(10))
def test_headline_per_pixel_kwargs_api(headline, b):
    from bench import row_to_kwargs

    c = CASES["headline_%02d" % b]
    kl, ks, kll = row_to_kwargs(c.inp["row"])
    lin, im = headline["lin"], headline["im"]
    assert_close(im.image(kl, ks, kll, unconvolved=True), c.out["image_unconvolved"], what="unconvolved", **BAR)
    A = lin.linear_response_matrix(kl, ks, kll)
    assert A.shape == c.out["A"].shape == (1, 32400)
    assert_close(A, c.out["A"], what="A", **BAR)
    model, _, _, param = lin.image_linear_solve(kl, ks, kll)
    assert_close(param, c.out["param"], what="param", rtol=1e-10, atol_frac=0)
    logL, param2 = lin.likelihood_data_given_model(kl, ks, kll)
    assert_close(logL, c.out["logL_linear"], what="logL_linear", rtol=1e-10, atol_frac=0)
    assert_close(im.likelihood_data_given_model(kl, ks, kll), c.out["logL"], what="logL", rtol=1e-10, atol_frac=0)
    if "wls_model" in c.out:
        assert_close(model, c.out["wls_model"], what="wls model", **BAR)
        assert_close(im.image(kl, ks, kll), c.out["image"], what="image", **BAR)


def test_headline_batch_entry_points_all_sets_per_pixel(headline):
    """One [10, 17] matrix with nmax 10 ... 100 mixed in one launch: basis images per pixel, amplitudes, logL, WLS model."""
    import torch

    lin, rows = headline["lin"], headline["rows"]
    params = torch.as_tensor(rows, device="cuda")
    basis = lin.response_batch(params).cpu().numpy()
    unconv = lin.response_batch(params, unconvolved=True).cpu().numpy()
    logl, amps, model = lin.logL_linear_batch(params, return_model=True)
    logl, amps, model = logl.cpu().numpy(), amps.cpu().numpy(), model.cpu().numpy()
    logl_j = headline["im"].logL_batch(params).cpu().numpy()
    for b in range(10):
        c = CASES["headline_%02d" % b]
        assert_close(basis[b, 0].reshape(-1), c.out["A"][0], what="basis %d" % b, **BAR)
        # rows carry amp = 1: the unconvolved basis image is the unconvolved image
        assert_close(unconv[b, 0], c.out["image_unconvolved"], what="unconvolved %d" % b, **BAR)
        assert_close(amps[b], c.out["param"], what="amps %d" % b, rtol=1e-10, atol_frac=0)
        assert_close(logl[b], c.out["logL_linear"], what="logL %d" % b, rtol=1e-10, atol_frac=0)
        assert_close(logl_j[b], c.out["logL"], what="logL (no solve) %d" % b, rtol=1e-10, atol_frac=0)
        if "wls_model" in c.out:
            assert_close(model[b], c.out["wls_model"], what="wls %d" % b, **BAR)
    # the same rows scattered through a batch that spans several chunks, in a different order: identical bits
    reps = 250
    order = np.random.default_rng(0).permutation(10 * reps)
    big = torch.as_tensor(np.tile(rows, (reps, 1))[order], device="cuda")
    l_big, a_big, _ = lin.logL_linear_batch(big)
    np.testing.assert_array_equal(l_big.cpu().numpy(), np.tile(logl, reps)[order])
    np.testing.assert_array_equal(a_big.cpu().numpy(), np.tile(amps, (reps, 1))[order])


def test_headline_through_the_sampler_boundary(headline):
    """prepare_logL_func("gpu", LikelihoodModule)(positions: np.ndarray[B, n]) -> np.ndarray[B]
    (Sampling/sampler.py:215-251): the call bench.py's e2e times."""
    from bench import headline_likelihood_module
    from jaxtronomy_b200.Sampling.sampler import prepare_logL_func

    s = CASES["headline_static"]
    like, positions_of = headline_likelihood_module(headline["lin"], e_bound=0.95)  # the fixture rows reach |e| = 0.8
    f = prepare_logL_func("gpu", like)
    pos = positions_of(headline["rows"])
    out = f(pos)
    assert isinstance(out, np.ndarray) and out.shape == (10,)
    for b in range(10):
        assert_close(out[b], CASES["headline_%02d" % b].out["logL_linear"], what="logL %d" % b, rtol=1e-10, atol_frac=0)
    # a row outside the prior box gets the reference's bound penalty (likelihood.py:374-405), its neighbours are untouched
    pos2 = pos.copy()
    pos2[3, 0] = -1.0
    out2 = f(pos2)
    assert out2[3] <= -1e17 and np.array_equal(np.delete(out2, 3), np.delete(out, 3))
    assert s.inp["rows"].shape == (10, 17)


# ---------------------------------------------------------------------------------------------------------------
# C3
# ---------------------------------------------------------------------------------------------------------------
def test_c3_multipole_two_components_per_pixel():
    import torch

    s = CASES["c3_static"]
    args = (s.inp["image_data"], s.inp["kernel"], s.meta["lens_model_list"], s.meta["source_model_list"],
            s.meta["lens_light_model_list"], s.meta["kwargs_numerics"])
    lin, im = _linear_fit(*args), _linear_fit(*args, linear=False)
    rows = []
    for b in range(3):
        c = CASES["c3_%02d" % b]
        kl, ks, kll = c.meta["kwargs_lens"], c.meta["kwargs_source"], c.meta["kwargs_lens_light"]
        A = lin.linear_response_matrix(kl, ks, kll)
        assert A.shape == (2, 32400)
        assert_close(A, c.out["A"], what="A %d" % b, **BAR)
        assert_close(im.image(kl, ks, kll, unconvolved=True), c.out["image_unconvolved"], what="unconvolved %d" % b, **BAR)
        cond = np.linalg.cond(c.out["A"] @ c.out["A"].T)
        ptol = max(1e-10, 10 * 2.2e-16 * cond)
        logL, param = lin.likelihood_data_given_model(kl, ks, kll)
        assert_close(param, c.out["param"], what="param %d" % b, rtol=ptol, atol_frac=ptol)
        assert_close(logL, c.out["logL_linear"], what="logL %d" % b, rtol=ptol, atol_frac=0)
        rows.append(lin._pack(kl, ks, kll)[0][0])
    # batched: the three sets in one launch
    logl, amps, _ = lin.logL_linear_batch(torch.as_tensor(np.stack(rows), device="cuda"))
    for b in range(3):
        c = CASES["c3_%02d" % b]
        assert_close(logl[b].item(), c.out["logL_linear"], what="batched logL %d" % b, rtol=1e-9, atol_frac=0)
        assert_close(amps[b].cpu().numpy(), c.out["param"], what="batched param %d" % b, rtol=1e-8, atol_frac=0)


# ---------------------------------------------------------------------------------------------------------------
# C4
# ---------------------------------------------------------------------------------------------------------------
def test_c4_thousand_subhalos():
    from jaxtronomy_b200.LensModel.lens_model_bulk import LensModelBulk
    from tests.test_oracle_baseline_shapes import c4_kwargs

    c = CASES["c4_substructure"]
    names, kws = c4_kwargs()
    lmb = LensModelBulk(c.meta["unique_lens_model_list"], multi_plane=False)
    rs = lmb.prepare_ray_shooting_kwargs(names, kws)
    np.testing.assert_array_equal(rs["index_list"], c.inp["index_list"])
    for k, v in rs["all_kwargs"].items():
        np.testing.assert_array_equal(v, c.inp["kw_" + k])
    bx, by = lmb.ray_shooting(c.inp["x"], c.inp["y"], rs)
    # 1,001 deflections of either sign add up: the absolute companion scales with the summed magnitude, not the net
    assert_close(bx, c.out["beta_x"], what="beta_x", rtol=1e-10, atol_frac=1e-11)
    assert_close(by, c.out["beta_y"], what="beta_y", rtol=1e-10, atol_frac=1e-11)


# ---------------------------------------------------------------------------------------------------------------
# C5
# ---------------------------------------------------------------------------------------------------------------
def _c5_multi_linear():
    from jaxtronomy_b200.ImSim.MultiBand.multi_linear import MultiLinear

    s = CASES["c5_static"]
    kd = _static()["kwargs_data"]
    bands = [[dict(kd, image_data=s.inp["image_data_%d" % b]),
              dict(psf_type="PIXEL", kernel_point_source=s.inp["kernel_%d" % b]), s.meta["band_numerics"][b]]
             for b in range(4)]
    ml = MultiLinear(bands, s.meta["kwargs_model"])
    for m in ml._imageModel_list:  # the fixture's cosmological constants (the reference takes them from astropy)
        lm = m.image_model.LensModel
        assert lm._sorted_redshift_index == [int(i) for i in s.inp["sorted_index"]]
        lm.set_multi_plane_constants(s.inp["T_ij_list"], s.inp["T_z_list"], s.inp["reduced2physical"],
                                     float(s.inp["T_ij_stop"]), float(s.inp["T_z_source"]))
    return ml


def test_c5_multi_plane_multi_band_per_pixel():
    ml = _c5_multi_linear()
    for n in range(3):
        c = CASES["c5_%02d" % n]
        kl, ks, kll = c.meta["kwargs_lens"], c.meta["kwargs_source"], c.meta["kwargs_lens_light"]
        logL, params = ml.likelihood_data_given_model(kl, ks, kll)
        worst = 1e-10
        for b in range(4):
            band = ml._imageModel_list[b]
            if "A_%d" % b in c.out:
                A = band.image_model.linear_response_matrix(kl, ks, kll)
                assert A.shape == (2, 32400)
                assert_close(A, c.out["A_%d" % b], what="set %d band %d A" % (n, b), **BAR)
                cond = np.linalg.cond(c.out["A_%d" % b] @ c.out["A_%d" % b].T)
                worst = max(worst, 10 * 2.2e-16 * cond)
            lb, pb = band.likelihood_data_given_model(kl, ks, kll)
            assert_close(pb, c.out["param_%d" % b], what="set %d band %d param" % (n, b), rtol=1e-8, atol_frac=0)
            assert_close(lb, c.out["logL_%d" % b], what="set %d band %d logL" % (n, b), rtol=1e-9, atol_frac=0)
            np.testing.assert_array_equal(params[b], pb)
        assert_close(logL, c.out["logL"], what="set %d logL" % n, rtol=1e-9, atol_frac=0)


def test_c5_batched_through_the_likelihood_module():
    """The path tools/bench_configs.py c5 and bench.py's `secondary` time: LikelihoodModule(MultiLinear, Param) on a
    [B, n] matrix -> one launch sequence per band, summed."""
    import torch

    from jaxtronomy_b200.Sampling.likelihood import LikelihoodModule
    from jaxtronomy_b200.Sampling.parameters import Param

    ml = _c5_multi_linear()
    km = CASES["c5_static"].meta["kwargs_model"]
    param = Param(km)
    like = LikelihoodModule(ml, param)
    pos = np.stack([param.kwargs2args(CASES["c5_%02d" % n].meta["kwargs_lens"], CASES["c5_%02d" % n].meta["kwargs_source"],
                                      CASES["c5_%02d" % n].meta["kwargs_lens_light"]) for n in range(3)])
    out = like.logL_batch_device(torch.as_tensor(np.tile(pos, (50, 1)), device="cuda")).cpu().numpy()
    for n in range(3):
        assert_close(out[n], CASES["c5_%02d" % n].out["logL"], what="logL %d" % n, rtol=1e-9, atol_frac=0)
    np.testing.assert_array_equal(out, np.tile(out[:3], 50))


# ---------------------------------------------------------------------------------------------------------------
# k_image_fold (linear profiles and light transforms composed per set, affine grid) against the plain k_image
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("case", ["headline", "epl_shear_convergence_two_lights_ss2", "sie_as_epl_ss1"])
def test_folded_image_kernel_matches_the_plain_one(case, monkeypatch):
    """The same plan built with and without LSIM_NO_FOLD: basis images per pixel at 1e-12 of the peak (the two kernels
    differ in operation order only), logL and amplitudes at 1e-11; the headline case is also pinned to the reference
    by the tests above, which run the folded kernel."""
    import torch

    from jaxtronomy_b200.Data.imaging_data import ImageData
    from jaxtronomy_b200.Data.psf import PSF
    from jaxtronomy_b200.ImSim.image_linear_solve import ImageLinearFit
    from jaxtronomy_b200.LensModel.lens_model import LensModel
    from jaxtronomy_b200.LightModel.light_model import LightModel

    s = CASES["headline_static"]
    rng = np.random.default_rng(17)
    kd = _static()["kwargs_data"]
    if case == "headline":
        lens, pk, src, ll, ss, n = ["EPL", "SHEAR"], None, ["SERSIC_ELLIPSE"], [], 3, 64
    elif case == "epl_shear_convergence_two_lights_ss2":
        lens, pk, src, ll, ss, n = ["SHEAR", "EPL", "CONVERGENCE"], None, ["SERSIC_ELLIPSE"], ["SERSIC"], 2, 16
    else:
        lens, pk, src, ll, ss, n = ["SIE", "SHEAR"], [dict(NIE=False, gamma=2.0), {}], ["SERSIC"], [], 1, 16

    def build():
        return ImageLinearFit(ImageData(image_data=s.inp["image_data"], **kd),
                              PSF(psf_type="PIXEL", kernel_point_source=s.inp["kernel"]),
                              LensModel(lens, profile_kwargs_list=pk), LightModel(src), LightModel(ll),
                              kwargs_numerics=dict(supersampling_factor=ss))

    folded = build()
    names = folded.param_names
    lo = dict(theta_E=0.8, gamma=1.7, e1=-0.3, e2=-0.3, center_x=-0.1, center_y=-0.1, gamma1=-0.08, gamma2=-0.08,
              ra_0=-0.2, dec_0=-0.2, kappa=-0.05, amp=1.0, R_sersic=0.1, n_sersic=1.0)
    hi = dict(theta_E=1.6, gamma=2.3, e1=0.3, e2=0.3, center_x=0.1, center_y=0.1, gamma1=0.08, gamma2=0.08,
              ra_0=0.2, dec_0=0.2, kappa=0.05, amp=1.0, R_sersic=0.5, n_sersic=4.0)
    mat = np.stack([rng.uniform(lo[nm], hi[nm], n) for _, _, nm in names], axis=1)
    params = torch.as_tensor(mat, device="cuda")
    img_f = folded.response_batch(params).cpu().numpy()
    l_f, a_f, _ = folded.logL_linear_batch(params)
    monkeypatch.setenv("LSIM_NO_FOLD", "1")
    plain = build()
    img_p = plain.response_batch(params).cpu().numpy()
    l_p, a_p, _ = plain.logL_linear_batch(params)
    assert not np.array_equal(img_f, img_p), "both plans ran the same kernel"
    scale = np.abs(img_p).max(axis=(2, 3), keepdims=True)
    assert np.all(np.abs(img_f - img_p) <= 1e-11 * np.abs(img_p) + 1e-12 * scale), \
        float(np.max(np.abs(img_f - img_p) / scale))
    np.testing.assert_allclose(l_f.cpu().numpy(), l_p.cpu().numpy(), rtol=1e-10)
    np.testing.assert_allclose(a_f.cpu().numpy(), a_p.cpu().numpy(), rtol=1e-9)
