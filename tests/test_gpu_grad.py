"""GPU: lsim_logl_grad (log-likelihood and its gradient w.r.t. the parameter matrix) through the
C ABI against the NumPy oracle.

The reference obtains the gradient from jax autodiff, which cannot run in this environment, so the pin is:
  * value: logL / amplitudes of the gradient call == the forward call (same kernels) and the oracle at the 1e-10 bar;
  * gradient: 4th-order central finite differences of the ORACLE's log-likelihood in every parameter, at the steps
    1e-4, 1e-5 and 1e-6 (the closest one counts: a step that is too large sees the curvature of the Sersic cusp at the
    source centre and of the singular isothermal centre, one that is too small sees fp64 round-off).  Stated
    tolerance: |g - g_fd| <= 1e-7 max|g| for the smooth configurations (measured: ~1e-10), 2e-4 for the fixture
    whose rays cross two steep (n ~ 4) Sersic cusps behind a singular SIS centre, where the finite difference itself
    converges only slowly towards the analytic value as the step shrinks (tools/grad_debug.py prints the table);
  * at the headline size (180 x 180, ss 3, 21 x 21 PSF): central differences of the GPU forward path itself."""

import numpy as np
import pytest

from tests.golden_io import load

pytestmark = pytest.mark.gpu

IMSIM = load("imsim.npz")
GRAD_TOL = 1e-7
CUSP_TOL = {"linear_sis_3comp": 2e-4}
STEPS = (1e-4, 1e-5, 1e-6)


@pytest.fixture(scope="module", autouse=True)
def _native_loaded():
    from jaxtronomy_b200 import _lib

    _lib.load()
    import torch

    assert torch.cuda.is_available()
    before = _lib.launch_count()
    yield
    assert _lib.launch_count() > before


def _base_row(c, names):
    kws = {"lens": c.meta["kwargs_lens"], "source": c.meta["kwargs_source"], "lens_light": c.meta["kwargs_lens_light"]}
    return np.array([kws[comp][i][n] for comp, i, n in names], dtype=float)


def _row_to_kwargs(c, names, row):
    kws = {"lens": [dict() for _ in c.meta["kwargs_lens"]], "source": [dict() for _ in c.meta["kwargs_source"]],
           "lens_light": [dict() for _ in c.meta["kwargs_lens_light"]]}
    for j, (comp, i, n) in enumerate(names):
        kws[comp][i][n] = row[j]
    return kws["lens"], kws["source"], kws["lens_light"]


def _oracle_logl(om, c, names, row, linear):
    kl, ks, kll = _row_to_kwargs(c, names, row)
    if linear:
        return om.likelihood_data_given_model_linear(kl, ks, kll)[0]
    return om.likelihood_data_given_model(kl, ks, kll)


def _fd(fun, row, j, h):
    def at(s):
        r = row.copy()
        r[j] += s
        return fun(r)
    return (8.0 * (at(h) - at(-h)) - (at(2 * h) - at(-2 * h))) / (12.0 * h)


def _perturbed(base, names, rng, B):
    mat = np.tile(base, (B, 1))
    for j, (comp, i, n) in enumerate(names):
        if n in ("theta_E", "gamma", "R_sersic", "n_sersic"):
            mat[1:, j] *= 1 + rng.normal(0, 0.02, B - 1)
        elif n in ("e1", "e2", "gamma1", "gamma2", "center_x", "center_y", "a1_a", "a3_a", "a4_a", "delta_phi_m1",
                   "delta_phi_m3", "delta_phi_m4"):
            mat[1:, j] += rng.normal(0, 0.01, B - 1)
    return mat


GRAD_CASES = ["headline_epl_shear_pixel", "c1_sie_shear_gauss", "linear_sis_3comp", "subframe_epl_shear",
              "c3_multipole_noise_map"]


@pytest.mark.parametrize("linear", [True, False])
@pytest.mark.parametrize("name", GRAD_CASES)
def test_gradient_matches_oracle_finite_differences(name, linear):
    import torch

    from tests.test_gpu_parity import build_image_model
    from tests.test_oracle_golden import build_oracle_image_model

    c = IMSIM[name]
    lin = build_image_model(c, linear=True)
    om = build_oracle_image_model(c)
    names = lin.param_names
    rng = np.random.default_rng(5)
    B = 3
    mat = _perturbed(_base_row(c, names), names, rng, B)
    t = torch.as_tensor(mat, device="cuda")
    if linear:
        logl, grad, amps = lin.logL_linear_grad_batch(t)
        logl_f, amps_f, _ = lin.logL_linear_batch(t)
        torch.testing.assert_close(amps, amps_f, rtol=1e-13, atol=0)  # other kernel, other reduction order
    else:
        logl, grad = lin.logL_grad_batch(t)
        logl_f = lin.logL_batch(t)
    # the value of the gradient call is the forward path's value
    np.testing.assert_allclose(logl.cpu().numpy(), logl_f.cpu().numpy(), rtol=1e-13)
    logl, grad = logl.cpu().numpy(), grad.cpu().numpy()
    assert np.all(np.isfinite(grad))
    worst = 0.0
    for b in (0, B - 1):
        row = mat[b]
        ol = _oracle_logl(om, c, names, row, linear)
        assert abs(logl[b] - ol) <= 1e-9 * abs(ol), (b, logl[b], ol)
        scale = np.abs(grad[b]).max()
        for j, (comp, i, n) in enumerate(names):
            if linear and n == "amp":
                assert grad[b, j] == 0.0  # solved per set, not a parameter
                continue
            fds = [_fd(lambda r: _oracle_logl(om, c, names, r, linear), row, j, h * max(1.0, abs(row[j]))) for h in STEPS]
            err = min(abs(grad[b, j] - f) for f in fds) / scale
            worst = max(worst, err)
            assert err <= CUSP_TOL.get(name, GRAD_TOL), (name, linear, b, n, err, grad[b, j], fds)
    print("gradient vs oracle finite differences: %s linear=%s worst relative error %.2e" % (name, linear, worst))


def test_gradient_of_a_gated_linear_solve_is_zero():
    """A degenerate basis (two identical source components) fails the reference's stability gate: the amplitudes are
    forced to 0 (de_lens.py:91-105), the model is identically 0 and so is the gradient (jnp.where passes the tangent
    of the selected, constant branch)."""
    import torch

    from tests.test_gpu_parity import build_image_model
    from tests.test_oracle_golden import build_oracle_image_model

    c = IMSIM["linear_degenerate"]
    lin = build_image_model(c, linear=True)
    names = lin.param_names
    row = _base_row(c, names)
    logl, grad, amps = lin.logL_linear_grad_batch(torch.as_tensor(row[None, :], device="cuda"))
    assert not amps.any() and not grad.any() and torch.isfinite(logl).all()
    om = build_oracle_image_model(c)
    ol = _oracle_logl(om, c, names, row, True)
    assert abs(float(logl[0]) - ol) <= 1e-10 * abs(ol)
    bumped = np.tile(row, (2, 1))
    bumped[1, 0] += 1e-3  # the likelihood is flat there: the forward path agrees with the zero gradient
    lb = lin.logL_linear_batch(torch.as_tensor(bumped, device="cuda"))[0]
    assert float(lb[0]) == float(lb[1]) == float(logl[0])


def test_gradient_is_reproducible_and_chunked_batches_agree():
    import torch

    from tests.test_gpu_parity import build_image_model

    c = IMSIM["headline_epl_shear_pixel"]
    lin = build_image_model(c, linear=True)
    names = lin.param_names
    mat = _perturbed(_base_row(c, names), names, np.random.default_rng(9), 37)
    t = torch.as_tensor(mat, device="cuda")
    l1, g1, a1 = lin.logL_linear_grad_batch(t)
    l2, g2, a2 = lin.logL_linear_grad_batch(t)
    assert torch.equal(l1, l2) and torch.equal(g1, g2) and torch.equal(a1, a2)  # fixed reduction order, no atomics
    # a set's gradient does not depend on its position in the batch up to the tile split (different CTA partition)
    l3, g3, _ = lin.logL_linear_grad_batch(t[5:6])
    torch.testing.assert_close(g3[0], g1[5], rtol=1e-11, atol=1e-9 * float(g1[5].abs().max()))
    assert lin.logL_linear_grad_batch(t[:0])[1].shape == (0, len(names))


def test_gradient_batches_larger_than_one_chunk_and_small_workspaces():
    """The C call walks the batch in chunks of 8 x SMs sets (fewer with a small workspace): rows on both sides of a
    chunk boundary must equal the same rows evaluated alone, and a workspace for ONE set must give the same numbers."""
    import ctypes as C

    import torch

    from jaxtronomy_b200 import _lib
    from tests.test_gpu_parity import build_image_model

    c = IMSIM["subframe_epl_shear"]
    lin = build_image_model(c, linear=True)
    names = lin.param_names
    n_sm = torch.cuda.get_device_properties(0).multi_processor_count
    B = 8 * n_sm + 37
    mat = _perturbed(_base_row(c, names), names, np.random.default_rng(4), B)
    t = torch.as_tensor(mat, device="cuda")
    logl, grad, amps = lin.logL_linear_grad_batch(t)
    assert torch.isfinite(grad).all()
    for b in (0, 8 * n_sm - 1, 8 * n_sm, B - 1):
        l1, g1, a1 = lin.logL_linear_grad_batch(t[b:b + 1])
        assert torch.equal(l1[0], logl[b]) and torch.equal(a1[0], amps[b])
        torch.testing.assert_close(g1[0], grad[b], rtol=1e-11, atol=1e-11 * float(grad[b].abs().max()))
    # the same call through the raw C ABI with the smallest workspace the library accepts (chunks of one set)
    plan = lin._split_plan()
    lib = _lib.load()
    one = int(lib.lsim_logl_grad_workspace_bytes(plan._handle, 1))
    ws = torch.empty(one // 8 + 1, dtype=torch.float64, device="cuda")
    n = 5
    l2 = torch.empty(n, dtype=torch.float64, device="cuda")
    g2 = torch.empty((n, len(names)), dtype=torch.float64, device="cuda")
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    rc = lib.lsim_logl_grad(plan._handle, t.data_ptr(), n, 1, l2.data_ptr(), g2.data_ptr(), None, ws.data_ptr(), one,
                            stream)
    assert rc == 0
    assert torch.equal(l2, logl[:n])
    torch.testing.assert_close(g2, grad[:n], rtol=1e-11, atol=1e-11 * float(grad[:n].abs().max()))
    # too small a workspace is an error code, not a crash
    assert lib.lsim_logl_grad(plan._handle, t.data_ptr(), n, 1, l2.data_ptr(), g2.data_ptr(), None, ws.data_ptr(),
                              one // 2, stream) == -4


def test_gradient_with_likelihood_mask_and_noise_map():
    """Masked pixels carry no gradient; a noise map makes C_D independent of the model."""
    import torch

    from bench import build_headline

    cfg = build_headline(num_pix=48, ss=2, kernel_size=9, n_sets=4, seed=11)
    from jaxtronomy_b200.Data.imaging_data import ImageData
    from jaxtronomy_b200.Data.psf import PSF
    from jaxtronomy_b200.ImSim.image_linear_solve import ImageLinearFit
    from jaxtronomy_b200.LensModel.lens_model import LensModel
    from jaxtronomy_b200.LightModel.light_model import LightModel

    st = cfg["static"]
    rng = np.random.default_rng(3)
    mask = (rng.uniform(size=(48, 48)) > 0.3).astype(float)
    kd = dict(st["kwargs_data"])
    kd.pop("exposure_time")
    kd.pop("background_rms")
    data = ImageData(image_data=cfg["data_image"], noise_map=0.05 + 0.02 * rng.uniform(size=(48, 48)), **kd)
    lin = ImageLinearFit(data, PSF(psf_type="PIXEL", kernel_point_source=st["kernel"]), LensModel(["EPL", "SHEAR"]),
                         LightModel(["SERSIC_ELLIPSE"]), LightModel([]),
                         kwargs_numerics=dict(supersampling_factor=2), likelihood_mask=mask)
    t = cfg["params"].clone()
    t[:, 10] = 15.0
    for linear in (True, False):
        def value(m):
            return (lin.logL_linear_batch(m)[0] if linear else lin.logL_batch(m)).cpu().numpy()
        grad = (lin.logL_linear_grad_batch(t)[1] if linear else lin.logL_grad_batch(t)[1]).cpu().numpy()
        scale = np.abs(grad).max(axis=1)
        for j in (0, 1, 2, 5, 6, 11, 12, 13, 15) + (() if linear else (10,)):
            err = np.full(len(scale), np.inf)
            for h in STEPS:
                d = torch.zeros_like(t)
                d[:, j] = h
                fd = (8 * (value(t + d) - value(t - d)) - (value(t + 2 * d) - value(t - 2 * d))) / (12 * h)
                err = np.minimum(err, np.abs(grad[:, j] - fd) / scale)
            assert np.all(err <= 1e-6), (linear, j, err)


def test_headline_size_gradient_against_gpu_central_differences():
    """180 x 180, ss 3, 21 x 21 PSF: the gradient against central differences of the forward path on the same device."""
    import torch

    from bench import build_headline

    cfg = build_headline(n_sets=8, seed=21)
    lin, t = cfg["model"], cfg["params"]
    logl, grad, amps = lin.logL_linear_grad_batch(t)
    logl_f, amps_f, _ = lin.logL_linear_batch(t)
    torch.testing.assert_close(amps, amps_f, rtol=1e-13, atol=0)
    np.testing.assert_allclose(logl.cpu().numpy(), logl_f.cpu().numpy(), rtol=1e-13)
    grad = grad.cpu().numpy()
    assert np.all(np.isfinite(grad)) and not grad[:, 10].any()  # amp is solved per set, not a parameter
    scale = np.abs(grad).max(axis=1)
    worst = 0.0
    v = lambda m: lin.logL_linear_batch(m)[0].cpu().numpy()  # noqa: E731
    for j in [0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 11, 12, 13, 14, 15, 16]:
        err = np.full(len(scale), np.inf)
        for h in STEPS:
            d = torch.zeros_like(t)
            d[:, j] = h
            fd = (8 * (v(t + d) - v(t - d)) - (v(t + 2 * d) - v(t - 2 * d))) / (12 * h)
            err = np.minimum(err, np.abs(grad[:, j] - fd) / scale)
        worst = max(worst, err.max())
        assert np.all(err <= 1e-6), (j, err)
    print("headline-size gradient vs central differences of the forward path: worst relative error %.2e" % worst)


def test_models_outside_the_gradient_scope_are_refused():
    import torch

    from jaxtronomy_b200 import _lib
    from tests.test_gpu_parity import build_image_model

    c = IMSIM["ss_conv_sie_shear"]  # supersampled convolution
    lin = build_image_model(c, linear=True)
    row = _base_row(c, lin.param_names)
    with pytest.raises(_lib.LsimError):
        lin.logL_linear_grad_batch(torch.as_tensor(row[None, :], device="cuda"))


def test_dual_number_profiles_on_the_device():
    """NFW + TNFW + NFW_ELLIPSE_CSE + SHEAR (four lens profiles: the generic instantiation that evaluates every profile
    twice, all on dual numbers) against central differences of the forward path on the device."""
    import torch

    from bench import build_headline
    from jaxtronomy_b200.Data.imaging_data import ImageData
    from jaxtronomy_b200.Data.psf import PSF
    from jaxtronomy_b200.ImSim.image_linear_solve import ImageLinearFit
    from jaxtronomy_b200.LensModel.lens_model import LensModel
    from jaxtronomy_b200.LightModel.light_model import LightModel

    cfg = build_headline(num_pix=32, ss=2, kernel_size=7, n_sets=1, seed=1)
    st = cfg["static"]
    lin = ImageLinearFit(ImageData(image_data=cfg["data_image"], **st["kwargs_data"]),
                         PSF(psf_type="PIXEL", kernel_point_source=st["kernel"]),
                         LensModel(["NFW", "TNFW", "NFW_ELLIPSE_CSE", "SHEAR"]),
                         LightModel(["SERSIC_ELLIPSE"]), LightModel(["GAUSSIAN"]),
                         kwargs_numerics=dict(supersampling_factor=2))
    names = lin.param_names
    vals = {"lens": [dict(Rs=0.6, alpha_Rs=0.5, center_x=0.02, center_y=-0.01),
                     dict(Rs=0.3, alpha_Rs=0.2, r_trunc=0.9, center_x=0.3, center_y=0.2),
                     dict(Rs=0.8, alpha_Rs=0.4, e1=0.1, e2=-0.05, center_x=-0.02, center_y=0.03),
                     dict(gamma1=0.03, gamma2=-0.02, ra_0=0.0, dec_0=0.0)],
            "source": [dict(amp=1.0, R_sersic=0.25, n_sersic=1.5, e1=0.1, e2=0.05, center_x=0.03, center_y=0.06)],
            "lens_light": [dict(amp=1.0, sigma=0.3, center_x=0.0, center_y=0.0)]}
    row = np.array([vals[comp][i][n] for comp, i, n in names])
    t = torch.as_tensor(np.tile(row, (3, 1)), device="cuda")
    t[1:] += 0.01 * torch.as_tensor(np.random.default_rng(2).normal(size=(2, len(row))), device="cuda")
    logl, grad, _ = lin.logL_linear_grad_batch(t)
    assert torch.isfinite(logl).all() and torch.isfinite(grad).all()
    grad = grad.cpu().numpy()
    scale = np.abs(grad).max(axis=1)
    v = lambda m: lin.logL_linear_batch(m)[0].cpu().numpy()  # noqa: E731
    for j, (comp, i, n) in enumerate(names):
        if n == "amp":
            continue
        err = np.full(3, np.inf)
        for h in STEPS:
            d = torch.zeros_like(t)
            d[:, j] = h
            fdv = (8 * (v(t + d) - v(t - d)) - (v(t + 2 * d) - v(t - 2 * d))) / (12 * h)
            err = np.minimum(err, np.abs(grad[:, j] - fdv) / scale)
        assert np.all(err <= 1e-6), (comp, i, n, err)


def test_gradient_descent_sampler_recovers_the_truth():
    """OptaxMinimizer (interface of jaxtronomy/Sampling/Samplers/optax.py) on a 48 x 48 EPL + SHEAR + SERSIC_ELLIPSE
    fit: every chain starts a few per cent off and must end at a log-likelihood at least as good as the truth's."""
    import torch

    from bench import TRUTH, build_headline
    from jaxtronomy_b200.Sampling.likelihood import LikelihoodModule
    from jaxtronomy_b200.Sampling.parameters import Param
    from jaxtronomy_b200.Sampling.Samplers.optax import OptaxMinimizer

    cfg = build_headline(num_pix=48, ss=2, kernel_size=9, n_sets=1, seed=1)
    lin = cfg["model"]
    kwargs_model = dict(lens_model_list=["EPL", "SHEAR"], source_light_model_list=["SERSIC_ELLIPSE"])
    fixed_lens = [{}, {"ra_0": 0.0, "dec_0": 0.0}]
    param = Param(kwargs_model, kwargs_fixed_lens=fixed_lens, kwargs_fixed_source=[{}], linear_solver=True)
    like = LikelihoodModule(lin, param, check_bounds=True, linear_solver=True)
    from bench import row_to_kwargs

    kl, ks, _ = row_to_kwargs(TRUTH)
    truth = np.asarray(param.kwargs2args(kwargs_lens=kl, kwargs_source=ks))
    l_truth, g_truth = like.value_and_grad(truth)
    assert np.isfinite(l_truth) and np.all(np.isfinite(g_truth)) and len(g_truth) == len(truth)
    # gradient of the module == central differences of its own logL
    for j in range(len(truth)):
        h = 1e-5
        e = np.zeros_like(truth)
        e[j] = h
        fd = (like.logL(truth + e) - like.logL(truth - e)) / (2 * h)
        assert abs(g_truth[j] - fd) <= 1e-5 * np.abs(g_truth).max() + 1e-6, (j, g_truth[j], fd)
    # Gaussian priors add their value and their gradient
    like_p = LikelihoodModule(lin, param, prior_lens=[[0, "gamma", 2.05, 0.05]], prior_source=[[0, "n_sersic", 2.5, 0.5]])
    l_p, g_p = like_p.value_and_grad(truth)
    names = param.num_param()[1]
    jg, jn = names.index("gamma_lens0"), names.index("n_sersic_source_light0")
    dl = -(truth[jg] - 2.05) ** 2 / 0.05**2 / 2 - (truth[jn] - 2.5) ** 2 / 0.5**2 / 2
    assert abs(l_p - (l_truth + dl)) <= 1e-10 * abs(l_truth) and abs(like_p.logL(truth) - l_p) <= 1e-10 * abs(l_p)
    dg = np.zeros_like(truth)
    dg[jg], dg[jn] = -(truth[jg] - 2.05) / 0.05**2, -(truth[jn] - 2.5) / 0.5**2
    np.testing.assert_allclose(g_p, g_truth + dg, rtol=1e-12, atol=1e-12 * np.abs(g_truth).max())
    lo, hi = param.param_limits()
    sigma = 0.02 * np.maximum(np.abs(truth), 0.05)
    opt = OptaxMinimizer(like, truth, sigma, np.asarray(lo), np.asarray(hi), maxiter=200)
    best = opt.run(num_chains=4, tol=1e-7, rng_seed=0)
    l_best = like.logL(best)
    assert l_best >= l_truth - 1e-6, (l_best, l_truth)
    assert np.all(opt.final_logL >= l_truth - 0.5), opt.final_logL
    print("gradient descent: logL(truth) %.4f -> best %.4f in %s iterations (%d batched evaluations)"
          % (l_truth, l_best, opt.num_iterations, opt.num_evaluations))


def test_multi_plane_gradient_against_the_oracle_and_the_forward_path():
    """Multi-plane lensing (EPL z = 0.5, NFW_ELLIPSE_CSE z = 0.3, SIS z = 0.8; source z = 1.5 -- BASELINE config 5's
    lens, with the reference fixture's cosmological constants) inside ImageLinearFit on a 32 x 32 frame, ss 2, pixel
    PSF: the gradient of logL through k_image_grad<0, 2, true> (reverse sweep over the planes) against 4th-order central
    differences of (a) the NumPy oracle's multi-plane likelihood and (b) the device forward path."""
    import torch

    from jaxtronomy_b200.Data.imaging_data import ImageData
    from jaxtronomy_b200.Data.psf import PSF
    from jaxtronomy_b200.ImSim.image_linear_solve import ImageLinearFit
    from jaxtronomy_b200.LensModel.lens_model import LensModel
    from jaxtronomy_b200.LightModel.light_model import LightModel
    from oracle import imsim
    from oracle.lenstronomy_port import PSF as OPSF
    from tests.golden_io import load

    s = load("baseline_shapes.npz")["c5_static"]
    c = load("baseline_shapes.npz")["c5_00"]
    km = s.meta["kwargs_model"]
    n, dp = 32, 0.12
    half = dp * (n - 1) / 2.0
    kd = dict(exposure_time=100.0, background_rms=0.05, ra_at_xy_0=-half, dec_at_xy_0=-half,
              transform_pix2angle=np.array([[dp, 0.0], [0.0, dp]]))
    kernel = s.inp["kernel_0"]
    rng = np.random.default_rng(8)
    kl, ks, kll = c.meta["kwargs_lens"], c.meta["kwargs_source"], c.meta["kwargs_lens_light"]

    def make(image):
        lens = LensModel(km["lens_model_list"], z_source=km["z_source"], lens_redshift_list=km["lens_redshift_list"],
                         multi_plane=True)
        lens.set_multi_plane_constants(s.inp["T_ij_list"], s.inp["T_z_list"], s.inp["reduced2physical"],
                                       float(s.inp["T_ij_stop"]), float(s.inp["T_z_source"]))
        return ImageLinearFit(ImageData(image_data=image, **kd), PSF(psf_type="PIXEL", kernel_point_source=kernel), lens,
                              LightModel(km["source_light_model_list"]), LightModel(km["lens_light_model_list"]),
                              kwargs_numerics=dict(supersampling_factor=2))

    truth = make(np.zeros((n, n))).image(kl, [dict(ks[0], amp=15.0)], [dict(kll[0], amp=25.0)])
    image = truth + rng.normal(0, 0.05, truth.shape)
    lin = make(image)
    om = imsim.ImageModel(imsim.ImageData(image_data=image, **kd), OPSF(psf_type="PIXEL", kernel_point_source=kernel),
                          km["lens_model_list"], km["source_light_model_list"], km["lens_light_model_list"],
                          dict(supersampling_factor=2), multi_plane=s.inp)
    names = lin.param_names
    row = np.array([{"lens": kl, "source": ks, "lens_light": kll}[comp][i][nm] for comp, i, nm in names], dtype=float)
    row[[j for j, (_, _, nm) in enumerate(names) if nm == "theta_E"][0]] *= 1.03  # off the truth: a non-trivial gradient
    t = torch.as_tensor(row[None, :], device="cuda")
    logl, grad, amps = lin.logL_linear_grad_batch(t)
    logl_f, amps_f, _ = lin.logL_linear_batch(t)
    np.testing.assert_allclose(logl.cpu().numpy(), logl_f.cpu().numpy(), rtol=1e-12)
    grad = grad.cpu().numpy()[0]
    assert np.all(np.isfinite(grad))

    def oracle_logl(r):
        kws = {"lens": [dict() for _ in kl], "source": [dict() for _ in ks], "lens_light": [dict() for _ in kll]}
        for j, (comp, i, nm) in enumerate(names):
            kws[comp][i][nm] = r[j]
        return om.likelihood_data_given_model_linear(kws["lens"], kws["source"], kws["lens_light"])[0]

    def device_logl(r):
        return float(lin.logL_linear_batch(torch.as_tensor(r[None, :], device="cuda"))[0][0])

    assert abs(oracle_logl(row) - float(logl[0])) <= 1e-9 * abs(float(logl[0]))
    gscale = np.abs(grad).max()
    checked = 0
    for j, (comp, i, nm) in enumerate(names):
        if nm == "amp":
            assert grad[j] == 0.0  # solved linearly
            continue
        h = 1e-5 * max(1.0, abs(row[j]))
        fd_dev = _fd(device_logl, row, j, h)
        assert abs(grad[j] - fd_dev) <= 2e-6 * gscale, (comp, i, nm, grad[j], fd_dev)
        if comp == "lens":  # the oracle costs ~0.1 s per evaluation: the lens parameters are the ones the sweep touches
            fd_or = _fd(oracle_logl, row, j, h)
            assert abs(grad[j] - fd_or) <= 2e-6 * gscale, (comp, i, nm, grad[j], fd_or)
        checked += 1
    assert checked >= 20


@pytest.mark.parametrize("linear", [True, False])
def test_point_source_gradient(linear):
    """EPL + SHEAR, two Sersic components, a lensed quad and a field star (the reference fixture `image_quad_and_star`,
    40 x 40, ss 2, 9 x 9 pixel PSF): d logL / d(ra_image, dec_image[, point_amp]) from k_point_sources_grad and every
    other column from the extended-light path, against 4th-order central differences of the device forward path and, for
    the point-source columns, of the NumPy oracle.  The stamps are bilinear interpolations of the PSF kernel
    (point_source_rendering.py:33-77): logL is piecewise smooth in the positions, the step stays far inside a cell."""
    import torch

    from tests.golden_io import load
    from tests.test_gpu_parity import _ps_model
    from tests.test_oracle_golden import _oracle_ps_model

    c = load("point_sources.npz")["image_quad_and_star"]
    kl, ks, kll = c.meta["kwargs_lens"], c.meta["kwargs_source"], c.meta["kwargs_lens_light"]
    kps = [{k: np.asarray(v, dtype=float) for k, v in kw.items()} for kw in c.meta["kwargs_ps"]]
    lin = _ps_model(c, linear=True)
    lin.num_point_source_images = [4, 1]
    om = _oracle_ps_model(c)
    mat, _, _ = lin._pack(kl, ks, kll, kps, [4, 1])
    row = np.array(mat[0], dtype=float)
    names = lin.param_names
    t = torch.as_tensor(row[None, :], device="cuda")
    if linear:
        logl, grad, _ = lin.logL_linear_grad_batch(t)
        fwd = lambda r: float(lin.logL_linear_batch(torch.as_tensor(r[None, :], device="cuda"))[0][0])  # noqa: E731
    else:
        logl, grad = lin.logL_grad_batch(t)
        fwd = lambda r: float(lin.logL_batch(torch.as_tensor(r[None, :], device="cuda"))[0])  # noqa: E731
    assert abs(float(logl[0]) - fwd(row)) <= 1e-12 * abs(fwd(row))
    grad = grad.cpu().numpy()[0]
    assert np.all(np.isfinite(grad))

    def oracle(r):
        kws = {"lens": [dict() for _ in kl], "source": [dict() for _ in ks], "lens_light": [dict() for _ in kll]}
        ps = [dict(ra_image=np.zeros(4), dec_image=np.zeros(4), point_amp=np.zeros(4)),
              dict(ra_image=np.zeros(1), dec_image=np.zeros(1), point_amp=np.zeros(1))]
        for j, (comp, i, nm) in enumerate(names):
            if comp == "point_source":
                base, e = nm.split("[")
                ps[i][base][int(e[:-1])] = r[j]
            else:
                kws[comp][i][nm] = r[j]
        if linear:
            return om.likelihood_data_given_model_linear(kws["lens"], kws["source"], kws["lens_light"], kwargs_ps=ps)[0]
        return om.likelihood_data_given_model(kws["lens"], kws["source"], kws["lens_light"], kwargs_ps=ps)

    assert abs(oracle(row) - float(logl[0])) <= 1e-9 * abs(float(logl[0]))
    gscale = np.abs(grad).max()
    n_ps = 0
    for j, (comp, i, nm) in enumerate(names):
        if linear and (nm == "amp" or nm.startswith("point_amp")):
            assert grad[j] == 0.0  # solved linearly
            continue
        h = 1e-6 * max(1.0, abs(row[j]))
        fd = _fd(fwd, row, j, h)
        assert abs(grad[j] - fd) <= 5e-6 * gscale, (comp, i, nm, grad[j], fd)
        if comp == "point_source":
            fo = _fd(oracle, row, j, h)
            assert abs(grad[j] - fo) <= 5e-6 * gscale, (comp, i, nm, grad[j], fo)
            n_ps += 1
    assert n_ps == (10 if linear else 15)


def test_likelihood_module_gradient_with_point_sources_and_image_position_term():
    """LikelihoodModule.logL_and_grad_batch for the Einstein-ring notebook's model class (EPL + SHEAR, two Sersic
    components, a 4-image LENSED_POSITION point source) with the image-position likelihood on and the source-position
    penalty off: value == logL_batch, gradient == central differences of logL_batch; with the penalty on it raises."""
    from tests.golden_io import load
    from tests.test_gpu_parity import _ps_model
    from jaxtronomy_b200.Sampling.likelihood import LikelihoodModule
    from jaxtronomy_b200.Sampling.parameters import Param

    c = load("point_sources.npz")["image_quad_and_star"]
    kl, ks, kll = c.meta["kwargs_lens"], c.meta["kwargs_source"], c.meta["kwargs_lens_light"]
    kps = [{k: np.asarray(v, dtype=float) for k, v in kw.items()} for kw in c.meta["kwargs_ps"]]
    lin = _ps_model(c, linear=True)
    km = dict(lens_model_list=c.meta["lens_model_list"], source_light_model_list=c.meta["source_model_list"],
              lens_light_model_list=c.meta["lens_light_model_list"], point_source_model_list=c.meta["point_source_type_list"])
    param = Param(km, kwargs_fixed_lens=[{}, {"ra_0": 0.0, "dec_0": 0.0}], num_point_source_list=[4, 1])
    ra_list = [list(kps[0]["ra_image"] + 0.002), list(kps[1]["ra_image"])]
    dec_list = [list(kps[0]["dec_image"] - 0.001), list(kps[1]["dec_image"])]
    like = LikelihoodModule(lin, param, source_position_likelihood=False, image_position_likelihood=True,
                            image_position_uncertainty=0.004, ra_image_list=ra_list, dec_image_list=dec_list)
    x0 = param.kwargs2args(kl, ks, kll, kwargs_ps=kps)
    pos = np.stack([x0, x0 * (1 + 1e-3)])
    l, g = like.logL_and_grad_batch(pos)
    np.testing.assert_allclose(l, like.logL_batch(pos), rtol=1e-12)
    assert np.all(np.isfinite(g)) and g.shape == pos.shape
    gs = np.abs(g[0]).max()
    for j in range(len(x0)):
        h = 1e-6 * max(1.0, abs(x0[j]))
        e = np.zeros_like(x0)
        e[j] = h
        fd = (8 * (like.logL(x0 + e) - like.logL(x0 - e)) - (like.logL(x0 + 2 * e) - like.logL(x0 - 2 * e))) / (12 * h)
        assert abs(g[0, j] - fd) <= 5e-6 * gs, (j, param.num_param()[1][j], g[0, j], fd)
    with pytest.raises(ValueError):
        LikelihoodModule(lin, param).logL_and_grad_batch(pos)  # source-position penalty on (the default): value only


GRADS = load("gradients.npz")


def _gradient_fixture_model(c):
    """The product's ImageLinearFit for a tests/golden/gradients.npz case (built from the fixture alone)."""
    from jaxtronomy_b200.Data.imaging_data import ImageData
    from jaxtronomy_b200.Data.psf import PSF
    from jaxtronomy_b200.ImSim.image_linear_solve import ImageLinearFit
    from jaxtronomy_b200.LensModel.lens_model import LensModel
    from jaxtronomy_b200.LightModel.light_model import LightModel
    from jaxtronomy_b200.PointSource.point_source import PointSource

    lm = c.meta["lens_meta"]
    if lm.get("multi_plane"):
        lens = LensModel(lm["lens_model_list"], z_source=lm["z_source"], lens_redshift_list=lm["lens_redshift_list"],
                         multi_plane=True)
        lens.set_multi_plane_constants(c.inp["T_ij_list"], c.inp["T_z_list"], c.inp["reduced2physical"],
                                       float(c.inp["T_ij_stop"]), float(c.inp["T_z_source"]))
    else:
        lens = LensModel(lm["lens_model_list"])
    kd = dict(image_data=c.inp["image_data"], transform_pix2angle=c.inp["transform_pix2angle"],
              ra_at_xy_0=float(c.inp["ra_at_xy_0"]), dec_at_xy_0=float(c.inp["dec_at_xy_0"]),
              background_rms=float(c.inp["background_rms"]), exposure_time=float(c.inp["exposure_time"]))
    ps_types = c.meta["point_source_type_list"]
    lin = ImageLinearFit(ImageData(**kd), PSF(psf_type="PIXEL", kernel_point_source=c.inp["kernel"]), lens,
                         LightModel(c.meta["source_model_list"]), LightModel(c.meta["lens_light_model_list"]),
                         point_source_class=PointSource(ps_types, lens_model=lens) if ps_types else None,
                         kwargs_numerics=c.meta["kwargs_numerics"])
    if ps_types:
        lin.num_point_source_images = c.meta["num_point_source_images"]
    return lin


@pytest.mark.parametrize("linear", [True, False])
@pytest.mark.parametrize("name", sorted(GRADS))
def test_gradient_matches_reference_finite_differences(name, linear):
    """lsim_logl_grad against 4th-order central differences of the UNMODIFIED reference's own log-likelihood
    (tests/golden/gradients.npz, make_golden.py gen_gradients: ImageModel / ImageLinearFit under refshim), for a
    single-plane extended model, the same with a lensed quad + field star, and BASELINE config 5's multi-plane lens.
    Value: 1e-10.  Gradient: the closest of the steps 1e-4 / 1e-5 / 1e-6 within 1e-8 max|g| (a finite difference of an
    fp64 forward value carries ~1e-16 |logL| / h of round-off and h^4 of truncation; measured agreement 1e-11 .. 2e-11)."""
    import torch

    c = GRADS[name]
    lin = _gradient_fixture_model(c)
    fixed = c.meta["fixed_lens"]
    fx = {tuple(n): v for n, v in zip(c.meta["names"], c.inp["row"])}
    for i, kw in enumerate(fixed):
        for k, v in kw.items():
            fx[("lens", i, k)] = float(v)
    names = lin.param_names
    row = np.array([fx[(comp, i, n)] for comp, i, n in names], dtype=float)
    col = {tuple(n): j for j, n in enumerate(c.meta["names"])}
    t = torch.as_tensor(row[None, :], device="cuda")
    if linear:
        logl, grad, _ = lin.logL_linear_grad_batch(t)
        ref_l, fds = float(c.out["logL_linear"]), c.out["fd_linear"]
    else:
        logl, grad = lin.logL_grad_batch(t)
        ref_l, fds = float(c.out["logL"]), c.out["fd_nonlinear"]
    assert abs(float(logl[0]) - ref_l) <= 1e-10 * abs(ref_l), (float(logl[0]), ref_l)
    grad = grad.cpu().numpy()[0]
    assert np.all(np.isfinite(grad))
    gscale = np.abs(grad).max()
    checked, worst = 0, 0.0
    for j, (comp, i, n) in enumerate(names):
        key = (comp, i, n)
        if key not in col:  # SHEAR's ra_0 / dec_0: fixed in the fixture
            continue
        fd = fds[:, col[key]]
        if linear and (n == "amp" or n.startswith("point_amp")):
            assert grad[j] == 0.0 and np.all(np.isnan(fd))
            continue
        err = np.min(np.abs(grad[j] - fd)) / gscale
        worst = max(worst, err)
        assert err <= 1e-8, (name, key, grad[j], fd)
        checked += 1
    assert checked == len(col) - (sum(1 for k in col if k[2] == "amp" or k[2].startswith("point_amp")) if linear else 0)
    print("%s linear=%s: %d columns, worst |g - fd_ref| / max|g| = %.2e" % (name, linear, checked, worst))
