"""CPU: host-side logic of the product (grid bookkeeping, kwargs packing, selection lists, sharding)."""

import os
import subprocess
import sys

import numpy as np
import pytest

from jaxtronomy_b200 import _model, parallel
from jaxtronomy_b200.Data.imaging_data import ImageData
from jaxtronomy_b200.Data.psf import PSF
from jaxtronomy_b200.ImSim.Numerics.grid import RegularGrid
from jaxtronomy_b200.ImSim.Numerics.numerics import Numerics, gaussian_pixel_kernel
from jaxtronomy_b200.Util import kernel_util, simulation_util, util
from tests.golden_io import load

IMSIM = load("imsim.npz")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("ss", [1, 2, 3])
def test_regular_grid_bit_exact(ss):
    """coordinates and index maps are array_equal to the reference's (test_grid.py)."""
    c = IMSIM["grid_ss%d" % ss]
    g = RegularGrid(6, 6, c.inp["transform_pix2angle"], float(c.inp["ra_at_xy_0"]), float(c.inp["dec_at_xy_0"]), ss)
    ra, dec = g.coordinates_evaluate
    np.testing.assert_array_equal(g._compute_indexes, c.out["compute_indexes"])
    np.testing.assert_array_equal(ra, c.out["ra"])
    np.testing.assert_array_equal(dec, c.out["dec"])
    assert g.num_grid_points_axes == (6 * ss, 6 * ss)


def test_regular_grid_mask_scatter():
    mask = np.zeros((4, 4))
    mask[1:3, 2] = 1
    g = RegularGrid(4, 4, np.eye(2) * 0.1, 0.0, 0.0, 2, flux_evaluate_indexes=mask)
    assert g._compute_indexes.sum() == 2 * 4
    flux = np.arange(8.0)
    full = g.flux_to_full_grid(flux)
    assert full.shape == (64,)
    np.testing.assert_array_equal(full[g._compute_indexes], flux)
    assert full.sum() == flux.sum()


@pytest.mark.parametrize("name", [n for n in IMSIM if n.startswith("numerics_")])
def test_numerics_coordinates_and_kernel(name):
    c = IMSIM[name]
    from jaxtronomy_b200.Data.pixel_grid import PixelGrid

    pg = PixelGrid(c.meta["nx"], c.meta["ny"], c.inp["transform_pix2angle"], float(c.inp["ra_at_xy_0"]),
                   float(c.inp["dec_at_xy_0"]))
    kw = dict(c.meta["kwargs_psf"])
    if kw["psf_type"] == "PIXEL":
        kw["kernel_point_source"] = c.inp["kernel"]
    psf = PSF(**kw)
    psf.set_pixel_size(pg.pixel_width)
    if c.meta.get("subframe"):
        from jaxtronomy_b200.ImSim.Numerics.numerics import NumericsSubFrame

        num = NumericsSubFrame(pg, psf, flux_evaluate_indexes=c.inp["flux_evaluate_indexes"],
                               **c.meta["kwargs_numerics"])
        assert num.subframe_calc and num.grid_dict()["nx"] < c.meta["nx"]
    else:
        num = Numerics(pg, psf, **c.meta["kwargs_numerics"])
    ra, dec = num.coordinates_evaluate
    np.testing.assert_array_equal(ra, c.out["ra"])
    np.testing.assert_array_equal(dec, c.out["dec"])
    if kw["psf_type"] == "GAUSSIAN":
        ss_conv = c.meta["kwargs_numerics"].get("supersampling_convolution", False)
        inner = num._numerics_subframe if c.meta.get("subframe") else num
        k2, k1 = (inner._kernel_hi, inner._kernel_hi_1d) if ss_conv else (inner._kernel, inner._kernel_1d)
        assert k2.shape[0] % 2 == 1 and abs(k2.sum() - 1) < 1e-14
        # the separable factorisation the kernels use reproduces the reference's 2-d kernel to rounding
        assert np.max(np.abs(np.outer(k1, k1) - k2)) < 1e-16


def test_gaussian_kernel_matches_reference_definition():
    k = gaussian_pixel_kernel(1.7, 9)
    x, y = util.make_grid(9, 1)
    ref = np.exp(-(x**2 + y**2) / (2 * 1.7**2)).reshape(9, 9)
    np.testing.assert_allclose(k, ref / ref.sum(), rtol=1e-14)
    with pytest.raises(ValueError):
        gaussian_pixel_kernel(1.0, 4)


def test_unsupported_numerics_raise():
    kd = simulation_util.data_configure_simple(10, 0.1, 100, 0.1)
    data = ImageData(**kd)
    psf = PSF(psf_type="PIXEL", kernel_point_source=np.ones((3, 3)))
    with pytest.raises(NotImplementedError):  # pixel PSF not given at the supersampled resolution
        Numerics(data, psf, supersampling_factor=3, supersampling_convolution=True)
    from jaxtronomy_b200.ImSim.Numerics.numerics import NumericsSubFrame

    rect = np.zeros((10, 10))
    rect[2:5, 1:8] = 1
    with pytest.raises(ValueError):  # rectangular sub-frame: the reference cannot place it back either
        NumericsSubFrame(data, psf, flux_evaluate_indexes=rect)
    with pytest.raises(ValueError):
        Numerics(data, psf, compute_mode="adaptive")
    with pytest.raises(TypeError):
        Numerics(data, psf, supersampling_factor=2.0)
    with pytest.raises(ValueError):
        PSF(psf_type="PIXEL", kernel_point_source=np.ones((4, 4)))
    with pytest.raises(ValueError):
        PSF(psf_type="FOO")


def test_image_data_noise_model():
    img = np.array([[1.0, -2.0], [0.5, 4.0]])
    d = ImageData(img, exposure_time=100.0, background_rms=0.1, transform_pix2angle=np.eye(2) * 0.05)
    np.testing.assert_allclose(d.C_D, np.maximum(img, 0) / 100.0 + 0.01)
    assert d.num_pixel_axes == (2, 2) and abs(d.pixel_width - 0.05) < 1e-15
    d2 = ImageData(img, noise_map=np.full((2, 2), 0.3))
    np.testing.assert_allclose(d2.C_D, 0.09)
    with pytest.raises(ValueError):
        ImageData(img, background_rms=0.1)
    d3 = ImageData(img, exposure_time=np.array([[0.0, 1.0], [1.0, 1.0]]), background_rms=0.1)
    assert d3.exposure_map[0, 0] == 1e-10


def test_convert_bool_list():
    assert util.convert_bool_list(3, None) == [True, True, True]
    assert util.convert_bool_list(3, 1) == [False, True, False]
    assert util.convert_bool_list(4, [0, 3]) == [True, False, False, True]
    assert util.convert_bool_list(2, [True, False]) == [True, False]
    with pytest.raises(ValueError):
        util.convert_bool_list(2, [True, False, True])


def test_kwargs_packing():
    names = [_model.LENS_PARAM_NAMES["SIS"], _model.LENS_PARAM_NAMES["SHEAR"]]
    kw = [dict(theta_E=np.array([1.0, 1.1, 1.2]), center_x=0.1), dict(gamma1=0.01, gamma2=np.array([0.0, 0.1, 0.2]))]
    values = _model.gather_values(names, kw)
    mat, B, batched = _model.pack_columns(values)
    assert mat.shape == (3, 7) and B == 3 and batched
    np.testing.assert_array_equal(mat[:, 0], [1.0, 1.1, 1.2])
    np.testing.assert_array_equal(mat[:, 1], 0.1)      # broadcast scalar
    np.testing.assert_array_equal(mat[:, 2], 0.0)      # default center_y
    np.testing.assert_array_equal(mat[:, 4], [0.0, 0.1, 0.2])
    mat1, B1, batched1 = _model.pack_columns(_model.gather_values(names, [dict(theta_E=1.0), dict(gamma1=0.0, gamma2=0.0)]))
    assert mat1.shape == (1, 7) and not batched1
    with pytest.raises(ValueError):
        _model.gather_values(names, [dict(theta_E=1.0, foo=2.0), dict(gamma1=0, gamma2=0)])
    with pytest.raises(ValueError):
        _model.gather_values(names, [dict(center_x=1.0), dict(gamma1=0, gamma2=0)])
    with pytest.raises(ValueError):
        _model.gather_values(names, [dict(theta_E=1.0)])
    with pytest.raises(ValueError):
        _model.pack_columns([np.zeros(3), np.zeros(4)])


def test_lens_desc_options():
    from jaxtronomy_b200 import _lib

    d = _model.lens_desc("SIE", [("cst", 1.0)] * 5, dict(NIE=False, gamma=2.1))
    assert d.flags & _lib.FLAG_SIE_NIE_FALSE and d.opt[1] == 2.1
    d = _model.lens_desc("CSE", [("col", k) for k in range(6)], dict(axis="major"))
    assert d.flags & _lib.FLAG_CSE_MAJOR and list(d.col[:6]) == list(range(6))
    with pytest.raises(ValueError):
        _model.lens_desc("CSE", [("cst", 0.0)] * 6, dict(axis="minor"))
    with pytest.raises(ValueError):
        _model.lens_desc("FOO", [])
    with pytest.raises(Exception):
        _model.lens_desc("NFW", [("cst", 1.0)] * 4, dict(interpol=True))


def test_cut_psf():
    k = np.arange(49.0).reshape(7, 7)
    c = kernel_util.cut_psf(k, 3, normalisation=False)
    np.testing.assert_array_equal(c, k[2:5, 2:5])
    assert abs(kernel_util.cut_psf(k, 5).sum() - 1) < 1e-15
    with pytest.raises(ValueError):
        kernel_util.cut_psf(k, 4)


def test_shard_bounds():
    for n, w in ((65536, 8), (10, 3), (5, 8), (0, 2)):
        got = [parallel.shard_bounds(n, w, r) for r in range(w)]
        assert got[0][0] == 0 and got[-1][1] == n
        for (a, b), (c2, d) in zip(got[:-1], got[1:]):
            assert b == c2 and b >= a
        sizes = parallel.shard_sizes(n, w)
        assert sum(sizes) == n and max(sizes) - min(sizes) <= 1


@pytest.mark.parametrize("n", [16, 11])
def test_all_gather_logl_gloo_world2(n):
    """N > 1 path on CPU: two gloo ranks shard a batch, evaluate a toy logL, all-gather it."""
    script = os.path.join(ROOT, "tests", "gloo_worker.py")
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT=str(29500 + n), PYTHONPATH=ROOT)
    procs = [subprocess.Popen([sys.executable, script, str(r), "2", str(n)], env=env, stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=120)[0] for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o
        assert "OK" in o, o


def test_param_mapping_round_trip():
    from jaxtronomy_b200.Sampling.parameters import Param

    km = dict(lens_model_list=["EPL", "SHEAR"], source_light_model_list=["SERSIC_ELLIPSE"],
              lens_light_model_list=["SERSIC"])
    p = Param(km, kwargs_fixed_lens=[{}, {"ra_0": 0, "dec_0": 0}], kwargs_fixed_lens_light=[{"n_sersic": 4.0}],
              kwargs_lower_lens=[{"theta_E": 0.5}, {}], kwargs_upper_lens=[{"theta_E": 2.0}, {}])
    n, names = p.num_param()
    assert n == 6 + 2 + 6 + 3 and names[0] == "theta_E_lens0" and names[-1] == "center_y_lens_light0"
    args = np.arange(n, dtype=float)
    kw = p.args2kwargs(args)
    assert kw["kwargs_lens"][1] == {"gamma1": 6.0, "gamma2": 7.0, "ra_0": 0, "dec_0": 0}
    assert kw["kwargs_source"][0]["amp"] == 1 and kw["kwargs_lens_light"][0]["n_sersic"] == 4.0
    np.testing.assert_array_equal(p.kwargs2args(kw["kwargs_lens"], kw["kwargs_source"], kw["kwargs_lens_light"]), args)
    lo, hi = p.param_limits()
    assert lo[0] == 0.5 and hi[0] == 2.0 and lo[1] == 1.5 and hi[1] == 2.5
    # batched: [B, n] matrix -> [B] arrays per parameter
    kwb = p.args2kwargs(np.stack([args, args + 1]))
    np.testing.assert_array_equal(kwb["kwargs_lens"][0]["gamma"], [1.0, 2.0])
    fn = p.column_source_fn()
    assert fn("lens", 1, ["gamma1", "gamma2", "ra_0", "dec_0"]) == [("col", 6), ("col", 7), ("cst", 0.0), ("cst", 0.0)]
    assert fn("source", 0, ["amp", "R_sersic"]) == [("cst", 1.0), ("col", 8)]
    # band sub-model: local profile 0 of the band is global profile 1
    fn2 = p.column_source_fn({"lens": [1], "source": None, "lens_light": None})
    assert fn2("lens", 0, ["gamma1", "gamma2", "ra_0", "dec_0"])[0] == ("col", 6)
    with pytest.raises(ValueError):
        Param(km, kwargs_fixed_lens=[{}])
    nolin = Param(km, linear_solver=False)
    assert nolin.num_param()[0] == 6 + 4 + 7 + 5


def test_pso_runs_on_a_toy_likelihood():
    from jaxtronomy_b200.Sampling.sampler import ParticleSwarmOptimizer, prepare_logL_func

    pso = ParticleSwarmOptimizer(lambda x: -np.sum((x - 0.3) ** 2, axis=1), [-1, -1], [1, 1], particle_count=32, seed=1)
    best, hist, _ = pso.optimize(max_iter=60)
    assert hist[-1] >= hist[0] and np.allclose(best, 0.3, atol=1e-2)
    with pytest.raises(ValueError):
        prepare_logL_func("cpu", None)
    with pytest.raises(ValueError):
        prepare_logL_func("tpu", None)


def test_oracle_pso_step_bookkeeping():
    """oracle/pso.py: first-wins ties, nan never wins, non-improving particles keep their personal best."""
    from oracle import pso as opso

    B, P = 6, 2
    pos = np.arange(B * P, dtype=float).reshape(B, P)
    vel = np.zeros((B, P))
    pb, pf = pos.copy() + 100, np.full(B, -10.0)
    gb, gf = np.zeros(P), np.array([-5.0])
    fit = np.array([-20.0, -1.0, np.nan, -1.0, -9.0, -30.0])
    rnd = np.zeros((3, B, P))
    before = pos.copy()
    opso.pso_step(pos, vel, pb, pf, gb, gf, fit, rnd)
    assert gf[0] == -1.0 and np.array_equal(gb, before[1])
    assert np.array_equal(pf, [-10.0, -1.0, -10.0, -1.0, -9.0, -10.0])
    assert np.array_equal(pb[1], before[1]) and np.array_equal(pb[0], before[0] + 100)
    assert np.array_equal(vel, np.zeros((B, P))) and np.array_equal(pos, before)  # all draws 0: w v = 0, no pull


def test_gaussian_psf_point_source_kernel():
    """Pixelised kernel of a GAUSSIAN PSF for point sources: odd size from truncation * fwhm / pixel size, unit sum,
    second moment = sigma^2 (in pixels) up to the truncation, supersampled version consistent."""
    from jaxtronomy_b200.Data.psf import PSF
    from jaxtronomy_b200.Util.util import fwhm2sigma

    psf = PSF(psf_type="GAUSSIAN", fwhm=0.3, truncation=6, pixel_size=0.05)
    k = psf.kernel_point_source
    assert k.shape == (37, 37) and abs(k.sum() - 1) < 1e-14 and np.allclose(k, k.T) and np.allclose(k, k[::-1, ::-1])
    a = np.arange(37) - 18
    var = (k.sum(0) * a**2).sum()
    assert abs(var - (fwhm2sigma(0.3) / 0.05) ** 2) < 2e-3 * var
    k3 = psf.kernel_point_source_supersampled(3)
    assert k3.shape[0] % 2 == 1 and abs(k3.sum() - 1) < 1e-14 and abs(k3.shape[0] - 3 * 36) <= 1


def test_bench_reference_arm_prints_one_json_line():
    """`bench.py --impl reference` (the CPU arm the driver runs beside the GPU arm) on a tiny sample: one JSON line
    with the contract's keys, no GPU needed."""
    import json
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0", "--cpu-images", "4"], capture_output=True, text=True, timeout=300, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "images/s" and d["value"] > 0 and d["higher_is_better"] is True
    assert d["cpu_baseline"]["kind"] == "port" and d["e2e"]["h2d_bytes_per_step"] == 0 and d["gpu_launches"] == 0
    # under torchrun only rank 0 works
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1"],
                         capture_output=True, text=True, timeout=60, cwd=root, env=dict(os.environ, RANK="1"))
    assert out.returncode == 0 and out.stdout.strip() == ""


# ---------------------------------------------------------------------------------------------------------------
# gradient-descent sampler (host logic of Sampling/Samplers/optax.py; the device gradient is tested in test_gpu_grad)
# ---------------------------------------------------------------------------------------------------------------
def _rosenbrock_logl(x):
    f = np.sum(100 * (x[:, 1:] - x[:, :-1] ** 2) ** 2 + (1 - x[:, :-1]) ** 2, axis=1)
    g = np.zeros_like(x)
    g[:, :-1] += -400 * x[:, :-1] * (x[:, 1:] - x[:, :-1] ** 2) - 2 * (1 - x[:, :-1])
    g[:, 1:] += 200 * (x[:, 1:] - x[:, :-1] ** 2)
    return -f, -g


def test_optax_minimizer_batched_lbfgs_converges_inside_the_bounds():
    from jaxtronomy_b200.Sampling.Samplers.optax import OptaxMinimizer

    n = 5
    opt = OptaxMinimizer(_rosenbrock_logl, np.zeros(n), np.ones(n), -3 * np.ones(n), 3 * np.ones(n), maxiter=500)
    best = opt.run(num_chains=6, tol=1e-12, rng_seed=0)
    np.testing.assert_allclose(best, np.ones(n), atol=1e-5)
    assert np.all(opt.num_iterations <= 500) and opt.final_logL.shape == (6,)
    # the chains are advanced together: far fewer batched evaluations than iterations x chains
    assert opt.num_evaluations < opt.num_iterations.sum()
    # same seed, same result (NumPy generator, deterministic line search)
    opt2 = OptaxMinimizer(_rosenbrock_logl, np.zeros(n), np.ones(n), -3 * np.ones(n), 3 * np.ones(n), maxiter=500)
    np.testing.assert_array_equal(opt2.run(num_chains=6, tol=1e-12, rng_seed=0), best)


def test_optax_minimizer_reparametrisation_and_stop_rule():
    from jaxtronomy_b200.Sampling.Samplers.optax import OptaxMinimizer

    lo, hi = np.array([-1.0, 0.0, 2.0]), np.array([1.0, 10.0, 2.5])
    opt = OptaxMinimizer(lambda x: (-np.sum((x - 0.3) ** 2, axis=1), -2 * (x - 0.3)), np.zeros(3), np.ones(3), lo, hi,
                         maxiter=7)
    x = np.array([[0.2, 3.0, 2.1], [-0.99, 9.99, 2.49]])
    np.testing.assert_allclose(opt.constrain(opt.unconstrain(x)), x, rtol=1e-12)
    # every unconstrained point maps inside the box, however far out (numpyro's sigmoid-affine interval transform)
    far = opt.constrain(np.array([[-800.0, 800.0, 0.0]]))
    assert np.all(far >= lo) and np.all(far <= hi) and far[0, 2] == 2.25
    # start points come from the truncated normal: inside the box, reproducible
    init = opt._draw_init_params(100, 3)
    assert np.all(init > lo) and np.all(init < hi)
    np.testing.assert_array_equal(init, opt._draw_init_params(100, 3))
    # the optimum of the third parameter lies outside its box: the chain walks to the lower edge; maxiter caps the loop
    u, iters = opt.run_single(opt.unconstrain(np.array([0.0, 5.0, 2.2])), tol=1e-14)
    assert iters <= 7
    # loss gradient in the unconstrained space = chain rule through the sigmoid
    f0, g0 = opt._loss_and_grad(u[None, :])
    e = 1e-6
    for j in range(3):
        d = np.zeros(3)
        d[j] = e
        fd = (opt._loss(u + d)[0] - opt._loss(u - d)[0]) / (2 * e)
        assert abs(g0[0, j] - fd) < 1e-6 * max(1.0, abs(fd))
    with pytest.raises(ValueError):
        OptaxMinimizer(_rosenbrock_logl, np.zeros(2), np.ones(2), np.ones(2), np.ones(2))


def test_likelihood_module_gaussian_priors_value_and_gradient():
    """Prior term of LikelihoodModule on the columns of the positions matrix (device-agnostic tensor code)."""
    import torch

    from jaxtronomy_b200.Sampling.likelihood import LikelihoodModule
    from jaxtronomy_b200.Sampling.parameters import Param

    class NoImage:  # the imaging likelihood needs a GPU; the prior bookkeeping does not
        def sampler_plans(self, *a, **k):
            return []

    km = dict(lens_model_list=["EPL", "SHEAR"], source_light_model_list=["SERSIC_ELLIPSE"])
    param = Param(km, kwargs_fixed_lens=[{"gamma": 2.0}, {"ra_0": 0, "dec_0": 0}])
    like = LikelihoodModule(NoImage(), param, prior_lens=[[0, "theta_E", 1.2, 0.1], [0, "gamma", 2.1, 0.2],
                                                          [1, "gamma1", 0.0, 0.05]],
                            prior_source=[[0, "R_sersic", 0.3, 0.1]])
    n, names = param.num_param()
    rng = np.random.default_rng(0)
    pos = torch.as_tensor(rng.normal(1.0, 0.3, (5, n)))
    val, grad = like._prior(pos, want_grad=True)
    j = {nm: i for i, nm in enumerate(names)}
    x = pos.numpy()
    want = -((x[:, j["theta_E_lens0"]] - 1.2) ** 2 / 0.1**2 + (x[:, j["gamma1_lens1"]]) ** 2 / 0.05**2
             + (x[:, j["R_sersic_source_light0"]] - 0.3) ** 2 / 0.1**2) / 2 - (2.0 - 2.1) ** 2 / 0.2**2 / 2
    np.testing.assert_allclose(val.numpy(), want, rtol=1e-13)
    g = np.zeros_like(x)
    g[:, j["theta_E_lens0"]] = -(x[:, j["theta_E_lens0"]] - 1.2) / 0.1**2
    g[:, j["gamma1_lens1"]] = -x[:, j["gamma1_lens1"]] / 0.05**2
    g[:, j["R_sersic_source_light0"]] = -(x[:, j["R_sersic_source_light0"]] - 0.3) / 0.1**2
    np.testing.assert_allclose(grad.numpy(), g, rtol=1e-13)
    np.testing.assert_array_equal(like._prior(pos).numpy(), val.numpy())
    with pytest.raises(ValueError):
        LikelihoodModule(NoImage(), param, prior_lens=[[0, "not_a_parameter", 0.0, 1.0]])
    with pytest.raises(ValueError):
        LikelihoodModule(NoImage(), param, time_delay_likelihood=True)


def test_likelihood_module_image_position_term_on_the_columns():
    """Position-likelihood wiring of LikelihoodModule that needs no GPU: the image positions are gathered from the
    columns of `positions` (sampled model) or from fixed constants, and the image-position term
    (position_likelihood.py:214-247) is formed on them; checked against the oracle row by row.  The source-position
    term needs the kernel (tests/test_gpu_zposition.py)."""
    import torch

    from jaxtronomy_b200.LensModel.lens_model import LensModel
    from jaxtronomy_b200.PointSource.point_source import PointSource
    from jaxtronomy_b200.Sampling.likelihood import LikelihoodModule
    from jaxtronomy_b200.Sampling.parameters import Param
    from oracle import position_likelihood as opl

    lens = LensModel(["SIS"])

    class NoImage:
        def sampler_plans(self, *a, **k):
            return []

    img = NoImage()
    img.PointSource = PointSource(["LENSED_POSITION", "UNLENSED"], lens_model=lens)
    km = dict(lens_model_list=["SIS"], point_source_model_list=["LENSED_POSITION", "UNLENSED"])
    star = dict(ra_image=np.array([-1.21]), dec_image=np.array([1.07]))
    param = Param(km, num_point_source_list=[3, 1], kwargs_fixed_ps=[{}, star])
    ra_data = [[1.0, -0.7, 0.4, 9.9], [-1.2]]       # one measured position more than the model has images: truncated
    dec_data = [[0.4, -0.8, -1.0, 9.9], [1.1]]
    like = LikelihoodModule(img, param, source_position_likelihood=False, image_position_likelihood=True,
                            image_position_uncertainty=0.01, ra_image_list=ra_data, dec_image_list=dec_data)
    n = param.num_param()[0]
    rng = np.random.default_rng(2)
    pos = rng.normal(0.5, 0.4, (6, n))
    got = like._position_logl(torch.as_tensor(pos)).numpy()
    for b in range(6):
        kw = param.args2kwargs(pos[b])
        want = opl.image_position_likelihood(kw["kwargs_ps"], ra_data, dec_data, 0.01)
        assert abs(got[b] - want) <= 1e-12 * abs(want)
    # the kwargs-level class gives the same number, and counts the data like the reference (num_data, :285-296)
    pl = like._position_likelihood
    kw = param.args2kwargs(pos[0])
    assert abs(pl.image_position_likelihood(kw["kwargs_ps"], None, 0.01) - got[0]) <= 1e-12 * abs(got[0])
    assert pl.num_data == 10
    # no position term at all -> None; the reference's unsupported switches raise
    assert LikelihoodModule(img, param, source_position_likelihood=False)._position_logl(torch.as_tensor(pos)) is None
    with pytest.raises(ValueError):
        LikelihoodModule(img, param, source_position_likelihood=False, force_no_add_image=True)
    with pytest.raises(ValueError):
        LikelihoodModule(img, param, source_position_likelihood=False, astrometric_likelihood=True)
    with pytest.raises(ValueError):
        LikelihoodModule(NoImage(), Param(dict(lens_model_list=["SIS"])), image_position_likelihood=True)
    with pytest.raises(ValueError):
        LikelihoodModule(img, param, source_position_likelihood=False, image_position_likelihood=True)  # no data lists


def test_device_ensemble_sampler_samples_a_correlated_gaussian():
    """The stretch-move sampler (the emcee call behind Sampler.mcmc_emcee) on CPU tensors: a 3-d correlated Gaussian."""
    import torch

    from jaxtronomy_b200.Sampling.sampler import DeviceEnsembleSampler, sample_ball_truncated

    mu = torch.tensor([0.5, -1.0, 2.0], dtype=torch.float64)
    cov = torch.tensor([[1.0, 0.6, 0.0], [0.6, 2.0, -0.5], [0.0, -0.5, 0.5]], dtype=torch.float64)
    icov = torch.linalg.inv(cov)
    calls = []

    def logp(x):
        calls.append(x.shape[0])
        d = x - mu
        return -0.5 * torch.einsum("bi,ij,bj->b", d, icov, d)

    W = 64
    init = sample_ball_truncated(mu.numpy(), np.ones(3), -10 * np.ones(3), 10 * np.ones(3), size=W, seed=1)
    assert init.shape == (W, 3) and np.all(np.abs(init) < 10)
    s = DeviceEnsembleSampler(logp, W, 3, seed=2, device="cpu")
    s.run_mcmc(init, 1500)
    assert calls[0] == W and set(calls[1:]) == {W // 2} and len(calls) == 1 + 2 * 1500  # two half-ensembles per step
    flat = s.get_chain(discard=300, flat=True)
    assert flat.shape == (1200 * W, 3) and s.get_log_prob(discard=300, flat=True).shape == (1200 * W,)
    np.testing.assert_allclose(flat.mean(axis=0), mu.numpy(), atol=0.06)
    np.testing.assert_allclose(np.cov(flat.T), cov.numpy(), atol=0.12)
    acc = s.acceptance_fraction
    assert acc.shape == (W,) and 0.3 < acc.mean() < 0.9
    # the stored log-probabilities belong to the stored positions
    last = torch.as_tensor(s.get_chain()[-1])
    np.testing.assert_allclose(s.get_log_prob()[-1], logp(last).numpy(), rtol=1e-12)
    # deterministic for a given seed
    s2 = DeviceEnsembleSampler(logp, W, 3, seed=2, device="cpu")
    s2.run_mcmc(init, 20)
    np.testing.assert_array_equal(s2.get_chain(), s.get_chain()[:20])
    with pytest.raises(ValueError):
        DeviceEnsembleSampler(logp, 7, 3, device="cpu")
