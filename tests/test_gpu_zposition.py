"""GPU parity of PositionLikelihood (lsim_source_position_logl, k_source_position) through the C ABI against fixtures
produced by the unmodified reference (tests/golden/position_likelihood.npz), the host-compiled kernel arithmetic, and
its use inside LikelihoodModule.  Tolerance: tests/test_hostmath_position.py::conditioning."""

import numpy as np
import pytest

from tests.test_hostmath_position import POS, SRC_CASES, _opts, conditioning

pytestmark = pytest.mark.gpu


def _position_likelihood(c, **extra):
    from jaxtronomy_b200.LensModel.lens_model import LensModel
    from jaxtronomy_b200.PointSource.point_source import PointSource
    from jaxtronomy_b200.Sampling.Likelihoods.position_likelihood import PositionLikelihood

    lens = LensModel(c.meta["lens_model_list"])
    ps = PointSource(c.meta["point_source_type_list"], lens_model=lens)
    return PositionLikelihood(ps, **dict(c.meta["options"], **extra))


@pytest.mark.parametrize("name", [n for n in POS if n != "astrometric"])
def test_position_likelihood_matches_the_reference(name):
    c = POS[name]
    pl = _position_likelihood(c)
    got = pl.logL(c.meta["kwargs_lens"], c.meta["kwargs_ps"], {})
    want = float(c.out["logL"])
    assert abs(got - want) <= 1e-10 * conditioning(c) * abs(want), (name, got, want)
    if name in SRC_CASES:
        sigma, hb = _opts(c)
        got = pl.source_position_likelihood(c.meta["kwargs_lens"], c.meta["kwargs_ps"], sigma, hard_bound_rms=hb)
        want = float(c.out["logL_source_position"])
        assert abs(got - want) <= 1e-10 * conditioning(c) * abs(want), (name, got, want)


@pytest.mark.parametrize("name", [n for n in POS if n != "astrometric"])
def test_point_source_source_position(name):
    """PointSource.source_position (point_source.py:188-202): mean of the ray-shot images / the positions themselves."""
    from jaxtronomy_b200.LensModel.lens_model import LensModel
    from jaxtronomy_b200.PointSource.point_source import PointSource

    c = POS[name]
    ps = PointSource(c.meta["point_source_type_list"], lens_model=LensModel(c.meta["lens_model_list"]))
    xs, ys = ps.source_position(c.meta["kwargs_ps"], c.meta["kwargs_lens"])
    for i in range(len(xs)):
        np.testing.assert_allclose(xs[i], c.out["x_source_%d" % i], rtol=1e-10, atol=1e-12)
        np.testing.assert_allclose(ys[i], c.out["y_source_%d" % i], rtol=1e-10, atol=1e-12)


def test_astrometric_likelihood_and_scope():
    from jaxtronomy_b200.Sampling.Likelihoods.position_likelihood import PositionLikelihood

    a = POS["astrometric"]
    quad = POS["quad"].meta["kwargs_ps"]
    got = PositionLikelihood.astrometric_likelihood(quad, a.meta["kwargs_special"], a.meta["sigma"])
    assert got == pytest.approx(float(a.out["logL"]), rel=1e-13)
    assert PositionLikelihood.astrometric_likelihood(quad, {}, 0.004) == 0
    with pytest.raises(ValueError):
        _position_likelihood(POS["quad"], force_no_add_image=True)
    with pytest.raises(ValueError):
        _position_likelihood(POS["quad"], restrict_image_number=True)
    assert _position_likelihood(POS["quad_image_positions"]).num_data == 8


def test_batched_sets_match_single_calls_and_the_host_build():
    """[B] lens parameters and [B, n_img] image positions: one launch; every row bit-equal to its single call and within
    1e-10 of the same C++ compiled for the host."""
    import torch

    from tests import hostmath_util as hm

    c = POS["quad"]
    pl = _position_likelihood(c)
    rng = np.random.default_rng(17)
    B = 6
    kw0 = c.meta["kwargs_lens"]
    th, g1 = kw0[0]["theta_E"] + rng.normal(0, 0.02, B), kw0[1]["gamma1"] + rng.normal(0, 0.01, B)
    ra = np.asarray(c.meta["kwargs_ps"][0]["ra_image"]) + rng.normal(0, 0.01, (B, 4))
    dec = np.asarray(c.meta["kwargs_ps"][0]["dec_image"]) + rng.normal(0, 0.01, (B, 4))
    kw_b = [dict(kw0[0], theta_E=th), dict(kw0[1], gamma1=g1)]
    got = pl.source_position_likelihood(kw_b, [dict(ra_image=ra, dec_image=dec)], 0.002, hard_bound_rms=0.3)
    assert got.shape == (B,)
    got_t = pl.source_position_likelihood(kw_b, [dict(ra_image=torch.as_tensor(ra, device="cuda"),
                                                      dec_image=torch.as_tensor(dec, device="cuda"))], 0.002,
                                          hard_bound_rms=0.3)
    assert isinstance(got_t, torch.Tensor) and got_t.is_cuda and np.array_equal(got_t.cpu().numpy(), got)
    for b in range(B):
        kw = [dict(kw0[0], theta_E=th[b]), dict(kw0[1], gamma1=g1[b])]
        one = pl.source_position_likelihood(kw, [dict(ra_image=ra[b], dec_image=dec[b])], 0.002, hard_bound_rms=0.3)
        assert one == got[b]
        host = hm.source_position_logl(c.meta["lens_model_list"], kw, ra[b], dec[b], 0.002, 0.3)
        # nvcc contracts more FMAs than the host build; the penalty's 1 / det(A)^2 amplifies that (measured 3e-12)
        assert abs(one - host) <= 1e-10 * abs(host)
    # shared image positions, batched lens parameters
    shared = pl.source_position_likelihood(kw_b, [dict(ra_image=ra[0], dec_image=dec[0])], 0.002)
    one = pl.source_position_likelihood([dict(kw0[0], theta_E=th[3]), dict(kw0[1], gamma1=g1[3])],
                                        [dict(ra_image=ra[0], dec_image=dec[0])], 0.002)
    assert shared.shape == (B,) and shared[3] == one


def test_c_abi_argument_checks():
    from jaxtronomy_b200 import _lib
    from jaxtronomy_b200.LensModel.lens_model import LensModel

    lm = LensModel(["SIS"])
    plan = lm._plan(None)
    p = np.array([[1.0, 0.0, 0.0]])
    x, y = np.array([1.1, -0.9]), np.array([0.0, 0.0])
    l = plan.source_position_logl(p, x, y, 0.001)
    assert abs(float(l[0])) < 1e-12          # exact images of a source at (0.1, 0): no spread, no penalty
    with pytest.raises(_lib.LsimError):
        plan.source_position_logl(p, x, y, 0.0)     # sigma must be positive
    mp = LensModel(["SIS"], z_source=1.5, lens_redshift_list=[0.5], multi_plane=True)
    with pytest.raises(_lib.LsimError):
        mp._plan(None).source_position_logl(p, x, y, 0.001)   # multi-plane: outside the scope
    assert float(plan.source_position_logl(p, x[:0], y[:0], 0.001)[0]) == 0.0


def test_likelihood_module_adds_the_position_terms():
    """LikelihoodModule.logL = imaging + position likelihood (likelihood.py:366-368) on the rows of `positions`: the
    image positions are columns of the matrix (model 0) or fixed constants (model 1)."""
    from jaxtronomy_b200.Sampling.Likelihoods.position_likelihood import PositionLikelihood
    from jaxtronomy_b200.Sampling.likelihood import LikelihoodModule
    from jaxtronomy_b200.Sampling.parameters import Param
    from tests.test_gpu_parity import POINTS, _ps_model

    c = POINTS["image_quad_and_star"]
    lin = _ps_model(c, linear=True)
    km = dict(lens_model_list=c.meta["lens_model_list"], source_light_model_list=c.meta["source_model_list"],
              lens_light_model_list=c.meta["lens_light_model_list"],
              point_source_model_list=c.meta["point_source_type_list"])
    kps = [{k: np.asarray(v, dtype=float) for k, v in kw.items()} for kw in c.meta["kwargs_ps"]]
    param = Param(km, kwargs_fixed_lens=[{}, {"ra_0": 0.0, "dec_0": 0.0}], num_point_source_list=[4, 1],
                  kwargs_fixed_ps=[{}, {"ra_image": kps[1]["ra_image"], "dec_image": kps[1]["dec_image"]}])
    opts = dict(source_position_sigma=0.002, source_position_tolerance=0.2, image_position_likelihood=True,
                image_position_uncertainty=0.004, ra_image_list=[kps[0]["ra_image"] + 0.001, kps[1]["ra_image"]],
                dec_image_list=[kps[0]["dec_image"] - 0.002, kps[1]["dec_image"] + 0.003])
    like_img = LikelihoodModule(lin, param, source_position_likelihood=False)
    like_all = LikelihoodModule(lin, param, **opts)
    like_default = LikelihoodModule(lin, param)          # source-position penalty on by default (LENSED_POSITION)
    pl = PositionLikelihood(lin.PointSource, **opts)
    pl_default = PositionLikelihood(lin.PointSource)
    x0 = param.kwargs2args(c.meta["kwargs_lens"], c.meta["kwargs_source"], c.meta["kwargs_lens_light"], kwargs_ps=kps)
    rng = np.random.default_rng(5)
    pos = x0[None, :] + rng.normal(0, 0.002, (5, x0.size))
    img, tot, dft = like_img.logL_batch(pos), like_all.logL_batch(pos), like_default.logL_batch(pos)
    for b in range(5):
        kw = param.args2kwargs(pos[b])
        want = img[b] + pl.logL(kw["kwargs_lens"], kw["kwargs_ps"], {})
        assert abs(tot[b] - want) <= 1e-12 * abs(want), (b, tot[b], want)
        want = img[b] + pl_default.logL(kw["kwargs_lens"], kw["kwargs_ps"], {})
        assert abs(dft[b] - want) <= 1e-12 * abs(want), (b, dft[b], want)
    with pytest.raises(ValueError):
        like_all.logL_and_grad_batch(pos)
    with pytest.raises(ValueError):
        LikelihoodModule(lin, param, astrometric_likelihood=True)
