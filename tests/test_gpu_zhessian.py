"""GPU parity of LensModel.hessian / kappa / gamma / curl / magnification (lsim_hessian, k_hessian) through the C ABI,
against fixtures produced by the unmodified reference (tests/golden/hessians.npz) and the oracle.  Tolerances and the
two reference facts behind them (NIE forward difference, core clamps): tests/hessian_cases.py."""

import numpy as np
import pytest

from tests.hessian_cases import FLEXION_CASES, HESSIANS, KEYS, MODEL_CASES, PROFILE_CASES, check_profile, keep_mask, profile_case

pytestmark = pytest.mark.gpu


def _LensModel():
    from jaxtronomy_b200.LensModel.lens_model import LensModel

    return LensModel


def _k(c):
    k = c.meta["k"]
    if isinstance(k, list) and not isinstance(k[0], bool):
        k = tuple(k)
    return k


@pytest.mark.parametrize("name", PROFILE_CASES)
def test_profile_hessian(name):
    c, prof, kw, pk = profile_case(name)
    lm = _LensModel()([prof], profile_kwargs_list=[pk])
    got = lm.hessian(c.inp["x"], c.inp["y"], [kw])
    check_profile(name, got, c, prof, kw)


def test_profile_hessian_matches_the_host_compiled_kernel_code():
    """Device against the same C++ compiled for the host: 1e-12 of the scale (FMA contraction only)."""
    from tests import hostmath_util as hm

    for name in ("epl_03", "nie_1", "tnfw_1", "epl_m1m3m4_0", "nfw_ellipse_cse_1"):
        c, prof, kw, pk = profile_case(name)
        x, y = c.inp["x"], c.inp["y"]
        keep = keep_mask(x, y, [kw])
        got = _LensModel()([prof], profile_kwargs_list=[pk]).hessian(x, y, [kw])
        want = hm.lens_hessian([prof], [kw], x, y, profile_kwargs_list=[pk])
        scale = max(np.max(np.abs(w[keep])) for w in want)
        for g, w, k in zip(got, want, KEYS):
            assert np.max(np.abs(g - w)[keep]) <= 1e-12 * scale, (name, k)


@pytest.mark.parametrize("name", MODEL_CASES)
def test_model_hessian_kappa_gamma_magnification(name):
    c = HESSIANS[name]
    names, kw, k = c.meta["lens_model_list"], c.meta["kwargs"], _k(c)
    x, y = c.inp["x"], c.inp["y"]
    sel = [kw[k]] if isinstance(k, int) else kw
    keep = keep_mask(x, y, sel)
    lm = _LensModel()(names)
    got = dict(zip(KEYS, lm.hessian(x, y, kw, k=k)))
    got["kappa"] = lm.kappa(x, y, kw, k=k)
    got["gamma1"], got["gamma2"] = lm.gamma(x, y, kw, k=k)
    got["curl"] = lm.curl(x, y, kw, k=k)
    got["magnification"] = lm.magnification(x, y, kw, k=k)
    scale = max(np.max(np.abs(c.out[q][keep])) for q in KEYS)
    for q in KEYS + ("kappa", "gamma1", "gamma2", "curl"):
        assert np.max(np.abs(got[q] - c.out[q])[keep]) <= 1e-10 * scale, (name, q)
    # mu = 1/det A: relative to |mu| times the conditioning of the determinant (|mu| scale / |det|)
    mu = c.out["magnification"]
    assert np.all(np.abs(got["magnification"] - mu)[keep] <= 1e-9 * np.abs(mu[keep]) * np.maximum(1.0, np.abs(mu[keep])))


@pytest.mark.parametrize("method", ["square", "cross"])
def test_differential_hessian(method):
    """diff != None: the reference differences alpha itself (lens_model.py:617-700); same stencil here.  The stencil
    divides rounding errors of alpha by diff = 1e-4, so the bar is 1e-10 |alpha| / diff."""
    c = HESSIANS["model_diff_%s" % method]
    names, kw = c.meta["lens_model_list"], c.meta["kwargs"]
    x, y = c.inp["x"], c.inp["y"]
    keep = keep_mask(x, y, kw)
    lm = _LensModel()(names)
    got = lm.hessian(x, y, kw, diff=c.meta["diff"], diff_method=method)
    ax, ay = lm.alpha(x, y, kw)
    tol = 1e-10 * max(np.max(np.abs(ax[keep])), np.max(np.abs(ay[keep]))) / c.meta["diff"]
    for g, q in zip(got, KEYS):
        assert np.max(np.abs(g - c.out[q])[keep]) <= tol, (method, q)
    with pytest.raises(ValueError):
        lm.hessian(x, y, kw, diff=1e-4, diff_method="circle")


@pytest.mark.parametrize("name", FLEXION_CASES)
def test_flexion(name):
    """lens_model.py:529-570: central differences of the Hessian over diff = 1e-4 (and of alpha over diff / 4 with
    hessian_diff).  The stencil divides the Hessian's rounding by diff (and alpha's by diff^2 / 4 with hessian_diff):
    the bar is 1e-10 of the Hessian scale / diff, resp. 1e-10 |alpha| / (diff^2 / 4)."""
    c = HESSIANS[name]
    names, kw, diff = c.meta["lens_model_list"], c.meta["kwargs"], c.meta["diff"]
    x, y = c.inp["x"], c.inp["y"]
    keep = keep_mask(x, y, kw)
    lm = _LensModel()(names)
    got = lm.flexion(x, y, kw, diff=diff, hessian_diff=c.meta["hessian_diff"])
    h = lm.hessian(x, y, kw)
    ax, ay = lm.alpha(x, y, kw)
    hs = max(np.max(np.abs(a[keep])) for a in h)
    tol = 1e-10 * (max(np.max(np.abs(ax[keep])), np.max(np.abs(ay[keep]))) / (diff * diff / 4)
                   if c.meta["hessian_diff"] else hs / diff)
    for g, q in zip(got, ("f_xxx", "f_xxy", "f_xyy", "f_yyy")):
        assert np.max(np.abs(g - c.out[q])[keep]) <= tol, (name, q, np.max(np.abs(g - c.out[q])[keep]), tol)


def test_batched_parameter_sets_torch_io_and_oracle_difference():
    """[B] parameter arrays = one launch over B sets; 2-d coordinates keep their shape; torch in -> torch out; every set
    against the oracle's 4th-order difference of alpha (1e-8 of the scale)."""
    import torch

    from oracle import lens_model as olm

    rng = np.random.default_rng(3)
    x, y = rng.uniform(0.4, 2.5, (6, 11)) * rng.choice([-1, 1], (6, 11)), rng.uniform(0.4, 2.5, (6, 11))
    B = 7
    th, g1, gam = rng.uniform(0.8, 1.5, B), rng.uniform(-0.05, 0.05, B), rng.uniform(1.8, 2.2, B)
    names = ["EPL_MULTIPOLE_M1M3M4", "SHEAR"]
    kw = [dict(theta_E=th, gamma=gam, e1=0.1, e2=-0.2, center_x=0.0, center_y=0.05, a1_a=0.01, delta_phi_m1=0.3,
               a3_a=0.005, delta_phi_m3=-0.2, a4_a=0.01, delta_phi_m4=0.1),
          dict(gamma1=g1, gamma2=0.02, ra_0=0.0, dec_0=0.0)]
    lm = _LensModel()(names)
    got = lm.hessian(torch.as_tensor(x, device="cuda"), torch.as_tensor(y, device="cuda"), kw)
    assert all(isinstance(g, torch.Tensor) and g.is_cuda and tuple(g.shape) == (B, 6, 11) for g in got)
    for b in range(B):
        kw_b = [dict(kw[0], theta_E=th[b], gamma=gam[b]), dict(kw[1], gamma1=g1[b])]
        want = olm.hessian_differential(names, x, y, kw_b, h=1e-3)
        one = lm.hessian(x, y, kw_b)
        scale = max(np.max(np.abs(w)) for w in want)
        for g, o, w in zip(got, one, want):
            assert np.array_equal(g[b].cpu().numpy(), o)          # batch row == single call, bit for bit
            assert np.max(np.abs(o - w)) <= 1e-8 * scale


def test_multi_plane_alpha_and_hessian():
    """Multi-plane: alpha = theta - beta (multi_plane.py:512-534) at 1e-10 of |theta|; the Hessian is the reference's
    forward difference of alpha with diff = 1e-8 (multi_plane.py:537-589), whose rounding noise is eps |theta| / diff on
    BOTH sides (measured 1.5e-7 on the host build of the kernel arithmetic, tests/test_hostmath_hessian.py): bar
    64 eps max(|theta|, 1) / diff per point; the diff = 1e-4 square stencil at 64 eps max(|theta|, 1) / 1e-4."""
    c = HESSIANS["multi_plane"]
    lm = _LensModel()(c.meta["lens_model_list"], z_source=c.meta["z_source"],
                      lens_redshift_list=c.meta["lens_redshift_list"], multi_plane=True)
    assert lm._sorted_redshift_index == [int(i) for i in c.inp["sorted_index"]]
    lm.set_multi_plane_constants(c.inp["T_ij_list"], c.inp["T_z_list"], c.inp["reduced2physical"],
                                 float(c.inp["T_ij_stop"]), float(c.inp["T_z_source"]))
    x, y, kw = c.inp["x"], c.inp["y"], c.meta["kwargs"]
    theta = np.maximum(np.maximum(np.abs(x), np.abs(y)), 1.0)
    ax, ay = lm.alpha(x, y, kw)
    assert np.all(np.abs(ax - c.out["alpha_x"]) <= 1e-10 * theta) and np.all(np.abs(ay - c.out["alpha_y"]) <= 1e-10 * theta)
    eps = np.finfo(float).eps
    for tag, got, diff in (("d8_", lm.hessian(x, y, kw), 1e-8),
                           ("d4_", lm.hessian(x, y, kw, diff=1e-4, diff_method="square"), 1e-4)):
        for g, q in zip(got, KEYS):
            assert np.all(np.abs(g - c.out[tag + q]) <= 64 * eps * theta / diff), (tag, q)
    with pytest.raises(ValueError):
        lm.hessian(x, y, kw, k=0)          # no subset selection in multi-plane mode, like the reference


def test_scope_errors_and_alpha_scaling():
    LensModel = _LensModel()
    x, y = np.array([0.7, -1.1]), np.array([0.4, 0.9])
    kw = [dict(theta_E=1.0, center_x=0.0, center_y=0.0)]
    mp = LensModel(["SIS"], z_source=1.5, lens_redshift_list=[0.5], multi_plane=True)
    # empty input, and the raw C ABI on a multi-plane plan
    lm = LensModel(["SIS"])
    assert all(h.shape == (0,) for h in lm.hessian(np.zeros(0), np.zeros(0), kw))
    from jaxtronomy_b200 import _lib

    with pytest.raises(_lib.LsimError):
        mp._plan(None).hessian(np.array([[1.0, 0.0, 0.0]]), x, y)
    # SIS closed form: f_xx = theta_E y^2 / r^3, f_xy = -theta_E x y / r^3
    h = lm.hessian(x, y, kw)
    r3 = np.hypot(x, y) ** 3
    np.testing.assert_allclose(h[0], y * y / r3, rtol=1e-13)
    np.testing.assert_allclose(h[1], -x * y / r3, rtol=1e-13)
    np.testing.assert_allclose(h[3], x * x / r3, rtol=1e-13)
