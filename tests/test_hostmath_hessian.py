"""CPU: k_hessian's per-point arithmetic (lens_prepare_hess / lens_hessian_add of jaxtronomy_b200/csrc/lsim_grad.cuh,
compiled for the host by tests/hostmath) against the reference-generated fixtures and the oracle."""

import numpy as np
import pytest

from oracle import lens_model as olm
from tests import hostmath_util as hm
from tests.hessian_cases import HESSIANS, KEYS, PROFILE_CASES, check_profile, keep_mask, profile_case


@pytest.mark.parametrize("name", PROFILE_CASES)
def test_profile_hessian_matches_the_reference_fixture(name):
    c, prof, kw, pk = profile_case(name)
    got = hm.lens_hessian([prof], [kw], c.inp["x"], c.inp["y"], profile_kwargs_list=[pk])
    check_profile(name, got, c, prof, kw)


@pytest.mark.parametrize("name", PROFILE_CASES)
def test_profile_hessian_is_the_derivative_of_the_oracle_deflection(name):
    """Independent of the fixtures: 4th-order central difference of the oracle's alpha, 1e-8 of the scale.
    This is the check that carries NIE / SIE, whose reference Hessian is itself a noisy forward difference."""
    c, prof, kw, pk = profile_case(name)
    x, y = c.inp["x"], c.inp["y"]
    keep = keep_mask(x, y, [kw]) & (np.hypot(x - kw.get("center_x", kw.get("ra_0")), y - kw.get("center_y", kw.get("dec_0"))) > 0.3)
    x, y = x[keep], y[keep]
    got = hm.lens_hessian([prof], [kw], x, y, profile_kwargs_list=[pk])
    want = olm.hessian_differential([prof], x, y, [kw], profile_kwargs_list=[pk], h=1e-3)
    scale = max(np.max(np.abs(w)) for w in want)
    # sie_2 / nie_2: |e| clamps to 0.9999 (q = 5e-5), the deflection varies on the scale q R and the finite
    # difference's own truncation error grows accordingly
    tol = 1e-6 if name in ("sie_2", "nie_2") else 1e-8
    for g, w, k in zip(got, want, KEYS):
        assert np.max(np.abs(g - w)) <= tol * scale, "%s %s: %.3g of scale %.3g" % (name, k, np.max(np.abs(g - w)), scale)


def test_model_hessian_sums_the_profiles():
    c = HESSIANS["model_knull"]
    names, kw = c.meta["lens_model_list"], c.meta["kwargs"]
    x, y = c.inp["x"], c.inp["y"]
    keep = keep_mask(x, y, kw)
    got = hm.lens_hessian(names, kw, x, y)
    scale = max(np.max(np.abs(c.out[k][keep])) for k in KEYS)
    for g, k in zip(got, KEYS):
        assert np.max(np.abs(g - c.out[k])[keep]) <= 1e-10 * scale, k


def test_multi_plane_forward_difference_noise_is_inside_the_stated_bar():
    """The multi-plane Hessian of the reference is a forward difference of alpha = theta - beta with diff = 1e-8
    (multi_plane.py:537-589).  The same stencil on the kernel's host-compiled ray shooting agrees with the reference
    fixture to the rounding of the ray shooting over diff: bar 64 eps max(|theta|, 1) / diff (tests/test_gpu_zhessian.py
    uses the same bar for the device)."""
    c = HESSIANS["multi_plane"]
    order = [int(i) for i in c.inp["sorted_index"]]
    names = [c.meta["lens_model_list"][i] for i in order]
    kws = [c.meta["kwargs"][i] for i in order]
    mp = dict(delta_T=c.inp["T_ij_list"], T_z=c.inp["T_z_list"], red2phys=c.inp["reduced2physical"],
              T_ij_end=float(c.inp["T_ij_stop"]), T_z_source=float(c.inp["T_z_source"]))
    x, y = c.inp["x"], c.inp["y"]

    def alpha(xx, yy):
        bx, by = hm.ray_shoot(names, kws, xx, yy, multi_plane=mp)
        return xx - bx, yy - by

    theta = np.maximum(np.maximum(np.abs(x), np.abs(y)), 1.0)
    ax, ay = alpha(x, y)
    assert np.all(np.abs(ax - c.out["alpha_x"]) <= 1e-10 * theta) and np.all(np.abs(ay - c.out["alpha_y"]) <= 1e-10 * theta)
    d = 1e-8
    axd, ayd = alpha(x + d, y)
    axe, aye = alpha(x, y + d)
    eps = np.finfo(float).eps
    for g, q in zip(((axd - ax) / d, (axe - ax) / d, (ayd - ay) / d, (aye - ay) / d), KEYS):
        assert np.all(np.abs(g - c.out["d8_" + q]) <= 64 * eps * theta / d), q


def test_multi_plane_facade_arithmetic_on_the_host_ray_shooting(monkeypatch):
    """LensModel.alpha / hessian / magnification of a multi-plane model are façade arithmetic over ray shooting.  With
    `_run` (the one GPU launch inside) replaced by the host build of the same kernel code, the whole façade path runs on
    the CPU against the reference fixture -- same bars as the GPU test."""
    from jaxtronomy_b200.LensModel.lens_model import LensModel

    c = HESSIANS["multi_plane"]
    order = [int(i) for i in c.inp["sorted_index"]]
    names = [c.meta["lens_model_list"][i] for i in order]
    mp = dict(delta_T=c.inp["T_ij_list"], T_z=c.inp["T_z_list"], red2phys=c.inp["reduced2physical"],
              T_ij_end=float(c.inp["T_ij_stop"]), T_z_source=float(c.inp["T_z_source"]))

    def host_run(self, x, y, kwargs, k, what):
        assert what == 0 and k is None
        return hm.ray_shoot(names, [kwargs[i] for i in order], np.asarray(x, dtype=float), np.asarray(y, dtype=float),
                            multi_plane=mp)

    monkeypatch.setattr(LensModel, "_run", host_run)
    lm = LensModel(c.meta["lens_model_list"], z_source=c.meta["z_source"],
                   lens_redshift_list=c.meta["lens_redshift_list"], multi_plane=True)
    x, y, kw = c.inp["x"], c.inp["y"], c.meta["kwargs"]
    theta = np.maximum(np.maximum(np.abs(x), np.abs(y)), 1.0)
    eps = np.finfo(float).eps
    ax, ay = lm.alpha(x, y, kw)
    assert np.all(np.abs(ax - c.out["alpha_x"]) <= 1e-10 * theta)
    for tag, got, diff in (("d8_", lm.hessian(x, y, kw), 1e-8),
                           ("d4_", lm.hessian(x, y, kw, diff=1e-4, diff_method="square"), 1e-4)):
        for g, q in zip(got, KEYS):
            assert np.all(np.abs(g - c.out[tag + q]) <= 64 * eps * theta / diff), (tag, q)
    # mu = 1 / det A from the noisy forward difference: relative agreement degrades with |mu| (d mu = mu^2 d det)
    mu = lm.magnification(x, y, kw)
    want = c.out["magnification"]
    assert np.all(np.abs(mu - want) <= 4 * 64 * eps * theta / 1e-8 * want**2 * 20 + 1e-9 * np.abs(want))
    f = lm.flexion(x, y, kw, diff=1e-2, hessian_diff=True)   # runs through hessian(diff=...) -> alpha: shapes only
    assert all(np.shape(a) == np.shape(x) for a in f)
