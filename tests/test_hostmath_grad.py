"""CPU: the dual-number profile arithmetic of the gradient path (jaxtronomy_b200/csrc/lsim_grad.cuh compiled for the
host by tests/hostmath) against the oracle.

The reference obtains these derivatives from jax autodiff, which cannot run here (no jax).  The pin is therefore
(i) values: the dual's value part against the oracle restatement (itself pinned to the reference's fixtures) at the
1e-10 bar, and (ii) tangents: 4th-order central finite differences of the oracle in every parameter, stated tolerance
1e-8 of the largest |derivative| of the profile over the test points (the finite-difference error floor, not the
kernel's: the two agree to ~1e-9 where the profile is smooth)."""

import numpy as np
import pytest

from oracle import lens_profiles as olp
from oracle import light_profiles as oli
from tests import hostmath_util as hm

FD_TOL = 1e-8


def fd(fun, p0, i, h):
    """4th-order central difference of fun(params) in parameter i."""
    def at(s):
        p = list(p0)
        p[i] = p0[i] + s
        return np.asarray(fun(p))
    return (8.0 * (at(h) - at(-h)) - (at(2 * h) - at(-2 * h))) / (12.0 * h)


def grid(n=9, lim=2.3):
    g = np.linspace(-lim, lim, n)
    x, y = np.meshgrid(g + 0.0137, g - 0.0291)
    return x.ravel(), y.ravel()


LENS_CASES = [
    ("EPL", dict(theta_E=1.2, gamma=2.1, e1=0.21, e2=-0.13, center_x=0.05, center_y=-0.03), None),
    ("EPL", dict(theta_E=0.9, gamma=1.7, e1=-0.02, e2=0.05, center_x=-0.1, center_y=0.2), None),  # nmax 10
    ("EPL", dict(theta_E=1.4, gamma=2.4, e1=0.45, e2=0.3, center_x=0.0, center_y=0.0), None),  # f = 0.54: nmax 35
    ("EPL", dict(theta_E=1.1, gamma=1.9, e1=-0.55, e2=0.5, center_x=0.0, center_y=0.1), None),  # f = 0.74: nmax 100
    ("EPL", dict(theta_E=1.0, gamma=2.05, e1=0.7, e2=-0.55, center_x=0.03, center_y=0.0), None),  # f = 0.89: nmax 200
    ("SIE", dict(theta_E=1.1, e1=0.15, e2=0.1, center_x=0.02, center_y=0.01), None),
    ("SIE", dict(theta_E=1.1, e1=0.15, e2=0.1, center_x=0.02, center_y=0.01), {"NIE": False}),
    ("NIE", dict(theta_E=1.3, e1=-0.2, e2=0.12, s_scale=0.3, center_x=0.0, center_y=0.04), None),
    ("SIS", dict(theta_E=0.8, center_x=0.1, center_y=-0.2), None),
    ("SHEAR", dict(gamma1=0.05, gamma2=-0.03, ra_0=0.1, dec_0=0.2), None),
    ("CONVERGENCE", dict(kappa=0.1, ra_0=-0.1, dec_0=0.3), None),
    ("NFW", dict(Rs=0.9, alpha_Rs=0.7, center_x=0.05, center_y=-0.02), None),
    ("TNFW", dict(Rs=0.8, alpha_Rs=0.6, r_trunc=2.5, center_x=-0.04, center_y=0.03), None),
    ("MULTIPOLE", dict(m=3, a_m=0.05, phi_m=0.3, center_x=0.02, center_y=0.01, r_E=1.0), None),
    ("MULTIPOLE", dict(m=1, a_m=0.04, phi_m=-0.2, center_x=0.0, center_y=0.05, r_E=1.3), None),
    ("MULTIPOLE", dict(m=4, a_m=0.0, phi_m=0.1, center_x=0.0, center_y=0.0, r_E=1.0), None),  # zero amplitude
    ("CSE", dict(a=1.3, s=0.4, e1=0.15, e2=-0.1, center_x=0.01, center_y=0.02), None),
    ("CSE", dict(a=0.9, s=0.6, e1=-0.2, e2=0.05, center_x=0.0, center_y=0.0), {"axis": "major"}),
    ("NFW_ELLIPSE_CSE", dict(Rs=1.1, alpha_Rs=0.8, e1=0.2, e2=0.1, center_x=0.03, center_y=-0.01), None),
    ("NFW_ELLIPSE_CSE", dict(Rs=0.7, alpha_Rs=0.5, e1=-0.1, e2=0.15, center_x=0.0, center_y=0.02),
     {"high_accuracy": False}),
    ("HERNQUIST_ELLIPSE_CSE", dict(sigma0=1.2, Rs=0.8, e1=0.1, e2=-0.2, center_x=-0.02, center_y=0.0), None),
]


@pytest.mark.parametrize("case", range(len(LENS_CASES)))
def test_lens_dual_against_oracle(case):
    name, kw, pk = LENS_CASES[case]
    names = olp.PARAM_NAMES[name]
    x, y = grid()
    ax, ay, dax, day = hm.lens_jvp([name], [kw], x, y, [pk or {}])
    p0 = [kw[n] for n in names]

    def fun(p):
        return olp.derivatives(name, x, y, profile_kwargs=pk, **dict(zip(names, p)))

    rx, ry = fun(p0)
    scale = max(np.abs(rx).max(), np.abs(ry).max())
    assert np.allclose(ax, rx, rtol=1e-10, atol=1e-12 * scale)
    assert np.allclose(ay, ry, rtol=1e-10, atol=1e-12 * scale)
    for i, n in enumerate(names):
        if n == "m":  # the multipole order is an integer
            assert not dax[:, 0, i].any() and not day[:, 0, i].any()
            continue
        d = fd(fun, p0, i, 1e-4)
        ref = np.stack([d[0], d[1]])
        got = np.stack([dax[:, 0, i], day[:, 0, i]])
        tol = FD_TOL * max(np.abs(ref).max(), 1e-3)
        assert np.abs(got - ref).max() <= tol, (name, n, np.abs(got - ref).max(), tol)
    for i in range(len(names), 6):
        assert not dax[:, 0, i].any() and not day[:, 0, i].any()


def test_lens_dual_sum_and_fixed_parameters():
    x, y = grid(5)
    models = ["EPL", "SHEAR", "SIS"]
    kws = [LENS_CASES[0][1], LENS_CASES[9][1], LENS_CASES[8][1]]
    ax, ay, dax, day = hm.lens_jvp(models, kws, x, y)
    rx = sum(olp.derivatives(m, x, y, **k)[0] for m, k in zip(models, kws))
    ry = sum(olp.derivatives(m, x, y, **k)[1] for m, k in zip(models, kws))
    assert np.allclose(ax, rx, rtol=1e-10) and np.allclose(ay, ry, rtol=1e-10)
    # every profile carries the derivatives of its own parameters only
    _, _, d1, _ = hm.lens_jvp(["SHEAR"], [kws[1]], x, y)
    assert np.array_equal(dax[:, 1], d1[:, 0])


LIGHT_CASES = [
    ("SERSIC_ELLIPSE", dict(amp=12.0, R_sersic=0.4, n_sersic=2.3, e1=0.2, e2=-0.1, center_x=0.05, center_y=-0.1)),
    ("SERSIC_ELLIPSE", dict(amp=3.0, R_sersic=1.1, n_sersic=0.9, e1=-0.3, e2=0.25, center_x=0.0, center_y=0.3)),
    ("SERSIC", dict(amp=5.0, R_sersic=0.7, n_sersic=3.6, center_x=-0.2, center_y=0.1)),
    ("GAUSSIAN", dict(amp=20.0, sigma=0.6, center_x=0.1, center_y=0.0)),
]


@pytest.mark.parametrize("case", range(len(LIGHT_CASES)))
def test_light_dual_against_oracle(case):
    name, kw = LIGHT_CASES[case]
    names = oli.PARAM_NAMES[name]
    x, y = grid(9, 1.7)
    func = oli._FUNCS[name]
    val, dv = hm.light_jvp(name, kw, x, y)
    p0 = [kw[n] for n in names]
    ref = func(x, y, **kw)
    assert np.allclose(val, ref, rtol=1e-10, atol=1e-12 * np.abs(ref).max())
    for i, n in enumerate(names):
        d = fd(lambda p: func(x, y, **dict(zip(names, p))), p0, i, 1e-4 * max(1.0, abs(p0[i])))
        tol = FD_TOL * max(np.abs(d).max(), 1e-3)
        assert np.abs(dv[:, 2 + i] - d).max() <= tol, (name, n, np.abs(dv[:, 2 + i] - d).max(), tol)
    h = 2e-4
    dx = (8 * (func(x + h, y, **kw) - func(x - h, y, **kw)) - (func(x + 2 * h, y, **kw) - func(x - 2 * h, y, **kw))) / (12 * h)
    dy = (8 * (func(x, y + h, **kw) - func(x, y - h, **kw)) - (func(x, y + 2 * h, **kw) - func(x, y - 2 * h, **kw))) / (12 * h)
    tol = FD_TOL * max(np.abs(dx).max(), np.abs(dy).max())
    assert np.abs(dv[:, 0] - dx).max() <= tol and np.abs(dv[:, 1] - dy).max() <= tol
    # amplitude solved linearly: amp = 1 and carries no tangent
    val1, dv1 = hm.light_jvp(name, kw, x, y, use_amp=False)
    assert np.allclose(val1 * kw["amp"], val, rtol=1e-13) and not dv1[:, 2].any()


@pytest.mark.parametrize("case", range(len(LENS_CASES)))
def test_hand_derived_lens_gradient_equals_the_dual_numbers(case):
    """The kernel's fast path (closed-form Jacobian of the EPL / shear, lens_point + lens_vjp) against the dual-number
    evaluation of the same profile: same function, two derivations -- they must agree to round-off."""
    name, kw, pk = LENS_CASES[case]
    x, y = grid(11)
    rng = np.random.default_rng(case)
    vx, vy = rng.normal(size=x.size), rng.normal(size=x.size)
    ax, ay, dax, day = hm.lens_jvp([name], [kw], x, y, [pk or {}])
    bx, by, g = hm.lens_vjp([name], [kw], x, y, vx, vy, [pk or {}])
    scale = max(np.abs(ax).max(), np.abs(ay).max())
    assert np.allclose(bx, ax, rtol=1e-12, atol=1e-13 * scale) and np.allclose(by, ay, rtol=1e-12, atol=1e-13 * scale)
    ref = vx[:, None] * dax[:, 0] + vy[:, None] * day[:, 0]
    for k in range(6):
        tol = 1e-11 * max(np.abs(ref[:, k]).max(), 1e-6)
        assert np.abs(g[:, 0, k] - ref[:, k]).max() <= tol, (name, k, np.abs(g[:, 0, k] - ref[:, k]).max(), tol)


@pytest.mark.parametrize("case", range(len(LIGHT_CASES)))
@pytest.mark.parametrize("use_amp", [True, False])
def test_hand_derived_light_gradient_equals_the_dual_numbers(case, use_amp):
    name, kw = LIGHT_CASES[case]
    x, y = grid(11, 1.9)
    v0, d0 = hm.light_jvp(name, kw, x, y, use_amp=use_amp)
    v1, d1 = hm.light_jvp(name, kw, x, y, use_amp=use_amp, fast=True)
    assert np.allclose(v1, v0, rtol=1e-12, atol=1e-300)
    n = len(oli.PARAM_NAMES[name])
    for k in range(2 + n):
        if k == 2 and not use_amp:
            continue  # the amplitude is not a parameter then; the kernel never reads that entry
        tol = 1e-11 * max(np.abs(d0[:, k]).max(), 1e-300)
        assert np.abs(d1[:, k] - d0[:, k]).max() <= tol, (name, k, np.abs(d1[:, k] - d0[:, k]).max(), tol)


def test_epl_multipole_gradient_through_the_expansion():
    """EPL_MULTIPOLE_M1M3M4 = EPL (closed form) + three multipole pseudo-profiles (dual numbers, unit amplitude) whose
    gradients land in the parent's twelve parameters: against finite differences of the oracle."""
    kw = dict(theta_E=1.15, gamma=2.08, e1=0.18, e2=-0.11, center_x=0.04, center_y=-0.02, a1_a=0.03, delta_phi_m1=0.4,
              a3_a=-0.02, delta_phi_m3=0.25, a4_a=0.015, delta_phi_m4=-0.3)
    names = olp.PARAM_NAMES["EPL_MULTIPOLE_M1M3M4"]
    x, y = grid(9)
    rng = np.random.default_rng(12)
    vx, vy = rng.normal(size=x.size), rng.normal(size=x.size)
    ax, ay, g = hm.eplm_vjp(kw, x, y, vx, vy)
    p0 = [kw[n] for n in names]

    def fun(p):
        return olp.derivatives("EPL_MULTIPOLE_M1M3M4", x, y, **dict(zip(names, p)))

    rx, ry = fun(p0)
    assert np.allclose(ax, rx, rtol=1e-10, atol=1e-12) and np.allclose(ay, ry, rtol=1e-10, atol=1e-12)
    for i, n in enumerate(names):
        d = fd(fun, p0, i, 1e-4)
        ref = vx * d[0] + vy * d[1]
        tol = FD_TOL * max(np.abs(ref).max(), 1e-3)
        assert np.abs(g[:, i] - ref).max() <= tol, (n, np.abs(g[:, i] - ref).max(), tol)


# ---------------------------------------------------------------------------------------------------------------
# multi-plane: forward recursion and its reverse sweep (k_image_grad<0, 2, true>'s per-point code)
# ---------------------------------------------------------------------------------------------------------------
def test_multi_plane_reverse_sweep_matches_finite_differences():
    """beta through mp_ray_forward equals the forward path's multi-plane ray shooting, and the reverse sweep's
    d(v . beta)/d(parameter) equals a 4th-order central difference of the forward path in every parameter of every
    profile -- four planes, a same-plane pair, closed-form (EPL, SHEAR) and dual-number (NFW_ELLIPSE_CSE, SIS) profiles."""
    from tests.golden_io import load

    c = load("lens_models.npz")["multi_plane"]
    names_all, kw_all = c.meta["lens_model_list"], c.meta["kwargs"]
    order = [int(i) for i in c.inp["sorted_index"]]
    names = [names_all[i] for i in order]
    kws = [dict(kw_all[i]) for i in order]
    mp = dict(delta_T=list(c.inp["T_ij_list"]), T_z=list(c.inp["T_z_list"]), red2phys=list(c.inp["reduced2physical"]),
              T_ij_end=float(c.inp["T_ij_stop"]), T_z_source=float(c.inp["T_z_source"]))
    rng = np.random.default_rng(3)
    x, y = rng.uniform(-2, 2, 40), rng.uniform(-2, 2, 40)
    vx, vy = rng.normal(size=40), rng.normal(size=40)
    bx, by, g = hm.mp_ray_vjp(names, kws, x, y, vx, vy, mp)
    fx, fy = hm.ray_shoot(names, kws, x, y, multi_plane=mp)
    assert np.max(np.abs(bx - fx)) < 1e-13 and np.max(np.abs(by - fy)) < 1e-13
    # and against the reference fixture itself (same points? no: the fixture's own points)
    rx, ry = hm.ray_shoot(names, kws, c.inp["x"], c.inp["y"], multi_plane=mp)
    assert np.max(np.abs(rx - c.out["beta_x"])) < 1e-10 * np.max(np.abs(c.out["beta_x"]))

    def s_of(kws2):
        ox, oy = hm.ray_shoot(names, kws2, x, y, multi_plane=mp)
        return vx * ox + vy * oy

    from jaxtronomy_b200 import _model

    for l, name in enumerate(names):
        for k, pname in enumerate(_model.LENS_PARAM_NAMES[name]):
            h = 2e-5 * max(1.0, abs(kws[l][pname]))
            vals = []
            for step in (-2 * h, -h, h, 2 * h):
                k2 = [dict(kk) for kk in kws]
                k2[l][pname] = kws[l][pname] + step
                vals.append(s_of(k2))
            fd = (8.0 * (vals[2] - vals[1]) - (vals[3] - vals[0])) / (12.0 * h)
            scale = max(np.max(np.abs(fd)), np.max(np.abs(g[:, l, :])), 1e-300)
            err = np.max(np.abs(g[:, l, k] - fd)) / scale
            assert err < 2e-8, (name, pname, err)
