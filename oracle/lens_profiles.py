"""Oracle: deflection angles (``derivatives``) of the hot-path lens profiles.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Each function restates the reference's arithmetic (clamps and sentinels
included) in NumPy and cites the file:line it follows under
/root/reference/jaxtronomy/LensModel/Profiles/.
"""

import numpy as np

from . import cse_tables
from .util import (
    cart2polar,
    ellipticity2phi_q,
    epl_nmax,
    hyp2f1_lopez_temme_8,
    nan_to_num,
    rotate,
    shift_center,
)

# param_names per profile string (the kwargs contract)
PARAM_NAMES = {
    "EPL": ["theta_E", "gamma", "e1", "e2", "center_x", "center_y"],  # epl.py:50
    "SIE": ["theta_E", "e1", "e2", "center_x", "center_y"],  # sie.py:40
    "NIE": ["theta_E", "e1", "e2", "s_scale", "center_x", "center_y"],  # nie.py:19
    "SIS": ["theta_E", "center_x", "center_y"],  # sis.py:20
    "SHEAR": ["gamma1", "gamma2", "ra_0", "dec_0"],  # shear.py:16
    "CONVERGENCE": ["kappa", "ra_0", "dec_0"],  # convergence.py:15
    "NFW": ["Rs", "alpha_Rs", "center_x", "center_y"],  # nfw.py:48
    "TNFW": ["Rs", "alpha_Rs", "r_trunc", "center_x", "center_y"],  # tnfw.py:27
    "MULTIPOLE": ["m", "a_m", "phi_m", "center_x", "center_y", "r_E"],  # multipole.py:23
    "EPL_MULTIPOLE_M1M3M4": [
        "theta_E", "gamma", "e1", "e2", "center_x", "center_y",
        "a1_a", "delta_phi_m1", "a3_a", "delta_phi_m3", "a4_a", "delta_phi_m4",
    ],  # epl_multipole_m1m3m4.py:415-428
    "CSE": ["a", "s", "e1", "e2", "center_x", "center_y"],  # cored_steep_ellipsoid.py:37
    "NFW_ELLIPSE_CSE": ["Rs", "alpha_Rs", "e1", "e2", "center_x", "center_y"],  # nfw_ellipse_cse.py:172
    "HERNQUIST_ELLIPSE_CSE": ["sigma0", "Rs", "e1", "e2", "center_x", "center_y"],  # hernquist_ellipse_cse.py:114
    "NULL": [],  # null.py:16
}


# --- EPL ---------------------------------------------------------------------------------------
def epl_major_axis(x, y, b, t, q):
    """EPLMajorAxis.derivatives, epl.py:366-394; _hyp2f1_evaluate epl.py:297-319."""
    x = np.asarray(x, dtype=float)
    y = np.asarray(y, dtype=float)
    Z = q * x + y * 1j
    R = np.maximum(np.abs(Z), 0.000000001)
    f = (1.0 - q) / (1.0 + q)
    with np.errstate(all="ignore"):
        z = -f * Z / np.conj(Z)
        nmax = epl_nmax(float(f))
        if nmax == 0:
            h = np.ones_like(z)
        else:
            B = t / 2.0
            C = 2.0 - B
            h = hyp2f1_lopez_temme_8(1, B, C, z, nmax=nmax)
        R_omega = Z * h
        alpha = 2 / (1 + q) * (b / R) ** t * R_omega
    ar = nan_to_num(alpha.real, posinf=10**10, neginf=-(10**10))
    ai = nan_to_num(alpha.imag, posinf=10**10, neginf=-(10**10))
    return ar, ai


def epl(x, y, theta_E, gamma, e1, e2, center_x=0, center_y=0):
    """EPL.derivatives, epl.py:178-200; _param_conv epl.py:109-124."""
    t = gamma - 1
    phi_G, q = ellipticity2phi_q(e1, e2)
    b = theta_E * np.sqrt(q)
    x_, y_ = shift_center(x, y, center_x, center_y)
    x_, y_ = rotate(x_, y_, phi_G)
    fx, fy = epl_major_axis(x_, y_, b, t, q)
    return rotate(fx, fy, -phi_G)


# --- NIE / SIE ---------------------------------------------------------------------------------
def nie_major_axis(x, y, b, s, q):
    """NIEMajorAxis.derivatives / _psi, nie.py:280-294,333-342."""
    x = np.asarray(x, dtype=float)
    y = np.asarray(y, dtype=float)
    q = np.where(q >= 1, 0.99999999, q)
    psi = np.sqrt(q**2 * (s**2 + x**2) + y**2)
    with np.errstate(all="ignore"):
        f_x = b / np.sqrt(1.0 - q**2) * np.arctan(np.sqrt(1.0 - q**2) * x / (psi + s))
        f_y = b / np.sqrt(1.0 - q**2) * np.arctanh(np.sqrt(1.0 - q**2) * y / (psi + q**2 * s))
    return f_x, f_y


def nie(x, y, theta_E, e1, e2, s_scale, center_x=0, center_y=0):
    """NIE.derivatives / _param_conv, nie.py:85-107,181-200,230-246."""
    phi_G, q = ellipticity2phi_q(e1, e2)
    theta_E_conv = theta_E / (np.sqrt((1.0 + q**2) / (2.0 * q)))
    b = theta_E_conv * np.sqrt((1 + q**2) / 2)
    s = s_scale / np.sqrt(q)
    x_, y_ = shift_center(x, y, center_x, center_y)
    x_, y_ = rotate(x_, y_, phi_G)
    fx, fy = nie_major_axis(x_, y_, b, s, q)
    return rotate(fx, fy, -phi_G)


def sie(x, y, theta_E, e1, e2, center_x=0, center_y=0, s_scale=0.0000000001, gamma=2, NIE=True):
    """SIE.derivatives, sie.py:56-63,104-124 (default NIE=True, s_scale=1e-10)."""
    if NIE:
        return nie(x, y, theta_E, e1, e2, s_scale, center_x, center_y)
    return epl(x, y, theta_E, gamma, e1, e2, center_x, center_y)


# --- SIS / SHEAR / CONVERGENCE / NULL --------------------------------------------------------
def sis(x, y, theta_E, center_x=0, center_y=0):
    """SIS.derivatives, sis.py:35-43 (_epsilon=1e-6, sis.py:24)."""
    eps = 0.000001
    x_, y_ = shift_center(x, y, center_x, center_y)
    R = np.sqrt(x_**2 + y_**2)
    R = np.where(R < eps, eps, R)
    a = theta_E / np.maximum(R, eps)
    return a * x_, a * y_


def shear(x, y, gamma1, gamma2, ra_0=0, dec_0=0):
    """Shear.derivatives, shear.py:39-53."""
    x_, y_ = shift_center(x, y, ra_0, dec_0)
    return gamma1 * x_ + gamma2 * y_, gamma2 * x_ - gamma1 * y_


def convergence(x, y, kappa, ra_0=0, dec_0=0):
    """Convergence.derivatives, convergence.py:35-46."""
    x_, y_ = shift_center(x, y, ra_0, dec_0)
    return kappa * x_, kappa * y_


def null(x, y, **kwargs):
    """Null.derivatives, null.py:30-38."""
    x = np.asarray(x, dtype=float)
    return np.zeros_like(x), np.zeros_like(x)


# --- NFW ---------------------------------------------------------------------------------------
def nfw_g(X):
    """NFW.g, nfw.py:347-365."""
    c = 0.000001
    X = np.where(X < c, c, X)
    with np.errstate(all="ignore"):
        a = np.where(
            X < 1,
            np.log(X / 2.0) + 1 / np.sqrt(1 - X**2) * np.arccosh(1.0 / X),
            np.where(
                X == 1,
                1 + np.log(1.0 / 2.0),
                np.log(X / 2) + 1 / np.sqrt(X**2 - 1) * np.arccos(1.0 / X),
            ),
        )
    return a


def nfw_alpha2rho0(alpha_Rs, Rs):
    """NFW.alpha2rho0, nfw.py:386-395."""
    return alpha_Rs / (4.0 * Rs**2 * (1.0 + np.log(1.0 / 2.0)))


def nfw(x, y, Rs, alpha_Rs, center_x=0, center_y=0):
    """NFW.derivatives / nfw_alpha, nfw.py:88-105,259-279."""
    rho0 = nfw_alpha2rho0(alpha_Rs, Rs)
    Rs = np.where(Rs < 0.0000001, 0.0000001, Rs)
    x_, y_ = shift_center(x, y, center_x, center_y)
    R = np.sqrt(x_**2 + y_**2)
    R = np.maximum(R, 0.00000001)
    X = R / Rs
    gx = nfw_g(X)
    a = 4 * rho0 * Rs * gx / X**2
    return a * x_, a * y_


# --- TNFW --------------------------------------------------------------------------------------
_TNFW_S = 0.001  # tnfw.py:43


def tnfw_F(x):
    """TNFW.F, tnfw.py:290-304."""
    x = np.maximum(x, _TNFW_S)
    with np.errstate(all="ignore"):
        result = np.where(x < 1, (1 - x**2) ** -0.5 * np.arctanh((1 - x**2) ** 0.5), x)
        result = np.where(x > 1, (x**2 - 1) ** -0.5 * np.arctan((x**2 - 1) ** 0.5), result)
    return result


def tnfw_L(x, tau):
    """TNFW._L, tnfw.py:278-286."""
    x = np.maximum(x, _TNFW_S)
    return np.log(x * (tau + np.sqrt(tau**2 + x**2)) ** -1)


def tnfw_g(x, tau):
    """TNFW._g, tnfw.py:331-349."""
    x = np.maximum(x, _TNFW_S)
    return (
        tau**2
        / (tau**2 + 1) ** 2
        * (
            (tau**2 + 1 + 2 * (x**2 - 1)) * tnfw_F(x)
            + tau * np.pi
            + (tau**2 - 1) * np.log(tau)
            + np.sqrt(tau**2 + x**2) * (-np.pi + tnfw_L(x, tau) * (tau**2 - 1) / tau)
        )
    )


def tnfw(x, y, Rs, alpha_Rs, r_trunc, center_x=0, center_y=0):
    """TNFW.derivatives / tnfw_alpha, tnfw.py:69-87,228-248."""
    rho0 = nfw_alpha2rho0(alpha_Rs, Rs)
    x_, y_ = shift_center(x, y, center_x, center_y)
    R = np.sqrt(x_**2 + y_**2)
    R = np.maximum(R, _TNFW_S * Rs)
    X = np.maximum(R / Rs, _TNFW_S)
    tau = r_trunc / Rs
    gx = tnfw_g(X, tau)
    a = 4 * rho0 * Rs * gx / X**2
    return a * x_, a * y_


# --- multipoles --------------------------------------------------------------------------------
def multipole(x, y, m, a_m, phi_m, center_x=0, center_y=0, r_E=1):
    """Multipole.derivatives, multipole.py:72-119."""
    r, phi = cart2polar(x, y, center_x=center_x, center_y=center_y)
    if m == 1:
        r = np.maximum(r, 0.000001)
        f_x = a_m / 2 * (np.cos(phi_m) * np.log(r / r_E) + np.cos(phi - phi_m) * np.cos(phi))
        f_y = a_m / 2 * (np.sin(phi_m) * np.log(r / r_E) + np.cos(phi - phi_m) * np.sin(phi))
        return f_x, f_y
    f_x = np.cos(phi) * a_m / (1 - m**2) * np.cos(m * (phi - phi_m)) + np.sin(phi) * m * a_m / (
        1 - m**2
    ) * np.sin(m * (phi - phi_m))
    f_y = np.sin(phi) * a_m / (1 - m**2) * np.cos(m * (phi - phi_m)) - np.cos(phi) * m * a_m / (
        1 - m**2
    ) * np.sin(m * (phi - phi_m))
    return f_x, f_y


def epl_multipole_m1m3m4(
    x, y, theta_E, gamma, e1, e2, center_x=0, center_y=0,
    a1_a=0, delta_phi_m1=0, a3_a=0, delta_phi_m3=0, a4_a=0, delta_phi_m4=0,
):
    """EPL_MULTIPOLE_M1M3M4.derivatives/_param_split, epl_multipole_m1m3m4.py:463-535,608-678.

    Keyed by name (kwargs path, single_plane.py:135); sum order epl + m3 + m4 + m1.
    """
    phi, q = ellipticity2phi_q(e1, e2)
    rescale = theta_E / np.sqrt(q)
    fx, fy = epl(x, y, theta_E, gamma, e1, e2, center_x, center_y)
    fx3, fy3 = multipole(x, y, 3, a3_a * rescale, phi + delta_phi_m3, center_x, center_y)
    fx4, fy4 = multipole(x, y, 4, a4_a * rescale, phi + delta_phi_m4, center_x, center_y)
    fx1, fy1 = multipole(x, y, 1, a1_a * rescale, phi + delta_phi_m1, center_x, center_y, r_E=theta_E)
    return fx + fx3 + fx4 + fx1, fy + fy3 + fy4 + fy1


# --- CSE family --------------------------------------------------------------------------------
def cse_major_axis(x, y, a, s, q):
    """CSEMajorAxis.derivatives, cored_steep_ellipsoid.py:227-245."""
    x = np.asarray(x, dtype=float)
    y = np.asarray(y, dtype=float)
    psi = np.sqrt(q**2 * (s**2 + x**2) + y**2)
    Phi = (psi + s) ** 2 + (1 - q**2) * x**2
    f_x = q * x * (psi + q**2 * s) / (s * psi * Phi)
    f_y = q * y * (psi + s) / (s * psi * Phi)
    return a * f_x, a * f_y


def cse_product_avg(x, y, a, s, q):
    """CSEProductAvg.derivatives/_convert2prodavg, cored_steep_ellipsoid.py:434-443,463-477."""
    x = np.asarray(x, dtype=float)
    y = np.asarray(y, dtype=float)
    a = a / q
    x = x * np.sqrt(q)
    y = y * np.sqrt(q)
    fx, fy = cse_major_axis(x, y, a, s, q)
    return np.sqrt(q) * fx, np.sqrt(q) * fy


def cse(x, y, a, s, e1, e2, center_x, center_y, axis="product_avg"):
    """CSE.derivatives, cored_steep_ellipsoid.py:107-132."""
    phi_q, q = ellipticity2phi_q(e1, e2)
    x_, y_ = shift_center(x, y, center_x, center_y)
    x_, y_ = rotate(x_, y_, phi_q)
    if axis == "major":
        fx, fy = cse_major_axis(x_, y_, a, s, q)
    else:
        fx, fy = cse_product_avg(x_, y_, a, s, q)
    return rotate(fx, fy, -phi_q)


def _cse_set(func, x, y, a_list, s_list, q):
    """CSEMajorAxisSet / CSEProductAvgSet.derivatives, cored_steep_ellipsoid.py:330-359,536-564."""
    x = np.asarray(x, dtype=float)
    y = np.asarray(y, dtype=float)
    f_x, f_y = np.zeros_like(x), np.zeros_like(y)
    for a, s in zip(a_list, s_list):
        fx_, fy_ = func(x, y, a, s, q)
        f_x = f_x + fx_
        f_y = f_y + fy_
    return f_x, f_y


def nfw_ellipse_cse(x, y, Rs, alpha_Rs, e1, e2, center_x=0, center_y=0, high_accuracy=True):
    """NFW_ELLIPSE_CSE.derivatives/_normalization, nfw_ellipse_cse.py:250-280,338-349."""
    phi_q, q = ellipticity2phi_q(e1, e2)
    x_, y_ = shift_center(x, y, center_x, center_y)
    x_, y_ = rotate(x_, y_, phi_q)
    if high_accuracy:
        A, S = cse_tables.NFW_CSE_HIGH_A, cse_tables.NFW_CSE_HIGH_S
    else:
        A, S = cse_tables.NFW_CSE_LOW_A, cse_tables.NFW_CSE_LOW_S
    fx, fy = _cse_set(cse_product_avg, x_ / Rs, y_ / Rs, A, S, q)
    fx, fy = rotate(fx, fy, -phi_q)
    rho0 = nfw_alpha2rho0(alpha_Rs, Rs)
    const = 4 * rho0 * Rs**3 / Rs
    return const * fx, const * fy


def hernquist_ellipse_cse(x, y, sigma0, Rs, e1, e2, center_x=0, center_y=0):
    """HernquistEllipseCSE.derivatives/_normalization, hernquist_ellipse_cse.py:151-164,190-202."""
    phi_q, q = ellipticity2phi_q(e1, e2)
    x_, y_ = shift_center(x, y, center_x, center_y)
    x_, y_ = rotate(x_, y_, phi_q)
    fx, fy = _cse_set(
        cse_major_axis, x_ / Rs, y_ / Rs, cse_tables.HERNQUIST_CSE_A, cse_tables.HERNQUIST_CSE_S, q
    )
    fx, fy = rotate(fx, fy, -phi_q)
    rs_ = Rs / np.sqrt(q)
    const = sigma0 * rs_**2 / Rs
    return const * fx, const * fy


_FUNCS = {
    "EPL": epl,
    "SIE": sie,
    "NIE": nie,
    "SIS": sis,
    "SHEAR": shear,
    "CONVERGENCE": convergence,
    "NULL": null,
    "NFW": nfw,
    "TNFW": tnfw,
    "MULTIPOLE": multipole,
    "EPL_MULTIPOLE_M1M3M4": epl_multipole_m1m3m4,
    "CSE": cse,
    "NFW_ELLIPSE_CSE": nfw_ellipse_cse,
    "HERNQUIST_ELLIPSE_CSE": hernquist_ellipse_cse,
}


def derivatives(profile, x, y, profile_kwargs=None, **kwargs):
    """Deflection of one named profile; ``profile_kwargs`` are the init options
    (profile_list_base.py:161-232), ``kwargs`` the ``param_names`` values."""
    opts = dict(profile_kwargs or {})
    return _FUNCS[profile](x, y, **kwargs, **opts)


# --- Hessians (closed forms of the reference; the remaining profiles are covered by the numerical derivative of the
# --- deflection, oracle/lens_model.py: hessian_differential, and directly by tests/golden/hessians.npz) -----------
def epl_major_axis_hessian(x, y, b, t, q):
    """EPLMajorAxis.hessian, epl.py:398-435."""
    x = np.asarray(x, dtype=float)
    y = np.asarray(y, dtype=float)
    with np.errstate(all="ignore"):
        R = np.maximum(np.hypot(q * x, y), 0.00000001)
        r = np.hypot(x, y)
        cos, sin = x / r, y / r
        cos2, sin2 = cos * cos * 2 - 1, sin * cos * 2
        kappa = nan_to_num((2 - t) / 2 * (b / R) ** t, posinf=10**10, neginf=-(10**10))
        alpha_x, alpha_y = epl_major_axis(x, y, b, t, q)
        gamma_1 = (1 - t) * (alpha_x * cos - alpha_y * sin) / r - kappa * cos2
        gamma_2 = (1 - t) * (alpha_y * cos + alpha_x * sin) / r - kappa * sin2
        gamma_1 = nan_to_num(gamma_1, posinf=10**10, neginf=-(10**10))
        gamma_2 = nan_to_num(gamma_2, posinf=10**10, neginf=-(10**10))
    return kappa + gamma_1, gamma_2, gamma_2, kappa - gamma_1


def epl_hessian(x, y, theta_E, gamma, e1, e2, center_x=0, center_y=0):
    """EPL.hessian, epl.py:203-232: major-axis Hessian, shear rotated back by 2 phi_G."""
    t = gamma - 1
    phi_G, q = ellipticity2phi_q(e1, e2)
    b = theta_E * np.sqrt(q)
    x_, y_ = shift_center(x, y, center_x, center_y)
    x_, y_ = rotate(x_, y_, phi_G)
    f__xx, f__xy, _, f__yy = epl_major_axis_hessian(x_, y_, b, t, q)
    kappa = 1.0 / 2 * (f__xx + f__yy)
    gamma1__ = 1.0 / 2 * (f__xx - f__yy)
    gamma2__ = f__xy
    gamma1 = np.cos(2 * phi_G) * gamma1__ - np.sin(2 * phi_G) * gamma2__
    gamma2 = np.sin(2 * phi_G) * gamma1__ + np.cos(2 * phi_G) * gamma2__
    return kappa + gamma1, gamma2, gamma2, kappa - gamma1


def sis_hessian(x, y, theta_E, center_x=0, center_y=0):
    """SIS.hessian, sis.py:47-57."""
    eps = 0.000001
    x_, y_ = shift_center(x, y, center_x, center_y)
    R = (x_**2 + y_**2) ** (3.0 / 2)
    R = np.where(R < eps, eps, R)
    prefac = theta_E / np.maximum(eps, R)
    return y_ * y_ * prefac, -x_ * y_ * prefac, -x_ * y_ * prefac, x_ * x_ * prefac


def shear_hessian(x, y, gamma1, gamma2, ra_0=0, dec_0=0):
    """Shear.hessian, shear.py:57-74."""
    o = np.ones_like(np.asarray(x, dtype=float))
    return gamma1 * o, gamma2 * o, gamma2 * o, -gamma1 * o


def convergence_hessian(x, y, kappa, ra_0=0, dec_0=0):
    """Convergence.hessian, convergence.py:50-66."""
    o = np.ones_like(np.asarray(x, dtype=float))
    return kappa * o, 0 * o, 0 * o, kappa * o


_HESSIANS = {"EPL": epl_hessian, "SIS": sis_hessian, "SHEAR": shear_hessian, "CONVERGENCE": convergence_hessian}


def hessian(profile, x, y, profile_kwargs=None, **kwargs):
    """Closed-form Hessian of one named profile, or None where only the numerical derivative is restated."""
    opts = dict(profile_kwargs or {})
    if profile == "SIE" and opts.get("NIE") is False:
        return epl_hessian(x, y, kwargs["theta_E"], opts.get("gamma", 2), kwargs["e1"], kwargs["e2"],
                           kwargs.get("center_x", 0), kwargs.get("center_y", 0))
    if profile not in _HESSIANS:
        return None
    return _HESSIANS[profile](x, y, **kwargs)
