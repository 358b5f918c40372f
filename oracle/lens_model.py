"""Oracle: profile-list summation, ray shooting (single plane, bulk, multi-plane).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Follows
  jaxtronomy/LensModel/lens_model.py:236-252          LensModel.ray_shooting
  jaxtronomy/LensModel/single_plane.py:38-56,112-139  SinglePlane.ray_shooting / alpha
  jaxtronomy/LensModel/single_plane_bulk.py:47-131    SinglePlaneBulk
  jaxtronomy/LensModel/MultiPlane/multi_plane.py:236-275, 685-696
  jaxtronomy/LensModel/MultiPlane/multi_plane_base.py:89-115,159-233,473-544
"""

import numpy as np

from . import lens_profiles
from .light_profiles import _bool_list


def alpha(lens_model_list, x, y, kwargs, k=None, profile_kwargs_list=None, alpha_scaling=1):
    x = np.asarray(x, dtype=float)
    y = np.asarray(y, dtype=float)
    if profile_kwargs_list is None:
        profile_kwargs_list = [{} for _ in lens_model_list]
    if isinstance(k, (int, np.integer)) and not isinstance(k, bool):
        return lens_profiles.derivatives(lens_model_list[k], x, y, profile_kwargs_list[k], **kwargs[k])
    bl = _bool_list(len(lens_model_list), k)
    f_x, f_y = np.zeros_like(x), np.zeros_like(x)
    for i, name in enumerate(lens_model_list):
        if bl[i]:
            fx_i, fy_i = lens_profiles.derivatives(name, x, y, profile_kwargs_list[i], **kwargs[i])
            f_x = f_x + fx_i
            f_y = f_y + fy_i
    return f_x * alpha_scaling, f_y * alpha_scaling


def ray_shooting(lens_model_list, x, y, kwargs, k=None, profile_kwargs_list=None, alpha_scaling=1):
    x = np.asarray(x, dtype=float)
    y = np.asarray(y, dtype=float)
    dx, dy = alpha(lens_model_list, x, y, kwargs, k, profile_kwargs_list, alpha_scaling)
    return x - dx, y - dy


def ray_shooting_bulk(unique_lens_model_list, x, y, all_kwargs, index_list, profile_kwargs_list=None):
    """SinglePlaneBulk.ray_shooting (single_plane_bulk.py:108-131): Sum_i switch(index_i) then beta = theta - sum."""
    x = np.asarray(x, dtype=float)
    y = np.asarray(y, dtype=float)
    names = list(unique_lens_model_list) + ["NULL"]
    if profile_kwargs_list is None:
        profile_kwargs_list = [{} for _ in names]
    sx, sy = np.zeros_like(x), np.zeros_like(y)
    for d, idx in enumerate(index_list):
        name = names[int(idx)]
        kw = {p: all_kwargs[p][d] for p in lens_profiles.PARAM_NAMES[name]}
        fx, fy = lens_profiles.derivatives(name, x, y, profile_kwargs_list[int(idx)] if int(idx) < len(profile_kwargs_list) else {}, **kw)
        sx = sx + fx
        sy = sy + fy
    return x - sx, y - sy


def multi_plane_constants(lens_redshift_list, z_source, z_source_convention, T_xy, d_xy):
    """Host constants of MultiPlaneBase.set_T_zs_and_T_ijs (multi_plane_base.py:89-115) and
    MultiPlane._set_source_distances (multi_plane.py:190-204).

    ``T_xy(z1,z2)`` / ``d_xy(z1,z2)`` are the transverse comoving / angular-diameter
    distance callables (astropy in the reference; any consistent pair works).
    Returns (sorted_index, T_ij_list, T_z_list, reduced2physical, T_ij_stop, T_z_source).
    """
    order = [int(i) for i in np.argsort(np.asarray(lens_redshift_list))]  # multi_plane_base.py:409-420
    # planes behind the source are never visited (multi_plane_base.py:195-201); sorted => a suffix
    order = [i for i in order if 0 < lens_redshift_list[i] <= z_source]
    T_ij, T_z_list, fac = [], [], []
    z_before, T_z = 0, 0
    for idx in order:
        z_lens = lens_redshift_list[idx]
        if z_before == z_lens:
            delta_T = 0
        else:
            T_z = T_xy(0, z_lens)
            delta_T = T_xy(z_before, z_lens)
        T_ij.append(delta_T)
        T_z_list.append(T_z)
        fac.append(d_xy(0, z_source_convention) / d_xy(z_lens, z_source_convention))
        z_before = z_lens
    # transverse_distance_start_stop(0, z_source): last lens with z <= z_source -> source
    z_last = 0
    for idx in order:
        if lens_redshift_list[idx] <= z_source:
            z_last = lens_redshift_list[idx]
    T_ij_stop = T_xy(z_last, z_source)
    return order, T_ij, T_z_list, fac, T_ij_stop, T_xy(0, z_source)


def ray_shooting_multi_plane(
    lens_model_list, theta_x, theta_y, kwargs_lens, sorted_index, T_ij_list, T_z_list,
    reduced2physical, T_ij_end, T_z_source, profile_kwargs_list=None,
):
    """MultiPlane.ray_shooting (multi_plane.py:236-275) with the recursion of
    ray_shooting_partial_comoving (multi_plane_base.py:159-233), all planes with z <= z_source."""
    if profile_kwargs_list is None:
        profile_kwargs_list = [{} for _ in lens_model_list]
    alpha_x = np.asarray(theta_x, dtype=float)
    alpha_y = np.asarray(theta_y, dtype=float)
    x = np.zeros_like(alpha_x)
    y = np.zeros_like(alpha_y)
    for i, k in enumerate(sorted_index):
        delta_T = T_ij_list[i]
        x = x + alpha_x * delta_T
        y = y + alpha_y * delta_T
        th_x, th_y = x / T_z_list[i], y / T_z_list[i]
        ax_red, ay_red = lens_profiles.derivatives(
            lens_model_list[k], th_x, th_y, profile_kwargs_list[k], **kwargs_lens[k]
        )
        alpha_x = alpha_x - ax_red * reduced2physical[i]
        alpha_y = alpha_y - ay_red * reduced2physical[i]
    x = x + alpha_x * T_ij_end
    y = y + alpha_y * T_ij_end
    return x / T_z_source, y / T_z_source


def hessian_differential(lens_model_list, x, y, kwargs, k=None, profile_kwargs_list=None, alpha_scaling=1, h=2e-3):
    """Second derivatives of the potential as a 4th-order central difference of the restated deflection `alpha`.

    NOT a restatement of the reference's per-profile ``hessian`` functions (LensModel/single_plane.py:142-180 sums
    them; most are closed forms, NIE/SIE a forward difference with step 1e-10, Profiles/nie.py:298-312): the closed
    forms are pinned directly on fixtures produced by the unmodified reference (tests/golden/hessians.npz).  This
    numerical derivative is the independent check that the Hessian IS the derivative of the deflection the oracle
    already pins; it is good to ~1e-9 of the profile's scale where the profile is smooth.
    -> (f_xx, f_xy, f_yx, f_yy)"""
    x = np.asarray(x, dtype=float)
    y = np.asarray(y, dtype=float)

    def a(dx, dy):
        return alpha(lens_model_list, x + dx, y + dy, kwargs, k=k, profile_kwargs_list=profile_kwargs_list,
                     alpha_scaling=alpha_scaling)

    def d(axis):
        s = (h, 0.0) if axis == 0 else (0.0, h)
        p1, m1 = a(*s), a(-s[0], -s[1])
        p2, m2 = a(2 * s[0], 2 * s[1]), a(-2 * s[0], -2 * s[1])
        return [(8.0 * (p1[c] - m1[c]) - (p2[c] - m2[c])) / (12.0 * h) for c in (0, 1)]

    (f_xx, f_yx), (f_xy, f_yy) = d(0), d(1)
    return f_xx, f_xy, f_yx, f_yy


def hessian(lens_model_list, x, y, kwargs, k=None, profile_kwargs_list=None, alpha_scaling=1):
    """SinglePlane.hessian, single_plane.py:142-180: sum of the profiles' Hessians (an int ``k`` returns that profile's
    own, unscaled).  Profiles whose closed form is restated (EPL family, SIS, SHEAR, CONVERGENCE) use it; the others fall
    back to `hessian_differential` of that profile alone."""
    x = np.asarray(x, dtype=float)
    y = np.asarray(y, dtype=float)
    if profile_kwargs_list is None:
        profile_kwargs_list = [{} for _ in lens_model_list]

    def one(i):
        h = lens_profiles.hessian(lens_model_list[i], x, y, profile_kwargs_list[i], **kwargs[i])
        if h is None:
            h = hessian_differential([lens_model_list[i]], x, y, [kwargs[i]], profile_kwargs_list=[profile_kwargs_list[i]],
                                     h=1e-3)
        return h

    if isinstance(k, (int, np.integer)) and not isinstance(k, bool):
        return one(k)
    bl = _bool_list(len(lens_model_list), k)
    out = [np.zeros_like(x) for _ in range(4)]
    for i in range(len(lens_model_list)):
        if bl[i]:
            for o, h in zip(out, one(i)):
                o += h
    return tuple(o * alpha_scaling for o in out)
