"""Build + bind tests/hostmath (the product's host+device profile arithmetic compiled with g++).  Test-only."""

import ctypes as C
import os
import subprocess

import numpy as np

from jaxtronomy_b200 import _lib, _model

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "hostmath", "hostmath.cpp")
SO = os.path.join(HERE, "hostmath", "libhostmath.so")
_DEPS = [SRC] + [os.path.join(HERE, "..", "jaxtronomy_b200", "csrc", f)
                 for f in ("lsim_profiles.cuh", "lsim_math.cuh", "lsim_cse_tables.h", "lsim_grad.cuh")] + \
    [os.path.join(HERE, "..", "include", "lsim.h")]

_dp = C.POINTER(C.c_double)
_h = None


def build():
    if os.path.exists(SO) and all(os.path.getmtime(SO) >= os.path.getmtime(d) for d in _DEPS):
        return SO
    # -ffp-contract=off: explicit fma() calls stay, nothing else is contracted -- the device build
    # contracts more (nvcc -fmad=true), which is exactly the kind of 1-ulp freedom the parity bar allows
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-shared", "-fPIC", "-o", SO, SRC])
    return SO


def lib():
    global _h
    if _h is None:
        h = C.CDLL(build())
        h.hm_ray_shoot.argtypes = [C.POINTER(_lib.LensDesc), C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, _dp,
                                   _dp, _dp, C.c_int64, _dp, _dp, C.c_int, C.c_double]
        h.hm_ray_shoot_epl_shear.argtypes = [C.POINTER(_lib.LensDesc), C.c_double, _dp, _dp, _dp, C.c_int64, _dp, _dp,
                                             C.c_double]
        h.hm_surface_brightness.argtypes = [C.POINTER(_lib.LightDesc), C.c_int, C.c_int, _dp, _dp, _dp, C.c_int64, _dp]
        h.hm_log2.argtypes = [_dp, C.c_int64, _dp]
        h.hm_exp2.argtypes = [_dp, C.c_int64, _dp]
        h.hm_exp2_neg.argtypes = [_dp, C.c_int64, _dp]
        h.hm_exp2_s.argtypes = [_dp, C.c_int64, _dp]
        h.hm_epl_terms.argtypes = [C.POINTER(_lib.LensDesc), _dp, C.c_double]
        h.hm_lens_jvp.argtypes = [C.POINTER(_lib.LensDesc), C.c_int, _dp, _dp, _dp, C.c_int64, _dp, _dp, _dp, _dp]
        h.hm_light_jvp.argtypes = [C.POINTER(_lib.LightDesc), _dp, _dp, _dp, C.c_int64, _dp, _dp]
        h.hm_light_point.argtypes = [C.POINTER(_lib.LightDesc), _dp, _dp, _dp, C.c_int64, _dp, _dp]
        h.hm_lens_vjp.argtypes = [C.POINTER(_lib.LensDesc), C.c_int, _dp, _dp, _dp, C.c_int64, _dp, _dp, _dp, _dp, _dp]
        h.hm_mp_ray_vjp.argtypes = [C.POINTER(_lib.LensDesc), C.c_int, C.c_double, C.c_double, _dp, _dp, _dp, C.c_int64, _dp,
                                    _dp, _dp, _dp, _dp]
        h.hm_eplm_vjp.argtypes = [C.POINTER(_lib.LensDesc), _dp, _dp, _dp, C.c_int64, _dp, _dp, _dp, _dp, _dp]
        h.hm_lens_hessian.argtypes = [C.POINTER(_lib.LensDesc), C.c_int, _dp, _dp, _dp, C.c_int64, _dp]
        h.hm_source_position_logl.argtypes = [C.POINTER(_lib.LensDesc), C.c_int, _dp, _dp, _dp, C.c_int, C.c_double,
                                              C.c_double, C.c_double, _dp]
        _h = h
    return _h


def _p(a):
    return a.ctypes.data_as(_dp)


def ray_shoot(lens_model_list, kwargs_list, x, y, what=0, profile_kwargs_list=None, multi_plane=None,
              alpha_scaling=1.0, series_tol=1e-14):
    """All parameters go through a one-row parameter matrix (columns), like the façade does."""
    names = [_model.LENS_PARAM_NAMES[n] for n in lens_model_list]
    values = _model.gather_values(names, kwargs_list)
    row = np.asarray(values, dtype=np.float64)
    layout = _model.ColumnLayout()
    descs = []
    for i, n in enumerate(lens_model_list):
        src = layout.add_profile("lens", i, names[i])
        extra = {}
        if multi_plane is not None:
            extra = dict(delta_T=multi_plane["delta_T"][i], T_z=multi_plane["T_z"][i], red2phys=multi_plane["red2phys"][i])
        descs.append(_model.lens_desc(n, src, (profile_kwargs_list or [{}] * len(names))[i], **extra))
    arr = (_lib.LensDesc * max(1, len(descs)))(*descs)
    x = np.ascontiguousarray(x, dtype=np.float64)
    y = np.ascontiguousarray(y, dtype=np.float64)
    ox, oy = np.empty_like(x), np.empty_like(y)
    mp = multi_plane is not None
    lib().hm_ray_shoot(arr, len(descs), int(mp), alpha_scaling, multi_plane["T_ij_end"] if mp else 0.0,
                       multi_plane["T_z_source"] if mp else 0.0, _p(row), _p(x), _p(y), x.size, _p(ox), _p(oy), what,
                       series_tol)
    return ox, oy


def ray_shoot_epl_shear(kwargs_list, x, y, alpha_scaling=1.0, series_tol=1e-14):
    """beta through the compile-time TypeList<EPL, SHEAR> path of the specialised image kernels."""
    names = [_model.LENS_PARAM_NAMES[n] for n in ("EPL", "SHEAR")]
    row = np.asarray(_model.gather_values(names, kwargs_list), dtype=np.float64)
    layout = _model.ColumnLayout()
    descs = [_model.lens_desc(n, layout.add_profile("lens", i, names[i]), {}) for i, n in enumerate(("EPL", "SHEAR"))]
    arr = (_lib.LensDesc * 2)(*descs)
    x = np.ascontiguousarray(x, dtype=np.float64)
    y = np.ascontiguousarray(y, dtype=np.float64)
    ox, oy = np.empty_like(x), np.empty_like(y)
    rc = lib().hm_ray_shoot_epl_shear(arr, alpha_scaling, _p(row), _p(x), _p(y), x.size, _p(ox), _p(oy), series_tol)
    assert rc == 0, rc
    return ox, oy


def surface_brightness(light_model_list, kwargs_list, x, y, split=False):
    names = [_model.LIGHT_PARAM_NAMES[n] for n in light_model_list]
    values = _model.gather_values(names, kwargs_list, "light")
    row = np.asarray(values, dtype=np.float64)
    layout = _model.ColumnLayout()
    descs = []
    for i, n in enumerate(light_model_list):
        src = layout.add_profile("light", i, names[i])
        descs.append(_model.light_desc(n, src, lensed=False, slot=i if split else 0, use_amp=not split))
    n_slots = len(descs) if split else 1
    arr = (_lib.LightDesc * max(1, len(descs)))(*descs)
    x = np.ascontiguousarray(x, dtype=np.float64)
    y = np.ascontiguousarray(y, dtype=np.float64)
    out = np.empty((n_slots, x.size))
    lib().hm_surface_brightness(arr, len(descs), n_slots, _p(row), _p(x), _p(y), x.size, _p(out))
    return out


def log2(x):
    x = np.ascontiguousarray(x, dtype=np.float64)
    out = np.empty_like(x)
    lib().hm_log2(_p(x), x.size, _p(out))
    return out


def exp2(x):
    x = np.ascontiguousarray(x, dtype=np.float64)
    out = np.empty_like(x)
    lib().hm_exp2(_p(x), x.size, _p(out))
    return out


def exp2_s(y):
    """exp2 of a pre-scaled argument y = 2048 x (the exp2 table size) with the kernel-parameter coefficients (MathK)."""
    y = np.ascontiguousarray(y, dtype=np.float64)
    out = np.empty_like(y)
    lib().hm_exp2_s(_p(y), y.size, _p(out))
    return out


def exp2_neg(x):
    x = np.ascontiguousarray(x, dtype=np.float64)
    out = np.empty_like(x)
    lib().hm_exp2_neg(_p(x), x.size, _p(out))
    return out


def lens_jvp(lens_model_list, kwargs_list, x, y, profile_kwargs_list=None):
    """Dual-number deflection of the gradient path: (alpha_x, alpha_y, d alpha_x [N, n_lens, 6], d alpha_y)."""
    names = [_model.LENS_PARAM_NAMES[n] for n in lens_model_list]
    row = np.asarray(_model.gather_values(names, kwargs_list), dtype=np.float64)
    layout = _model.ColumnLayout()
    descs = [_model.lens_desc(n, layout.add_profile("lens", i, names[i]), (profile_kwargs_list or [{}] * len(names))[i])
             for i, n in enumerate(lens_model_list)]
    arr = (_lib.LensDesc * max(1, len(descs)))(*descs)
    x = np.ascontiguousarray(x, dtype=np.float64)
    y = np.ascontiguousarray(y, dtype=np.float64)
    ax, ay = np.empty_like(x), np.empty_like(y)
    dax, day = np.empty((x.size, len(descs), 6)), np.empty((x.size, len(descs), 6))
    rc = lib().hm_lens_jvp(arr, len(descs), _p(row), _p(x), _p(y), x.size, _p(ax), _p(ay), _p(dax), _p(day))
    assert rc == 0, rc
    return ax, ay, dax, day


def lens_hessian(lens_model_list, kwargs_list, x, y, profile_kwargs_list=None):
    """k_hessian's per-point code on the host: (f_xx, f_xy, f_yx, f_yy), unscaled."""
    names = [_model.LENS_PARAM_NAMES[n] for n in lens_model_list]
    row = np.asarray(_model.gather_values(names, kwargs_list), dtype=np.float64)
    layout = _model.ColumnLayout()
    descs = [_model.lens_desc(n, layout.add_profile("lens", i, names[i]), (profile_kwargs_list or [{}] * len(names))[i])
             for i, n in enumerate(lens_model_list)]
    arr = (_lib.LensDesc * max(1, len(descs)))(*descs)
    x = np.ascontiguousarray(x, dtype=np.float64)
    y = np.ascontiguousarray(y, dtype=np.float64)
    h = np.empty((4, x.size))
    rc = lib().hm_lens_hessian(arr, len(descs), _p(row), _p(x), _p(y), x.size, _p(h))
    assert rc == 0, rc
    return tuple(h[k].reshape(x.shape) for k in range(4))


def source_position_logl(lens_model_list, kwargs_list, ra, dec, sigma, hard_bound_rms=None, profile_kwargs_list=None):
    """k_source_position's per-set arithmetic on the host -> logL (float)."""
    names = [_model.LENS_PARAM_NAMES[n] for n in lens_model_list]
    row = np.asarray(_model.gather_values(names, kwargs_list), dtype=np.float64)
    layout = _model.ColumnLayout()
    descs = [_model.lens_desc(n, layout.add_profile("lens", i, names[i]), (profile_kwargs_list or [{}] * len(names))[i])
             for i, n in enumerate(lens_model_list)]
    arr = (_lib.LensDesc * max(1, len(descs)))(*descs)
    ra = np.ascontiguousarray(ra, dtype=np.float64)
    dec = np.ascontiguousarray(dec, dtype=np.float64)
    out = np.zeros(1)
    rc = lib().hm_source_position_logl(arr, len(descs), _p(row), _p(ra), _p(dec), ra.size, 1.0, float(sigma),
                                       -1.0 if hard_bound_rms is None else float(hard_bound_rms), _p(out))
    assert rc == 0, rc
    return float(out[0])


def lens_vjp(lens_model_list, kwargs_list, x, y, vx, vy, profile_kwargs_list=None):
    """Hand-derived fast path: (alpha_x, alpha_y, g [N, n_lens, 6]) with g = d(vx alpha_x + vy alpha_y)/d(param)."""
    names = [_model.LENS_PARAM_NAMES[n] for n in lens_model_list]
    row = np.asarray(_model.gather_values(names, kwargs_list), dtype=np.float64)
    layout = _model.ColumnLayout()
    descs = [_model.lens_desc(n, layout.add_profile("lens", i, names[i]), (profile_kwargs_list or [{}] * len(names))[i])
             for i, n in enumerate(lens_model_list)]
    arr = (_lib.LensDesc * max(1, len(descs)))(*descs)
    x, y, vx, vy = (np.ascontiguousarray(a, dtype=np.float64) for a in (x, y, vx, vy))
    ax, ay = np.empty_like(x), np.empty_like(y)
    g = np.empty((x.size, len(descs), 6))
    rc = lib().hm_lens_vjp(arr, len(descs), _p(row), _p(x), _p(y), x.size, _p(vx), _p(vy), _p(ax), _p(ay), _p(g))
    assert rc == 0, rc
    return ax, ay, g


def mp_ray_vjp(lens_model_list, kwargs_list, x, y, vx, vy, multi_plane, profile_kwargs_list=None):
    """Multi-plane ray shooting + reverse sweep of the gradient path: (beta_x, beta_y, g [N, n_lens, 6]) with
    g = d(vx beta_x + vy beta_y)/d(parameter), profiles given in plane order with their (delta_T, T_z, red2phys)."""
    names = [_model.LENS_PARAM_NAMES[n] for n in lens_model_list]
    row = np.asarray(_model.gather_values(names, kwargs_list), dtype=np.float64)
    layout = _model.ColumnLayout()
    descs = [_model.lens_desc(n, layout.add_profile("lens", i, names[i]), (profile_kwargs_list or [{}] * len(names))[i],
                              delta_T=multi_plane["delta_T"][i], T_z=multi_plane["T_z"][i],
                              red2phys=multi_plane["red2phys"][i]) for i, n in enumerate(lens_model_list)]
    arr = (_lib.LensDesc * len(descs))(*descs)
    x, y, vx, vy = (np.ascontiguousarray(a, dtype=np.float64) for a in (x, y, vx, vy))
    bx, by = np.empty_like(x), np.empty_like(y)
    g = np.zeros((x.size, len(descs), 6))
    rc = lib().hm_mp_ray_vjp(arr, len(descs), multi_plane["T_ij_end"], multi_plane["T_z_source"], _p(row), _p(x), _p(y),
                             x.size, _p(vx), _p(vy), _p(bx), _p(by), _p(g))
    assert rc == 0, rc
    return bx, by, g


def light_jvp(light_model, kwargs, x, y, use_amp=True, fast=False):
    """Dual-number surface brightness: (I [N], dI [N, 9]) with dI[:, 0:2] = d/dx, d/dy and dI[:, 2 + k] = d/d(param k)."""
    names = [_model.LIGHT_PARAM_NAMES[light_model]]
    row = np.asarray(_model.gather_values(names, [kwargs], "light"), dtype=np.float64)
    layout = _model.ColumnLayout()
    desc = _model.light_desc(light_model, layout.add_profile("light", 0, names[0]), lensed=True, slot=0, use_amp=use_amp)
    x = np.ascontiguousarray(x, dtype=np.float64)
    y = np.ascontiguousarray(y, dtype=np.float64)
    out, dout = np.empty_like(x), np.empty((x.size, 9))
    fn = lib().hm_light_point if fast else lib().hm_light_jvp
    rc = fn(C.byref(desc), _p(row), _p(x), _p(y), x.size, _p(out), _p(dout))
    assert rc == 0, rc
    return out, dout


def eplm_vjp(kwargs, x, y, vx, vy):
    """EPL_MULTIPOLE_M1M3M4 through the gradient path's expansion: (alpha_x, alpha_y, g [N, 12])."""
    names = [_model.LENS_PARAM_NAMES["EPL_MULTIPOLE_M1M3M4"]]
    row = np.asarray(_model.gather_values(names, [kwargs]), dtype=np.float64)
    layout = _model.ColumnLayout()
    desc = _model.lens_desc("EPL_MULTIPOLE_M1M3M4", layout.add_profile("lens", 0, names[0]), {})
    x, y, vx, vy = (np.ascontiguousarray(a, dtype=np.float64) for a in (x, y, vx, vy))
    ax, ay = np.empty_like(x), np.empty_like(y)
    g = np.empty((x.size, 12))
    rc = lib().hm_eplm_vjp(C.byref(desc), _p(row), _p(x), _p(y), x.size, _p(vx), _p(vy), _p(ax), _p(ay), _p(g))
    assert rc == 0, rc
    return ax, ay, g
