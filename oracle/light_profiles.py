"""Oracle: surface brightness of the hot-path light profiles and the LightModel sum/split.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Follows
  jaxtronomy/LightModel/Profiles/sersic_ellipse.py:61-93   SersicElliptic.function
  jaxtronomy/LightModel/Profiles/sersic.py:54-76           Sersic.function
  jaxtronomy/LensModel/Profiles/sersic_utils.py:67-110,238-267  b_n, get_distance_from_center, _r_sersic
  jaxtronomy/LightModel/Profiles/gaussian.py:27-42         Gaussian.function
  jaxtronomy/LightModel/light_model_base.py:230-252        LightModelBase.surface_brightness
  jaxtronomy/LightModel/linear_basis.py:24-91,161-219      functions_split, update_linear
"""

import numpy as np

from .util import nan_to_num, transform_e1e2_product_average

PARAM_NAMES = {
    "SERSIC_ELLIPSE": ["amp", "R_sersic", "n_sersic", "e1", "e2", "center_x", "center_y"],
    "SERSIC": ["amp", "R_sersic", "n_sersic", "center_x", "center_y"],
    "GAUSSIAN": ["amp", "sigma", "center_x", "center_y"],
}

_SMOOTHING = 0.0001  # sersic_utils.py:13


def _r_sersic(R, R_sersic, n_sersic, max_R_frac=1000.0):
    """SersicUtil._r_sersic, sersic_utils.py:247-267."""
    R_ = np.maximum(_SMOOTHING, R)
    R_sersic_ = np.maximum(_SMOOTHING, R_sersic)
    bn = np.maximum(1.9992 * n_sersic - 0.3271, 0.00001)
    R_frac = R_ / R_sersic_
    with np.errstate(all="ignore"):
        exponent = -bn * (R_frac ** (1.0 / n_sersic) - 1.0)
        result = np.where(R_frac <= max_R_frac, np.exp(exponent), 0)
    return nan_to_num(result)


def sersic_ellipse(x, y, amp, R_sersic, n_sersic, e1, e2, center_x=0, center_y=0, max_R_frac=1000.0):
    R_sersic = np.maximum(0, R_sersic)
    x_, y_ = transform_e1e2_product_average(x, y, e1, e2, center_x, center_y)
    R = np.sqrt(x_**2 + y_**2)
    return amp * _r_sersic(R, R_sersic, n_sersic, max_R_frac)


def sersic(x, y, amp, R_sersic, n_sersic, center_x=0, center_y=0, max_R_frac=1000.0):
    x_, y_ = transform_e1e2_product_average(x, y, 0, 0, center_x, center_y)
    R = np.sqrt(x_**2 + y_**2)
    return amp * _r_sersic(R, R_sersic, n_sersic, max_R_frac)


def gaussian(x, y, amp, sigma, center_x=0, center_y=0):
    x = np.asarray(x, dtype=float)
    y = np.asarray(y, dtype=float)
    c = amp / (2 * np.pi * sigma**2)
    r2 = (x - center_x) ** 2 / sigma**2 + (y - center_y) ** 2 / sigma**2
    return c * np.exp(-r2 / 2.0)


_FUNCS = {"SERSIC_ELLIPSE": sersic_ellipse, "SERSIC": sersic, "GAUSSIAN": gaussian}


def _bool_list(n, k=None):
    """lenstronomy.Util.util.convert_bool_list."""
    if k is None:
        return [True] * n
    if isinstance(k, (int, np.integer)) and not isinstance(k, bool):
        out = [False] * n
        out[k] = True
        return out
    k = list(k)
    if len(k) > 0 and isinstance(k[0], (bool, np.bool_)):
        if len(k) != n:
            raise ValueError("length of selected lens models in format of boolean list is %s and does not match the models of this class instance %s." % (len(k), n))
        return [bool(v) for v in k]
    out = [False] * n
    for i in k:
        out[i] = True
    return out


def surface_brightness(light_model_list, x, y, kwargs_list, k=None):
    x = np.array(x, dtype=float)
    y = np.array(y, dtype=float)
    flux = np.zeros_like(x)
    bl = _bool_list(len(light_model_list), k)
    for i, name in enumerate(light_model_list):
        if bl[i]:
            flux = flux + _FUNCS[name](x, y, **kwargs_list[i])
    return flux


def functions_split(light_model_list, x, y, kwargs_list, k=None):
    """Basis images with amp=1 (linear_basis.py:24-91)."""
    response = []
    bl = _bool_list(len(light_model_list), k)
    for i, name in enumerate(light_model_list):
        if bl[i]:
            kw = dict(kwargs_list[i])
            kw["amp"] = 1
            response.append(_FUNCS[name](x, y, **kw))
    return response, len(response)
