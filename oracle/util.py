"""Oracle: coordinate helpers and the Lopez-Temme 2F1 series.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Follows
  jaxtronomy/Util/util.py:102-139        rotate, shift_center
  jaxtronomy/Util/util.py:18-37,72-82    array2image, image2array
  jaxtronomy/Util/param_util.py:6-24     cart2polar
  jaxtronomy/Util/param_util.py:73-99    phi_q2_ellipticity, ellipticity2phi_q
  jaxtronomy/Util/param_util.py:102-123  transform_e1e2_product_average
  jaxtronomy/Util/hyp2f1_util.py:135-176 hyp2f1_lopez_temme_8
  jaxtronomy/Util/image_util.py:78-98    re_size
"""

import numpy as np


def shift_center(x, y, center_x=0.0, center_y=0.0):
    x = np.asarray(x, dtype=float)
    y = np.asarray(y, dtype=float)
    return x - center_x, y - center_y


def rotate(x, y, angle):
    c, s = np.cos(angle), np.sin(angle)
    return x * c + y * s, -x * s + y * c


def cart2polar(x, y, center_x=0.0, center_y=0.0):
    xs, ys = shift_center(x, y, center_x, center_y)
    return np.sqrt(xs**2 + ys**2), np.arctan2(ys, xs)


def phi_q2_ellipticity(phi, q):
    e1 = (1.0 - q) / (1.0 + q) * np.cos(2 * phi)
    e2 = (1.0 - q) / (1.0 + q) * np.sin(2 * phi)
    return e1, e2


def ellipticity2phi_q(e1, e2):
    phi = np.arctan2(e2, e1) / 2
    c = np.minimum(np.sqrt(e1**2 + e2**2), 0.9999)
    q = (1 - c) / (1 + c)
    return phi, q


def transform_e1e2_product_average(x, y, e1, e2, center_x, center_y):
    xs, ys = shift_center(x, y, center_x, center_y)
    norm = np.maximum(np.sqrt(np.abs(1 - e1**2 - e2**2)), 0.000001)
    x_ = ((1 - e1) * xs - e2 * ys) / norm
    y_ = (-e2 * xs + (1 + e1) * ys) / norm
    return x_, y_


def nan_to_num(a, posinf=None, neginf=None):
    """jnp.nan_to_num: nan -> 0, +-inf -> given (default: largest finite)."""
    return np.nan_to_num(a, nan=0.0, posinf=posinf, neginf=neginf)


def array2image(array, nx=0, ny=0):
    array = np.asarray(array)
    if nx == 0 or ny == 0:
        n = int(np.sqrt(array.size))
        if n * n != array.size:
            raise ValueError("length of input array given as %s is not square of integer number!" % array.size)
        nx, ny = n, n
    return array.reshape(ny, nx)


def image2array(image):
    image = np.asarray(image)
    return image.reshape(image.size)


def re_size(image, factor=1):
    if factor < 1:
        raise ValueError("scaling factor in re-sizing %s < 1" % factor)
    if factor == 1:
        return image
    f = int(factor)
    nx, ny = image.shape
    if nx % f or ny % f:
        raise ValueError("scaling with factor %s is not possible with grid size %s, %s" % (f, nx, ny))
    return image.reshape(nx // f, f, ny // f, f).mean(3).mean(1)


def hyp2f1_lopez_temme_8(a, b, c, z, nmax=75):
    """2F1(a,b;c;z) by eq. 8 of Lopez & Temme (2013), truncated at nmax terms.

    Loop structure and operation order as hyp2f1_util.py:152-176.
    """
    z = np.asarray(z, dtype=complex)
    w = z / (z - 2.0)
    pref = a * w
    ppF = 1.0
    pF = 1.0 - 2.0 * b / c
    s = 1 + pref * pF
    for n in range(2, nmax):
        pref = pref * ((a + n - 1.0) / n * w)
        nF = ((n - 1.0) * ppF - (2.0 * b - c) * pF) / (c + n - 1.0)
        s = s + pref * nF
        ppF, pF = pF, nF
    return (1.0 - z / 2.0) ** (-a) * s


def epl_nmax(f):
    """nmax schedule of EPLMajorAxis._hyp2f1_evaluate (epl.py:308-319).  0 means 'return 1'."""
    if f == 0:
        return 0
    for thr, n in ((0.12, 10), (0.4, 20), (0.6, 35), (0.73, 50), (0.86, 100), (0.92, 200)):
        if f < thr:
            return n
    return 500
