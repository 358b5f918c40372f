"""Ellipticity conversions on the host (NumPy), jaxtronomy/Util/param_util.py:6-99."""

import numpy as np


def cart2polar(x, y, center_x=0, center_y=0):
    xs, ys = np.asarray(x, dtype=float) - center_x, np.asarray(y, dtype=float) - center_y
    return np.sqrt(xs**2 + ys**2), np.arctan2(ys, xs)


def polar2cart(r, phi, center):
    return r * np.cos(phi) - center[0], r * np.sin(phi) - center[1]


def phi_q2_ellipticity(phi, q):
    e1 = (1.0 - q) / (1.0 + q) * np.cos(2 * phi)
    e2 = (1.0 - q) / (1.0 + q) * np.sin(2 * phi)
    return e1, e2


def ellipticity2phi_q(e1, e2):
    phi = np.arctan2(e2, e1) / 2
    c = np.minimum(np.sqrt(e1**2 + e2**2), 0.9999)
    return phi, (1 - c) / (1 + c)


def shear_polar2cartesian(phi, gamma):
    return gamma * np.cos(2 * phi), gamma * np.sin(2 * phi)


def shear_cartesian2polar(gamma1, gamma2):
    return np.arctan2(gamma2, gamma1) / 2, np.sqrt(gamma1**2 + gamma2**2)
