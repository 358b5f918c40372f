"""Oracle: PositionLikelihood (source-position penalty, image-position and astrometric terms).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Follows
  jaxtronomy/Sampling/Likelihoods/position_likelihood.py:178-212  astrometric_likelihood
  jaxtronomy/Sampling/Likelihoods/position_likelihood.py:214-247  image_position_likelihood
  jaxtronomy/Sampling/Likelihoods/position_likelihood.py:249-283  source_position_likelihood
  jaxtronomy/Sampling/Likelihoods/position_likelihood.py:296-352  image2source_covariance, _compute_penalty
  jaxtronomy/PointSource/Types/lensed_position.py:65-90           source position = mean of the ray-shot images

Pinned on tests/golden/position_likelihood.npz (outputs of the unmodified reference).  The Hessian inside is
oracle/lens_model.py: hessian -- the restated closed forms for the EPL family, SIS, SHEAR and CONVERGENCE (1e-10 against
the fixtures), the 4th-order difference of the restated deflection for the other profiles (~1e-8; of the fixture cases
only quad_m1m3m4 needs it).
"""

import numpy as np

from . import lens_model


def compute_penalty(f_xx, f_xy, f_yx, f_yy, x_source_avg, y_source_avg, x_source, y_source, sigma, hard_bound_rms):
    """One image (scalars) -> penalty."""
    A = np.array([[1 - f_xx, -f_xy], [-f_yx, 1 - f_yy]], dtype=float)
    Sigma_theta = np.array([[1, 0], [0, 1]], dtype=float) * sigma**2
    Sigma_beta = np.matmul(np.matmul(A.T, Sigma_theta), A)
    delta = np.array([x_source_avg - x_source, y_source_avg - y_source], dtype=float)
    a, b, c, d = Sigma_beta[0][0], Sigma_beta[0][1], Sigma_beta[1][0], Sigma_beta[1][1]
    det = a * d - b * c
    Sigma_inv = np.array([[d, -b], [-c, a]])
    penalty = 10.0**15 if det == 0 else delta.T.dot(Sigma_inv.dot(delta)) / (2 * det)
    if hard_bound_rms is not None and delta[0] ** 2 + delta[1] ** 2 > hard_bound_rms**2:
        penalty = penalty + 10**3
    return penalty


def source_position_likelihood(lens_model_list, kwargs_lens, point_source_type_list, kwargs_ps, sigma,
                               hard_bound_rms=None, hessian=None):
    logL = 0.0
    for k, kw in enumerate(kwargs_ps):
        if "ra_image" not in kw or point_source_type_list[k] != "LENSED_POSITION":
            continue
        x, y = np.asarray(kw["ra_image"], dtype=float), np.asarray(kw["dec_image"], dtype=float)
        xs, ys = lens_model.ray_shooting(lens_model_list, x, y, kwargs_lens)
        h = hessian(x, y) if hessian is not None else lens_model.hessian(lens_model_list, x, y, kwargs_lens)
        for i in range(x.size):
            logL -= compute_penalty(h[0][i], h[1][i], h[2][i], h[3][i], np.mean(xs), np.mean(ys), xs[i], ys[i], sigma,
                                    hard_bound_rms)
    return logL


def image_position_likelihood(kwargs_ps, ra_image_list, dec_image_list, sigma):
    logL = 0.0
    for i, kw in enumerate(kwargs_ps):
        ra, dec = np.atleast_1d(np.asarray(kw["ra_image"], dtype=float)), np.atleast_1d(np.asarray(kw["dec_image"], dtype=float))
        n = min(len(ra_image_list[i]), len(ra))
        logL += -np.sum(((ra[:n] - np.asarray(ra_image_list[i], dtype=float)[:n]) ** 2
                         + (dec[:n] - np.asarray(dec_image_list[i], dtype=float)[:n]) ** 2) / sigma**2 / 2)
    return logL


def astrometric_likelihood(kwargs_ps, kwargs_special, sigma):
    if len(kwargs_ps) == 0 or "ra_image" not in kwargs_ps[0] or "delta_x_image" not in kwargs_special:
        return 0.0
    dx, dy = np.asarray(kwargs_special["delta_x_image"], dtype=float), np.asarray(kwargs_special["delta_y_image"], dtype=float)
    return float(np.nan_to_num(-np.sum((dx**2 + dy**2) / sigma**2 / 2), nan=-(10.0**15)))
