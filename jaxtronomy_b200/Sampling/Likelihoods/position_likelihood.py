"""PositionLikelihood: likelihood of the positions of multiply imaged point sources,
jaxtronomy/Sampling/Likelihoods/position_likelihood.py:10-283 (penalty: _compute_penalty, :296-352).

* source-position likelihood: for every LENSED_POSITION model the image positions are ray-shot to the source plane,
  the spread around their mean is weighted with the image-plane uncertainty carried through the lensing Jacobian
  A = 1 - Hessian (Birrer & Treu 2019, eq. 13).  ONE kernel launch per point-source model (`lsim_source_position_logl`,
  k_source_position): deflection and Hessian come from the same dual-number evaluation, the penalty and the sum over
  images are formed in the kernel -- batched over parameter sets, nothing but logL[B] leaves the GPU.
* image-position likelihood / astrometric likelihood: quadratic forms of the sampled positions themselves (no lensing);
  plain array arithmetic on the caller's arrays (NumPy in -> NumPy out, torch in -> torch out).

Like the reference: `force_no_add_image` and `restrict_image_number` raise.  Point-source frames (k_list) are outside the
scope of jaxtronomy_b200 (PointSource refuses them at construction)."""

import warnings

import numpy as np

from jaxtronomy_b200 import _arrays, _model

__all__ = ["PositionLikelihood"]


class PositionLikelihood(object):
    def __init__(self, point_source_class, image_position_uncertainty=0.005, astrometric_likelihood=False,
                 image_position_likelihood=False, ra_image_list=None, dec_image_list=None,
                 source_position_likelihood=None, source_position_tolerance=None, source_position_sigma=0.001,
                 force_no_add_image=False, restrict_image_number=False, max_num_images=None):
        self._pointSource = point_source_class
        self._lensModel = point_source_class._lens_model
        types = point_source_class.point_source_type_list
        if source_position_likelihood is None:
            source_position_likelihood = "LENSED_POSITION" in types
        if force_no_add_image:
            raise ValueError("force_no_add_image is not supported in jaxtronomy yet")
        if restrict_image_number:
            raise ValueError("restrict_image_number is not supported in jaxtronomy yet")
        self._max_num_images = max_num_images
        self._astrometric_likelihood = astrometric_likelihood
        self._source_position_likelihood = source_position_likelihood
        self._source_position_sigma = source_position_sigma
        self._bound_source_position_tolerance = source_position_tolerance
        if source_position_tolerance is not None and source_position_likelihood is False:
            warnings.warn("source_position_tolerance has been set but source_position_likelihood is False. \n"
                          "In order to use the source_position_tolerance, set source_position_likelihood to True")
        self._image_position_likelihood = image_position_likelihood
        self._image_position_sigma = image_position_uncertainty
        self._ra_image_list = [np.asarray(a, dtype=float) for a in (ra_image_list or [])]
        self._dec_image_list = [np.asarray(a, dtype=float) for a in (dec_image_list or [])]

    # -- the reference's kwargs-level API -------------------------------------------------------------------------
    def logL(self, kwargs_lens, kwargs_ps, kwargs_special=None, verbose=False):
        """position_likelihood.py:100-176"""
        logL = 0
        if self._astrometric_likelihood is True:
            logL = logL + self.astrometric_likelihood(kwargs_ps, kwargs_special or {}, self._image_position_sigma)
        if self._source_position_likelihood:
            logL = logL + self.source_position_likelihood(kwargs_lens, kwargs_ps, self._source_position_sigma,
                                                          hard_bound_rms=self._bound_source_position_tolerance)
        if self._image_position_likelihood is True:
            logL = logL + self.image_position_likelihood(kwargs_ps, kwargs_lens, self._image_position_sigma)
        return logL

    @staticmethod
    def astrometric_likelihood(kwargs_ps, kwargs_special, sigma):
        """position_likelihood.py:178-212: -sum (delta_x^2 + delta_y^2) / (2 sigma^2) of the astrometric offsets."""
        if len(kwargs_ps) == 0 or "ra_image" not in kwargs_ps[0] or "delta_x_image" not in kwargs_special:
            return 0
        dx, dy = kwargs_special["delta_x_image"], kwargs_special["delta_y_image"]
        if not _arrays.is_torch(dx):
            dx, dy = np.asarray(dx, dtype=float), np.asarray(dy, dtype=float)
        logL = -((dx**2 + dy**2) / sigma**2 / 2).sum(-1)
        if _arrays.is_torch(logL):
            return logL.nan_to_num(nan=-1e15)
        return np.nan_to_num(logL, nan=-(10.0**15))

    def image_position_likelihood(self, kwargs_ps, kwargs_lens, sigma):
        """position_likelihood.py:214-247: model image positions against the measured ones (ra_image_list)."""
        logL = 0
        for i, kw in enumerate(kwargs_ps):
            ra, dec = kw["ra_image"], kw["dec_image"]
            if not _arrays.is_torch(ra):
                ra, dec = np.atleast_1d(np.asarray(ra, dtype=float)), np.atleast_1d(np.asarray(dec, dtype=float))
            n = min(len(self._ra_image_list[i]), ra.shape[-1])
            ra0, dec0 = self._ra_image_list[i][:n], self._dec_image_list[i][:n]
            if _arrays.is_torch(ra):
                import torch

                ra0, dec0 = (torch.as_tensor(a, dtype=ra.dtype, device=ra.device) for a in (ra0, dec0))
            logL = logL - (((ra[..., :n] - ra0) ** 2 + (dec[..., :n] - dec0) ** 2) / sigma**2 / 2).sum(-1)
        return logL

    def source_position_likelihood(self, kwargs_lens, kwargs_ps, sigma, hard_bound_rms=None, verbose=False):
        """position_likelihood.py:249-283.  Parameters given as ``[B]`` arrays (image positions: ``[B, n_img]``) make
        the call batched: one launch per LENSED_POSITION model over all B sets."""
        if len(kwargs_ps) < 1:
            return 0
        lm = self._lensModel
        if lm is None:
            raise ValueError("the source-position likelihood needs the lens model: build PointSource with lens_model=...")
        values = _model.gather_values(lm.param_name_list, kwargs_lens, "lens")
        mat, B, batched = _model.pack_columns(values)
        plan = lm._plan(None)
        total, want_torch = None, False
        for k, kw in enumerate(kwargs_ps):
            if "ra_image" not in kw or self._pointSource.point_source_type_list[k] != "LENSED_POSITION":
                continue
            ra, dec = kw["ra_image"], kw["dec_image"]
            want_torch = want_torch or _arrays.is_torch(ra)
            if not _arrays.is_torch(ra):
                ra, dec = np.atleast_1d(np.asarray(ra, dtype=np.float64)), np.atleast_1d(np.asarray(dec, dtype=np.float64))
            if ra.ndim == 2 and not batched:  # batched image positions with scalar lens parameters
                mat, B, batched = np.repeat(np.asarray(mat), ra.shape[0], axis=0), ra.shape[0], True
            l = plan.source_position_logl(mat, ra, dec, sigma, hard_bound_rms)
            total = l if total is None else total + l
        if total is None:
            return 0
        if not batched:
            total = total[0]
        return total if want_torch else (total.cpu().numpy() if batched else float(total))

    @property
    def num_data(self):
        """position_likelihood.py:285-296"""
        num = 0
        if self._image_position_likelihood is True:
            for a in self._ra_image_list:
                num += len(a) * 2
        return num
