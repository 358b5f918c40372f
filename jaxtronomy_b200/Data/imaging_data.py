"""ImageData (host container of the static data / noise configuration), jaxtronomy/Data/imaging_data.py:11-220.
The likelihood itself is evaluated on the GPU by lsim_logl (k_likelihood)."""

import numpy as np

from jaxtronomy_b200.Data.image_noise import ImageNoise
from jaxtronomy_b200.Data.pixel_grid import PixelGrid

__all__ = ["ImageData"]


class ImageData(PixelGrid, ImageNoise):
    def __init__(self, image_data, exposure_time=None, background_rms=None, noise_map=None, gradient_boost_factor=None,
                 ra_at_xy_0=0, dec_at_xy_0=0, transform_pix2angle=None, ra_shift=0, dec_shift=0, phi_rot=0,
                 log_likelihood_constant=0, antenna_primary_beam=None, likelihood_method="diagonal", flux_scaling=1):
        if antenna_primary_beam is not None:
            raise ValueError("primary beam is not supported in jaxtronomy")
        nx, ny = np.shape(image_data)
        if transform_pix2angle is None:
            transform_pix2angle = np.array([[1, 0], [0, 1]])
        cos_phi, sin_phi = np.cos(phi_rot), np.sin(phi_rot)
        rot = np.array([[cos_phi, -sin_phi], [sin_phi, cos_phi]])
        PixelGrid.__init__(self, nx, ny, np.dot(transform_pix2angle, rot), ra_at_xy_0 + ra_shift,
                           dec_at_xy_0 + dec_shift, antenna_primary_beam)
        ImageNoise.__init__(self, image_data, exposure_time=exposure_time, background_rms=background_rms,
                            noise_map=noise_map, gradient_boost_factor=gradient_boost_factor, verbose=False,
                            flux_scaling=flux_scaling)
        self._logL_constant = log_likelihood_constant
        if likelihood_method == "interferometry_natwt":
            raise ValueError("interferometry_natwt is outside the scope of jaxtronomy_b200 "
                             "(ImageLinearFit rejects it in the reference as well, image_linear_solve.py:89-92)")
        if likelihood_method != "diagonal":
            raise ValueError("likelihood_method %s not supported! likelihood_method can only be 'diagonal' or "
                             "'interferometry_natwt'!" % likelihood_method)
        self._logL_method = likelihood_method

    def update_data(self, image_data):
        raise ValueError("Updating data class not supported. Create a new instance of ImageData / ImageModel instead.")

    def log_likelihood(self, model, mask, additional_error_map=0):
        """The chi^2 of imaging_data.py:159-182 is evaluated on the GPU, fused with the image simulation
        (lsim_logl / k_likelihood); there is deliberately no host implementation."""
        raise NotImplementedError("use ImageModel.likelihood_data_given_model / ImageLinearFit."
                                  "likelihood_data_given_model: the likelihood runs on the GPU (lsim_logl)")

    def likelihood_method(self):
        return self._logL_method
