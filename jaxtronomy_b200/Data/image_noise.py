"""Diagonal noise model (host side), jaxtronomy/Data/image_noise.py:16-137."""

import numpy as np


def covariance_matrix(data, background_rms, exposure_map):
    """image_noise.py:113-137: max(data, 0) / t_exp + sigma_bkg^2."""
    d_pos = np.where(data >= 0, data, 0)
    return d_pos / exposure_map + background_rms**2


class ImageNoise(object):
    def __init__(self, image_data, exposure_time=None, background_rms=None, noise_map=None, flux_scaling=1,
                 gradient_boost_factor=None, verbose=True):
        if exposure_time is None:
            if noise_map is None:
                raise ValueError("Exposure map has not been specified in Noise() class!")
        else:
            exposure_time = np.asarray(exposure_time, dtype=float)
            self.exposure_map = np.where(exposure_time <= 10 ** (-10), 10 ** (-10), exposure_time)
        if background_rms is None:
            if noise_map is None:
                raise ValueError("rms background value as 'background_rms' not specified!")
            self.background_rms = np.median(noise_map)
        else:
            self.background_rms = background_rms
        self.data = np.array(image_data, dtype=float)
        self.flux_scaling = flux_scaling
        if noise_map is not None:
            assert np.shape(noise_map) == np.shape(image_data)
            self._noise_map = np.array(noise_map, dtype=float)
        else:
            self._noise_map = None
        if self._noise_map is not None:
            self.C_D = self._noise_map**2
        else:
            self.C_D = covariance_matrix(self.data, self.background_rms, self.exposure_map)
        if gradient_boost_factor is not None:
            raise ValueError("gradient_boost_factor not supported in JAXtronomy. Please use lenstronomy instead")
