"""PSF container (host): lenstronomy.Data.psf.PSF as consumed by jaxtronomy/ImSim/Numerics/numerics.py:78-139."""

import numpy as np

from jaxtronomy_b200.Util import util


class PSF(object):
    def __init__(self, psf_type="NONE", fwhm=None, truncation=5, pixel_size=None, kernel_point_source=None,
                 psf_variance_map=None, point_source_supersampling_factor=1, kernel_point_source_init=None,
                 kernel_point_source_normalisation=True):
        self.psf_type = psf_type
        self._pixel_size = pixel_size
        self.kernel_point_source_init = kernel_point_source_init
        self.point_source_supersampling_factor = point_source_supersampling_factor
        if psf_type == "GAUSSIAN":
            if fwhm is None:
                raise ValueError("fwhm must be set for GAUSSIAN psf type!")
            self._fwhm = fwhm
            self._sigma_gaussian = util.fwhm2sigma(fwhm)
            self._truncation = truncation
        elif psf_type == "PIXEL":
            if kernel_point_source is None:
                raise ValueError("kernel_point_source needs to be specified for PIXEL PSF type!")
            kernel = np.array(kernel_point_source, dtype=float)
            if len(kernel) % 2 == 0:
                raise ValueError("kernel needs to have odd axis number, not ", np.shape(kernel))
            if point_source_supersampling_factor > 1:
                # the given kernel is sampled at `point_source_supersampling_factor` points per pixel
                from jaxtronomy_b200.Util import kernel_util

                self._kernel_point_source_supersampled = kernel / np.sum(kernel) \
                    if kernel_point_source_normalisation else kernel
                kernel = kernel_util.degrade_kernel(kernel, point_source_supersampling_factor)
            if kernel_point_source_normalisation:
                kernel = kernel / np.sum(kernel)
            self._kernel_point_source = kernel
        elif psf_type == "NONE":
            k = np.zeros((3, 3))
            k[1, 1] = 1
            self._kernel_point_source = k
        else:
            raise ValueError("psf_type %s not supported!" % psf_type)
        # variance of the PSF model per kernel pixel, relative to a point source of amplitude 1 (lenstronomy
        # Data/psf.py: same shape as kernel_point_source); feeds ImageModel._error_map_psf
        self.psf_variance_map_bool = psf_variance_map is not None
        self._psf_variance_map = None
        if psf_variance_map is not None:
            if psf_type != "PIXEL":
                raise ValueError("psf_variance_map needs a PIXEL PSF")
            vm = np.array(psf_variance_map, dtype=float)
            if vm.shape != self._kernel_point_source.shape:
                raise ValueError("psf_variance_map must have the shape of the pixel-resolution kernel_point_source")
            self._psf_variance_map = vm

    def _gaussian_kernel(self, supersampling_factor):
        """Pixelised kernel of a GAUSSIAN PSF (point sources only; the extended light uses GaussianConvolution):
        lenstronomy.Data.psf.PSF.kernel_point_source / kernel_point_source_supersampled -- a Gaussian of the given
        fwhm sampled on round(truncation * fwhm / pixel_size * factor) (made odd) points of pixel_size / factor,
        normalised.  lenstronomy is not under /root/reference: restated from its published behaviour, parity
        unpinned."""
        from jaxtronomy_b200.Util import kernel_util

        if self._pixel_size is None:
            raise ValueError("set_pixel_size() must be called before the kernel of a GAUSSIAN PSF is requested")
        n = int(round(self._truncation * self._fwhm / self._pixel_size * supersampling_factor))
        if n % 2 == 0:
            n += 1
        k1 = kernel_util.kernel_gaussian_1d(n, self._sigma_gaussian / (self._pixel_size / supersampling_factor))
        return np.outer(k1, k1)

    @property
    def kernel_point_source(self):
        if self.psf_type == "GAUSSIAN":
            return self._gaussian_kernel(1)
        return self._kernel_point_source

    def kernel_point_source_supersampled(self, supersampling_factor, updata_cache=True):
        """Kernel sampled at `supersampling_factor` points per pixel.  For a PIXEL PSF only available when the PSF
        was given supersampled at that factor: lenstronomy otherwise interpolates the pixel kernel iteratively
        (kernel_util.subgrid_kernel), which is not reproduced here."""
        if self.psf_type == "GAUSSIAN":
            return self._gaussian_kernel(supersampling_factor)
        if self.psf_type == "NONE":
            return self._kernel_point_source
        if supersampling_factor == 1:
            return self._kernel_point_source
        if getattr(self, "_kernel_point_source_supersampled", None) is not None and \
                self.point_source_supersampling_factor == supersampling_factor:
            return self._kernel_point_source_supersampled
        raise NotImplementedError(
            "supersampling_convolution=True needs the PSF at the supersampled resolution: pass "
            "PSF(psf_type='PIXEL', kernel_point_source=<supersampled kernel>, point_source_supersampling_factor=%d)"
            % supersampling_factor)

    @property
    def psf_variance_map(self):
        if self._psf_variance_map is None:
            return np.zeros_like(self.kernel_point_source)
        return self._psf_variance_map

    @property
    def fwhm(self):
        if self.psf_type == "GAUSSIAN":
            return self._fwhm
        raise ValueError("fwhm of a %s PSF is not used on the forward-simulation path" % self.psf_type)

    @property
    def truncation(self):
        return self._truncation

    def set_pixel_size(self, deltaPix):
        self._pixel_size = deltaPix
