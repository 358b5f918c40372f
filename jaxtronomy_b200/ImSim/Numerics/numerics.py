"""Numerics: wiring of grid, re-binning and PSF convolution, jaxtronomy/ImSim/Numerics/numerics.py:24-229 and
convolution.py.  On the host this class only *describes* the numerics (grid arrays, convolution mode, kernel);
the arithmetic happens in the CUDA kernels (k_image / k_rebin / k_conv_*).

Mapping of the reference's convolution classes onto kernel modes:
  PixelKernelConvolution  ("fft", "fft_static" and "grid" alike)  -> LSIM_CONV_ZERO, direct "same" convolution
  GaussianConvolution                                            -> LSIM_CONV_EDGE with the exact [2r+1]^2 kernel;
                                                                    the kernel is outer(k1, k1), so two 1-d passes
  psf_type NONE / unconvolved=True                                -> LSIM_CONV_NONE
  SubgridKernelConvolution / Gaussian with supersampling_convolution=True
                                                                  -> kernel_hi on the supersampled grid, re-bin,
                                                                     + low-resolution "hole" kernel on the pixel grid
Not built: flux_evaluate_indexes sub-frames, compute_mode="adaptive" (not in the reference either).
"""

import numpy as np

from jaxtronomy_b200 import _lib
from jaxtronomy_b200.ImSim.Numerics.grid import RegularGrid
from jaxtronomy_b200.Util import kernel_util, util

__all__ = ["Numerics", "NumericsSubFrame"]


def gaussian_pixel_kernel(sigma_scaled, num_pix):
    """GaussianConvolution.pixel_kernel (convolution.py:254-273): normalised Gaussian.function on make_grid(n, 1)."""
    if num_pix % 2 == 0:
        raise ValueError("num_pix must be an odd integer")
    if num_pix < 3:
        raise ValueError("psf kernel size must be 3 or greater")
    x, y = util.make_grid(numPix=num_pix, deltapix=1)
    c = 1.0 / (2 * np.pi * sigma_scaled**2)
    r2 = x**2 / sigma_scaled**2 + y**2 / sigma_scaled**2
    kernel = util.array2image(c * np.exp(-r2 / 2.0))
    return kernel / np.sum(kernel)


class Numerics(object):
    def __init__(self, pixel_grid, psf, supersampling_factor=1, compute_mode="regular",
                 supersampling_convolution=False, supersampling_kernel_size=5, flux_evaluate_indexes=None,
                 supersampled_indexes=None, compute_indexes=None, point_source_supersampling_factor=1,
                 convolution_kernel_size=None, convolution_type="fft", truncation_conv=None):
        if compute_mode != "regular":
            if compute_mode == "adaptive":
                raise ValueError("AdaptiveConvolution not implemented in Jaxtronomy. Please use lenstronomy instead.")
            raise ValueError('compute_mode specified as %s not valid. Options are "regular" and "adaptive" '
                             "(adaptive only supported in lenstronomy)" % compute_mode)
        self._psf_type = psf.psf_type
        if not isinstance(supersampling_factor, int):
            raise TypeError("supersampling_factor needs to be an integer! Current type is %s"
                            % type(supersampling_factor))
        if supersampling_factor == 1:
            supersampling_convolution = False
        if convolution_type not in ("fft", "fft_static", "grid"):
            raise ValueError("convolution_type %s not supported!" % convolution_type)
        self._pixel_width = pixel_grid.pixel_width
        # PointSourceRendering(pixel_grid, supersampling_factor=point_source_supersampling_factor, psf)
        # (numerics.py:141-147, point_source_rendering.py:22-24)
        self._ps_supersampling = psf.point_source_supersampling_factor if point_source_supersampling_factor is None \
            else int(point_source_supersampling_factor)
        nx, ny = pixel_grid.num_pixel_axes
        ra0, dec0 = pixel_grid.radec_at_xy_0
        self._nx, self._ny = nx, ny
        self._grid = RegularGrid(nx, ny, pixel_grid.transform_pix2angle, ra0, dec0, supersampling_factor,
                                 flux_evaluate_indexes)
        self._eval_mask = None
        if flux_evaluate_indexes is not None and not bool(np.all(flux_evaluate_indexes)):
            self._eval_mask = util.image2array(np.asarray(flux_evaluate_indexes)).astype(bool)
        self._kernel = None
        self._kernel_1d = None
        self._kernel_hi = None     # supersampled convolution: kernel on the [ny ss, nx ss] grid
        self._kernel_hi_1d = None
        self._high_res_return = bool(supersampling_convolution)
        if self._psf_type == "PIXEL" and supersampling_convolution is True:
            # SubgridKernelConvolution (convolution.py:99-166)
            kernel_super = psf.kernel_point_source_supersampled(supersampling_factor)
            if convolution_kernel_size is not None:
                kernel_super = self._supersampling_cut_kernel(kernel_super, convolution_kernel_size,
                                                              supersampling_factor)
            self._conv_mode = _lib.CONV_ZERO
            if supersampling_kernel_size is None:
                self._kernel_hi = np.array(kernel_super, dtype=float)
            else:
                low, high = kernel_util.split_kernel(kernel_super, supersampling_kernel_size, supersampling_factor)
                self._kernel, self._kernel_hi = np.array(low, dtype=float), np.array(high, dtype=float)
        elif self._psf_type == "PIXEL":
            kernel = psf.kernel_point_source
            kernel = self._supersampling_cut_kernel(kernel, convolution_kernel_size, supersampling_factor=1)
            if kernel.shape[0] != kernel.shape[1] or kernel.shape[0] % 2 == 0:
                raise ValueError("PSF kernel must be square with an odd side, got %s" % (kernel.shape,))
            self._conv_mode, self._kernel = _lib.CONV_ZERO, np.array(kernel, dtype=float)
        elif self._psf_type == "GAUSSIAN":
            sigma = util.fwhm2sigma(psf.fwhm)
            if truncation_conv is None:
                truncation_conv = psf.truncation
            sigma_scaled = sigma / pixel_grid.pixel_width
            if supersampling_convolution is True:
                sigma_scaled *= supersampling_factor  # convolution.py:199-200
            kernel_radius = max(round(sigma_scaled * truncation_conv), 1)
            num_pix = 2 * int(kernel_radius) + 1
            self._conv_mode = _lib.CONV_EDGE
            k2d = gaussian_pixel_kernel(sigma_scaled, num_pix)
            k1d = kernel_util.kernel_gaussian_1d(num_pix, sigma_scaled)
            if supersampling_convolution is True:
                self._kernel_hi, self._kernel_hi_1d = k2d, k1d
            else:
                self._kernel, self._kernel_1d = k2d, k1d
        elif self._psf_type == "NONE":
            self._conv_mode = _lib.CONV_NONE
        else:
            raise ValueError("psf_type %s not valid! Chose either NONE, GAUSSIAN or PIXEL." % self._psf_type)
        self._plans = {}

    # -- descriptors for _lib.Plan -----------------------------------------------------------------------
    def grid_dict(self):
        xs, ys = self._grid.coordinates_full
        return dict(nx=self._nx, ny=self._ny, ss=self._grid.supersampling_factor, x_sub=xs, y_sub=ys,
                    eval_mask=self._eval_mask)

    def psf_dict(self, unconvolved=False):
        if unconvolved or self._conv_mode == _lib.CONV_NONE:
            return dict(mode=_lib.CONV_NONE)
        return dict(mode=self._conv_mode, kernel=self._kernel, kernel_1d=self._kernel_1d,
                    kernel_hi=self._kernel_hi, kernel_hi_1d=self._kernel_hi_1d)

    @property
    def pixel_area(self):
        return self._pixel_width**2

    def re_size_convolve(self, flux_array, unconvolved=False):
        """flux on ``coordinates_evaluate`` (last axis) -> convolved image on the pixel grid [..., ny, nx] * pixel
        area (numerics.py:157-177).  NumPy in -> NumPy out; torch CUDA in -> torch CUDA out."""
        want_torch = type(flux_array).__module__.startswith("torch")
        full = self._grid.flux_to_full_grid(flux_array)
        lead = tuple(full.shape[:-1])
        flat = full.reshape((-1, full.shape[-1]))
        key = bool(unconvolved)
        if key not in self._plans:
            self._plans[key] = _lib.Plan(0, grid=self.grid_dict(), psf=self.psf_dict(unconvolved),
                                         image_scale=self.pixel_area)
        out = self._plans[key].resize_convolve(flat, unconvolved=unconvolved)
        out = out.reshape(lead + (self._ny, self._nx))
        return out if want_torch else out.cpu().numpy()

    @property
    def point_source_supersampling_factor(self):
        return self._ps_supersampling

    @property
    def grid_supersampling_factor(self):
        return self._grid.supersampling_factor

    @property
    def coordinates_evaluate(self):
        return self._grid.coordinates_evaluate

    @staticmethod
    def _supersampling_cut_kernel(kernel_super, convolution_kernel_size, supersampling_factor):
        """numerics.py:197-214."""
        if convolution_kernel_size is not None:
            size = convolution_kernel_size * supersampling_factor
            if size % 2 == 0:
                size += 1
            return kernel_util.cut_psf(kernel_super, size, normalisation=False)
        return kernel_super

    @property
    def convolution_kernel(self):
        return self._kernel

    @property
    def grid_class(self):
        return self._grid


class NumericsSubFrame(object):
    """Finds the rectangular sub-frame that contains all flux_evaluate_indexes, runs the numerics only there and
    pads zeros around it (jaxtronomy/ImSim/Numerics/numerics_subframe.py:15-199).

    The reference cuts ``image[x_min:x_max+1, y_min:y_max+1]`` with x := array axis 0 and builds the sub-frame grid as
    ``PixelGrid(nx_sub = x extent, ny_sub = y extent, ra/dec at map_pix2coord(x_min, y_min))``; that is reproduced
    verbatim.  It only works when the sub-frame is square (otherwise ``_complete_frame`` cannot place the result),
    which is checked here."""

    def __init__(self, pixel_grid, psf, supersampling_factor=1, compute_mode="regular",
                 supersampling_convolution=False, supersampling_kernel_size=5, flux_evaluate_indexes=None,
                 supersampled_indexes=None, compute_indexes=None, point_source_supersampling_factor=1,
                 convolution_kernel_size=None, convolution_type="fft", truncation_conv=None):
        from jaxtronomy_b200.Data.pixel_grid import PixelGrid

        self._nx, self._ny = pixel_grid.num_pixel_axes
        self._init_sub_frame(flux_evaluate_indexes)
        if self._subframe_calc:
            nx_sub = self._x_max_sub - self._x_min_sub + 1
            ny_sub = self._y_max_sub - self._y_min_sub + 1
            if nx_sub != ny_sub:
                raise ValueError("the sub-frame enclosing flux_evaluate_indexes must be square (%d x %d): the "
                                 "reference cannot place a rectangular sub-frame back into the full frame"
                                 % (nx_sub, ny_sub))
            ra0, dec0 = pixel_grid.map_pix2coord(self._x_min_sub, self._y_min_sub)
            pixel_grid_sub = PixelGrid(nx=nx_sub, ny=ny_sub, transform_pix2angle=pixel_grid.transform_pix2angle,
                                       ra_at_xy_0=ra0, dec_at_xy_0=dec0)
        else:
            pixel_grid_sub = pixel_grid
        self._numerics_subframe = Numerics(
            pixel_grid=pixel_grid_sub, psf=psf, supersampling_factor=supersampling_factor, compute_mode=compute_mode,
            supersampling_convolution=supersampling_convolution, supersampling_kernel_size=supersampling_kernel_size,
            flux_evaluate_indexes=self._cut_frame(flux_evaluate_indexes),
            point_source_supersampling_factor=point_source_supersampling_factor,
            convolution_kernel_size=convolution_kernel_size, convolution_type=convolution_type,
            truncation_conv=truncation_conv)

    def _init_sub_frame(self, flux_evaluate_indexes):
        if flux_evaluate_indexes is None:
            self._subframe_calc = False
            self._x_min_sub, self._y_min_sub = 0, 0
            self._x_max_sub, self._y_max_sub = self._nx, self._ny
        else:
            self._subframe_calc = True
            idx = np.where(flux_evaluate_indexes)
            self._x_min_sub, self._x_max_sub = int(np.min(idx[0])), int(np.max(idx[0]))
            self._y_min_sub, self._y_max_sub = int(np.min(idx[1])), int(np.max(idx[1]))

    def _cut_frame(self, image):
        """Sub-frame of a full-frame 2d array (numerics_subframe.py:163-176); works on the last two axes."""
        if self._subframe_calc is True and image is not None:
            return image[..., self._x_min_sub:self._x_max_sub + 1, self._y_min_sub:self._y_max_sub + 1]
        return image

    def _complete_frame(self, image_sub_frame):
        """Zero-pad a sub-frame image (last two axes) back to the full frame (numerics_subframe.py:134-148)."""
        if not self._subframe_calc:
            return image_sub_frame
        lead = tuple(image_sub_frame.shape[:-2])
        if type(image_sub_frame).__module__.startswith("torch"):
            import torch

            image = torch.zeros(lead + (self._nx, self._ny), dtype=image_sub_frame.dtype,
                                device=image_sub_frame.device)
        else:
            image = np.zeros(lead + (self._nx, self._ny))
        image[..., self._x_min_sub:self._x_max_sub + 1, self._y_min_sub:self._y_max_sub + 1] = image_sub_frame
        return image

    @property
    def subframe_calc(self):
        return self._subframe_calc

    def cut_frame(self, image):
        return self._cut_frame(image)

    def complete_frame(self, image_sub_frame):
        return self._complete_frame(image_sub_frame)

    def grid_dict(self):
        return self._numerics_subframe.grid_dict()

    def psf_dict(self, unconvolved=False):
        return self._numerics_subframe.psf_dict(unconvolved)

    @property
    def pixel_area(self):
        return self._numerics_subframe.pixel_area

    def re_size_convolve(self, flux_array, unconvolved=False):
        return self._complete_frame(self._numerics_subframe.re_size_convolve(flux_array, unconvolved=unconvolved))

    @property
    def point_source_supersampling_factor(self):
        return self._numerics_subframe.point_source_supersampling_factor

    @property
    def grid_supersampling_factor(self):
        return self._numerics_subframe.grid_supersampling_factor

    @property
    def coordinates_evaluate(self):
        return self._numerics_subframe.coordinates_evaluate

    @property
    def convolution_kernel(self):
        return self._numerics_subframe.convolution_kernel

    @property
    def grid_class(self):
        return self._numerics_subframe.grid_class
