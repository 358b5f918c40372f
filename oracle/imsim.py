"""Oracle: supersampled grid, re-binning, PSF convolution, image assembly, WLS solve, chi^2.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Follows (all under /root/reference/jaxtronomy/)
  ImSim/Numerics/grid.py:22-64,101-148            RegularGrid
  ImSim/Numerics/convolution.py:73-96,153-166,177-273   Pixel/Subgrid/Gaussian convolution
  ImSim/Numerics/numerics.py:24-177               Numerics.__init__ / re_size_convolve
  Data/image_noise.py:16-137, Data/imaging_data.py:159-182   noise model, log-likelihood
  ImSim/image_model.py:27-150,162-221,223-393,431-492,542-595   ImageModel
  ImSim/image_linear_solve.py:97-181,228-339,361-455            ImageLinearFit
  ImSim/de_lens.py:10-36,39-49,77-105             WLS solve, marginalisation constant
  ImSim/Numerics/numerics_subframe.py:15-199      NumericsSubFrame (sub-frame cut / zero padding)
"""

import sys

import numpy as np
from scipy import signal

from . import lens_model as _lens
from . import light_profiles as _light
from .lenstronomy_port import PixelGrid, cut_psf, fwhm2sigma, make_grid, make_subgrid, split_kernel
from .util import array2image, image2array, nan_to_num, re_size

EPSILON = sys.float_info.epsilon


class RegularGrid:
    def __init__(self, nx, ny, transform_pix2angle, ra_at_xy_0, dec_at_xy_0, supersampling_factor=1,
                 flux_evaluate_indexes=None):
        self._ss = supersampling_factor
        self._nx, self._ny = nx, ny
        pg = PixelGrid(nx, ny, transform_pix2angle, ra_at_xy_0, dec_at_xy_0)
        self._x_grid, self._y_grid = pg.coordinate_grid_1d()
        if flux_evaluate_indexes is None:
            fei = np.ones_like(self._x_grid)
        else:
            fei = image2array(flux_evaluate_indexes)
        self._compute_indexes = self._subgrid_index(fei, self._ss, nx, ny).astype(bool)
        xs, ys = make_subgrid(self._x_grid, self._y_grid, self._ss, nx, ny)
        self._ra_subgrid = xs[self._compute_indexes]
        self._dec_subgrid = ys[self._compute_indexes]

    @staticmethod
    def _subgrid_index(idex_mask, subgrid_res, nx, ny):
        # grid.py:119-131 with lenstronomy's array2image, which reshapes to (nx, ny) -- unlike
        # jaxtronomy's own (ny, nx) -- so each pixel's flag is repeated along x, then whole rows are repeated
        idex_sub = np.repeat(idex_mask, subgrid_res, axis=0)
        idex_sub = np.reshape(idex_sub, (int(nx), int(ny * subgrid_res)))
        idex_sub = np.repeat(idex_sub, subgrid_res, axis=0)
        return image2array(idex_sub)

    @property
    def coordinates_evaluate(self):
        return self._ra_subgrid, self._dec_subgrid

    def _array2image(self, array):
        nx, ny = self._nx * self._ss, self._ny * self._ss
        grid1d = np.zeros(nx * ny)
        grid1d[self._compute_indexes] = array
        return array2image(grid1d, nx, ny)

    def flux_array2image_low_high(self, flux_array):
        image = self._array2image(flux_array)
        if self._ss > 1:
            return re_size(image, self._ss), image
        return image, None


class PixelKernelConvolution:
    def __init__(self, kernel, convolution_type="fft_static"):
        if convolution_type not in ("fft", "fft_static", "grid"):
            raise ValueError("convolution_type %s not supported!" % convolution_type)
        self._kernel = np.asarray(kernel, dtype=float)
        self.convolution_type = convolution_type

    def convolution2d(self, image):
        if self.convolution_type in ("fft", "fft_static"):
            return signal.fftconvolve(image, self._kernel, mode="same")
        return signal.convolve2d(image, self._kernel, mode="same")

    def re_size_convolve(self, image_low_res, image_high_res=None):
        return self.convolution2d(image_low_res)


class SubgridKernelConvolution:
    """convolution.py:99-166."""

    def __init__(self, kernel_supersampled, supersampling_factor, supersampling_kernel_size=None,
                 convolution_type="fft_static"):
        self._ss = supersampling_factor
        if supersampling_kernel_size is None:
            kernel_low_res, kernel_high_res = np.zeros((3, 3)), kernel_supersampled
            self._low_res_convolution = False
        else:
            kernel_low_res, kernel_high_res = split_kernel(kernel_supersampled, supersampling_kernel_size, self._ss)
            self._low_res_convolution = True
        self._low_res_conv = PixelKernelConvolution(kernel_low_res, convolution_type=convolution_type)
        self._high_res_conv = PixelKernelConvolution(kernel_high_res, convolution_type=convolution_type)

    def re_size_convolve(self, image_low_res, image_high_res):
        out = re_size(self._high_res_conv.convolution2d(image_high_res), self._ss)
        if self._low_res_convolution is True:
            out = out + self._low_res_conv.convolution2d(image_low_res)
        return out


class GaussianConvolution:
    def __init__(self, sigma, pixel_scale, supersampling_factor=1, supersampling_convolution=False, truncation=2):
        self._sigma_scaled = sigma / pixel_scale
        self._ss = supersampling_factor
        self._ss_conv = supersampling_convolution
        if supersampling_convolution is True:
            self._sigma_scaled *= supersampling_factor
        kernel_radius = max(round(self._sigma_scaled * truncation), 1)
        self._pad = int(kernel_radius)
        self.kernel = self.pixel_kernel(2 * kernel_radius + 1)
        self._conv = PixelKernelConvolution(self.kernel, "fft_static")

    def pixel_kernel(self, num_pix):
        x, y = make_grid(numPix=num_pix, deltapix=1)
        kernel = array2image(_light.gaussian(x, y, amp=1, sigma=self._sigma_scaled))
        return kernel / np.sum(kernel)

    def convolution2d(self, image):
        image = np.pad(image, pad_width=self._pad, mode="edge")
        out = self._conv.convolution2d(image)
        return out[self._pad:-self._pad, self._pad:-self._pad]

    def re_size_convolve(self, image_low_res, image_high_res):
        if self._ss_conv is True:
            return re_size(self.convolution2d(image_high_res), self._ss)
        return self.convolution2d(image_low_res)


class Numerics:
    def __init__(self, pixel_grid, psf, supersampling_factor=1, supersampling_convolution=False,
                 convolution_kernel_size=None, convolution_type="fft", truncation_conv=None,
                 flux_evaluate_indexes=None, supersampling_kernel_size=5, **unused):
        if not isinstance(supersampling_factor, int):
            raise TypeError("supersampling_factor needs to be an integer! Current type is %s" % type(supersampling_factor))
        if supersampling_factor == 1:
            supersampling_convolution = False
        self._psf_type = psf.psf_type
        self._pixel_width = pixel_grid.pixel_width
        nx, ny = pixel_grid.num_pixel_axes
        ra0, dec0 = pixel_grid.radec_at_xy_0
        self._grid = RegularGrid(nx, ny, pixel_grid.transform_pix2angle, ra0, dec0, supersampling_factor,
                                 flux_evaluate_indexes)
        if self._psf_type == "PIXEL":
            if supersampling_convolution is True:
                kernel_super = psf.kernel_point_source_supersampled(supersampling_factor)
                if convolution_kernel_size is not None:
                    size = convolution_kernel_size * supersampling_factor
                    if size % 2 == 0:
                        size += 1
                    kernel_super = cut_psf(kernel_super, size, normalisation=False)
                self._conv = SubgridKernelConvolution(kernel_super, supersampling_factor,
                                                      supersampling_kernel_size=supersampling_kernel_size,
                                                      convolution_type=convolution_type)
            else:
                kernel = psf.kernel_point_source
                if convolution_kernel_size is not None:
                    size = convolution_kernel_size
                    if size % 2 == 0:
                        size += 1
                    kernel = cut_psf(kernel, size, normalisation=False)
                self._conv = PixelKernelConvolution(kernel, convolution_type=convolution_type)
        elif self._psf_type == "GAUSSIAN":
            sigma = fwhm2sigma(psf.fwhm)
            if truncation_conv is None:
                truncation_conv = psf.truncation
            self._conv = GaussianConvolution(sigma, pixel_grid.pixel_width, supersampling_factor,
                                             supersampling_convolution, truncation=truncation_conv)
        elif self._psf_type == "NONE":
            self._conv = None
        else:
            raise ValueError("psf_type %s not valid! Chose either NONE, GAUSSIAN or PIXEL." % self._psf_type)

    @property
    def coordinates_evaluate(self):
        return self._grid.coordinates_evaluate

    def re_size_convolve(self, flux_array, unconvolved=False):
        low, high = self._grid.flux_array2image_low_high(flux_array)
        if unconvolved is True or self._psf_type == "NONE":
            conv = low
        else:
            conv = self._conv.re_size_convolve(low, high)
        return conv * self._pixel_width**2


class NumericsSubFrame:
    """numerics_subframe.py:15-199 (x := array axis 0, as in the reference)."""

    def __init__(self, pixel_grid, psf, flux_evaluate_indexes=None, **kwargs):
        self._nx, self._ny = pixel_grid.num_pixel_axes
        if flux_evaluate_indexes is None:
            self._subframe_calc = False
            pixel_grid_sub = pixel_grid
            cut = None
        else:
            self._subframe_calc = True
            idx = np.where(flux_evaluate_indexes)
            self._x_min, self._x_max = np.min(idx[0]), np.max(idx[0])
            self._y_min, self._y_max = np.min(idx[1]), np.max(idx[1])
            ra0, dec0 = pixel_grid.map_pix2coord(self._x_min, self._y_min)
            pixel_grid_sub = PixelGrid(self._x_max - self._x_min + 1, self._y_max - self._y_min + 1,
                                       pixel_grid.transform_pix2angle, ra0, dec0)
            cut = np.asarray(flux_evaluate_indexes)[self._x_min:self._x_max + 1, self._y_min:self._y_max + 1]
        self._numerics = Numerics(pixel_grid_sub, psf, flux_evaluate_indexes=cut, **kwargs)

    @property
    def coordinates_evaluate(self):
        return self._numerics.coordinates_evaluate

    def re_size_convolve(self, flux_array, unconvolved=False):
        sub = self._numerics.re_size_convolve(flux_array, unconvolved=unconvolved)
        if not self._subframe_calc:
            return sub
        image = np.zeros((self._nx, self._ny))
        image[self._x_min:self._x_max + 1, self._y_min:self._y_max + 1] = sub
        return image


def covariance_matrix(data, background_rms, exposure_map):
    """image_noise.py:113-137."""
    d_pos = np.where(data >= 0, data, 0)
    return d_pos / exposure_map + background_rms**2


class ImageData(PixelGrid):
    def __init__(self, image_data, exposure_time=None, background_rms=None, noise_map=None,
                 ra_at_xy_0=0, dec_at_xy_0=0, transform_pix2angle=None, ra_shift=0, dec_shift=0,
                 flux_scaling=1, **unused):
        nx, ny = np.shape(image_data)
        if transform_pix2angle is None:
            transform_pix2angle = np.array([[1, 0], [0, 1]])
        PixelGrid.__init__(self, nx, ny, transform_pix2angle, ra_at_xy_0 + ra_shift, dec_at_xy_0 + dec_shift)
        if exposure_time is None:
            if noise_map is None:
                raise ValueError("Exposure map has not been specified in Noise() class!")
        else:
            self.exposure_map = np.where(np.asarray(exposure_time) <= 10 ** (-10), 10 ** (-10), exposure_time)
        if background_rms is None:
            if noise_map is None:
                raise ValueError("rms background value as 'background_rms' not specified!")
            self.background_rms = np.median(noise_map)
        else:
            self.background_rms = background_rms
        self.data = np.array(image_data, dtype=float)
        self.flux_scaling = flux_scaling
        self._noise_map = None if noise_map is None else np.array(noise_map, dtype=float)
        if self._noise_map is not None:
            self.C_D = self._noise_map**2
        else:
            self.C_D = covariance_matrix(self.data, self.background_rms, self.exposure_map)

    def C_D_model(self, model):
        if self._noise_map is not None:
            return self._noise_map**2
        return covariance_matrix(model, self.background_rms, self.exposure_map)

    def log_likelihood(self, model, mask, additional_error_map=0):
        """imaging_data.py:159-182."""
        c_d = self.C_D_model(model)
        chi2 = (model - self.data) ** 2
        chi2 = chi2 / (c_d + np.abs(additional_error_map))
        chi2 = chi2 * mask
        return -np.sum(chi2) / 2


def get_param_WLS(A, C_D_inv, d, inv_bool=True):
    """de_lens.py:10-36 (A is [N_d, n_lin])."""
    A = np.asarray(A, dtype=float)
    M = A.T.dot(np.multiply(C_D_inv, A.T).T)
    stable = np.linalg.cond(M) < 5 / EPSILON
    R = A.T.dot(np.multiply(C_D_inv, d))
    if inv_bool:
        if stable:
            M_inv = nan_to_num(np.linalg.inv(M), posinf=0, neginf=0)
        else:
            M_inv = np.zeros_like(M)
        B = M_inv.dot(R)
    else:
        M_inv = None
        B = np.zeros(A.shape[1])
        if stable:
            b = np.linalg.solve(M, R).T
            if np.all(np.isfinite(b)):
                B = b
    return B, M_inv, A.dot(B)


def marginalisation_const(M_inv):
    sign, log_det = np.linalg.slogdet(M_inv)
    return -(10**15) if sign == 0 else sign * log_det / 2


class ImageModel:
    """ImageModel / ImageLinearFit restated for extended sources only (no point sources,
    image_model.py:73-76 default)."""

    def __init__(self, data_class, psf_class, lens_model_list=None, source_model_list=None,
                 lens_light_model_list=None, kwargs_numerics=None, likelihood_mask=None,
                 lens_profile_kwargs_list=None, point_source_type_list=None, psf_variance_map=None,
                 psf_error_map_bool_list=None, multi_plane=None):
        # multi_plane: None, or the host constants of MultiPlane (sorted_index, T_ij_list, T_z_list, reduced2physical,
        # T_ij_stop, T_z_source) -- LensModel(multi_plane=True) inside the image model, multi_plane.py:236-275
        self._multi_plane = multi_plane
        # point sources: UNLENSED / LENSED_POSITION without fixed magnification -- positions and amplitudes are the
        # parameters themselves (PointSource/Types/unlensed.py, lensed_position.py)
        self.point_source_type_list = list(point_source_type_list or [])
        # PSF error map (image_model.py:127-132,615-648): variance per kernel pixel for a unit point source
        self._psf_variance_map = None if psf_variance_map is None else np.asarray(psf_variance_map, dtype=float)
        self._psf_error_map_bool_list = psf_error_map_bool_list
        self._ps_supersampling = (kwargs_numerics or {}).get("point_source_supersampling_factor", 1)
        self.Data = data_class
        self.PSF = psf_class
        self._flux_scaling = getattr(data_class, "flux_scaling", 1)
        self.PSF.set_pixel_size(self.Data.pixel_width)
        kwargs_numerics = {k: v for k, v in (kwargs_numerics or {}).items() if k != "point_source_supersampling_factor"}
        self.ImageNumerics = NumericsSubFrame(pixel_grid=self.Data, psf=self.PSF, **kwargs_numerics)
        self.lens_model_list = list(lens_model_list or [])
        self.lens_profile_kwargs_list = lens_profile_kwargs_list
        self.source_model_list = list(source_model_list or [])
        self.lens_light_model_list = list(lens_light_model_list or [])
        if likelihood_mask is None:
            likelihood_mask = np.ones(data_class.num_pixel_axes)
        self.likelihood_mask = np.array(likelihood_mask, dtype=bool)
        self.num_data_evaluate = int(np.sum(self.likelihood_mask))
        self._mask1d = image2array(self.likelihood_mask)

    # --- image_model.py:223-393 ---
    def _ray_shoot(self, x, y, kwargs_lens):
        if self._multi_plane is not None:
            mp = self._multi_plane
            return _lens.ray_shooting_multi_plane(
                self.lens_model_list, x, y, kwargs_lens, [int(i) for i in mp["sorted_index"]], mp["T_ij_list"],
                mp["T_z_list"], mp["reduced2physical"], float(mp["T_ij_stop"]), float(mp["T_z_source"]),
                profile_kwargs_list=self.lens_profile_kwargs_list)
        return _lens.ray_shooting(self.lens_model_list, x, y, kwargs_lens or [],
                                  profile_kwargs_list=self.lens_profile_kwargs_list)

    def source_surface_brightness(self, kwargs_source, kwargs_lens=None, unconvolved=False, de_lensed=False, k=None):
        if len(self.source_model_list) == 0:
            return np.zeros(self.Data.num_pixel_axes)
        x, y = self.ImageNumerics.coordinates_evaluate
        if not de_lensed:
            x, y = self._ray_shoot(x, y, kwargs_lens)
        flux = _light.surface_brightness(self.source_model_list, x, y, kwargs_source, k=k)
        return self.ImageNumerics.re_size_convolve(flux * self._flux_scaling, unconvolved=unconvolved)

    def lens_surface_brightness(self, kwargs_lens_light, unconvolved=False, k=None):
        x, y = self.ImageNumerics.coordinates_evaluate
        flux = _light.surface_brightness(self.lens_light_model_list, x, y, kwargs_lens_light or [], k=k)
        return self.ImageNumerics.re_size_convolve(flux, unconvolved=unconvolved) * self._flux_scaling

    # --- point sources: image_model.py:395-429, PointSource/point_source.py:255-284,347-384 ---
    def _ps_render(self, ra, dec, amp, unconvolved=False):
        from oracle.point_source import point_source_rendering

        kernel = None if unconvolved else self.PSF.kernel_point_source_supersampled(self._ps_supersampling)
        return point_source_rendering(self.Data, kernel, self._ps_supersampling, ra, dec, amp, unconvolved=unconvolved)

    def error_map_psf(self, kwargs_ps):
        """ImageModel._error_map_psf, image_model.py:615-648."""
        from oracle.point_source import psf_variance_map

        error_map = np.zeros(self.Data.num_pixel_axes)
        if self._psf_variance_map is None or not self.point_source_type_list or not kwargs_ps:
            return error_map
        flags = self._psf_error_map_bool_list or [True] * len(self.point_source_type_list)
        for k, kw in enumerate(kwargs_ps):
            if flags[k] and np.size(kw["ra_image"]) > 0:
                error_map = error_map + psf_variance_map(self.Data, self.PSF.kernel_point_source,
                                                         self._psf_variance_map, kw["ra_image"], kw["dec_image"],
                                                         self.Data.data)
        return error_map

    def point_source(self, kwargs_ps, unconvolved=False, k=None):
        image = np.zeros(self.Data.num_pixel_axes)
        ra, dec, amp = [], [], []
        for i, kw in enumerate(kwargs_ps or []):
            if k is None or k == i:
                ra += list(np.atleast_1d(kw["ra_image"]))
                dec += list(np.atleast_1d(kw["dec_image"]))
                amp += list(np.atleast_1d(kw["point_amp"]))
        image = image + self._ps_render(np.array(ra), np.array(dec), np.array(amp), unconvolved=unconvolved)
        return image * self._flux_scaling

    def image(self, kwargs_lens=None, kwargs_source=None, kwargs_lens_light=None, unconvolved=False,
              source_add=True, lens_light_add=True, kwargs_ps=None, point_source_add=True):
        model = np.zeros(self.Data.num_pixel_axes)
        if source_add:
            model = model + self.source_surface_brightness(kwargs_source, kwargs_lens, unconvolved=unconvolved)
        if lens_light_add:
            model = model + self.lens_surface_brightness(kwargs_lens_light, unconvolved=unconvolved)
        if point_source_add and self.point_source_type_list:
            model = model + self.point_source(kwargs_ps, unconvolved=unconvolved)
        return model

    def likelihood_data_given_model(self, kwargs_lens=None, kwargs_source=None, kwargs_lens_light=None,
                                    kwargs_ps=None):
        """image_model.py:162-221 (no linear solve)."""
        im_sim = self.image(kwargs_lens, kwargs_source, kwargs_lens_light, kwargs_ps=kwargs_ps)
        model_error = self.error_map_psf(kwargs_ps)
        return self.Data.log_likelihood(im_sim, self.likelihood_mask, model_error)

    # --- masks ---
    def image2array_masked(self, image):
        return image2array(image)[self._mask1d]

    def array_masked2image(self, array):
        nx, ny = self.Data.num_pixel_axes
        grid1d = np.zeros(nx * ny)
        grid1d[self._mask1d] = array
        return array2image(grid1d, nx, ny)

    # --- image_linear_solve.py ---
    def linear_response_matrix(self, kwargs_lens, kwargs_source, kwargs_lens_light, unconvolved=False,
                               kwargs_ps=None):
        x, y = self.ImageNumerics.coordinates_evaluate
        rows = []
        if len(self.source_model_list) > 0:
            bx, by = self._ray_shoot(x, y, kwargs_lens)
            resp, _ = _light.functions_split(self.source_model_list, bx, by, kwargs_source)
            for image in resp:
                image = self.ImageNumerics.re_size_convolve(image, unconvolved=unconvolved)
                rows.append(nan_to_num(self.image2array_masked(image)))
        resp, _ = _light.functions_split(self.lens_light_model_list, x, y, kwargs_lens_light or [])
        for image in resp:
            image = self.ImageNumerics.re_size_convolve(image, unconvolved=unconvolved)
            rows.append(nan_to_num(self.image2array_masked(image)))
        for kw in (kwargs_ps or []) if self.point_source_type_list else []:  # one basis image per point-source image
            for ra, dec in zip(np.atleast_1d(kw["ra_image"]), np.atleast_1d(kw["dec_image"])):
                image = self._ps_render(np.array([ra]), np.array([dec]), np.array([1.0]))
                rows.append(nan_to_num(self.image2array_masked(image)))
        A = np.array(rows).reshape(len(rows), self.num_data_evaluate)
        return A * self._flux_scaling

    def image_linear_solve(self, kwargs_lens=None, kwargs_source=None, kwargs_lens_light=None, inv_bool=False,
                           kwargs_ps=None):
        A = self.linear_response_matrix(kwargs_lens, kwargs_source, kwargs_lens_light, kwargs_ps=kwargs_ps)
        model_error = self.error_map_psf(kwargs_ps)
        C_D_response = self.image2array_masked(self.Data.C_D + model_error)
        d = self.image2array_masked(self.Data.data)
        param, cov_param, wls_model = get_param_WLS(A.T, 1 / C_D_response, d, inv_bool=inv_bool)
        return self.array_masked2image(wls_model), model_error, cov_param, param

    def likelihood_data_given_model_linear(self, kwargs_lens=None, kwargs_source=None, kwargs_lens_light=None,
                                           source_marg=False, kwargs_ps=None):
        """ImageLinearFit.likelihood_data_given_model, image_linear_solve.py:228-339 -> (logL, param)."""
        model, model_error, cov, param = self.image_linear_solve(kwargs_lens, kwargs_source, kwargs_lens_light,
                                                                 inv_bool=source_marg, kwargs_ps=kwargs_ps)
        logL = self.Data.log_likelihood(model, self.likelihood_mask, model_error)
        if cov is not None and source_marg:
            logL = logL + marginalisation_const(cov)
        return logL, param
