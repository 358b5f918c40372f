"""Oracle: host-side lenstronomy helpers the reference imports on the hot path.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

lenstronomy (pinned ``>=1.13.6`` in /root/reference/setup.py:4, ``main`` in
requirements.txt:1) is NOT under /root/reference and not installable here, so
these are restated from lenstronomy's published behaviour (SURVEY.md Appendix
B) and anchored on the reference's call sites:
  make_grid / make_subgrid                     jaxtronomy/Util/util.py:12, ImSim/Numerics/grid.py:60
  Coordinates1D.coordinate_grid                ImSim/Numerics/grid.py:44-48
  PixelGrid                                    Data/imaging_data.py:102-110
  PSF                                          Util/class_creator.py:8, ImSim/Numerics/numerics.py:78-139
  kernel_util.cut_psf                          ImSim/Numerics/numerics.py:210
  fwhm2sigma                                   ImSim/Numerics/numerics.py:130
  simulation_util.data_configure_simple        test/test_ImSim/test_image_model.py setup
PARITY UNPINNED for this file (init-time bookkeeping only; no per-pixel arithmetic).
"""

import numpy as np

from .util import array2image, image2array


def fwhm2sigma(fwhm):
    return fwhm / (2 * np.sqrt(2 * np.log(2)))


def array2image_lenstronomy(array, nx=0, ny=0):
    """lenstronomy.Util.util.array2image reshapes to (nx, ny); jaxtronomy's own array2image (Util/util.py:18-37)
    reshapes to (ny, nx).  The reference mixes both (grid.py:128 uses this one)."""
    array = np.asarray(array)
    if nx == 0 or ny == 0:
        n = int(np.sqrt(len(array)))
        if n**2 != len(array):
            raise ValueError("length of input array given as %s is not square of integer number!" % len(array))
        nx, ny = n, n
    return array.reshape(int(nx), int(ny))


def make_grid(numPix, deltapix, subgrid_res=1, left_lower=False):
    if isinstance(numPix, (tuple, list, np.ndarray)):
        numPix = np.asarray(numPix, dtype=int)
    else:
        numPix = np.array([int(numPix), int(numPix)])
    numPix_eff = (numPix * subgrid_res).astype(int)
    deltapix_eff = deltapix / float(subgrid_res)
    x_grid = np.tile(np.arange(numPix_eff[0]), numPix_eff[1]) * deltapix_eff
    y_grid = np.repeat(np.arange(numPix_eff[1]), numPix_eff[0]) * deltapix_eff
    if left_lower:
        shift = -1.0 / 2 + 1.0 / (2 * subgrid_res) * np.array([1, 1])
    else:
        shift = deltapix_eff * (numPix_eff - 1) / 2
    return x_grid - shift[0], y_grid - shift[1]


def map_coord2pix(ra, dec, x_0, y_0, M):
    x, y = np.asarray(M).dot(np.array([ra, dec]))
    return x + x_0, y + y_0


def make_grid_with_coordtransform(numPix, deltapix, subgrid_res=1, center_ra=0, center_dec=0,
                                  left_lower=False, inverse=True):
    numPix_eff = numPix * subgrid_res
    deltapix_eff = deltapix / float(subgrid_res)
    a = np.arange(numPix_eff)
    matrix = np.dstack(np.meshgrid(a, a)).reshape(-1, 2)
    delta_x = -deltapix_eff if inverse else deltapix_eff
    if left_lower:
        ra_grid = matrix[:, 0] * delta_x
        dec_grid = matrix[:, 1] * deltapix_eff
    else:
        ra_grid = (matrix[:, 0] - (numPix_eff - 1) / 2.0) * delta_x
        dec_grid = (matrix[:, 1] - (numPix_eff - 1) / 2.0) * deltapix_eff
    shift = (subgrid_res - 1) / (2.0 * subgrid_res) * deltapix
    ra_grid = ra_grid - shift + center_ra
    dec_grid = dec_grid - shift + center_dec
    ra_at_xy_0, dec_at_xy_0 = ra_grid[0], dec_grid[0]
    Mpix2coord = np.array([[delta_x, 0], [0, deltapix_eff]])
    Mcoord2pix = np.linalg.inv(Mpix2coord)
    x0, y0 = map_coord2pix(-ra_at_xy_0, -dec_at_xy_0, 0, 0, Mcoord2pix)
    return ra_grid, dec_grid, ra_at_xy_0, dec_at_xy_0, x0, y0, Mpix2coord, Mcoord2pix


def data_configure_simple(numPix, deltaPix, exposure_time=None, background_rms=None,
                          center_ra=0, center_dec=0, inverse=False):
    (_, _, ra0, dec0, _, _, M, _) = make_grid_with_coordtransform(
        numPix=numPix, deltapix=deltaPix, center_ra=center_ra, center_dec=center_dec,
        subgrid_res=1, inverse=inverse)
    return {
        "background_rms": background_rms,
        "exposure_time": exposure_time,
        "ra_at_xy_0": ra0,
        "dec_at_xy_0": dec0,
        "transform_pix2angle": M,
        "image_data": np.zeros((numPix, numPix)),
    }


def grid_from_coordinate_transform(nx, ny, Mpix2coord, ra_at_xy_0, dec_at_xy_0):
    matrix = np.dstack(np.meshgrid(np.arange(nx), np.arange(ny))).reshape(-1, 2)
    xg, yg = matrix[:, 0], matrix[:, 1]
    ra = xg * Mpix2coord[0, 0] + yg * Mpix2coord[0, 1] + ra_at_xy_0
    dec = xg * Mpix2coord[1, 0] + yg * Mpix2coord[1, 1] + dec_at_xy_0
    return ra, dec


def make_subgrid(ra_coord, dec_coord, subgrid_res=2, nx=0, ny=0):
    ra_array = array2image(ra_coord, nx, ny)
    dec_array = array2image(dec_coord, nx, ny)
    ny_, nx_ = ra_array.shape
    d_ra_x = ra_array[0][1] - ra_array[0][0]
    d_ra_y = ra_array[1][0] - ra_array[0][0]
    d_dec_x = dec_array[0][1] - dec_array[0][0]
    d_dec_y = dec_array[1][0] - dec_array[0][0]
    ra_new = np.zeros((ny_ * subgrid_res, nx_ * subgrid_res))
    dec_new = np.zeros((ny_ * subgrid_res, nx_ * subgrid_res))
    s = subgrid_res
    for i in range(s):
        for j in range(s):
            ra_new[i::s, j::s] = (ra_array + d_ra_x * (-1 / 2.0 + 1 / (2.0 * s) + j / float(s))
                                  + d_ra_y * (-1 / 2.0 + 1 / (2.0 * s) + i / float(s)))
            dec_new[i::s, j::s] = (dec_array + d_dec_x * (-1 / 2.0 + 1 / (2.0 * s) + j / float(s))
                                   + d_dec_y * (-1 / 2.0 + 1 / (2.0 * s) + i / float(s)))
    return image2array(ra_new), image2array(dec_new)


class PixelGrid:
    def __init__(self, nx, ny, transform_pix2angle, ra_at_xy_0, dec_at_xy_0):
        self._nx, self._ny = int(nx), int(ny)
        self._Mpix2a = np.asarray(transform_pix2angle, dtype=float)
        self._Ma2pix = np.linalg.inv(self._Mpix2a)
        self._ra_at_xy_0, self._dec_at_xy_0 = ra_at_xy_0, dec_at_xy_0
        self._x_at_radec_0, self._y_at_radec_0 = map_coord2pix(-ra_at_xy_0, -dec_at_xy_0, 0, 0, self._Ma2pix)

    @property
    def num_pixel_axes(self):
        return self._nx, self._ny

    @property
    def transform_pix2angle(self):
        return self._Mpix2a

    @property
    def radec_at_xy_0(self):
        return self._ra_at_xy_0, self._dec_at_xy_0

    @property
    def pixel_width(self):
        return np.sqrt(np.abs(np.linalg.det(self._Mpix2a)))

    def coordinate_grid_1d(self):
        return grid_from_coordinate_transform(self._nx, self._ny, self._Mpix2a, self._ra_at_xy_0, self._dec_at_xy_0)

    def map_pix2coord(self, x, y):
        return map_coord2pix(x, y, self._ra_at_xy_0, self._dec_at_xy_0, self._Mpix2a)


def cut_edges(image, num_pix):
    nx, ny = image.shape
    if nx < num_pix or ny < num_pix:
        raise ValueError("image can not be resized, in routine cut_edges with image shape (%s %s) and desired new shape (%s %s)" % (nx, ny, num_pix, num_pix))
    if (nx % 2) != (ny % 2) or (nx % 2) != (num_pix % 2):
        raise ValueError("image with odd and even axis / mismatching parity is not supported")
    x_min = int((nx - num_pix) / 2)
    y_min = int((ny - num_pix) / 2)
    return np.array(image[x_min:nx - x_min, y_min:ny - y_min])


def cut_psf(psf_data, psf_size, normalisation=True):
    kernel = cut_edges(psf_data, psf_size)
    if normalisation:
        kernel = kernel / np.sum(kernel)
    return kernel


def averaging(grid, numGrid, numPix):
    n = int(numGrid / numPix)
    return np.reshape(grid, (numPix, n, numPix, n)).mean(-1).mean(1)


def degrade_kernel(kernel_super, degrading_factor):
    """lenstronomy.Util.kernel_util.degrade_kernel, odd factors (averaging_odd_kernel)."""
    if degrading_factor == 1:
        return kernel_super
    assert degrading_factor % 2 == 1
    n_high = len(kernel_super)
    num_pix = int(round(n_high / degrading_factor + 0.5))
    if num_pix % 2 == 0:
        num_pix += 1
    n_pad = num_pix * degrading_factor
    padded = np.zeros((n_pad, n_pad))
    i0 = int((n_pad - n_high) / 2)
    padded[i0:i0 + n_high, i0:i0 + n_high] = kernel_super
    return averaging(padded, n_pad, num_pix) * degrading_factor**2


def split_kernel(kernel_super, supersampling_kernel_size, supersampling_factor, normalized=True):
    """lenstronomy.Util.kernel_util.split_kernel (convolution.py:123)."""
    kernel_super = np.array(kernel_super, dtype=float)
    n_super = len(kernel_super)
    n_sub = supersampling_kernel_size * supersampling_factor
    if n_sub % 2 == 0:
        n_sub += 1
    if n_sub > n_super:
        n_sub = n_super
    hole = kernel_super.copy()
    n_min = int((n_super - 1) / 2 - (n_sub - 1) / 2)
    n_max = int((n_super - 1) / 2 + (n_sub - 1) / 2 + 1)
    hole[n_min:n_max, n_min:n_max] = 0
    hole_resized = degrade_kernel(hole, supersampling_factor)
    cut = kernel_super[n_min:n_max, n_min:n_max].copy()
    if normalized:
        cut = cut * (1.0 - np.sum(hole_resized)) / np.sum(cut)
    else:
        hole_resized = hole_resized / supersampling_factor**2
    return hole_resized, cut


class PSF:
    def __init__(self, psf_type="NONE", fwhm=None, truncation=5, pixel_size=None,
                 kernel_point_source=None, kernel_point_source_normalisation=True,
                 point_source_supersampling_factor=1):
        self.psf_type = psf_type
        self._pixel_size = pixel_size
        if psf_type == "GAUSSIAN":
            if fwhm is None:
                raise ValueError("fwhm must be set for GAUSSIAN psf type!")
            self.fwhm = fwhm
            self.truncation = truncation
        elif psf_type == "PIXEL":
            if kernel_point_source is None:
                raise ValueError("kernel_point_source needs to be specified for PIXEL PSF type!")
            k = np.array(kernel_point_source, dtype=float)
            if len(k) % 2 == 0:
                raise ValueError("kernel needs to have odd axis number, not ", np.shape(k))
            self.point_source_supersampling_factor = point_source_supersampling_factor
            if point_source_supersampling_factor > 1:
                self._kernel_super = k / np.sum(k) if kernel_point_source_normalisation else k
                k = degrade_kernel(k, point_source_supersampling_factor)
            if kernel_point_source_normalisation:
                k = k / np.sum(k)
            self.kernel_point_source = k
        elif psf_type == "NONE":
            k = np.zeros((3, 3))
            k[1, 1] = 1
            self.kernel_point_source = k
        else:
            raise ValueError("psf_type %s not supported!" % psf_type)

    def kernel_point_source_supersampled(self, supersampling_factor, updata_cache=True):
        if supersampling_factor == 1:
            return self.kernel_point_source
        if getattr(self, "_kernel_super", None) is not None and \
                self.point_source_supersampling_factor == supersampling_factor:
            return self._kernel_super
        raise NotImplementedError("oracle: kernel_util.subgrid_kernel interpolation is not restated")

    def set_pixel_size(self, deltaPix):
        self._pixel_size = deltaPix
