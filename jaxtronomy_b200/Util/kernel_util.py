"""PSF kernel helpers (host), after lenstronomy.Util.kernel_util as used at ImSim/Numerics/numerics.py:210."""

import numpy as np


def cut_edges(image, num_pix):
    nx, ny = image.shape
    if nx < num_pix or ny < num_pix:
        raise ValueError("image can not be resized, in routine cut_edges with image shape (%s %s) and desired new "
                         "shape (%s %s)" % (nx, ny, num_pix, num_pix))
    if (nx % 2) != (ny % 2) or (nx % 2) != (num_pix % 2):
        raise ValueError("image with odd and even axis / mismatching parity is not supported for cutting")
    x_min, y_min = (nx - num_pix) // 2, (ny - num_pix) // 2
    return np.array(image[x_min:nx - x_min, y_min:ny - y_min])


def cut_psf(psf_data, psf_size, normalisation=True):
    kernel = cut_edges(psf_data, psf_size)
    if normalisation:
        kernel = kernel / np.sum(kernel)
    return kernel


def kernel_gaussian_1d(num_pix, sigma_pix):
    """Normalised 1-d Gaussian on make_grid(num_pix, 1): outer(k, k) is GaussianConvolution.pixel_kernel
    (ImSim/Numerics/convolution.py:254-273) up to rounding."""
    a = np.arange(num_pix) - (num_pix - 1) / 2.0
    k = np.exp(-(a**2) / (2 * sigma_pix**2))
    return k / np.sum(k)


def averaging(grid, numGrid, numPix):
    """Re-size a [numGrid, numGrid] grid to [numPix, numPix] by block averaging (lenstronomy.Util.util.averaging)."""
    n = int(numGrid / numPix)
    return np.reshape(grid, (numPix, n, numPix, n)).mean(-1).mean(1)


def averaging_odd_kernel(kernel_high_res, degrading_factor):
    """Lower-resolution kernel of an odd high-resolution kernel and an odd degrading factor (zero-padded so that
    the high-resolution centre pixel sits in the centre of the centre low-resolution pixel)."""
    n_high = len(kernel_high_res)
    num_pix = int(round(n_high / degrading_factor + 0.5))
    if num_pix % 2 == 0:
        num_pix += 1
    n_pad = num_pix * degrading_factor
    padded = np.zeros((n_pad, n_pad))
    i0 = int((n_pad - n_high) / 2)
    padded[i0:i0 + n_high, i0:i0 + n_high] = kernel_high_res
    return averaging(padded, n_pad, num_pix)


def degrade_kernel(kernel_super, degrading_factor):
    """Supersampled kernel -> pixel-resolution kernel, flux conserving (lenstronomy.Util.kernel_util.degrade_kernel).
    Only odd degrading factors are supported here."""
    if degrading_factor == 1:
        return kernel_super
    if degrading_factor % 2 == 0:
        raise NotImplementedError("even supersampling factors of the PSF kernel are not supported by jaxtronomy_b200")
    return averaging_odd_kernel(kernel_super, degrading_factor) * degrading_factor**2


def split_kernel(kernel_super, supersampling_kernel_size, supersampling_factor, normalized=True):
    """(low-resolution kernel with a central hole, central high-resolution cut), flux conserving
    (lenstronomy.Util.kernel_util.split_kernel as used at jaxtronomy/ImSim/Numerics/convolution.py:123)."""
    kernel_super = np.array(kernel_super, dtype=float)
    n_super = len(kernel_super)
    n_sub = supersampling_kernel_size * supersampling_factor
    if n_sub % 2 == 0:
        n_sub += 1
    if n_sub > n_super:
        n_sub = n_super
    kernel_hole = kernel_super.copy()
    n_min = int((n_super - 1) / 2 - (n_sub - 1) / 2)
    n_max = int((n_super - 1) / 2 + (n_sub - 1) / 2 + 1)
    kernel_hole[n_min:n_max, n_min:n_max] = 0
    kernel_hole_resized = degrade_kernel(kernel_hole, degrading_factor=supersampling_factor)
    kernel_subgrid_cut = kernel_super[n_min:n_max, n_min:n_max].copy()
    if normalized:
        flux_subsampled = np.sum(kernel_subgrid_cut)
        flux_hole = np.sum(kernel_hole_resized)
        kernel_subgrid_cut = kernel_subgrid_cut * (1.0 - flux_hole) / flux_subsampled
    else:
        kernel_hole_resized = kernel_hole_resized / supersampling_factor**2
    return kernel_hole_resized, kernel_subgrid_cut
