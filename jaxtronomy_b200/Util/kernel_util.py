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
