"""Host-side array / grid helpers (NumPy; init-time bookkeeping only).

Mirrors jaxtronomy/Util/util.py:18-139 and the lenstronomy.Util.util functions it re-exports
(make_grid, make_subgrid; jaxtronomy/Util/util.py:12).
"""

import numpy as np


def array2image(array, nx=0, ny=0):
    """1-d array -> [ny, nx] image (x fastest), jaxtronomy/Util/util.py:18-37."""
    array = np.asarray(array)
    if nx == 0 or ny == 0:
        n = int(np.sqrt(array.size))
        if n**2 != array.size:
            raise ValueError("length of input array given as %s is not square of integer number!" % array.size)
        nx, ny = n, n
    return np.reshape(array, (ny, nx))


def image2array(image):
    """jaxtronomy/Util/util.py:72-82."""
    image = np.asarray(image)
    return np.reshape(image, image.size)


def fwhm2sigma(fwhm):
    return fwhm / (2 * np.sqrt(2 * np.log(2)))


def sigma2fwhm(sigma):
    return sigma * (2 * np.sqrt(2 * np.log(2)))


def map_coord2pix(ra, dec, x_0, y_0, M):
    x, y = np.asarray(M).dot(np.array([ra, dec]))
    return x + x_0, y + y_0


def make_grid(numPix, deltapix, subgrid_res=1, left_lower=False):
    """Centred (or lower-left anchored) square/rectangular coordinate grid, flattened x-fastest."""
    if isinstance(numPix, (tuple, list, np.ndarray)):
        if len(numPix) != 2:
            raise ValueError("numPix must be an int or a pair of ints")
        n = np.asarray(numPix, dtype=int)
    else:
        n = np.array([int(numPix), int(numPix)])
    n_eff = n * int(subgrid_res)
    d_eff = deltapix / float(subgrid_res)
    x = np.tile(np.arange(n_eff[0]), n_eff[1]) * d_eff
    y = np.repeat(np.arange(n_eff[1]), n_eff[0]) * d_eff
    if left_lower:
        shift = -1.0 / 2 + 1.0 / (2 * subgrid_res) * np.array([1, 1])
    else:
        shift = d_eff * (n_eff - 1) / 2
    return x - shift[0], y - shift[1]


def make_grid_with_coordtransform(numPix, deltapix, subgrid_res=1, center_ra=0, center_dec=0, left_lower=False,
                                  inverse=True):
    n_eff = numPix * subgrid_res
    d_eff = deltapix / float(subgrid_res)
    idx = np.arange(n_eff)
    ix, iy = np.meshgrid(idx, idx)
    ix, iy = ix.ravel(), iy.ravel()
    dx = -d_eff if inverse else d_eff
    if left_lower:
        ra, dec = ix * dx, iy * d_eff
    else:
        ra = (ix - (n_eff - 1) / 2.0) * dx
        dec = (iy - (n_eff - 1) / 2.0) * d_eff
    shift = (subgrid_res - 1) / (2.0 * subgrid_res) * deltapix
    ra = ra - shift + center_ra
    dec = dec - shift + center_dec
    Mpix2coord = np.array([[dx, 0], [0, d_eff]])
    Mcoord2pix = np.linalg.inv(Mpix2coord)
    x0, y0 = map_coord2pix(-ra[0], -dec[0], 0, 0, Mcoord2pix)
    return ra, dec, ra[0], dec[0], x0, y0, Mpix2coord, Mcoord2pix


def grid_from_coordinate_transform(nx, ny, Mpix2coord, ra_at_xy_0, dec_at_xy_0):
    ix, iy = np.meshgrid(np.arange(nx), np.arange(ny))
    ix, iy = ix.ravel(), iy.ravel()
    ra = ix * Mpix2coord[0, 0] + iy * Mpix2coord[0, 1] + ra_at_xy_0
    dec = ix * Mpix2coord[1, 0] + iy * Mpix2coord[1, 1] + dec_at_xy_0
    return ra, dec


def make_subgrid(ra_coord, dec_coord, subgrid_res=2, nx=0, ny=0):
    """Sub-pixel coordinates of a regular grid; sub-pixel (i, j) of a pixel sits at the pixel centre plus
    d_x (-1/2 + 1/(2 s) + j/s) + d_y (-1/2 + 1/(2 s) + i/s)."""
    ra = array2image(ra_coord, nx, ny)
    dec = array2image(dec_coord, nx, ny)
    s = int(subgrid_res)
    d_ra_x, d_ra_y = ra[0][1] - ra[0][0], ra[1][0] - ra[0][0]
    d_dec_x, d_dec_y = dec[0][1] - dec[0][0], dec[1][0] - dec[0][0]
    ra_new = np.zeros((ra.shape[0] * s, ra.shape[1] * s))
    dec_new = np.zeros_like(ra_new)
    for i in range(s):
        for j in range(s):
            oj = -1 / 2.0 + 1 / (2.0 * s) + j / float(s)
            oi = -1 / 2.0 + 1 / (2.0 * s) + i / float(s)
            ra_new[i::s, j::s] = ra + d_ra_x * oj + d_ra_y * oi
            dec_new[i::s, j::s] = dec + d_dec_x * oj + d_dec_y * oi
    return image2array(ra_new), image2array(dec_new)


def convert_bool_list(n, k=None):
    """lenstronomy.Util.util.convert_bool_list (profile_list_base.py:244, light_model_base.py:428)."""
    if k is None:
        return [True] * n
    if isinstance(k, (int, np.integer)) and not isinstance(k, (bool, np.bool_)):
        out = [False] * n
        out[k] = True
        return out
    k = list(k)
    if len(k) == 0:
        return [False] * n
    if isinstance(k[0], (bool, np.bool_)):
        if len(k) != n:
            raise ValueError("length of selected lens models in format of boolean list is %s and does not match the "
                             "models of this class instance %s." % (len(k), n))
        return [bool(v) for v in k]
    out = [False] * n
    for i in k:
        if i >= n:
            raise ValueError("k as set by %s is not convertable in a bool string!" % k)
        out[i] = True
    return out
