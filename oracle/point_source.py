"""ORACLE (test infrastructure): point-source rendering of the reference, restated in NumPy.

Follows jaxtronomy/ImSim/Numerics/point_source_rendering.py:33-77 (PointSourceRendering.point_source_rendering),
jaxtronomy/Util/image_util.py:9-40 (add_layer2image) and jaxtronomy/Util/util.py:86-99 (map_coord2pix).  The
sub-pixel shift is ``jax.scipy.ndimage.map_coordinates(kernel, [y_grid, x_grid], order=1)``: jax (third party, not
under /root/reference; jax/_src/scipy/ndimage.py, jax >= 0.4) interpolates linearly between floor and floor + 1 and
replaces every neighbour that falls outside the array by cval = 0, neighbour by neighbour (scipy's "grid-constant", NOT
scipy's "constant"); its published algorithm is restated in map_coordinates_linear below.
"""

import numpy as np


def map_coordinates_linear(inp, coordinates, cval=0.0):
    """jax.scipy.ndimage.map_coordinates(inp, coordinates, order=1, mode='constant', cval=cval) for a 2-d input."""
    inp = np.asarray(inp, dtype=float)
    nodes = []
    for c, size in zip(coordinates, inp.shape):
        c = np.asarray(c, dtype=float)
        lower = np.floor(c)
        upper_w = c - lower
        lower_w = 1 - upper_w
        idx = lower.astype(np.int64)
        nodes.append([(idx, lower_w, size), (idx + 1, upper_w, size)])
    out = None
    for (iy, wy, sy) in nodes[0]:
        for (ix, wx, sx) in nodes[1]:
            valid = (iy >= 0) & (iy < sy) & (ix >= 0) & (ix < sx)
            val = np.where(valid, inp[np.clip(iy, 0, sy - 1), np.clip(ix, 0, sx - 1)], cval)
            term = (wy * wx) * val
            out = term if out is None else out + term
    return out


def add_layer2image(grid2d, x_pos, y_pos, kernel):
    """image_util.py:9-40 (order = 1)."""
    k_rows, k_cols = kernel.shape
    n_row, n_col = grid2d.shape
    xrange = np.arange(n_col) + k_cols // 2 - x_pos
    yrange = np.arange(n_row) + k_rows // 2 - y_pos
    x_grid, y_grid = np.meshgrid(xrange, yrange)
    return map_coordinates_linear(kernel, [y_grid, x_grid]) + grid2d


def re_size(image, factor):
    if factor == 1:
        return image
    f = int(factor)
    nx, ny = image.shape
    return image.reshape(nx // f, f, ny // f, f).mean(3).mean(1)


def point_source_rendering(pixel_grid, kernel_supersampled, supersampling_factor, ra_pos, dec_pos, amp,
                           unconvolved=False):
    """point_source_rendering.py:33-77.  pixel_grid: oracle.lenstronomy_port.PixelGrid (needs _x_at_radec_0,
    _y_at_radec_0, _Ma2pix, num_pixel_axes)."""
    subgrid = int(supersampling_factor)
    nx, ny = pixel_grid.num_pixel_axes
    M = np.asarray(pixel_grid._Ma2pix)
    ra_pos, dec_pos, amp = np.atleast_1d(ra_pos), np.atleast_1d(dec_pos), np.atleast_1d(amp)
    x_pos = M[0, 0] * ra_pos + M[0, 1] * dec_pos + pixel_grid._x_at_radec_0
    y_pos = M[1, 0] * ra_pos + M[1, 1] * dec_pos + pixel_grid._y_at_radec_0
    x_sub = x_pos * subgrid + (subgrid - 1) / 2.0
    y_sub = y_pos * subgrid + (subgrid - 1) / 2.0
    if unconvolved:
        kernel = np.zeros((3, 3))
        kernel[1, 1] = 1
    else:
        kernel = np.asarray(kernel_supersampled, dtype=float)
    if len(x_pos) > len(amp):
        raise ValueError("there are %s images appearing but only %s amplitudes provided!" % (len(x_pos), len(amp)))
    subgrid2d = np.zeros((nx * subgrid, ny * subgrid))
    for i in range(len(x_pos)):
        subgrid2d = add_layer2image(subgrid2d, x_sub[i], y_sub[i], amp[i] * kernel)
    return re_size(subgrid2d, subgrid) * subgrid**2


def estimate_amp(data, x_pos, y_pos, psf_kernel):
    """jaxtronomy/Util/kernel_util.py:7-47."""
    numPix_x, numPix_y = data.shape
    x_int = int(np.round(x_pos - 0.49999))
    y_int = int(np.round(y_pos - 0.49999))
    ok = not (x_int <= 2 or x_int >= numPix_x - 2 or y_int <= 2 or y_int >= numPix_y - 2)
    if not ok:
        return 0.0
    mean_image = max(np.sum(data[y_int - 2:y_int + 3, x_int - 2:x_int + 3]), 0)
    num = len(psf_kernel)
    center = int((num - 0.5) / 2)
    mean_kernel = np.sum(psf_kernel[center - 2:center + 3, center - 2:center + 3])
    return mean_image / mean_kernel


def psf_variance_map(pixel_grid, psf_kernel, variance_kernel, ra_pos, dec_pos, data):
    """PointSourceRendering.psf_variance_map with fix_psf_variance_map=False
    (jaxtronomy/ImSim/Numerics/point_source_rendering.py:79-111)."""
    M = np.asarray(pixel_grid._Ma2pix)
    ra_pos, dec_pos = np.atleast_1d(ra_pos), np.atleast_1d(dec_pos)
    x_pos = M[0, 0] * ra_pos + M[0, 1] * dec_pos + pixel_grid._x_at_radec_0
    y_pos = M[1, 0] * ra_pos + M[1, 1] * dec_pos + pixel_grid._y_at_radec_0
    out = np.zeros_like(data, dtype=float)
    for i in range(len(x_pos)):
        amp = estimate_amp(data, x_pos[i], y_pos[i], psf_kernel)
        out = add_layer2image(out, x_pos[i], y_pos[i], variance_kernel * amp**2)
    return out
