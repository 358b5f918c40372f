"""RegularGrid: host-side bookkeeping of the supersampled evaluation grid, jaxtronomy/ImSim/Numerics/grid.py:15-148.

Coordinates and index maps are computed once on the host exactly as the reference does (NumPy, same operation
order -> bit-identical) and uploaded with the plan; the kernels never regenerate them.
"""

import numpy as np

from jaxtronomy_b200.Data.pixel_grid import Coordinates1D
from jaxtronomy_b200.Util import util

__all__ = ["RegularGrid"]


class RegularGrid(Coordinates1D):
    """Manages a super-sampled grid on the partial image."""

    def __init__(self, nx, ny, transform_pix2angle, ra_at_xy_0, dec_at_xy_0, supersampling_factor=1,
                 flux_evaluate_indexes=None):
        super(RegularGrid, self).__init__(transform_pix2angle, ra_at_xy_0, dec_at_xy_0)
        self._supersampling_factor = supersampling_factor
        self._nx, self._ny = nx, ny
        self._x_grid, self._y_grid = self.coordinate_grid(nx, ny)
        if flux_evaluate_indexes is None:
            flux_evaluate_indexes = np.ones_like(self._x_grid)
        else:
            flux_evaluate_indexes = util.image2array(flux_evaluate_indexes)
        self._compute_indexes = self._subgrid_index(flux_evaluate_indexes, supersampling_factor, nx, ny).astype(bool)
        # full supersampled grid (what the kernels index); the evaluation set is its boolean compression
        self._ra_sub_full, self._dec_sub_full = util.make_subgrid(self._x_grid, self._y_grid, supersampling_factor,
                                                                  nx, ny)
        self._ra_subgrid = self._ra_sub_full[self._compute_indexes]
        self._dec_subgrid = self._dec_sub_full[self._compute_indexes]

    @property
    def coordinates_evaluate(self):
        """1d arrays of all coordinates being evaluated to perform the image computation."""
        return self._ra_subgrid, self._dec_subgrid

    @property
    def coordinates_full(self):
        """1d arrays of the full [ny*ss, nx*ss] supersampled grid (row-major)."""
        return self._ra_sub_full, self._dec_sub_full

    @property
    def grid_points_spacing(self):
        return self.pixel_width / self._supersampling_factor

    @property
    def num_grid_points_axes(self):
        return self._nx * self._supersampling_factor, self._ny * self._supersampling_factor

    @property
    def supersampling_factor(self):
        return self._supersampling_factor

    @staticmethod
    def _subgrid_index(idex_mask, subgrid_res, nx, ny):
        """1d mask at pixel resolution -> equivalent 1d mask at subgrid resolution (grid.py:119-131)."""
        # the reference calls lenstronomy's array2image here, which reshapes to (nx, ny) (jaxtronomy's own uses
        # (ny, nx)): each pixel's flag is repeated along x, then whole rows are repeated
        idex_sub = np.repeat(idex_mask, subgrid_res, axis=0)
        idex_sub = np.reshape(idex_sub, (int(nx), int(ny * subgrid_res)))
        idex_sub = np.repeat(idex_sub, subgrid_res, axis=0)
        return util.image2array(idex_sub)

    def flux_to_full_grid(self, flux_array):
        """Scatter a flux given on ``coordinates_evaluate`` into the full supersampled grid, zeros elsewhere
        (grid.py:133-148); works on the last axis of NumPy arrays and torch tensors."""
        if bool(np.all(self._compute_indexes)):
            return flux_array
        n_full = self._compute_indexes.size
        if type(flux_array).__module__.startswith("torch"):
            import torch

            out = torch.zeros(tuple(flux_array.shape[:-1]) + (n_full,), dtype=flux_array.dtype, device=flux_array.device)
            out[..., torch.as_tensor(self._compute_indexes, device=flux_array.device)] = flux_array
            return out
        flux_array = np.asarray(flux_array, dtype=float)
        out = np.zeros(flux_array.shape[:-1] + (n_full,))
        out[..., self._compute_indexes] = flux_array
        return out
