"""PixelGrid / Coordinates (host): lenstronomy.Data.pixel_grid / coord_transforms as used by
jaxtronomy/Data/imaging_data.py:102-110 and jaxtronomy/ImSim/Numerics/grid.py:44-48."""

import numpy as np

from jaxtronomy_b200.Util import util


class Coordinates(object):
    def __init__(self, transform_pix2angle, ra_at_xy_0, dec_at_xy_0):
        self._Mpix2a = np.asarray(transform_pix2angle, dtype=float)
        self._Ma2pix = np.linalg.inv(self._Mpix2a)
        self._ra_at_xy_0 = ra_at_xy_0
        self._dec_at_xy_0 = dec_at_xy_0
        self._x_at_radec_0, self._y_at_radec_0 = util.map_coord2pix(-ra_at_xy_0, -dec_at_xy_0, 0, 0, self._Ma2pix)

    @property
    def transform_angle2pix(self):
        return self._Ma2pix

    @property
    def transform_pix2angle(self):
        return self._Mpix2a

    @property
    def xy_at_radec_0(self):
        return self._x_at_radec_0, self._y_at_radec_0

    @property
    def radec_at_xy_0(self):
        return self._ra_at_xy_0, self._dec_at_xy_0

    def map_coord2pix(self, ra, dec):
        return util.map_coord2pix(ra, dec, self._x_at_radec_0, self._y_at_radec_0, self._Ma2pix)

    def map_pix2coord(self, x, y):
        return util.map_coord2pix(x, y, self._ra_at_xy_0, self._dec_at_xy_0, self._Mpix2a)

    @property
    def pixel_area(self):
        return np.abs(np.linalg.det(self._Mpix2a))

    @property
    def pixel_width(self):
        return np.sqrt(self.pixel_area)

    def coordinate_grid(self, nx, ny):
        ra, dec = util.grid_from_coordinate_transform(nx, ny, self._Mpix2a, self._ra_at_xy_0, self._dec_at_xy_0)
        return util.array2image(ra, nx, ny), util.array2image(dec, nx, ny)


class Coordinates1D(Coordinates):
    def coordinate_grid(self, nx, ny):
        return util.grid_from_coordinate_transform(nx, ny, self._Mpix2a, self._ra_at_xy_0, self._dec_at_xy_0)


class PixelGrid(Coordinates):
    def __init__(self, nx, ny, transform_pix2angle, ra_at_xy_0, dec_at_xy_0, antenna_primary_beam=None):
        super(PixelGrid, self).__init__(transform_pix2angle, ra_at_xy_0, dec_at_xy_0)
        self._nx, self._ny = int(nx), int(ny)
        self._pb = antenna_primary_beam

    @property
    def num_pixel(self):
        return self._nx * self._ny

    @property
    def num_pixel_axes(self):
        return self._nx, self._ny

    @property
    def width(self):
        return self._nx * self.pixel_width, self._ny * self.pixel_width

    @property
    def center(self):
        x, y = self.pixel_coordinates
        return np.mean(x), np.mean(y)

    @property
    def pixel_coordinates(self):
        return self.coordinate_grid(self._nx, self._ny)

    @property
    def primary_beam(self):
        return self._pb
