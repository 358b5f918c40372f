"""lenstronomy.Util.simulation_util.data_configure_simple (used by the reference's tests and notebooks)."""

import numpy as np

from jaxtronomy_b200.Util import util


def data_configure_simple(numPix, deltaPix, exposure_time=None, background_rms=None, center_ra=0, center_dec=0,
                          inverse=False):
    _, _, ra0, dec0, _, _, Mpix2coord, _ = util.make_grid_with_coordtransform(
        numPix=numPix, deltapix=deltaPix, center_ra=center_ra, center_dec=center_dec, subgrid_res=1, inverse=inverse)
    return {
        "background_rms": background_rms,
        "exposure_time": exposure_time,
        "ra_at_xy_0": ra0,
        "dec_at_xy_0": dec0,
        "transform_pix2angle": Mpix2coord,
        "image_data": np.zeros((numPix, numPix)),
    }
