"""Cosmological distances on the host (init-time constants of the multi-plane recursion).

The reference takes them from astropy through lenstronomy.Cosmo.background.Background
(jaxtronomy/LensModel/MultiPlane/multi_plane_base.py:50-52,100-115).  astropy is not a dependency here:
* if the user passes an astropy cosmology (anything with ``comoving_transverse_distance``), it is used as is;
* if the user passes an object with ``T_xy`` / ``d_xy`` methods, it is used as is;
* otherwise a flat LCDM (matter + Lambda, H0 = 67.66, Om0 = 0.30966 -- the Planck18 values without
  radiation/neutrinos) is integrated numerically.  Distances only enter as per-plane constants of the kernels.
"""

import numpy as np
from scipy.integrate import quad

_C_KM_S = 299792.458


class FlatLambdaCDM(object):
    def __init__(self, H0=67.66, Om0=0.30966):
        self.H0, self.Om0 = float(H0), float(Om0)

    def _chi(self, z):
        f = lambda zz: 1.0 / np.sqrt(self.Om0 * (1 + zz) ** 3 + 1 - self.Om0)  # noqa: E731
        return _C_KM_S / self.H0 * quad(f, 0, z, epsabs=1e-13, epsrel=1e-13)[0]

    def T_xy(self, z_observer, z_source):
        return float(self._chi(z_source) - self._chi(z_observer))

    def d_xy(self, z_observer, z_source):
        return self.T_xy(z_observer, z_source) / (1.0 + z_source)


class Background(object):
    def __init__(self, cosmo=None, **unused):
        if cosmo is None:
            cosmo = FlatLambdaCDM()
        self.cosmo = cosmo

    def T_xy(self, z_observer, z_source):
        """transverse comoving distance z_observer -> z_source [Mpc]"""
        c = self.cosmo
        if hasattr(c, "T_xy"):
            return float(c.T_xy(z_observer, z_source))
        if hasattr(c, "_comoving_transverse_distance_z1z2"):
            return float(c._comoving_transverse_distance_z1z2(z_observer, z_source).value)
        raise ValueError("cosmo must provide T_xy/d_xy or be an astropy cosmology")

    def d_xy(self, z_observer, z_source):
        """angular diameter distance z_observer -> z_source [Mpc]"""
        c = self.cosmo
        if hasattr(c, "d_xy"):
            return float(c.d_xy(z_observer, z_source))
        if hasattr(c, "angular_diameter_distance_z1z2"):
            return float(c.angular_diameter_distance_z1z2(z_observer, z_source).value)
        raise ValueError("cosmo must provide T_xy/d_xy or be an astropy cosmology")
