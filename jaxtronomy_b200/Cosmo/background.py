"""Cosmological distances on the host (init-time constants of the multi-plane recursion).

The reference takes them from astropy through lenstronomy.Cosmo.background.Background
(jaxtronomy/LensModel/MultiPlane/multi_plane_base.py:50-52,100-115).  astropy is not a dependency here:
* if the user passes an astropy cosmology (anything with ``comoving_transverse_distance``), it is used as is;
* if the user passes an object with ``T_xy`` / ``d_xy`` methods, it is used as is;
* otherwise the reference's default (astropy's Planck18: flat LCDM with H0 = 67.66, Om0 = 0.30966, Tcmb0 = 2.7255 K,
  Neff = 3.046, one massive neutrino species of 0.06 eV) is integrated numerically: photons, massless and massive
  neutrinos enter E(z) the way astropy's FlatLambdaCDM does (Komatsu et al. 2011 fitting formula for the massive
  species).  astropy is not importable here, so this restatement is unpinned; it differs from a matter + Lambda
  model at the 1e-4 level in the multi-plane constants.  Distances only enter as per-plane constants of the kernels.
"""

import numpy as np
from scipy.integrate import quad

_C_KM_S = 299792.458


class FlatLambdaCDM(object):
    """Flat LCDM with radiation: E(z)^2 = Om0 (1+z)^3 + Ogamma0 (1 + nu(z)) (1+z)^4 + Ode0, Ode0 = 1 - Om0 - Ogamma0 (1 + nu(0)).

    Tcmb0 = 0 switches the radiation off (pure matter + Lambda)."""

    def __init__(self, H0=67.66, Om0=0.30966, Tcmb0=2.7255, Neff=3.046, m_nu=(0.0, 0.0, 0.06)):
        self.H0, self.Om0, self.Tcmb0, self.Neff = float(H0), float(Om0), float(Tcmb0), float(Neff)
        self.m_nu = tuple(float(m) for m in m_nu)
        # photon density today: rho_gamma = 4 sigma_SB T^4 / c^3; critical density 3 H0^2 / (8 pi G)  (SI)
        sigma_sb, c, G, mpc = 5.670374419e-8, 299792458.0, 6.67430e-11, 3.085677581491367e22
        h0_si = self.H0 * 1e3 / mpc
        rho_crit = 3.0 * h0_si**2 / (8.0 * np.pi * G)
        self.Ogamma0 = 4.0 * sigma_sb * self.Tcmb0**4 / c**3 / rho_crit if self.Tcmb0 > 0 else 0.0
        # neutrino temperature and mass-to-temperature ratios (astropy: Tnu0 = (4/11)^(1/3) Tcmb0, kB in eV/K)
        kb_ev = 8.617333262e-5
        self._nu_y = [m / (kb_ev * 0.7137658555036082 * self.Tcmb0) for m in self.m_nu if m > 0] if self.Tcmb0 > 0 else []
        self._n_massless = len(self.m_nu) - len(self._nu_y)
        self.Ode0 = 1.0 - self.Om0 - self.Ogamma0 * (1.0 + self._nu_rel(0.0))

    def _nu_rel(self, z):
        """Neutrino energy density relative to the photons' (astropy nu_relative_density, Komatsu et al. 2011 eq. 26)."""
        if self.Tcmb0 <= 0:
            return 0.0
        prefac, p, invp, k = 0.22710731766, 1.83, 0.54644808743, 0.3173
        rel = float(self._n_massless)
        for y in self._nu_y:
            rel += (1.0 + (k * y / (1.0 + z)) ** p) ** invp
        return prefac * (self.Neff / max(len(self.m_nu), 1)) * rel

    def _inv_E(self, z):
        zp = 1.0 + z
        return 1.0 / np.sqrt(self.Om0 * zp**3 + self.Ogamma0 * (1.0 + self._nu_rel(z)) * zp**4 + self.Ode0)

    def _chi(self, z):
        return _C_KM_S / self.H0 * quad(self._inv_E, 0, z, epsabs=1e-13, epsrel=1e-13)[0]

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
