"""LensModelBulk façade (jaxtronomy/LensModel/lens_model_bulk.py:10-128, single_plane_bulk.py:47-131,
MultiPlane/multi_plane_bulk.py): many deflectors (substructure), struct-of-arrays kwargs.

The reference materialises a [N_deflectors, N_rays] array per component and sums it (lax.map + switch);
here k_ray_shoot streams the deflector list through shared memory in chunks and accumulates in registers.
Parameters are wired by NAME (the reference's positional call mis-wires EPL_MULTIPOLE_M1M3M4).
"""

import numpy as np

from jaxtronomy_b200 import _arrays, _lib, _model
from jaxtronomy_b200.Cosmo.background import Background

__all__ = ["LensModelBulk"]


class LensModelBulk(object):
    def __init__(self, unique_lens_model_list, multi_plane, profile_kwargs_list=None, cosmo=None,
                 cosmology_model="FlatLambdaCDM"):
        for name in unique_lens_model_list:
            if name not in _lib.LENS_TYPES:
                raise ValueError("%s is not a valid lens model. Supported are: %s." % (name, sorted(_lib.LENS_TYPES)))
        self.unique_lens_model_list = list(unique_lens_model_list)
        if profile_kwargs_list is None:
            profile_kwargs_list = [{} for _ in range(len(unique_lens_model_list))]
        self.profile_kwargs_list = list(profile_kwargs_list) + [{}]
        self._names = self.unique_lens_model_list + ["NULL"]
        self.param_name_list = [_model.LENS_PARAM_NAMES[n] for n in self._names]
        self.multi_plane = bool(multi_plane)
        self.type = "MultiPlaneBulk" if self.multi_plane else "SinglePlaneBulk"
        self._bkg = Background(cosmo) if self.multi_plane else None
        self._plan_cache = {}

    def prepare_ray_shooting_kwargs(self, lens_model_list, kwargs_lens, z_source=None, lens_redshift_list=None,
                                    num_deflectors=None):
        """lenstronomy-convention (lens_model_list, kwargs_lens) -> struct of arrays, padded with NULL profiles to
        num_deflectors (single_plane_bulk.py:47-106, multi_plane_bulk.py:54-140)."""
        if len(lens_model_list) != len(kwargs_lens):
            raise ValueError("length of lens model list %d and length of kwargs_lens %d do not match"
                             % (len(lens_model_list), len(kwargs_lens)))
        if num_deflectors is None:
            num_deflectors = len(lens_model_list)
        elif num_deflectors < len(lens_model_list):
            raise ValueError("The provided num_deflectors %d is smaller than the length of lens_model_list %d"
                             % (num_deflectors, len(lens_model_list)))
        out = {}
        if self.multi_plane:
            if lens_redshift_list is None:
                raise ValueError("In multi-plane lensing, you need to specify the redshifts of the lensing planes.")
            if z_source is None:
                raise ValueError("z_source needs to be set for multi-plane lens modelling.")
            if len(lens_redshift_list) != len(lens_model_list):
                raise ValueError("length of lens model list and length of redshift list should match")
            order = np.argsort(np.asarray(lens_redshift_list), kind="stable")
            lens_model_list = [lens_model_list[i] for i in order]
            kwargs_lens = [kwargs_lens[i] for i in order]
            z_sorted = [lens_redshift_list[i] for i in order]
            T_ij, T_z, fac = [], [], []
            z_before, Tz = 0, 0
            for z in z_sorted:
                if z_before == z:
                    dT = 0
                else:
                    Tz = self._bkg.T_xy(0, z)
                    dT = self._bkg.T_xy(z_before, z)
                T_ij.append(dT)
                T_z.append(Tz)
                fac.append(self._bkg.d_xy(0, z_source) / self._bkg.d_xy(z, z_source))
                z_before = z
            out.update(T_ij_list=np.array(T_ij, dtype=float), T_z_list=np.array(T_z, dtype=float),
                       reduced2physical_factor=np.array(fac, dtype=float),
                       T_ij_end=self._bkg.T_xy(z_before, z_source), T_z_source=self._bkg.T_xy(0, z_source))
        num_filler = num_deflectors - len(lens_model_list)
        index_list = [self.unique_lens_model_list.index(m) for m in lens_model_list]
        index_list += [len(self.unique_lens_model_list)] * num_filler
        all_kwargs = {}
        for kwarg in sorted(set(p for names in self.param_name_list for p in names)):
            vals = [kw.get(kwarg, 0) for kw in kwargs_lens] + [0] * num_filler
            all_kwargs[kwarg] = np.array(vals, dtype=float)
        out.update(all_kwargs=all_kwargs, index_list=np.array(index_list))
        return out

    def _plan(self, rs):
        """Deflector parameters are per-plan constants here (B = 1 ray-shooting of a fixed realisation)."""
        index_list = np.asarray(rs["index_list"]).astype(int)
        all_kwargs = rs["all_kwargs"]
        descs = []
        for d, idx in enumerate(index_list):
            name = self._names[idx]
            if name == "NULL":
                continue
            src = [("cst", float(np.asarray(all_kwargs[p])[d])) for p in _model.LENS_PARAM_NAMES[name]]
            extra = {}
            if self.multi_plane:
                extra = dict(delta_T=float(rs["T_ij_list"][d]), T_z=float(rs["T_z_list"][d]),
                             red2phys=float(rs["reduced2physical_factor"][d]))
            descs.append(_model.lens_desc(name, src, self.profile_kwargs_list[idx], **extra))
        kw = dict(multi_plane=False)
        if self.multi_plane:
            kw = dict(multi_plane=True, T_ij_end=float(rs["T_ij_end"]), T_z_source=float(rs["T_z_source"]))
        return _lib.Plan(0, lens_descs=descs, **kw)

    def ray_shooting(self, x, y, ray_shooting_kwargs):
        """Maps image to source position through all deflectors (lens_model_bulk.py:113-128)."""
        key = id(ray_shooting_kwargs)
        if self._plan_cache.get("key") != key:
            self._plan_cache = {"key": key, "plan": self._plan(ray_shooting_kwargs), "ref": ray_shooting_kwargs}
        plan = self._plan_cache["plan"]
        xf, yf, shape, want_torch = _arrays.flatten_coords(x, y)
        ox, oy = plan.ray_shoot(np.zeros((1, 0)), xf, yf, what=0)
        return _arrays.finish(ox, shape, False, want_torch), _arrays.finish(oy, shape, False, want_torch)
