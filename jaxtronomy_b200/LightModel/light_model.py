"""LightModel façade: jaxtronomy/LightModel/light_model.py, light_model_base.py:230-252, linear_basis.py:24-219.
Surface brightness is evaluated by lsim_surface_brightness (k_surface_brightness)."""

import numpy as np

from jaxtronomy_b200 import _arrays, _lib, _model
from jaxtronomy_b200.Util.util import convert_bool_list

__all__ = ["LightModel"]


class LightModel(object):
    """Class to handle extended surface brightness profiles (source and lens light)."""

    def __init__(self, light_model_list, deflection_scaling_list=None, source_redshift_list=None,
                 profile_kwargs_list=None):
        for name in light_model_list:
            if name not in _lib.LIGHT_TYPES:
                raise ValueError("No light model of type %s found! Supported models in jaxtronomy_b200 are %s"
                                 % (name, sorted(_lib.LIGHT_TYPES)))
        if source_redshift_list is not None and source_redshift_list != []:
            raise ValueError("multi plane sources not supported in jaxtronomy yet, source_redshift_list must be None")
        if profile_kwargs_list is not None and any(pk for pk in profile_kwargs_list):
            raise ValueError("light profile_kwargs (e.g. sersic_major_axis) are outside the scope of jaxtronomy_b200")
        self.profile_type_list = list(light_model_list)
        self.deflection_scaling_list = deflection_scaling_list
        self.redshift_list = source_redshift_list
        self.param_name_list = [_model.LIGHT_PARAM_NAMES[n] for n in self.profile_type_list]
        self._num_func = len(self.profile_type_list)
        self._plans = {}

    def _bool_list(self, k=None):
        return convert_bool_list(n=self._num_func, k=k)

    def check_parameters(self, kwargs_list):
        _model.gather_values(self.param_name_list, kwargs_list, "light")

    # -- descriptors (shared with ImageModel) -----------------------------------------------------------
    def light_descs(self, layout, component, lensed, slots, use_amp, source_fn=None):
        """slots: per profile, the output slot or -1.  ``source_fn``: see LensModel.lens_descs."""
        descs = []
        for i, name in enumerate(self.profile_type_list):
            if source_fn is not None:
                src = source_fn(component, i, self.param_name_list[i])
            else:
                src = layout.add_profile(component, i, self.param_name_list[i])
            descs.append(_model.light_desc(name, src, lensed=lensed, slot=slots[i], use_amp=use_amp))
        return descs

    def _plan(self, bool_list, split):
        key = (tuple(bool_list), split)
        if key not in self._plans:
            layout = _model.ColumnLayout()
            if split:
                slots, n = [], 0
                for b in bool_list:
                    slots.append(n if b else -1)
                    n += 1 if b else 0
                if n > _lib.MAX_SLOTS:
                    raise ValueError("at most %d linear components are supported" % _lib.MAX_SLOTS)
            else:
                slots, n = [0 if b else -1 for b in bool_list], 1
            descs = self.light_descs(layout, "light", lensed=False, slots=slots, use_amp=not split)
            self._plans[key] = (_lib.Plan(layout.n_cols, light_descs=descs, n_slots=max(n, 1)), n)
        return self._plans[key]

    def surface_brightness(self, x, y, kwargs_list, k=None):
        """Sum of the selected profiles at (x, y), light_model_base.py:230-252."""
        xf, yf, shape, want_torch = _arrays.flatten_coords(x, y)
        values = _model.gather_values(self.param_name_list, kwargs_list, "light")
        mat, B, batched = _model.pack_columns(values)
        plan, _ = self._plan(self._bool_list(k), split=False)
        out = plan.surface_brightness(mat, xf, yf)  # [B, 1, N]
        return _arrays.finish(out[:, 0, :], shape, batched, want_torch)

    def functions_split(self, x, y, kwargs_list, k=None):
        """amp = 1 response of every selected linear component -> (list of arrays, n), linear_basis.py:24-91."""
        xf, yf, shape, want_torch = _arrays.flatten_coords(x, y)
        values = _model.gather_values(self.param_name_list, kwargs_list, "light")
        mat, B, batched = _model.pack_columns(values)
        plan, n = self._plan(self._bool_list(k), split=True)
        if n == 0:
            return [], 0
        out = plan.surface_brightness(mat, xf, yf)  # [B, n, N]
        return [_arrays.finish(out[:, i, :], shape, batched, want_torch) for i in range(n)], n

    def num_param_linear(self, kwargs_list=None, list_return=False):
        n_list = [1] * self._num_func
        return n_list if list_return else int(np.sum(n_list))

    def update_linear(self, param, i, kwargs_list):
        """Write linear amplitudes param[i:] into kwargs_list (linear_basis.py:161-219)."""
        for k in range(self._num_func):
            kwargs_list[k]["amp"] = param[..., i] if hasattr(param, "ndim") and param.ndim > 1 else param[i]
            i += 1
        return kwargs_list, i

    def linear_param_from_kwargs(self, kwargs_list):
        return [kw["amp"] for kw in kwargs_list]
