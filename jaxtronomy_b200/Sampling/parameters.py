"""Param: flat sampler vector <-> kwargs lists, after lenstronomy.Sampling.parameters.Param as used at
jaxtronomy/Sampling/likelihood.py:305.

Order of the flat vector: lens -> source light -> lens light, profile by profile, each profile in its
``param_names`` order, skipping fixed parameters.  With ``linear_solver=True`` every ``amp`` is fixed (to 1) and
solved linearly instead (LightParam / LinearBasis.add_fixed_linear, linear_basis.py:221-233).

Beyond the reference: ``column_source_fn()`` turns the same bookkeeping into the per-parameter column / constant
map of a kernel plan, so that a batch of sampler positions is consumed by the GPU without any per-particle Python.
Joint-parameter links, point sources and special parameters are outside the scope.
"""

import numpy as np

from jaxtronomy_b200 import _model

_DEFAULT_LOWER = {"theta_E": 0, "gamma": 1.5, "e1": -0.5, "e2": -0.5, "center_x": -100, "center_y": -100,
                  "gamma1": -0.5, "gamma2": -0.5, "ra_0": -100, "dec_0": -100, "s_scale": 0, "kappa": -10,
                  "Rs": 0, "alpha_Rs": 0, "r_trunc": 0, "m": 1, "a_m": 0, "phi_m": -np.pi, "r_E": 0,
                  "a1_a": -0.2, "delta_phi_m1": -np.pi, "a3_a": -0.2, "delta_phi_m3": -np.pi / 6, "a4_a": -0.2,
                  "delta_phi_m4": -np.pi / 8, "a": -1000, "s": 0, "sigma0": 0, "amp": 0, "R_sersic": 0,
                  "n_sersic": 0.5, "sigma": 0}
_DEFAULT_UPPER = {"theta_E": 100, "gamma": 2.5, "e1": 0.5, "e2": 0.5, "center_x": 100, "center_y": 100,
                  "gamma1": 0.5, "gamma2": 0.5, "ra_0": 100, "dec_0": 100, "s_scale": 100, "kappa": 10,
                  "Rs": 100, "alpha_Rs": 10, "r_trunc": 100, "m": 100, "a_m": 100, "phi_m": np.pi, "r_E": 100,
                  "a1_a": 0.2, "delta_phi_m1": np.pi, "a3_a": 0.2, "delta_phi_m3": np.pi / 6, "a4_a": 0.2,
                  "delta_phi_m4": np.pi / 8, "a": 1000, "s": 10000, "sigma0": 100, "amp": 100, "R_sersic": 100,
                  "n_sersic": 8, "sigma": 100}

_COMPONENTS = (("lens", "lens_model_list", "kwargs_lens", "lens"),
               ("source", "source_light_model_list", "kwargs_source", "source_light"),
               ("lens_light", "lens_light_model_list", "kwargs_lens_light", "lens_light"))


class Param(object):
    def __init__(self, kwargs_model, kwargs_fixed_lens=None, kwargs_fixed_source=None, kwargs_fixed_lens_light=None,
                 kwargs_lower_lens=None, kwargs_lower_source=None, kwargs_lower_lens_light=None,
                 kwargs_upper_lens=None, kwargs_upper_source=None, kwargs_upper_lens_light=None, linear_solver=True,
                 **unsupported):
        for k, v in unsupported.items():
            if v:
                raise ValueError("Param option %s is outside the scope of jaxtronomy_b200" % k)
        self._linear_solver = linear_solver
        self._names = {
            "lens": [_model.LENS_PARAM_NAMES[m] for m in kwargs_model.get("lens_model_list", [])],
            "source": [_model.LIGHT_PARAM_NAMES[m] for m in kwargs_model.get("source_light_model_list", [])],
            "lens_light": [_model.LIGHT_PARAM_NAMES[m] for m in kwargs_model.get("lens_light_model_list", [])],
        }

        def fill(lst, comp):
            n = len(self._names[comp])
            lst = [dict(d) for d in lst] if lst is not None else [{} for _ in range(n)]
            if len(lst) != n:
                raise ValueError("%s: %d dicts given for %d profiles" % (comp, len(lst), n))
            return lst

        self._fixed = {"lens": fill(kwargs_fixed_lens, "lens"), "source": fill(kwargs_fixed_source, "source"),
                       "lens_light": fill(kwargs_fixed_lens_light, "lens_light")}
        if linear_solver:
            for comp in ("source", "lens_light"):
                for d in self._fixed[comp]:
                    d.setdefault("amp", 1)
        self._lower = {"lens": fill(kwargs_lower_lens, "lens"), "source": fill(kwargs_lower_source, "source"),
                       "lens_light": fill(kwargs_lower_lens_light, "lens_light")}
        self._upper = {"lens": fill(kwargs_upper_lens, "lens"), "source": fill(kwargs_upper_source, "source"),
                       "lens_light": fill(kwargs_upper_lens_light, "lens_light")}
        # flat index of every sampled parameter
        self._index = {}
        self._flat_names = []
        for comp, _, _, label in _COMPONENTS:
            for i, names in enumerate(self._names[comp]):
                for n in names:
                    if n not in self._fixed[comp][i]:
                        self._index[(comp, i, n)] = len(self._flat_names)
                        self._flat_names.append("%s_%s%d" % (n, label, i))

    def num_param(self):
        return len(self._flat_names), list(self._flat_names)

    def param_limits(self):
        lower, upper = np.empty(len(self._flat_names)), np.empty(len(self._flat_names))
        for (comp, i, n), j in self._index.items():
            lower[j] = self._lower[comp][i].get(n, _DEFAULT_LOWER[n])
            upper[j] = self._upper[comp][i].get(n, _DEFAULT_UPPER[n])
        return lower, upper

    def args2kwargs(self, args):
        """flat vector (or [B, n] matrix: values become [B] arrays) -> kwargs lists."""
        args = np.asarray(args) if not type(args).__module__.startswith("torch") else args
        out = {}
        for comp, _, key, _ in _COMPONENTS:
            lst = []
            for i, names in enumerate(self._names[comp]):
                kw = {}
                for n in names:
                    if n in self._fixed[comp][i]:
                        kw[n] = self._fixed[comp][i][n]
                    else:
                        kw[n] = args[..., self._index[(comp, i, n)]]
                lst.append(kw)
            out[key] = lst
        out.update(kwargs_ps=[], kwargs_special={}, kwargs_extinction=[], kwargs_tracer_source=[])
        return out

    def kwargs2args(self, kwargs_lens=None, kwargs_source=None, kwargs_lens_light=None, **unused):
        given = {"lens": kwargs_lens or [], "source": kwargs_source or [], "lens_light": kwargs_lens_light or []}
        args = np.empty(len(self._flat_names))
        for (comp, i, n), j in self._index.items():
            args[j] = given[comp][i][n]
        return args

    def column_source_fn(self, index_lists=None):
        """-> source_fn(component, i_local, names) for LensModel.lens_descs / LightModel.light_descs.
        ``index_lists``: {"lens": [...], "source": [...], "lens_light": [...]} local -> global profile index of a
        band's sub-model (SingleBandMultiModel.select_kwargs); None = identity."""

        def fn(component, i_local, names):
            idx = (index_lists or {}).get(component)
            i = i_local if idx is None else idx[i_local]
            src = []
            for n in names:
                if n in self._fixed[component][i]:
                    src.append(("cst", float(self._fixed[component][i][n])))
                else:
                    src.append(("col", self._index[(component, i, n)]))
            return src

        return fn
