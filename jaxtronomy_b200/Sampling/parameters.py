"""Param: flat sampler vector <-> kwargs lists, after lenstronomy.Sampling.parameters.Param as used at
jaxtronomy/Sampling/likelihood.py:305.

Order of the flat vector: lens -> source light -> lens light, profile by profile, each profile in its
``param_names`` order, skipping fixed parameters.  With ``linear_solver=True`` every ``amp`` is fixed (to 1) and
solved linearly instead (LightParam / LinearBasis.add_fixed_linear, linear_basis.py:221-233).

Beyond the reference: ``column_source_fn()`` turns the same bookkeeping into the per-parameter column / constant
map of a kernel plan, so that a batch of sampler positions is consumed by the GPU without any per-particle Python.
Point sources (UNLENSED / LENSED_POSITION: ra_image, dec_image, point_amp per image, after the light components) are
sampled too; joint-parameter links and special parameters are outside the scope.
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
                 kwargs_fixed_ps=None, kwargs_lower_ps=None, kwargs_upper_ps=None, num_point_source_list=None,
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
        # point sources after the light components (lenstronomy Param order); per model ra_image[0..n),
        # dec_image[0..n), point_amp[0..n) (PointSourceParam.set_params).  The image count per model is static.
        self._ps_models = list(kwargs_model.get("point_source_model_list", []))
        for m in self._ps_models:
            if m not in ("UNLENSED", "LENSED_POSITION"):
                raise ValueError("point-source model %s is outside the scope of jaxtronomy_b200" % m)
        if self._ps_models and (num_point_source_list is None or len(num_point_source_list) != len(self._ps_models)):
            raise ValueError("num_point_source_list (images per point-source model) is required")
        self._ps_num = [int(n) for n in (num_point_source_list or [])]
        fill_ps = lambda lst: [dict(d) for d in lst] if lst is not None else [{} for _ in self._ps_models]  # noqa: E731
        self._ps_fixed, self._ps_lower, self._ps_upper = fill_ps(kwargs_fixed_ps), fill_ps(kwargs_lower_ps), \
            fill_ps(kwargs_upper_ps)
        if linear_solver:
            for d, n in zip(self._ps_fixed, self._ps_num):
                d.setdefault("point_amp", np.ones(n))
        for k, n in enumerate(self._ps_num):
            for name in ("ra_image", "dec_image", "point_amp"):
                if name not in self._ps_fixed[k]:
                    for j in range(n):
                        self._index[("point_source", k, name, j)] = len(self._flat_names)
                        self._flat_names.append("%s_%d_point_source%d" % (name, j, k))

    def num_param(self):
        return len(self._flat_names), list(self._flat_names)

    def param_limits(self):
        lower, upper = np.empty(len(self._flat_names)), np.empty(len(self._flat_names))
        ps_lo, ps_hi = {"ra_image": -100, "dec_image": -100, "point_amp": 0}, \
            {"ra_image": 100, "dec_image": 100, "point_amp": 100}
        for key, j in self._index.items():
            if key[0] == "point_source":
                _, k, n, e = key
                lo, hi = self._ps_lower[k].get(n, ps_lo[n]), self._ps_upper[k].get(n, ps_hi[n])
                lower[j], upper[j] = np.broadcast_to(lo, (self._ps_num[k],))[e], np.broadcast_to(hi, (self._ps_num[k],))[e]
                continue
            comp, i, n = key
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
        kwargs_ps = []
        for k, n in enumerate(self._ps_num):
            kw = {}
            for name in ("ra_image", "dec_image", "point_amp"):
                if name in self._ps_fixed[k]:
                    kw[name] = np.broadcast_to(np.asarray(self._ps_fixed[k][name], dtype=float), (n,)).copy()
                else:
                    cols = [self._index[("point_source", k, name, j)] for j in range(n)]
                    kw[name] = args[..., cols]
            kwargs_ps.append(kw)
        out.update(kwargs_ps=kwargs_ps, kwargs_special={}, kwargs_extinction=[], kwargs_tracer_source=[])
        return out

    def kwargs2args(self, kwargs_lens=None, kwargs_source=None, kwargs_lens_light=None, kwargs_ps=None, **unused):
        given = {"lens": kwargs_lens or [], "source": kwargs_source or [], "lens_light": kwargs_lens_light or []}
        args = np.empty(len(self._flat_names))
        for key, j in self._index.items():
            if key[0] == "point_source":
                _, k, n, e = key
                args[j] = np.atleast_1d(kwargs_ps[k][n])[e]
            else:
                comp, i, n = key
                args[j] = given[comp][i][n]
        return args

    def column_source_fn(self, index_lists=None):
        """-> source_fn(component, i_local, names) for LensModel.lens_descs / LightModel.light_descs.
        ``index_lists``: {"lens": [...], "source": [...], "lens_light": [...]} local -> global profile index of a
        band's sub-model (SingleBandMultiModel.select_kwargs); None = identity."""

        def fn(component, i_local, names, element=None):
            if component == "point_source":  # one parameter (`names`) of image `element` of model i_local
                fixed = self._ps_fixed[i_local]
                if names in fixed:
                    return ("cst", float(np.broadcast_to(np.asarray(fixed[names], dtype=float),
                                                         (self._ps_num[i_local],))[element]))
                return ("col", self._index[("point_source", i_local, names, element)])
            idx = (index_lists or {}).get(component)
            i = i_local if idx is None else idx[i_local]
            src = []
            for n in names:
                if n in self._fixed[component][i]:
                    src.append(("cst", float(self._fixed[component][i][n])))
                else:
                    src.append(("col", self._index[(component, i, n)]))
            return src

        fn.num_point_images = list(self._ps_num)
        return fn
