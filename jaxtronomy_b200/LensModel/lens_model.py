"""LensModel façade: same constructor / method signatures as jaxtronomy/LensModel/lens_model.py:20-252 for the
ray-shooting path; the arithmetic runs in lsim_ray_shoot (k_ray_shoot)."""

import numpy as np

from jaxtronomy_b200 import _arrays, _lib, _model
from jaxtronomy_b200.Cosmo.background import Background
from jaxtronomy_b200.Util.util import convert_bool_list

__all__ = ["LensModel"]


class LensModel(object):
    """Class to handle an arbitrary list of lens models (single plane or multi plane)."""

    def __init__(self, lens_model_list, z_lens=None, z_source=None, lens_redshift_list=None, cosmo=None,
                 multi_plane=False, observed_convention_index=None, z_source_convention=None, cosmo_interp=False,
                 z_interp_stop=None, num_z_interp=100, profile_kwargs_list=None, decouple_multi_plane=False,
                 kwargs_multiplane_model=None, distance_ratio_sampling=False, cosmology_sampling=False,
                 cosmology_model="FlatLambdaCDM"):
        if cosmology_sampling:
            raise ValueError("Cosmology sampling not supported in JAXtronomy")
        if distance_ratio_sampling:
            raise ValueError("Distance ratio sampling not supported in JAXtronomy")
        if decouple_multi_plane:
            raise ValueError("decoupled multi-plane is outside the scope of jaxtronomy_b200")
        if observed_convention_index is not None:
            raise ValueError("observed_convention_index is outside the scope of jaxtronomy_b200")
        for name in lens_model_list:
            if name in ("LOS", "LOS_MINIMAL"):
                raise ValueError("line-of-sight models are outside the scope of jaxtronomy_b200")
            if name not in _lib.LENS_TYPES:
                raise ValueError("%s is not a valid lens model. Supported are: %s." % (name, sorted(_lib.LENS_TYPES)))
        self.lens_model_list = list(lens_model_list)
        self.z_lens, self.z_source = z_lens, z_source
        self._z_source_convention = z_source_convention
        self.redshift_list = lens_redshift_list
        if profile_kwargs_list is None:
            profile_kwargs_list = [{} for _ in range(len(lens_model_list))]
        self.profile_kwargs_list = [pk or {} for pk in profile_kwargs_list]
        self.cosmo = cosmo
        self.multi_plane = bool(multi_plane)
        self.param_name_list = [_model.LENS_PARAM_NAMES[n] for n in self.lens_model_list]
        self._num_func = len(self.lens_model_list)
        self._plans = {}
        self.alpha_scaling = 1.0
        if self.multi_plane:
            if lens_redshift_list is None:
                raise ValueError("In multi-plane lensing, you need to specify the redshifts of the lensing planes.")
            if z_source is None:
                raise ValueError("z_source needs to be set for multi-plane lens modelling.")
            if len(lens_redshift_list) != len(lens_model_list):
                raise ValueError("The length of lens_model_list does not correspond to redshift_list")
            self.type = "MultiPlane"
            self._set_multi_plane(Background(cosmo))
        else:
            self.type = "SinglePlane"
            if z_source is not None and z_source_convention is not None and z_source != z_source_convention:
                if z_lens is None:
                    raise ValueError("a lens redshift needs to provided when z_source != z_source_convention")
                bkg = Background(cosmo)
                # LensCosmo.beta_double_source_plane: D_ds1 / D_s1 * D_s2 / D_ds2
                self.alpha_scaling = (bkg.d_xy(z_lens, z_source) / bkg.d_xy(0, z_source)
                                      * bkg.d_xy(0, z_source_convention) / bkg.d_xy(z_lens, z_source_convention))

    # -- multi-plane host constants: MultiPlaneBase.set_T_zs_and_T_ijs (multi_plane_base.py:89-115),
    #    MultiPlane._set_source_distances (multi_plane.py:190-204)
    def _set_multi_plane(self, bkg):
        z_list = list(self.redshift_list)
        z_conv = self._z_source_convention if self._z_source_convention is not None else self.z_source
        if len(z_list) > 0 and np.max(z_list) >= z_conv:
            raise ValueError("deflector redshifts higher or equal the source redshift convention (%s >= %s for the "
                             "reduced lens model quantities not allowed (leads to negative reduced deflection "
                             "angles!" % (np.max(z_list), z_conv))
        order = [int(i) for i in np.argsort(np.array(z_list))]
        self._sorted_redshift_index = order
        self._T_ij_list, self._T_z_list, self._red2phys = [], [], []
        z_before, T_z = 0, 0
        for idx in order:
            z = z_list[idx]
            if z_before == z:
                delta_T = 0
            else:
                T_z = bkg.T_xy(0, z)
                delta_T = bkg.T_xy(z_before, z)
            self._T_ij_list.append(delta_T)
            self._T_z_list.append(T_z)
            self._red2phys.append(bkg.d_xy(0, z_conv) / bkg.d_xy(z, z_conv))
            z_before = z
        z_last = 0
        for idx in order:
            if 0 < z_list[idx] <= self.z_source:
                z_last = z_list[idx]
        self._T_ij_stop = bkg.T_xy(z_last, self.z_source)
        self._T_z_source = bkg.T_xy(0, self.z_source)

    def set_multi_plane_constants(self, T_ij_list, T_z_list, reduced2physical, T_ij_stop, T_z_source):
        """Override the cosmological constants (in sorted-redshift order), e.g. with astropy-derived values."""
        self._T_ij_list, self._T_z_list = list(map(float, T_ij_list)), list(map(float, T_z_list))
        self._red2phys = list(map(float, reduced2physical))
        self._T_ij_stop, self._T_z_source = float(T_ij_stop), float(T_z_source)
        self._plans = {}

    def info(self):
        for i, name in enumerate(self.lens_model_list):
            print("Lens model %s is %s with parameters %s" % (i, name, self.param_name_list[i]))

    def check_parameters(self, kwargs_list):
        _model.gather_values(self.param_name_list, kwargs_list, "lens")

    def _bool_list(self, k=None):
        return convert_bool_list(n=self._num_func, k=k)

    # -- descriptors (shared with ImageModel) -----------------------------------------------------------
    def lens_descs(self, layout, bool_list=None, source_fn=None):
        """Append this model's parameters to `layout` (one column each) and return the descriptors of the
        selected profiles, in evaluation order (redshift-sorted for multi plane).  With ``source_fn(component, i,
        names) -> [("col", j) | ("cst", value), ...]`` the caller decides where every parameter comes from
        (sampler boundary: sampled parameters are columns of `positions`, fixed ones constants)."""
        if source_fn is not None:
            sources = [source_fn("lens", i, names) for i, names in enumerate(self.param_name_list)]
        else:
            sources = [layout.add_profile("lens", i, names) for i, names in enumerate(self.param_name_list)]
        if bool_list is None:
            bool_list = [True] * self._num_func
        descs = []
        if self.multi_plane:
            for pos, i in enumerate(self._sorted_redshift_index):
                if not (0 < self.redshift_list[i] <= self.z_source):
                    continue
                descs.append(_model.lens_desc(self.lens_model_list[i], sources[i], self.profile_kwargs_list[i],
                                              delta_T=self._T_ij_list[pos], T_z=self._T_z_list[pos],
                                              red2phys=self._red2phys[pos]))
        else:
            for i, name in enumerate(self.lens_model_list):
                if bool_list[i]:
                    descs.append(_model.lens_desc(name, sources[i], self.profile_kwargs_list[i]))
        return descs

    def plan_kwargs(self):
        if self.multi_plane:
            return dict(multi_plane=True, T_ij_end=self._T_ij_stop, T_z_source=self._T_z_source)
        return dict(multi_plane=False, alpha_scaling=self.alpha_scaling)

    def _plan(self, k):
        if self.multi_plane and k is not None:
            raise ValueError("no specific selection of a subset of lens models supported in multi-plane mode. Please"
                             "use single plane mode or generate new instance of LensModel of the subset of profiles.")
        bl = tuple(self._bool_list(k))
        if bl not in self._plans:
            layout = _model.ColumnLayout()
            descs = self.lens_descs(layout, bl)
            self._plans[bl] = _lib.Plan(layout.n_cols, lens_descs=descs, **self.plan_kwargs())
        return self._plans[bl]

    def _run(self, x, y, kwargs, k, what):
        xf, yf, shape, want_torch = _arrays.flatten_coords(x, y)
        values = _model.gather_values(self.param_name_list, kwargs, "lens")
        mat, B, batched = _model.pack_columns(values)
        ox, oy = self._plan(k).ray_shoot(mat, xf, yf, what=what)
        return _arrays.finish(ox, shape, batched, want_torch), _arrays.finish(oy, shape, batched, want_torch)

    def ray_shooting(self, x, y, kwargs, k=None):
        """Maps image to source position (inverse deflection), lens_model.py:236-252."""
        return self._run(x, y, kwargs, k, what=0)

    def alpha(self, x, y, kwargs, k=None):
        """Deflection angles, lens_model.py (alpha) / single_plane.py:112-139."""
        if self.multi_plane:  # reduced deflection theta - beta, MultiPlane/multi_plane.py:512-534
            x, y = (v if _arrays.is_torch(v) else np.asarray(v, dtype=np.float64) for v in (x, y))
            bx, by = self._run(x, y, kwargs, k, what=0)
            if _arrays.is_torch(bx):  # a CPU tensor in, a CUDA tensor out: subtract on the result's device
                x, y = (v.to(bx.device) if _arrays.is_torch(v) else v for v in (x, y))
            return x - bx, y - by
        return self._run(x, y, kwargs, k, what=1)

    # -- second derivatives (lens_model.py:384-527) ------------------------------------------------------------
    def hessian(self, x, y, kwargs, k=None, diff=None, diff_method="square"):
        """Hessian of the lensing potential (f_xx, f_xy, f_yx, f_yy), lens_model.py:384-414.

        ``diff=None``: the exact differentials, one `lsim_hessian` launch (single plane; the reference sums the
        profiles' own ``hessian`` functions, single_plane.py:142-180).  With ``diff`` set, the finite differences
        of the deflection field on a square or a cross around (x, y), lens_model.py:617-700 -- four `alpha`
        launches.  Like `alpha`, the deflection scaling of the model is applied for every ``k``.  Multi-plane models
        have no exact form: ``diff=None`` is the reference's forward difference with 1e-8 (`_hessian_multi_plane`)."""
        if diff is None:
            if self.multi_plane:
                return self._hessian_multi_plane(x, y, kwargs, k=k)
            xf, yf, shape, want_torch = _arrays.flatten_coords(x, y)
            values = _model.gather_values(self.param_name_list, kwargs, "lens")
            mat, B, batched = _model.pack_columns(values)
            h = self._plan(k).hessian(mat, xf, yf)
            return tuple(_arrays.finish(h[:, i], shape, batched, want_torch) for i in range(4))
        if diff_method == "square":
            return self._hessian_differential_square(x, y, kwargs, k=k, diff=diff)
        if diff_method == "cross":
            return self._hessian_differential_cross(x, y, kwargs, k=k, diff=diff)
        raise ValueError('diff_method %s not supported. Chose among "square" or "cross".' % diff_method)

    def _hessian_multi_plane(self, x, y, kwargs, k=None, diff=0.00000001):
        """MultiPlane.hessian, MultiPlane/multi_plane.py:537-589: FORWARD differences of alpha = theta - beta with
        diff = 1e-8 (three ray-shooting launches).  As the reference notes, the result carries the rounding of the ray
        shooting divided by diff (~1e-7); pass ``diff`` to `hessian` for the centred stencils."""
        x, y = (v if _arrays.is_torch(v) else np.asarray(v, dtype=np.float64) for v in (x, y))
        ax, ay = self.alpha(x, y, kwargs, k=k)
        ax_dx, ay_dx = self.alpha(x + diff, y, kwargs, k=k)
        ax_dy, ay_dy = self.alpha(x, y + diff, kwargs, k=k)
        f_xx = (ax_dx - ax) / diff
        f_xy = (ax_dy - ax) / diff
        f_yx = (ay_dx - ay) / diff
        f_yy = (ay_dy - ay) / diff
        return f_xx, f_xy, f_yx, f_yy

    def _hessian_differential_cross(self, x, y, kwargs, k=None, diff=0.00001):
        """Differentials of alpha along the cross centred at (x, y), lens_model.py:617-649."""
        x, y = (v if _arrays.is_torch(v) else np.asarray(v, dtype=np.float64) for v in (x, y))
        h = diff / 2
        ax_dx, ay_dx = self.alpha(x + h, y, kwargs, k=k)
        ax_dy, ay_dy = self.alpha(x, y + h, kwargs, k=k)
        ax_dx_, ay_dx_ = self.alpha(x - h, y, kwargs, k=k)
        ax_dy_, ay_dy_ = self.alpha(x, y - h, kwargs, k=k)
        f_xx = (ax_dx - ax_dx_) / diff
        f_xy = (ax_dy - ax_dy_) / diff
        f_yx = (ay_dx - ay_dx_) / diff
        f_yy = (ay_dy - ay_dy_) / diff
        return f_xx, f_xy, f_yx, f_yy

    def _hessian_differential_square(self, x, y, kwargs, k=None, diff=0.00001):
        """Differentials of alpha on the square around (x, y) (minimises the curl), lens_model.py:651-688."""
        x, y = (v if _arrays.is_torch(v) else np.asarray(v, dtype=np.float64) for v in (x, y))
        h = diff / 2
        ax_pp, ay_pp = self.alpha(x + h, y + h, kwargs, k=k)
        ax_pn, ay_pn = self.alpha(x + h, y - h, kwargs, k=k)
        ax_np, ay_np = self.alpha(x - h, y + h, kwargs, k=k)
        ax_nn, ay_nn = self.alpha(x - h, y - h, kwargs, k=k)
        f_xx = (ax_pp - ax_np + ax_pn - ax_nn) / diff / 2
        f_xy = (ax_pp - ax_pn + ax_np - ax_nn) / diff / 2
        f_yx = (ay_pp - ay_np + ay_pn - ay_nn) / diff / 2
        f_yy = (ay_pp - ay_pn + ay_np - ay_nn) / diff / 2
        return f_xx, f_xy, f_yx, f_yy

    def kappa(self, x, y, kwargs, k=None, diff=None, diff_method="square"):
        """Convergence 1/2 (f_xx + f_yy), lens_model.py:416-441."""
        f_xx, f_xy, f_yx, f_yy = self.hessian(x, y, kwargs, k=k, diff=diff, diff_method=diff_method)
        return 1.0 / 2 * (f_xx + f_yy)

    def curl(self, x, y, kwargs, k=None, diff=None, diff_method="square"):
        """f_xy - f_yx, lens_model.py:443-467."""
        f_xx, f_xy, f_yx, f_yy = self.hessian(x, y, kwargs, k=k, diff=diff, diff_method=diff_method)
        return f_xy - f_yx

    def gamma(self, x, y, kwargs, k=None, diff=None, diff_method="square"):
        """Shear (1/2 (f_xx - f_yy), f_xy), lens_model.py:469-497."""
        f_xx, f_xy, f_yx, f_yy = self.hessian(x, y, kwargs, k=k, diff=diff, diff_method=diff_method)
        return 1.0 / 2 * (f_xx - f_yy), f_xy

    def magnification(self, x, y, kwargs, k=None, diff=None, diff_method="square"):
        """1 / det(1 - Hessian), lens_model.py:499-526 (no guard against det = 0, like the reference)."""
        f_xx, f_xy, f_yx, f_yy = self.hessian(x, y, kwargs, k=k, diff=diff, diff_method=diff_method)
        det_A = (1 - f_xx) * (1 - f_yy) - f_xy * f_yx
        return 1.0 / det_A

    def flexion(self, x, y, kwargs, k=None, diff=0.0001, hessian_diff=False):
        """Third derivatives (f_xxx, f_xxy, f_xyy, f_yyy) as central differences of the Hessian over ``diff``,
        lens_model.py:529-570; ``hessian_diff=True`` also differences the Hessian itself (step diff / 4)."""
        x, y = (v if _arrays.is_torch(v) else np.asarray(v, dtype=np.float64) for v in (x, y))
        hd = diff / 4 if hessian_diff else None
        f_xx_dx, f_xy_dx, f_yx_dx, f_yy_dx = self.hessian(x + diff / 2, y, kwargs, k=k, diff=hd)
        f_xx_dy, f_xy_dy, f_yx_dy, f_yy_dy = self.hessian(x, y + diff / 2, kwargs, k=k, diff=hd)
        f_xx_dx_, f_xy_dx_, f_yx_dx_, f_yy_dx_ = self.hessian(x - diff / 2, y, kwargs, k=k, diff=hd)
        f_xx_dy_, f_xy_dy_, f_yx_dy_, f_yy_dy_ = self.hessian(x, y - diff / 2, kwargs, k=k, diff=hd)
        f_xxx = (f_xx_dx - f_xx_dx_) / diff
        f_xxy = (f_xx_dy - f_xx_dy_) / diff
        f_xyy = (f_xy_dy - f_xy_dy_) / diff
        f_yyy = (f_yy_dy - f_yy_dy_) / diff
        return f_xxx, f_xxy, f_xyy, f_yyy

    def set_static(self, kwargs):
        """lens_model.py:572-581.  The reference's profiles keep no static cache under jit (profile_list_base.py:246-254
        forwards to no-op `set_static`s); here every call prepares its parameter set on the device anyway: returns the
        keyword arguments unchanged."""
        self.check_parameters(kwargs)
        return kwargs

    def set_dynamic(self):
        """lens_model.py:583-592: nothing is cached, nothing to free."""
        return None
