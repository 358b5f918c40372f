"""ImageModel façade, jaxtronomy/ImSim/image_model.py:23-694: same constructor and method signatures for the
extended-source path; every image / likelihood call is one lsim_image / lsim_logl launch sequence.

Batching: any kwargs value may be a ``[B]`` array instead of a scalar; outputs then carry a leading B axis
(the reference gets the same with ``jax.vmap``).  ``image_batch`` / ``logL_batch`` take a ready-made ``[B, P]``
parameter matrix in the column order of ``self.param_layout.names`` and skip all Python-side packing.
"""

import numpy as np

from jaxtronomy_b200 import _arrays, _lib, _model
from jaxtronomy_b200.ImSim.Numerics.numerics import NumericsSubFrame
from jaxtronomy_b200.LensModel.lens_model import LensModel
from jaxtronomy_b200.LightModel.light_model import LightModel
from jaxtronomy_b200.Util import util

__all__ = ["ImageModel"]


class ImageModel(object):
    """This class uses functions of lens_model and source_model to make a lensed image."""

    def __init__(self, data_class, psf_class, lens_model_class=None, source_model_class=None,
                 lens_light_model_class=None, point_source_class=None, extinction_class=None, kwargs_numerics=None,
                 likelihood_mask=None, psf_error_map_bool_list=None, kwargs_pixelbased=None):
        self.type = "single-band"
        self.num_bands = 1
        self.PSF = psf_class
        self.Data = data_class
        self._flux_scaling = getattr(self.Data, "flux_scaling", 1)
        self.PSF.set_pixel_size(self.Data.pixel_width)
        if kwargs_numerics is None:
            kwargs_numerics = {}
        self.ImageNumerics = NumericsSubFrame(pixel_grid=self.Data, psf=self.PSF, **kwargs_numerics)
        if lens_model_class is None:
            lens_model_class = LensModel(lens_model_list=[])
        self.LensModel = lens_model_class
        if point_source_class is None:
            from jaxtronomy_b200.PointSource.point_source import PointSource

            point_source_class = PointSource(point_source_type_list=[])
        self.PointSource = point_source_class
        self._n_ps_models = len(self.PointSource.point_source_type_list)
        if psf_error_map_bool_list is None:
            psf_error_map_bool_list = [True] * self._n_ps_models
        self._psf_error_map_bool_list = [bool(b) for b in psf_error_map_bool_list]
        self._psf_error_map = bool(getattr(self.PSF, "psf_variance_map_bool", False))  # image_model.py:127-132
        # images per point-source model: fixes the column layout of the batch entry points; the kwargs entry points
        # read it from the lengths of ra_image
        self._num_point_images = [0] * self._n_ps_models
        if source_model_class is None:
            source_model_class = LightModel(light_model_list=[])
        self.SourceModel = source_model_class
        if lens_light_model_class is None:
            lens_light_model_class = LightModel(light_model_list=[])
        self.LensLightModel = lens_light_model_class
        self._kwargs_numerics = kwargs_numerics
        if extinction_class is not None:
            raise ValueError("Extinction is not implemented in JAXtronomy yet")
        if kwargs_pixelbased is not None:
            raise ValueError("pixelbased solver not supported in JAXtronomy")
        self._pixelbased_bool = False
        if likelihood_mask is None:
            likelihood_mask = np.ones(data_class.num_pixel_axes)
        self.likelihood_mask = np.array(likelihood_mask, dtype=bool)
        self.num_data_evaluate = int(np.sum(self.likelihood_mask))
        self._mask1d = util.image2array(self.likelihood_mask)
        if data_class.primary_beam is not None:
            raise ValueError("primary beam not supported in jaxtronomy")
        # one column per parameter, in the order lens -> source -> lens light -> point sources (Param.args2kwargs)
        self.param_layout = self._full_layout(self._num_point_images)
        self._plans = {}

    # ----------------------------------------------------------------------------------------------------
    # plans
    # ----------------------------------------------------------------------------------------------------
    def _data_dict(self):
        """Data / noise / mask arrays on the frame the kernels work on (the evaluated sub-frame when
        flux_evaluate_indexes is set).  Likelihood pixels outside the sub-frame see a model of exactly zero; their
        chi^2 does not depend on the parameters and enters as a constant offset."""
        d, num = self.Data, self.ImageNumerics
        nm2 = None if d._noise_map is None else d._noise_map**2
        full = lambda a: None if a is None else np.broadcast_to(np.asarray(a, dtype=float), d.data.shape)  # noqa: E731
        exposure = full(getattr(d, "exposure_map", None))
        offset = 0.0
        if num.subframe_calc:
            outside = np.array(self.likelihood_mask, dtype=bool)
            inside = np.zeros_like(outside)
            num.cut_frame(inside)[...] = True
            outside &= ~inside
            cd0 = nm2 if nm2 is not None else np.full(d.data.shape, float(d.background_rms) ** 2)
            offset = -0.5 * float(np.sum((d.data**2 / cd0)[outside]))
        cut = num.cut_frame
        mask = cut(self.likelihood_mask)
        return dict(data=cut(d.data), cd_data=cut(np.asarray(d.C_D)), noise_map2=cut(nm2), exposure=cut(exposure),
                    background_rms=d.background_rms, mask=None if bool(np.all(mask)) else mask, logl_offset=offset)

    # -- point sources ----------------------------------------------------------------------------------------
    @property
    def num_point_source_images(self):
        return list(self._num_point_images)

    @num_point_source_images.setter
    def num_point_source_images(self, num_images):
        """Images per point-source model for the [B, P] batch entry points (image_batch, logL_batch, the sampler
        plans): the parameter matrix then carries ra_image[0..n), dec_image[0..n), point_amp[0..n) per model after
        the lens, source and lens-light columns."""
        num_images = [int(n) for n in num_images]
        if len(num_images) != self._n_ps_models:
            raise ValueError("%d image counts given for %d point-source models" % (len(num_images), self._n_ps_models))
        if num_images != self._num_point_images:
            self._num_point_images = num_images
            self._plans = {}
            self.param_layout = self._full_layout(num_images)

    def _full_layout(self, num_images):
        layout = _model.ColumnLayout()
        self.LensModel.lens_descs(layout)
        self.SourceModel.light_descs(layout, "source", True, [0] * self.SourceModel._num_func, True)
        self.LensLightModel.light_descs(layout, "lens_light", False, [0] * self.LensLightModel._num_func, True)
        self.PointSource.point_descs(layout, num_images, [0] * sum(num_images), True)
        return layout

    def _ps_dict(self, unconvolved):
        """Rendering constants of PointSourceRendering (point_source_rendering.py:13-62)."""
        if self.ImageNumerics.subframe_calc:
            raise ValueError("point sources together with flux_evaluate_indexes sub-frames are not supported")
        ss = self.ImageNumerics.point_source_supersampling_factor
        if unconvolved:
            kernel = np.zeros((3, 3))
            kernel[1, 1] = 1
        else:
            kernel = self.PSF.kernel_point_source_supersampled(ss)
        x0, y0 = self.Data.map_coord2pix(0.0, 0.0)
        out = dict(kernel=kernel, supersampling=ss, x_at_radec_0=x0, y_at_radec_0=y0,
                   Ma2pix=self.Data.transform_angle2pix, scale=self._flux_scaling)
        if self._psf_error_map and any(self._psf_error_map_bool_list):
            k = np.asarray(self.PSF.kernel_point_source)
            c = int((len(k) - 0.5) / 2)  # kernel_util.estimate_amp, Util/kernel_util.py:42-44
            out.update(variance_map=self.PSF.psf_variance_map, mean_kernel=float(np.sum(k[c - 2:c + 3, c - 2:c + 3])))
        return out

    def _build_plan(self, source_slots, lens_light_slots, n_slots, use_amp, unconvolved, lensed=True,
                    source_fn=None, n_cols=None, ps=None):
        """ps: None or dict(num_images, slots (flat over images, -1 = left out), k)"""
        layout = _model.ColumnLayout()
        lens_descs = self.LensModel.lens_descs(layout, source_fn=source_fn)  # always registers the lens columns
        if not lensed:
            lens_descs = []
        light = self.SourceModel.light_descs(layout, "source", lensed=lensed, slots=source_slots, use_amp=use_amp,
                                             source_fn=source_fn)
        light += self.LensLightModel.light_descs(layout, "lens_light", lensed=False, slots=lens_light_slots,
                                                 use_amp=use_amp, source_fn=source_fn)
        points = None
        if ps is not None and sum(ps["num_images"]) > 0:
            points = self._ps_dict(unconvolved)
            points["descs"] = self.PointSource.point_descs(
                layout, ps["num_images"], ps["slots"], use_amp, k=ps.get("k"), source_fn=source_fn,
                error_map_bool_list=self._psf_error_map_bool_list if self._psf_error_map else None)
        return _lib.Plan(layout.n_cols if n_cols is None else n_cols, lens_descs=lens_descs, light_descs=light,
                         n_slots=n_slots, points=points,
                         grid=self.ImageNumerics.grid_dict(), psf=self.ImageNumerics.psf_dict(unconvolved),
                         image_scale=self.ImageNumerics.pixel_area * self._flux_scaling, data=self._data_dict(),
                         **self.LensModel.plan_kwargs())

    def _joint_plan(self, source_bool, lens_light_bool, unconvolved=False, de_lensed=False, num_images=None,
                    ps_add=True, ps_k=None):
        """All selected components summed with their amplitudes into one image.  The parameter matrix always
        carries the point-source columns of ``num_images`` (default: the model's current image counts)."""
        num_images = list(self._num_point_images if num_images is None else num_images)
        key = ("joint", tuple(source_bool), tuple(lens_light_bool), bool(unconvolved), bool(de_lensed),
               tuple(num_images), bool(ps_add), ps_k)
        if key not in self._plans:
            ps = dict(num_images=num_images, slots=[0 if ps_add else -1] * sum(num_images), k=ps_k)
            self._plans[key] = self._build_plan([0 if b else -1 for b in source_bool],
                                                [0 if b else -1 for b in lens_light_bool], 1, True, unconvolved,
                                                lensed=not de_lensed, ps=ps)
        return self._plans[key]

    def sampler_plan(self, source_fn, n_cols, linear_solver):
        """Plan for the sampler boundary: ``source_fn`` maps every model parameter to a column of the sampler's
        `positions` matrix or to a constant (Param.args2kwargs, Sampling/likelihood.py:305)."""
        ns, nl = self.SourceModel._num_func, self.LensLightModel._num_func
        num_images = list(getattr(source_fn, "num_point_images", [0] * self._n_ps_models))
        if len(num_images) != self._n_ps_models:
            raise ValueError("the Param class describes %d point-source models, the image model has %d"
                             % (len(num_images), self._n_ps_models))
        n_ps = sum(num_images)
        if linear_solver:
            ps_slots, n_ps_lin = self.PointSource.linear_slots(num_images, ns + nl)
            n_slots = max(ns + nl + n_ps_lin, 1)
            if n_slots > _lib.MAX_SLOTS:
                raise ValueError("at most %d linear components are supported" % _lib.MAX_SLOTS)
            ps = dict(num_images=num_images, slots=ps_slots, k=None)
            return self._build_plan(list(range(ns)), [ns + i for i in range(nl)], n_slots, False, False,
                                    source_fn=source_fn, n_cols=n_cols, ps=ps)
        ps = dict(num_images=num_images, slots=[0] * n_ps, k=None)
        return self._build_plan([0] * ns, [0] * nl, 1, True, False, source_fn=source_fn, n_cols=n_cols, ps=ps)

    def sampler_plans(self, param_class, n_cols, linear_solver, index_lists=None):
        """List of plans whose log-likelihoods add up to this model's (one entry for a single band)."""
        return [self.sampler_plan(param_class.column_source_fn(index_lists), n_cols, linear_solver)]

    def _pack(self, kwargs_lens, kwargs_source, kwargs_lens_light, kwargs_ps=None, num_images=None):
        values = _model.gather_values(self.LensModel.param_name_list, kwargs_lens or [], "lens")
        values += _model.gather_values(self.SourceModel.param_name_list, kwargs_source or [], "source light")
        values += _model.gather_values(self.LensLightModel.param_name_list, kwargs_lens_light or [], "lens light")
        if num_images is not None and sum(num_images) > 0:
            values += self.PointSource.gather_values(kwargs_ps, num_images)
        return _model.pack_columns(values)

    def _ps_counts(self, kwargs_ps):
        """Images per point-source model of a kwargs call ([] / None: the model has no point sources)."""
        if self._n_ps_models == 0:
            return []
        if not kwargs_ps:
            raise ValueError("kwargs_ps is required: the model has %d point-source model(s)" % self._n_ps_models)
        return self.PointSource.num_images(kwargs_ps)

    def _dummy_ps(self, num_images):
        return [dict(ra_image=np.zeros(n), dec_image=np.zeros(n), point_amp=np.zeros(n)) for n in num_images]

    @staticmethod
    def _want_torch(*kwargs_lists):
        for kl in kwargs_lists:
            for kw in kl or []:
                for v in kw.values():
                    if _arrays.is_torch(v):
                        return True
        return False

    @staticmethod
    def _finish(t, batched, want_torch):
        if not batched:
            t = t[0]
        return t if want_torch else t.cpu().numpy()

    # ----------------------------------------------------------------------------------------------------
    # images
    # ----------------------------------------------------------------------------------------------------
    def image(self, kwargs_lens=None, kwargs_source=None, kwargs_lens_light=None, kwargs_ps=None,
              kwargs_extinction=None, kwargs_special=None, unconvolved=False, source_add=True, lens_light_add=True,
              point_source_add=True):
        """Make an image with a realisation of linear parameter values (image_model.py:431-492)."""
        if kwargs_extinction:
            raise ValueError("Extinction is not implemented in JAXtronomy yet")
        counts = self._ps_counts(kwargs_ps)
        mat, B, batched = self._pack(kwargs_lens, kwargs_source, kwargs_lens_light, kwargs_ps, counts)
        plan = self._joint_plan([bool(source_add)] * self.SourceModel._num_func,
                                [bool(lens_light_add)] * self.LensLightModel._num_func, unconvolved,
                                num_images=counts, ps_add=bool(point_source_add))
        out = self.ImageNumerics.complete_frame(plan.image(mat)[:, 0])
        return self._finish(out, batched, self._want_torch(kwargs_lens, kwargs_source, kwargs_lens_light))

    def point_source(self, kwargs_ps, kwargs_lens=None, kwargs_special=None, unconvolved=False, k=None):
        """Point-source images painted with the PSF at their sub-pixel positions (image_model.py:395-429)."""
        if kwargs_special:
            raise ValueError("astrometric displacements (kwargs_special) are outside the scope of jaxtronomy_b200")
        if self._n_ps_models == 0:
            return np.zeros(self.Data.num_pixel_axes)
        counts = self._ps_counts(kwargs_ps)
        mat, B, batched = self._pack(self._dummy_lens(), self._dummy_source(), self._dummy_lens_light(), kwargs_ps,
                                     counts)
        plan = self._joint_plan([False] * self.SourceModel._num_func, [False] * self.LensLightModel._num_func,
                                unconvolved, num_images=counts, ps_add=True, ps_k=k)
        return self._finish(plan.image(mat)[:, 0], batched, False)

    def source_surface_brightness(self, kwargs_source, kwargs_lens=None, kwargs_extinction=None, kwargs_special=None,
                                  unconvolved=False, de_lensed=False, k=None, update_pixelbased_mapping=True):
        """image_model.py:223-368."""
        if len(self.SourceModel.profile_type_list) == 0:
            return np.zeros(self.Data.num_pixel_axes)
        if kwargs_extinction:
            raise ValueError("Extinction is not implemented in JAXtronomy yet")
        mat, B, batched = self._pack(kwargs_lens, kwargs_source, self._dummy_lens_light())
        plan = self._joint_plan(self.SourceModel._bool_list(k), [False] * self.LensLightModel._num_func, unconvolved,
                                de_lensed=de_lensed, num_images=[0] * self._n_ps_models)
        return self._finish(self.ImageNumerics.complete_frame(plan.image(mat)[:, 0]), batched,
                            self._want_torch(kwargs_lens, kwargs_source))

    def lens_surface_brightness(self, kwargs_lens_light, unconvolved=False, k=None):
        """image_model.py:370-393."""
        mat, B, batched = self._pack(self._dummy_lens(), self._dummy_source(), kwargs_lens_light)
        plan = self._joint_plan([False] * self.SourceModel._num_func, self.LensLightModel._bool_list(k), unconvolved,
                                num_images=[0] * self._n_ps_models)
        return self._finish(self.ImageNumerics.complete_frame(plan.image(mat)[:, 0]), batched,
                            self._want_torch(kwargs_lens_light))

    def _dummy(self, names_list):
        return [{n: 1.0 for n in names} for names in names_list]

    def _dummy_lens(self):
        return self._dummy(self.LensModel.param_name_list)

    def _dummy_source(self):
        return self._dummy(self.SourceModel.param_name_list)

    def _dummy_lens_light(self):
        return self._dummy(self.LensLightModel.param_name_list)

    # ----------------------------------------------------------------------------------------------------
    # likelihood
    # ----------------------------------------------------------------------------------------------------
    def likelihood_data_given_model(self, kwargs_lens=None, kwargs_source=None, kwargs_lens_light=None,
                                    kwargs_ps=None, kwargs_extinction=None, kwargs_special=None, source_marg=False,
                                    linear_prior=None, check_positive_flux=False):
        """log-likelihood of the data given the model, amplitudes taken from the kwargs
        (image_model.py:162-221)."""
        if check_positive_flux:
            raise ValueError("check positive flux is not supported in jaxtronomy")
        if kwargs_special:
            raise ValueError("kwargs_special is outside the scope of jaxtronomy_b200")
        counts = self._ps_counts(kwargs_ps)
        mat, B, batched = self._pack(kwargs_lens, kwargs_source, kwargs_lens_light, kwargs_ps, counts)
        plan = self._joint_plan([True] * self.SourceModel._num_func, [True] * self.LensLightModel._num_func,
                                num_images=counts)
        logl = plan.logl(mat, linear_solve=False)[0]
        return self._finish(logl, batched, self._want_torch(kwargs_lens, kwargs_source, kwargs_lens_light))

    def logL_batch(self, params, out=None):
        """[B, P] parameter matrix (columns = self.param_layout order: lens, source, lens light) -> logL [B]
        as a CUDA tensor on the current stream."""
        plan = self._joint_plan([True] * self.SourceModel._num_func, [True] * self.LensLightModel._num_func)
        return plan.logl(params, linear_solve=False, out=out)[0]

    def logL_grad_batch(self, params):
        """[B, P] -> (logL [B], d logL / d params [B, P]) as CUDA tensors: what jax.value_and_grad of
        likelihood_data_given_model gives the reference (lsim_logl_grad; amplitudes are parameters here)."""
        plan = self._joint_plan([True] * self.SourceModel._num_func, [True] * self.LensLightModel._num_func)
        logl, grad, _ = plan.logl_grad(params, linear_solve=False)
        return logl, grad

    def image_batch(self, params):
        """[B, P] -> model images [B, ny, nx] (CUDA tensor)."""
        plan = self._joint_plan([True] * self.SourceModel._num_func, [True] * self.LensLightModel._num_func)
        return self.ImageNumerics.complete_frame(plan.image(params)[:, 0])

    @property
    def param_names(self):
        """(component, profile index, parameter name) of every column of the batch parameter matrix."""
        return self.param_layout.names

    # ----------------------------------------------------------------------------------------------------
    # masks (host bookkeeping, bit-exact: image_model.py:542-574)
    # ----------------------------------------------------------------------------------------------------
    def image2array_masked(self, image):
        if _arrays.is_torch(image):
            import torch

            return image.reshape(image.shape[:-2] + (-1,))[..., torch.as_tensor(self._mask1d, device=image.device)]
        image = np.asarray(image)
        return image.reshape(image.shape[:-2] + (-1,))[..., self._mask1d]

    def array_masked2image(self, array):
        nx, ny = self.Data.num_pixel_axes
        array = np.asarray(array)
        grid1d = np.zeros(array.shape[:-1] + (nx * ny,))
        grid1d[..., self._mask1d] = array
        return grid1d.reshape(array.shape[:-1] + (ny, nx))

    @property
    def data_response(self):
        return self.image2array_masked(self.Data.data)

    def _error_map_psf(self, kwargs_lens, kwargs_ps, kwargs_special=None):
        """sigma^2 expected from inaccuracies of the PSF model at the point-source positions
        (image_model.py:615-648); zeros without a psf_variance_map."""
        if not (self._psf_error_map and self._n_ps_models and kwargs_ps):
            return np.zeros(self.Data.num_pixel_axes)
        counts = self._ps_counts(kwargs_ps)
        mat, B, batched = self._pack(self._dummy_lens(), self._dummy_source(), self._dummy_lens_light(), kwargs_ps,
                                     counts)
        plan = self._joint_plan([True] * self.SourceModel._num_func, [True] * self.LensLightModel._num_func,
                                num_images=counts)
        return self._finish(plan.psf_error_map(mat), batched, False)

    def _error_map_model(self, kwargs_lens, kwargs_ps, kwargs_special=None):
        return self._error_map_psf(kwargs_lens, kwargs_ps, kwargs_special)

    def error_response(self, kwargs_lens, kwargs_ps, kwargs_special):
        """C_D of the data on the unmasked pixels + model error (the PSF error map), image_model.py:576-595."""
        model_error = self._error_map_model(kwargs_lens, kwargs_ps, kwargs_special)
        return self.image2array_masked(self.Data.C_D + model_error), model_error

    def update_psf(self, psf_class):
        raise ValueError("Updating psf class not supported in jaxtronomy. Create a new instance of ImageModel instead.")

    def update_data(self, data_class):
        raise ValueError("Updating data class not supported in jaxtronomy. Create a new instance of ImageModel instead.")
