"""PointSource façade, jaxtronomy/PointSource/point_source.py:10-500 with the types the reference supports on the
forward path: UNLENSED (PointSource/Types/unlensed.py) and LENSED_POSITION without fixed magnification
(PointSource/Types/lensed_position.py) -- image positions ``ra_image``/``dec_image`` and amplitudes ``point_amp`` are
model parameters.  SOURCE_POSITION, additional images and fixed magnification need the lens-equation solver /
magnification and raise, as SOURCE_POSITION and additional images do in the reference.

Host-side bookkeeping only: which parameter feeds which point-source image of which output slot.  The rendering is
``k_point_sources`` behind ``lsim_image`` / ``lsim_logl`` (include/lsim.h, ``lsim_point_desc``).
"""

import copy

import numpy as np

from jaxtronomy_b200 import _lib

__all__ = ["PointSource"]

_SUPPORTED_MODELS = ["UNLENSED", "LENSED_POSITION", "SOURCE_POSITION"]
PARAM_NAMES = ["ra_image", "dec_image", "point_amp"]


class PointSource(object):
    def __init__(self, point_source_type_list, lens_model=None, fixed_magnification_list=None,
                 additional_images_list=None, flux_from_point_source_list=None, magnification_limit=None,
                 save_cache=False, kwargs_lens_eqn_solver=None, index_lens_model_list=None,
                 point_source_frame_list=None, redshift_list=None):
        if save_cache:
            raise ValueError("saving cache is not supported in jaxtronomy")
        if index_lens_model_list is not None or point_source_frame_list is not None:
            raise ValueError("point-source frames (index_lens_model_list) are outside the scope of jaxtronomy_b200")
        self.point_source_type_list = list(point_source_type_list)
        n = len(self.point_source_type_list)
        fixed_magnification_list = fixed_magnification_list or [False] * n
        additional_images_list = additional_images_list or [False] * n
        if flux_from_point_source_list is None:
            flux_from_point_source_list = [True] * n
        self._flux_from_point_source_list = list(flux_from_point_source_list)
        self._fixed_magnification_list = list(fixed_magnification_list)
        self._lens_model = lens_model
        for i, model in enumerate(self.point_source_type_list):
            if model == "SOURCE_POSITION":
                raise ValueError("source position not supported in jaxtronomy")
            if model not in ("UNLENSED", "LENSED_POSITION"):
                raise ValueError("Point-source model %s not available. Supported models are %s ."
                                 % (model, _SUPPORTED_MODELS))
            if additional_images_list[i]:
                raise ValueError("additional images not supported in jaxtronomy")
            if fixed_magnification_list[i]:
                raise ValueError("fixed_magnification needs the lensing magnification (Hessian), which is outside the "
                                 "scope of jaxtronomy_b200")

    # -- positions / amplitudes straight from the keyword arguments ----------------------------------------------
    def image_position(self, kwargs_ps, kwargs_lens=None, k=None, original_position=False, additional_images=False):
        """point_source.py:205-253 -> (list of ra arrays, list of dec arrays)"""
        if additional_images:
            raise ValueError("additional images not supported in jaxtronomy")
        xs, ys = [], []
        for i in range(len(self.point_source_type_list)):
            if k is None or k == i:
                xs.append(np.atleast_1d(np.asarray(kwargs_ps[i]["ra_image"], dtype=float)))
                ys.append(np.atleast_1d(np.asarray(kwargs_ps[i]["dec_image"], dtype=float)))
        return xs, ys

    def source_position(self, kwargs_ps, kwargs_lens=None):
        """Intrinsic source positions, point_source.py:188-202: LENSED_POSITION = mean of the ray-shot image positions
        (Types/lensed_position.py:65-90, one `lsim_ray_shoot` launch per model), UNLENSED = the positions themselves
        (Types/unlensed.py).  -> (list of x_source, list of y_source)"""
        xs, ys = [], []
        for i, model in enumerate(self.point_source_type_list):
            ra = np.atleast_1d(np.asarray(kwargs_ps[i]["ra_image"], dtype=float))
            dec = np.atleast_1d(np.asarray(kwargs_ps[i]["dec_image"], dtype=float))
            if model == "LENSED_POSITION":
                if self._lens_model is None:
                    raise ValueError("LENSED_POSITION source positions need the lens model (lens_model=...)")
                bx, by = self._lens_model.ray_shooting(ra, dec, kwargs_lens)
                xs.append(np.mean(bx, axis=-1))
                ys.append(np.mean(by, axis=-1))
            else:
                xs.append(ra)
                ys.append(dec)
        return xs, ys

    def image_amplitude(self, kwargs_ps, kwargs_lens=None, k=None):
        """point_source.py:304-326"""
        amps = []
        for i in range(len(self.point_source_type_list)):
            if k is None or k == i:
                a = np.atleast_1d(np.asarray(kwargs_ps[i]["point_amp"], dtype=float))
                amps.append(a if self._flux_from_point_source_list[i] else np.zeros_like(a))
        return amps

    def point_source_list(self, kwargs_ps, kwargs_lens=None, k=None, with_amp=True):
        """All image positions and amplitudes in single arrays (point_source.py:255-284)."""
        ra_list, dec_list = self.image_position(kwargs_ps, kwargs_lens, k=k)
        amp_list = self.image_amplitude(kwargs_ps, kwargs_lens, k=k) if with_amp else [np.ones_like(r) for r in ra_list]
        cat = lambda lst: np.concatenate(lst) if lst else np.zeros(0)  # noqa: E731
        return cat(ra_list), cat(dec_list), cat(amp_list)

    def num_images(self, kwargs_ps):
        return [int(np.size(kw["ra_image"])) for kw in kwargs_ps]

    def linear_slots(self, num_images, first_slot):
        """Output slot per image, flat over (model, image): a running index from `first_slot` over the images of
        the models whose flux is part of the linear solve, -1 for the others (they own no basis image:
        point_source.py:296,361,398 skip them).  -> (slots, number of slots used)"""
        slots, used = [], 0
        for n, f in zip(num_images, self._flux_from_point_source_list):
            for _ in range(int(n)):
                if f:
                    slots.append(first_slot + used)
                    used += 1
                else:
                    slots.append(-1)
        return slots, used

    def num_basis(self, kwargs_ps, kwargs_lens=None):
        """point_source.py:286-302"""
        return sum(n for n, f in zip(self.num_images(kwargs_ps), self._flux_from_point_source_list) if f)

    def linear_response_set(self, kwargs_ps, kwargs_lens=None, with_amp=False):
        """One entry per linear basis = per image (point_source.py:347-384)."""
        ra_pos, dec_pos, amp = [], [], []
        xs, ys = self.image_position(kwargs_ps, kwargs_lens)
        amps = self.image_amplitude(kwargs_ps, kwargs_lens)
        for i in range(len(self.point_source_type_list)):
            if self._flux_from_point_source_list[i]:
                for j in range(len(xs[i])):
                    ra_pos.append(np.array([xs[i][j]]))
                    dec_pos.append(np.array([ys[i][j]]))
                    amp.append(np.array([amps[i][j] if with_amp else 1.0]))
        return ra_pos, dec_pos, amp, len(ra_pos)

    def update_linear(self, param, i, kwargs_ps, kwargs_lens=None):
        """point_source.py:386-409"""
        param = np.asarray(param)
        for k in range(len(self.point_source_type_list)):
            if self._flux_from_point_source_list[k]:
                n = int(np.size(kwargs_ps[k]["ra_image"]))
                kwargs_ps[k]["point_amp"] = np.array(param[i:i + n])
                i += n
        return kwargs_ps, i

    def linear_param_from_kwargs(self, kwargs_list):
        """point_source.py:411-430"""
        param = []
        for k in range(len(self.point_source_type_list)):
            if self._flux_from_point_source_list[k]:
                param += [a for a in np.atleast_1d(kwargs_list[k]["point_amp"])]
        return param

    def set_amplitudes(self, amp_list, kwargs_ps):
        """point_source.py:461-486"""
        kwargs_list = copy.deepcopy(kwargs_ps)
        for i in range(len(self.point_source_type_list)):
            if self._flux_from_point_source_list[i]:
                kwargs_list[i]["point_amp"] = amp_list[i]
        return kwargs_list

    @staticmethod
    def check_positive_flux(kwargs_ps):
        """point_source.py:488-502"""
        return bool(all(np.all(np.asarray(kw["point_amp"]) >= 0) for kw in kwargs_ps if "point_amp" in kw))

    # -- kernel descriptors --------------------------------------------------------------------------------------
    def point_descs(self, layout, num_images, slots, use_amp, k=None, source_fn=None, error_map_bool_list=None):
        """lsim_point_desc list for a plan.

        :param layout: _model.ColumnLayout; 3 n columns are registered per model: ra_image[0..n), dec_image[0..n),
            point_amp[0..n) (lenstronomy's Param order)
        :param num_images: images per point-source model (static: it fixes the column layout)
        :param slots: output slot per image, flat over (model, image), or -1 to leave an image out
        :param source_fn: None (every parameter is its own column) or the sampler's (component, index, name,
            element) -> ("col", j) | ("const", v) map
        """
        descs, flat = [], 0
        for i, n in enumerate(num_images):
            cols = layout.add_profile("point_source", i, ["%s[%d]" % (name, j) for name in PARAM_NAMES for j in range(n)])
            for j in range(n):
                d = _lib.PointDesc()
                srcs = []
                for t, name in enumerate(PARAM_NAMES):
                    src = cols[t * n + j] if source_fn is None else source_fn("point_source", i, name, j)
                    srcs.append(src)
                for (kind, v), col_attr, cst_attr in zip(srcs, ("col_ra", "col_dec", "col_amp"),
                                                         ("cst_ra", "cst_dec", "cst_amp")):
                    setattr(d, col_attr, int(v) if kind == "col" else -1)
                    setattr(d, cst_attr, 0.0 if kind == "col" else float(v))
                include = (k is None or k == i) and self._flux_from_point_source_list[i]
                d.slot = int(slots[flat]) if include else -1
                d.use_amp = 1 if use_amp else 0
                # ImageModel._error_map_psf loops over ALL models with a True flag (image_model.py:632-647)
                d.error_map = 1 if (error_map_bool_list is not None and error_map_bool_list[i]) else 0
                descs.append(d)
                flat += 1
        return descs

    def gather_values(self, kwargs_ps, num_images):
        """Flat value list in the column order of point_descs."""
        if len(kwargs_ps) != len(self.point_source_type_list):
            raise ValueError("length of kwargs_ps %d does not match the number of point-source models %d"
                             % (len(kwargs_ps), len(self.point_source_type_list)))
        values = []
        for kw, n in zip(kwargs_ps, num_images):
            for name in PARAM_NAMES:
                v = np.atleast_1d(np.asarray(kw.get(name, np.ones(n)), dtype=float))
                if v.size != n:
                    raise ValueError("%s holds %d values for %d point-source images" % (name, v.size, n))
                values += [float(x) for x in v]
        return values
