"""ImageLinearFit façade, jaxtronomy/ImSim/image_linear_solve.py:13-481 + de_lens.py: the linear amplitudes of all
light components are solved by weighted least squares inside the likelihood kernel (k_likelihood): the normal
equations M = A W A^T, R = A W d are reduced per CTA straight from the convolved basis images, the n_lin x n_lin
system is gated on cond_2(M) < 5/eps and solved in registers, and chi^2 is evaluated on sum_k B_k A_k -- the
response matrix A never exists as a [n_lin, N_d] array unless the caller asks for it.
"""

import numpy as np

from jaxtronomy_b200 import _arrays, _lib
from jaxtronomy_b200.ImSim.image_model import ImageModel

__all__ = ["ImageLinearFit"]


class ImageLinearFit(ImageModel):
    """Linear version class, inherits ImageModel."""

    def __init__(self, data_class, psf_class, lens_model_class=None, source_model_class=None,
                 lens_light_model_class=None, point_source_class=None, extinction_class=None, kwargs_numerics=None,
                 likelihood_mask=None, psf_error_map_bool_list=None, kwargs_pixelbased=None):
        super(ImageLinearFit, self).__init__(
            data_class, psf_class=psf_class, lens_model_class=lens_model_class, source_model_class=source_model_class,
            lens_light_model_class=lens_light_model_class, point_source_class=point_source_class,
            extinction_class=extinction_class, kwargs_numerics=kwargs_numerics, likelihood_mask=likelihood_mask,
            psf_error_map_bool_list=psf_error_map_bool_list, kwargs_pixelbased=kwargs_pixelbased)
        n = self.SourceModel._num_func + self.LensLightModel._num_func
        if n > _lib.MAX_SLOTS:
            raise ValueError("at most %d linear components are supported by jaxtronomy_b200, got %d" % (_lib.MAX_SLOTS, n))
        self._n_lin = n

    def _split_plan(self, unconvolved=False, num_images=None):
        """Every linear component in its own slot with amp = 1: source components first, then lens light, then one
        slot per point-source image (linear_response_matrix order, image_linear_solve.py:391-453)."""
        num_images = list(self._num_point_images if num_images is None else num_images)
        key = ("split", bool(unconvolved), tuple(num_images))
        if key not in self._plans:
            ns, n_light = self.SourceModel._num_func, self._n_lin
            ps_slots, n_ps = self.PointSource.linear_slots(num_images, n_light)
            n_slots = n_light + n_ps
            if n_slots > _lib.MAX_SLOTS:
                raise ValueError("at most %d linear components are supported by jaxtronomy_b200, got %d"
                                 % (_lib.MAX_SLOTS, n_slots))
            ps = dict(num_images=num_images, slots=ps_slots, k=None)
            self._plans[key] = self._build_plan(list(range(ns)), [ns + i for i in range(self.LensLightModel._num_func)],
                                                max(n_slots, 1), False, unconvolved, ps=ps)
        return self._plans[key]

    def num_param_linear(self, kwargs_lens=None, kwargs_source=None, kwargs_lens_light=None, kwargs_ps=None):
        counts = self.PointSource.num_images(kwargs_ps) if (kwargs_ps and self._n_ps_models) else self._num_point_images
        return self._n_lin + self.PointSource.linear_slots(counts, 0)[1]

    # -- batch entry points (CUDA tensors, no Python-side packing) ----------------------------------------------
    def logL_linear_batch(self, params, source_marg=False, return_model=False, out=None, num_images=None):
        """[B, P] parameter matrix -> (logL [B], amplitudes [B, n_lin], model [B, ny, nx] or None).  With point
        sources the matrix carries their columns for ``num_images`` (default: num_point_source_images)."""
        logl, amps, model = self._split_plan(num_images=num_images).logl(
            params, linear_solve=True, source_marg=source_marg, return_model=return_model, out=out)
        if model is not None:
            model = self.ImageNumerics.complete_frame(model)
        return logl, amps, model

    def logL_linear_grad_batch(self, params):
        """[B, P] -> (logL [B], d logL / d params [B, P], amplitudes [B, n_lin]) with the amplitudes solved per set
        and differentiated through (jax.value_and_grad of ImageLinearFit.likelihood_data_given_model in the
        reference; lsim_logl_grad).  The `amp` columns of the matrix are not read, their gradient is 0."""
        return self._split_plan().logl_grad(params, linear_solve=True)

    def response_batch(self, params, unconvolved=False, num_images=None):
        """[B, P] -> amp=1 basis images [B, n_lin, ny, nx] (full frame): convolved extended components, then the
        point-source images."""
        return self.ImageNumerics.complete_frame(self._split_plan(unconvolved, num_images).image(params))

    # -- reference API --------------------------------------------------------------------------------------------
    def linear_response_matrix(self, kwargs_lens, kwargs_source, kwargs_lens_light, kwargs_ps=None,
                               kwargs_extinction=None, kwargs_special=None, unconvolved=False):
        """Response matrix [n_lin, N_d] on the unmasked pixels (image_linear_solve.py:361-455)."""
        counts = self._ps_counts(kwargs_ps)
        mat, B, batched = self._pack(kwargs_lens, kwargs_source, kwargs_lens_light, kwargs_ps, counts)
        if self._n_lin + self.PointSource.linear_slots(counts, 0)[1] == 0:
            return np.zeros(((B,) if batched else ()) + (0, self.num_data_evaluate))
        import torch

        # flux scaling is part of image_scale / ps_scale; point sources keep the PSF kernel also when
        # unconvolved=True (image_linear_solve.py:444-449 calls point_source_rendering without the flag)
        images = torch.nan_to_num(self.response_batch(mat, unconvolved, counts))
        A = self.image2array_masked(images)
        return self._finish(A, batched, self._want_torch(kwargs_lens, kwargs_source, kwargs_lens_light))

    def image_linear_solve(self, kwargs_lens=None, kwargs_source=None, kwargs_lens_light=None, kwargs_ps=None,
                           kwargs_extinction=None, kwargs_special=None, inv_bool=False):
        """-> (model [ny, nx], model_error, cov_param, param): WLS amplitudes and the model they give
        (image_linear_solve.py:97-181).  ``cov_param`` (M^-1) is only formed when inv_bool is True."""
        counts = self._ps_counts(kwargs_ps)
        mat, B, batched = self._pack(kwargs_lens, kwargs_source, kwargs_lens_light, kwargs_ps, counts)
        want_torch = self._want_torch(kwargs_lens, kwargs_source, kwargs_lens_light)
        _, amps, model = self.logL_linear_batch(mat, return_model=True, num_images=counts)
        cov = None
        if inv_bool:
            cov = self._finish(self._cov_param(mat, counts), batched, want_torch)
        model_error = self._error_map_model(kwargs_lens, kwargs_ps, kwargs_special)
        return (self._finish(model, batched, want_torch), model_error, cov, self._finish(amps, batched, want_torch))

    def _cov_param(self, mat, num_images=None):
        """M^-1 = (A W A^T)^-1 with the cond gate of de_lens.py:26-27 (only for callers that want the covariance;
        the likelihood kernel never forms it)."""
        import torch

        images = torch.nan_to_num(self.response_batch(mat, num_images=num_images))
        A = self.image2array_masked(images)  # [B, n, N_d]
        cd = torch.as_tensor(self.image2array_masked(self.Data.C_D), device=A.device)
        if self._psf_error_map and sum(num_images or self._num_point_images) > 0:  # C_D + model error
            err = self._split_plan(num_images=num_images).psf_error_map(mat)
            w = 1.0 / (cd[None, :] + self.image2array_masked(err))
        else:
            w = (1.0 / cd)[None, :].expand(A.shape[0], -1)
        M = torch.einsum("bin,bn,bjn->bij", A, w, A)
        stable = torch.linalg.cond(M) < 5 / np.finfo(float).eps
        Minv = torch.linalg.pinv(M)
        return torch.where(stable[:, None, None], torch.nan_to_num(Minv, nan=0.0, posinf=0.0, neginf=0.0),
                           torch.zeros_like(M))

    def likelihood_data_given_model(self, kwargs_lens=None, kwargs_source=None, kwargs_lens_light=None, kwargs_ps=None,
                                    kwargs_extinction=None, kwargs_special=None, source_marg=False, linear_prior=None,
                                    check_positive_flux=False):
        """-> (logL, param): linear inversion + chi^2 (+ marginalisation constant) (image_linear_solve.py:228-339)."""
        if check_positive_flux:
            raise ValueError("check positive flux is not supported in jaxtronomy")
        if linear_prior is not None:
            raise ValueError("linear_prior (d_prior marginalisation, de_lens.py:52-74) is outside the scope of "
                             "jaxtronomy_b200")
        counts = self._ps_counts(kwargs_ps)
        mat, B, batched = self._pack(kwargs_lens, kwargs_source, kwargs_lens_light, kwargs_ps, counts)
        want_torch = self._want_torch(kwargs_lens, kwargs_source, kwargs_lens_light)
        logl, amps, _ = self.logL_linear_batch(mat, source_marg=source_marg, num_images=counts)
        return self._finish(logl, batched, want_torch), self._finish(amps, batched, want_torch)

    def update_linear_kwargs(self, param, kwargs_lens, kwargs_source, kwargs_lens_light, kwargs_ps=None):
        """Write the solved amplitudes into the kwargs lists (image_linear_solve.py:457-481)."""
        i = 0
        kwargs_source, i = self.SourceModel.update_linear(param, i, kwargs_source)
        kwargs_lens_light, i = self.LensLightModel.update_linear(param, i, kwargs_lens_light)
        if kwargs_ps and self._n_ps_models:
            kwargs_ps, i = self.PointSource.update_linear(param, i, kwargs_ps, kwargs_lens)
        return kwargs_lens, kwargs_source, kwargs_lens_light, kwargs_ps

    def linear_param_from_kwargs(self, kwargs_source, kwargs_lens_light, kwargs_ps=None):
        out = self.SourceModel.linear_param_from_kwargs(kwargs_source) + \
            self.LensLightModel.linear_param_from_kwargs(kwargs_lens_light)
        if kwargs_ps and self._n_ps_models:
            out = out + self.PointSource.linear_param_from_kwargs(kwargs_ps)
        return out
