"""LikelihoodModule: flat sampler vector -> log-likelihood, jaxtronomy/Sampling/likelihood.py:293-405 and
Sampling/Likelihoods/image_likelihood.py:56-96 for the imaging likelihood, plus Gaussian priors on the sampled
parameters (position / time-delay / flux-ratio / kinematic likelihoods and KDE priors: out of scope).

The reference evaluates ``jit(vmap(logL))`` over the particle axis (Sampling/sampler.py:241-246).  Here the whole
``positions[B, n]`` matrix goes to the GPU as is: the plan's column map *is* ``Param.args2kwargs``, the bound check
is a reduction over the matrix, and one lsim_logl call returns ``logL[B]``.
"""

import numpy as np

from jaxtronomy_b200 import _lib

__all__ = ["LikelihoodModule"]


class LikelihoodModule(object):
    def __init__(self, image_model, param_class, check_bounds=True, source_marg=False, linear_solver=True,
                 prior_lens=None, prior_source=None, prior_lens_light=None, image_position_uncertainty=0.005,
                 astrometric_likelihood=False, image_position_likelihood=False, ra_image_list=None, dec_image_list=None,
                 source_position_likelihood=None, source_position_tolerance=None, source_position_sigma=0.001,
                 force_no_add_image=False, restrict_image_number=False, max_num_images=None, **unsupported):
        """
        :param image_model: ImageModel / ImageLinearFit / MultiLinear instance (the imaging likelihood)
        :param param_class: jaxtronomy_b200.Sampling.parameters.Param
        :param prior_lens, prior_source, prior_lens_light: Gaussian priors ``[[index, param_name, mean, sigma], ...]``
            (the reference forwards them to lenstronomy's PriorLikelihood, Sampling/likelihood.py:155-176, and adds
            its value to logL, :345-349): logL -= (value - mean)^2 / (2 sigma^2).  Evaluated on the columns of the
            `positions` matrix on the device; a prior on a fixed parameter is a constant.  lenstronomy is not under
            /root/reference: the formula is restated (parity unpinned).  KDE / log-normal priors, the
            time-delay, flux-ratio, kinematic and tracer likelihoods are outside the scope: passing
            them raises.
        :param image_position_uncertainty ... max_num_images: the reference's ``kwargs_position`` (likelihood.py:196-214),
            forwarded to PositionLikelihood.  As in the reference the source-position penalty is ON by default when the
            model has a LENSED_POSITION point source (position_likelihood.py:62-66).  ``astrometric_likelihood`` needs
            ``kwargs_special`` (delta_x_image), which Param does not sample: raises.
        """
        for k, v in unsupported.items():
            if v:
                raise ValueError("LikelihoodModule option %s is outside the scope of jaxtronomy_b200" % k)
        self.imSim = image_model
        self.param = param_class
        self._check_bounds = check_bounds
        self._source_marg = source_marg
        self._linear_solver = linear_solver
        self._lower_limit, self._upper_limit = self.param.param_limits()
        self._n = self.param.num_param()[0]
        self._bounds_dev = None
        self._plans = self.imSim.sampler_plans(self.param, self._n, linear_solver)
        # Gaussian priors: (column, mean, sigma) of sampled parameters + the constant of priors on fixed ones
        self._prior_cols, self._prior_const = [], 0.0
        for comp, priors in (("lens", prior_lens), ("source", prior_source), ("lens_light", prior_lens_light)):
            for index, name, mean, sigma in (priors or []):
                key = (comp, int(index), name)
                if key in self.param._index:
                    self._prior_cols.append((self.param._index[key], float(mean), float(sigma)))
                elif name in self.param._fixed[comp][int(index)]:
                    self._prior_const -= (float(self.param._fixed[comp][int(index)][name]) - mean) ** 2 / sigma**2 / 2
                else:
                    raise ValueError("prior on unknown parameter %s of %s model %d" % (name, comp, index))
        self._prior_dev = None
        # position likelihood (likelihood.py:286-288, 366-368)
        if astrometric_likelihood:
            raise ValueError("astrometric_likelihood needs kwargs_special, which is outside the scope of jaxtronomy_b200's "
                             "Param; use PositionLikelihood.astrometric_likelihood on kwargs")
        self._position_likelihood, self._pos_models, self._pos_plan, self._pos_dev = None, [], None, None
        ps = getattr(image_model, "PointSource", None)
        if ps is not None and len(ps.point_source_type_list) > 0:
            from jaxtronomy_b200.Sampling.Likelihoods.position_likelihood import PositionLikelihood

            pl = PositionLikelihood(ps, image_position_uncertainty=image_position_uncertainty,
                                    image_position_likelihood=image_position_likelihood, ra_image_list=ra_image_list,
                                    dec_image_list=dec_image_list, source_position_likelihood=source_position_likelihood,
                                    source_position_tolerance=source_position_tolerance,
                                    source_position_sigma=source_position_sigma, force_no_add_image=force_no_add_image,
                                    restrict_image_number=restrict_image_number, max_num_images=max_num_images)
            self._position_likelihood = pl
            fn = self.param.column_source_fn()
            for k, n in enumerate(fn.num_point_images):
                src = [[fn("point_source", k, name, element=j) for j in range(n)] for name in ("ra_image", "dec_image")]
                self._pos_models.append((ps.point_source_type_list[k], n, src))
            if pl._image_position_likelihood and (len(pl._ra_image_list) < len(self._pos_models)
                                                  or len(pl._dec_image_list) < len(self._pos_models)):
                raise ValueError("image_position_likelihood needs ra_image_list / dec_image_list with one entry per "
                                 "point-source model")
            if pl._source_position_likelihood and "LENSED_POSITION" in ps.point_source_type_list:
                # the reference's class_creator always hands the lens model to PointSource; a PointSource built without
                # one falls back to the image model's
                lm = ps._lens_model if ps._lens_model is not None else getattr(image_model, "LensModel", None)
                if lm is None:
                    raise ValueError("the source-position likelihood needs the lens model: build PointSource with "
                                     "lens_model=..., or pass source_position_likelihood=False")
                if getattr(lm, "multi_plane", False):
                    # the penalty needs the lensing Jacobian at every image; lsim_source_position_logl has it for
                    # single-plane models only (the reference differences MultiPlane.alpha, multi_plane.py:537-589)
                    raise ValueError("the source-position likelihood of a multi-plane lens model is not supported: "
                                     "pass source_position_likelihood=False")
                self._pos_plan = _lib.Plan(self._n, lens_descs=lm.lens_descs(None, source_fn=fn), **lm.plan_kwargs())
        elif source_position_likelihood or image_position_likelihood:
            raise ValueError("position likelihoods need an image model with a point-source class")

    def _image_positions(self, positions):
        """Per point-source model: (ra [B, n], dec [B, n]) device tensors, columns of `positions` or fixed constants."""
        import torch

        if self._pos_dev is None or self._pos_dev[0][0][0].device != positions.device:
            dev = []
            for _, n, src in self._pos_models:
                per = []
                for lst in src:
                    idx = torch.as_tensor([v if kind == "col" else 0 for kind, v in lst], dtype=torch.long,
                                          device=positions.device)
                    is_col = torch.as_tensor([kind == "col" for kind, _ in lst], device=positions.device)
                    cst = torch.as_tensor([0.0 if kind == "col" else v for kind, v in lst], dtype=positions.dtype,
                                          device=positions.device)
                    per.append((idx, is_col, cst))
                dev.append(per)
            self._pos_dev = dev
        out = []
        for per in self._pos_dev:
            out.append(tuple(torch.where(is_col, positions[:, idx], cst).contiguous() for idx, is_col, cst in per))
        return out

    def _position_logl(self, positions):
        """PositionLikelihood.logL over the rows of `positions` (device tensor) -> [B] or None."""
        pl = self._position_likelihood
        if pl is None or not self._pos_models or not (self._pos_plan is not None or pl._image_position_likelihood):
            return None
        total = None
        coords = self._image_positions(positions)
        for k, (kind, n, _) in enumerate(self._pos_models):
            ra, dec = coords[k]
            if self._pos_plan is not None and kind == "LENSED_POSITION" and n > 0:
                l = self._pos_plan.source_position_logl(positions, ra, dec, pl._source_position_sigma,
                                                        pl._bound_source_position_tolerance)
                total = l if total is None else total + l
            if pl._image_position_likelihood:  # position_likelihood.py:214-247 on the columns themselves
                import torch

                m = min(len(pl._ra_image_list[k]), n)
                ra0 = torch.as_tensor(pl._ra_image_list[k][:m], dtype=positions.dtype, device=positions.device)
                dec0 = torch.as_tensor(pl._dec_image_list[k][:m], dtype=positions.dtype, device=positions.device)
                l = -(((ra[:, :m] - ra0) ** 2 + (dec[:, :m] - dec0) ** 2) / pl._image_position_sigma**2 / 2).sum(dim=1)
                total = l if total is None else total + l
        return total

    def _prior(self, positions, want_grad=False):
        """Gaussian prior term of every row of `positions` (device tensor) [+ its gradient]."""
        import torch

        if not self._prior_cols:
            z = torch.full((positions.shape[0],), self._prior_const, dtype=positions.dtype, device=positions.device)
            return (z, None) if want_grad else z
        if self._prior_dev is None or self._prior_dev[0].device != positions.device:
            cols = torch.as_tensor([c for c, _, _ in self._prior_cols], device=positions.device)
            mean = torch.as_tensor([m for _, m, _ in self._prior_cols], dtype=positions.dtype, device=positions.device)
            isig2 = torch.as_tensor([1.0 / s**2 for _, _, s in self._prior_cols], dtype=positions.dtype,
                                    device=positions.device)
            self._prior_dev = (cols, mean, isig2)
        cols, mean, isig2 = self._prior_dev
        d = positions[:, cols] - mean
        val = -(d * d * isig2).sum(dim=1) / 2 + self._prior_const
        if not want_grad:
            return val
        grad = torch.zeros_like(positions)
        grad.index_add_(1, cols, -d * isig2)
        return val, grad

    def __call__(self, a):
        return self.logL(a)

    @staticmethod
    def check_bounds(args, lowerLimit, upperLimit, verbose=False):
        """likelihood.py:374-405 -> (penalty, bound_hit)"""
        args = np.atleast_1d(np.asarray(args))
        hit = bool(np.any((args < np.atleast_1d(lowerLimit)) | (args > np.atleast_1d(upperLimit))))
        return (10.0**5 if hit else 0.0), hit

    def logL(self, args, verbose=False):
        """One parameter vector -> float (likelihood.py:293-319)."""
        return float(self.logL_batch(np.asarray(args, dtype=float)[None, :])[0])

    def logL_batch_device(self, positions, out=None):
        """positions: [B, n] CUDA tensor -> logL [B] CUDA tensor on the current stream."""
        import torch

        total = None
        for plan in self._plans:
            l, _, _ = plan.logl(positions, linear_solve=self._linear_solver, source_marg=self._source_marg)
            l = torch.where(torch.isnan(l), torch.full_like(l, -(10.0**15)), l)  # image_likelihood.py:95
            total = l if total is None else total + l
        if self._prior_cols or self._prior_const:
            total = total + self._prior(positions)
        pos_l = self._position_logl(positions)
        if pos_l is not None:
            total = total + pos_l
        if self._check_bounds:
            if self._bounds_dev is None or self._bounds_dev[0].device != positions.device:
                self._bounds_dev = (torch.as_tensor(self._lower_limit, device=positions.device),
                                    torch.as_tensor(self._upper_limit, device=positions.device))
            lo, hi = self._bounds_dev
            hit = ((positions < lo) | (positions > hi)).any(dim=1)
            total = torch.where(hit, torch.full_like(total, -(10.0**18)), total)
        total = torch.where(torch.isnan(total), torch.full_like(total, 1e-18), total)  # likelihood.py:370
        if out is not None:
            out.copy_(total)
            return out
        return total

    def logL_and_grad_batch_device(self, positions):
        """positions: [B, n] CUDA tensor -> (logL [B], d logL / d positions [B, n]) on the current stream: the batched
        counterpart of jax.value_and_grad(LikelihoodModule.logL) (Sampling/Samplers/optax.py:39).  The penalties are
        constants (jnp.where branches without a parameter dependence): their gradient is 0."""
        import torch

        if self._source_marg:
            raise ValueError("the gradient does not cover the source marginalisation constant")
        pl = self._position_likelihood
        if pl is not None and self._pos_plan is not None:
            # the penalty is built on the lensing Jacobian at the images: its gradient needs third derivatives of the
            # potential, which no kernel here carries (the imaging likelihood's gradient does cover the positions)
            raise ValueError("the gradient does not cover the source-position penalty of the position likelihood: "
                             "pass source_position_likelihood=False")
        total, gtotal = None, None
        for plan in self._plans:
            l, g, _ = plan.logl_grad(positions, linear_solve=self._linear_solver)
            bad = torch.isnan(l)
            l = torch.where(bad, torch.full_like(l, -(10.0**15)), l)
            g = torch.where(bad[:, None], torch.zeros_like(g), g)
            total = l if total is None else total + l
            gtotal = g if gtotal is None else gtotal + g
        if self._prior_cols or self._prior_const:
            pv, pg = self._prior(positions, want_grad=True)
            total = total + pv
            if pg is not None:
                gtotal = gtotal + pg
        if pl is not None and pl._image_position_likelihood and self._pos_models:
            # image-position term (position_likelihood.py:214-247): a quadratic form of the sampled columns themselves
            for k, (kind, n, src) in enumerate(self._pos_models):
                m = min(len(pl._ra_image_list[k]), n)
                for lst, ref in ((src[0], pl._ra_image_list[k]), (src[1], pl._dec_image_list[k])):
                    for e in range(m):
                        kind_e, v = lst[e]
                        if kind_e == "col":
                            d = positions[:, v] - float(ref[e])
                            total = total - d * d / pl._image_position_sigma**2 / 2
                            gtotal[:, v] -= d / pl._image_position_sigma**2
                        else:
                            total = total - (v - float(ref[e])) ** 2 / pl._image_position_sigma**2 / 2
        if self._check_bounds:
            if self._bounds_dev is None or self._bounds_dev[0].device != positions.device:
                self._bounds_dev = (torch.as_tensor(self._lower_limit, device=positions.device),
                                    torch.as_tensor(self._upper_limit, device=positions.device))
            lo, hi = self._bounds_dev
            hit = ((positions < lo) | (positions > hi)).any(dim=1)
            total = torch.where(hit, torch.full_like(total, -(10.0**18)), total)
            gtotal = torch.where(hit[:, None], torch.zeros_like(gtotal), gtotal)
        return total, gtotal

    def logL_and_grad_batch(self, positions):
        """positions: [B, n] NumPy array -> (logL [B], grad [B, n]) NumPy arrays."""
        import torch

        pos = torch.as_tensor(np.ascontiguousarray(positions, dtype=np.float64)).to("cuda")
        if pos.ndim != 2 or pos.shape[1] != self._n:
            raise _lib.LsimError("positions must be [B, %d]" % self._n)
        l, g = self.logL_and_grad_batch_device(pos)
        return l.cpu().numpy(), g.cpu().numpy()

    def value_and_grad(self, args):
        """One parameter vector -> (logL, d logL / d args)."""
        l, g = self.logL_and_grad_batch(np.asarray(args, dtype=float)[None, :])
        return float(l[0]), g[0]

    def logL_batch(self, positions):
        """positions: [B, n] NumPy array -> logL NumPy array [B] (the `new_logL_func` of prepare_logL_func)."""
        import torch

        pos = torch.as_tensor(np.ascontiguousarray(positions, dtype=np.float64)).to("cuda")
        if pos.ndim != 2 or pos.shape[1] != self._n:
            raise _lib.LsimError("positions must be [B, %d]" % self._n)
        return self.logL_batch_device(pos).cpu().numpy()

    @property
    def num_data(self):
        return self.imSim.num_data_evaluate
