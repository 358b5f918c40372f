"""MultiLinear: joint likelihood of several bands that share the non-linear parameters,
jaxtronomy/ImSim/MultiBand/multi_linear.py:20-161 and multi_data_base.py: logL = sum over bands of the
per-band ImageLinearFit likelihood (each band solves its own linear amplitudes)."""

from jaxtronomy_b200.ImSim.MultiBand.single_band_multi_model import SingleBandMultiModel

__all__ = ["MultiLinear"]


class MultiLinear(object):
    def __init__(self, multi_band_list, kwargs_model, likelihood_mask_list=None, compute_bool=None,
                 kwargs_pixelbased=None, linear_solver=True):
        self.type = "multi-linear"
        if linear_solver is False and len(multi_band_list) > 1:
            raise ValueError('Multi-linear mode with more than one band does not support "linear_solver" = False.')
        self._imageModel_list = [
            SingleBandMultiModel(multi_band_list, kwargs_model, likelihood_mask_list=likelihood_mask_list,
                                 band_index=i, kwargs_pixelbased=kwargs_pixelbased, linear_solver=linear_solver)
            for i in range(len(multi_band_list))]
        self._num_bands = len(self._imageModel_list)
        if compute_bool is None:
            compute_bool = [True] * self._num_bands
        elif len(compute_bool) != self._num_bands:
            raise ValueError("compute_bool statement has not the same range as number of bands available!")
        self._compute_bool = [bool(b) for b in compute_bool]
        self._linear_solver = linear_solver

    @property
    def num_bands(self):
        return self._num_bands

    @property
    def num_data_evaluate(self):
        return sum(m.num_data_evaluate for m, c in zip(self._imageModel_list, self._compute_bool) if c)

    def image_linear_solve(self, kwargs_lens=None, kwargs_source=None, kwargs_lens_light=None, kwargs_ps=None,
                           kwargs_extinction=None, kwargs_special=None, inv_bool=False):
        """-> lists (one entry per band) of wls model, error map, cov_param, param (multi_linear.py:62-111)."""
        wls, err, cov, par = [], [], [], []
        for m, c in zip(self._imageModel_list, self._compute_bool):
            if c:
                a, b, cc, d = m.image_linear_solve(kwargs_lens, kwargs_source, kwargs_lens_light, inv_bool=inv_bool)
            else:
                a = b = cc = d = None
            wls.append(a)
            err.append(b)
            cov.append(cc)
            par.append(d)
        return wls, err, cov, par

    def likelihood_data_given_model(self, kwargs_lens=None, kwargs_source=None, kwargs_lens_light=None, kwargs_ps=None,
                                    kwargs_extinction=None, kwargs_special=None, source_marg=False, linear_prior=None,
                                    check_positive_flux=False):
        """-> (sum of the band log-likelihoods, [linear parameters per band]) (multi_linear.py:112-161)."""
        logL, param_list = 0, []
        for m, c in zip(self._imageModel_list, self._compute_bool):
            if c:
                logL_i, param_i = m.likelihood_data_given_model(kwargs_lens, kwargs_source, kwargs_lens_light,
                                                                source_marg=source_marg)
                logL = logL + logL_i
                param_list.append(param_i)
            else:
                param_list.append(None)
        return logL, param_list

    def sampler_plans(self, param_class, n_cols, linear_solver):
        plans = []
        for m, c in zip(self._imageModel_list, self._compute_bool):
            if c:
                plans += m.sampler_plans(param_class, n_cols, linear_solver)
        return plans
