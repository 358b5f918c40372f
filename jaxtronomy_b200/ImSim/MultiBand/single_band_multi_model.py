"""SingleBandMultiModel: one band of a multi-band model, jaxtronomy/ImSim/MultiBand/single_band_multi_model.py:
the band uses the sub-lists of lens / source / lens-light models named by ``index_*_model_list[band_index]`` in
kwargs_model (class_creator.create_class_instances with band_index) and selects the matching kwargs
(select_kwargs, :639-694)."""

from jaxtronomy_b200.Data.imaging_data import ImageData
from jaxtronomy_b200.Data.psf import PSF
from jaxtronomy_b200.ImSim.image_linear_solve import ImageLinearFit
from jaxtronomy_b200.ImSim.image_model import ImageModel
from jaxtronomy_b200.LensModel.lens_model import LensModel
from jaxtronomy_b200.LightModel.light_model import LightModel

__all__ = ["SingleBandMultiModel"]


def _select(lst, index):
    if lst is None or index is None:
        return lst
    return [lst[k] for k in index]


class SingleBandMultiModel(object):
    """Holds an ImageLinearFit (linear_solver=True) or ImageModel for one band and forwards to it with the
    band's kwargs subset.  (The reference inherits from the image class; composition keeps the plan caches of
    the two variants apart.)"""

    def __init__(self, multi_band_list, kwargs_model, likelihood_mask_list=None, band_index=0, kwargs_pixelbased=None,
                 linear_solver=True):
        self.type = "single-band-multi-model"
        self._band_index = band_index
        n_bands = len(multi_band_list)
        if likelihood_mask_list is None:
            likelihood_mask_list = [None for _ in range(n_bands)]
        none_list = [None for _ in range(n_bands)]
        self._index_lens_model = kwargs_model.get("index_lens_model_list", none_list)[band_index]
        self._index_source = kwargs_model.get("index_source_light_model_list", none_list)[band_index]
        self._index_lens_light = kwargs_model.get("index_lens_light_model_list", none_list)[band_index]
        if kwargs_model.get("point_source_model_list"):
            raise ValueError("point sources are outside the scope of jaxtronomy_b200")
        lens_list = kwargs_model.get("lens_model_list", [])
        lens_pk = kwargs_model.get("lens_profile_kwargs_list")
        lens_z = kwargs_model.get("lens_redshift_list")
        lens_model = LensModel(
            _select(lens_list, self._index_lens_model), profile_kwargs_list=_select(lens_pk, self._index_lens_model),
            z_source=kwargs_model.get("z_source"), lens_redshift_list=_select(lens_z, self._index_lens_model),
            multi_plane=kwargs_model.get("multi_plane", False), cosmo=kwargs_model.get("cosmo"),
            z_source_convention=kwargs_model.get("z_source_convention"), z_lens=kwargs_model.get("z_lens"))
        source_model = LightModel(_select(kwargs_model.get("source_light_model_list", []), self._index_source))
        lens_light_model = LightModel(_select(kwargs_model.get("lens_light_model_list", []), self._index_lens_light))
        kwargs_data, kwargs_psf, kwargs_numerics = multi_band_list[band_index][:3]
        self.linear_solver = linear_solver
        cls = ImageLinearFit if linear_solver else ImageModel
        self.image_model = cls(ImageData(**kwargs_data), PSF(**kwargs_psf), lens_model, source_model, lens_light_model,
                               kwargs_numerics=kwargs_numerics, likelihood_mask=likelihood_mask_list[band_index],
                               kwargs_pixelbased=kwargs_pixelbased)
        self.num_data_evaluate = self.image_model.num_data_evaluate

    def select_kwargs(self, kwargs_lens=None, kwargs_source=None, kwargs_lens_light=None, kwargs_ps=None,
                      kwargs_extinction=None, kwargs_special=None):
        return (_select(kwargs_lens, self._index_lens_model), _select(kwargs_source, self._index_source),
                _select(kwargs_lens_light, self._index_lens_light), kwargs_ps, None)

    @property
    def index_lists(self):
        return {"lens": self._index_lens_model, "source": self._index_source, "lens_light": self._index_lens_light}

    def image(self, kwargs_lens=None, kwargs_source=None, kwargs_lens_light=None, kwargs_ps=None, **kwargs):
        kl, ks, kll, _, _ = self.select_kwargs(kwargs_lens, kwargs_source, kwargs_lens_light)
        return self.image_model.image(kl, ks, kll, **kwargs)

    def image_linear_solve(self, kwargs_lens=None, kwargs_source=None, kwargs_lens_light=None, kwargs_ps=None,
                           kwargs_extinction=None, kwargs_special=None, inv_bool=False):
        kl, ks, kll, _, _ = self.select_kwargs(kwargs_lens, kwargs_source, kwargs_lens_light)
        return self.image_model.image_linear_solve(kl, ks, kll, inv_bool=inv_bool)

    def likelihood_data_given_model(self, kwargs_lens=None, kwargs_source=None, kwargs_lens_light=None, kwargs_ps=None,
                                    kwargs_extinction=None, kwargs_special=None, source_marg=False, linear_prior=None,
                                    check_positive_flux=False):
        kl, ks, kll, _, _ = self.select_kwargs(kwargs_lens, kwargs_source, kwargs_lens_light)
        out = self.image_model.likelihood_data_given_model(kl, ks, kll, source_marg=source_marg,
                                                           linear_prior=linear_prior,
                                                           check_positive_flux=check_positive_flux)
        return out if self.linear_solver else (out, None)

    def sampler_plans(self, param_class, n_cols, linear_solver):
        return self.image_model.sampler_plans(param_class, n_cols, linear_solver, index_lists=self.index_lists)
