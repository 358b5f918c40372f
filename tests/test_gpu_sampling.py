"""GPU: the sampler boundary (LikelihoodModule / prepare_logL_func) and multi-band likelihood, through the C ABI."""

import numpy as np
import pytest

from tests.golden_io import load

pytestmark = pytest.mark.gpu
IMSIM = load("imsim.npz")


def _setup(case="headline_epl_shear_pixel"):
    from jaxtronomy_b200.Sampling.likelihood import LikelihoodModule
    from jaxtronomy_b200.Sampling.parameters import Param
    from tests.test_gpu_parity import build_image_model

    c = IMSIM[case]
    lin = build_image_model(c, linear=True)
    km = dict(lens_model_list=c.meta["lens_model_list"], source_light_model_list=c.meta["source_model_list"],
              lens_light_model_list=c.meta["lens_light_model_list"])
    fixed_lens = [{} for _ in c.meta["lens_model_list"]]
    fixed_lens[-1] = {"ra_0": 0.0, "dec_0": 0.0}
    param = Param(km, kwargs_fixed_lens=fixed_lens, kwargs_lower_lens=[{"theta_E": 0.6}, {}],
                  kwargs_upper_lens=[{"theta_E": 1.6}, {}])
    return c, lin, param, LikelihoodModule(lin, param)


def test_likelihood_module_batch_matches_kwargs_path():
    c, lin, param, like = _setup()
    kl, ks, kll = c.meta["kwargs_lens"], c.meta["kwargs_source"], c.meta["kwargs_lens_light"]
    x0 = param.kwargs2args(kl, ks, kll)
    rng = np.random.default_rng(4)
    B = 37
    pos = x0[None, :] + rng.normal(0, 0.01, (B, x0.size))
    pos[5, 0] = 5.0    # theta_E above its upper bound
    pos[9, 1] = 0.2    # gamma below the default lower bound 1.5
    logl = like.logL_batch(pos)
    assert logl.shape == (B,)
    assert logl[5] == -1e18 and logl[9] == -1e18          # likelihood.py:312-315
    for b in (0, 6, 36):
        kw = param.args2kwargs(pos[b])
        want, _ = lin.likelihood_data_given_model(kw["kwargs_lens"], kw["kwargs_source"], kw["kwargs_lens_light"])
        assert abs(logl[b] - want) <= 1e-12 * abs(want), (b, logl[b], want)
    assert like.logL(pos[0]) == logl[0]
    # against the reference fixture at the fixture's own parameters
    assert abs(like.logL(x0) - float(c.out["logL_linear"])) <= 1e-10 * abs(float(c.out["logL_linear"]))


def test_prepare_logL_func_and_pso_improve_the_fit():
    from jaxtronomy_b200.Sampling.sampler import ParticleSwarmOptimizer, prepare_logL_func

    c, lin, param, like = _setup()
    f = prepare_logL_func("gpu", like)
    x0 = param.kwargs2args(c.meta["kwargs_lens"], c.meta["kwargs_source"], c.meta["kwargs_lens_light"])
    lo, hi = x0 - 0.05, x0 + 0.05
    pso = ParticleSwarmOptimizer(f, lo, hi, particle_count=64, seed=0)
    best, hist, _ = pso.optimize(max_iter=15)
    assert np.all(np.diff(hist) >= 0)
    assert hist[-1] > hist[0]
    assert hist[-1] >= like.logL(x0) - 50.0


def test_multi_band_likelihood_is_the_sum_of_bands():
    """Two bands with different data / PSF / supersampling sharing the lens and source, each with its own lens
    light (index lists), against the oracle evaluated per band (multi_linear.py:112-161)."""
    from jaxtronomy_b200.ImSim.MultiBand.multi_linear import MultiLinear
    from jaxtronomy_b200.Sampling.likelihood import LikelihoodModule
    from jaxtronomy_b200.Sampling.parameters import Param
    from jaxtronomy_b200.Util import simulation_util
    from oracle import imsim as oimsim
    from oracle.lenstronomy_port import PSF as OPSF

    rng = np.random.default_rng(21)
    kl = [dict(theta_E=1.1, gamma=2.05, e1=0.2, e2=-0.1, center_x=0.02, center_y=-0.01),
          dict(gamma1=0.03, gamma2=-0.02, ra_0=0.0, dec_0=0.0)]
    ks = [dict(amp=12.0, R_sersic=0.25, n_sersic=2.0, e1=0.1, e2=0.1, center_x=0.05, center_y=0.1)]
    kll = [dict(amp=20.0, R_sersic=0.5, n_sersic=3.0, e1=-0.1, e2=0.05, center_x=0.02, center_y=-0.01),
           dict(amp=9.0, R_sersic=0.4, n_sersic=2.5, e1=0.0, e2=0.1, center_x=0.02, center_y=-0.01)]
    km = dict(lens_model_list=["EPL", "SHEAR"], source_light_model_list=["SERSIC_ELLIPSE"],
              lens_light_model_list=["SERSIC_ELLIPSE", "SERSIC_ELLIPSE"],
              index_lens_light_model_list=[[0], [1]])
    a = np.arange(7) - 3.0
    kern = np.exp(-(a[:, None] ** 2 + a[None, :] ** 2) / 3.0)
    bands, oracles = [], []
    for band, (npx, dpx, ss, psf_kw) in enumerate([
            (30, 0.1, 2, dict(psf_type="PIXEL", kernel_point_source=kern)),
            (24, 0.12, 1, dict(psf_type="GAUSSIAN", fwhm=0.25, truncation=4))]):
        kd = simulation_util.data_configure_simple(npx, dpx, 100.0, 0.05)
        kd["image_data"] = rng.uniform(0, 1, (npx, npx))
        bands.append([kd, psf_kw, dict(supersampling_factor=ss)])
        od = oimsim.ImageData(**kd)
        oracles.append(oimsim.ImageModel(od, OPSF(**psf_kw), km["lens_model_list"], km["source_light_model_list"],
                                         [km["lens_light_model_list"][band]], dict(supersampling_factor=ss)))
    ml = MultiLinear(bands, km)
    logL, params = ml.likelihood_data_given_model(kl, ks, kll)
    want, want_p = 0.0, []
    for band, om in enumerate(oracles):
        l, p = om.likelihood_data_given_model_linear(kl, ks, [kll[band]])
        want += l
        want_p.append(p)
    assert abs(logL - want) <= 1e-10 * abs(want), (logL, want)
    for p, wp in zip(params, want_p):
        np.testing.assert_allclose(p, wp, rtol=1e-9)
    # sampler boundary over both bands: one positions matrix, plans per band, sum on the device
    param = Param(km)
    like = LikelihoodModule(ml, param)
    x0 = param.kwargs2args(kl, ks, kll)
    got = like.logL_batch(np.stack([x0, x0]))
    assert abs(got[0] - want) <= 1e-10 * abs(want) and got[0] == got[1]
    # the gradient of the multi-band likelihood is the sum of the bands' gradients (pixel PSF band + Gaussian PSF band)
    val, grad = like.value_and_grad(x0)
    assert abs(val - want) <= 1e-10 * abs(want) and np.all(np.isfinite(grad))
    scale = np.abs(grad).max()
    for j in range(len(x0)):
        best = np.inf
        for h in (1e-4, 1e-5, 1e-6):
            e = np.zeros_like(x0)
            e[j] = h
            fd = (8 * (like.logL(x0 + e) - like.logL(x0 - e)) - (like.logL(x0 + 2 * e) - like.logL(x0 - 2 * e))) / (12 * h)
            best = min(best, abs(grad[j] - fd))
        assert best <= 1e-6 * scale, (j, grad[j], best)


def test_pso_step_is_bit_exact_against_the_oracle():
    """lsim_pso_step vs oracle/pso.py on the same state and the same uniform numbers: array_equal, including ties
    (first particle wins), a nan fitness (never wins) and particles that do not improve."""
    import torch

    from jaxtronomy_b200 import _lib
    from oracle import pso as opso

    rng = np.random.default_rng(11)
    for B, P in ((1, 3), (37, 11), (5000, 17)):
        st = dict(position=rng.normal(0, 1, (B, P)), velocity=rng.normal(0, 0.1, (B, P)),
                  pbest_pos=rng.normal(0, 1, (B, P)), pbest_fit=rng.normal(-100, 10, B),
                  gbest_pos=rng.normal(0, 1, P), gbest_fit=np.array([-95.0]))
        fitness = rng.normal(-100, 10, B)
        if B > 4:
            fitness[3] = fitness[1] = fitness.max() + 1.0   # tie for the best: index 1 must win
            fitness[2] = np.nan
        rnd = rng.uniform(0, 1, (3, B, P))
        dev = {k: torch.as_tensor(v.copy(), device="cuda") for k, v in st.items()}
        for _ in range(2):  # two consecutive iterations on the evolving state
            opso.pso_step(st["position"], st["velocity"], st["pbest_pos"], st["pbest_fit"], st["gbest_pos"],
                          st["gbest_fit"], fitness, rnd)
            _lib.pso_step(dev["position"], dev["velocity"], dev["pbest_pos"], dev["pbest_fit"], dev["gbest_pos"],
                          dev["gbest_fit"], torch.as_tensor(fitness, device="cuda"), torch.as_tensor(rnd, device="cuda"))
            for k in st:
                np.testing.assert_array_equal(dev[k].cpu().numpy(), st[k], err_msg="%s B=%d" % (k, B))
            fitness = fitness - 1.0  # second pass: nobody improves, only the move changes the state


def test_device_pso_improves_the_fit_without_host_round_trips():
    from jaxtronomy_b200.Sampling.sampler import DeviceParticleSwarmOptimizer

    c, lin, param, like = _setup()
    x0 = param.kwargs2args(c.meta["kwargs_lens"], c.meta["kwargs_source"], c.meta["kwargs_lens_light"])
    pso = DeviceParticleSwarmOptimizer(like, x0 - 0.05, x0 + 0.05, particle_count=64, seed=0)
    best, hist = pso.optimize(max_iter=15)
    assert best.shape == x0.shape
    assert np.all(np.diff(hist) >= 0) and hist[-1] > hist[0]
    assert hist[-1] >= like.logL(x0) - 50.0
    assert abs(like.logL(best) - hist[-1]) <= 1e-9 * abs(hist[-1])


def test_likelihood_module_with_point_sources():
    """Sampler boundary with point sources: the image positions are columns of the `positions` matrix, the point
    amplitudes are solved linearly (one per image) together with the extended light."""
    from jaxtronomy_b200.Sampling.likelihood import LikelihoodModule
    from jaxtronomy_b200.Sampling.parameters import Param
    from tests.test_gpu_parity import POINTS, _ps_model

    c = POINTS["image_quad_and_star"]
    lin = _ps_model(c, linear=True)
    km = dict(lens_model_list=c.meta["lens_model_list"], source_light_model_list=c.meta["source_model_list"],
              lens_light_model_list=c.meta["lens_light_model_list"],
              point_source_model_list=c.meta["point_source_type_list"])
    kps = [{k: np.asarray(v, dtype=float) for k, v in kw.items()} for kw in c.meta["kwargs_ps"]]
    param = Param(km, kwargs_fixed_lens=[{}, {"ra_0": 0.0, "dec_0": 0.0}], num_point_source_list=[4, 1],
                  kwargs_fixed_ps=[{}, {"ra_image": kps[1]["ra_image"], "dec_image": kps[1]["dec_image"]}])
    # imaging likelihood alone (the reference turns the source-position penalty on by default for LENSED_POSITION
    # models, position_likelihood.py:62-66; that term is covered by tests/test_gpu_zposition.py)
    like = LikelihoodModule(lin, param, source_position_likelihood=False)
    kl, ks, kll = c.meta["kwargs_lens"], c.meta["kwargs_source"], c.meta["kwargs_lens_light"]
    x0 = param.kwargs2args(kl, ks, kll, kwargs_ps=kps)
    assert abs(like.logL(x0) - float(c.out["logL_linear"])) <= 1e-9 * abs(float(c.out["logL_linear"]))
    rng = np.random.default_rng(3)
    pos = x0[None, :] + rng.normal(0, 0.003, (9, x0.size))
    logl = like.logL_batch(pos)
    for b in (0, 8):
        kw = param.args2kwargs(pos[b])
        want, _ = lin.likelihood_data_given_model(kw["kwargs_lens"], kw["kwargs_source"], kw["kwargs_lens_light"],
                                                  kw["kwargs_ps"])
        assert abs(logl[b] - want) <= 1e-11 * abs(want), (b, logl[b], want)


def test_device_ensemble_sampler_on_the_image_likelihood():
    """Sampler.mcmc_emcee's boundary (sampler.py:112-212): walkers on the device, two batched likelihood calls per
    step; the chain must stay inside the prior box, keep its stored log-probabilities consistent with the likelihood
    and move (acceptance fraction) while improving on a deliberately offset start."""
    import torch

    from jaxtronomy_b200.Sampling.sampler import mcmc_emcee

    c, lin, param, like = _setup()
    x0 = np.asarray(param.kwargs2args(c.meta["kwargs_lens"], c.meta["kwargs_source"], c.meta["kwargs_lens_light"]))
    lo, hi = param.param_limits()
    n_walkers, n_run, n_burn = 32, 30, 10
    samples, dist = mcmc_emcee(like, n_walkers, n_run, n_burn, x0, 1e-3 * np.maximum(np.abs(x0), 0.1), seed=4)
    assert samples.shape == (n_walkers * n_run, x0.size) and dist.shape == (n_walkers * n_run,)
    assert np.all(samples >= lo) and np.all(samples <= hi) and np.all(np.isfinite(dist))
    s = mcmc_emcee.last_sampler
    acc = s.acceptance_fraction
    assert 0.05 < acc.mean() < 0.95, acc.mean()
    # stored log-probabilities == the likelihood of the stored positions
    last = torch.as_tensor(s.get_chain()[-1], device="cuda")
    np.testing.assert_allclose(s.get_log_prob()[-1], like.logL_batch_device(last).cpu().numpy(), rtol=1e-12)
    with pytest.raises(ValueError):
        mcmc_emcee(like, 8, 2, 0, x0, 1e-3 * np.ones_like(x0), mpi=True)


def test_sampler_front_end_pso_then_mcmc():
    """Sampler(likelihoodModule).pso / .mcmc_emcee with the reference's signatures (sampler.py:16-212)."""
    from jaxtronomy_b200.Sampling.sampler import Sampler

    c, lin, param, like = _setup()
    x0 = np.asarray(param.kwargs2args(c.meta["kwargs_lens"], c.meta["kwargs_source"], c.meta["kwargs_lens_light"]))
    sampler = Sampler(like)
    sig = 0.02 * np.maximum(np.abs(x0), 0.1)
    start = x0 + 0.5 * sig
    result, (logl_list, pos_list, vel_list) = sampler.pso(32, 25, lower_start=x0 - sig, upper_start=x0 + sig,
                                                           init_pos=start, seed=2)
    assert result.shape == x0.shape and len(logl_list) == 25 and len(pos_list) == 25 and vel_list == []
    assert all(b >= a for a, b in zip(logl_list[:-1], logl_list[1:]))  # the global best never gets worse
    assert logl_list[-1] >= like.logL(start) and abs(like.logL(np.asarray(pos_list[-1])) - logl_list[-1]) <= 1e-9 * abs(logl_list[-1])
    np.testing.assert_array_equal(np.asarray(pos_list[-1]), result)
    samples, dist = sampler.mcmc_emcee(16, 10, 5, result, 1e-3 * np.maximum(np.abs(result), 0.1), seed=1)
    assert samples.shape == (160, x0.size) and np.all(np.isfinite(dist))
    with pytest.raises(ValueError):
        sampler.pso(8, 2, mpi=True)
