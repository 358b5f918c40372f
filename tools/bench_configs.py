#!/usr/bin/env python
"""Secondary measurements on the other BASELINE.json configurations (one GPU).  Not the driver's bench line
(that is bench.py); these feed profiles/r1_configs.md.

    python tools/bench_configs.py [c2] [c3] [c4] [c5]
"""

import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from bench import psf_kernel, static_config  # noqa: E402
from jaxtronomy_b200 import _lib  # noqa: E402
from jaxtronomy_b200.Data.imaging_data import ImageData  # noqa: E402
from jaxtronomy_b200.Data.psf import PSF  # noqa: E402
from jaxtronomy_b200.ImSim.image_linear_solve import ImageLinearFit  # noqa: E402
from jaxtronomy_b200.ImSim.MultiBand.multi_linear import MultiLinear  # noqa: E402
from jaxtronomy_b200.ImSim.Numerics.grid import RegularGrid  # noqa: E402
from jaxtronomy_b200.LensModel.lens_model import LensModel  # noqa: E402
from jaxtronomy_b200.LensModel.lens_model_bulk import LensModelBulk  # noqa: E402
from jaxtronomy_b200.LightModel.light_model import LightModel  # noqa: E402
from jaxtronomy_b200.Sampling.likelihood import LikelihoodModule  # noqa: E402
from jaxtronomy_b200.Sampling.parameters import Param  # noqa: E402


def timed(fn, steps=3, warmup=2):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def c2():
    """EPL+SHEAR ray shooting, 180x180 ss 3 grid, 1024 parameter sets."""
    st = static_config()
    grid = RegularGrid(180, 180, st["kwargs_data"]["transform_pix2angle"], st["kwargs_data"]["ra_at_xy_0"],
                       st["kwargs_data"]["dec_at_xy_0"], 3)
    x, y = grid.coordinates_evaluate
    xt, yt = torch.as_tensor(x, device="cuda"), torch.as_tensor(y, device="cuda")
    rng = np.random.default_rng(1)
    B = 1024
    em, ea = rng.uniform(0.15, 0.35, B), rng.uniform(0, 2 * np.pi, B)
    kw = [dict(theta_E=torch.as_tensor(rng.uniform(0.8, 1.6, B), device="cuda"),
               gamma=torch.as_tensor(rng.uniform(1.7, 2.3, B), device="cuda"),
               e1=torch.as_tensor(em * np.cos(ea), device="cuda"), e2=torch.as_tensor(em * np.sin(ea), device="cuda"),
               center_x=0.0, center_y=0.0),
          dict(gamma1=torch.as_tensor(rng.uniform(-0.08, 0.08, B), device="cuda"), gamma2=0.01, ra_0=0.0, dec_0=0.0)]
    lm = LensModel(["EPL", "SHEAR"])
    ms = timed(lambda: lm.ray_shooting(xt, yt, kw))
    n = B * x.size
    return dict(config="C2 EPL+SHEAR ray shooting 180x180 ss3 x 1024 sets", ms=ms, rays_per_s=n / ms * 1e3,
                sets_per_s=B / ms * 1e3, out_GBps=n * 16 / ms / 1e6,
                fp64_TFLOPs_algorithmic=n * 418 / ms / 1e9)


def c3():
    """EPL_MULTIPOLE_M1M3M4+SHEAR, SERSIC_ELLIPSE source + lens light, 21x21 pixel PSF, ImageLinearFit, 8192 sets."""
    st = static_config()
    rng = np.random.default_rng(2)
    data = ImageData(image_data=rng.normal(0, 0.05, (180, 180)) + 0.1, **st["kwargs_data"])
    psf = PSF(psf_type="PIXEL", kernel_point_source=psf_kernel(21))
    km = dict(lens_model_list=["EPL_MULTIPOLE_M1M3M4", "SHEAR"], source_light_model_list=["SERSIC_ELLIPSE"],
              lens_light_model_list=["SERSIC_ELLIPSE"])
    lin = ImageLinearFit(data, psf, LensModel(km["lens_model_list"]), LightModel(km["source_light_model_list"]),
                         LightModel(km["lens_light_model_list"]), kwargs_numerics=dict(supersampling_factor=3))
    param = Param(km, kwargs_fixed_lens=[{}, {"ra_0": 0.0, "dec_0": 0.0}])
    like = LikelihoodModule(lin, param)
    lo, hi = param.param_limits()
    B = 8192
    kl = [dict(theta_E=1.2, gamma=2.0, e1=0.2, e2=-0.1, center_x=0.0, center_y=0.0, a1_a=0.01, delta_phi_m1=0.1,
               a3_a=0.01, delta_phi_m3=0.1, a4_a=0.01, delta_phi_m4=0.1), dict(gamma1=0.03, gamma2=0.01, ra_0=0, dec_0=0)]
    ks = [dict(amp=1, R_sersic=0.3, n_sersic=2.0, e1=0.1, e2=0.1, center_x=0.05, center_y=0.0)]
    kll = [dict(amp=1, R_sersic=0.6, n_sersic=3.0, e1=0.05, e2=0.0, center_x=0.0, center_y=0.0)]
    x0 = param.kwargs2args(kl, ks, kll)
    pos = torch.as_tensor(x0[None, :] + rng.normal(0, 0.02, (B, x0.size)), device="cuda")
    ms = timed(lambda: like.logL_batch_device(pos))
    flop = 291600 * (713 + 3 * 80) + 32400 * (1764 + 10 + 24 + 4)
    return dict(config="C3 EPL_MULTIPOLE_M1M3M4+SHEAR, 2 SERSIC_ELLIPSE (n_lin=2), 21x21 PSF, 8192 sets", ms=ms,
                images_per_s=B / ms * 1e3, fp64_TFLOPs_algorithmic=B * flop / ms / 1e9)


def c3grad():
    """C3's model through the gradient path: logL and d logL / d 27 sampled parameters, 1184 sets."""
    st = static_config()
    rng = np.random.default_rng(2)
    data = ImageData(image_data=rng.normal(0, 0.05, (180, 180)) + 0.1, **st["kwargs_data"])
    psf = PSF(psf_type="PIXEL", kernel_point_source=psf_kernel(21))
    km = dict(lens_model_list=["EPL_MULTIPOLE_M1M3M4", "SHEAR"], source_light_model_list=["SERSIC_ELLIPSE"],
              lens_light_model_list=["SERSIC_ELLIPSE"])
    lin = ImageLinearFit(data, psf, LensModel(km["lens_model_list"]), LightModel(km["source_light_model_list"]),
                         LightModel(km["lens_light_model_list"]), kwargs_numerics=dict(supersampling_factor=3))
    param = Param(km, kwargs_fixed_lens=[{}, {"ra_0": 0.0, "dec_0": 0.0}])
    like = LikelihoodModule(lin, param)
    B = 1184
    kl = [dict(theta_E=1.2, gamma=2.0, e1=0.2, e2=-0.1, center_x=0.0, center_y=0.0, a1_a=0.01, delta_phi_m1=0.1,
               a3_a=0.01, delta_phi_m3=0.1, a4_a=0.01, delta_phi_m4=0.1), dict(gamma1=0.03, gamma2=0.01, ra_0=0, dec_0=0)]
    ks = [dict(amp=1, R_sersic=0.3, n_sersic=2.0, e1=0.1, e2=0.1, center_x=0.05, center_y=0.0)]
    kll = [dict(amp=1, R_sersic=0.6, n_sersic=3.0, e1=0.05, e2=0.0, center_x=0.0, center_y=0.0)]
    x0 = param.kwargs2args(kl, ks, kll)
    pos = torch.as_tensor(x0[None, :] + rng.normal(0, 0.02, (B, x0.size)), device="cuda")
    ms = timed(lambda: like.logL_and_grad_batch_device(pos))
    ms_f = timed(lambda: like.logL_batch_device(pos))
    return dict(config="C3 model, logL + gradient (%d sampled parameters), %d sets" % (x0.size, B), ms=ms,
                gradients_per_s=B / ms * 1e3, forward_only_ms=ms_f, grad_over_forward=ms / ms_f)


def c4():
    """1 EPL + 1000 TNFW on a 1000x1000 ray grid (substructure notebook shape)."""
    rng = np.random.default_rng(3)
    nh = 1000
    names = ["EPL"] + ["TNFW"] * nh
    kws = [dict(theta_E=1.0, gamma=2.0, e1=0.1, e2=0.1, center_x=0.0, center_y=0.0)]
    kws += [dict(Rs=float(rng.uniform(0.01, 0.1)), alpha_Rs=float(rng.uniform(1e-4, 5e-3)),
                 r_trunc=float(rng.uniform(0.05, 0.5)), center_x=float(rng.uniform(-2, 2)),
                 center_y=float(rng.uniform(-2, 2))) for _ in range(nh)]
    g = np.linspace(-2, 2, 1000)
    xx, yy = np.meshgrid(g, g)
    xt, yt = torch.as_tensor(xx.ravel(), device="cuda"), torch.as_tensor(yy.ravel(), device="cuda")
    lmb = LensModelBulk(["EPL", "TNFW"], multi_plane=False)
    rs = lmb.prepare_ray_shooting_kwargs(names, kws)
    ms = timed(lambda: lmb.ray_shooting(xt, yt, rs))
    return dict(config="C4 EPL + 1000 TNFW, 1000x1000 rays", ms=ms, deflection_evals_per_s=1e6 * 1001 / ms * 1e3,
                fp64_TFLOPs_algorithmic=1e6 * (404 + 1000 * 230) / ms / 1e9)


def c5():
    """3 planes (EPL, NFW_ELLIPSE_CSE, SIS), 4 bands of 180x180 (ss 1, 1, 2, 1), 4096 sets, multi-band linear
    likelihood; model, redshifts, PSFs and numerics are the ones of the reference fixture c5_* (parity:
    tests/test_gpu_baseline_shapes.py)."""
    rng = np.random.default_rng(4)
    km = dict(lens_model_list=["EPL", "NFW_ELLIPSE_CSE", "SIS"], lens_redshift_list=[0.5, 0.3, 0.8], z_source=1.5,
              multi_plane=True, source_light_model_list=["SERSIC_ELLIPSE"], lens_light_model_list=["SERSIC_ELLIPSE"])
    st = static_config()
    bands = []
    for b in range(4):
        kd = dict(st["kwargs_data"], image_data=rng.normal(0, 0.05, (180, 180)) + 0.1)
        # numerics per band as in tests/golden/baseline_shapes.npz (c5_static): band 2 is supersampled
        bands.append([kd, dict(psf_type="PIXEL", kernel_point_source=psf_kernel(11, 1.5 + 0.2 * b)),
                      dict(supersampling_factor=2 if b == 2 else 1)])
    ml = MultiLinear(bands, km)
    param = Param(km)
    like = LikelihoodModule(ml, param)
    kl = [dict(theta_E=1.1, gamma=2.0, e1=0.1, e2=-0.1, center_x=0.0, center_y=0.0),
          dict(Rs=1.5, alpha_Rs=0.2, e1=0.05, e2=0.05, center_x=0.3, center_y=-0.2),
          dict(theta_E=0.1, center_x=-0.5, center_y=0.4)]
    ks = [dict(amp=1, R_sersic=0.3, n_sersic=2.0, e1=0.1, e2=0.1, center_x=0.05, center_y=0.0)]
    kll = [dict(amp=1, R_sersic=0.6, n_sersic=3.0, e1=0.05, e2=0.0, center_x=0.0, center_y=0.0)]
    x0 = param.kwargs2args(kl, ks, kll)
    B = 4096
    pos = torch.as_tensor(x0[None, :] + rng.normal(0, 0.01, (B, x0.size)), device="cuda")
    ms = timed(lambda: like.logL_batch_device(pos), steps=2, warmup=1)
    return dict(config="C5 multi-plane EPL+NFW_ELLIPSE_CSE+SIS, 4 bands 180x180 ss 1/1/2/1, 4096 sets", ms=ms,
                band_images_per_s=4 * B / ms * 1e3, sets_per_s=B / ms * 1e3)


def c1(mcmc=False, lbfgs=False):
    """The Einstein-ring notebook's PSO (notebooks/modeling_a_simple_Einstein_ring.ipynb: 256 particles x 200
    iterations, 100x100 px, EPL+SHEAR, SERSIC_ELLIPSE source + lens light, 4-image point source, Gaussian PSF, ss 1,
    linear solver; published: lenstronomy 6 min 43 s, JAXtronomy on a CPU node 12 s) with the GPU-resident swarm."""
    from jaxtronomy_b200.PointSource.point_source import PointSource
    from jaxtronomy_b200.Sampling.sampler import DeviceParticleSwarmOptimizer
    from jaxtronomy_b200.Util import simulation_util as sim_util

    rng = np.random.default_rng(5)
    kd = sim_util.data_configure_simple(100, 0.05, 100.0, 0.5, inverse=False)
    km = dict(lens_model_list=["EPL", "SHEAR"], source_light_model_list=["SERSIC_ELLIPSE"],
              lens_light_model_list=["SERSIC_ELLIPSE"], point_source_model_list=["LENSED_POSITION"])
    lens = LensModel(km["lens_model_list"])
    psf = PSF(psf_type="GAUSSIAN", fwhm=0.1, truncation=5)
    kl = [dict(theta_E=1.0, gamma=2.0, e1=0.1, e2=-0.05, center_x=0.0, center_y=0.0),
          dict(gamma1=0.03, gamma2=-0.01, ra_0=0, dec_0=0)]
    ks = [dict(amp=16.0, R_sersic=0.1, n_sersic=1.0, e1=0.1, e2=0.1, center_x=0.1, center_y=0.05)]
    kll = [dict(amp=16.0, R_sersic=0.6, n_sersic=2.0, e1=-0.1, e2=0.1, center_x=0.0, center_y=0.0)]
    kps = [dict(ra_image=np.array([1.02, -0.71, 0.43, -0.52]), dec_image=np.array([0.41, -0.83, -0.98, 0.95]),
                point_amp=np.array([100.0, 80.0, 60.0, 40.0]))]

    def make(image):
        return ImageLinearFit(ImageData(**dict(kd, image_data=image)), psf, lens, LightModel(km["source_light_model_list"]),
                              LightModel(km["lens_light_model_list"]), point_source_class=PointSource(["LENSED_POSITION"], lens_model=lens),
                              kwargs_numerics=dict(supersampling_factor=1))

    truth = make(np.zeros((100, 100))).image(kl, ks, kll, kps)
    lin = make(truth + rng.normal(0, 0.5, truth.shape))
    param = Param(km, kwargs_fixed_lens=[{}, {"ra_0": 0.0, "dec_0": 0.0}], num_point_source_list=[4])
    if lbfgs:
        return _c1_lbfgs(LikelihoodModule(lin, param, source_position_likelihood=False), param,
                         param.kwargs2args(kl, ks, kll, kwargs_ps=kps))
    like = LikelihoodModule(lin, param)
    x0 = param.kwargs2args(kl, ks, kll, kwargs_ps=kps)
    sig = np.full(x0.size, 0.05)
    if mcmc:
        return _c1_mcmc(like, x0)
    n_part, n_iter = 256, 200

    def run():
        pso = DeviceParticleSwarmOptimizer(like, x0 - sig, x0 + sig, particle_count=n_part, seed=1)
        best, hist = pso.optimize(max_iter=n_iter)
        return hist

    run()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    hist = run()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    return dict(config="C1 notebook PSO: 256 particles x 200 iterations, 100x100, EPL+SHEAR, 2 SERSIC_ELLIPSE, "
                       "4-image point source, Gaussian PSF, linear solver (n_lin = 6), GPU-resident swarm",
                wall_s=dt, likelihoods_per_s=n_part * n_iter / dt, logL_first=hist[0], logL_last=hist[-1],
                logL_truth=like.logL(x0), published="lenstronomy 403 s, JAXtronomy (CPU node) 12 s")


def _c1_mcmc(like, x0):
    """The same notebook's MCMC: 256 walkers x 100 steps (published: lenstronomy 211.1 s, JAXtronomy on a CPU node
    8.59 s) with the device-resident stretch-move sampler."""
    from jaxtronomy_b200.Sampling.sampler import mcmc_emcee

    n_walkers, n_run = 256, 100
    sig = np.full(x0.size, 0.005)

    def run():
        return mcmc_emcee(like, n_walkers, n_run, 0, x0, sig, seed=3)

    run()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    samples, dist = run()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    acc = float(mcmc_emcee.last_sampler.acceptance_fraction.mean())
    return dict(config="C1 notebook MCMC: 256 walkers x 100 steps (2 half-ensemble likelihood batches per step), "
                       "same model as the PSO run, device-resident stretch move",
                wall_s=dt, likelihoods_per_s=n_walkers * n_run / dt, acceptance_fraction=acc,
                logL_max=float(dist.max()), logL_truth=like.logL(x0), samples=list(samples.shape),
                published="lenstronomy 211.1 s, JAXtronomy (CPU node) 8.59 s")


def c1mcmc():
    return c1(mcmc=True)


def c1lbfgs():
    """The same notebook's gradient descent (notebooks/modeling_a_simple_Einstein_ring.ipynb:913-926: optax L-BFGS,
    maxiter 1000, 1 chain, tolerance 0.01; published 730 iterations in 36.5 s with JAXtronomy on a CPU node) through
    OptaxMinimizer on lsim_logl_grad: EPL + SHEAR, two SERSIC_ELLIPSE, a 4-image LENSED_POSITION point source whose
    positions are differentiated by k_point_sources_grad.  The source-position penalty of the position likelihood is
    switched off (its gradient is outside the scope); 8 chains are run at once to show what the batch axis buys."""
    return c1(lbfgs=True)


def _c1_lbfgs(like, param, x0):
    from jaxtronomy_b200.Sampling.Samplers.optax import OptaxMinimizer

    lo, hi = param.param_limits()
    sigma = 0.02 * np.maximum(np.abs(x0), 0.05)
    out = {}
    for chains in (1, 8):
        opt = OptaxMinimizer(like, x0, sigma, np.asarray(lo), np.asarray(hi), maxiter=1000)
        opt.run(num_chains=chains, tol=0.01, rng_seed=1)  # warm-up (plan creation, first launches)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        best = opt.run(num_chains=chains, tol=0.01, rng_seed=1)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        out["chains_%d" % chains] = dict(wall_s=dt, iterations=[int(i) for i in opt.num_iterations],
                                         evaluations=int(opt.num_evaluations), logL_best=float(like.logL(best)))
    return dict(config="C1 notebook L-BFGS (OptaxMinimizer over lsim_logl_grad): 100x100, EPL+SHEAR, 2 SERSIC_ELLIPSE, "
                       "4-image point source, Gaussian PSF, linear solver, maxiter 1000, tol 0.01",
                logL_truth=like.logL(x0), published="JAXtronomy (CPU node) 36.5 s for 730 iterations, 1 chain", **out)


def mixed():
    """Headline workload with the default prior box e1, e2 ~ U(-0.5, 0.5) (SURVEY.md 8(d) "mixed-nmax" run): the
    series length differs from set to set (nmax 10 ... 50), so CTAs of one launch no longer cost the same."""
    from bench import build_headline

    B = 16384
    cfg = build_headline(n_sets=B, seed=1234)
    rng = np.random.default_rng(7)
    p = cfg["params"].clone()
    p[:, 2] = torch.as_tensor(rng.uniform(-0.5, 0.5, B), device="cuda")
    p[:, 3] = torch.as_tensor(rng.uniform(-0.5, 0.5, B), device="cuda")
    emod = torch.sqrt(p[:, 2] ** 2 + p[:, 3] ** 2).cpu().numpy()
    nmax = np.select([emod < 0.12, emod < 0.4, emod < 0.6, emod < 0.73], [10, 20, 35, 50], 100)
    ms = timed(lambda: cfg["model"].logL_linear_batch(p))
    # what a launch that never mixed series lengths could reach: every bracket timed on its own (all 16,384 sets moved
    # into that bracket, angles kept), combined with the histogram's weights
    ang = torch.atan2(p[:, 3], p[:, 2])
    per = {}
    for k, (lo, hi) in {10: (0.02, 0.12), 20: (0.12, 0.4), 35: (0.4, 0.6), 50: (0.6, 0.73), 100: (0.73, 0.86)}.items():
        if not (nmax == k).any():
            continue
        sel = torch.as_tensor(emod[nmax == k], device="cuda")
        mod = sel[torch.arange(B, device="cuda") % sel.numel()]  # the bracket's own |e| distribution, tiled
        pk = p.clone()
        pk[:, 2], pk[:, 3] = mod * torch.cos(ang), mod * torch.sin(ang)
        per[k] = timed(lambda: cfg["model"].logL_linear_batch(pk))
    expect = sum(per[int(k)] * int((nmax == k).sum()) / B for k in np.unique(nmax))
    return dict(config="headline with e1,e2 ~ U(-0.5,0.5) (mixed nmax), 16384 sets", ms=ms, images_per_s=B / ms * 1e3,
                nmax_histogram={int(k): int((nmax == k).sum()) for k in np.unique(nmax)},
                ms_single_bracket={int(k): v for k, v in per.items()}, ms_expected_if_never_mixed=expect,
                mixed_over_expected=ms / expect)


CONFIGS = {"c1_pso": c1, "c1_lbfgs": c1lbfgs, "c2": c2, "c3": c3, "c4": c4, "c5": c5, "mixed": mixed}  # what bench.py reports under "secondary"


if __name__ == "__main__":
    which = sys.argv[1:] or ["c2", "c3", "c4", "c5"]
    t0 = time.time()
    for name in which:
        out = globals()[name]()
        out["launches_total"] = _lib.launch_count()
        print(json.dumps(out), flush=True)
    print("# wall %.1f s" % (time.time() - t0))
