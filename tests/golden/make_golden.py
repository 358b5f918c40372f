#!/usr/bin/env python
"""Generate the committed golden fixtures from the UNMODIFIED reference source.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

It imports ``jaxtronomy`` from /root/reference through ``refshim`` (NumPy-backed jax, restated
lenstronomy host helpers -- see refshim.py for what is and is not substituted) and writes

    tests/golden/profiles.npz     deflections of every hot-path lens profile + light profiles + 2F1 sweep
    tests/golden/lens_models.npz  LensModel / LensModelBulk / MultiPlane ray shooting
    tests/golden/imsim.npz        grid bookkeeping, re-bin + convolution, ImageModel, ImageLinearFit
    tests/golden/gradients.npz    central differences of the reference log-likelihood (python ... make_golden.py gradients)
    tests/golden/point_sources.npz  point-source rendering, ImageModel / ImageLinearFit with point sources
                                  (python tests/golden/make_golden.py point_sources  regenerates only this file)

Parameter sets follow the reference's own tests (test/test_LensModel/test_Profiles/*.py,
test/test_ImSim/*.py) plus the clamps/edge cases listed in SURVEY.md 7 "Hard parts".
Each case is stored as ``<case>/in/<name>`` and ``<case>/out/<name>`` arrays; scalar kwargs are
stored as 0-d arrays, strings as JSON in ``<case>/meta``.
"""

import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import refshim  # noqa: E402

refshim.install()

from jaxtronomy.LensModel.lens_model import LensModel  # noqa: E402
from jaxtronomy.LensModel.lens_model_bulk import LensModelBulk  # noqa: E402
from jaxtronomy.LightModel.light_model import LightModel  # noqa: E402
from jaxtronomy.Util.hyp2f1_util import hyp2f1_lopez_temme_8  # noqa: E402
from jaxtronomy.LensModel.Profiles.epl import EPLMajorAxis  # noqa: E402
from jaxtronomy.ImSim.Numerics.grid import RegularGrid  # noqa: E402
from jaxtronomy.ImSim.Numerics.numerics import Numerics  # noqa: E402
from jaxtronomy.ImSim.image_model import ImageModel  # noqa: E402
from jaxtronomy.ImSim.image_linear_solve import ImageLinearFit  # noqa: E402
from jaxtronomy.Data.imaging_data import ImageData  # noqa: E402
from lenstronomy.Data.psf import PSF  # noqa: E402
from lenstronomy.Data.pixel_grid import PixelGrid  # noqa: E402
import lenstronomy.Util.simulation_util as sim_util  # noqa: E402


class Store:
    def __init__(self):
        self.d = {}

    def add(self, case, inputs, outputs, meta=None):
        for k, v in inputs.items():
            self.d["%s/in/%s" % (case, k)] = np.asarray(v)
        for k, v in outputs.items():
            self.d["%s/out/%s" % (case, k)] = np.asarray(v)
        self.d["%s/meta" % case] = np.asarray(json.dumps(meta or {}, default=lambda o: np.asarray(o).tolist()))

    def save(self, name):
        path = os.path.join(HERE, name)
        np.savez_compressed(path, **self.d)
        print("wrote %s: %d arrays, %.1f KiB" % (name, len(self.d), os.path.getsize(path) / 1024))


def points(rng, n=97, span=3.0):
    """Scattered points plus the awkward ones: the origin, axis points (zero crossings), far field."""
    x = rng.uniform(-span, span, n)
    y = rng.uniform(-span, span, n)
    x = np.concatenate([x, [0.0, 0.0, 1.0, -1.0, 0.0, 1e-7, 25.0, 1e-12]])
    y = np.concatenate([y, [0.0, 1.0, 0.0, 0.0, -1.3, -1e-7, -40.0, 0.0]])
    return x, y


def e_from(mod, ang):
    return mod * np.cos(ang), mod * np.sin(ang)


# ------------------------------------------------------------------------------------------------
def gen_profiles():
    st = Store()
    rng = np.random.default_rng(20261017)
    x, y = points(rng)

    def lens_case(case, name, kw, xx=x, yy=y, profile_kwargs=None):
        lm = LensModel([name], profile_kwargs_list=[profile_kwargs or {}])
        fx, fy = lm.lens_model.alpha(xx, yy, [kw])
        st.add(case, dict(x=xx, y=yy, **kw), dict(f_x=fx, f_y=fy),
               dict(profile=name, profile_kwargs=profile_kwargs or {}))

    # EPL: one case per nmax bracket (|e| thresholds epl.py:311-317), plus e==0, q->clamp, gamma extremes
    n = 0
    for emod, ang, gamma in [(0.0, 0.0, 2.0), (0.05, 0.3, 1.9), (0.119, -2.0, 1.5), (0.12, 1.0, 2.0),
                             (0.25, 0.7, 2.1), (0.399, 2.5, 2.5), (0.45, -0.4, 1.7), (0.65, 1.9, 2.2),
                             (0.8, 0.1, 2.0), (0.9, -1.1, 1.8), (0.95, 0.6, 2.3), (0.99995, 0.2, 2.0)]:
        e1, e2 = e_from(emod, ang)
        lens_case("epl_%02d" % n, "EPL", dict(theta_E=1.3, gamma=gamma, e1=e1, e2=e2, center_x=0.1, center_y=-0.2))
        n += 1
    # the reference's own sweep (test_epl.py:199-206) on the major-axis function
    ys = np.linspace(-10, 10, 101)
    xs = np.ones_like(ys)
    fx_all, fy_all = [], []
    qs = np.linspace(0.03, 1, 200)[::-1]
    for q in qs[::8]:
        fx, fy = EPLMajorAxis.derivatives(xs, ys, 1.7, 0.9, q)
        fx_all.append(fx)
        fy_all.append(fy)
    st.add("epl_major_sweep", dict(x=xs, y=ys, b=1.7, t=0.9, q=qs[::8]), dict(f_x=np.array(fx_all), f_y=np.array(fy_all)))

    for i, (e1, e2) in enumerate([(0.1, -0.05), (0.0, 0.0), (-0.4, 0.3), (0.7, 0.7)]):
        lens_case("sie_%d" % i, "SIE", dict(theta_E=1.1, e1=e1, e2=e2, center_x=-0.1, center_y=0.3))
        lens_case("nie_%d" % i, "NIE", dict(theta_E=0.9, e1=e1, e2=e2, s_scale=0.1 + 0.2 * i, center_x=0.2, center_y=0.0))
    lens_case("sie_epl", "SIE", dict(theta_E=1.1, e1=0.2, e2=0.1, center_x=0.0, center_y=0.0),
              profile_kwargs=dict(NIE=False, gamma=2))
    lens_case("sis_0", "SIS", dict(theta_E=1.4, center_x=0.0, center_y=0.0))
    lens_case("sis_1", "SIS", dict(theta_E=0.7, center_x=0.3, center_y=-0.6))
    lens_case("shear_0", "SHEAR", dict(gamma1=0.05, gamma2=-0.03, ra_0=0.0, dec_0=0.0))
    lens_case("shear_1", "SHEAR", dict(gamma1=-0.2, gamma2=0.4, ra_0=0.5, dec_0=-1.0))
    lens_case("convergence_0", "CONVERGENCE", dict(kappa=0.12, ra_0=0.1, dec_0=0.2))
    # NFW: X<1, X==1 exactly, X>1, X below the 1e-6 clamp, tiny Rs clamp
    xn = np.concatenate([x, [2.4, 0.0, 1e-9]])
    yn = np.concatenate([y, [0.0, 2.4, 0.0]])
    lens_case("nfw_0", "NFW", dict(Rs=2.4, alpha_Rs=1.5, center_x=0.0, center_y=0.0), xn, yn)
    lens_case("nfw_1", "NFW", dict(Rs=0.3, alpha_Rs=0.7, center_x=1.0, center_y=2.0))
    lens_case("nfw_2", "NFW", dict(Rs=1e-8, alpha_Rs=1e-9, center_x=0.0, center_y=0.0))
    # TNFW: reference test parameters (test_tnfw.py:48-79) and x==1, tau<1, tau>1
    lens_case("tnfw_0", "TNFW", dict(Rs=1.4, alpha_Rs=3.5, r_trunc=3.2, center_x=0.0, center_y=0.0),
              np.concatenate([x, [1.4]]), np.concatenate([y, [0.0]]))
    lens_case("tnfw_1", "TNFW", dict(Rs=2.4, alpha_Rs=1.5, r_trunc=1.2, center_x=1.0, center_y=2.0))
    lens_case("tnfw_2", "TNFW", dict(Rs=4.0, alpha_Rs=0.5, r_trunc=8.0, center_x=1.0, center_y=2.0))
    lens_case("tnfw_3", "TNFW", dict(Rs=0.01, alpha_Rs=0.002, r_trunc=0.07, center_x=-0.5, center_y=0.25))
    for m in (1, 2, 3, 4):
        lens_case("multipole_m%d" % m, "MULTIPOLE",
                  dict(m=m, a_m=0.05, phi_m=0.3 * m, center_x=0.1, center_y=-0.1, r_E=1.2))
    for i, (emod, ang) in enumerate([(0.2, 0.4), (0.0, 0.0), (0.45, -1.0)]):
        e1, e2 = e_from(emod, ang)
        lens_case("epl_m1m3m4_%d" % i, "EPL_MULTIPOLE_M1M3M4",
                  dict(theta_E=1.2, gamma=2.05, e1=e1, e2=e2, center_x=0.05, center_y=0.02, a1_a=0.02,
                       delta_phi_m1=0.5, a3_a=-0.01, delta_phi_m3=0.2, a4_a=0.015, delta_phi_m4=-0.1))
    for axis in ("product_avg", "major"):
        lens_case("cse_%s" % axis, "CSE", dict(a=1.3, s=0.4, e1=0.15, e2=-0.2, center_x=0.1, center_y=0.0),
                  profile_kwargs=dict(axis=axis))
    for hi in (True, False):
        lens_case("nfw_ellipse_cse_%d" % hi, "NFW_ELLIPSE_CSE",
                  dict(Rs=1.7, alpha_Rs=1.1, e1=0.12, e2=0.08, center_x=-0.05, center_y=0.1),
                  profile_kwargs=dict(high_accuracy=hi))
    lens_case("hernquist_ellipse_cse", "HERNQUIST_ELLIPSE_CSE",
              dict(sigma0=0.9, Rs=1.2, e1=-0.1, e2=0.2, center_x=0.0, center_y=0.05))

    # light profiles (test_sersic_ellipse.py / test_gaussian.py style) incl. clamps: R<1e-4, R/Rs > 1000
    def light_case(case, name, kw, xx=x, yy=y):
        lm = LightModel([name])
        st.add(case, dict(x=xx, y=yy, **kw), dict(flux=lm.surface_brightness(xx, yy, [kw])), dict(profile=name))

    xl = np.concatenate([x, [0.05, 0.05 + 1e-6, 200.0]])
    yl = np.concatenate([y, [0.1, 0.1, 0.0]])
    light_case("sersic_ellipse_0", "SERSIC_ELLIPSE",
               dict(amp=10.0, R_sersic=0.2, n_sersic=2.0, e1=0.1, e2=0.1, center_x=0.05, center_y=0.1), xl, yl)
    light_case("sersic_ellipse_1", "SERSIC_ELLIPSE",
               dict(amp=1.4, R_sersic=0.6, n_sersic=7.0, e1=0.3, e2=-0.4, center_x=0.0, center_y=0.0), xl, yl)
    light_case("sersic_ellipse_2", "SERSIC_ELLIPSE",
               dict(amp=2.0, R_sersic=0.1, n_sersic=0.5, e1=-0.2, e2=0.0, center_x=0.3, center_y=-0.3), xl, yl)
    light_case("sersic_ellipse_3", "SERSIC_ELLIPSE",
               dict(amp=2.0, R_sersic=-0.1, n_sersic=1.0, e1=0.0, e2=0.0, center_x=0.0, center_y=0.0), xl, yl)
    light_case("sersic_0", "SERSIC", dict(amp=1.0, R_sersic=0.1, n_sersic=2.0, center_x=0.0, center_y=0.0), xl, yl)
    light_case("sersic_1", "SERSIC", dict(amp=3.0, R_sersic=1.0, n_sersic=4.0, center_x=0.5, center_y=0.5), xl, yl)
    light_case("gaussian_0", "GAUSSIAN", dict(amp=4.0, sigma=0.3, center_x=0.1, center_y=-0.2), xl, yl)

    # 2F1 sweep in the EPL configuration (a=1, b=t/2, c=2-t/2, |z|=f)
    t = np.array([0.5, 0.9, 1.0, 1.4])
    ang = np.linspace(-np.pi, np.pi, 41)
    for nmax, f in [(10, 0.1), (20, 0.39), (35, 0.59), (50, 0.72), (100, 0.85), (200, 0.91), (500, 0.99)]:
        z = f * np.exp(1j * ang)
        out = np.array([hyp2f1_lopez_temme_8(1, tt / 2, 2 - tt / 2, z, nmax=nmax) for tt in t])
        st.add("hyp2f1_nmax%d" % nmax, dict(t=t, z=z, nmax=nmax), dict(h=out))
    st.save("profiles.npz")


# ------------------------------------------------------------------------------------------------
def gen_lens_models():
    st = Store()
    rng = np.random.default_rng(7)
    x, y = points(rng, n=61)
    names = ["EPL", "SHEAR", "SIS", "NFW"]
    kw = [dict(theta_E=1.2, gamma=2.1, e1=0.2, e2=-0.1, center_x=0.05, center_y=-0.02),
          dict(gamma1=0.03, gamma2=-0.02, ra_0=0.0, dec_0=0.0),
          dict(theta_E=0.2, center_x=0.8, center_y=0.5),
          dict(Rs=0.5, alpha_Rs=0.1, center_x=-0.7, center_y=0.3)]
    lm = LensModel(names)
    for k in (None, 0, (0, 2), [True, True, False, True]):
        bx, by = lm.ray_shooting(x, y, kw, k=k)
        st.add("single_plane_k%s" % json.dumps(k).replace(" ", ""), dict(x=x, y=y), dict(beta_x=bx, beta_y=by),
               dict(lens_model_list=names, kwargs=kw, k=k))

    # bulk single plane: 1 EPL + 40 TNFW (+ NULL padding to 48), substructure-notebook shape
    nh = 40
    bulk_names = ["EPL"] + ["TNFW"] * nh
    bulk_kw = [kw[0]] + [dict(Rs=float(rng.uniform(0.01, 0.1)), alpha_Rs=float(rng.uniform(1e-4, 5e-3)),
                              r_trunc=float(rng.uniform(0.05, 0.5)), center_x=float(rng.uniform(-2, 2)),
                              center_y=float(rng.uniform(-2, 2))) for _ in range(nh)]
    lmb = LensModelBulk(["EPL", "TNFW"], multi_plane=False)
    rs_kw = lmb.prepare_ray_shooting_kwargs(bulk_names, bulk_kw, num_deflectors=48)
    bx, by = lmb.ray_shooting(x, y, rs_kw)
    st.add("bulk_single_plane", dict(x=x, y=y, index_list=rs_kw["index_list"],
                                     **{"kw_" + k: v for k, v in rs_kw["all_kwargs"].items()}),
           dict(beta_x=bx, beta_y=by), dict(unique_lens_model_list=["EPL", "TNFW"], lens_model_list=bulk_names,
                                            kwargs=bulk_kw, num_deflectors=48))

    # multi-plane, config-5 shape: EPL (z=0.5), NFW_ELLIPSE_CSE (z=0.5, same plane), SIS (z=0.2), SHEAR (z=0.8)
    mp_names = ["EPL", "NFW_ELLIPSE_CSE", "SIS", "SHEAR"]
    mp_z = [0.5, 0.5, 0.2, 0.8]
    mp_kw = [kw[0], dict(Rs=1.5, alpha_Rs=0.3, e1=0.1, e2=0.05, center_x=0.0, center_y=0.1),
             dict(theta_E=0.15, center_x=0.6, center_y=-0.4), kw[1]]
    z_source = 1.5
    lmm = LensModel(mp_names, z_source=z_source, lens_redshift_list=mp_z, multi_plane=True)
    bx, by = lmm.ray_shooting(x, y, mp_kw)
    mp = lmm.lens_model
    base = mp.multi_plane_base
    st.add("multi_plane", dict(x=x, y=y, sorted_index=np.asarray(base.sorted_redshift_index),
                               T_ij_list=np.asarray(base.T_ij_list, dtype=float), T_z_list=np.asarray(base.T_z_list, dtype=float),
                               reduced2physical=np.asarray(base._reduced2physical_factor, dtype=float),
                               T_ij_stop=float(mp.T_ij_stop), T_z_source=float(mp._T_z_source)),
           dict(beta_x=bx, beta_y=by),
           dict(lens_model_list=mp_names, lens_redshift_list=mp_z, z_source=z_source, kwargs=mp_kw,
                cosmology="flat LCDM H0=70 Om0=0.3 stand-in (refshim.Background)"))
    st.save("lens_models.npz")


# ------------------------------------------------------------------------------------------------
def gauss_kernel(n, sigma):
    a = np.arange(n) - (n - 1) / 2
    k = np.exp(-(a[:, None] ** 2 + a[None, :] ** 2) / (2 * sigma**2))
    return k / k.sum()


def gen_imsim():
    st = Store()
    rng = np.random.default_rng(11)

    # grid bookkeeping: coordinates and index maps must be bit-exact (test_grid.py)
    M = np.array([[0.05, 0.002], [-0.001, 0.05]])
    for ss in (1, 2, 3):
        g = RegularGrid(6, 6, M, -0.15, -0.14, supersampling_factor=ss)
        ra, dec = g.coordinates_evaluate
        flux = rng.uniform(0, 1, ra.shape)
        low, high = g.flux_array2image_low_high(flux)
        st.add("grid_ss%d" % ss, dict(nx=6, ny=6, transform_pix2angle=M, ra_at_xy_0=-0.15, dec_at_xy_0=-0.14, ss=ss,
                                      flux=flux),
               dict(ra=ra, dec=dec, compute_indexes=g._compute_indexes, low=low,
                    high=(high if high is not None else np.zeros(0))))

    # re_size_convolve: PIXEL psf (fft and grid), GAUSSIAN psf (edge padded), NONE
    kd = sim_util.data_configure_simple(20, 0.08, 100, 0.05, inverse=True)
    pg = PixelGrid(20, 20, kd["transform_pix2angle"], kd["ra_at_xy_0"], kd["dec_at_xy_0"])
    kern = gauss_kernel(7, 1.3) + 0.02 * rng.uniform(0, 1, (7, 7)) / 49  # not separable, not symmetric
    for name, psf_kw, num_kw in [
        ("pixel_fft", dict(psf_type="PIXEL", kernel_point_source=kern), dict(supersampling_factor=3, convolution_type="fft")),
        ("pixel_grid", dict(psf_type="PIXEL", kernel_point_source=kern), dict(supersampling_factor=2, convolution_type="grid")),
        ("pixel_cut", dict(psf_type="PIXEL", kernel_point_source=kern), dict(supersampling_factor=1, convolution_kernel_size=5)),
        ("gaussian", dict(psf_type="GAUSSIAN", fwhm=0.2, truncation=5), dict(supersampling_factor=2)),
        ("gaussian_trunc", dict(psf_type="GAUSSIAN", fwhm=0.15, truncation=3), dict(supersampling_factor=1, truncation_conv=4)),
        ("none", dict(psf_type="NONE"), dict(supersampling_factor=2)),
    ]:
        psf = PSF(**psf_kw)
        psf.set_pixel_size(pg.pixel_width)
        num = Numerics(pg, psf, **num_kw)
        ra, dec = num.coordinates_evaluate
        flux = np.exp(-((ra - 0.1) ** 2 + (dec + 0.2) ** 2) / 0.05) + 0.01 * rng.uniform(0, 1, ra.shape)
        out = num.re_size_convolve(flux)
        outu = num.re_size_convolve(flux, unconvolved=True)
        ins = dict(flux=flux, transform_pix2angle=kd["transform_pix2angle"], ra_at_xy_0=kd["ra_at_xy_0"],
                   dec_at_xy_0=kd["dec_at_xy_0"])
        if "kernel_point_source" in psf_kw:
            ins["kernel"] = kern
        meta_psf = {k: v for k, v in psf_kw.items() if k != "kernel_point_source"}
        st.add("numerics_%s" % name, ins, dict(image=out, image_unconvolved=outu, ra=ra, dec=dec),
               dict(nx=20, ny=20, kwargs_psf=meta_psf, kwargs_numerics=num_kw))

    # ImageModel / ImageLinearFit cases
    def image_case(case, numPix, deltaPix, lens_names, kw_lens, src_names, kw_src, ll_names, kw_ll, psf_kw, num_kw,
                   mask=None, noise_map=False, exposure_map=False, inverse=False, profile_kwargs=None):
        kd = sim_util.data_configure_simple(numPix, deltaPix, 100.0, 0.05, inverse=inverse)
        if exposure_map:
            kd["exposure_time"] = 80.0 + 40.0 * rng.uniform(0, 1, (numPix, numPix))
        data = ImageData(**kd)
        psf = PSF(**psf_kw)
        lens = LensModel(lens_names, profile_kwargs_list=profile_kwargs)
        src = LightModel(src_names)
        ll = LightModel(ll_names)
        im = ImageModel(data, psf, lens, src, ll, kwargs_numerics=num_kw, likelihood_mask=mask)
        truth = np.asarray(im.image(kw_lens, kw_src, kw_ll))
        noisy = truth + rng.normal(0, 0.05, truth.shape) + rng.normal(0, 1, truth.shape) * np.sqrt(np.maximum(truth, 0) / 100.0)
        kd2 = dict(kd)
        kd2["image_data"] = noisy
        if noise_map:
            kd2["noise_map"] = 0.05 + 0.02 * rng.uniform(0, 1, truth.shape)
        data = ImageData(**kd2)
        im = ImageModel(data, psf, lens, src, ll, kwargs_numerics=num_kw, likelihood_mask=mask)
        lin = ImageLinearFit(data, psf, lens, src, ll, kwargs_numerics=num_kw, likelihood_mask=mask)
        # perturbed parameters -> the model differs from the data
        kw_lens_p = [dict(k) for k in kw_lens]
        for k in kw_lens_p:
            if "theta_E" in k:
                k["theta_E"] *= 1.02
        image = np.asarray(im.image(kw_lens_p, kw_src, kw_ll))
        image_unc = np.asarray(im.image(kw_lens_p, kw_src, kw_ll, unconvolved=True))
        src_only = np.asarray(im.source_surface_brightness(kw_src, kw_lens_p))
        ll_only = np.asarray(im.lens_surface_brightness(kw_ll))
        logL = float(im.likelihood_data_given_model(kw_lens_p, kw_src, kw_ll))
        # the reference's update_linear_kwargs writes the solved amplitudes into the kwargs dicts: pass copies
        cp = lambda kws: [dict(k) for k in kws]  # noqa: E731
        A = np.asarray(lin.linear_response_matrix(kw_lens_p, cp(kw_src), cp(kw_ll), None))
        model, _, _, param = lin.image_linear_solve(kw_lens_p, cp(kw_src), cp(kw_ll))
        logL_lin, param2 = lin.likelihood_data_given_model(kw_lens_p, cp(kw_src), cp(kw_ll))
        logL_marg, _ = lin.likelihood_data_given_model(kw_lens_p, cp(kw_src), cp(kw_ll), source_marg=True)
        ins = dict(image_data=noisy, transform_pix2angle=kd["transform_pix2angle"], ra_at_xy_0=kd["ra_at_xy_0"],
                   dec_at_xy_0=kd["dec_at_xy_0"], exposure_time=kd["exposure_time"], background_rms=0.05)
        num_meta = dict(num_kw)
        if "flux_evaluate_indexes" in num_meta:
            ins["flux_evaluate_indexes"] = num_meta.pop("flux_evaluate_indexes")
        if mask is not None:
            ins["likelihood_mask"] = mask
        if noise_map:
            ins["noise_map"] = kd2["noise_map"]
        if "kernel_point_source" in psf_kw:
            ins["kernel"] = psf_kw["kernel_point_source"]
        st.add(case, ins,
               dict(image=image, image_unconvolved=image_unc, source=src_only, lens_light=ll_only, logL=logL, A=A,
                    wls_model=np.asarray(model), param=np.asarray(param), logL_linear=float(logL_lin),
                    logL_marg=float(logL_marg)),
               dict(numPix=numPix, deltaPix=deltaPix, lens_model_list=lens_names, kwargs_lens=kw_lens_p,
                    source_model_list=src_names, kwargs_source=kw_src, lens_light_model_list=ll_names,
                    kwargs_lens_light=kw_ll, kwargs_psf={k: v for k, v in psf_kw.items() if k != "kernel_point_source"},
                    kwargs_numerics=num_meta, lens_profile_kwargs=profile_kwargs))

    sersic_src = dict(amp=12.0, R_sersic=0.25, n_sersic=2.0, e1=0.1, e2=0.1, center_x=0.05, center_y=0.1)
    sersic_ll = dict(amp=20.0, R_sersic=0.5, n_sersic=3.0, e1=-0.1, e2=0.05, center_x=0.02, center_y=-0.01)
    shear = dict(gamma1=0.03, gamma2=-0.02, ra_0=0.0, dec_0=0.0)
    # C1: SIE+SHEAR, SERSIC_ELLIPSE source, Gaussian PSF (Einstein-ring notebook shape, shrunk to 40x40)
    image_case("c1_sie_shear_gauss", 40, 0.1, ["SIE", "SHEAR"],
               [dict(theta_E=1.2, e1=0.1, e2=-0.05, center_x=0.02, center_y=-0.01), shear],
               ["SERSIC_ELLIPSE"], [sersic_src], ["SERSIC_ELLIPSE"], [sersic_ll],
               dict(psf_type="GAUSSIAN", fwhm=0.2, truncation=5), dict(supersampling_factor=1))
    # headline shape shrunk: EPL+SHEAR, ss 3, 11x11 non-separable pixel PSF, likelihood mask
    kern = gauss_kernel(11, 1.6) + 0.05 * rng.uniform(0, 1, (11, 11)) / 121
    mask = np.ones((36, 36))
    mask[:4, :] = 0
    mask[20:24, 10:15] = 0
    image_case("headline_epl_shear_pixel", 36, 0.09, ["EPL", "SHEAR"],
               [dict(theta_E=1.1, gamma=2.05, e1=0.2, e2=-0.1, center_x=0.02, center_y=-0.01), shear],
               ["SERSIC_ELLIPSE"], [sersic_src], ["SERSIC_ELLIPSE"], [sersic_ll],
               dict(psf_type="PIXEL", kernel_point_source=kern), dict(supersampling_factor=3), mask=mask)
    # test_image_linear_solve.py shape: SIS, 2 x SERSIC_ELLIPSE source + SERSIC lens light, Gaussian PSF truncation 5, ss 2
    e1, e2 = 0.1 / 1.9 * np.cos(0.4), 0.1 / 1.9 * np.sin(0.4)
    image_case("linear_sis_3comp", 40, 0.1, ["SIS"], [dict(theta_E=1.0, center_x=0.0, center_y=0.0)],
               ["SERSIC_ELLIPSE", "SERSIC_ELLIPSE"],
               [dict(amp=1.4, R_sersic=0.6, n_sersic=7.0, center_x=0.0, center_y=0.0, e1=e1, e2=e2),
                dict(amp=2.1, R_sersic=0.3, n_sersic=6.0, center_x=0.0, center_y=0.0, e1=e1, e2=e2)],
               ["SERSIC"], [dict(amp=1.0, R_sersic=0.1, n_sersic=2.0, center_x=0.0, center_y=0.0)],
               dict(psf_type="GAUSSIAN", fwhm=0.5, truncation=5), dict(supersampling_factor=2), inverse=True)
    # config-3 shape shrunk: EPL_MULTIPOLE_M1M3M4 + SHEAR, noise map, per-pixel exposure map, GAUSSIAN light
    image_case("c3_multipole_noise_map", 32, 0.1, ["EPL_MULTIPOLE_M1M3M4", "SHEAR"],
               [dict(theta_E=1.0, gamma=1.95, e1=-0.15, e2=0.1, center_x=0.0, center_y=0.03, a1_a=0.02,
                     delta_phi_m1=0.4, a3_a=0.01, delta_phi_m3=0.1, a4_a=-0.015, delta_phi_m4=0.05), shear],
               ["SERSIC_ELLIPSE", "GAUSSIAN"],
               [sersic_src, dict(amp=0.5, sigma=0.15, center_x=-0.1, center_y=0.0)],
               ["SERSIC_ELLIPSE"], [sersic_ll],
               dict(psf_type="PIXEL", kernel_point_source=gauss_kernel(5, 0.9)),
               dict(supersampling_factor=2, convolution_type="grid"), noise_map=True, exposure_map=True)
    # degenerate linear system: two identical basis functions -> cond gate -> amplitudes 0 (de_lens.py:26-33)
    image_case("linear_degenerate", 24, 0.1, ["SIS"], [dict(theta_E=0.8, center_x=0.0, center_y=0.0)],
               ["SERSIC_ELLIPSE", "SERSIC_ELLIPSE"], [sersic_src, sersic_src], [], [],
               dict(psf_type="NONE"), dict(supersampling_factor=1))
    # ---- supersampling_convolution=True (appended last so that the random stream of the cases above is unchanged)
    kd = sim_util.data_configure_simple(20, 0.08, 100, 0.05, inverse=True)
    pg = PixelGrid(20, 20, kd["transform_pix2angle"], kd["ra_at_xy_0"], kd["dec_at_xy_0"])
    kern_super = gauss_kernel(27, 3.5) + 0.05 * rng.uniform(0, 1, (27, 27)) / 729  # sampled at 3 points per pixel
    for name, psf_kw, num_kw in [
        ("ss_pixel_split", dict(psf_type="PIXEL", kernel_point_source=kern_super, point_source_supersampling_factor=3),
         dict(supersampling_factor=3, supersampling_convolution=True, supersampling_kernel_size=5)),
        ("ss_pixel_full", dict(psf_type="PIXEL", kernel_point_source=kern_super, point_source_supersampling_factor=3),
         dict(supersampling_factor=3, supersampling_convolution=True, supersampling_kernel_size=None,
              convolution_type="grid")),
        ("ss_gaussian", dict(psf_type="GAUSSIAN", fwhm=0.2, truncation=4),
         dict(supersampling_factor=2, supersampling_convolution=True)),
    ]:
        psf = PSF(**psf_kw)
        psf.set_pixel_size(pg.pixel_width)
        num = Numerics(pg, psf, **num_kw)
        ra, dec = num.coordinates_evaluate
        flux = np.exp(-((ra - 0.1) ** 2 + (dec + 0.2) ** 2) / 0.05) + 0.01 * rng.uniform(0, 1, ra.shape)
        out = num.re_size_convolve(flux)
        outu = num.re_size_convolve(flux, unconvolved=True)
        ins = dict(flux=flux, transform_pix2angle=kd["transform_pix2angle"], ra_at_xy_0=kd["ra_at_xy_0"],
                   dec_at_xy_0=kd["dec_at_xy_0"])
        if "kernel_point_source" in psf_kw:
            ins["kernel"] = kern_super
        meta_psf = {k: v for k, v in psf_kw.items() if k != "kernel_point_source"}
        st.add("numerics_%s" % name, ins, dict(image=out, image_unconvolved=outu, ra=ra, dec=dec),
               dict(nx=20, ny=20, kwargs_psf=meta_psf, kwargs_numerics=num_kw))
    image_case("ss_conv_sie_shear", 30, 0.1, ["SIE", "SHEAR"],
               [dict(theta_E=1.0, e1=0.1, e2=-0.05, center_x=0.02, center_y=-0.01), shear],
               ["SERSIC_ELLIPSE"], [sersic_src], ["SERSIC"],
               [dict(amp=31.0, R_sersic=0.1, n_sersic=2.0, center_x=0.0, center_y=0.0)],
               dict(psf_type="PIXEL", kernel_point_source=kern_super, point_source_supersampling_factor=3),
               dict(supersampling_factor=3, supersampling_convolution=True))
    # ---- flux_evaluate_indexes: sub-frame evaluation + zero padding (numerics_subframe.py), appended last
    from jaxtronomy.ImSim.Numerics.numerics_subframe import NumericsSubFrame

    yy, xx = np.mgrid[:20, :20]
    disk = ((xx - 9.5) ** 2 + (yy - 9.5) ** 2 < 7.2**2)  # symmetric: square sub-frame [3:17, 3:17]
    for name, psf_kw, num_kw in [
        ("subframe_pixel", dict(psf_type="PIXEL", kernel_point_source=kern), dict(supersampling_factor=2)),
        ("subframe_gaussian", dict(psf_type="GAUSSIAN", fwhm=0.2, truncation=5), dict(supersampling_factor=1)),
    ]:
        psf = PSF(**psf_kw)
        psf.set_pixel_size(pg.pixel_width)
        num = NumericsSubFrame(pg, psf, flux_evaluate_indexes=disk, **num_kw)
        ra, dec = num.coordinates_evaluate
        flux = np.exp(-((ra - 0.1) ** 2 + (dec + 0.2) ** 2) / 0.05) + 0.01 * rng.uniform(0, 1, ra.shape)
        out = np.asarray(num.re_size_convolve(flux))
        ins = dict(flux=flux, transform_pix2angle=kd["transform_pix2angle"], ra_at_xy_0=kd["ra_at_xy_0"],
                   dec_at_xy_0=kd["dec_at_xy_0"], flux_evaluate_indexes=disk)
        if "kernel_point_source" in psf_kw:
            ins["kernel"] = psf_kw["kernel_point_source"]
        st.add("numerics_%s" % name, ins,
               dict(image=out, image_unconvolved=np.asarray(num.re_size_convolve(flux, unconvolved=True)), ra=ra, dec=dec),
               dict(nx=20, ny=20, kwargs_psf={k: v for k, v in psf_kw.items() if k != "kernel_point_source"},
                    kwargs_numerics=num_kw, subframe=True))
    yy, xx = np.mgrid[:36, :36]
    disk36 = ((xx - 17.5) ** 2 + (yy - 17.5) ** 2 < 13.0**2)
    lmask = ((xx - 17.5) ** 2 + (yy - 17.5) ** 2 < 15.5**2).astype(float)  # likelihood mask larger than the sub-frame
    image_case("subframe_epl_shear", 36, 0.09, ["EPL", "SHEAR"],
               [dict(theta_E=1.1, gamma=2.05, e1=0.2, e2=-0.1, center_x=0.02, center_y=-0.01), shear],
               ["SERSIC_ELLIPSE"], [sersic_src], ["SERSIC_ELLIPSE"], [sersic_ll],
               dict(psf_type="PIXEL", kernel_point_source=gauss_kernel(7, 1.2)),
               dict(supersampling_factor=2, flux_evaluate_indexes=disk36), mask=lmask)
    st.save("imsim.npz")


def gen_point_sources():
    """Point sources: PointSourceRendering, ImageModel.point_source / image and ImageLinearFit with
    UNLENSED + LENSED_POSITION models, after test/test_ImSim/test_Numerics/test_point_source_rendering.py and
    test/test_ImSim/test_image_model.py (point_source_type_list=["LENSED_POSITION"], 4 images)."""
    from jaxtronomy.ImSim.Numerics.point_source_rendering import PointSourceRendering
    from jaxtronomy.PointSource.point_source import PointSource

    st = Store()
    rng = np.random.default_rng(23)
    # 1. rendering alone: supersampling 1 and 3, positions inside, on the edge and outside the frame
    numPix, deltaPix = 24, 0.07
    kd = sim_util.data_configure_simple(numPix, deltaPix, 100.0, 0.05, inverse=False)
    pg = PixelGrid(numPix, numPix, kd["transform_pix2angle"], kd["ra_at_xy_0"], kd["dec_at_xy_0"])
    ra = np.array([0.11, -0.4321, 0.77, -0.83, 1.2])  # the last one lies outside the frame
    dec = np.array([-0.2, 0.3, 0.79, -0.8299, 0.1])
    amp = np.array([10.0, 3.0, 7.5, 1.0, 4.0])
    for ss in (1, 3):
        K = 7 * ss if (7 * ss) % 2 == 1 else 7 * ss + 1
        kern = gauss_kernel(K, 1.1 * ss) + 0.02 * rng.uniform(0, 1, (K, K)) / K**2
        psf = PSF(psf_type="PIXEL", kernel_point_source=kern, point_source_supersampling_factor=ss)
        psr = PointSourceRendering(pg, supersampling_factor=ss, psf=psf)
        out = np.asarray(psr.point_source_rendering(ra, dec, amp))
        outu = np.asarray(psr.point_source_rendering(ra, dec, amp, unconvolved=True))
        st.add("rendering_ss%d" % ss, dict(ra=ra, dec=dec, amp=amp, kernel=kern, transform_pix2angle=kd["transform_pix2angle"],
                                         ra_at_xy_0=kd["ra_at_xy_0"], dec_at_xy_0=kd["dec_at_xy_0"]),
               dict(image=out, image_unconvolved=outu), dict(numPix=numPix, supersampling=ss))

    # 2. ImageModel / ImageLinearFit with a lensed quad and a field star
    numPix, deltaPix = 40, 0.08
    kern = gauss_kernel(9, 1.4) + 0.02 * rng.uniform(0, 1, (9, 9)) / 81
    kd = sim_util.data_configure_simple(numPix, deltaPix, 100.0, 0.05, inverse=False)
    kw_lens = [dict(theta_E=1.1, gamma=2.0, e1=0.1, e2=-0.05, center_x=0.02, center_y=-0.01),
               dict(gamma1=0.03, gamma2=-0.01, ra_0=0, dec_0=0)]
    kw_src = [dict(amp=12.0, R_sersic=0.2, n_sersic=1.5, e1=0.1, e2=0.05, center_x=0.08, center_y=0.03)]
    kw_ll = [dict(amp=20.0, R_sersic=0.5, n_sersic=3.0, e1=0.05, e2=0.0, center_x=0.02, center_y=-0.01)]
    kw_ps = [dict(ra_image=np.array([1.013, -0.734, 0.412, -0.551]), dec_image=np.array([0.402, -0.846, -1.002, 0.978]),
                  point_amp=np.array([30.0, 25.0, 18.0, 12.0])),
             dict(ra_image=np.array([-1.21]), dec_image=np.array([1.07]), point_amp=np.array([40.0]))]
    ps_types = ["LENSED_POSITION", "UNLENSED"]
    psf_kw = dict(psf_type="PIXEL", kernel_point_source=kern)
    num_kw = dict(supersampling_factor=2, point_source_supersampling_factor=1)
    lens, src, ll = LensModel(["EPL", "SHEAR"]), LightModel(["SERSIC_ELLIPSE"]), LightModel(["SERSIC_ELLIPSE"])
    cp = lambda kws: [{k: (np.array(v) if isinstance(v, np.ndarray) else v) for k, v in kw.items()} for kw in kws]  # noqa: E731
    psf = PSF(**psf_kw)
    ps = PointSource(ps_types, lens_model=lens)
    data = ImageData(**kd)
    im = ImageModel(data, psf, lens, src, ll, point_source_class=ps, kwargs_numerics=num_kw)
    truth = np.asarray(im.image(kw_lens, kw_src, kw_ll, cp(kw_ps)))
    noisy = truth + rng.normal(0, 0.05, truth.shape) + rng.normal(0, 1, truth.shape) * np.sqrt(np.maximum(truth, 0) / 100.0)
    kd2 = dict(kd, image_data=noisy)
    data = ImageData(**kd2)
    im = ImageModel(data, psf, lens, src, ll, point_source_class=ps, kwargs_numerics=num_kw)
    lin = ImageLinearFit(data, psf, lens, src, ll, point_source_class=ps, kwargs_numerics=num_kw)
    kw_ps_p = cp(kw_ps)
    kw_ps_p[0]["ra_image"] = kw_ps_p[0]["ra_image"] + 0.004  # perturbed: the model differs from the data
    outs = dict(
        image=np.asarray(im.image(kw_lens, kw_src, kw_ll, cp(kw_ps_p))),
        image_unconvolved=np.asarray(im.image(kw_lens, kw_src, kw_ll, cp(kw_ps_p), unconvolved=True)),
        image_no_ps=np.asarray(im.image(kw_lens, kw_src, kw_ll, cp(kw_ps_p), point_source_add=False)),
        point_source=np.asarray(im.point_source(cp(kw_ps_p), kw_lens)),
        point_source_k1=np.asarray(im.point_source(cp(kw_ps_p), kw_lens, k=1)),
        logL=float(im.likelihood_data_given_model(kw_lens, kw_src, kw_ll, cp(kw_ps_p))),
        A=np.asarray(lin.linear_response_matrix(kw_lens, cp(kw_src), cp(kw_ll), cp(kw_ps_p))))
    model, _, _, param = lin.image_linear_solve(kw_lens, cp(kw_src), cp(kw_ll), cp(kw_ps_p))
    logL_lin, param2 = lin.likelihood_data_given_model(kw_lens, cp(kw_src), cp(kw_ll), cp(kw_ps_p))
    outs.update(wls_model=np.asarray(model), param=np.asarray(param), logL_linear=float(logL_lin))
    ins = dict(image_data=noisy, transform_pix2angle=kd["transform_pix2angle"], ra_at_xy_0=kd["ra_at_xy_0"],
               dec_at_xy_0=kd["dec_at_xy_0"], exposure_time=kd["exposure_time"], background_rms=0.05, kernel=kern)
    st.add("image_quad_and_star", ins, outs,
           dict(numPix=numPix, deltaPix=deltaPix, lens_model_list=["EPL", "SHEAR"], kwargs_lens=kw_lens,
                source_model_list=["SERSIC_ELLIPSE"], kwargs_source=kw_src, lens_light_model_list=["SERSIC_ELLIPSE"],
                kwargs_lens_light=kw_ll, point_source_type_list=ps_types, kwargs_ps=kw_ps_p,
                kwargs_psf=dict(psf_type="PIXEL"), kwargs_numerics=num_kw))
    # 3. the same model with a PSF variance map: error map, likelihood and linear solve with C_D + model error
    var_map = 0.02 * kern**2 / np.max(kern**2) + 1e-4 * rng.uniform(0, 1, kern.shape)
    psf_v = PSF(psf_type="PIXEL", kernel_point_source=kern, psf_variance_map=var_map)
    for case, flags in (("error_map_all", None), ("error_map_first_only", [True, False])):
        im = ImageModel(data, psf_v, lens, src, ll, point_source_class=ps, kwargs_numerics=num_kw,
                        psf_error_map_bool_list=flags)
        lin = ImageLinearFit(data, psf_v, lens, src, ll, point_source_class=ps, kwargs_numerics=num_kw,
                             psf_error_map_bool_list=flags)
        outs = dict(error_map=np.asarray(im._error_map_psf(kw_lens, cp(kw_ps_p))),
                    logL=float(im.likelihood_data_given_model(kw_lens, kw_src, kw_ll, cp(kw_ps_p))))
        model, model_error, _, param = lin.image_linear_solve(kw_lens, cp(kw_src), cp(kw_ll), cp(kw_ps_p))
        logL_lin, _ = lin.likelihood_data_given_model(kw_lens, cp(kw_src), cp(kw_ll), cp(kw_ps_p))
        outs.update(wls_model=np.asarray(model), model_error=np.asarray(model_error), param=np.asarray(param),
                    logL_linear=float(logL_lin))
        st.add(case, dict(ins, psf_variance_map=var_map), outs,
               dict(numPix=numPix, deltaPix=deltaPix, lens_model_list=["EPL", "SHEAR"], kwargs_lens=kw_lens,
                    source_model_list=["SERSIC_ELLIPSE"], kwargs_source=kw_src, lens_light_model_list=["SERSIC_ELLIPSE"],
                    kwargs_lens_light=kw_ll, point_source_type_list=ps_types, kwargs_ps=kw_ps_p,
                    kwargs_psf=dict(psf_type="PIXEL"), kwargs_numerics=num_kw, psf_error_map_bool_list=flags))
    st.save("point_sources.npz")


# ------------------------------------------------------------------------------------------------
def hessian_cases():
    """(case, profile, kwargs, profile_kwargs) of the Hessian fixtures: the deflection fixtures' parameter sets."""
    cases = []
    for n, (emod, ang, gamma) in enumerate([(0.0, 0.0, 2.0), (0.05, 0.3, 1.9), (0.12, 1.0, 2.0), (0.25, 0.7, 2.1),
                                            (0.399, 2.5, 2.5), (0.45, -0.4, 1.7), (0.65, 1.9, 2.2), (0.8, 0.1, 2.0)]):
        e1, e2 = e_from(emod, ang)
        cases.append(("epl_%02d" % n, "EPL", dict(theta_E=1.3, gamma=gamma, e1=e1, e2=e2, center_x=0.1, center_y=-0.2), {}))
    for i, (e1, e2) in enumerate([(0.1, -0.05), (-0.4, 0.3), (0.7, 0.7)]):
        cases.append(("sie_%d" % i, "SIE", dict(theta_E=1.1, e1=e1, e2=e2, center_x=-0.1, center_y=0.3), {}))
        cases.append(("nie_%d" % i, "NIE", dict(theta_E=0.9, e1=e1, e2=e2, s_scale=0.1 + 0.2 * i, center_x=0.2,
                                               center_y=0.0), {}))
    cases.append(("sie_epl", "SIE", dict(theta_E=1.1, e1=0.2, e2=0.1, center_x=0.0, center_y=0.0), dict(NIE=False, gamma=2)))
    cases.append(("sis_0", "SIS", dict(theta_E=0.7, center_x=0.3, center_y=-0.6), {}))
    cases.append(("shear_0", "SHEAR", dict(gamma1=-0.2, gamma2=0.4, ra_0=0.5, dec_0=-1.0), {}))
    cases.append(("convergence_0", "CONVERGENCE", dict(kappa=0.12, ra_0=0.1, dec_0=0.2), {}))
    cases.append(("nfw_0", "NFW", dict(Rs=2.4, alpha_Rs=1.5, center_x=0.0, center_y=0.0), {}))
    cases.append(("nfw_1", "NFW", dict(Rs=0.3, alpha_Rs=0.7, center_x=1.0, center_y=2.0), {}))
    cases.append(("tnfw_0", "TNFW", dict(Rs=1.4, alpha_Rs=3.5, r_trunc=3.2, center_x=0.0, center_y=0.0), {}))
    cases.append(("tnfw_1", "TNFW", dict(Rs=2.4, alpha_Rs=1.5, r_trunc=1.2, center_x=1.0, center_y=2.0), {}))
    for m in (1, 2, 3, 4):
        cases.append(("multipole_m%d" % m, "MULTIPOLE",
                      dict(m=m, a_m=0.05, phi_m=0.3 * m, center_x=0.1, center_y=-0.1, r_E=1.2), {}))
    for i, (emod, ang) in enumerate([(0.2, 0.4), (0.0, 0.0), (0.45, -1.0)]):
        e1, e2 = e_from(emod, ang)
        cases.append(("epl_m1m3m4_%d" % i, "EPL_MULTIPOLE_M1M3M4",
                      dict(theta_E=1.2, gamma=2.05, e1=e1, e2=e2, center_x=0.05, center_y=0.02, a1_a=0.02,
                           delta_phi_m1=0.5, a3_a=-0.01, delta_phi_m3=0.2, a4_a=0.015, delta_phi_m4=-0.1), {}))
    for axis in ("product_avg", "major"):
        cases.append(("cse_%s" % axis, "CSE", dict(a=1.3, s=0.4, e1=0.15, e2=-0.2, center_x=0.1, center_y=0.0),
                      dict(axis=axis)))
    for hi in (True, False):
        cases.append(("nfw_ellipse_cse_%d" % hi, "NFW_ELLIPSE_CSE",
                      dict(Rs=1.7, alpha_Rs=1.1, e1=0.12, e2=0.08, center_x=-0.05, center_y=0.1), dict(high_accuracy=hi)))
    cases.append(("hernquist_ellipse_cse", "HERNQUIST_ELLIPSE_CSE",
                  dict(sigma0=0.9, Rs=1.2, e1=-0.1, e2=0.2, center_x=0.0, center_y=0.05), {}))
    return cases


def gen_hessians():
    """LensModel.hessian / kappa / gamma / magnification (lens_model.py:384-527) of the unmodified reference."""
    st = Store()
    rng = np.random.default_rng(11)
    x, y = points(rng, n=61)
    skipped = []
    for case, name, kw, pk in hessian_cases():
        try:
            lm = LensModel([name], profile_kwargs_list=[pk])
            h = lm.hessian(x, y, [kw])
        except Exception as e:  # a profile whose reference hessian needs something the shim lacks
            skipped.append((case, repr(e)))
            continue
        st.add(case, dict(x=x, y=y, **kw), dict(zip(("f_xx", "f_xy", "f_yx", "f_yy"), (np.asarray(a) for a in h))),
               dict(profile=name, profile_kwargs=pk))
    names = ["EPL", "SHEAR", "SIS", "NFW"]
    kw = [dict(theta_E=1.2, gamma=2.1, e1=0.2, e2=-0.1, center_x=0.05, center_y=-0.02),
          dict(gamma1=0.03, gamma2=-0.02, ra_0=0.0, dec_0=0.0),
          dict(theta_E=0.2, center_x=0.8, center_y=0.5),
          dict(Rs=0.5, alpha_Rs=0.1, center_x=-0.7, center_y=0.3)]
    lm = LensModel(names)
    for k in (None, 0, (0, 2), [True, True, False, True]):
        h = lm.hessian(x, y, kw, k=k)
        g1, g2 = lm.gamma(x, y, kw, k=k)
        outs = dict(zip(("f_xx", "f_xy", "f_yx", "f_yy"), (np.asarray(a) for a in h)))
        outs.update(kappa=np.asarray(lm.kappa(x, y, kw, k=k)), gamma1=np.asarray(g1), gamma2=np.asarray(g2),
                    magnification=np.asarray(lm.magnification(x, y, kw, k=k)), curl=np.asarray(lm.curl(x, y, kw, k=k)))
        st.add("model_k%s" % json.dumps(k).replace(" ", ""), dict(x=x, y=y), outs,
               dict(lens_model_list=names, kwargs=kw, k=k))
    for method in ("square", "cross"):
        h = lm.hessian(x, y, kw, diff=1e-4, diff_method=method)
        st.add("model_diff_%s" % method, dict(x=x, y=y),
               dict(zip(("f_xx", "f_xy", "f_yx", "f_yy"), (np.asarray(a) for a in h))),
               dict(lens_model_list=names, kwargs=kw, k=None, diff=1e-4, diff_method=method))
    for hd in (False, True):
        f = lm.flexion(x, y, kw, diff=1e-4, hessian_diff=hd)
        st.add("model_flexion_%d" % hd, dict(x=x, y=y),
               dict(zip(("f_xxx", "f_xxy", "f_xyy", "f_yyy"), (np.asarray(a) for a in f))),
               dict(lens_model_list=names, kwargs=kw, diff=1e-4, hessian_diff=hd))
    # multi-plane (same system as lens_models.npz "multi_plane"): alpha = theta - beta (multi_plane.py:512-534), the
    # Hessian as the reference's forward difference of alpha with diff = 1e-8 (multi_plane.py:537-589), and the
    # square stencil with diff = 1e-4 through LensModel.hessian
    mp_names = ["EPL", "NFW_ELLIPSE_CSE", "SIS", "SHEAR"]
    mp_z = [0.5, 0.5, 0.2, 0.8]
    mp_kw = [kw[0], dict(Rs=1.5, alpha_Rs=0.3, e1=0.1, e2=0.05, center_x=0.0, center_y=0.1),
             dict(theta_E=0.15, center_x=0.6, center_y=-0.4), kw[1]]
    lmm = LensModel(mp_names, z_source=1.5, lens_redshift_list=mp_z, multi_plane=True)
    mp = lmm.lens_model
    base = mp.multi_plane_base
    ax, ay = lmm.alpha(x, y, mp_kw)
    h8 = lmm.hessian(x, y, mp_kw)
    h4 = lmm.hessian(x, y, mp_kw, diff=1e-4, diff_method="square")
    outs = dict(alpha_x=np.asarray(ax), alpha_y=np.asarray(ay), magnification=np.asarray(lmm.magnification(x, y, mp_kw)))
    outs.update({"d8_" + q: np.asarray(a) for q, a in zip(("f_xx", "f_xy", "f_yx", "f_yy"), h8)})
    outs.update({"d4_" + q: np.asarray(a) for q, a in zip(("f_xx", "f_xy", "f_yx", "f_yy"), h4)})
    st.add("multi_plane", dict(x=x, y=y, sorted_index=np.asarray(base.sorted_redshift_index),
                               T_ij_list=np.asarray(base.T_ij_list, dtype=float), T_z_list=np.asarray(base.T_z_list, dtype=float),
                               reduced2physical=np.asarray(base._reduced2physical_factor, dtype=float),
                               T_ij_stop=float(mp.T_ij_stop), T_z_source=float(mp._T_z_source)), outs,
           dict(lens_model_list=mp_names, lens_redshift_list=mp_z, z_source=1.5, kwargs=mp_kw))
    st.save("hessians.npz")
    for c in skipped:
        print("skipped", c)


def gen_position_likelihood():
    """PositionLikelihood (Sampling/Likelihoods/position_likelihood.py) of the unmodified reference: source-position
    penalty (with and without the hard bound), image-position likelihood, astrometric likelihood, their sum."""
    from jaxtronomy.PointSource.point_source import PointSource
    from jaxtronomy.Sampling.Likelihoods.position_likelihood import PositionLikelihood

    st = Store()
    kw_epl = dict(theta_E=1.1, gamma=2.0, e1=0.1, e2=-0.05, center_x=0.02, center_y=-0.01)
    kw_shear = dict(gamma1=0.03, gamma2=-0.01, ra_0=0, dec_0=0)
    quad = dict(ra_image=np.array([1.013, -0.734, 0.412, -0.551]), dec_image=np.array([0.402, -0.846, -1.002, 0.978]),
                point_amp=np.array([30.0, 25.0, 18.0, 12.0]))
    star = dict(ra_image=np.array([-1.21]), dec_image=np.array([1.07]), point_amp=np.array([40.0]))
    double = dict(ra_image=np.array([0.93, -0.61]), dec_image=np.array([0.81, -0.52]), point_amp=np.array([9.0, 4.0]))
    cases = [
        ("quad", ["EPL", "SHEAR"], [kw_epl, kw_shear], ["LENSED_POSITION"], [quad], dict()),
        ("quad_hard_bound", ["EPL", "SHEAR"], [kw_epl, kw_shear], ["LENSED_POSITION"], [quad],
         dict(source_position_tolerance=0.05, source_position_sigma=0.003)),
        ("quad_star_double", ["EPL", "SHEAR"], [kw_epl, kw_shear], ["LENSED_POSITION", "UNLENSED", "LENSED_POSITION"],
         [quad, star, double], dict(source_position_sigma=0.002)),
        ("quad_m1m3m4", ["EPL_MULTIPOLE_M1M3M4", "SHEAR", "SIS"],
         [dict(kw_epl, a1_a=0.01, delta_phi_m1=0.3, a3_a=0.005, delta_phi_m3=-0.2, a4_a=0.01, delta_phi_m4=0.1), kw_shear,
          dict(theta_E=0.1, center_x=0.9, center_y=-0.7)], ["LENSED_POSITION"], [quad], dict()),
        ("quad_image_positions", ["EPL", "SHEAR"], [kw_epl, kw_shear], ["LENSED_POSITION"], [quad],
         dict(image_position_likelihood=True, image_position_uncertainty=0.004,
              ra_image_list=[[1.01, -0.73, 0.41, -0.55]], dec_image_list=[[0.40, -0.85, -1.00, 0.98]])),
        # SIS, theta_E = 1, source at (0.1, 0): exact images (1.1, 0) and (-0.9, 0), perturbed by ~1e-3
        ("sis_double_near", ["SIS"], [dict(theta_E=1.0, center_x=0.0, center_y=0.0)], ["LENSED_POSITION"],
         [dict(ra_image=np.array([1.1008, -0.9003]), dec_image=np.array([0.0005, -0.0002]), point_amp=np.array([1.0, 1.0]))],
         dict(source_position_tolerance=0.001)),
        ("sis_double_near_bound_hit", ["SIS"], [dict(theta_E=1.0, center_x=0.0, center_y=0.0)], ["LENSED_POSITION"],
         [dict(ra_image=np.array([1.1008, -0.9003]), dec_image=np.array([0.0005, -0.0002]), point_amp=np.array([1.0, 1.0]))],
         dict(source_position_tolerance=0.0001)),
        ("star_only", ["EPL", "SHEAR"], [kw_epl, kw_shear], ["UNLENSED"], [star], dict()),
    ]
    cp = lambda kws: [{k: (np.array(v) if isinstance(v, np.ndarray) else v) for k, v in kw.items()} for kw in kws]  # noqa: E731
    for case, names, kw_lens, types, kw_ps, opts in cases:
        lens = LensModel(names)
        ps = PointSource(types, lens_model=lens)
        pl = PositionLikelihood(ps, **opts)
        outs = dict(logL=float(pl.logL(kw_lens, cp(kw_ps), {})))
        if "LENSED_POSITION" in types:
            outs["logL_source_position"] = float(pl.source_position_likelihood(
                kw_lens, cp(kw_ps), opts.get("source_position_sigma", 0.001),
                hard_bound_rms=opts.get("source_position_tolerance")))
        sx, sy = ps.source_position(cp(kw_ps), kw_lens)
        for i in range(len(types)):
            outs["x_source_%d" % i], outs["y_source_%d" % i] = np.asarray(sx[i], dtype=float), np.asarray(sy[i], dtype=float)
        st.add(case, dict(), outs, dict(lens_model_list=names, kwargs_lens=kw_lens, point_source_type_list=types,
                                        kwargs_ps=[{k: np.asarray(v).tolist() for k, v in kw.items()} for kw in kw_ps],
                                        options=opts))
    # astrometric likelihood (static): kwargs_special offsets
    special = dict(delta_x_image=[0.002, -0.001, 0.0005, 0.003], delta_y_image=[-0.0015, 0.0, 0.002, -0.0005])
    st.add("astrometric", dict(), dict(logL=float(PositionLikelihood.astrometric_likelihood([quad], special, 0.004))),
           dict(kwargs_special=special, sigma=0.004))
    st.save("position_likelihood.npz")


# ------------------------------------------------------------------------------------------------
def gen_baseline_shapes():
    """Fixtures at the BASELINE.json shapes themselves (not shrunk), from the unmodified reference:
    headline (EPL+SHEAR, SERSIC_ELLIPSE, 180x180, ss 3, 21x21 pixel PSF, ImageLinearFit) for parameter sets in every
    nmax bracket of epl.py:308-319; C3 (EPL_MULTIPOLE_M1M3M4+SHEAR, n_lin 2); C4 (1 EPL + 1,000 TNFW, 10^4 rays of
    the 1000x1000 grid); C5 (multi-plane EPL / NFW_ELLIPSE_CSE / SIS inside MultiLinear over 4 bands).
    The static inputs are the ones bench.py / tools/bench_configs.py use, so the instantiations checked are the ones timed."""
    sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
    from bench import LENS_LIST, SOURCE_LIST, TRUTH, psf_kernel, row_to_kwargs, sample_params, static_config
    from jaxtronomy.ImSim.MultiBand.multi_linear import MultiLinear

    st = Store()
    cp = lambda kws: [dict(k) for k in kws]  # noqa: E731
    sc = static_config()
    psf = PSF(psf_type="PIXEL", kernel_point_source=sc["kernel"])
    num_kw = dict(supersampling_factor=3, supersampling_convolution=False)

    # ---- headline: 10 parameter sets, |e| chosen per nmax bracket (10, 10, 20, 20, 20, 35, 35, 50, 50, 100 terms)
    lens, src, none = LensModel(LENS_LIST), LightModel(SOURCE_LIST), LightModel([])
    im0 = ImageModel(ImageData(image_data=np.zeros((180, 180)), **sc["kwargs_data"]), psf, lens, src, none,
                     kwargs_numerics=num_kw)
    truth = np.asarray(im0.image(*row_to_kwargs(TRUTH)))
    rng = np.random.default_rng(0)
    noisy = truth + rng.normal(0, 0.05, truth.shape) + rng.normal(0, 1, truth.shape) * np.sqrt(np.maximum(truth, 0) / 100.0)
    data = ImageData(image_data=noisy, **sc["kwargs_data"])
    im = ImageModel(data, psf, lens, src, none, kwargs_numerics=num_kw)
    lin = ImageLinearFit(data, psf, lens, src, none, kwargs_numerics=num_kw)
    rows = sample_params(10, seed=2026)
    emods = [0.05, 0.11, 0.2, 0.3, 0.39, 0.45, 0.55, 0.65, 0.72, 0.8]
    ang = np.random.default_rng(5).uniform(0, 2 * np.pi, 10)
    rows[:, 2], rows[:, 3] = emods * np.cos(ang), emods * np.sin(ang)
    st.add("headline_static", dict(image_data=noisy, kernel=sc["kernel"], rows=rows, truth_row=TRUTH), dict(truth=truth),
           dict(lens_model_list=LENS_LIST, source_model_list=SOURCE_LIST, kwargs_numerics=num_kw, e_moduli=emods))
    for b, row in enumerate(rows):
        kl, ks, kll = row_to_kwargs(row)
        outs = dict(image_unconvolved=np.asarray(im.image(kl, ks, kll, unconvolved=True)),
                    A=np.asarray(lin.linear_response_matrix(kl, cp(ks), cp(kll), None)),
                    logL=float(im.likelihood_data_given_model(kl, ks, kll)))
        model, _, _, param = lin.image_linear_solve(kl, cp(ks), cp(kll))
        logL_lin, _ = lin.likelihood_data_given_model(kl, cp(ks), cp(kll))
        outs.update(param=np.asarray(param), logL_linear=float(logL_lin))
        if b in (0, 7):
            outs.update(wls_model=np.asarray(model), image=np.asarray(im.image(kl, ks, kll)))
        st.add("headline_%02d" % b, dict(row=row), outs)

    # ---- C3: EPL_MULTIPOLE_M1M3M4 + SHEAR, SERSIC_ELLIPSE source + lens light (n_lin 2), same frame / PSF
    lens3, ll3 = LensModel(["EPL_MULTIPOLE_M1M3M4", "SHEAR"]), LightModel(["SERSIC_ELLIPSE"])
    kl3 = [dict(theta_E=1.2, gamma=2.0, e1=0.2, e2=-0.1, center_x=0.0, center_y=0.0, a1_a=0.01, delta_phi_m1=0.1,
                a3_a=0.01, delta_phi_m3=0.1, a4_a=0.01, delta_phi_m4=0.1), dict(gamma1=0.03, gamma2=0.01, ra_0=0, dec_0=0)]
    ks3 = [dict(amp=1.0, R_sersic=0.3, n_sersic=2.0, e1=0.1, e2=0.1, center_x=0.05, center_y=0.0)]
    kll3 = [dict(amp=1.0, R_sersic=0.6, n_sersic=3.0, e1=0.05, e2=0.0, center_x=0.0, center_y=0.0)]
    im0 = ImageModel(ImageData(image_data=np.zeros((180, 180)), **sc["kwargs_data"]), psf, lens3, src, ll3,
                     kwargs_numerics=num_kw)
    truth3 = np.asarray(im0.image(kl3, [dict(ks3[0], amp=20.0)], [dict(kll3[0], amp=30.0)]))
    noisy3 = truth3 + rng.normal(0, 0.05, truth3.shape) + rng.normal(0, 1, truth3.shape) * np.sqrt(np.maximum(truth3, 0) / 100.0)
    data3 = ImageData(image_data=noisy3, **sc["kwargs_data"])
    lin3 = ImageLinearFit(data3, psf, lens3, src, ll3, kwargs_numerics=num_kw)
    im3 = ImageModel(data3, psf, lens3, src, ll3, kwargs_numerics=num_kw)
    st.add("c3_static", dict(image_data=noisy3, kernel=sc["kernel"]), dict(),
           dict(lens_model_list=["EPL_MULTIPOLE_M1M3M4", "SHEAR"], source_model_list=["SERSIC_ELLIPSE"],
                lens_light_model_list=["SERSIC_ELLIPSE"], kwargs_numerics=num_kw))
    rng3 = np.random.default_rng(33)
    for b in range(3):
        kl = [dict(kl3[0]), dict(kl3[1])]
        for k, s in (("theta_E", 0.03), ("gamma", 0.05), ("e1", 0.1), ("e2", 0.1), ("a1_a", 0.01), ("a3_a", 0.01),
                     ("a4_a", 0.01), ("delta_phi_m1", 0.3), ("delta_phi_m3", 0.3), ("delta_phi_m4", 0.3),
                     ("center_x", 0.02), ("center_y", 0.02)):
            kl[0][k] = float(kl[0][k] + s * rng3.normal())
        ks = [dict(ks3[0], R_sersic=float(0.3 + 0.05 * rng3.normal()), n_sersic=float(2.0 + 0.3 * rng3.normal()))]
        kll = [dict(kll3[0], R_sersic=float(0.6 + 0.05 * rng3.normal()), e1=float(0.05 + 0.05 * rng3.normal()))]
        model, _, _, param = lin3.image_linear_solve(kl, cp(ks), cp(kll))
        logL_lin, _ = lin3.likelihood_data_given_model(kl, cp(ks), cp(kll))
        st.add("c3_%02d" % b, dict(), dict(A=np.asarray(lin3.linear_response_matrix(kl, cp(ks), cp(kll), None)),
                                          image_unconvolved=np.asarray(im3.image(kl, ks, kll, unconvolved=True)),
                                          param=np.asarray(param), logL_linear=float(logL_lin)),
               dict(kwargs_lens=kl, kwargs_source=ks, kwargs_lens_light=kll))

    # ---- C4: 1 EPL + 1,000 TNFW, every 10th row / column of the 1000 x 1000 ray grid (10^4 rays)
    rng4 = np.random.default_rng(3)
    nh = 1000
    names = ["EPL"] + ["TNFW"] * nh
    kws = [dict(theta_E=1.0, gamma=2.0, e1=0.1, e2=0.1, center_x=0.0, center_y=0.0)]
    kws += [dict(Rs=float(rng4.uniform(0.01, 0.1)), alpha_Rs=float(rng4.uniform(1e-4, 5e-3)),
                 r_trunc=float(rng4.uniform(0.05, 0.5)), center_x=float(rng4.uniform(-2, 2)),
                 center_y=float(rng4.uniform(-2, 2))) for _ in range(nh)]
    g = np.linspace(-2, 2, 1000)[5::10]
    xx, yy = np.meshgrid(g, g)
    lmb = LensModelBulk(["EPL", "TNFW"], multi_plane=False)
    rs_kw = lmb.prepare_ray_shooting_kwargs(names, kws)
    bx, by = lmb.ray_shooting(xx.ravel(), yy.ravel(), rs_kw)
    st.add("c4_substructure", dict(x=xx.ravel(), y=yy.ravel(), index_list=rs_kw["index_list"],
                                   **{"kw_" + k: v for k, v in rs_kw["all_kwargs"].items()}),
           dict(beta_x=np.asarray(bx), beta_y=np.asarray(by)),
           dict(unique_lens_model_list=["EPL", "TNFW"], n_halos=nh, kwargs_epl=kws[0]))

    # ---- C5: multi-plane (EPL z=0.5, NFW_ELLIPSE_CSE z=0.3, SIS z=0.8; source z=1.5) inside MultiLinear, 4 bands
    rng5 = np.random.default_rng(4)
    km = dict(lens_model_list=["EPL", "NFW_ELLIPSE_CSE", "SIS"], lens_redshift_list=[0.5, 0.3, 0.8], z_source=1.5,
              multi_plane=True, source_light_model_list=["SERSIC_ELLIPSE"], lens_light_model_list=["SERSIC_ELLIPSE"])
    kl5 = [dict(theta_E=1.1, gamma=2.0, e1=0.1, e2=-0.1, center_x=0.0, center_y=0.0),
           dict(Rs=1.5, alpha_Rs=0.2, e1=0.05, e2=0.05, center_x=0.3, center_y=-0.2),
           dict(theta_E=0.1, center_x=-0.5, center_y=0.4)]
    ks5 = [dict(amp=1.0, R_sersic=0.3, n_sersic=2.0, e1=0.1, e2=0.1, center_x=0.05, center_y=0.0)]
    kll5 = [dict(amp=1.0, R_sersic=0.6, n_sersic=3.0, e1=0.05, e2=0.0, center_x=0.0, center_y=0.0)]
    band_kernels = [psf_kernel(11, 1.5 + 0.2 * b) for b in range(4)]
    band_num = [dict(supersampling_factor=1), dict(supersampling_factor=1), dict(supersampling_factor=2),
                dict(supersampling_factor=1)]

    def bands_for(images):
        return [[dict(sc["kwargs_data"], image_data=images[b]), dict(psf_type="PIXEL", kernel_point_source=band_kernels[b]),
                 band_num[b]] for b in range(4)]

    ml0 = MultiLinear(bands_for([np.zeros((180, 180))] * 4), km)
    images = []
    for b in range(4):
        t = np.asarray(ml0._imageModel_list[b].image(kl5, [dict(ks5[0], amp=15.0 + 3 * b)], [dict(kll5[0], amp=25.0 - 2 * b)]))
        images.append(t + rng5.normal(0, 0.05, t.shape) + rng5.normal(0, 1, t.shape) * np.sqrt(np.maximum(t, 0) / 100.0))
    ml = MultiLinear(bands_for(images), km)
    base = ml._imageModel_list[0].LensModel.lens_model.multi_plane_base
    mp = ml._imageModel_list[0].LensModel.lens_model
    ins = dict(sorted_index=np.asarray(base.sorted_redshift_index), T_ij_list=np.asarray(base.T_ij_list, dtype=float),
               T_z_list=np.asarray(base.T_z_list, dtype=float),
               reduced2physical=np.asarray(base._reduced2physical_factor, dtype=float),
               T_ij_stop=float(mp.T_ij_stop), T_z_source=float(mp._T_z_source))
    for b in range(4):
        ins["image_data_%d" % b], ins["kernel_%d" % b] = images[b], band_kernels[b]
    st.add("c5_static", ins, dict(), dict(kwargs_model=km, band_numerics=band_num))
    for s in range(3):
        kl = [dict(k) for k in kl5]
        for i, k, sig in ((0, "theta_E", 0.02), (0, "gamma", 0.05), (0, "e1", 0.05), (0, "center_x", 0.01),
                          (1, "alpha_Rs", 0.02), (1, "Rs", 0.1), (1, "e2", 0.03), (2, "theta_E", 0.02), (2, "center_y", 0.05)):
            kl[i][k] = float(kl[i][k] + sig * rng5.normal())
        ks = [dict(ks5[0], R_sersic=float(0.3 + 0.03 * rng5.normal()), center_x=float(0.05 + 0.02 * rng5.normal()))]
        logL, params = ml.likelihood_data_given_model(kl, cp(ks), cp(kll5))
        outs = dict(logL=float(logL))
        for b in range(4):
            outs["param_%d" % b] = np.asarray(params[b])
            lb, _ = ml._imageModel_list[b].likelihood_data_given_model(kl, cp(ks), cp(kll5))
            outs["logL_%d" % b] = float(lb)
        if s == 0:
            for b in range(4):
                outs["A_%d" % b] = np.asarray(ml._imageModel_list[b].linear_response_matrix(kl, cp(ks), cp(kll5), None))
        else:
            outs["A_2"] = np.asarray(ml._imageModel_list[2].linear_response_matrix(kl, cp(ks), cp(kll5), None))
        st.add("c5_%02d" % s, dict(), outs, dict(kwargs_lens=kl, kwargs_source=ks, kwargs_lens_light=kll5))
    st.save("baseline_shapes.npz")

# ------------------------------------------------------------------------------------------------
def gen_gradients():
    """Gradient pins: 4th-order central differences of the UNMODIFIED reference's log-likelihood
    (ImageModel.likelihood_data_given_model and ImageLinearFit.likelihood_data_given_model under refshim) in every
    non-linear parameter, at the steps 1e-4, 1e-5, 1e-6 (x max(1, |value|)), for
      * epl_shear_sersic: EPL + SHEAR, SERSIC_ELLIPSE source and lens light, 40 x 40, ss 2, 9 x 9 pixel PSF;
      * quad_and_star: the same plus a 4-image LENSED_POSITION point source and an UNLENSED star (the class of the
        Einstein-ring notebook's L-BFGS model, image_model.py:395-429, point_source_rendering.py:33-77);
      * multi_plane: BASELINE config 5's lens (EPL z=0.5, NFW_ELLIPSE_CSE z=0.3, SIS z=0.8, source z=1.5) on 32 x 32.
    jax autodiff itself cannot run here (no jax); the finite difference of the reference's own forward value is the
    closest pin available, and it is independent of the oracle.  tests/test_gpu_grad.py compares lsim_logl_grad with
    the closest of the three steps."""
    from jaxtronomy.PointSource.point_source import PointSource

    st = Store()
    steps = (1e-4, 1e-5, 1e-6)
    gk = lambda n, s: gauss_kernel(n, s)  # noqa: E731

    def run_case(case, kd, kern, num_kw, lens, lens_meta, src_names, ll_names, kw0, ps_types, truth_amp, seed, extra_in):
        rng = np.random.default_rng(seed)
        src, ll = LightModel(src_names), LightModel(ll_names)
        psf = PSF(psf_type="PIXEL", kernel_point_source=kern)
        ps = PointSource(ps_types, lens_model=lens) if ps_types else None
        n = kd["transform_pix2angle"].shape[0] and int(kd["numPix"])
        kdd = {k: v for k, v in kd.items() if k != "numPix"}
        cp = lambda kws: [{k: (np.array(v) if isinstance(v, np.ndarray) else v) for k, v in kw.items()} for kw in kws]  # noqa: E731
        kl, ks, kll, kps = kw0
        mk = lambda cls, image: cls(ImageData(image_data=image, **kdd), psf, lens, src, ll, point_source_class=ps,  # noqa: E731
                                    kwargs_numerics=num_kw)
        ks_t = [dict(k, amp=a) for k, a in zip(ks, truth_amp["source"])]
        kll_t = [dict(k, amp=a) for k, a in zip(kll, truth_amp["lens_light"])]
        truth = np.asarray(mk(ImageModel, np.zeros((n, n))).image(cp(kl), ks_t, kll_t, cp(kps) if kps else None))
        noisy = truth + rng.normal(0, 0.05, truth.shape) + rng.normal(0, 1, truth.shape) * np.sqrt(np.maximum(truth, 0) / 100.0)
        im, lin = mk(ImageModel, noisy), mk(ImageLinearFit, noisy)
        # the parameter vector, in the order lens / source / lens_light / point_source; the evaluation point is a few
        # per cent off the truth so that no gradient component vanishes
        names, row = [], []
        for comp, kws in (("lens", kl), ("source", ks_t), ("lens_light", kll_t)):
            for i, kw in enumerate(kws):
                for k, v in kw.items():
                    if comp == "lens" and k in ("ra_0", "dec_0"):
                        continue
                    names.append([comp, i, k])
                    row.append(float(v))
        for i, kw in enumerate(kps or []):
            for k in ("ra_image", "dec_image", "point_amp"):
                for e, v in enumerate(np.atleast_1d(kw[k])):
                    names.append(["point_source", i, "%s[%d]" % (k, e)])
                    row.append(float(v))
        row = np.array(row)
        for j, (comp, i, k) in enumerate(names):
            if k in ("theta_E", "gamma", "R_sersic", "n_sersic", "Rs", "alpha_Rs"):
                row[j] *= 1 + 0.02 * rng.normal()
            elif k in ("e1", "e2", "gamma1", "gamma2", "center_x", "center_y") or k.startswith(("ra_image", "dec_image")):
                row[j] += 0.004 * rng.normal()

        def kwargs_of(r):
            out = {"lens": [dict() for _ in kl], "source": [dict() for _ in ks], "lens_light": [dict() for _ in kll]}
            for i, kw in enumerate(kl):
                for k in ("ra_0", "dec_0"):
                    if k in kw:
                        out["lens"][i][k] = kw[k]
            pso = [{k: np.zeros(len(np.atleast_1d(kw[k]))) for k in ("ra_image", "dec_image", "point_amp")} for kw in (kps or [])]
            for j, (comp, i, k) in enumerate(names):
                if comp == "point_source":
                    base, e = k.split("[")
                    pso[i][base][int(e[:-1])] = r[j]
                else:
                    out[comp][i][k] = float(r[j])
            return out["lens"], out["source"], out["lens_light"], (pso if kps else None)

        f_non = lambda r: float(im.likelihood_data_given_model(*kwargs_of(r)))  # noqa: E731
        f_lin = lambda r: float(lin.likelihood_data_given_model(*kwargs_of(r))[0])  # noqa: E731

        def fd(fun, j, h):
            def at(s):
                r = row.copy()
                r[j] += s
                return fun(r)
            return (8.0 * (at(h) - at(-h)) - (at(2 * h) - at(-2 * h))) / (12.0 * h)

        P = len(row)
        fd_non, fd_lin = np.full((len(steps), P), np.nan), np.full((len(steps), P), np.nan)
        for j, (comp, i, k) in enumerate(names):
            linear_col = k == "amp" or k.startswith("point_amp")
            for a, h0 in enumerate(steps):
                h = h0 * max(1.0, abs(row[j]))
                fd_non[a, j] = fd(f_non, j, h)
                if not linear_col:
                    fd_lin[a, j] = fd(f_lin, j, h)
        ins = dict(image_data=noisy, kernel=kern, row=row, steps=np.array(steps), **kdd, **extra_in)
        st.add(case, ins, dict(fd_nonlinear=fd_non, fd_linear=fd_lin, logL=f_non(row), logL_linear=f_lin(row)),
               dict(names=names, lens_meta=lens_meta, source_model_list=src_names, lens_light_model_list=ll_names,
                    point_source_type_list=ps_types, num_point_source_images=[len(np.atleast_1d(kw["ra_image"])) for kw in (kps or [])],
                    kwargs_numerics=num_kw, fixed_lens=[{k: kw[k] for k in ("ra_0", "dec_0") if k in kw} for kw in kl]))
        print(case, "P =", P, "logL", f_non(row), f_lin(row))

    numPix, deltaPix = 40, 0.08
    kd = sim_util.data_configure_simple(numPix, deltaPix, 100.0, 0.05, inverse=False)
    kd = dict(numPix=numPix, transform_pix2angle=kd["transform_pix2angle"], ra_at_xy_0=kd["ra_at_xy_0"],
              dec_at_xy_0=kd["dec_at_xy_0"], exposure_time=100.0, background_rms=0.05)
    rngk = np.random.default_rng(31)
    kern = gk(9, 1.4) + 0.02 * rngk.uniform(0, 1, (9, 9)) / 81
    kl = [dict(theta_E=1.1, gamma=2.0, e1=0.1, e2=-0.05, center_x=0.02, center_y=-0.01),
          dict(gamma1=0.03, gamma2=-0.01, ra_0=0, dec_0=0)]
    ks = [dict(amp=12.0, R_sersic=0.2, n_sersic=1.5, e1=0.1, e2=0.05, center_x=0.08, center_y=0.03)]
    kll = [dict(amp=20.0, R_sersic=0.5, n_sersic=3.0, e1=0.05, e2=0.0, center_x=0.02, center_y=-0.01)]
    kps = [dict(ra_image=np.array([1.013, -0.734, 0.412, -0.551]), dec_image=np.array([0.402, -0.846, -1.002, 0.978]),
                point_amp=np.array([30.0, 25.0, 18.0, 12.0])),
           dict(ra_image=np.array([-1.21]), dec_image=np.array([1.07]), point_amp=np.array([40.0]))]
    amps = dict(source=[12.0], lens_light=[20.0])
    num_kw = dict(supersampling_factor=2, point_source_supersampling_factor=1)
    lens = LensModel(["EPL", "SHEAR"])
    run_case("epl_shear_sersic", kd, kern, dict(supersampling_factor=2), lens, dict(lens_model_list=["EPL", "SHEAR"]),
             ["SERSIC_ELLIPSE"], ["SERSIC_ELLIPSE"], (kl, ks, kll, None), [], amps, 41, {})
    run_case("quad_and_star", kd, kern, num_kw, lens, dict(lens_model_list=["EPL", "SHEAR"]), ["SERSIC_ELLIPSE"],
             ["SERSIC_ELLIPSE"], (kl, ks, kll, kps), ["LENSED_POSITION", "UNLENSED"], amps, 42, {})

    n, dp = 32, 0.12
    half = dp * (n - 1) / 2.0
    kd = dict(numPix=n, transform_pix2angle=np.array([[dp, 0.0], [0.0, dp]]), ra_at_xy_0=-half, dec_at_xy_0=-half,
              exposure_time=100.0, background_rms=0.05)
    km = dict(lens_model_list=["EPL", "NFW_ELLIPSE_CSE", "SIS"], lens_redshift_list=[0.5, 0.3, 0.8], z_source=1.5,
              multi_plane=True)
    lmm = LensModel(km["lens_model_list"], z_source=km["z_source"], lens_redshift_list=km["lens_redshift_list"],
                    multi_plane=True)
    mp = lmm.lens_model
    base = mp.multi_plane_base
    consts = dict(sorted_index=np.asarray(base.sorted_redshift_index), T_ij_list=np.asarray(base.T_ij_list, dtype=float),
                  T_z_list=np.asarray(base.T_z_list, dtype=float),
                  reduced2physical=np.asarray(base._reduced2physical_factor, dtype=float),
                  T_ij_stop=float(mp.T_ij_stop), T_z_source=float(mp._T_z_source))
    kl5 = [dict(theta_E=1.1, gamma=2.0, e1=0.1, e2=-0.1, center_x=0.0, center_y=0.0),
           dict(Rs=1.5, alpha_Rs=0.2, e1=0.05, e2=0.05, center_x=0.3, center_y=-0.2),
           dict(theta_E=0.1, center_x=-0.5, center_y=0.4)]
    ks5 = [dict(amp=15.0, R_sersic=0.3, n_sersic=2.0, e1=0.1, e2=0.1, center_x=0.05, center_y=0.0)]
    kll5 = [dict(amp=25.0, R_sersic=0.6, n_sersic=3.0, e1=0.05, e2=0.0, center_x=0.0, center_y=0.0)]
    run_case("multi_plane", kd, gk(11, 1.5), dict(supersampling_factor=2), lmm, km, ["SERSIC_ELLIPSE"], ["SERSIC_ELLIPSE"],
             (kl5, ks5, kll5, None), [], dict(source=[15.0], lens_light=[25.0]), 43, consts)
    st.save("gradients.npz")


if __name__ == "__main__":
    if sys.argv[1:] == ["gradients"]:
        gen_gradients()
        sys.exit(0)
    if sys.argv[1:] == ["baseline_shapes"]:
        gen_baseline_shapes()
        sys.exit(0)
    if sys.argv[1:] == ["position_likelihood"]:
        gen_position_likelihood()
        sys.exit(0)
    if sys.argv[1:] == ["hessians"]:
        gen_hessians()
        sys.exit(0)
    if sys.argv[1:] == ["point_sources"]:
        gen_point_sources()
        sys.exit(0)
    gen_profiles()
    gen_lens_models()
    gen_imsim()
    gen_point_sources()
    gen_hessians()
    gen_position_likelihood()
    gen_baseline_shapes()
    gen_gradients()
