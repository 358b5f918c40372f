"""CPU: the NumPy oracle against tests/golden/gradients.npz (the unmodified reference's log-likelihood and its central
differences, make_golden.py gen_gradients).  The value must agree at the 1e-10 bar; one central difference of the
oracle per case must reproduce the reference's (so the oracle-based gradient pins in test_gpu_grad.py and the
reference-based ones measure the same function)."""

import numpy as np
import pytest

from oracle import imsim
from oracle.lenstronomy_port import PSF
from tests.golden_io import load

GRADS = load("gradients.npz")


def _oracle_model(c):
    kd = dict(image_data=c.inp["image_data"], transform_pix2angle=c.inp["transform_pix2angle"],
              ra_at_xy_0=float(c.inp["ra_at_xy_0"]), dec_at_xy_0=float(c.inp["dec_at_xy_0"]),
              background_rms=float(c.inp["background_rms"]), exposure_time=float(c.inp["exposure_time"]))
    lm = c.meta["lens_meta"]
    return imsim.ImageModel(imsim.ImageData(**kd), PSF(psf_type="PIXEL", kernel_point_source=c.inp["kernel"]),
                            lm["lens_model_list"], c.meta["source_model_list"], c.meta["lens_light_model_list"],
                            c.meta["kwargs_numerics"], point_source_type_list=c.meta["point_source_type_list"] or None,
                            multi_plane=c.inp if lm.get("multi_plane") else None)


def _kwargs(c, r):
    names = [tuple(n) for n in c.meta["names"]]
    n_of = lambda comp: 1 + max(i for cc, i, _ in names if cc == comp)  # noqa: E731
    out = {comp: [dict() for _ in range(n_of(comp))] for comp in ("lens", "source", "lens_light")}
    for i, kw in enumerate(c.meta["fixed_lens"]):
        out["lens"][i].update(kw)
    ps = [{k: np.zeros(n) for k in ("ra_image", "dec_image", "point_amp")} for n in c.meta["num_point_source_images"]]
    for j, (comp, i, k) in enumerate(names):
        if comp == "point_source":
            base, e = k.split("[")
            ps[i][base][int(e[:-1])] = r[j]
        else:
            out[comp][i][k] = float(r[j])
    return out["lens"], out["source"], out["lens_light"], (ps or None)


@pytest.mark.parametrize("name", sorted(GRADS))
def test_oracle_value_and_one_difference_match_the_reference(name):
    c = GRADS[name]
    om = _oracle_model(c)
    row = c.inp["row"]

    def f_non(r):
        kl, ks, kll, kps = _kwargs(c, r)
        return om.likelihood_data_given_model(kl, ks, kll, kwargs_ps=kps)

    def f_lin(r):
        kl, ks, kll, kps = _kwargs(c, r)
        return om.likelihood_data_given_model_linear(kl, ks, kll, kwargs_ps=kps)[0]

    assert abs(f_non(row) - float(c.out["logL"])) <= 1e-10 * abs(float(c.out["logL"]))
    assert abs(f_lin(row) - float(c.out["logL_linear"])) <= 1e-10 * abs(float(c.out["logL_linear"]))
    names = [tuple(n) for n in c.meta["names"]]
    j = [k for _, _, k in names].index("theta_E")
    h = 1e-5 * max(1.0, abs(row[j]))

    def fd(fun):
        def at(s):
            r = row.copy()
            r[j] += s
            return fun(r)
        return (8.0 * (at(h) - at(-h)) - (at(2 * h) - at(-2 * h))) / (12.0 * h)

    for fun, key in ((f_non, "fd_nonlinear"), (f_lin, "fd_linear")):
        ref = c.out[key][1, j]
        assert abs(fd(fun) - ref) <= 1e-6 * abs(ref), (key, fd(fun), ref)
