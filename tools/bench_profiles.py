#!/usr/bin/env python
"""Per-profile throughput on one B200, after the reference README's per-profile table (README.rst:100-244): every
lens profile of SURVEY.md 8(a) through `LensModel.alpha` and every light profile through
`LightModel.surface_brightness`, 180 x 180 supersampling 3 grid (291,600 points) x 256 parameter sets per call, with
the NumPy port (oracle/, one core, one parameter set) timed beside it.  Feeds profiles/r1_profiles.md.

    python tools/bench_profiles.py [--no-cpu]
"""

import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from bench import static_config  # noqa: E402
from jaxtronomy_b200.ImSim.Numerics.grid import RegularGrid  # noqa: E402
from jaxtronomy_b200.LensModel.lens_model import LensModel  # noqa: E402
from jaxtronomy_b200.LightModel.light_model import LightModel  # noqa: E402

B = 256
LENS = {
    "EPL": dict(theta_E=1.2, gamma=2.1, e1=0.2, e2=-0.1, center_x=0.01, center_y=-0.02),
    "SIE": dict(theta_E=1.2, e1=0.2, e2=-0.1, center_x=0.01, center_y=-0.02),
    "NIE": dict(theta_E=1.2, e1=0.2, e2=-0.1, s_scale=0.05, center_x=0.01, center_y=-0.02),
    "SIS": dict(theta_E=1.2, center_x=0.01, center_y=-0.02),
    "SHEAR": dict(gamma1=0.05, gamma2=-0.02, ra_0=0.0, dec_0=0.0),
    "CONVERGENCE": dict(kappa=0.1, ra_0=0.0, dec_0=0.0),
    "NFW": dict(Rs=1.5, alpha_Rs=0.8, center_x=0.01, center_y=-0.02),
    "TNFW": dict(Rs=1.5, alpha_Rs=0.8, r_trunc=4.0, center_x=0.01, center_y=-0.02),
    "MULTIPOLE": dict(m=4, a_m=0.02, phi_m=0.3, center_x=0.01, center_y=-0.02, r_E=1.0),
    "EPL_MULTIPOLE_M1M3M4": dict(theta_E=1.2, gamma=2.1, e1=0.2, e2=-0.1, center_x=0.01, center_y=-0.02, a1_a=0.01,
                                 delta_phi_m1=0.1, a3_a=0.01, delta_phi_m3=0.1, a4_a=0.01, delta_phi_m4=0.1),
    "CSE": dict(a=1.0, s=0.5, e1=0.2, e2=-0.1, center_x=0.01, center_y=-0.02),
    "NFW_ELLIPSE_CSE": dict(Rs=1.5, alpha_Rs=0.8, e1=0.1, e2=0.05, center_x=0.01, center_y=-0.02),
    "HERNQUIST_ELLIPSE_CSE": dict(sigma0=1.0, Rs=1.5, e1=0.1, e2=0.05, center_x=0.01, center_y=-0.02),
}
LIGHT = {
    "SERSIC": dict(amp=10.0, R_sersic=0.3, n_sersic=2.0, center_x=0.05, center_y=0.0),
    "SERSIC_ELLIPSE": dict(amp=10.0, R_sersic=0.3, n_sersic=2.0, e1=0.1, e2=-0.1, center_x=0.05, center_y=0.0),
    "GAUSSIAN": dict(amp=10.0, sigma=0.3, center_x=0.05, center_y=0.0),
}


def timed(fn, steps=5, warmup=2):
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


def batched(kw, rng):
    """First parameter of the profile varies over the batch, everything else is shared."""
    out = dict(kw)
    first = next(k for k in kw if k not in ("m",))
    out[first] = torch.as_tensor(kw[first] * (1 + 0.05 * rng.uniform(-1, 1, B)), device="cuda")
    return out


def main():
    cpu = "--no-cpu" not in sys.argv
    st = static_config()
    grid = RegularGrid(180, 180, st["kwargs_data"]["transform_pix2angle"], st["kwargs_data"]["ra_at_xy_0"],
                       st["kwargs_data"]["dec_at_xy_0"], 3)
    x, y = grid.coordinates_evaluate
    xt, yt = torch.as_tensor(x, device="cuda"), torch.as_tensor(y, device="cuda")
    rng = np.random.default_rng(0)
    n = x.size
    rows = []
    for name, kw in LENS.items():
        lm = LensModel([name])
        kwb = [batched(kw, rng)]
        ms = timed(lambda: lm.alpha(xt, yt, kwb))
        row = dict(kind="lens", profile=name, ms=ms, points_per_s=B * n / ms * 1e3)
        if cpu:
            from oracle import lens_model as olm

            t0 = time.perf_counter()
            olm.alpha([name], x, y, [kw]) if hasattr(olm, "alpha") else olm.ray_shooting([name], x, y, [kw])
            row["numpy_points_per_s"] = n / (time.perf_counter() - t0)
        rows.append(row)
        print(json.dumps(row), flush=True)
    for name, kw in LIGHT.items():
        lm = LightModel([name])
        kwb = [batched(kw, rng)]
        ms = timed(lambda: lm.surface_brightness(xt, yt, kwb))
        row = dict(kind="light", profile=name, ms=ms, points_per_s=B * n / ms * 1e3)
        if cpu:
            from oracle import light_profiles as olp

            t0 = time.perf_counter()
            olp.surface_brightness([name], x, y, [kw])
            row["numpy_points_per_s"] = n / (time.perf_counter() - t0)
        rows.append(row)
        print(json.dumps(row), flush=True)
    print("| profile | B200 points/s (256 sets x 291,600 points per call, incl. Python packing) | NumPy port, 1 core | ratio |")
    print("|---|---|---|---|")
    for r in rows:
        c = r.get("numpy_points_per_s")
        print("| %s `%s` | %.3g | %s | %s |" % (r["kind"], r["profile"], r["points_per_s"], "%.3g" % c if c else "-",
                                             "%.0fx" % (r["points_per_s"] / c) if c else "-"))


if __name__ == "__main__":
    main()
