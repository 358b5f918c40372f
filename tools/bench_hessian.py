"""Throughput of the two kernels added past SURVEY.md 8: k_hessian and k_source_position (CUDA events, warm-up 3).

    python tools/bench_hessian.py            # prints one JSON line per kernel

k_hessian writes 4 doubles per (set, point): algorithmic bytes = 32 B per point (coordinates are shared by the sets and
L2-resident); with EPL in the model it is fp64-bound like k_ray_shoot.  k_source_position is latency / preparation
bound (one warp per set, 4 images): reported as sets/s.
"""

import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def timed(fn, steps=10, warmup=3):
    import torch

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


def main():
    import torch

    from jaxtronomy_b200.LensModel.lens_model import LensModel

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))
    except OSError:
        pass
    hbm = float(peaks.get("hbm_gbs", 6550.0))
    rng = np.random.default_rng(1234)
    for names in (["EPL", "SHEAR"], ["SIS", "SHEAR"]):
        lm = LensModel(names)
        plan = lm._plan(None)
        B, n = 1024, 180
        g = (np.arange(n) - (n - 1) / 2) * 0.05
        x, y = (torch.as_tensor(a.ravel(), device="cuda") for a in np.meshgrid(g, g))
        P = plan.n_param_cols
        params = np.zeros((B, P))
        if names[0] == "EPL":
            params[:, 0] = rng.uniform(0.8, 1.6, B); params[:, 1] = rng.uniform(1.7, 2.3, B)
            params[:, 2] = rng.uniform(-0.2, 0.2, B); params[:, 3] = rng.uniform(-0.2, 0.2, B)
            params[:, 6:8] = rng.uniform(-0.08, 0.08, (B, 2))
        else:
            params[:, 0] = rng.uniform(0.8, 1.6, B); params[:, 3:5] = rng.uniform(-0.08, 0.08, (B, 2))
        pt = torch.as_tensor(params, device="cuda")
        ms = timed(lambda: plan.hessian(pt, x, y))
        by = B * x.numel() * 32
        print(json.dumps({"kernel": "k_hessian", "model": "+".join(names), "sets": B, "points": int(x.numel()),
                          "ms": round(ms, 4), "sets_per_s": round(B / ms * 1e3, 1),
                          "hessians_per_s": round(B * x.numel() / ms * 1e3, 1),
                          "GBps_written": round(by / ms / 1e6, 1), "frac_hbm": round(by / ms / 1e6 / hbm, 4),
                          "note": "includes the [B,4,N] output allocation of Plan.hessian"}))
    lm = LensModel(["EPL", "SHEAR"])
    plan = lm._plan(None)
    B = 65536
    params = np.zeros((B, plan.n_param_cols))
    params[:, 0] = rng.uniform(0.8, 1.6, B); params[:, 1] = rng.uniform(1.7, 2.3, B)
    params[:, 2] = rng.uniform(-0.2, 0.2, B); params[:, 3] = rng.uniform(-0.2, 0.2, B)
    params[:, 6:8] = rng.uniform(-0.08, 0.08, (B, 2))
    pt = torch.as_tensor(params, device="cuda")
    ra = torch.as_tensor(np.array([1.013, -0.734, 0.412, -0.551]) + rng.normal(0, 0.01, (B, 4)), device="cuda")
    dec = torch.as_tensor(np.array([0.402, -0.846, -1.002, 0.978]) + rng.normal(0, 0.01, (B, 4)), device="cuda")
    ms = timed(lambda: plan.source_position_logl(pt, ra, dec, 0.001, 0.1))
    print(json.dumps({"kernel": "k_source_position", "model": "EPL+SHEAR", "sets": B, "images": 4, "ms": round(ms, 4),
                      "sets_per_s": round(B / ms * 1e3, 1)}))


if __name__ == "__main__":
    main()
