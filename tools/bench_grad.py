"""Throughput of lsim_logl_grad (logL + d logL / d params) at the headline shape:
EPL+SHEAR, SERSIC_ELLIPSE source, 180x180, ss 3, 21x21 pixel PSF, ImageLinearFit, B parameter sets resident in HBM.

    python tools/bench_grad.py [--batch 1184] [--steps 5] [--warmup 3]

Prints one JSON line: gradients/s, the forward-only time of the same batch, and the per-kernel-class CUDA-event
times of one call (image = k_image + k_image_grad, conv = forward + transposed PSF, like = k_like_adjoint)."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch

    from bench import build_headline
    from jaxtronomy_b200 import _lib

    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=1184)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    args = ap.parse_args()
    cfg = build_headline(n_sets=args.batch, seed=1234)
    model, t = cfg["model"], cfg["params"]

    def timed(fn):
        for _ in range(args.warmup):
            fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(args.steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / args.steps

    ms_grad = timed(lambda: model.logL_linear_grad_batch(t))
    ms_fwd = timed(lambda: model.logL_linear_batch(t))
    _lib.set_timing(True)
    model.logL_linear_grad_batch(t)
    torch.cuda.synchronize()
    tm = _lib.get_timing()
    _lib.set_timing(False)
    print(json.dumps({"metric": "likelihood gradients/s (logL + d logL/d 16 parameters per set)",
                      "value": args.batch / (ms_grad * 1e-3), "batch": args.batch, "ms_per_call": ms_grad,
                      "forward_only_ms": ms_fwd, "grad_over_forward": ms_grad / ms_fwd,
                      "kernel_ms": {k: v["ms"] for k, v in tm.items()},
                      "kernel_launches": {k: v["count"] for k, v in tm.items()}}))


if __name__ == "__main__":
    main()
