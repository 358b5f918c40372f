#!/usr/bin/env python
"""Headline benchmark: simulated images/s for EPL + SHEAR lens, SERSIC_ELLIPSE source, 180x180 pixels,
supersampling 3, 21x21 pixel PSF, ImageLinearFit likelihood (image + linear amplitude + chi^2), on a
65,536-particle batch sharded over the GPUs of one box (BASELINE.json metric; SURVEY.md 8(d) inputs).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--impl reference]

One step = one pass of the hot path over the whole particle batch.  Rank 0 prints ONE JSON line.
`--impl reference` times the CPU restatement of the reference (oracle/, NumPy) on all host cores on a bounded
sample of the same workload (the reference itself needs jax + lenstronomy, which cannot be installed offline).
"""

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "simulated images/sec (EPL+SHEAR+SERSIC_ELLIPSE, 180x180 ss=3, 21x21 PSF, linear-solve likelihood)"
N_SUB_PER_IMAGE = 180 * 180 * 9
# 291,600 super-pixels x 566 flop + 32,400 pixels x (882 + 4 + 24 + 2) flop
FLOP_PER_SUPERPIXEL = 566.0
FLOP_PER_PIXEL_POST = 2 * 21 * 21 + 4 + 24 + 2
FLOP_PER_IMAGE = N_SUB_PER_IMAGE * FLOP_PER_SUPERPIXEL + 180 * 180 * FLOP_PER_PIXEL_POST

LENS_LIST = ["EPL", "SHEAR"]
SOURCE_LIST = ["SERSIC_ELLIPSE"]


def sample_params(n_sets, seed=1234):
    """[B, 17] parameter matrix, columns in ImageModel.param_layout order; SURVEY.md 8(d) priors
    (|e| in U(0.15, 0.35): every set uses the nmax = 20 series branch)."""
    rng = np.random.default_rng(seed)
    emod, eang = rng.uniform(0.15, 0.35, n_sets), rng.uniform(0, 2 * np.pi, n_sets)
    cols = [
        rng.uniform(0.8, 1.6, n_sets), rng.uniform(1.7, 2.3, n_sets), emod * np.cos(eang), emod * np.sin(eang),
        rng.uniform(-0.1, 0.1, n_sets), rng.uniform(-0.1, 0.1, n_sets),                      # EPL
        rng.uniform(-0.08, 0.08, n_sets), rng.uniform(-0.08, 0.08, n_sets), np.zeros(n_sets), np.zeros(n_sets),  # SHEAR
        np.ones(n_sets), rng.uniform(0.1, 0.5, n_sets), rng.uniform(1.0, 4.0, n_sets), rng.uniform(-0.3, 0.3, n_sets),
        rng.uniform(-0.3, 0.3, n_sets), rng.uniform(-0.2, 0.2, n_sets), rng.uniform(-0.2, 0.2, n_sets),  # SERSIC_ELLIPSE
    ]
    return np.ascontiguousarray(np.stack(cols, axis=1))


TRUTH = np.array([1.2, 2.0, 0.2, -0.1, 0.02, -0.03, 0.04, -0.02, 0.0, 0.0, 20.0, 0.3, 2.0, 0.1, -0.15, 0.05, 0.1])


def row_to_kwargs(row):
    kl = [dict(theta_E=row[0], gamma=row[1], e1=row[2], e2=row[3], center_x=row[4], center_y=row[5]),
          dict(gamma1=row[6], gamma2=row[7], ra_0=row[8], dec_0=row[9])]
    ks = [dict(amp=row[10], R_sersic=row[11], n_sersic=row[12], e1=row[13], e2=row[14], center_x=row[15],
               center_y=row[16])]
    return kl, ks, []


def psf_kernel(size, sigma=2.0):
    a = np.arange(size) - (size - 1) / 2.0
    k = np.exp(-(a[:, None] ** 2 + a[None, :] ** 2) / (2 * sigma**2))
    return k / k.sum()


def static_config(num_pix=180, delta_pix=0.05, ss=3, kernel_size=21):
    half = delta_pix * (num_pix - 1) / 2.0
    return dict(num_pix=num_pix, delta_pix=delta_pix, ss=ss, kernel=psf_kernel(kernel_size),
                kwargs_data=dict(exposure_time=100.0, background_rms=0.05, ra_at_xy_0=-half, dec_at_xy_0=-half,
                                 transform_pix2angle=np.array([[delta_pix, 0.0], [0.0, delta_pix]])))


def build_headline(num_pix=180, ss=3, kernel_size=21, n_sets=1024, seed=1234, data_image=None):
    """GPU model (ImageLinearFit) + parameter matrix for the headline workload."""
    import torch

    from jaxtronomy_b200.Data.imaging_data import ImageData
    from jaxtronomy_b200.Data.psf import PSF
    from jaxtronomy_b200.ImSim.image_linear_solve import ImageLinearFit
    from jaxtronomy_b200.LensModel.lens_model import LensModel
    from jaxtronomy_b200.LightModel.light_model import LightModel

    st = static_config(num_pix=num_pix, ss=ss, kernel_size=kernel_size)

    def make(image):
        data = ImageData(image_data=image, **st["kwargs_data"])
        psf = PSF(psf_type="PIXEL", kernel_point_source=st["kernel"])
        return ImageLinearFit(data, psf, LensModel(LENS_LIST), LightModel(SOURCE_LIST), LightModel([]),
                              kwargs_numerics=dict(supersampling_factor=ss, supersampling_convolution=False))

    if data_image is None:
        # data = reference-parameter simulation + seeded Gaussian + Poisson-like noise
        truth = make(np.zeros((num_pix, num_pix))).image_batch(TRUTH[None, :])[0].cpu().numpy()
        rng = np.random.default_rng(0)
        data_image = truth + rng.normal(0, 0.05, truth.shape) + rng.normal(0, 1, truth.shape) * np.sqrt(
            np.maximum(truth, 0) / 100.0)
    model = make(data_image)
    params_host = sample_params(n_sets, seed)
    return dict(model=model, params=torch.as_tensor(params_host, device="cuda"), params_host=params_host,
                static=st, data_image=data_image, num_pix=num_pix, ss=ss)


def build_oracle(st, data_image, ss):
    from oracle import imsim
    from oracle.lenstronomy_port import PSF as OPSF

    data = imsim.ImageData(image_data=data_image, **st["kwargs_data"])
    psf = OPSF(psf_type="PIXEL", kernel_point_source=st["kernel"])
    return imsim.ImageModel(data, psf, LENS_LIST, SOURCE_LIST, [], dict(supersampling_factor=ss))


def oracle_headline_logl(cfg, b):
    om = build_oracle(cfg["static"], cfg["data_image"], cfg["ss"])
    kl, ks, kll = row_to_kwargs(cfg["params_host"][b])
    return om.likelihood_data_given_model_linear(kl, ks, kll)


# ---------------------------------------------------------------------------------------------------------------
# CPU arm: the oracle restatement on the host cores
# ---------------------------------------------------------------------------------------------------------------
_W = {}


def _cpu_init(st, data_image, ss):
    os.environ["OMP_NUM_THREADS"] = "1"
    _W["om"] = build_oracle(st, data_image, ss)


def _cpu_work(rows):
    out = []
    for row in rows:
        kl, ks, kll = row_to_kwargs(row)
        out.append(_W["om"].likelihood_data_given_model_linear(kl, ks, kll)[0])
    return out


def cpu_images_per_s(n_images, cores, st=None, data_image=None, ss=3):
    """Time the NumPy oracle on `n_images` parameter sets over `cores` processes -> (images/s, seconds)."""
    import multiprocessing as mp

    st = st or static_config()
    if data_image is None:
        data_image = np.random.default_rng(0).normal(0, 0.05, (st["num_pix"], st["num_pix"]))
    rows = sample_params(n_images, seed=99)
    chunks = [rows[i::cores] for i in range(cores)]
    ctx = mp.get_context("fork")
    with ctx.Pool(cores, initializer=_cpu_init, initargs=(st, data_image, ss)) as pool:
        pool.map(_cpu_work, [c[:1] for c in chunks])  # warm-up (imports, FFT plans)
        t0 = time.perf_counter()
        pool.map(_cpu_work, chunks)
        dt = time.perf_counter() - t0
    return n_images / dt, dt


# ---------------------------------------------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock + throttle reasons sampled DURING the timed region: NVML polled from a thread every 10 ms (no process
    start-up latency, so even a 0.2 s region at 8 GPUs gets samples); `nvidia-smi -lms` is the fallback."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.rows, self.proc, self._stop, self.thread = index, [], None, threading.Event(), None
        self._nvml = None

    def _visible_index(self):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            try:
                return int(vis.split(",")[self.index])
            except (ValueError, IndexError):
                pass
        return self.index

    def __enter__(self):
        try:
            import pynvml

            pynvml.nvmlInit()
            self._nvml = pynvml
            self._handle = pynvml.nvmlDeviceGetHandleByIndex(self._visible_index())
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return self
        except Exception:
            self._nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self._visible_index()), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None
        return self

    def _poll(self):
        nv = self._nvml
        flags = (("hw_slowdown", getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8)),
                 ("hw_thermal_slowdown", getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40)),
                 ("sw_thermal_slowdown", getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20)),
                 ("sw_power_cap", getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4)))
        get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
            getattr(nv, "nvmlDeviceGetCurrentClocksThrottleReasons")
        smax = nv.nvmlDeviceGetMaxClockInfo(self._handle, nv.NVML_CLOCK_SM)
        while not self._stop.is_set():
            try:
                sm = nv.nvmlDeviceGetClockInfo(self._handle, nv.NVML_CLOCK_SM)
                mask = get_reasons(self._handle)
                self.rows.append([str(sm), str(smax), ""] + ["Active" if mask & bit else "Not Active" for _, bit in flags])
            except Exception:
                pass
            self._stop.wait(0.01)

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *a):
        self._stop.set()
        if self._nvml is not None:
            if self.thread is not None:
                self.thread.join(timeout=2)
            try:
                self._nvml.nvmlShutdown()
            except Exception:
                pass
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                smax.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for n, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(np.max(smax)), "reasons": sorted(reasons),
                "samples": len(sm), "source": "nvml" if self._nvml is not None else "nvidia-smi"}


# ---------------------------------------------------------------------------------------------------------------
def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    n = args.cpu_images if args.cpu_images else 8 * cores
    vals = []
    for _ in range(args.warmup):
        cpu_images_per_s(cores, cores)
    for _ in range(args.steps):
        v, dt = cpu_images_per_s(n, cores)
        vals.append(v)
    value = float(np.mean(vals))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "images/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * n / value, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "headline: EPL+SHEAR, SERSIC_ELLIPSE source, 180x180, ss=3, 21x21 pixel PSF, "
                               "ImageLinearFit likelihood", "sample_images_per_step": n},
        "cpu_baseline": {"value": value, "unit": "images/s", "cores": cores, "kind": "port",
                         "sample": "%d parameter sets per step, NumPy oracle restatement (jax + lenstronomy are not "
                                   "installable offline), one process per core" % n},
        "e2e": {"value": value, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def run_gpu(args):
    import torch
    import torch.distributed as dist

    from jaxtronomy_b200 import _lib, parallel

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    B = args.batch
    if args.precision == "f32":
        import jaxtronomy_b200

        jaxtronomy_b200.config.update("enable_x64", False)
    cfg = build_headline(n_sets=B, seed=1234)
    model = cfg["model"]
    lo, hi = parallel.shard_bounds(B, world, rank)
    shard = cfg["params"][lo:hi].contiguous()
    host_params = torch.as_tensor(cfg["params_host"][lo:hi]).pin_memory()
    logl_all = torch.empty(B, dtype=torch.float64, device="cuda")
    host_logl = torch.empty(B, dtype=torch.float64).pin_memory()

    def step_device():
        logl, _, _ = model.logL_linear_batch(shard, out=logl_all[lo:hi])
        parallel.all_gather_logl(logl_all, world, rank, B)
        return logl

    def step_e2e():
        dev = host_params.to("cuda", non_blocking=True)
        model.logL_linear_batch(dev, out=logl_all[lo:hi])
        parallel.all_gather_logl(logl_all, world, rank, B)
        host_logl.copy_(logl_all, non_blocking=True)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    for _ in range(max(args.warmup, 3)):
        step_device()
    launches0 = _lib.launch_count()
    with ClockSampler(local_rank) as clk:
        ms_total = timed(step_device, args.steps)
    launches = _lib.launch_count() - launches0
    for _ in range(2):
        step_e2e()
    ms_e2e = timed(step_e2e, args.steps)
    value = B * args.steps / (ms_total * 1e-3)
    e2e_value = B * args.steps / (ms_e2e * 1e-3)

    # roofline of the dominant kernel (k_image), timed live with CUDA events on the launch stream
    roof, cpu = None, None
    if rank == 0:
        peak = _lib.measure_fp64_peak(400000)
        _lib.set_timing(True)
        model.logL_linear_batch(shard, out=logl_all[lo:hi])  # local work only: no collective outside the timed steps
        torch.cuda.synchronize()
        tm = _lib.get_timing()
        _lib.set_timing(False)
        n_img_launch = max(tm["image"]["count"], 1)
        sets_per_launch = (hi - lo) / n_img_launch
        flop_per_launch = sets_per_launch * N_SUB_PER_IMAGE * FLOP_PER_SUPERPIXEL
        avg_ms = tm["image"]["ms"] / n_img_launch
        achieved = flop_per_launch / (avg_ms * 1e-3) / 1e12
        total_ms = sum(v["ms"] for v in tm.values())
        roof = {"bound": "fp64", "kernel": "k_image (fused ray-shoot + light + re-bin)", "achieved": achieved,
                "peak": peak / 1e12, "unit": "TFLOP/s", "frac": achieved / (peak / 1e12),
                # dram__bytes_read.sum + dram__bytes_write.sum of one k_image launch of 1184 sets, from the ncu
                # --set full capture profiles/image_r1h_ncu_summary.txt (5.7 MB + 255.7 MB: the re-binned images,
                # 1184 x 259 KB = 306.9 MB algorithmic; the tail of the writes is still in L2 when the kernel ends)
                "traffic": 261.4e6 * sets_per_launch / 1184.0,
                "peak_source": "measured DFMA-chain microbenchmark (lsim_measure_fp64_peak); MEASURED_PEAKS.json has "
                               "no fp64 entry; nominal 37.2 TFLOP/s",
                "flop_per_launch": flop_per_launch, "avg_launch_ms": avg_ms,
                "kernel_share_of_step": {k: v["ms"] / total_ms for k, v in tm.items()},
                "whole_step_frac": value / world * FLOP_PER_IMAGE / peak}
        if world == 1 and not args.no_cpu:
            cores = os.cpu_count() or 1
            n = args.cpu_images if args.cpu_images else 16 * cores
            v, dt = cpu_images_per_s(n, cores)
            cpu = {"value": v, "unit": "images/s", "cores": cores, "kind": "port",
                   "sample": "%d parameter sets of the same workload, NumPy oracle restatement, one process per core, "
                             "%.1f s" % (n, dt)}
    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "images/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_total / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": args.precision, "data": "synthetic",
            "config": {"workload": "headline: EPL+SHEAR, SERSIC_ELLIPSE source, 180x180, ss=3, 21x21 pixel PSF, "
                                   "ImageLinearFit likelihood", "particles": B, "parallelism": "particle shards x%d, "
                       "NCCL all-gather of logL" % world,
                       "l2": "no flush: each step streams %.1f GB of scratch images per GPU (>> 126 MB L2); the only "
                             "reused inputs are the 4.7 MB grid and the parameter matrix" %
                             ((hi - lo) * 2 * 180 * 180 * 8 / 1e9)},
            "e2e": {"value": e2e_value, "unit": "images/s", "h2d_bytes_per_step": int(host_params.numel() * 8),
                    "d2h_bytes_per_step": int(B * 8)},
            "gpu_launches": int(launches), "clocks": clk.summary(), "roofline": roof, "cpu_baseline": cpu,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=65536)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--cpu-images", type=int, default=0)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--precision", default="f64", choices=["f64", "f32"],
                    help="f32: the opt-in single-precision image kernel (NOT the headline; dtype is reported)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
