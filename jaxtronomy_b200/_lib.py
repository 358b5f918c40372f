"""ctypes binding of the C ABI in include/lsim.h (jaxtronomy_b200/csrc/liblsim.so).

This is the only place the product touches native code.  There is NO CPU fallback: if the shared library
is missing or no CUDA device is visible, every compute entry point raises.  PyTorch is used for device
memory, streams and (in parallel.py) torch.distributed only.
"""

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("LSIM_LIB") or os.path.join(_HERE, "csrc", "liblsim.so")  # LSIM_LIB: kernel experiments

MAX_PARAMS = 12
MAX_SLOTS = 8
ABI_VERSION = 8
PRECISION_F64, PRECISION_F32 = 0, 1

# enum mirrors (include/lsim.h)
LENS_TYPES = {
    "NULL": 0, "EPL": 1, "SIE": 2, "NIE": 3, "SIS": 4, "SHEAR": 5, "CONVERGENCE": 6, "NFW": 7, "TNFW": 8,
    "MULTIPOLE": 9, "EPL_MULTIPOLE_M1M3M4": 10, "CSE": 11, "NFW_ELLIPSE_CSE": 12, "HERNQUIST_ELLIPSE_CSE": 13,
}
LIGHT_TYPES = {"SERSIC": 0, "SERSIC_ELLIPSE": 1, "GAUSSIAN": 2}
FLAG_SIE_NIE_FALSE, FLAG_CSE_MAJOR, FLAG_CSE_LOW_ACCURACY = 1, 2, 4
CONV_NONE, CONV_ZERO, CONV_EDGE = 0, 1, 2


class LensDesc(C.Structure):
    _fields_ = [
        ("type", C.c_int32), ("flags", C.c_uint32), ("col", C.c_int32 * MAX_PARAMS), ("cst", C.c_double * MAX_PARAMS),
        ("opt", C.c_double * 2), ("delta_T", C.c_double), ("T_z", C.c_double), ("red2phys", C.c_double),
    ]


class LightDesc(C.Structure):
    _fields_ = [
        ("type", C.c_int32), ("lensed", C.c_int32), ("slot", C.c_int32), ("use_amp", C.c_int32),
        ("col", C.c_int32 * MAX_PARAMS), ("cst", C.c_double * MAX_PARAMS),
    ]


_dp = C.POINTER(C.c_double)


class PointDesc(C.Structure):
    _fields_ = [("col_ra", C.c_int32), ("col_dec", C.c_int32), ("col_amp", C.c_int32), ("slot", C.c_int32),
                ("cst_ra", C.c_double), ("cst_dec", C.c_double), ("cst_amp", C.c_double), ("use_amp", C.c_int32),
                ("error_map", C.c_int32)]


class PlanDesc(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32), ("n_param_cols", C.c_int32),
        ("n_lens", C.c_int32), ("lens", C.POINTER(LensDesc)), ("multi_plane", C.c_int32),
        ("alpha_scaling", C.c_double), ("T_ij_end", C.c_double), ("T_z_source", C.c_double),
        ("n_light", C.c_int32), ("light", C.POINTER(LightDesc)), ("n_slots", C.c_int32),
        ("nx", C.c_int32), ("ny", C.c_int32), ("ss", C.c_int32), ("x_sub", _dp), ("y_sub", _dp),
        ("eval_mask", C.POINTER(C.c_uint8)), ("conv_mode", C.c_int32), ("kernel_size", C.c_int32), ("kernel", _dp), ("kernel_separable", C.c_int32),
        ("kernel_1d", _dp), ("image_scale", C.c_double),
        ("supersampled_conv", C.c_int32), ("kernel_hi_size", C.c_int32), ("kernel_hi", _dp),
        ("kernel_hi_separable", C.c_int32), ("kernel_hi_1d", _dp),
        ("data", _dp), ("cd_data", _dp), ("noise_map2", _dp), ("exposure", _dp), ("background_rms", C.c_double),
        ("mask", C.POINTER(C.c_uint8)), ("logl_offset", C.c_double), ("precision", C.c_int32),
        ("n_points", C.c_int32), ("points", C.POINTER(PointDesc)), ("ps_kernel_size", C.c_int32),
        ("ps_supersampling", C.c_int32), ("ps_kernel", _dp), ("ps_x_at_radec_0", C.c_double),
        ("ps_y_at_radec_0", C.c_double), ("ps_Ma2pix", C.c_double * 4), ("ps_scale", C.c_double),
        ("ps_variance_map", _dp), ("ps_variance_size", C.c_int32), ("ps_mean_kernel", C.c_double),
    ]


class LsimError(RuntimeError):
    pass


_lib = None

# every symbol include/lsim.h declares (tests check the library exports all of them)
SYMBOLS = [
    "lsim_plan_create", "lsim_plan_destroy", "lsim_ray_shoot", "lsim_surface_brightness", "lsim_workspace_bytes",
    "lsim_image", "lsim_resize_convolve", "lsim_logl", "lsim_abi_version", "lsim_strerror", "lsim_last_cuda_error", "lsim_launch_count",
    "lsim_measure_fp64_peak", "lsim_timing_enable", "lsim_timing_read", "lsim_pso_step", "lsim_psf_error_map",
    "lsim_logl_grad", "lsim_logl_grad_workspace_bytes", "lsim_hessian", "lsim_source_position_logl",
]


def load():
    """Load liblsim.so (once).  Raises LsimError if it has not been built: there is no fallback path."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise LsimError(
            "native library %s not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or jaxtronomy_b200/csrc/build.py). jaxtronomy_b200 has no CPU fallback." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    vp, i64, i32 = C.c_void_p, C.c_int64, C.c_int
    lib.lsim_plan_create.argtypes = [C.POINTER(PlanDesc), C.POINTER(vp)]
    lib.lsim_plan_create.restype = i32
    lib.lsim_plan_destroy.argtypes = [vp]
    lib.lsim_plan_destroy.restype = None
    lib.lsim_ray_shoot.argtypes = [vp, vp, i64, vp, vp, i64, i32, vp, vp, i32, vp]
    lib.lsim_ray_shoot.restype = i32
    lib.lsim_hessian.argtypes = [vp, vp, i64, vp, vp, i64, i32, vp, vp]
    lib.lsim_hessian.restype = i32
    lib.lsim_source_position_logl.argtypes = [vp, vp, i64, vp, vp, i32, i32, C.c_double, C.c_double, vp, vp]
    lib.lsim_source_position_logl.restype = i32
    lib.lsim_surface_brightness.argtypes = [vp, vp, i64, vp, vp, i64, i32, vp, vp]
    lib.lsim_surface_brightness.restype = i32
    lib.lsim_workspace_bytes.argtypes = [vp, i64]
    lib.lsim_workspace_bytes.restype = C.c_size_t
    lib.lsim_image.argtypes = [vp, vp, i64, vp, vp, C.c_size_t, vp]
    lib.lsim_image.restype = i32
    lib.lsim_resize_convolve.argtypes = [vp, vp, i64, i32, vp, vp, C.c_size_t, vp]
    lib.lsim_resize_convolve.restype = i32
    lib.lsim_logl.argtypes = [vp, vp, i64, i32, i32, vp, vp, vp, vp, C.c_size_t, vp]
    lib.lsim_logl.restype = i32
    lib.lsim_logl_grad_workspace_bytes.argtypes = [vp, i64]
    lib.lsim_logl_grad_workspace_bytes.restype = C.c_size_t
    lib.lsim_logl_grad.argtypes = [vp, vp, i64, i32, vp, vp, vp, vp, C.c_size_t, vp]
    lib.lsim_logl_grad.restype = i32
    lib.lsim_abi_version.restype = i32
    lib.lsim_strerror.argtypes = [i32]
    lib.lsim_strerror.restype = C.c_char_p
    lib.lsim_last_cuda_error.restype = C.c_char_p
    lib.lsim_launch_count.restype = i64
    lib.lsim_measure_fp64_peak.argtypes = [i64, C.POINTER(C.c_double), vp]
    lib.lsim_measure_fp64_peak.restype = i32
    lib.lsim_psf_error_map.argtypes = [vp, vp, i64, vp, vp]
    lib.lsim_psf_error_map.restype = i32
    lib.lsim_pso_step.argtypes = [vp] * 8 + [i64, i32, C.c_double, C.c_double, vp]
    lib.lsim_pso_step.restype = i32
    lib.lsim_timing_enable.argtypes = [i32]
    lib.lsim_timing_enable.restype = i32
    lib.lsim_timing_read.argtypes = [C.POINTER(C.c_double * 4), C.POINTER(C.c_int64 * 4)]
    lib.lsim_timing_read.restype = i32
    if lib.lsim_abi_version() != ABI_VERSION:
        raise LsimError("liblsim.so ABI version %d != binding %d: rebuild" % (lib.lsim_abi_version(), ABI_VERSION))
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        lib = load()
        msg = lib.lsim_strerror(rc).decode()
        if rc == -3:
            msg += ": " + lib.lsim_last_cuda_error().decode()
        raise LsimError("lsim error %d: %s" % (rc, msg))


def launch_count():
    return int(load().lsim_launch_count())


def _torch():
    import torch

    if not torch.cuda.is_available():
        raise LsimError("no CUDA device visible: jaxtronomy_b200 runs on the GPU only (no CPU fallback)")
    return torch


def _f64(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64))


def _ptr(a):
    return a.ctypes.data_as(_dp)


class Plan:
    """Owns one lsim_plan (device copies of the static configuration) on the current CUDA device."""

    def __init__(self, n_param_cols, lens_descs=(), light_descs=(), n_slots=1, multi_plane=False, alpha_scaling=1.0,
                 T_ij_end=0.0, T_z_source=0.0, grid=None, psf=None, image_scale=1.0, data=None, points=None):
        """
        :param grid: None or dict(nx, ny, ss, x_sub, y_sub)
        :param psf: None or dict(mode=CONV_*, kernel=[K,K], kernel_1d=[K] or None)
        :param data: None or dict(data, cd_data, noise_map2 or None, exposure or None, background_rms, mask or None)
        :param points: None or dict(descs=[PointDesc], kernel=[Kp,Kp], supersampling, x_at_radec_0, y_at_radec_0,
            Ma2pix=[2,2], scale): point-source images painted into the slot images after the convolution
        """
        torch = _torch()
        lib = load()
        self._lib = lib
        self.device = torch.device("cuda", torch.cuda.current_device())
        self.n_param_cols = int(n_param_cols)
        self.n_slots = int(n_slots)
        d = PlanDesc()
        d.abi_version = ABI_VERSION
        d.n_param_cols = self.n_param_cols
        keep = []
        lens_arr = (LensDesc * max(1, len(lens_descs)))(*lens_descs)
        light_arr = (LightDesc * max(1, len(light_descs)))(*light_descs)
        d.n_lens, d.lens = len(lens_descs), lens_arr
        d.n_light, d.light = len(light_descs), light_arr
        d.n_slots = self.n_slots
        d.multi_plane = 1 if multi_plane else 0
        d.alpha_scaling, d.T_ij_end, d.T_z_source = float(alpha_scaling), float(T_ij_end), float(T_z_source)
        self.nx = self.ny = self.ss = 0
        if grid is not None:
            self.nx, self.ny, self.ss = int(grid["nx"]), int(grid["ny"]), int(grid["ss"])
            xs, ys = _f64(grid["x_sub"]).ravel(), _f64(grid["y_sub"]).ravel()
            n_sub = self.nx * self.ny * self.ss * self.ss
            if xs.size != n_sub or ys.size != n_sub:
                raise LsimError("evaluation grid must hold nx*ny*ss^2 = %d points, got %d" % (n_sub, xs.size))
            keep += [xs, ys]
            d.nx, d.ny, d.ss, d.x_sub, d.y_sub = self.nx, self.ny, self.ss, _ptr(xs), _ptr(ys)
            d.image_scale = float(image_scale)
            from jaxtronomy_b200 import config

            d.precision = PRECISION_F64 if config.read("enable_x64") else PRECISION_F32
            if grid.get("eval_mask") is not None:
                em = np.ascontiguousarray(np.asarray(grid["eval_mask"]).astype(np.uint8)).ravel()
                if em.size != self.nx * self.ny:
                    raise LsimError("eval_mask must hold nx*ny values")
                keep.append(em)
                d.eval_mask = em.ctypes.data_as(C.POINTER(C.c_uint8))
            d.conv_mode = CONV_NONE
            if psf is not None and psf["mode"] != CONV_NONE:
                d.conv_mode = int(psf["mode"])
                if psf.get("kernel") is not None:
                    k = _f64(psf["kernel"])
                    if k.ndim != 2 or k.shape[0] != k.shape[1]:
                        raise LsimError("PSF kernel must be square, got %s" % (k.shape,))
                    keep.append(k)
                    d.kernel_size, d.kernel = k.shape[0], _ptr(k)
                    if psf.get("kernel_1d") is not None:
                        k1 = _f64(psf["kernel_1d"])
                        keep.append(k1)
                        d.kernel_separable, d.kernel_1d = 1, _ptr(k1)
                if psf.get("kernel_hi") is not None:  # supersampled convolution
                    kh = _f64(psf["kernel_hi"])
                    if kh.ndim != 2 or kh.shape[0] != kh.shape[1]:
                        raise LsimError("supersampled PSF kernel must be square, got %s" % (kh.shape,))
                    keep.append(kh)
                    d.supersampled_conv, d.kernel_hi_size, d.kernel_hi = 1, kh.shape[0], _ptr(kh)
                    if psf.get("kernel_hi_1d") is not None:
                        kh1 = _f64(psf["kernel_hi_1d"])
                        keep.append(kh1)
                        d.kernel_hi_separable, d.kernel_hi_1d = 1, _ptr(kh1)
            if data is not None:
                npix = self.nx * self.ny
                for name in ("data", "cd_data", "noise_map2", "exposure"):
                    v = data.get(name)
                    if v is None:
                        continue
                    a = _f64(np.broadcast_to(np.asarray(v, dtype=np.float64), (self.ny, self.nx))).ravel()
                    assert a.size == npix
                    keep.append(a)
                    setattr(d, name, _ptr(a))
                d.background_rms = float(data.get("background_rms") or 0.0)
                d.logl_offset = float(data.get("logl_offset") or 0.0)
                if data.get("mask") is not None:
                    m = np.ascontiguousarray(np.asarray(data["mask"]).astype(np.uint8)).ravel()
                    assert m.size == npix
                    keep.append(m)
                    d.mask = m.ctypes.data_as(C.POINTER(C.c_uint8))
        if points is not None and len(points["descs"]) > 0:
            if grid is None:
                raise LsimError("point sources need an evaluation grid")
            pts = (PointDesc * len(points["descs"]))(*points["descs"])
            pk = _f64(points["kernel"])
            keep += [pts, pk]
            d.n_points, d.points = len(points["descs"]), pts
            d.ps_kernel_size, d.ps_kernel, d.ps_supersampling = pk.shape[0], _ptr(pk), int(points["supersampling"])
            d.ps_x_at_radec_0, d.ps_y_at_radec_0 = float(points["x_at_radec_0"]), float(points["y_at_radec_0"])
            m = np.asarray(points["Ma2pix"], dtype=np.float64).ravel()
            for k in range(4):
                d.ps_Ma2pix[k] = float(m[k])
            d.ps_scale = float(points.get("scale", 1.0))
            if points.get("variance_map") is not None:
                vm = _f64(points["variance_map"])
                keep.append(vm)
                d.ps_variance_map, d.ps_variance_size = _ptr(vm), vm.shape[0]
                d.ps_mean_kernel = float(points["mean_kernel"])
        self.has_data = data is not None
        handle = C.c_void_p()
        with torch.cuda.device(self.device):
            check(lib.lsim_plan_create(C.byref(d), C.byref(handle)))
        self._handle = handle
        self._ws = None
        del keep, lens_arr, light_arr

    def __del__(self):
        h = getattr(self, "_handle", None)
        if h is not None and h.value:
            try:
                self._lib.lsim_plan_destroy(h)
            except Exception:  # pragma: no cover - interpreter shutdown
                pass
            self._handle = None

    # -- helpers ---------------------------------------------------------------------------------------
    def _params(self, params):
        torch = _torch()
        if isinstance(params, torch.Tensor):
            p = params.to(device=self.device, dtype=torch.float64).contiguous()
        else:
            p = torch.from_numpy(_f64(params)).to(self.device)
        if p.ndim != 2 or p.shape[1] != self.n_param_cols:
            raise LsimError("params must be [B, %d], got %s" % (self.n_param_cols, tuple(p.shape)))
        return p

    def _coords(self, a, B):
        torch = _torch()
        if isinstance(a, torch.Tensor):
            t = a.to(device=self.device, dtype=torch.float64).contiguous()
        else:
            t = torch.from_numpy(_f64(a)).to(self.device)
        return t

    def _stream(self):
        return C.c_void_p(_torch().cuda.current_stream(self.device).cuda_stream)

    def workspace(self, B):
        """A cached scratch buffer of the size lsim_workspace_bytes recommends for B sets."""
        torch = _torch()
        need = int(self._lib.lsim_workspace_bytes(self._handle, int(B)))
        if self._ws is None or self._ws.numel() * 8 < need:
            self._ws = torch.empty((need + 7) // 8, dtype=torch.float64, device=self.device)
        return self._ws

    # -- launches (all return torch CUDA tensors, enqueued on the current stream) -------------------
    def ray_shoot(self, params, x, y, what=0):
        torch = _torch()
        p = self._params(params)
        B = p.shape[0]
        xt, yt = self._coords(x, B), self._coords(y, B)
        if xt.shape != yt.shape:
            raise LsimError("x and y must have the same shape")
        batched = 1 if (xt.ndim == 2) else 0
        if batched and xt.shape[0] != B:
            raise LsimError("batched coordinates must be [B, N]")
        N = xt.shape[-1] if xt.ndim else 1
        ox = torch.empty((B, N), dtype=torch.float64, device=self.device)
        oy = torch.empty_like(ox)
        with torch.cuda.device(self.device):
            check(self._lib.lsim_ray_shoot(self._handle, p.data_ptr(), B, xt.data_ptr(), yt.data_ptr(), N, batched,
                                           ox.data_ptr(), oy.data_ptr(), int(what), self._stream()))
        return ox, oy

    def hessian(self, params, x, y):
        """-> [B, 4, N]: f_xx, f_xy, f_yx, f_yy (lsim_hessian; single-plane fp64 plans)."""
        torch = _torch()
        p = self._params(params)
        B = p.shape[0]
        xt, yt = self._coords(x, B), self._coords(y, B)
        if xt.shape != yt.shape:
            raise LsimError("x and y must have the same shape")
        batched = 1 if (xt.ndim == 2) else 0
        if batched and xt.shape[0] != B:
            raise LsimError("batched coordinates must be [B, N]")
        N = xt.shape[-1] if xt.ndim else 1
        out = torch.empty((B, 4, N), dtype=torch.float64, device=self.device)
        with torch.cuda.device(self.device):
            check(self._lib.lsim_hessian(self._handle, p.data_ptr(), B, xt.data_ptr(), yt.data_ptr(), N, batched,
                                         out.data_ptr(), self._stream()))
        return out

    def source_position_logl(self, params, x, y, sigma, hard_bound_rms=None):
        """-> logL [B] of PositionLikelihood.source_position_likelihood for one LENSED_POSITION model
        (lsim_source_position_logl); x, y: image positions [n_img] (shared) or [B, n_img]."""
        torch = _torch()
        p = self._params(params)
        B = p.shape[0]
        xt, yt = self._coords(x, B), self._coords(y, B)
        if xt.shape != yt.shape or xt.ndim not in (1, 2):
            raise LsimError("image positions must be [n_img] or [B, n_img], the same for x and y")
        batched = 1 if (xt.ndim == 2) else 0
        if batched and xt.shape[0] != B:
            raise LsimError("batched image positions must be [B, n_img]")
        out = torch.empty((B,), dtype=torch.float64, device=self.device)
        with torch.cuda.device(self.device):
            check(self._lib.lsim_source_position_logl(
                self._handle, p.data_ptr(), B, xt.data_ptr(), yt.data_ptr(), int(xt.shape[-1]), batched, float(sigma),
                -1.0 if hard_bound_rms is None else float(hard_bound_rms), out.data_ptr(), self._stream()))
        return out

    def surface_brightness(self, params, x, y):
        torch = _torch()
        p = self._params(params)
        B = p.shape[0]
        xt, yt = self._coords(x, B), self._coords(y, B)
        batched = 1 if (xt.ndim == 2) else 0
        N = xt.shape[-1] if xt.ndim else 1
        out = torch.empty((B, self.n_slots, N), dtype=torch.float64, device=self.device)
        with torch.cuda.device(self.device):
            check(self._lib.lsim_surface_brightness(self._handle, p.data_ptr(), B, xt.data_ptr(), yt.data_ptr(), N,
                                                    batched, out.data_ptr(), self._stream()))
        return out

    def image(self, params):
        torch = _torch()
        p = self._params(params)
        B = p.shape[0]
        out = torch.empty((B, self.n_slots, self.ny, self.nx), dtype=torch.float64, device=self.device)
        ws = self.workspace(B)
        with torch.cuda.device(self.device):
            check(self._lib.lsim_image(self._handle, p.data_ptr(), B, out.data_ptr(), ws.data_ptr(), ws.numel() * 8,
                                       self._stream()))
        return out

    def resize_convolve(self, flux, unconvolved=False):
        """flux [n, ny*ss*nx*ss] (device or host) -> images [n, ny, nx]"""
        torch = _torch()
        f = self._coords(flux, 0)
        if f.ndim == 1:
            f = f[None]
        n = f.shape[0]
        if f.shape[1] != self.nx * self.ny * self.ss * self.ss:
            raise LsimError("flux must hold nx*ny*ss^2 values per image")
        out = torch.empty((n, self.ny, self.nx), dtype=torch.float64, device=self.device)
        ws = self.workspace(max(n // max(self.n_slots, 1), 1))
        with torch.cuda.device(self.device):
            check(self._lib.lsim_resize_convolve(self._handle, f.data_ptr(), n, int(bool(unconvolved)), out.data_ptr(),
                                                 ws.data_ptr(), ws.numel() * 8, self._stream()))
        return out

    def logl(self, params, linear_solve=False, source_marg=False, return_model=False, out=None):
        """-> (logl [B], amps [B, n_slots] or None, model [B, ny, nx] or None)"""
        torch = _torch()
        p = self._params(params)
        B = p.shape[0]
        logl = out if out is not None else torch.empty(B, dtype=torch.float64, device=self.device)
        amps = torch.empty((B, self.n_slots), dtype=torch.float64, device=self.device) if linear_solve else None
        model = torch.empty((B, self.ny, self.nx), dtype=torch.float64, device=self.device) if return_model else None
        ws = self.workspace(B)
        with torch.cuda.device(self.device):
            check(self._lib.lsim_logl(self._handle, p.data_ptr(), B, int(bool(linear_solve)), int(bool(source_marg)),
                                      logl.data_ptr(), amps.data_ptr() if amps is not None else None,
                                      model.data_ptr() if model is not None else None, ws.data_ptr(), ws.numel() * 8,
                                      self._stream()))
        return logl, amps, model


def _plan_logl_grad(self, params, linear_solve=False):
    """[B, P] -> (logl [B], grad [B, P] = d logl / d params, amps [B, n_slots] or None)   (lsim_logl_grad)"""
    torch = _torch()
    p = self._params(params)
    B = p.shape[0]
    logl = torch.empty(B, dtype=torch.float64, device=self.device)
    grad = torch.empty((B, self.n_param_cols), dtype=torch.float64, device=self.device)
    amps = torch.empty((B, self.n_slots), dtype=torch.float64, device=self.device) if linear_solve else None
    need = int(self._lib.lsim_logl_grad_workspace_bytes(self._handle, max(int(B), 1)))
    if need == 0:
        check(-2)  # LSIM_E_UNSUPPORTED: the model is outside the gradient's scope (see include/lsim.h)
    if self._ws is None or self._ws.numel() * 8 < need:
        self._ws = torch.empty((need + 7) // 8, dtype=torch.float64, device=self.device)
    ws = self._ws
    with torch.cuda.device(self.device):
        check(self._lib.lsim_logl_grad(self._handle, p.data_ptr(), B, int(bool(linear_solve)), logl.data_ptr(),
                                       grad.data_ptr(), amps.data_ptr() if amps is not None else None, ws.data_ptr(),
                                       ws.numel() * 8, self._stream()))
    return logl, grad, amps


def _plan_psf_error_map(self, params):
    """[B, P] -> PSF error map sigma^2 [B, ny, nx] (lsim_psf_error_map)."""
    torch = _torch()
    p = self._params(params)
    B = p.shape[0]
    out = torch.empty((B, self.ny, self.nx), dtype=torch.float64, device=self.device)
    with torch.cuda.device(self.device):
        check(self._lib.lsim_psf_error_map(self._handle, p.data_ptr(), B, out.data_ptr(), self._stream()))
    return out


Plan.psf_error_map = _plan_psf_error_map
Plan.logl_grad = _plan_logl_grad


def pso_step(position, velocity, pbest_pos, pbest_fit, gbest_pos, gbest_fit, fitness, rnd, c1=1.193, c2=1.193):
    """In-place particle-swarm iteration on CUDA tensors (lsim_pso_step); rnd: [3, B, P] uniform numbers."""
    torch = _torch()
    B, P = position.shape
    for t, shape in ((position, (B, P)), (velocity, (B, P)), (pbest_pos, (B, P)), (pbest_fit, (B,)), (gbest_pos, (P,)),
                     (gbest_fit, (1,)), (fitness, (B,)), (rnd, (3, B, P))):
        if not (t.is_cuda and t.dtype == torch.float64 and t.is_contiguous() and tuple(t.shape) == shape):
            raise LsimError("pso_step needs contiguous float64 CUDA tensors; got %s %s for expected shape %s"
                            % (t.dtype, tuple(t.shape), shape))
    ptr = lambda t: C.c_void_p(t.data_ptr())  # noqa: E731
    check(load().lsim_pso_step(ptr(position), ptr(velocity), ptr(pbest_pos), ptr(pbest_fit), ptr(gbest_pos),
                               ptr(gbest_fit), ptr(fitness), ptr(rnd), B, P, float(c1), float(c2),
                               C.c_void_p(torch.cuda.current_stream().cuda_stream)))


def measure_fp64_peak(iters=200000):
    """DFMA-chain microbenchmark (lsim_measure_fp64_peak): achieved fp64 FLOP/s of the vector pipe."""
    torch = _torch()
    lib = load()
    out = C.c_double(0.0)
    check(lib.lsim_measure_fp64_peak(int(iters), C.byref(out), C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    return out.value


def set_timing(on):
    """Bracket every launch with CUDA events on its stream (see lsim_timing_enable); for bench.py's roofline."""
    check(load().lsim_timing_enable(1 if on else 0))


def get_timing():
    """-> {class: {"ms": summed device ms, "count": launches}} since the last call; synchronises."""
    ms, cnt = (C.c_double * 4)(), (C.c_int64 * 4)()
    check(load().lsim_timing_read(C.byref(ms), C.byref(cnt)))
    names = ["image", "conv", "likelihood", "other"]
    return {n: {"ms": float(ms[i]), "count": int(cnt[i])} for i, n in enumerate(names) if cnt[i] > 0 or i < 3}
