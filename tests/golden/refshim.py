"""Execute the UNMODIFIED reference source (/root/reference/jaxtronomy) without jax / lenstronomy / astropy.

Used only by tests/golden/make_golden.py, in the build container, to produce the committed
golden fixtures (tests/golden/*.npz).  Nothing here travels into the product or runs on the GPU box.

What is substituted, and why it is a faithful stand-in for generating fp64 reference values:

* ``jax`` -> a NumPy-backed shim.  ``jax.numpy`` is NumPy (same IEEE-754 fp64 elementwise ops; XLA may
  contract a*b+c into FMA and reorder reductions, which moves results at the 1e-16 level, far inside
  the 1e-10 parity bar).  ``jit`` is the identity, ``lax.fori_loop/switch/cond/map/scan`` are the
  obvious Python loops, arrays get a functional ``.at[...]`` and non-mutating in-place operators so
  that jax's immutable-array semantics are preserved.
* ``lenstronomy`` / ``astropy`` -> only host-side, init-time helpers are needed on the hot path
  (grids, PixelGrid, PSF, kernel cut, Image2SourceMapping, convert_bool_list, LensProfileBase).  They are
  restated from lenstronomy's published behaviour (oracle/lenstronomy_port.py);
  every other lenstronomy/astropy name resolves to an inert placeholder.

The per-pixel arithmetic -- every profile, the 2F1 series, re-binning, convolution wiring, the WLS
solve and the likelihood -- is the reference's own code.
"""

import importlib.abc
import importlib.machinery
import sys
import types

import numpy as np
import scipy.ndimage
import scipy.signal
import scipy.special

REFERENCE_ROOT = "/root/reference"


# ------------------------------------------------------------------------------------------------
# jax shim
# ------------------------------------------------------------------------------------------------
class _At:
    def __init__(self, arr):
        self._arr = arr

    def __getitem__(self, idx):
        return _AtIdx(self._arr, idx)


class _AtIdx:
    def __init__(self, arr, idx):
        self._arr, self._idx = arr, idx

    def set(self, value):
        out = np.array(self._arr, copy=True)
        out[self._idx] = value
        return _wrap(out)

    def add(self, value):
        out = np.array(self._arr, copy=True)
        np.add.at(out, self._idx, value)
        return _wrap(out)

    def get(self):
        return _wrap(np.asarray(self._arr)[self._idx])


class JArray(np.ndarray):
    """ndarray with jax's functional ``.at`` and immutable-style in-place operators."""

    @property
    def at(self):
        return _At(self)

    def __iadd__(self, o):
        return self + o

    def __isub__(self, o):
        return self - o

    def __imul__(self, o):
        return self * o

    def __itruediv__(self, o):
        return self / o


def _wrap(x):
    if isinstance(x, np.ndarray) and not isinstance(x, JArray):
        return x.view(JArray)
    if isinstance(x, (np.floating, np.complexfloating, np.integer, np.bool_)):
        return np.asarray(x).view(JArray)
    if isinstance(x, tuple):
        return tuple(_wrap(v) for v in x)
    if isinstance(x, list):
        return [_wrap(v) for v in x]
    return x


def _wrapped(fn):
    def call(*a, **k):
        return _wrap(fn(*a, **k))

    call.__name__ = getattr(fn, "__name__", "fn")
    return call


class _NumpyModule(types.ModuleType):
    def __getattr__(self, name):
        obj = getattr(np, name)
        if callable(obj) and not isinstance(obj, type):
            obj = _wrapped(obj)
        setattr(self, name, obj)
        return obj


def _jit(fun=None, **kwargs):
    if fun is None:
        return lambda f: f
    return fun


class _CustomJvp:
    def __init__(self, fun):
        self.fun = fun

    def defjvp(self, f):
        return f

    def __call__(self, *a, **k):
        fun = self.fun.__func__ if isinstance(self.fun, staticmethod) else self.fun
        return fun(*a, **k)


def _tree_index(tree, i):
    if isinstance(tree, dict):
        return {k: _tree_index(v, i) for k, v in tree.items()}
    if isinstance(tree, (tuple, list)):
        return type(tree)(_tree_index(v, i) for v in tree)
    return _wrap(np.asarray(tree)[i])


def _tree_len(tree):
    if isinstance(tree, dict):
        return _tree_len(next(iter(tree.values())))
    if isinstance(tree, (tuple, list)):
        return _tree_len(tree[0])
    return len(tree)


def _tree_stack(items):
    first = items[0]
    if isinstance(first, dict):
        return {k: _tree_stack([it[k] for it in items]) for k in first}
    if isinstance(first, (tuple, list)):
        return type(first)(_tree_stack([it[j] for it in items]) for j in range(len(first)))
    return _wrap(np.stack([np.asarray(it) for it in items]))


def _fori_loop(lo, hi, body, init):
    val = init
    for i in range(int(lo), int(hi)):
        val = body(i, val)
    return val


def _switch(index, branches, *operands):
    return branches[int(np.clip(int(index), 0, len(branches) - 1))](*operands)


def _cond(pred, true_fun, false_fun, *operands):
    return true_fun(*operands) if bool(pred) else false_fun(*operands)


def _map(f, xs):
    return _tree_stack([f(_tree_index(xs, i)) for i in range(_tree_len(xs))])


def _scan(f, init, xs, length=None):
    carry, ys = init, []
    n = _tree_len(xs) if xs is not None else length
    for i in range(n):
        carry, y = f(carry, _tree_index(xs, i) if xs is not None else None)
        ys.append(y)
    return carry, (_tree_stack(ys) if ys and ys[0] is not None else None)


def _dynamic_slice(operand, start, sizes):
    sl = tuple(slice(int(s), int(s) + int(n)) for s, n in zip(start, sizes))
    return _wrap(np.asarray(operand)[sl])


def _make_jax():
    jax = types.ModuleType("jax")
    jnp = _NumpyModule("jax.numpy")
    jnp.pi, jnp.inf, jnp.nan, jnp.newaxis = np.pi, np.inf, np.nan, np.newaxis
    jnp.float64, jnp.complex128, jnp.int32, jnp.int64, jnp.bool_ = np.float64, np.complex128, np.int32, np.int64, np.bool_
    jnp.ndarray = np.ndarray
    jnp.linalg = types.ModuleType("jax.numpy.linalg")
    for name in ("cond", "slogdet", "eig", "det", "norm"):
        setattr(jnp.linalg, name, _wrapped(getattr(np.linalg, name)))

    def _nan_on_singular(fn):
        # jax returns non-finite values for a singular system; NumPy raises
        def call(a, *rest):
            try:
                return _wrap(fn(a, *rest))
            except np.linalg.LinAlgError:
                shape = np.shape(rest[0]) if rest else np.shape(a)
                return _wrap(np.full(shape, np.nan))

        return call

    jnp.linalg.solve = _nan_on_singular(np.linalg.solve)
    jnp.linalg.inv = _nan_on_singular(np.linalg.inv)

    lax = types.ModuleType("jax.lax")
    lax.fori_loop, lax.switch, lax.cond, lax.map, lax.scan = _fori_loop, _switch, _cond, _map, _scan
    lax.dynamic_slice = _dynamic_slice
    lax.stop_gradient = lambda x: x

    tree_util = types.ModuleType("jax.tree_util")
    tree_util.register_pytree_node = lambda *a, **k: None
    tree_util.register_pytree_node_class = lambda cls: cls

    jsp = types.ModuleType("jax.scipy")
    jsp.signal = types.ModuleType("jax.scipy.signal")
    for name in ("fftconvolve", "convolve2d", "convolve"):
        setattr(jsp.signal, name, _wrapped(getattr(scipy.signal, name)))
    jsp.special = types.ModuleType("jax.scipy.special")
    jsp.special.gamma = _wrapped(scipy.special.gamma)
    jsp.special.gammaln = _wrapped(scipy.special.gammaln)
    jsp.special.eval_hermite = _wrapped(scipy.special.eval_hermite)
    jsp.ndimage = types.ModuleType("jax.scipy.ndimage")
    def _map_coordinates(input, coordinates, order, mode="constant", cval=0.0):
        # jax.scipy.ndimage.map_coordinates (jax/_src/scipy/ndimage.py) replaces every out-of-range NEIGHBOUR by
        # cval and keeps interpolating: that is scipy's "grid-constant", not scipy's "constant"
        if mode == "constant":
            mode = "grid-constant"
        return scipy.ndimage.map_coordinates(np.asarray(input, dtype=float), [np.asarray(c) for c in coordinates],
                                             order=order, mode=mode, cval=cval)

    jsp.ndimage.map_coordinates = _map_coordinates

    experimental = types.ModuleType("jax.experimental")
    experimental.checkify = types.ModuleType("jax.experimental.checkify")

    class _Config:
        def update(self, *a, **k):
            pass

    jax.numpy, jax.lax, jax.tree_util, jax.scipy, jax.experimental = jnp, lax, tree_util, jsp, experimental
    jax.jit = _jit
    jax.custom_jvp = _CustomJvp
    jax.jvp = lambda *a, **k: (_ for _ in ()).throw(NotImplementedError("jvp not available in refshim"))
    jax.grad = jax.jvp
    jax.vmap = _vmap
    jax.config = _Config()
    jax.debug = types.SimpleNamespace(print=lambda *a, **k: None)
    jax.devices = lambda *a: [object()]
    # imported at module level by the lens-equation solver (PointSource/Types/base_ps.py pulls it in); never called
    jax.random = types.ModuleType("jax.random")
    jsp.optimize = types.ModuleType("jax.scipy.optimize")
    jsp.optimize.minimize = lambda *a, **k: (_ for _ in ()).throw(NotImplementedError("minimize not in refshim"))
    mods = {
        "jax.random": jax.random, "jax.scipy.optimize": jsp.optimize,
        "jax": jax, "jax.numpy": jnp, "jax.numpy.linalg": jnp.linalg, "jax.lax": lax, "jax.tree_util": tree_util,
        "jax.scipy": jsp, "jax.scipy.signal": jsp.signal, "jax.scipy.special": jsp.special,
        "jax.scipy.ndimage": jsp.ndimage, "jax.experimental": experimental,
        "jax.experimental.checkify": experimental.checkify,
    }
    return mods


# ------------------------------------------------------------------------------------------------
# lenstronomy / astropy shim
# ------------------------------------------------------------------------------------------------
class _Placeholder:
    """Inert stand-in for any lenstronomy/astropy name that is off the hot path."""

    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return _Placeholder()

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Placeholder()

    def __iter__(self):
        return iter(())

    def __mro_entries__(self, bases):
        return (object,)


class _StubModule(types.ModuleType):
    __all__ = []
    __path__ = []

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        sub = sys.modules.get(self.__name__ + "." + name)
        if sub is not None:
            return sub
        cls = type(name, (_Placeholder,), {})
        setattr(self, name, cls)
        return cls


class _StubFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    prefixes = ("lenstronomy", "astropy")

    def find_spec(self, fullname, path, target=None):
        if fullname.split(".")[0] in self.prefixes:
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        return None

    def create_module(self, spec):
        return _StubModule(spec.name)

    def exec_module(self, module):
        pass


def _make_lenstronomy():
    sys.path.insert(0, "/root/repo")
    from oracle import lenstronomy_port as port
    from oracle import util as outil
    from oracle.light_profiles import _bool_list

    def mod(name, **attrs):
        m = _StubModule(name)
        for k, v in attrs.items():
            setattr(m, k, v)
        return m

    class LensProfileBase(object):
        def __init__(self, *args, **kwargs):
            self._static = False

        def set_static(self, **kwargs):
            pass

        def set_dynamic(self):
            pass

    def exporter():
        all_ = []

        def decorator(obj):
            all_.append(obj.__name__)
            return obj

        return decorator, all_

    class Coordinates(object):
        def __init__(self, transform_pix2angle, ra_at_xy_0, dec_at_xy_0):
            self._Mpix2a = np.asarray(transform_pix2angle, dtype=float)
            self._Ma2pix = np.linalg.inv(self._Mpix2a)
            self._ra_at_xy_0, self._dec_at_xy_0 = ra_at_xy_0, dec_at_xy_0
            self._x_at_radec_0, self._y_at_radec_0 = port.map_coord2pix(-ra_at_xy_0, -dec_at_xy_0, 0, 0, self._Ma2pix)

        @property
        def transform_pix2angle(self):
            return self._Mpix2a

        @property
        def radec_at_xy_0(self):
            return self._ra_at_xy_0, self._dec_at_xy_0

        @property
        def pixel_area(self):
            return np.abs(np.linalg.det(self._Mpix2a))

        @property
        def pixel_width(self):
            return np.sqrt(self.pixel_area)

        def map_pix2coord(self, x, y):
            return port.map_coord2pix(x, y, self._ra_at_xy_0, self._dec_at_xy_0, self._Mpix2a)

        def map_coord2pix(self, ra, dec):
            return port.map_coord2pix(ra, dec, self._x_at_radec_0, self._y_at_radec_0, self._Ma2pix)

        def coordinate_grid(self, nx, ny):
            ra, dec = port.grid_from_coordinate_transform(nx, ny, self._Mpix2a, self._ra_at_xy_0, self._dec_at_xy_0)
            return outil.array2image(ra, nx, ny), outil.array2image(dec, nx, ny)

    class Coordinates1D(Coordinates):
        def coordinate_grid(self, nx, ny):
            return port.grid_from_coordinate_transform(nx, ny, self._Mpix2a, self._ra_at_xy_0, self._dec_at_xy_0)

    class PixelGrid(Coordinates):
        def __init__(self, nx, ny, transform_pix2angle, ra_at_xy_0, dec_at_xy_0, antenna_primary_beam=None):
            Coordinates.__init__(self, transform_pix2angle, ra_at_xy_0, dec_at_xy_0)
            self._nx, self._ny = nx, ny
            self._pb = antenna_primary_beam

        @property
        def num_pixel_axes(self):
            return self._nx, self._ny

        @property
        def num_pixel(self):
            return self._nx * self._ny

        @property
        def primary_beam(self):
            return self._pb

        @property
        def width(self):
            return self._nx * self.pixel_width, self._ny * self.pixel_width

        @property
        def center(self):
            x, y = Coordinates.coordinate_grid(self, self._nx, self._ny)
            return np.mean(x), np.mean(y)

    class PSF(port.PSF):
        def __init__(self, psf_type="NONE", fwhm=None, truncation=5, pixel_size=None, kernel_point_source=None,
                     psf_variance_map=None, point_source_supersampling_factor=1, kernel_point_source_init=None,
                     kernel_point_source_normalisation=True):
            port.PSF.__init__(self, psf_type, fwhm, truncation, pixel_size, kernel_point_source,
                              kernel_point_source_normalisation, point_source_supersampling_factor)
            self.point_source_supersampling_factor = point_source_supersampling_factor
            self.psf_variance_map_bool = psf_variance_map is not None
            # lenstronomy PSF.psf_variance_map: the given map (shape of kernel_point_source), zeros if none
            self.psf_variance_map = np.array(psf_variance_map, dtype=float) if psf_variance_map is not None \
                else np.zeros_like(self.kernel_point_source) if psf_type != "GAUSSIAN" else None

        def kernel_point_source_supersampled(self, supersampling_factor, updata_cache=True):
            if self.psf_type == "GAUSSIAN":
                n = 3
                k = np.zeros((n, n))
                k[1, 1] = 1
                return k
            return port.PSF.kernel_point_source_supersampled(self, supersampling_factor)

    class Image2SourceMapping(object):
        """lenstronomy.ImSim.image2source_mapping, single source plane."""

        def __init__(self, lens_model, source_model):
            self._lens_model, self._light_model = lens_model, source_model

        def image2source(self, x, y, kwargs_lens, index_source=None, kwargs_special=None):
            return self._lens_model.ray_shooting(x, y, kwargs_lens)

        def image_flux_joint(self, x, y, kwargs_lens, kwargs_source, kwargs_special=None, k=None):
            bx, by = self._lens_model.ray_shooting(x, y, kwargs_lens)
            return self._light_model.surface_brightness(bx, by, kwargs_source, k=k)

        def image_flux_split(self, x, y, kwargs_lens, kwargs_source, kwargs_special=None):
            bx, by = self._lens_model.ray_shooting(x, y, kwargs_lens)
            return self._light_model.functions_split(bx, by, kwargs_source)

    class DifferentialExtinction(object):
        def __init__(self, optical_depth_model=None, tau0_index=0, **kwargs):
            pass

        @property
        def compute_bool(self):  # image_model.py consults this; no extinction model on the path
            return False

    class Background(object):
        """lenstronomy.Cosmo.background.Background for a flat LCDM (matter + Lambda) stand-in cosmology.
        Only produces host constants; any consistent (T_xy, d_xy) pair exercises the same recursion."""

        def __init__(self, cosmo=None, interp=False, **kwargs):
            self.H0, self.Om0 = 70.0, 0.3

        def _chi(self, z):
            from scipy.integrate import quad

            c = 299792.458
            return c / self.H0 * quad(lambda zz: 1.0 / np.sqrt(self.Om0 * (1 + zz) ** 3 + 1 - self.Om0), 0, z,
                                      epsabs=1e-13, epsrel=1e-13)[0]

        def T_xy(self, z_observer, z_source):
            return np.vectorize(lambda a, b: self._chi(b) - self._chi(a))(z_observer, z_source) + 0.0

        def d_xy(self, z_observer, z_source):
            return self.T_xy(z_observer, z_source) / (1.0 + np.asarray(z_source))

        def ddt(self, z_lens, z_source):
            return (1 + z_lens) * self.d_xy(0, z_lens) * self.d_xy(0, z_source) / self.d_xy(z_lens, z_source)

    lutil = mod(
        "lenstronomy.Util.util",
        make_grid=port.make_grid, make_subgrid=lambda ra, dec, subgrid_res=2: port.make_subgrid(ra, dec, subgrid_res),
        array2image=port.array2image_lenstronomy, image2array=outil.image2array, fwhm2sigma=port.fwhm2sigma,
        convert_bool_list=lambda n, k=None: _bool_list(n, k),
        local_minima_2d=None, selectBest=None, map_coord2pix=port.map_coord2pix,
        make_grid_with_coordtransform=port.make_grid_with_coordtransform,
    )
    kutil = mod("lenstronomy.Util.kernel_util", cut_psf=port.cut_psf, split_kernel=port.split_kernel,
                degrade_kernel=port.degrade_kernel)
    mods = {
        "lenstronomy.Util.util": lutil,
        "lenstronomy.Util.kernel_util": kutil,
        "lenstronomy.Util.package_util": mod("lenstronomy.Util.package_util", exporter=exporter),
        "lenstronomy.Util.simulation_util": mod("lenstronomy.Util.simulation_util",
                                                data_configure_simple=port.data_configure_simple),
        "lenstronomy.LensModel.Profiles.base_profile": mod("lenstronomy.LensModel.Profiles.base_profile",
                                                           LensProfileBase=LensProfileBase),
        "lenstronomy.Data.coord_transforms": mod("lenstronomy.Data.coord_transforms", Coordinates=Coordinates,
                                                 Coordinates1D=Coordinates1D),
        "lenstronomy.Data.pixel_grid": mod("lenstronomy.Data.pixel_grid", PixelGrid=PixelGrid),
        "lenstronomy.Data.psf": mod("lenstronomy.Data.psf", PSF=PSF),
        "lenstronomy.ImSim.image2source_mapping": mod("lenstronomy.ImSim.image2source_mapping",
                                                      Image2SourceMapping=Image2SourceMapping),
        "lenstronomy.ImSim.differential_extinction": mod("lenstronomy.ImSim.differential_extinction",
                                                         DifferentialExtinction=DifferentialExtinction),
        "lenstronomy.Conf.config_loader": mod("lenstronomy.Conf.config_loader", conventions_conf=lambda: {}),
        "lenstronomy.Cosmo.background": mod("lenstronomy.Cosmo.background", Background=Background),
        "astropy.cosmology": mod("astropy.cosmology",
                                 default_cosmology=types.SimpleNamespace(get=lambda: "flat-LCDM(H0=70,Om0=0.3) stand-in")),
    }
    # Util package needs attribute access lenstronomy.Util.util / kernel_util via "from lenstronomy.Util import util"
    mods["lenstronomy.Util"] = mod("lenstronomy.Util", util=lutil, kernel_util=kutil,
                                   simulation_util=mods["lenstronomy.Util.simulation_util"])
    return mods


_installed = False


def _vmap(f, in_axes=0, out_axes=0):
    """jax.vmap for positional arguments: a Python loop over the mapped axis (in_axes: int or a tuple with None for
    broadcast arguments); outputs (array or tuple of arrays) are stacked along axis 0."""
    def mapped(*args):
        axes = in_axes if isinstance(in_axes, (tuple, list)) else (in_axes,) * len(args)
        n = next(np.shape(a)[ax] for a, ax in zip(args, axes) if ax is not None)
        outs = [f(*[a if ax is None else np.take(np.asarray(a), i, axis=ax) for a, ax in zip(args, axes)])
                for i in range(n)]
        if isinstance(outs[0], tuple):
            return tuple(_wrap(np.stack([np.asarray(o[j]) for o in outs])) for j in range(len(outs[0])))
        return _wrap(np.stack([np.asarray(o) for o in outs]))
    return mapped


def install():
    """Make ``import jaxtronomy`` (from /root/reference) work in this process."""
    global _installed
    if _installed:
        return
    sys.modules.update(_make_jax())
    sys.meta_path.append(_StubFinder())
    sys.modules.update(_make_lenstronomy())
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    np.seterr(all="ignore")
    _installed = True
