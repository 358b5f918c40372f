"""Host-side packing of the reference's kwargs lists into the descriptors of include/lsim.h.

The reference passes parameters as lists of dicts of scalars and batches them with ``vmap``
(Sampling/sampler.py:241-246).  Here a batch is a dense ``[B, P]`` fp64 matrix plus, per profile
parameter, either a column index or a constant -- the same information ``Param.args2kwargs``
(Sampling/likelihood.py:305) encodes.  Parameter names and order are the reference classes'
``param_names``.
"""

import numpy as np

from . import _lib

LENS_PARAM_NAMES = {
    "EPL": ["theta_E", "gamma", "e1", "e2", "center_x", "center_y"],
    "SIE": ["theta_E", "e1", "e2", "center_x", "center_y"],
    "NIE": ["theta_E", "e1", "e2", "s_scale", "center_x", "center_y"],
    "SIS": ["theta_E", "center_x", "center_y"],
    "SHEAR": ["gamma1", "gamma2", "ra_0", "dec_0"],
    "CONVERGENCE": ["kappa", "ra_0", "dec_0"],
    "NFW": ["Rs", "alpha_Rs", "center_x", "center_y"],
    "TNFW": ["Rs", "alpha_Rs", "r_trunc", "center_x", "center_y"],
    "MULTIPOLE": ["m", "a_m", "phi_m", "center_x", "center_y", "r_E"],
    "EPL_MULTIPOLE_M1M3M4": ["theta_E", "gamma", "e1", "e2", "center_x", "center_y", "a1_a", "delta_phi_m1",
                             "a3_a", "delta_phi_m3", "a4_a", "delta_phi_m4"],
    "CSE": ["a", "s", "e1", "e2", "center_x", "center_y"],
    "NFW_ELLIPSE_CSE": ["Rs", "alpha_Rs", "e1", "e2", "center_x", "center_y"],
    "HERNQUIST_ELLIPSE_CSE": ["sigma0", "Rs", "e1", "e2", "center_x", "center_y"],
    "NULL": [],
}
LIGHT_PARAM_NAMES = {
    "SERSIC": ["amp", "R_sersic", "n_sersic", "center_x", "center_y"],
    "SERSIC_ELLIPSE": ["amp", "R_sersic", "n_sersic", "e1", "e2", "center_x", "center_y"],
    "GAUSSIAN": ["amp", "sigma", "center_x", "center_y"],
}
# defaults of the reference's function signatures (e.g. epl.py:179, shear.py:39, multipole.py:72)
PARAM_DEFAULTS = {"center_x": 0.0, "center_y": 0.0, "ra_0": 0.0, "dec_0": 0.0, "r_E": 1.0}

LOWER_LIMIT_DEFAULT = {
    "EPL": {"theta_E": 0, "gamma": 1.5, "e1": -0.5, "e2": -0.5, "center_x": -100, "center_y": -100},
    "SIE": {"theta_E": 0, "e1": -0.5, "e2": -0.5, "center_x": -100, "center_y": -100},
    "SIS": {"theta_E": 0, "center_x": -100, "center_y": -100},
    "SHEAR": {"gamma1": -0.5, "gamma2": -0.5, "ra_0": -100, "dec_0": -100},
}
UPPER_LIMIT_DEFAULT = {
    "EPL": {"theta_E": 100, "gamma": 2.5, "e1": 0.5, "e2": 0.5, "center_x": 100, "center_y": 100},
    "SIE": {"theta_E": 100, "e1": 0.5, "e2": 0.5, "center_x": 100, "center_y": 100},
    "SIS": {"theta_E": 100, "center_x": 100, "center_y": 100},
    "SHEAR": {"gamma1": 0.5, "gamma2": 0.5, "ra_0": 100, "dec_0": 100},
}


def _lens_flags_opts(name, profile_kwargs):
    """The reference's profile init options (profile_list_base.py:161-232) -> (flags, opt[2])."""
    pk = dict(profile_kwargs or {})
    flags, opt = 0, [0.0, 0.0]
    if name == "SIE":
        opt = [float(pk.pop("s_scale", 0.0000000001)), float(pk.pop("gamma", 2))]
        if not pk.pop("NIE", True):
            flags |= _lib.FLAG_SIE_NIE_FALSE
    elif name == "CSE":
        axis = pk.pop("axis", "product_avg")
        if axis not in ("major", "product_avg"):
            raise ValueError("axis must be set to 'major' or 'product_avg'. Input is %s ." % axis)
        if axis == "major":
            flags |= _lib.FLAG_CSE_MAJOR
    elif name == "NFW_ELLIPSE_CSE":
        if not pk.pop("high_accuracy", True):
            flags |= _lib.FLAG_CSE_LOW_ACCURACY
    elif name == "NFW":
        if pk.pop("interpol", False):
            raise Exception("This class no longer supports interpol functionality in JAXtronomy.")
    if pk:
        raise ValueError("unsupported profile_kwargs %s for lens model %s" % (sorted(pk), name))
    return flags, opt


def lens_desc(name, sources, profile_kwargs=None, delta_T=0.0, T_z=0.0, red2phys=0.0):
    """``sources``: per parameter (param_names order) either ("col", j) or ("cst", value)."""
    if name not in _lib.LENS_TYPES:
        raise ValueError("%s is not a valid lens model. Supported are: %s." % (name, sorted(_lib.LENS_TYPES)))
    d = _lib.LensDesc()
    d.type = _lib.LENS_TYPES[name]
    flags, opt = _lens_flags_opts(name, profile_kwargs)
    d.flags = flags
    d.opt[0], d.opt[1] = opt
    for k in range(_lib.MAX_PARAMS):
        d.col[k] = -1
        d.cst[k] = 0.0
    for k, (kind, v) in enumerate(sources):
        if kind == "col":
            d.col[k] = int(v)
        else:
            d.cst[k] = float(v)
    d.delta_T, d.T_z, d.red2phys = float(delta_T), float(T_z), float(red2phys)
    return d


def light_desc(name, sources, lensed, slot, use_amp):
    if name not in _lib.LIGHT_TYPES:
        raise ValueError("No light model of type %s found! Supported are %s" % (name, sorted(_lib.LIGHT_TYPES)))
    d = _lib.LightDesc()
    d.type = _lib.LIGHT_TYPES[name]
    d.lensed, d.slot, d.use_amp = int(bool(lensed)), int(slot), int(bool(use_amp))
    for k in range(_lib.MAX_PARAMS):
        d.col[k] = -1
        d.cst[k] = 0.0
    for k, (kind, v) in enumerate(sources):
        if kind == "col":
            d.col[k] = int(v)
        else:
            d.cst[k] = float(v)
    return d


class ColumnLayout:
    """Every parameter of every profile gets its own column: the plan then depends on the model structure
    only and can be cached, whatever values (scalars or [B] arrays) a call passes."""

    def __init__(self):
        self.names = []  # (component, profile index, param name)

    def add_profile(self, component, index, param_names):
        start = len(self.names)
        self.names += [(component, index, n) for n in param_names]
        return [("col", start + k) for k in range(len(param_names))]

    @property
    def n_cols(self):
        return len(self.names)


def _is_torch(v):
    return type(v).__module__.startswith("torch")


def pack_columns(values):
    """values: list of scalars / [B] arrays (numpy or torch) -> ([B, P] numpy array or torch tensor, B, batched?)."""
    B = 1
    batched = False
    any_torch = False
    for v in values:
        if _is_torch(v):
            any_torch = True
            n = v.numel() if v.ndim > 0 else 1
            if v.ndim > 0:
                batched = True
        else:
            a = np.asarray(v)
            n = a.size if a.ndim > 0 else 1
            if a.ndim > 0:
                batched = True
        if n != 1:
            if B != 1 and n != B:
                raise ValueError("inconsistent batch sizes among parameters: %d vs %d" % (B, n))
            B = n
    if any_torch:
        import torch

        dev = next(v.device for v in values if _is_torch(v))
        cols = [(v.to(device=dev, dtype=torch.float64).reshape(-1) if _is_torch(v)
                 else torch.as_tensor(np.asarray(v, dtype=np.float64).reshape(-1), device=dev)).expand(B)
                for v in values]
        mat = torch.stack(cols, dim=1) if cols else torch.zeros((B, 0), dtype=torch.float64, device=dev)
    else:
        mat = np.empty((B, len(values)), dtype=np.float64)
        for j, v in enumerate(values):
            mat[:, j] = np.asarray(v, dtype=np.float64).reshape(-1)
    return mat, B, batched


def gather_values(param_names_list, kwargs_list, what="lens"):
    """kwargs list (one dict per profile) -> flat list of values in layout order, with the reference's checks."""
    if len(kwargs_list) != len(param_names_list):
        raise ValueError("length of input parameter list %s does not match number of %s models %s"
                         % (len(kwargs_list), what, len(param_names_list)))
    values = []
    for i, names in enumerate(param_names_list):
        kw = kwargs_list[i]
        for key in kw:
            if key not in names:
                raise ValueError("parameter %s in %s model is not part of model %s. Parameters allowed are %s"
                                 % (key, what, i, names))
        for n in names:
            if n in kw:
                values.append(kw[n])
            elif n in PARAM_DEFAULTS:
                values.append(PARAM_DEFAULTS[n])
            else:
                raise ValueError("%s model %s requires parameter %s which is not provided in input." % (what, i, n))
    return values
