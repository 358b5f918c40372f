"""Global numerical configuration, after ``jax.config.update("jax_enable_x64", ...)`` which is how a JAXtronomy user
chooses between double and single precision.

    import jaxtronomy_b200
    jaxtronomy_b200.config.update("enable_x64", False)   # opt into the fp32 image kernel

``enable_x64`` is True by default: every kernel computes in fp64 and results match the reference to rtol 1e-10.  With
False, image plans created AFTERWARDS run the fused ray-shooting + light + re-bin kernel in fp32 (stated tolerance
2e-4 on images; PSF convolution, linear solve and chi^2 stay fp64).  Models outside the fp32 kernel's scope raise
LsimError at plan creation; there is no silent fallback in either direction.  Ray shooting and surface brightness on
arbitrary points (LensModel.ray_shooting, LightModel.surface_brightness) are always fp64.
"""

_values = {"enable_x64": True}


def update(name, value):
    if name not in _values:
        raise KeyError("unknown option %r (known: %s)" % (name, ", ".join(sorted(_values))))
    _values[name] = bool(value)


def read(name):
    return _values[name]
