"""jaxtronomy_b200 -- B200-native forward image simulation for strong lensing.

A drop-in for the forward-simulation hot path of JAXtronomy/lenstronomy (ray shooting -> light on the
supersampled grid -> re-bin -> PSF convolution -> linear amplitudes + likelihood), batched over parameter
sets, running as hand-written sm_100a CUDA kernels behind the C ABI of include/lsim.h.  The Python classes
keep the reference's names and signatures (LensModel.ray_shooting, LightModel.surface_brightness,
ImageModel.image, ImageLinearFit.likelihood_data_given_model); there is no CPU fallback.
"""

__version__ = "0.1.0"

from jaxtronomy_b200 import config  # noqa: E402,F401
