"""
fastbox_b200 -- B200-native (sm_100a) implementation of the FastBox field-generation
hot path behind the reference's Python API (``CosmoBox``, ``BeamModel.convolve_fft``,
``HaloDistribution.halo_count_field``, ``TracerModel`` / ``HITracer``).

Python host code calls ``libfastbox_b200.so`` (C ABI, ``include/fastbox_b200.h``)
through ctypes; there is no CPU fallback.
"""
from .box import CosmoBox, default_cosmo  # noqa: F401
from . import beams, filters, foregrounds, halos, noise, tracers  # noqa: F401

__version__ = "0.1.0"
