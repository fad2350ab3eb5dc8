"""
Foreground filters (reference ``fastbox/filters.py``).  Only ``mean_spectrum_filter``
(filters.py:35-55) is provided so far: per-channel mean over the pixels (float64 sums on the
device) subtracted from the cube.  The PCA / ICA / GPR filters need a float64 data path
 and are not part of this package yet.
"""
import numpy as np

from . import _lib

_plans = {}


def _plan_for(N):
    if N not in _plans:
        _plans[N] = _lib.Plan(N, 1.0, 1.0, 1.0)
    return _plans[N]


def mean_spectrum_filter(field, return_mean=False):
    """
    Subtract the mean from each frequency slice (filters.py:35-55); the 3rd axis is frequency.
    ``field`` must be a cube (N, N, N) with N a power of two in [8, 2048].  Returns float64 like
    the reference (computed in float32 on the device, the means in float64).
    """
    field = np.asarray(field)
    if field.ndim != 3 or not (field.shape[0] == field.shape[1] == field.shape[2]):
        raise ValueError("mean_spectrum_filter: field must have shape (N, N, N)")
    N = field.shape[0]
    plan = _plan_for(N)
    d_in = plan.upload_f32(field)
    d_out = plan.alloc(N ** 3 * 4)
    mean = plan.mean_spectrum_filter(d_in, d_out)
    out = plan.download_f64(d_out, (N, N, N))
    return (out, mean) if return_mean else out
