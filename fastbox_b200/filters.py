"""
Foreground filters (reference ``fastbox/filters.py``): ``mean_spectrum_filter``
(filters.py:35-55) and ``pca_filter`` (filters.py:93-183).  The N^3 steps run on the device;
the PCA filter works in float64 throughout (foreground-dominated cubes are
not a float32 problem) and leaves the N_f x N_f eigen-decomposition to NumPy on the host.
The ICA / GPR / band-pass filters are not part of this package.
"""
import numpy as np

from . import _lib

_plans = {}


def _plan_for(N, device=0):
    if (N, device) not in _plans:
        _plans[(N, device)] = _lib.Plan(N, 1.0, 1.0, 1.0, device)
    return _plans[(N, device)]


def mean_spectrum_filter(field, return_mean=False, device=0):
    """
    Subtract the mean from each frequency slice (filters.py:35-55); the 3rd axis is frequency.
    ``field`` must be a cube (N, N, N) with N a power of two in [8, 2048].  Returns float64 like
    the reference (computed in float32 on the device, the means in float64).
    """
    field = np.asarray(field)
    if field.ndim != 3 or not (field.shape[0] == field.shape[1] == field.shape[2]):
        raise ValueError("mean_spectrum_filter: field must have shape (N, N, N)")
    N = field.shape[0]
    plan = _plan_for(N, device)
    d_in = plan.upload_f32(field)
    d_out = plan.alloc(N ** 3 * 4)
    mean = plan.mean_spectrum_filter(d_in, d_out)
    out = plan.download_f64(d_out, (N, N, N))
    return (out, mean) if return_mean else out


def pca_filter(field, nmodes, fit_powerlaw=False, return_filter=False, device=0):
    """
    PCA foreground filter (filters.py:93-183): subtract the ``nmodes`` leading eigenmodes of the
    frequency-frequency covariance (plus the mean spectrum) from every line of sight.

    Device, float64: mean spectrum, covariance (``np.cov`` normalisation) and the projection.
    Host: ``np.linalg.eigh`` of the (N, N) covariance (the reference calls ``np.linalg.eig``; the
    eigenvectors agree up to sign, which cancels in ``cleaned_field``; ``U_fg`` / ``fg_amps`` may
    differ from the reference's by a sign per mode).  ``fit_powerlaw`` fits the mean spectrum with
    ``scipy.optimize.curve_fit`` exactly as the reference does (filters.py:146-155).
    """
    field = np.asarray(field, dtype=np.float64)
    if field.ndim != 3 or not (field.shape[0] == field.shape[1] == field.shape[2]):
        raise ValueError("pca_filter: field must have shape (N, N, N)")
    N = field.shape[0]
    nmodes = int(nmodes)
    if not 1 <= nmodes <= min(32, N):
        raise ValueError("pca_filter: nmodes must be in [1, 32]")
    plan = _plan_for(N, device)
    d_cube = plan.upload(np.ascontiguousarray(field))
    d_mean, cov = plan.pca_covariance(d_cube)
    if fit_powerlaw:
        from scipy.optimize import curve_fit
        freqs = np.linspace(1., 10., N)

        def fn(nu, amp, beta):
            return amp * (nu / nu[0]) ** beta
        pfit, _ = curve_fit(fn, freqs, d_mean, p0=[d_mean[0], -2.7])
        d_mean = fn(freqs, pfit[0], pfit[1])
    eigvals, eigvecs = np.linalg.eigh(cov)
    idxs = np.argsort(eigvals)[::-1]                     # biggest eigenvalue first (filters.py:165-167)
    U_fg = np.ascontiguousarray(eigvecs[:, idxs][:, :nmodes])
    d_clean = plan.alloc(N ** 3 * 8)
    d_amps = plan.alloc(nmodes * N * N * 8) if return_filter else None
    plan.pca_project(d_cube, d_mean, U_fg, d_clean, d_amps)
    cleaned = plan.download(d_clean, (N, N, N), np.float64)
    if return_filter:
        return cleaned, U_fg, plan.download(d_amps, (nmodes, N * N), np.float64)
    return cleaned
