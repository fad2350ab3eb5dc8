"""
Diffuse foregrounds (reference ``fastbox/foregrounds.py``, class ``ForegroundModel``).
The N^3 step, ``construct_cube`` (foregrounds.py:152-174), runs on the device
(``fb_fg_cube``).  The 2-D maps (``realise_foreground_amp``, ``realise_spectral_index``:
N^2 values, foregrounds.py:48-149) are evaluated with the reference's own NumPy / SciPy
calls on the host, so they reproduce the reference bit for bit under ``np.random.seed``.
"""
import numpy as np
import scipy.ndimage
from numpy import fft

from . import _lib
from .cosmology import get_backend


class ForegroundModel(object):

    def __init__(self, box):
        self.box = box

    def realise_foreground_amp(self, amp, beta, monopole, smoothing_scale=None, redshift=None):
        """2-D Gaussian random amplitude map with C_ell = amp (ell/1000)^beta (foregrounds.py:48-113)."""
        box = self.box
        ccl = get_backend()
        if redshift is None:
            redshift = box.redshift
        scale_factor = 1. / (1. + redshift)
        r = ccl.comoving_angular_distance(box.cosmo, scale_factor)
        k_perp = 2. * np.pi * np.sqrt((box.Kx[:, :, 0] / box.Lx) ** 2. + (box.Ky[:, :, 0] / box.Ly) ** 2.)
        with np.errstate(divide="ignore"):
            C_ell = amp * (0.5 * k_perp * r / 1000.) ** (beta)
        C_ell[np.isinf(C_ell)] = 0.
        C_ell *= (box.N ** 4.) / (box.Lx * box.Ly)
        re = np.random.normal(0.0, 1.0, np.shape(k_perp))
        im = np.random.normal(0.0, 1.0, np.shape(k_perp))
        fg_k = (re + 1.j * im) * np.sqrt(C_ell)
        fg_k[k_perp == 0.] = 0.
        fg_x = fft.ifftn(fg_k).real + monopole
        if smoothing_scale is not None:
            ang_x, ang_y = box.pixel_array(redshift=redshift)
            sigma = smoothing_scale / (ang_x[1] - ang_x[0])
            fg_x = scipy.ndimage.gaussian_filter(fg_x, sigma=sigma, mode='wrap')
        return fg_x

    def realise_spectral_index(self, mean_spec_idx, std_spec_idx, smoothing_scale, redshift=None):
        """Smoothed Gaussian random spectral-index map (foregrounds.py:116-149)."""
        box = self.box
        alpha = np.random.normal(mean_spec_idx, std_spec_idx, box.Kx[:, :, 0].shape)
        ang_x, ang_y = box.pixel_array(redshift=redshift)
        sigma = smoothing_scale / (ang_x[1] - ang_x[0])
        return scipy.ndimage.gaussian_filter(alpha, sigma=sigma, mode='wrap')

    def construct_cube(self, amps, spectral_idx, freq_ref=130., redshift=None, add_to=None):
        """
        cube[x,y,z] = amps[x,y] (freqs[z]/freq_ref)^spectral_idx[x,y] (foregrounds.py:152-174), float32
        on the device, returned as float64.  ``add_to`` (extension): an (N,N,N) cube the foregrounds
        are added to in the same pass.
        """
        box = self.box
        N = box.N
        plan = box._plan
        freqs = box.freq_array(redshift=redshift)
        l2 = np.log2(freqs / freq_ref)
        amps = np.asarray(amps, dtype=np.float64)
        if amps.shape != (N, N):
            raise ValueError("amps must have shape (N, N)")
        idx = spectral_idx if np.ndim(spectral_idx) == 0 else np.asarray(spectral_idx, dtype=np.float64)
        if np.ndim(idx) != 0 and idx.shape != (N, N):
            raise ValueError("spectral_idx must be a float or an (N, N) array")
        if add_to is None:
            out = plan.alloc(N ** 3 * 4)
            plan.fg_cube(amps, idx, l2, out)
        else:
            out = plan.upload_f32(add_to) if not isinstance(add_to, _lib.DeviceBuffer) else add_to
            plan.fg_cube(amps, idx, l2, out, accumulate=True)
        return plan.download_f64(out, (N, N, N))
