"""
Instrumental noise on top of a simulated cube (reference ``fastbox/noise.py``).
The N^3 step of ``NoiseModel.realise_radiometer_noise`` (noise.py:71-75) runs on the
device (``fb_radiometer_noise``); the per-channel rms (noise.py:55-69) is N numbers.
"""
import numpy as np

from . import _lib


class NoiseModel(object):

    def __init__(self, box):
        self.box = box

    def radiometer_rms(self, Tinst, tp, fov, Ndish, redshift=None):
        """Noise rms per frequency channel in mK (noise.py:55-69)."""
        freqs = self.box.freq_array(redshift=redshift)
        dnu = np.abs(freqs[1] - freqs[0])
        tp = tp * 3600.                                   # hrs to sec (the reference rescales its argument)
        ang_x, ang_y = self.box.pixel_array(redshift=redshift)
        dtheta = ang_x[1] - ang_x[0]
        t_res = tp * dtheta ** 2. / fov
        Tsky = 60e3 * (freqs / 300.) ** (-2.5)            # mK
        Tsys = Tinst * 1e3 + Tsky
        return Tsys / np.sqrt(Ndish * t_res * (dnu * 1e6))

    def realise_radiometer_noise(self, Tinst, tp, fov, Ndish, redshift=None, seed=None, normals=None):
        """
        White noise from the radiometer equation, in mK (noise.py:25-75).

        By default the unit normals are ``np.random.normal(0, 1, (N, N, N))`` exactly
        as the reference draws them (noise.py:73), so the cube equals the reference's
        under the same ``np.random.seed`` (to float32).  ``seed=`` draws them on the
        device instead (Philox4x32-10, nothing crosses the bus); ``normals=`` supplies
        them.  Returns float64 like the reference.
        """
        box = self.box
        N = box.N
        plan = box._plan
        sigma = self.radiometer_rms(Tinst, tp, fov, Ndish, redshift=redshift)
        out = plan.alloc(N ** 3 * 4)
        if seed is not None:
            plan.radiometer_noise(sigma, out, None, seed=int(seed))
        else:
            if normals is None:
                normals = np.random.normal(0., 1., (N, N, N))
            n32 = np.ascontiguousarray(normals, dtype=np.float32)
            if n32.shape != (N, N, N):
                raise ValueError("normals must have shape (N, N, N)")
            plan.radiometer_noise(sigma, out, n32)
        return plan.download_f64(out, (N, N, N))
