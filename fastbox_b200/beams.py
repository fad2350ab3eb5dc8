"""
Instrumental beams (reference ``fastbox/beams.py``): the FFT beam convolution of
``BeamModel.convolve_fft`` (beams.py:63-87) runs on the device.
"""
import numpy as np

from . import _lib


class BeamModel(object):

    def __init__(self, box):
        self.box = box

    def beam_cube(self, pol=None):
        """Beam value at each voxel; default is a unit beam (beams.py:26-38)."""
        return np.ones((self.box.N, self.box.N, self.box.N))

    def beam_value(self, x, y, freq, pol=None):
        assert x.shape == y.shape == freq.shape, \
            "x, y, and freq arrays should have the same shape"
        return 1. + 0. * x

    def convolve_fft(self, field_x, pol=None):
        """
        Per-channel zero-padded 2-D convolution of ``field_x`` with the beam,
        divided by the per-channel beam sum (beams.py:81-87), on the GPU.
        """
        box = self.box
        N = box.N
        plan = box._plan
        beam = self.beam_cube(pol=pol)
        # the transform of the beam cube is kept by the plan (fb_beam_set) and reused while the cube that
        # beam_cube() returns stays the same (content signature: two BLAS-speed reductions of the host array)
        if isinstance(beam, _lib.DeviceBuffer):
            sig, d_beam = ("device", beam.ptr), beam
        else:
            beam = np.asarray(beam)
            sig, d_beam = ("host", beam.shape, box._field_signature(beam)), None
        if getattr(plan, "beam_sig", None) != sig:
            plan.beam_sig = None
            plan.beam_set(d_beam if d_beam is not None else plan.upload_f32(beam))
            plan.beam_sig = sig
        d_field = box._to_device_field(field_x)
        out = plan.alloc(N ** 3 * 4)
        plan.beam_convolve(None, d_field, out)
        return plan.download_f64(out, (N, N, N))


class GaussianBeamModel(BeamModel):
    """
    EXTENSION: frequency-dependent Gaussian beam, FWHM(nu) = 1.22 lambda / D
    (dish diameter as in forecast.py:16), evaluated on ``box.pixel_array()`` /
    ``box.freq_array()``; used by the headline benchmark configuration.
    """

    def __init__(self, box, dish_diameter=13.5):
        super().__init__(box)
        self.D = dish_diameter

    def beam_cube(self, pol=None):
        box = self.box
        ang_x, ang_y = box.pixel_array()
        freqs = box.freq_array()
        lam = 299792458. / (freqs * 1e6)
        fwhm = np.degrees(1.22 * lam / self.D)
        sig = fwhm / 2.3548200450309493
        r2 = ang_x[:, None, None] ** 2 + ang_y[None, :, None] ** 2
        return np.exp(-0.5 * r2 / sig[None, None, :] ** 2)
