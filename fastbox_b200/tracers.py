"""
Biased tracers on top of a density field (reference ``fastbox/tracers.py``).
These are scalar functions of redshift; the field operation ``delta * bias`` is
fused into the last FFT pass on the device (``scale`` argument of the C ABI).
"""
import numpy as np

from .cosmology import get_backend

ccl = get_backend()


class TracerModel(object):

    def __init__(self, box):
        self.box = box

    def signal_amplitude(self, amp, redshift):
        """Constant amplitude model (tracers.py:25-41)."""
        return amp + 0. * redshift

    def linear_bias(self, b0, redshift):
        """b(z) = b0 sqrt(1 + z) (tracers.py:44-59)."""
        return b0 * np.sqrt(1. + redshift)


class HITracer(TracerModel):

    def __init__(self, box, OmegaHI0=0.000486, bHI0=0.677105):
        super().__init__(box)
        self.OmegaHI0 = OmegaHI0
        self.bHI0 = bHI0

    def signal_amplitude(self, redshift=None, formula='powerlaw'):
        """Brightness temperature Tb(z) in mK (tracers.py:88-126)."""
        if redshift is None:
            redshift = self.box.redshift
        z = redshift
        omegaHI = self.Omega_HI(redshift=redshift)
        if formula == 'powerlaw':
            Tb = 5.5919e-02 + 2.3242e-01 * z - 2.4136e-02 * z ** 2.
        elif formula == 'hall':
            E = ccl.h_over_h0(self.box.cosmo, 1. / (1. + z))
            Tb = 188. * self.box.cosmo['h'] * omegaHI * (1. + z) ** 2. / E
        else:
            raise ValueError("No formula found with name '%s'" % formula)
        return Tb

    def bias_HI(self, redshift=None):
        """HI bias fitting formula (tracers.py:129-144)."""
        if redshift is None:
            redshift = self.box.redshift
        z = redshift
        return (self.bHI0 / 0.677105) * (6.6655e-01 + 1.7765e-01 * z + 5.0223e-02 * z ** 2.)

    def Omega_HI(self, redshift=None, formula='powerlaw'):
        """Fractional HI density fitting formula (tracers.py:147-163)."""
        if redshift is None:
            redshift = self.box.redshift
        z = redshift
        return (self.OmegaHI0 / 0.000486) * (4.8304e-04 + 3.8856e-04 * z - 6.5119e-05 * z ** 2.)
