"""
Biased tracers on top of a density field (reference ``fastbox/tracers.py``).
These are scalar functions of redshift; the field operation ``delta * bias`` is
fused into the last FFT pass on the device (``scale`` argument of the C ABI).
"""
import numpy as np

from .cosmology import get_backend

ccl = get_backend()


def _quadratic(z, c0, c1, c2):
    """c0 + c1 z + c2 z^2, evaluated left to right like the reference's fitting formulae."""
    return c0 + c1 * z + c2 * z ** 2.


class TracerModel(object):
    """Constant-amplitude signal and b(z) = b0 sqrt(1+z) bias (tracers.py:11-59)."""

    def __init__(self, box):
        self.box = box

    def _z(self, redshift):
        return self.box.redshift if redshift is None else redshift

    def signal_amplitude(self, amp, redshift):
        return amp + 0. * redshift                        # shaped like `redshift` (tracers.py:41)

    def linear_bias(self, b0, redshift):
        return b0 * np.sqrt(1. + redshift)                # tracers.py:59


class HITracer(TracerModel):
    """Neutral-hydrogen intensity-mapping tracer: Tb(z), b_HI(z), Omega_HI(z) fits (tracers.py:62-163)."""

    TB_FIT = (5.5919e-02, 2.3242e-01, -2.4136e-02)        # mK, tracers.py:116
    BIAS_FIT = (6.6655e-01, 1.7765e-01, 5.0223e-02)       # tracers.py:144
    OMEGA_FIT = (4.8304e-04, 3.8856e-04, -6.5119e-05)     # tracers.py:163
    BIAS_REF, OMEGA_REF = 0.677105, 0.000486              # normalisation of the two fits at z = 0

    def __init__(self, box, OmegaHI0=0.000486, bHI0=0.677105):
        super().__init__(box)
        self.OmegaHI0, self.bHI0 = OmegaHI0, bHI0

    def signal_amplitude(self, redshift=None, formula='powerlaw'):
        """Mean brightness temperature in mK."""
        z = self._z(redshift)
        if formula == 'powerlaw':
            return _quadratic(z, *self.TB_FIT)
        if formula == 'hall':                             # Hall et al. form, tracers.py:118-121
            E = ccl.h_over_h0(self.box.cosmo, 1. / (1. + z))
            return 188. * self.box.cosmo['h'] * self.Omega_HI(redshift=z) * (1. + z) ** 2. / E
        raise ValueError("No formula found with name '%s'" % formula)

    def bias_HI(self, redshift=None):
        return (self.bHI0 / self.BIAS_REF) * _quadratic(self._z(redshift), *self.BIAS_FIT)

    def Omega_HI(self, redshift=None, formula='powerlaw'):
        return (self.OmegaHI0 / self.OMEGA_REF) * _quadratic(self._z(redshift), *self.OMEGA_FIT)
