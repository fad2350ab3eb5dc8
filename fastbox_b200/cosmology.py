"""
Cosmology provider used in place of ``pyccl`` when CCL is not installed.

The reference obtains every background/power-spectrum quantity from CCL
(``/root/reference/fastbox/box.py:62,163,165,280-281,345,406,781,820,851``).
CCL is a compiled third-party library that is absent from this image, so this
module supplies the same *call signatures* backed by closed-form fits:

* ``linear_matter_power`` / ``nonlin_matter_power`` -- Eisenstein & Hu (1998)
  zero-baryon ("no-wiggle") transfer function, normalised to ``sigma8`` with a
  real-space top-hat of 8 Mpc/h, scaled by the linear growth factor.  No
  halofit correction is applied, so the two functions coincide (documented in
  DESIGN.md as "parity unpinned w.r.t. CCL"; field/P(k) parity never depends on
  it because oracle and GPU path consume the *same* provider).
* ``h_over_h0``, ``growth_factor``, ``growth_rate``,
  ``comoving_angular_distance`` -- flat LCDM quadratures.

If the real ``pyccl`` is importable it is preferred (see ``get_backend``).
"""
import numpy as np

__all__ = ["Cosmology", "linear_matter_power", "nonlin_matter_power",
           "h_over_h0", "growth_factor", "growth_rate",
           "comoving_angular_distance", "get_backend"]

_C_KMS = 299792.458  # km/s


class Cosmology(dict):
    """Dictionary-like cosmology (supports ``cosmo['h']`` as box.py:280 needs)."""

    def __init__(self, **kw):
        kw.setdefault("T_CMB", 2.725)
        tf = kw.get("transfer_function", "eisenstein_hu")
        if tf not in (None, "eisenstein_hu", "eisenstein_hu_nowiggles"):
            raise ValueError("built-in cosmology provider: transfer_function=%r is not available (only the "
                             "Eisenstein-Hu fit); install pyccl for Boltzmann-code spectra" % (tf,))
        super().__init__(**kw)
        self._cache = {}

    # -- derived -----------------------------------------------------------
    @property
    def Omega_m(self):
        return self["Omega_c"] + self["Omega_b"]

    def _norm(self):
        """Amplitude A such that sigma_8(z=0) == self['sigma8']."""
        if "A" not in self._cache:
            k = np.logspace(-5.0, 2.5, 20000)             # Mpc^-1
            R = 8.0 / self["h"]
            x = k * R
            w = 3.0 * (np.sin(x) - x * np.cos(x)) / x ** 3
            integrand = k ** 3 * _pk_shape(self, k) * w ** 2 / (2.0 * np.pi ** 2)
            lnk = np.log(k)
            var = np.sum(0.5 * (integrand[1:] + integrand[:-1]) * np.diff(lnk))
            self._cache["A"] = self["sigma8"] ** 2 / var
        return self._cache["A"]


def _eh_nowiggle_transfer(cosmo, k):
    """EH98 eqs. 26, 28-31; k in Mpc^-1."""
    h = cosmo["h"]
    om = cosmo.Omega_m
    omh2 = om * h * h
    obh2 = cosmo["Omega_b"] * h * h
    fb = cosmo["Omega_b"] / om
    th = cosmo["T_CMB"] / 2.7
    s = 44.5 * np.log(9.83 / omh2) / np.sqrt(1.0 + 10.0 * obh2 ** 0.75)  # Mpc
    ag = 1.0 - 0.328 * np.log(431.0 * omh2) * fb + 0.38 * np.log(22.3 * omh2) * fb ** 2
    gam = om * h * (ag + (1.0 - ag) / (1.0 + (0.43 * k * s) ** 4))
    q = (k / h) * th * th / gam
    L0 = np.log(2.0 * np.e + 1.8 * q)
    C0 = 14.2 + 731.0 / (1.0 + 62.5 * q)
    return L0 / (L0 + C0 * q * q)


def _pk_shape(cosmo, k):
    return k ** cosmo["n_s"] * _eh_nowiggle_transfer(cosmo, k) ** 2


def h_over_h0(cosmo, a):
    """E(a) = H(a)/H0 for flat LCDM (matter + Lambda)."""
    a = np.asarray(a, dtype=np.float64)
    om = cosmo.Omega_m
    return np.sqrt(om / a ** 3 + (1.0 - om))


def _growth_unnorm(cosmo, a):
    a = float(a)
    om = cosmo.Omega_m
    x = np.linspace(1e-6, a, 4097)
    e = np.sqrt(om / x ** 3 + (1.0 - om))
    f = 1.0 / (x * e) ** 3
    integ = np.sum(0.5 * (f[1:] + f[:-1]) * np.diff(x))
    return 2.5 * om * float(h_over_h0(cosmo, a)) * integ


def growth_factor(cosmo, a):
    """Linear growth D(a), normalised to D(1) = 1."""
    av = np.atleast_1d(np.asarray(a, dtype=np.float64))
    d1 = _growth_unnorm(cosmo, 1.0)
    out = np.array([_growth_unnorm(cosmo, x) / d1 for x in av])
    return out if np.ndim(a) else float(out[0])


def growth_rate(cosmo, a):
    """f = dlnD/dlna ~ Omega_m(a)^0.55."""
    a = np.asarray(a, dtype=np.float64)
    om = cosmo.Omega_m
    oma = om / a ** 3 / (om / a ** 3 + (1.0 - om))
    out = oma ** 0.55
    return out if np.ndim(a) else float(out)


def comoving_angular_distance(cosmo, a):
    """Comoving distance in Mpc (flat universe)."""
    av = np.atleast_1d(np.asarray(a, dtype=np.float64))
    res = []
    for x in av:
        aa = np.linspace(x, 1.0, 4097)
        f = 1.0 / (aa ** 2 * h_over_h0(cosmo, aa))
        res.append(_C_KMS / (100.0 * cosmo["h"]) * np.sum(0.5 * (f[1:] + f[:-1]) * np.diff(aa)))
    res = np.array(res)
    return res if np.ndim(a) else float(res[0])


def linear_matter_power(cosmo, k, a):
    """P_lin(k, a) in Mpc^3, k in Mpc^-1.  NaN at k=0 like CCL (box.py:167)."""
    k = np.asarray(k, dtype=np.float64)
    with np.errstate(divide="ignore", invalid="ignore"):
        pk = cosmo._norm() * _pk_shape(cosmo, k) * growth_factor(cosmo, float(a)) ** 2
    return np.where(k > 0.0, pk, np.nan)


def nonlin_matter_power(cosmo, k, a):
    return linear_matter_power(cosmo, k, a)


class _Backend(object):
    """Namespace with the CCL call signatures the hot path uses."""

    def __init__(self, mod, name):
        self.name = name
        for fn in ("Cosmology", "linear_matter_power", "nonlin_matter_power",
                   "h_over_h0", "growth_factor", "growth_rate",
                   "comoving_angular_distance"):
            setattr(self, fn, getattr(mod, fn))


_WARNED = False


def get_backend(prefer_ccl=True):
    """Return CCL if importable (as the reference uses), else this module."""
    import sys
    if prefer_ccl:
        try:
            import pyccl  # noqa: F401
            if hasattr(pyccl, "comoving_angular_distance"):
                return _Backend(pyccl, "pyccl")
        except Exception:
            pass
    global _WARNED
    if _WARNED:
        return _Backend(sys.modules[__name__], "builtin-eh98")
    _WARNED = True
    import warnings
    warnings.warn("pyccl is not importable: fastbox_b200 uses its built-in cosmology provider (Eisenstein-Hu 1998 "
                  "zero-baryon P(k) normalised to sigma8, flat LCDM background).  nonlin_matter_power equals the "
                  "linear spectrum (no halofit), growth_rate = Omega_m(a)^0.55, and the `transfer_function` keyword "
                  "only accepts 'eisenstein_hu'.", RuntimeWarning, stacklevel=2)
    return _Backend(sys.modules[__name__], "builtin-eh98")
