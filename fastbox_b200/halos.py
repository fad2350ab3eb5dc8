"""
Poisson halo / galaxy counts on top of a density field (reference
``fastbox/halos.py``).  ``halo_count_field`` (halos.py:53-117) and
``realise_halo_catalogue`` (halos.py:120-176) run on the device.
"""
import numpy as np

from . import _lib


class HaloDistribution(object):

    def __init__(self, box, mass_range, mass_bins):
        self.box = box
        self.Mmin, self.Mmax = mass_range
        self.mass_bins = mass_bins

    @staticmethod
    def _classify(x, N):
        """(float32 array, kind): scalar -> 0, (N,) along z -> 1, (N,N,N) -> 2 (halos.py:91-99)."""
        a = np.atleast_1d(np.asarray(x, dtype=np.float64))
        if a.ndim == 1 and a.size == 1:
            return np.ascontiguousarray(a, dtype=np.float32), 0
        if a.ndim == 1 and a.size == N:
            return np.ascontiguousarray(a, dtype=np.float32), 1
        if a.shape == (N, N, N):
            return np.ascontiguousarray(a, dtype=np.float32), 2
        raise ValueError("nbar / bias must be a scalar, an (N,) array along z, or an (N,N,N) array")

    def halo_count_field(self, delta_x, nbar, bias, lognormal=False, uniforms=None, return_mean=False):
        """
        N_halo(x) ~ Poisson(vol_voxel * nbar * (1 + bias * delta_x)) (halos.py:53-117).

        One uniform per voxel drives a Poisson inversion on the device
        (``include/fb_poisson.h``).  By default the uniforms come from NumPy's
        global generator (``np.random.uniform``), so results are reproducible under
        ``np.random.seed`` but follow a different stream from the reference's
        ``np.random.poisson``; pass ``uniforms`` to control them.  Returns int64
        counts like the reference.
        """
        box = self.box
        N = box.N
        plan = box._plan
        nb, nbk = self._classify(nbar, N)
        bi, bik = self._classify(bias, N)
        if uniforms is None:
            uniforms = np.random.uniform(0., 1., (N, N, N))
        u = np.ascontiguousarray(uniforms, dtype=np.float64)
        d = box._to_device_field(delta_x)
        counts = np.empty((N, N, N), dtype=np.int32)
        mean = np.empty((N, N, N), dtype=np.float32) if return_mean else None
        plan.halo_counts(d, nb, nbk, bi, bik, lognormal, 0.0, u, counts, mean)
        if return_mean:
            return counts.astype(np.int64), mean
        return counts.astype(np.int64)

    def realise_halo_catalogue(self, Nhalo, scatter=False, scatter_type='uniform', uniforms=None):
        """
        Counts per voxel -> catalogue of comoving positions, shape (Nhalos, 3)
        (halos.py:120-176), built on the device (``fb_halo_catalogue``: per-tile
        histograms, exclusive scan, stable scatter).  Row order is the reference's:
        ascending count value, then C-order voxel index, every voxel repeated
        ``count`` times.  With ``scatter=True`` the offsets are
        ``np.random.uniform(0, 1-1e-8, 3*Nhalos)`` drawn exactly as the reference
        draws them (halos.py:166), or ``uniforms`` (Nhalos, 3) if given, so the
        catalogue is bit-identical to the reference's under the same seed.
        """
        if scatter and scatter_type != 'uniform':
            raise ValueError("scatter_type='%s' not recognised" % scatter_type)
        box = self.box
        N = box.N
        plan = box._plan
        cnt = np.asarray(Nhalo)
        if cnt.shape != (N, N, N):
            raise ValueError("Nhalo must have shape (N, N, N)")
        if cnt.size and (cnt.min() < 0 or cnt.max() > np.iinfo(np.int32).max):
            raise ValueError("halo counts must be non-negative 32-bit integers")
        cnt = plan.upload(np.ascontiguousarray(cnt, dtype=np.int32))     # one upload for the size query and the fill
        nh = plan.halo_catalogue(cnt)
        cat = np.empty((nh, 3), dtype=np.float64)
        if nh == 0:
            return cat
        u = None
        if scatter:
            if uniforms is None:
                uniforms = np.random.uniform(0., 1. - 1e-8, cat.size).reshape(cat.shape)
            u = np.ascontiguousarray(uniforms, dtype=np.float64)
            if u.shape != cat.shape:
                raise ValueError("uniforms must have shape (Nhalos, 3) = %s" % (cat.shape,))
        plan.halo_catalogue(cnt, u, cat, nh)
        return cat

