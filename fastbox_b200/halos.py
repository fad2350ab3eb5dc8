"""
Poisson halo / galaxy counts on top of a density field (reference
``fastbox/halos.py``).  ``halo_count_field`` (halos.py:53-117) runs on the device.
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

    def realise_halo_catalogue(self, Nhalo, scatter=False, scatter_type='uniform'):
        """Counts -> comoving positions (halos.py:120-176); host side (not on the hot path)."""
        Nhalo = np.asarray(Nhalo)
        idx = np.nonzero(Nhalo > 0)
        reps = Nhalo[idx]
        order = np.argsort(reps, kind="stable")         # reference groups voxels by count value
        cat = np.column_stack([np.repeat(ax[order], reps[order]) for ax in idx]).astype(np.float64)
        if scatter:
            if scatter_type == 'uniform':
                cat += np.random.uniform(0., 1. - 1e-8, cat.size).reshape(cat.shape)
            else:
                raise ValueError("scatter_type='%s' not recognised" % scatter_type)
        cat[:, 0] *= box_len(self.box.Lx, self.box.N)
        cat[:, 1] *= box_len(self.box.Ly, self.box.N)
        cat[:, 2] *= box_len(self.box.Lz, self.box.N)
        return cat


def box_len(L, N):
    return L / N
