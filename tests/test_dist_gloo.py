"""
world_size-2 gloo test (CPU) of the slab-decomposition bookkeeping in fastbox_b200/dist.py:
plane / row ownership, all_to_all split sizes and buffer layouts, P(k) moment all-reduce.
The CUDA passes are replaced by a NumPy engine built from the oracle (test infrastructure);
the assembled field must equal the single-process oracle realisation.
"""
import os
import socket

import numpy as np
import pytest

from fastbox_b200 import _lib
from fastbox_b200 import dist as fbd
from fastbox_b200 import kspace as ks
from oracle import restate as R

from _util import pk_function

N, L, SEED = 16, (5e2, 5e2, 5e2), 77


def test_slab_geometry_covers_everything():
    for n, w in [(16, 2), (64, 4), (2048, 8), (32, 1)]:
        planes, rows = [], []
        for r in range(w):
            a0, na, y0, ny = fbd.slab_geometry(n, w, r)
            planes += list(range(a0, a0 + na))
            rows += list(range(y0, y0 + ny))
        assert planes == list(range(n // 2 + 1)) and rows == list(range(n))
        assert sum(fbd.plane_counts(n, w)) == n // 2 + 1
    with pytest.raises(ValueError):
        fbd.slab_geometry(16, 3, 0)
    assert fbd.alltoall_bytes_per_rank(2048, 8)[0] == 8 * 128 * 256 * 2048 * 7
    # chunked exchange: every plane lands exactly once in the receive buffer
    for n, w, c in [(32, 2, 4), (64, 4, 2), (2048, 8, 4)]:
        off = fbd.plane_offsets(n, w, c, n // w)
        assert np.array_equal(np.sort(off), np.arange(n // 2 + 1) * (n // w) * n)
        assert sum(cnt for r in range(w) for _, cnt in fbd.chunk_planes(n, w, r, c)) == n // 2 + 1
    with pytest.raises(ValueError):
        fbd.chunk_planes(32, 2, 0, 3)


class NumpyEngine(object):
    """Same interface as dist.CudaEngine, NumPy float64 arithmetic from oracle/restate.py."""

    def __init__(self, rank, world, chunks=1):
        import torch
        self.torch = torch
        self.N, self.rank, self.world, self.chunks = N, rank, world, chunks
        self.a0, self.na, self.y0, self.ny = fbd.slab_geometry(N, world, rank)
        self.send = torch.empty((world, self.na, self.ny, N), dtype=torch.complex128)
        self.recv = torch.empty((N // 2 + 1, self.ny, N), dtype=torch.complex128)
        self.field = None
        if chunks > 1:
            self.my_chunks = fbd.chunk_planes(N, world, rank, chunks)
            self.send_c = [torch.empty((world, n, self.ny, N), dtype=torch.complex128) for _, n in self.my_chunks]
            self.plane_off = fbd.plane_offsets(N, world, chunks, self.ny)
        self.nedges = 20
        _, self.pkf = pk_function(0.5)
        self.edges = R.pk_bin_edges(N, *L, nbins=self.nedges)

    def realise_kspace_chunk(self, c, seed, flags, want_pk):
        a_first, n = self.my_chunks[c]
        saved = (self.a0, self.na, self.send)
        self.a0, self.na, self.send = a_first, n, self.send_c[c]
        try:
            return self.realise_kspace(seed, flags, want_pk)
        finally:
            self.a0, self.na, self.send = saved

    def x_to_real_gather(self, flags=0, scale=1.0):
        flat = self.recv.numpy().reshape(-1)
        planes = np.stack([flat[o:o + self.ny * N].reshape(self.ny, N) for o in self.plane_off])
        self.field = np.fft.irfft(planes, n=N, axis=0) * N * scale / float(N) ** 3
        return (float(self.field.sum()), float((self.field ** 2).sum()))

    def realise_kspace(self, seed, flags, want_pk):
        idx = np.arange(N ** 3, dtype=np.uint64).reshape(N, N, N)
        re, im = R.philox_normals(seed, idx, N)                      # every rank can evaluate any cell
        half = R.hermitian_half_from_noise(re, im, R.sqrt_pk_half(self.pkf, N, *L))
        local = half[self.a0:self.a0 + self.na]
        work = np.fft.ifft(np.fft.ifft(local, axis=2), axis=1) * N * N       # z rows, y columns (unnormalised)
        # [plane][y][z] -> [dest][plane][y'][z]
        self.send.copy_(self.torch.from_numpy(
            np.ascontiguousarray(work.reshape(self.na, self.world, self.ny, N).transpose(1, 0, 2, 3))))
        if not want_pk:
            return None
        idxb = R.digitize_half(N, *L, self.edges)[self.a0:self.a0 + self.na]
        p = (local * np.conj(local)).real / R.boxfactor(N, *L)
        w = np.broadcast_to(R.half_weights(N)[self.a0:self.a0 + self.na, None, None], p.shape)
        c, s1, s2 = R.pk_moments(p.ravel(), idxb.ravel(), self.nedges, w.ravel())
        return dict(count=c, sum1=s1, sum2=s2)

    def x_to_real(self, flags=0, scale=1.0):
        spec = self.recv.numpy()
        self.field = np.fft.irfft(spec, n=N, axis=0) * N * scale / float(N) ** 3
        return (float(self.field.sum()), float((self.field ** 2).sum()))

    def x_from_real(self):
        self.recv.copy_(self.torch.from_numpy(np.fft.rfft(self.field, axis=0)))      # [kx][y'][z]

    def forward_kspace(self, want_pk=True, spec_out=None, poles=False):
        blocks = self.send.numpy()                                   # [src][plane][y'][z]
        planes = np.concatenate([blocks[s] for s in range(self.world)], axis=1)       # [plane][y][z]
        spec = np.fft.fft(np.fft.fft(planes, axis=1), axis=2)
        idxb = R.digitize_half(N, *L, self.edges)[self.a0:self.a0 + self.na]
        p = (spec * np.conj(spec)).real / R.boxfactor(N, *L)
        w = np.broadcast_to(R.half_weights(N)[self.a0:self.a0 + self.na, None, None], p.shape)
        c, s1, s2 = R.pk_moments(p.ravel(), idxb.ravel(), self.nedges, w.ravel())
        return dict(count=c, sum1=s1, sum2=s2)

    def sync(self):
        pass

    def sync_exchange(self):
        pass

    def moments_tensor(self, res):
        n = self.nedges + 1
        return self.torch.from_numpy(np.concatenate([res["count"][:n].astype(np.float64), res["sum1"][:n],
                                                     res["sum2"][:n]]))


def _worker(rank, world, port, outdir, chunks):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    eng = NumpyEngine(rank, world, chunks)
    dr = fbd.DistributedRealiser(eng)
    if chunks > 1:
        field, pk, sums = dr.realise_overlapped(SEED, _lib.F_SQRTPK, want_pk=True)
    else:
        field, pk, sums = dr.realise(SEED, _lib.F_SQRTPK, want_pk=True)
    fwd = dr.power_spectrum()                                       # reverse exchange + forward passes
    np.savez(os.path.join(outdir, "r%d.npz" % rank), field=eng.field, y0=eng.y0, fcount=fwd["count"],
             fsum1=fwd["sum1"], **pk)
    dist.destroy_process_group()


@pytest.mark.parametrize("world,chunks", [(2, 1), (2, 2)])
def test_two_rank_realise_matches_single_process(tmp_path, world, chunks):
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_worker, args=(world, port, str(tmp_path), chunks), nprocs=world, join=True)
    idx = np.arange(N ** 3, dtype=np.uint64).reshape(N, N, N)
    re, im = R.philox_normals(SEED, idx, N)
    _, pkf = pk_function(0.5)
    ref, half = R.realise_density_lean(re, im, pkf, N, *L)
    full = np.empty((N, N, N))
    pks = []
    for r in range(world):
        d = np.load(os.path.join(str(tmp_path), "r%d.npz" % r))
        ny = N // world
        full[:, int(d["y0"]):int(d["y0"]) + ny, :] = d["field"]
        pks.append(d)
    assert np.abs(full - ref).max() < 1e-12 * np.abs(ref).max()
    kc, pk, err, cnt = R.binned_power_spectrum_lean(half, N, *L, nbins=20)
    for d in pks:                                            # every rank holds the global moments
        assert np.array_equal(d["count"][:20].astype(np.int64), cnt[:20])
        got = ks.moments_to_spectrum(R.pk_bin_edges(N, *L, nbins=20), d["count"], d["sum1"], d["sum2"])
        m = ~np.isnan(pk)
        assert np.allclose(got[1][m], pk[m], rtol=1e-12)
        # P(k) re-measured from the sharded real field agrees with the spectrum it was made from
        assert np.array_equal(d["fcount"][:20].astype(np.int64), cnt[:20])
        assert np.allclose(d["fsum1"][1:20][m] / np.maximum(d["fcount"][1:20][m], 1), pk[m], rtol=1e-9)
