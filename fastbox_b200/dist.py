"""
Slab-decomposed (multi-GPU) realise / P(k) pipelines: one process per GPU,
``torch.distributed`` (NCCL over NVLink) for the single exchange step of each 3-D FFT.

Decomposition (DESIGN.md section "Multi-GPU"):

* spectrum side: the kx planes a = 0..N/2 are split over the P ranks (N/(2P) planes each,
  the last rank also owns the Nyquist plane a = N/2); the z (rows) and y (columns) passes are
  local;
* real side: y rows are split over the ranks (N/P each); the x pass is local.

Between the two sits ONE all-to-all of the half spectrum per transform.  The y pass writes its
output already grouped by destination rank (``[dest][plane][y'][z]``) and the receive buffer of
``all_to_all_single`` is exactly the ``[kx][y'][z]`` array the x pass consumes, so no pack/unpack
kernels are needed.  P(k) moments are per-rank partial histograms summed with ``all_reduce``.

The reference (``fastbox/box.py``) is single process; this module has no counterpart there.

Two exchange mechanisms exist:

* ``NvlinkRealiser`` (default): the exchange lives INSIDE the C library (``csrc/fb_dist.cu``).  Every rank
  maps the receive buffers of its peers (CUDA IPC over NVLink) and the y pass stores each element straight
  into the buffer of the rank that owns it -- transform pass and all-to-all are one kernel; the P(k) moments
  are summed through per-rank slots and the ranks meet at device-side epoch flags.  ``torch.distributed`` is
  used once, at set-up, to publish the 128-byte handles.
* ``DistributedRealiser`` + ``CudaEngine``: the same passes with ``all_to_all_single`` (NCCL) between them;
  kept as the comparison point and for clusters where peer mapping is not available.

The local compute of the NCCL variant is behind ``engine`` so that the exchange bookkeeping can be exercised
on CPU (gloo) with a NumPy engine in ``tests/test_dist_gloo.py``; the product engines are C ABI only (no CPU
fallback).
"""
import numpy as np

from . import _lib


def slab_geometry(N, world, rank):
    """(a0, na, y0, ny): local kx planes [a0, a0+na) and local y rows [y0, y0+ny)."""
    if world < 1 or (N // 2) % world != 0 or N % world != 0:
        raise ValueError("world size %d must divide N/2 = %d" % (world, N // 2))
    per = (N // 2) // world
    a0 = rank * per
    na = per + (1 if rank == world - 1 else 0)
    ny = N // world
    return a0, na, rank * ny, ny


def plane_counts(N, world):
    return [slab_geometry(N, world, r)[1] for r in range(world)]


def alltoall_bytes_per_rank(N, world):
    """Bytes each rank sends (= receives) per transform, excluding its own block."""
    ny = N // world
    return [8 * na * ny * N * (world - 1) for na in plane_counts(N, world)]


def chunk_planes(N, world, rank, chunks):
    """[(first global plane, count)] for each exchange chunk of this rank (Nyquist plane in the last)."""
    a0, na, _, _ = slab_geometry(N, world, rank)
    per = (N // 2) // world
    if chunks < 1 or per % chunks != 0:
        raise ValueError("chunks=%d must divide the %d planes per rank" % (chunks, per))
    nc = per // chunks
    out = [(a0 + c * nc, nc) for c in range(chunks)]
    if rank == world - 1:
        out[-1] = (out[-1][0], nc + 1)
    return out


def recv_regions(N, world, chunks):
    """Plane offset of each chunk's region in the receive buffer, and planes per (chunk, source)."""
    counts = [[chunk_planes(N, world, s, chunks)[c][1] for s in range(world)] for c in range(chunks)]
    starts = np.concatenate([[0], np.cumsum([sum(row) for row in counts])])
    return starts[:-1].astype(np.int64), counts


def plane_offsets(N, world, chunks, ny):
    """
    Element offset (in complex64 units) of every global kx plane a = 0..N/2 inside the receive
    buffer when the exchange is done chunk by chunk: region(chunk) -> source rank -> plane.
    """
    starts, counts = recv_regions(N, world, chunks)
    off = np.empty(N // 2 + 1, dtype=np.int64)
    for c in range(chunks):
        pos = int(starts[c])
        for s in range(world):
            a_first, n = chunk_planes(N, world, s, chunks)[c]
            off[a_first:a_first + n] = (pos + np.arange(n)) * ny * N
            pos += n
    return off


class CudaEngine(object):
    """Local passes on one GPU through libfastbox_b200 (torch tensors carry the buffers)."""

    def __init__(self, N, L, rank, world, device, chunks=1):
        import torch
        self.torch = torch
        self.N, self.rank, self.world, self.chunks = N, rank, world, chunks
        self.a0, self.na, self.y0, self.ny = slab_geometry(N, world, rank)
        self.dev = torch.device("cuda", device)
        self.plan = _lib.Plan(N, L[0], L[1], L[2], device)
        self.plan.set_slab(self.a0, self.na, self.y0, self.ny)
        c64 = torch.complex64
        self.work = torch.empty((self.na, N, N), dtype=c64, device=self.dev)
        self.send = torch.empty((world, self.na, self.ny, N), dtype=c64, device=self.dev)
        self.recv = torch.empty((N // 2 + 1, self.ny, N), dtype=c64, device=self.dev)
        self.field = torch.empty((N, self.ny, N), dtype=torch.float32, device=self.dev)
        if chunks > 1:
            self.my_chunks = chunk_planes(N, world, rank, chunks)
            self.send_c = [torch.empty((world, n, self.ny, N), dtype=c64, device=self.dev) for _, n in self.my_chunks]
            self.plane_off = torch.from_numpy(plane_offsets(N, world, chunks, self.ny)).to(self.dev)

    def realise_kspace_chunk(self, c, seed, flags, want_pk):
        """Same as realise_kspace for exchange chunk c only (fills ``self.send_c[c]``)."""
        a_first, n = self.my_chunks[c]
        self.plan.set_slab(a_first, n, self.y0, self.ny)
        try:
            return self.plan.realise_local_kspace(seed, flags, self.work, self.send_c[c], self.ny, want_pk=want_pk)
        finally:
            self.plan.set_slab(self.a0, self.na, self.y0, self.ny)

    def x_to_real_gather(self, flags=0, scale=1.0):
        N = self.N
        return self.plan.fft_pass_x_c2r_gather(self.recv, self.plane_off, self.field, self.ny * N, flags=flags,
                                               scale=scale / float(N) ** 3)

    def realise_kspace(self, seed, flags, want_pk):
        """rows + columns on the local planes; fills ``self.send``; returns local P(k) moments."""
        return self.plan.realise_local_kspace(seed, flags, self.work, self.send, self.ny, want_pk=want_pk)

    def x_to_real(self, flags=0, scale=1.0):
        N = self.N
        return self.plan.fft_pass_x_c2r(self.recv, self.field, self.ny * N, flags=flags,
                                        scale=scale / float(N) ** 3)

    def x_from_real(self):
        """Forward x pass on the local y slab: ``self.field`` -> ``self.recv`` ([kx][y'][z])."""
        self.plan.fft_pass_x_r2c(self.field, self.recv, self.ny * self.N)

    def forward_kspace(self, want_pk=True, spec_out=None, poles=False):
        """y columns + z rows on the local kx planes of the received spectrum (``self.send`` buffer)."""
        return self.plan.forward_local_kspace(self.send, self.work, self.ny, spec_out=spec_out, want_pk=want_pk,
                                              poles=poles)

    def sync(self):
        self.plan.sync()

    def sync_exchange(self):
        """NCCL runs on torch's current stream: wait for it before the x pass (plan stream)."""
        self.torch.cuda.current_stream(self.dev).synchronize()

    def moments_tensor(self, res):
        t = self.torch
        n = self.plan.nedges + 1
        arr = np.concatenate([res["count"][:n].astype(np.float64), res["sum1"][:n], res["sum2"][:n]])
        return t.from_numpy(arr).to(self.dev)


class DistributedRealiser(object):
    """Realise (+ filter + P(k)) a Gaussian field on an N^3 grid sharded over the ranks."""

    def __init__(self, engine, group=None):
        import torch.distributed as dist
        self.dist = dist
        self.e = engine
        self.group = group
        N, world = engine.N, engine.world
        ny = engine.ny
        self.in_splits = [engine.na * ny * N] * world                 # elements sent to each peer
        self.out_splits = [na * ny * N for na in plane_counts(N, world)]
        self.chunks = getattr(engine, "chunks", 1)
        if self.chunks > 1:
            self.region_start, self.region_counts = recv_regions(N, world, self.chunks)

    def exchange(self):
        e = self.e
        if e.world == 1:
            e.recv.reshape(-1).copy_(e.send.reshape(-1))
            return
        self.dist.all_to_all_single(e.recv.reshape(-1), e.send.reshape(-1),
                                    output_split_sizes=self.out_splits, input_split_sizes=self.in_splits,
                                    group=self.group)

    def exchange_chunk(self, c):
        """Asynchronous all-to-all of chunk c into its region of the receive buffer."""
        e = self.e
        N, ny = e.N, e.ny
        n_mine = e.my_chunks[c][1]
        counts = self.region_counts[c]
        start = int(self.region_start[c]) * ny * N
        total = sum(counts) * ny * N
        out = e.recv.reshape(-1)[start:start + total]
        if e.world == 1:
            out.copy_(e.send_c[c].reshape(-1))
            return None
        return self.dist.all_to_all_single(out, e.send_c[c].reshape(-1),
                                           output_split_sizes=[n * ny * N for n in counts],
                                           input_split_sizes=[n_mine * ny * N] * e.world, group=self.group,
                                           async_op=True)

    def realise_overlapped(self, seed, flags, want_pk=False, scale=1.0):
        """
        Chunked pipeline: the all-to-all of chunk c runs (NCCL stream) while the z/y passes of chunk
        c+1 run on the library stream; the x pass gathers the planes through ``plane_off``.
        """
        e = self.e
        pending, acc = [], None
        for c in range(self.chunks):
            res = e.realise_kspace_chunk(c, seed, flags, want_pk)
            e.sync()
            pending.append(self.exchange_chunk(c))
            if want_pk:
                acc = res if acc is None else {k: acc[k] + res[k] for k in ("count", "sum1", "sum2")}
        for w in pending:
            if w is not None:
                w.wait()
        e.sync_exchange()
        sums = e.x_to_real_gather(flags=flags & _lib.F_EXP, scale=scale)
        return e.field, self._reduce_moments(acc) if want_pk else None, sums

    def _reduce_moments(self, res):
        e = self.e
        t = e.moments_tensor(res)
        if e.world > 1:
            self.dist.all_reduce(t, group=self.group)
        arr = t.cpu().numpy() if hasattr(t, "cpu") else np.asarray(t)
        n = arr.size // 3
        return dict(count=np.rint(arr[:n]).astype(np.uint64), sum1=arr[n:2 * n], sum2=arr[2 * n:])

    def exchange_back(self):
        """Real-side slabs -> spectrum-side planes (the reverse all-to-all, used by the forward transform)."""
        e = self.e
        if e.world == 1:
            e.send.reshape(-1).copy_(e.recv.reshape(-1))
            return
        self.dist.all_to_all_single(e.send.reshape(-1), e.recv.reshape(-1),
                                    output_split_sizes=self.in_splits, input_split_sizes=self.out_splits,
                                    group=self.group)

    def power_spectrum(self):
        """
        Binned P(k) moments of the sharded real field in ``engine.field`` (box.py:736-764):
        local x pass -> all-to-all -> local y and z passes with the histogram epilogue -> all-reduce.
        """
        e = self.e
        e.x_from_real()
        e.sync()
        self.exchange_back()
        e.sync_exchange()
        res = e.forward_kspace(want_pk=True)
        return self._reduce_moments(res)

    def realise(self, seed, flags, want_pk=False, scale=1.0):
        """
        Returns (local field slab handle, global P(k) moments or None).  The y pass and the
        exchange are ordered on the device: the plan stream is synchronised before NCCL runs
        on torch's stream, and torch's stream before the x pass.
        """
        e = self.e
        res = e.realise_kspace(seed, flags, want_pk)
        e.sync()
        self.exchange()
        e.sync_exchange()
        sums = e.x_to_real(flags=flags & _lib.F_EXP, scale=scale)
        pk = self._reduce_moments(res) if want_pk else None
        return e.field, pk, sums


def gather_handles(blob, world, group=None):
    """All ranks' ``fb_dist_get_handle`` blobs in rank order ((world, 128) uint8), via torch.distributed."""
    if world == 1:
        return np.ascontiguousarray(blob).reshape(1, -1)
    import torch
    import torch.distributed as dist
    on_gpu = dist.get_backend(group) == "nccl"
    t = torch.from_numpy(np.ascontiguousarray(blob))
    if on_gpu:
        t = t.cuda()
    out = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(out, t, group=group)
    return np.stack([o.cpu().numpy() for o in out])


class NvlinkRealiser(object):
    """
    Slab-decomposed realise (+ filter + P(k)) and forward P(k) with the exchange inside the library
    (``fb_dist_*``, peer stores over NVLink).  One instance per process / GPU; every method is collective.
    """

    def __init__(self, N, L, rank, world, device, group=None, with_forward=True, chunks=16):
        self.N, self.rank, self.world, self.chunks = N, rank, world, chunks
        self.plan = _lib.Plan(N, L[0], L[1], L[2], device)
        self.plan.dist_init(rank, world, with_forward)
        self.plan.dist_connect(gather_handles(self.plan.dist_handle(), world, group))
        info = self.plan.dist_info()
        self.a0, self.na, self.y0, self.ny = info["a0"], info["na"], info["y0"], info["ny"]
        self.block_bytes = info["block_bytes"]
        self.field = self.plan.alloc(N * self.ny * N * 4)              # float32 [N][ny][N]
        self.plan.dist_barrier()                                       # every rank has mapped every block

    def realise(self, seed, flags, want_pk=False, scale=1.0, want_sums=False):
        """Returns (device field slab [N][ny][N], global P(k) moments or None, local (sum, sum of squares))."""
        res, sums = self.plan.dist_realise(seed, flags, self.field, scale=scale, chunks=self.chunks,
                                           want_pk=want_pk, want_sums=want_sums)
        return self.field, res, sums

    def power_spectrum(self, field=None, poles=False):
        """Global binned P(k) moments of the sharded real field (default: the last realised one)."""
        res = self.plan.dist_power_spectrum(self.field if field is None else field, poles=poles)
        res.pop("_st", None)
        return res

    def field_host(self):
        return self.plan.download(self.field, (self.N, self.ny, self.N), np.float32)

    def close(self):
        self.plan.close()
