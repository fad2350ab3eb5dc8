"""
GPU tests of the slab-decomposed pipelines.

* building blocks + NCCL-style bookkeeping: world = 1 and an emulated 2-rank split on one GPU;
* the library-resident exchange (``fb_dist_*``: peer stores + epoch flags): world = 1, and 2 / 4 ranks driven
  from one process on one GPU phase by phase (every rank's stores and signal complete before any rank waits,
  so no kernel ever spins on another);
* real multi-process runs (``torchrun`` of ``tools/dist_check.py``: CUDA IPC over NVLink, and NCCL) when the
  box has at least two GPUs.
"""
import numpy as np
import pytest

from fastbox_b200 import _lib
from fastbox_b200 import dist as fbd

from _util import TOL, rel_l2, setup_plan, transfer_fn

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("N", [32, 64])
def test_world1_pipeline_equals_fused_call(gpu, N):
    import torch
    L = (1e3, 1e3, 1e3)
    plan, edges = setup_plan(N, L, 0.8, nbins=20, filt=transfer_fn)
    ref = np.empty((N, N, N), np.float32)
    flags = _lib.F_SQRTPK | _lib.F_FILTER
    res_ref, _ = plan.realise(None, None, seed=5, flags=flags, field_out=ref, want_pk=True)
    plan.close()
    eng = fbd.CudaEngine(N, L, 0, 1, 0)
    p2, _ = setup_plan(N, L, 0.8, nbins=20, filt=transfer_fn)      # same tables on the engine's plan
    p2.close()
    # re-create tables on the engine plan
    from fastbox_b200 import kspace as ks
    from oracle import restate as R
    from _util import pk_function
    _, pkf = pk_function(0.8)
    eng.plan.set_sqrt_pk(ks.sqrt_pk_int_lut(pkf, N, L[0], R.boxfactor(N, *L)), 1)
    ft = ks.filter_tables(transfer_fn, N, *L)
    eng.plan.set_filter(ft.tperp, ft.tpar, ft.tdense)
    eng.plan.set_pk_bins(ks.bin_thresholds(edges))
    dr = fbd.DistributedRealiser(eng)
    field, pk, sums = dr.realise(5, flags, want_pk=True)
    torch.cuda.synchronize()
    got = field.cpu().numpy().reshape(N, N, N)
    assert rel_l2(got, ref.astype(np.float64)) < 1e-6
    assert np.array_equal(pk["count"], res_ref["count"])
    assert np.allclose(pk["sum1"], res_ref["sum1"], rtol=1e-12)
    # forward direction through the same building blocks: P(k) of the realised (filtered) field
    pkf = dr.power_spectrum()
    assert np.array_equal(pkf["count"], res_ref["count"])
    m = res_ref["count"] > 0
    floor = 1e-12 * np.abs(res_ref["sum1"]).max()
    assert np.all(np.abs(pkf["sum1"][m] - res_ref["sum1"][m]) <= 2e-5 * np.abs(res_ref["sum1"][m]) + floor)


def test_two_slabs_emulated_on_one_gpu(gpu):
    """Run both ranks' local passes sequentially on one GPU and do the exchange by hand."""
    import torch
    N, L, world = 32, (1e3, 1e3, 1e3), 2
    flags = _lib.F_SQRTPK
    plan, edges = setup_plan(N, L, 0.8, nbins=20)
    ref = np.empty((N, N, N), np.float32)
    res_ref, _ = plan.realise(None, None, seed=9, flags=flags, field_out=ref, want_pk=True)
    plan.close()
    from fastbox_b200 import kspace as ks
    from oracle import restate as R
    from _util import pk_function
    _, pkf = pk_function(0.8)
    engines = []
    for r in range(world):
        e = fbd.CudaEngine(N, L, r, world, 0)
        e.plan.set_sqrt_pk(ks.sqrt_pk_int_lut(pkf, N, L[0], R.boxfactor(N, *L)), 1)
        e.plan.set_pk_bins(ks.bin_thresholds(edges))
        engines.append(e)
    moments = [e.realise_kspace(9, flags, True) for e in engines]
    for e in engines:
        e.sync()
    # all-to-all by hand: rank d receives block d of every rank's send buffer, in rank order
    for d, e in enumerate(engines):
        blocks = [src.send[d].reshape(-1) for src in engines]
        e.recv.reshape(-1).copy_(torch.cat(blocks))
    torch.cuda.synchronize()
    full = np.empty((N, N, N), np.float32)
    for e in engines:
        e.x_to_real()
        e.sync()
        full[:, e.y0:e.y0 + e.ny, :] = e.field.cpu().numpy()
    assert rel_l2(full, ref.astype(np.float64)) < 1e-6
    cnt = sum(m["count"].astype(np.int64) for m in moments)
    assert np.array_equal(cnt, res_ref["count"].astype(np.int64))
    s1 = sum(m["sum1"] for m in moments)
    assert np.allclose(s1, res_ref["sum1"], rtol=1e-12)


def test_chunked_pipeline_world1(gpu):
    """Chunked (overlap-capable) pipeline with the gathered x pass == fused single call."""
    import torch
    from fastbox_b200 import kspace as ks
    from oracle import restate as R
    from _util import pk_function
    N, L = 64, (1e3, 1e3, 1e3)
    flags = _lib.F_SQRTPK
    plan, edges = setup_plan(N, L, 0.8, nbins=20)
    ref = np.empty((N, N, N), np.float32)
    res_ref, _ = plan.realise(None, None, seed=3, flags=flags, field_out=ref, want_pk=True)
    plan.close()
    _, pkf = pk_function(0.8)
    eng = fbd.CudaEngine(N, L, 0, 1, 0, chunks=4)
    eng.plan.set_sqrt_pk(ks.sqrt_pk_int_lut(pkf, N, L[0], R.boxfactor(N, *L)), 1)
    eng.plan.set_pk_bins(ks.bin_thresholds(edges))
    dr = fbd.DistributedRealiser(eng)
    field, pk, sums = dr.realise_overlapped(3, flags, want_pk=True)
    torch.cuda.synchronize()
    assert rel_l2(field.cpu().numpy().reshape(N, N, N), ref.astype(np.float64)) < 1e-6
    assert np.array_equal(pk["count"], res_ref["count"])
    assert np.allclose(pk["sum1"], res_ref["sum1"], rtol=1e-12)


def _nvlink_tables(plan, N, L, edges, filt=None):
    from fastbox_b200 import kspace as ks
    from oracle import restate as R
    from _util import pk_function
    _, pkf = pk_function(0.8)
    plan.set_sqrt_pk(ks.sqrt_pk_int_lut(pkf, N, L[0], R.boxfactor(N, *L)), 1)
    plan.set_pk_bins(ks.bin_thresholds(edges))
    if filt is not None:
        ft = ks.filter_tables(filt, N, *L)
        plan.set_filter(ft.tperp, ft.tpar, ft.tdense)


@pytest.mark.parametrize("N,world,chunks,xmode", [(32, 1, 1, 0), (64, 1, 4, 0), (32, 2, 1, 0), (64, 2, 2, 0),
                                                  (64, 4, 2, 0), (128, 8, 2, 0), (64, 2, 2, 1), (128, 8, 4, 1),
                                                  (64, 2, 2, 2), (128, 8, 4, 2), (64, 4, 1, 2), (64, 2, 2, 3), (128, 8, 4, 3),
                                                  (256, 2, 4, 3)])
def test_library_exchange_emulated_ranks(gpu, monkeypatch, N, world, chunks, xmode):
    """
    fb_dist_realise / fb_dist_power_spectrum: `world` plans on one GPU, connected through their handles.
    xmode 0: the y pass stores into the peers' buffers; 1: local blocks + copy-engine transfers; 2: local blocks +
    the high-priority copy kernel; 3: the copy kernel drives the bulk copy engine (cp.async.bulk).
    """
    monkeypatch.setenv("FB_DIST_XMODE", str(xmode))
    L = (1e3, 1e3, 1e3)
    flags = _lib.F_SQRTPK | _lib.F_FILTER
    plan, edges = setup_plan(N, L, 0.8, nbins=20, filt=transfer_fn, exact_below=4096)
    ref = np.empty((N, N, N), np.float32)
    res_ref, sums_ref = plan.realise(None, None, seed=21, flags=flags, field_out=ref, want_pk=True)
    fwd_ref = plan.field_to_spectrum(ref, want_pk=True)
    plan.close()
    plans = []
    for r in range(world):
        pl = _lib.Plan(N, *L)
        _nvlink_tables(pl, N, L, edges, transfer_fn)
        pl.dist_init(r, world, True)
        plans.append(pl)
    handles = np.stack([pl.dist_handle() for pl in plans])
    for pl in plans:
        pl.dist_connect(handles)
    infos = [pl.dist_info() for pl in plans]
    assert [i["a0"] for i in infos] == [r * (N // 2 // world) for r in range(world)]
    assert sum(i["na"] for i in infos) == N // 2 + 1 and all(i["ny"] == N // world for i in infos)
    ny = N // world
    fields = [pl.alloc(N * ny * N * 4) for pl in plans]
    for step in range(3):                                   # three steps: both receive buffers get reused
        for pl, f in zip(plans, fields):
            pl.dist_realise(21, flags, f, chunks=chunks, phase=1, want_pk=True)
        for pl in plans:
            pl.sync()
        out = [pl.dist_realise(21, flags, f, chunks=chunks, phase=2, want_pk=True, want_sums=True)
               for pl, f in zip(plans, fields)]
    full = np.empty((N, N, N), np.float32)
    for r, (pl, f) in enumerate(zip(plans, fields)):
        full[:, r * ny:(r + 1) * ny, :] = pl.download(f, (N, ny, N), np.float32)
    assert np.array_equal(full, ref)                        # same kernels, same arithmetic: bit identical
    for res, sums in out:                                   # every rank holds the global moments
        assert np.array_equal(res["count"], res_ref["count"])
        assert np.allclose(res["sum1"], res_ref["sum1"], rtol=1e-12)
        assert np.allclose(res["sum2"], res_ref["sum2"], rtol=1e-12)
    assert abs(sum(s[1] for _, s in out) - sums_ref[1]) <= 1e-9 * sums_ref[1]
    # forward: P(k) of the sharded field, phases 1 (x pass into the peers), 2 (k-space, moments), 3 (sum)
    for rep in range(2):
        pend = [pl.dist_power_spectrum(f, phase=1) for pl, f in zip(plans, fields)]
        for pl in plans:
            pl.sync()
        for pl, f, res in zip(plans, fields, pend):
            pl.dist_power_spectrum(f, phase=2, res=res)
        for pl in plans:
            pl.sync()
        for pl, f, res in zip(plans, fields, pend):
            pl.dist_power_spectrum(f, phase=3, res=res)
    for res in pend:
        assert np.array_equal(res["count"], fwd_ref["count"])
        m = fwd_ref["count"] > 0
        assert np.allclose(res["sum1"][m], fwd_ref["sum1"][m], rtol=1e-6, atol=1e-12 * np.abs(fwd_ref["sum1"]).max())
    for pl in plans:
        pl.close()


def _gpu_count():
    import subprocess
    try:
        out = subprocess.run(["nvidia-smi", "-L"], capture_output=True, text=True, timeout=30).stdout
        return sum(1 for ln in out.splitlines() if ln.startswith("GPU "))
    except Exception:
        return 0


@pytest.mark.parametrize("mode", ["p2p", "nccl"])
def test_real_multiprocess_exchange(gpu, mode):
    """torchrun --nproc 2 of tools/dist_check.py: slab-decomposed field / P(k) == single-GPU result."""
    import os
    import socket
    import subprocess
    import sys
    if _gpu_count() < 2:
        pytest.skip("needs two GPUs (the driver's multi-GPU run covers it)")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", str(port), os.path.join(root, "tools", "dist_check.py"), "--mode", mode]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "DIST CHECK OK" in out.stdout, out.stdout[-2000:]
