"""GPU test of the slab building blocks (world = 1 and emulated 2-rank split on one GPU, no NCCL)."""
import numpy as np
import pytest

from fastbox_b200 import _lib
from fastbox_b200 import dist as fbd

from _util import TOL, rel_l2, setup_plan, transfer_fn

pytestmark = pytest.mark.gpu


def _configure(plan, ref_plan_tables):
    pass


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
