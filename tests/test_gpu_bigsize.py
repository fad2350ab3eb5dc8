"""
Parity at BASELINE.json's sizes against the float64 oracle, through the C ABI:

* 512^3 (configs[1]): fused realise + k_perp/k_par filter + P(k) on the FULL field, bias + log-normal, v_y, v_z and
  potential realisations from the stored spectrum, forward P(k) of the realised field;
* 1024^3 (configs[2,3], the headline instantiation: FftCfg<1024>, float-bit sqrt(P) table, 8/16-column tiles):
  16 x planes of the field, every bin population (bit exact) and every per-bin P(k) / error bar;
* 2048 (configs[4]): the N = 2048 kernels on kx slabs of the 2048^3 box (Philox rows pass + y pass + fused
  P(k) moments on planes 0, 1 and the Nyquist plane) and the 2048-point x passes.

Oracle = ``oracle.restate.realise_slabwise`` (float64, one kx plane at a time; validated on the CPU against the
restatements that are pinned bit-for-bit to the unmodified reference, tests/test_oracle_cpu.py).  The noise
is drawn on the host plane by plane (np.random.default_rng((seed, plane)), float32) and uploaded, so the oracle
and the device see identical input white noise.  Tolerances: north_star's 1e-5 relative L2 for fields, 1e-5
per bin for P(k), bit exact for bin populations.
"""
import os
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import pytest

from fastbox_b200 import _lib
from fastbox_b200 import kspace as ks
from oracle import restate as R

from _util import TOL, assert_pk_close, pk_function, rel_l2, setup_plan, transfer_fn

pytestmark = pytest.mark.gpu
F = _lib
F32_FLOOR = 1e-50            # float32 range: power this far below the largest bin underflows on the device
FWD_FLOOR = 1e-10            # P(k) re-measured from the float32 FIELD: its 1e-7 rounding noise is white, i.e. a
                             # flat ~1e-14 x sigma^2 pedestal under bins the filter has emptied
WORKERS = max(2, min(32, os.cpu_count() or 4))


def noise_plane(seed, a, N):
    rng = np.random.default_rng((seed, a))
    return rng.standard_normal((N, N), dtype=np.float32), rng.standard_normal((N, N), dtype=np.float32)


def upload_noise(plan, seed, N):
    """Device noise cubes re, im [N][N][N] float32 filled plane by plane."""
    d_re, d_im = plan.alloc(N ** 3 * 4), plan.alloc(N ** 3 * 4)
    plane_bytes = N * N * 4
    with ThreadPoolExecutor(WORKERS) as ex:
        for a0 in range(0, N, 64):
            planes = list(ex.map(lambda a: noise_plane(seed, a, N), range(a0, min(N, a0 + 64))))
            for i, (re, im) in enumerate(planes):
                _lib.check(plan.lib.fb_copy(plan.h, d_re.ptr + (a0 + i) * plane_bytes, re.ctypes.data, plane_bytes))
                _lib.check(plan.lib.fb_copy(plan.h, d_im.ptr + (a0 + i) * plane_bytes, im.ctypes.data, plane_bytes))
    return d_re, d_im


def download_planes(plan, buf, N, xs):
    out = np.empty((len(xs), N, N), np.float32)
    for j, x in enumerate(xs):
        _lib.check(plan.lib.fb_copy(plan.h, out[j].ctypes.data, buf.ptr + int(x) * N * N * 4, N * N * 4))
    return out


def test_512_full_field_and_derived_fields_vs_oracle(gpu):
    N, L, seed, nb = 512, (1e3, 1e3, 1e3), 41, 50
    _, pkf = pk_function(0.8)
    plan, edges = setup_plan(N, L, 0.8, nbins=nb, filt=transfer_fn)      # N >= 512: float-bit sqrt(P) table
    d_re, d_im = upload_noise(plan, seed, N)
    field = plan.alloc(N ** 3 * 4)
    spec = plan.alloc((N // 2 + 1) * N * N * 8)
    flags = F.F_SQRTPK | F.F_FILTER
    res, sums = plan.realise(d_re, d_im, flags=flags, field_out=field, spec_out=spec, want_pk=True)
    d_re.free()
    d_im.free()
    xs = list(range(0, N, 64)) + [1, N - 1]
    ref = R.realise_slabwise(lambda a: noise_plane(seed, a, N), pkf, N, *L, transfer_fn=transfer_fn, nbins=nb,
                             kinds=(None,), workers=WORKERS)
    full_ref = ref["fields"][None]
    got = plan.download(field, (N, N, N), np.float32)
    assert rel_l2(got, full_ref) < TOL
    assert abs(sums[1] / float((full_ref ** 2).sum()) - 1) < 1e-5 and abs(sums[0]) < 1e-3 * np.sqrt(sums[1] * N ** 3)
    # bin populations bit exact, per-bin P(k) and error bars to 1e-5
    assert np.array_equal(res["count"][:nb].astype(np.int64), ref["count"][:nb])
    assert int(res["count"].sum()) == N ** 3
    assert_pk_close(ks.moments_to_spectrum(edges, res["count"], res["sum1"], res["sum2"]),
                    ks.moments_to_spectrum(edges, ref["count"].astype(np.uint64), ref["sum1"], ref["sum2"]),
                    floor_rel=F32_FLOOR)
    # forward P(k) of the realised field (box.py:736-764) == moments of the spectrum it came from
    fwd = plan.field_to_spectrum(field, want_pk=True)
    assert np.array_equal(fwd["count"], res["count"])
    assert_pk_close(ks.moments_to_spectrum(edges, fwd["count"], fwd["sum1"], fwd["sum2"]),
                    ks.moments_to_spectrum(edges, ref["count"].astype(np.uint64), ref["sum1"], ref["sum2"]),
                    tol=2 * TOL, floor_rel=FWD_FLOOR)
    # bias + log-normal (tracers.py bias, box.py:457-459) on the full field
    b = 0.84081272
    s1, _ = plan.spectrum_to_field(spec, field, flags=F.F_EXP, scale=b)
    plan.affine(field, N ** 3, 1.0 / (s1 / N ** 3), -1.0)
    ln = plan.download(field, (N, N, N), np.float32)
    assert rel_l2(ln, R.lognormal(b * full_ref)) < TOL
    assert ln.min() >= -1.0
    del got, ln, full_ref
    # velocity components and potential from the stored (filtered) spectrum, on 10 x planes
    ref = R.realise_slabwise(lambda a: noise_plane(seed, a, N), pkf, N, *L, transfer_fn=transfer_fn,
                             kinds=("vel_y", "vel_z", "potential"), x_planes=xs, workers=WORKERS)
    fac = 57.3
    for kind, key, scale in ((F.KIND_VEL_Y, "vel_y", fac), (F.KIND_VEL_Z, "vel_z", fac), (F.KIND_POTENTIAL, "potential", 1.0)):
        plan.spectrum_to_field(spec, field, kind=kind, scale=scale)
        assert rel_l2(download_planes(plan, field, N, xs), scale * ref["fields"][key]) < TOL, key
    plan.close()


def test_1024_headline_pipeline_vs_oracle(gpu):
    """The bench.py workload itself: 1024^3, 2 Gpc, z = 0.8, filter of tests/test_box.py:88-90, 50 bins."""
    N, L, seed, nb = 1024, (2e3, 2e3, 2e3), 11, 50
    _, pkf = pk_function(0.8)
    plan, edges = setup_plan(N, L, 0.8, nbins=nb, filt=transfer_fn)
    d_re, d_im = upload_noise(plan, seed, N)
    field = plan.alloc(N ** 3 * 4)
    flags = F.F_SQRTPK | F.F_FILTER
    res, sums = plan.realise(d_re, d_im, flags=flags, field_out=field, want_pk=True)
    xs = list(range(0, N, 73)) + [N - 1]                      # 16 planes incl. both ends
    got = download_planes(plan, field, N, xs)
    fwd = plan.field_to_spectrum(field, want_pk=True)
    plan.close()
    del d_re, d_im, field
    ref = R.realise_slabwise(lambda a: noise_plane(seed, a, N), pkf, N, *L, transfer_fn=transfer_fn, nbins=nb,
                             x_planes=xs, workers=WORKERS)
    assert rel_l2(got, ref["fields"][None]) < TOL
    for j in range(len(xs)):                                 # no single plane hides behind the others
        assert rel_l2(got[j], ref["fields"][None][j]) < 2 * TOL, xs[j]
    assert np.array_equal(res["count"][:nb].astype(np.int64), ref["count"][:nb])
    assert int(res["count"].sum()) == N ** 3
    ref_pk = ks.moments_to_spectrum(edges, ref["count"].astype(np.uint64), ref["sum1"], ref["sum2"])
    assert_pk_close(ks.moments_to_spectrum(edges, res["count"], res["sum1"], res["sum2"]), ref_pk, floor_rel=F32_FLOOR)
    assert np.array_equal(fwd["count"], res["count"])
    assert_pk_close(ks.moments_to_spectrum(edges, fwd["count"], fwd["sum1"], fwd["sum2"]), ref_pk, tol=2 * TOL,
                    floor_rel=FWD_FLOOR)
    # Parseval against the oracle's float64 spectrum (box.py:944-946)
    total_k = float(ref["sum1"].sum()) * R.boxfactor(N, *L)
    assert abs(sums[1] * N ** 3 / total_k - 1) < 1e-5


def test_2048_kspace_kernels_on_slabs_vs_oracle(gpu):
    """k_rows_inv<2048, Philox> + k_cols_c2c<2048> + fused moments on planes 0, 1 and N/2 of a 2048^3 box."""
    N, L, seed, nb = 2048, (4e3, 4e3, 4e3), 0x5EED, 50
    _, pkf = pk_function(0.8)
    plan, edges = setup_plan(N, L, 0.8, nbins=nb, filt=transfer_fn)
    bf = R.boxfactor(N, *L)
    m = R.mode_numbers(N).astype(np.float64)
    # the Gaussian k_perp filter puts the Nyquist plane ~50 orders of magnitude below float32's range (it
    # comes out as exact zeros): that plane is checked without the filter
    for a0, na, filt in ((0, 2, True), (N // 2, 1, False)):
        plan.set_slab(a0, na, 0, N)
        work = plan.alloc(na * N * N * 8)
        res = plan.realise_local_kspace(seed, F.F_SQRTPK | (F.F_FILTER if filt else 0), work, None, 0, want_pk=True)
        got = plan.download(work, (na, N, N), np.complex64)
        cnt = np.zeros(nb + 1)
        s1 = np.zeros(nb + 1)
        for i in range(na):
            a = a0 + i
            idx = (np.uint64(a) * np.uint64(N) + np.arange(N, dtype=np.uint64)[:, None]) * np.uint64(N) \
                + np.arange(N, dtype=np.uint64)[None, :]
            re, im = R.philox_normals(seed, idx, N)          # Hermitian white noise H0 of this plane
            k = R.k_plane(a, N, *L)
            amp = np.sqrt(np.nan_to_num(pkf(k.ravel()).reshape(k.shape)) * bf)
            kperp = 2 * np.pi * np.sqrt((m[a] / L[0]) ** 2. + (m[:, None] / L[1]) ** 2.)
            if filt:
                amp = amp * np.nan_to_num(transfer_fn(np.broadcast_to(kperp, (N, N)), 2 * np.pi * m[None, :] / L[2]))
            H = (re + 1j * im) * amp
            want = np.fft.ifft2(H) * (N * N)                 # z rows and y columns, unnormalised
            assert rel_l2(got[i], want) < TOL, a
            w = 1.0 if a in (0, N // 2) else 2.0
            c, t1, _ = R.pk_moments((H * np.conj(H)).real.ravel() / bf, np.digitize(k.ravel(), edges), nb,
                                    np.full(N * N, w))
            cnt += c
            s1 += t1
        assert np.array_equal(res["count"][:nb + 1].astype(np.int64), np.rint(cnt).astype(np.int64))
        ok = cnt[:nb] > 0
        assert np.all(np.abs(res["sum1"][:nb][ok] - s1[:nb][ok]) <= TOL * np.abs(s1[:nb][ok]) + 1e-12 * np.abs(s1).max())
        work.free()
    plan.close()
