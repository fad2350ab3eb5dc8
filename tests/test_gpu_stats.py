"""
SURVEY 8(f) rank 4: P(k_perp, |k_par|) and xi(r) on the device against oracle/restate.py.

The reference delegates both to nbodykit (examples/example_endtoend.py:128-151), which is neither vendored nor
pinned: parity here is pinned to the NumPy restatement only (the restatement itself is tied to the pinned 1-D
estimator and to a direct pair average in tests/test_oracle_cpu.py).
Bin populations are integer work -> bit-exact; moments within the north_star's 1e-5.
"""
import numpy as np
import pytest

from fastbox_b200 import kspace as ks
from fastbox_b200.box import CosmoBox, default_cosmo
from oracle import restate as R

from _util import TOL, rel_l2, setup_plan

pytestmark = pytest.mark.gpu


def _field(N, seed):
    rng = np.random.default_rng(seed)
    f = rng.standard_normal((N, N, N))
    # some large-scale structure so that xi(r) is not pure noise; the spectrum spans two decades only, so every
    # bin stays well inside float32's 1e-7 of the largest mode
    k2 = np.fft.fftfreq(N)[:, None, None] ** 2 + np.fft.fftfreq(N)[None, :, None] ** 2 + np.fft.fftfreq(N)[None, None, :] ** 2
    return np.fft.ifftn(np.fft.fftn(f) / np.sqrt(1. + k2 / 0.01)).real.astype(np.float32).astype(np.float64)


@pytest.mark.parametrize("N,L", [(16, (1e2, 1e2, 1e2)), (32, (1e2, 1.5e2, 2e2)), (64, (3e2, 2e2, 1e2))])
def test_pk2d_matches_oracle(gpu, N, L):
    plan, _ = setup_plan(N, L, 0.8)
    f = _field(N, N)
    g = _field(N, N + 1)
    half = R.rfft3_axis0(f)
    halfg = R.rfft3_axis0(g)
    kp = np.concatenate([[0.0], np.logspace(np.log10(2 * np.pi / max(L[:2])), np.log10(1.5 * np.pi * N / min(L[:2])), 12)])
    kl = np.concatenate([[0.0], np.logspace(np.log10(2 * np.pi / L[2]), np.log10(np.pi * N / L[2]), 9)])
    m = R.mode_numbers(N).astype(np.float64)
    ipar = np.digitize(np.abs(2. * np.pi * m / L[2]), kl)
    thr = ks.bin_thresholds(kp)
    d_f = plan.upload_f32(f)
    spec = plan.alloc((N // 2 + 1) * N * N * 8)
    plan.field_to_spectrum(d_f, spec_out=spec)
    for other, hb in ((None, None), (plan.upload(halfg.astype(np.complex64)), halfg)):
        res = plan.pk2d_from_spectrum(spec, thr, ipar, kl.size, cross=other)
        _, _, mean, err, cnt = R.binned_power_spectrum_2d_lean(half, N, *L, kp, kl, half_b=hb)
        assert np.array_equal(res["count"].astype(np.int64), cnt)                    # bit-exact bin populations
        assert int(res["count"].sum()) == N ** 3
        c = res["count"].astype(np.float64)
        with np.errstate(all="ignore"):
            got = (res["sum1"] / c)[1:kp.size, 1:kl.size]
            gerr = (np.sqrt(np.maximum(res["sum2"] / c - (res["sum1"] / c) ** 2, 0.0)) / np.sqrt(c))[1:kp.size, 1:kl.size]
        assert np.array_equal(np.isnan(got), np.isnan(mean))
        ok = ~np.isnan(mean)
        scale = np.abs(mean[ok]).max()
        if hb is None:
            assert np.all(np.abs(got[ok] - mean[ok]) <= TOL * np.abs(mean[ok]))
        else:                                   # a cross spectrum changes sign: compare against the bin's scatter
            assert np.all(np.abs(got[ok] - mean[ok]) <= TOL * (np.abs(mean[ok]) + err[ok] * np.sqrt(cnt[1:kp.size, 1:kl.size][ok])))
        # error bars: a bin holding one Hermitian pair has stddev == 0 exactly on the device and float64 rounding
        # noise (~1e-8 of the bin's power) in NumPy's two-pass formula: floor tied to the bin's own power
        assert np.all(np.abs(gerr[ok] - err[ok]) <= 10 * TOL * err[ok] + 1e-6 * np.abs(mean[ok]) + 1e-9 * scale)
    # full-cube input gives the same table (every mode counted once instead of twice)
    full = np.fft.fftn(f).astype(np.complex64)
    res_full = plan.pk2d_from_spectrum(plan.upload(full), thr, ipar, kl.size, full_cube=True)
    res_half = plan.pk2d_from_spectrum(spec, thr, ipar, kl.size)
    assert np.array_equal(res_full["count"], res_half["count"])
    assert np.allclose(res_full["sum1"], res_half["sum1"], rtol=1e-5, atol=1e-5 * np.abs(res_half["sum1"]).max())


@pytest.mark.parametrize("N,L", [(16, (1e2, 1e2, 1e2)), (32, (1e2, 1.5e2, 2e2)), (64, (2e2, 2e2, 2e2))])
def test_correlation_function_matches_oracle(gpu, N, L):
    plan, _ = setup_plan(N, L, 0.8)
    f = _field(N, 3 * N)
    g = _field(N, 3 * N + 1)
    h = min(L) / N
    edges = np.concatenate([[0.0, 0.5 * h], np.arange(1.5 * h, 0.45 * max(L), 1.7 * h)])
    d_f, d_g = plan.upload_f32(f), plan.upload_f32(g)
    xi_cube = plan.alloc(N ** 3 * 4)
    for db, hb in ((None, None), (d_g, g)):
        res = plan.correlation_function(d_f, edges, field_b=db, xi_out=xi_cube)
        cent, mean, err, cnt = R.correlation_function_port(f, N, *L, edges, field_b=hb)
        assert np.array_equal(res["count"][:edges.size].astype(np.int64), cnt[:edges.size])  # lags per shell: bit-exact
        A = np.fft.fftn(f)
        B = A if hb is None else np.fft.fftn(hb)
        want_cube = np.fft.ifftn(A * np.conj(B)).real / float(N) ** 3
        assert rel_l2(plan.download(xi_cube, (N, N, N), np.float32), want_cube) < TOL
        c = res["count"][1:edges.size].astype(np.float64)
        got = res["sum1"][1:edges.size] / c
        ok = cnt[1:edges.size] > 0
        # xi(r) crosses zero: per-bin error relative to the rms of the lag cube over the square root of the shell size
        rms = np.sqrt(np.mean(want_cube ** 2))
        assert np.all(np.abs(got[ok] - mean[ok]) <= TOL * (np.abs(mean[ok]) + rms))
        assert rel_l2(got[ok], mean[ok]) < TOL
    # zero lag = variance of the field
    res = plan.correlation_function(d_f, edges)
    assert res["count"][1] == 1 and abs(res["sum1"][1] - np.mean(f * f)) < TOL * np.mean(f * f)


def test_shim_two_point_statistics(gpu):
    """CosmoBox.binned_power_spectrum_2d / correlation_function on a realised box, against the restatement."""
    np.random.seed(5)
    N, L = 32, (2e2, 2e2, 3e2)
    box = CosmoBox(cosmo=default_cosmo, box_scale=L, nsamp=N, realise_now=True)
    f = box.delta_x.astype(np.float32).astype(np.float64)
    kperp_c, kpar_c, pk2d, err2d = box.binned_power_spectrum_2d(delta_x=box.delta_x, nbins=(10, 8))
    kp = np.logspace(np.log10(2. * np.pi / max(L[:2])), np.log10(np.sqrt(2.) * np.pi * N / min(L[:2])), 10)
    kl = np.logspace(np.log10(2. * np.pi / L[2]), np.log10(np.pi * N / L[2]), 8)
    cp, cl, mean, err, _ = R.binned_power_spectrum_2d_lean(R.rfft3_axis0(f), N, *L, kp, kl)
    assert np.allclose(kperp_c, cp) and np.allclose(kpar_c, cl)
    assert pk2d.shape == mean.shape == (9, 7)
    assert np.array_equal(np.isnan(pk2d), np.isnan(mean))
    ok = ~np.isnan(mean)
    assert np.all(np.abs(pk2d[ok] - mean[ok]) <= TOL * np.abs(mean[ok]))
    # collapsing P(k_perp, k_par) with its populations gives back total power (Parseval), as does the 1-D spectrum
    r, xi, xerr = box.correlation_function(delta_x=box.delta_x, dr=10., rmin=10., rmax=120.)
    cent, want, werr, cnt = R.correlation_function_port(f, N, *L, np.arange(10., 125., 10.))
    assert np.allclose(r, cent)
    assert rel_l2(xi, want) < TOL
    assert np.all(np.abs(xerr - werr) <= 10 * TOL * werr + 1e-9 * np.abs(want).max())
    with pytest.raises(ValueError):
        box.binned_power_spectrum_2d(delta_x=box.delta_x, delta_k=box.delta_k)


def test_device_multipoles_kaiser_and_cross_at_256(gpu):
    """
    Multipoles and cross power are unpinned w.r.t. nbodykit (absent), so the DEVICE path is tied to known answers as
    well: a field with spectrum (1 + beta mu^2) d(k), |d|^2 = boxfactor, must give the Kaiser / Hamilton multipoles
    P_0 = 1 + 2 beta/3 + beta^2/5, P_2 = 4 beta/3 + 4 beta^2/7, P_4 = 8 beta^2/35 (to the quadrature error of the
    discrete shells, which falls as 1/N^2), and the NumPy restatement to the usual tolerance; cross(a, c a) = c auto(a).
    """
    N, L = 256, (5e2, 5e2, 5e2)
    rng = np.random.default_rng(5)
    half = R.rfft3_axis0(rng.standard_normal((N, N, N)))
    half = half / np.maximum(np.abs(half), 1e-300) * np.sqrt(R.boxfactor(N, *L))
    m = R.mode_numbers(N).astype(np.float64)
    h = N // 2 + 1
    kk = np.sqrt(m[:h, None, None] ** 2 + m[None, :, None] ** 2 + m[None, None, :] ** 2)
    with np.errstate(all="ignore"):
        mu = np.where(kk > 0, m[None, None, :] / kk, 0.0)
    del kk
    beta = 0.6
    half_s = half * (1.0 + beta * mu ** 2)
    del mu
    # a real field needs real self-conjugate modes; the planes a = 0 and a = N/2 are made Hermitian by the
    # transform pair itself (c2r drops the inconsistent parts), so measure what the device sees: go through the field
    field = R.irfft3_axis0(half_s).astype(np.float32)
    half_seen = R.rfft3_axis0(field.astype(np.float64))
    knyq = np.pi * N / L[0]
    edges = np.concatenate([[0.0], np.linspace(0.3 * knyq, 0.98 * knyq, 5)])
    plan = setup_plan(N, L, 0.0, nbins=20)[0]
    plan.set_pk_bins(ks.bin_thresholds(edges))
    res = plan.field_to_spectrum(field, want_pk=True, poles=True)
    cnt = res["count"][:edges.size].astype(np.float64)
    with np.errstate(all="ignore"):                             # bin 0 (below the first edge, k = 0 only) is unused
        p0 = (res["sum1"][:edges.size] / cnt)[2:]               # bins between the four upper edges
        p2 = (5.0 * res["sum_l2"][:edges.size] / cnt)[2:]
        p4 = (9.0 * res["sum_l4"][:edges.size] / cnt)[2:]
    _, poles = R.pk_multipoles(half_seen, N, *L, kbins=edges)
    for got, ell in ((p0, 0), (p2, 2), (p4, 4)):
        assert np.all(np.abs(got - poles[ell][1:]) <= 20 * TOL * np.abs(poles[0][1:])), ell
    want = {0: 1 + 2 * beta / 3 + beta ** 2 / 5, 2: 4 * beta / 3 + 4 * beta ** 2 / 7, 4: 8 * beta ** 2 / 35}
    for got, ell, tol in ((p0, 0, 2e-3), (p2, 2, 4e-3), (p4, 4, 8e-3)):
        assert np.all(np.abs(got - want[ell]) < tol), (ell, got, want[ell])
    # cross power of the field with 2.5 x itself = 2.5 x its auto power, bin by bin (and the same populations)
    spec_b = plan.alloc((N // 2 + 1) * N * N * 8)
    plan.field_to_spectrum((2.5 * field).astype(np.float32), spec_out=spec_b)
    resx = plan.field_to_spectrum(field, cross=spec_b, want_pk=True)
    assert np.array_equal(resx["count"], res["count"])
    ok = res["count"][:edges.size] > 0
    assert np.allclose(resx["sum1"][:edges.size][ok], 2.5 * res["sum1"][:edges.size][ok], rtol=1e-6)
    plan.close()
