"""
GPU parity of the individual FFT passes against numpy.fft (the library the
reference calls at fastbox/box.py:187,193,736), through the C ABI.
Tolerance: 1e-5 relative L2 (BASELINE.json north_star) -- observed ~1e-7.
"""
import numpy as np
import pytest

from fastbox_b200 import _lib

pytestmark = pytest.mark.gpu

SIZES = [8, 16, 32, 64, 128, 256]
TOL = 1e-5


from _util import rel_l2


@pytest.mark.parametrize("N", SIZES + [512, 1024, 2048])
@pytest.mark.parametrize("sign", [-1, 1])
def test_rows_and_cols_c2c(gpu, N, sign):
    rng = np.random.default_rng(N + sign)
    planes = 3
    x = (rng.standard_normal((planes, N, N)) + 1j * rng.standard_normal((planes, N, N))).astype(np.complex64)
    plan = _lib.Plan(N, 100., 100., 100.)
    for axis_pass, axis in ((0, 2), (1, 1)):
        buf = plan.upload(x)
        plan.fft_pass_c2c(buf, planes, axis_pass, sign)
        plan.sync()
        got = plan.download(buf, x.shape, np.complex64)
        ref = np.fft.fft(x.astype(np.complex128), axis=axis) if sign < 0 else \
            np.fft.ifft(x.astype(np.complex128), axis=axis) * N
        assert rel_l2(got, ref) < TOL, (N, sign, axis_pass)
    plan.close()


@pytest.mark.parametrize("N", SIZES + [512, 1024, 2048])
def test_x_real_passes(gpu, N):
    rng = np.random.default_rng(N)
    ncols = 64 if N <= 64 else 96
    ncols = max(ncols, 32)
    ncols = (ncols // 32) * 32
    f = rng.standard_normal((N, ncols)).astype(np.float32)
    plan = _lib.Plan(N, 100., 100., 100.)
    dfield = plan.upload(f)
    dspec = plan.alloc((N // 2 + 1) * ncols * 8)
    plan.fft_pass_x_r2c(dfield, dspec, ncols)
    plan.sync()
    spec = plan.download(dspec, (N // 2 + 1, ncols), np.complex64)
    ref = np.fft.rfft(f.astype(np.float64), axis=0)
    assert rel_l2(spec, ref) < TOL
    # and back
    dback = plan.alloc(N * ncols * 4)
    s1, s2 = plan.fft_pass_x_c2r(dspec, dback, ncols, scale=1.0 / N)
    back = plan.download(dback, (N, ncols), np.float32)
    assert rel_l2(back, f.astype(np.float64)) < TOL
    assert abs(s1 - back.astype(np.float64).sum()) < 1e-3 * max(1.0, abs(s1)) + 1e-2
    assert abs(s2 - (back.astype(np.float64) ** 2).sum()) < 1e-4 * s2
    plan.close()


@pytest.mark.parametrize("N", [16, 64, 128])
def test_3d_roundtrip_matches_numpy(gpu, N):
    """field -> half spectrum (vs rfftn over axes (1,2,0)) -> field."""
    rng = np.random.default_rng(7)
    f = rng.standard_normal((N, N, N)).astype(np.float32)
    plan = _lib.Plan(N, 100., 100., 100.)
    spec = np.empty((N // 2 + 1, N, N), dtype=np.complex64)
    plan.field_to_spectrum(f, spec_out=spec)
    ref = np.fft.fftn(f.astype(np.float64))[:N // 2 + 1]
    assert rel_l2(spec, ref) < TOL
    back = np.empty((N, N, N), dtype=np.float32)
    plan.spectrum_to_field(spec, back)
    assert rel_l2(back, f.astype(np.float64)) < TOL
    plan.close()


def test_y_pass_2048_wide_tiles(gpu, monkeypatch):
    """k_cols_c2c<2048, 8>: the 64-byte-row variant the multi-GPU exchange uses for its NVLink stores."""
    monkeypatch.setenv("FB_CZ_COLS", "8")
    N, planes = 2048, 2
    rng = np.random.default_rng(5)
    x = (rng.standard_normal((planes, N, N)) + 1j * rng.standard_normal((planes, N, N))).astype(np.complex64)
    plan = _lib.Plan(N, 100., 100., 100.)
    buf = plan.upload(x)
    plan.fft_pass_c2c(buf, planes, 1, +1)
    plan.sync()
    got = plan.download(buf, x.shape, np.complex64)
    assert rel_l2(got, np.fft.ifft(x.astype(np.complex128), axis=1) * N) < TOL
    plan.close()


@pytest.mark.parametrize("N,cz", [(512, 8), (512, 16), (1024, 4), (1024, 8), (2048, 4)])
@pytest.mark.parametrize("sign", [-1, 1])
def test_y_pass_tma_pipelined(gpu, monkeypatch, N, cz, sign):
    """Persistent TMA-pipelined y pass (fb_cols_tma.cu) == numpy.fft and == the per-thread kernel bit for bit."""
    planes = 5                                      # 5 * N/cz tiles: several tiles per CTA, both buffers reused
    rng = np.random.default_rng(N + cz)
    x = (rng.standard_normal((planes, N, N)) + 1j * rng.standard_normal((planes, N, N))).astype(np.complex64)
    plan = _lib.Plan(N, 100., 100., 100.)
    out = {}
    for tma in ("1", "0"):
        monkeypatch.setenv("FB_COLS_TMA", tma)
        monkeypatch.setenv("FB_CZ_TMA", str(cz))
        buf = plan.upload(x)
        plan.fft_pass_c2c(buf, planes, 1, sign)
        plan.sync()
        out[tma] = plan.download(buf, x.shape, np.complex64)
    ref = np.fft.fft(x.astype(np.complex128), axis=1) if sign < 0 else np.fft.ifft(x.astype(np.complex128), axis=1) * N
    assert rel_l2(out["1"], ref) < TOL
    assert np.array_equal(out["1"], out["0"])
    plan.close()
