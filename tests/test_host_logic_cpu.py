"""CPU tests of the host-side logic of the drop-in (no GPU): tables, bin thresholds, API surface."""
import ctypes
import os
import re

import numpy as np
import pytest

import fastbox_b200 as fb
from fastbox_b200 import _lib
from fastbox_b200 import kspace as ks
from oracle import restate as R

from _util import pk_function, transfer_fn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("N,L", [(64, (1e3, 1e3, 1e3)), (32, (1e2, 2e2, 1e3)), (128, (2e3, 2e3, 2e3))])
@pytest.mark.parametrize("nbins", [20, 50])
def test_bin_thresholds_reproduce_digitize_bit_exact(N, L, nbins):
    """Integer work must be bit exact: thresholds on s == np.digitize on the reference's k."""
    kmin, kmax = R.kmin_kmax(N, *L)
    edges = ks.pk_bin_edges(kmin, kmax, nbins)
    assert np.array_equal(edges, R.pk_bin_edges(N, *L, nbins))
    thr = ks.bin_thresholds(edges)
    s = (ks.axis_sq(N, L[0])[:, None, None] + ks.axis_sq(N, L[1])[None, :, None]) + ks.axis_sq(N, L[2])[None, None, :]
    assert np.array_equal(2. * np.pi * np.sqrt(s), R.k_grid(N, *L))                 # same float64 bits as box.k
    idx = np.searchsorted(thr, s.ravel(), side="right")
    assert np.array_equal(idx, R.digitize_modes(N, *L, edges))
    # thresholds are tight: the previous double maps below the edge
    below = np.nextafter(thr, -np.inf)
    ok = thr > 0
    assert np.all(ks.k_of_s(thr[ok]) >= edges[ok]) and np.all(ks.k_of_s(below[ok]) < edges[ok])


def test_custom_bins_and_degenerate_edges():
    edges = np.array([0.0, 0.01, 0.01, 0.5, 3.0])
    thr = ks.bin_thresholds(edges)
    k = np.array([0.0, 0.005, 0.01, 0.2, 0.5, 2.9999, 3.0, 10.0])
    s = (k / (2 * np.pi)) ** 2
    # round trip through sqrt may move a point by 1 ulp; use the mapped k for the reference side
    assert np.array_equal(np.searchsorted(thr, s, side="right"), np.digitize(ks.k_of_s(s), edges))


def test_sqrt_pk_tables():
    _, pkf = pk_function(0.5)
    N, L = 32, 500.0
    bf = R.boxfactor(N, L, L, L)
    lut = ks.sqrt_pk_int_lut(pkf, N, L, bf)
    assert lut[0] == 0.0 and lut.size == 3 * (N // 2) ** 2 + 1
    ref = R.sqrt_pk_half(pkf, N, L, L, L)
    m = ks.mode_numbers(N)
    n2 = (m[:N // 2 + 1, None, None] ** 2 + m[None, :, None] ** 2 + m[None, None, :] ** 2)
    assert np.allclose(lut[n2], ref, rtol=2e-7)
    tab, l0, dl = ks.sqrt_pk_log_table(pkf, 16, 1e2, 2e2, 1e3, R.boxfactor(16, 1e2, 2e2, 1e3))
    ref = R.sqrt_pk_half(pkf, 16, 1e2, 2e2, 1e3)
    mm = ks.mode_numbers(16).astype(np.float64)
    s = (mm[:9, None, None] / 1e2) ** 2 + (mm[None, :, None] / 2e2) ** 2 + (mm[None, None, :] / 1e3) ** 2
    interp = ks.eval_log_table(tab, l0, dl, s)
    ok = s > 0
    assert np.allclose(interp[ok], ref[ok], rtol=2e-6)
    assert np.all(interp[~ok] == 0.0)
    # the chooser validates the interpolated table and keeps small cubic grids exact
    assert ks.choose_sqrt_pk_table(pkf, 32, 500., 500., 500., bf)[0] == 1
    mode, t2, base, M = ks.choose_sqrt_pk_table(pkf, 32, 500., 500., 500., bf, exact_below=0)
    assert mode == 3 and M == 9.0                          # float-bit table, 9 mantissa bits
    n2 = np.arange(1, 3 * 16 ** 2 + 1, dtype=np.float64)
    exact = np.sqrt(pkf(2 * np.pi * np.sqrt(n2) / 500.) * bf)
    assert np.allclose(ks.eval_bit_table(t2, int(base), int(M), n2 / 500. ** 2), exact, rtol=2e-6)
    spiky = lambda k: pkf(k) * (1.0 + 0.9 * np.sin(k * 2.0e4))          # cannot be interpolated: falls back
    assert ks.choose_sqrt_pk_table(spiky, 32, 500., 500., 500., bf, exact_below=0)[0] == 1


def test_log_bin_model():
    thr = ks.bin_thresholds(ks.pk_bin_edges(2 * np.pi / 2e3, 2 * np.pi * np.sqrt(3.) * 1024 / 2e3, 50))
    l0, inv = ks.log_bin_model(thr)
    assert inv > 0
    s = np.exp(np.random.RandomState(0).uniform(np.log(thr[0]) - 1, np.log(thr[-1]) + 1, 100000))
    guess = np.clip(np.floor((np.log2(s.astype(np.float32)) - np.float32(l0)) * np.float32(inv)).astype(int) + 1, 0, 50)
    exact = np.searchsorted(thr, s, side="right")
    assert np.max(np.abs(guess - exact)) <= 1               # the device corrects by at most one step
    assert ks.log_bin_model(np.array([0.0, 1.0, 2.0])) == (0.0, 0.0)
    assert ks.log_bin_model(np.array([1.0, 2.0, 3.0, 10.0])) == (0.0, 0.0)


def test_filter_tables_separable_and_dense():
    N, L = 16, (1e2, 1e2, 1e2)
    ft = ks.filter_tables(transfer_fn, N, *L)
    assert ft.tdense is None and ft.even
    kperp, kpar = R.kperp_kpar(N, *L)
    full = transfer_fn(kperp[:N // 2 + 1], kpar)
    assert np.allclose(ft.tperp[:, :, None] * ft.tpar[None, None, :], full, rtol=2e-6, atol=1e-30)
    wedge = lambda kp, kl: (np.abs(kl) > 0.5 * kp).astype(float)
    fd = ks.filter_tables(wedge, N, *L)
    assert fd.tdense is not None and fd.even
    assert np.array_equal(fd.tdense, wedge(kperp[:N // 2 + 1], kpar + 0 * kperp[:N // 2 + 1]).astype(np.float32))
    odd = lambda kp, kl: 1.0 + 0.5 * np.tanh(kl / 0.2)
    assert not ks.filter_tables(odd, N, *L).even
    nanfn = lambda kp, kl: np.sin(kp) / kp
    assert np.isfinite(ks.filter_tables(nanfn, N, *L, force_dense=True).tdense).all()     # box.py:379


def test_moments_to_spectrum_matches_reference_estimator():
    rng = np.random.default_rng(0)
    N, L = 16, (1e2, 1e2, 1e2)
    dk = np.fft.fftn(rng.standard_normal((N, N, N)))
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        kc, pk, err, cnt, idx = R.binned_power_spectrum_port(dk, N, *L, nbins=20, return_raw=True)
        p = (dk * np.conj(dk)).real.ravel() / R.boxfactor(N, *L)
        c, s1, s2 = R.pk_moments(p, idx, 20)
        got = ks.moments_to_spectrum(R.pk_bin_edges(N, *L, 20), c, s1, s2)
    assert np.array_equal(got[0], kc)
    m = ~np.isnan(pk)
    assert np.array_equal(np.isnan(got[1]), np.isnan(pk))
    assert np.allclose(got[1][m], pk[m], rtol=1e-13)
    assert np.allclose(got[2][m], err[m], rtol=1e-6, atol=1e-9 * np.abs(pk[m]).max())


def test_cosmobox_host_surface_matches_reference_semantics():
    """Reference tests/test_box.py: test_box_errors, coordinates, lengths (CPU-only parts)."""
    with pytest.raises(TypeError):
        fb.CosmoBox(cosmo=[0.7, 0.3], box_scale=(1e2, 1e2, 1e2), nsamp=16, realise_now=False)
    box = fb.CosmoBox(cosmo=fb.default_cosmo, box_scale=(1e3, 1e3, 1e3), nsamp=16, realise_now=False, redshift=0.8)
    assert box.Lx == box.Ly == box.Lz == 1e3
    assert box.x.size == 16 and np.isclose(np.max(box.x) - np.min(box.x), 1e3)
    ang_x, ang_y = box.pixel_array()
    ang_x2, _ = box.pixel_array(redshift=0.82)
    assert np.isclose(ang_x[1] - ang_x[0], ang_y[1] - ang_y[0])
    assert ang_x[1] - ang_x[0] > ang_x2[1] - ang_x2[0]
    assert np.all(np.diff(box.freq_array()) < 0.) and np.all(np.diff(box.freq_array(redshift=2.)) < 0.)
    assert np.array_equal(box.k, R.k_grid(16, 1e3, 1e3, 1e3))
    assert np.array_equal(box.Kz[0, 0], ks.mode_numbers(16))
    with pytest.raises(ValueError):
        box.binned_power_spectrum(delta_x=np.zeros((16,) * 3), delta_k=np.zeros((16,) * 3))
    with pytest.raises(AssertionError):
        fb.CosmoBox(cosmo=fb.default_cosmo, box_scale=(1e2, 1e2), nsamp=16, realise_now=False)
    tr = fb.tracers.HITracer(box)
    assert np.isclose(tr.bias_HI(0.8), 0.84081272)
    assert np.isclose(fb.tracers.TracerModel(box).linear_bias(1.5, 3.0), 3.0)


def test_golden_coordinates_match_reference():
    from _util import load_golden
    g = load_golden("n32_gpc")
    box = fb.CosmoBox(cosmo=fb.default_cosmo, box_scale=1e3, nsamp=32, realise_now=False, redshift=0.8)
    assert np.allclose(box.freq_array(), g["freq"], rtol=1e-13)
    assert np.allclose(box.pixel_array()[0], g["pix_x"], rtol=1e-13)
    assert box.boxfactor == float(g["boxfactor"])
    assert np.isclose(fb.tracers.HITracer(box).signal_amplitude(), float(g["Tb"]))


def test_c_abi_library_loads_and_exports_every_declared_symbol():
    lib = _lib.load()
    hdr = open(os.path.join(ROOT, "include", "fastbox_b200.h")).read()
    declared = set(re.findall(r"\b(fb_[a-z0-9_]+)\s*\(", hdr))
    declared -= {"fb_plan", "fb_pk_result"}
    assert declared, "no declarations parsed"
    for name in sorted(declared):
        assert hasattr(lib, name), "library does not export %s" % name
        assert name in _lib.SIGNATURES, "ctypes binding missing for %s" % name
    assert lib.fb_version().startswith(b"fastbox_b200")


def test_no_cpu_fallback_without_gpu():
    """Without a CUDA device the product path must fail loudly, never compute on the CPU."""
    try:
        _lib.device_info(0)
        pytest.skip("a GPU is present")
    except _lib.FastBoxError:
        pass
    with pytest.raises(_lib.FastBoxError):
        _lib.Plan(16, 1e2, 1e2, 1e2)
    box = fb.CosmoBox(cosmo=fb.default_cosmo, box_scale=1e2, nsamp=16, realise_now=False)
    with pytest.raises(_lib.FastBoxError):
        box.realise_density()
    # the rows added from SURVEY 8(f) have no CPU path either
    cube = np.zeros((16, 16, 16))
    with pytest.raises(_lib.FastBoxError):
        fb.filters.mean_spectrum_filter(cube)
    with pytest.raises(_lib.FastBoxError):
        fb.filters.pca_filter(cube, nmodes=2)
    with pytest.raises(_lib.FastBoxError):
        fb.halos.HaloDistribution(box, (1e12, 1e15), 10).realise_halo_catalogue(np.ones((16, 16, 16), int))
    with pytest.raises(_lib.FastBoxError):
        fb.foregrounds.ForegroundModel(box).construct_cube(np.ones((16, 16)), -2.7)
    with pytest.raises(_lib.FastBoxError):
        fb.noise.NoiseModel(box).realise_radiometer_noise(18., 2., 1., 64, seed=1)


def test_host_maps_of_the_foreground_model_need_no_gpu():
    """The N^2 maps are host work with the reference's own NumPy / SciPy calls (foregrounds.py:48-149):
    bit-identical to the unmodified reference's golden output under the same seed."""
    g = np.load(os.path.join(ROOT, "tests", "golden", "fg_noise_cube.npz"))
    box = fb.CosmoBox(cosmo=fb.default_cosmo, box_scale=tuple(g["scale"]), nsamp=int(g["N"]),
                      redshift=float(g["redshift"]), realise_now=False)
    fg = fb.foregrounds.ForegroundModel(box)
    np.random.seed(77)
    assert np.array_equal(fg.realise_foreground_amp(amp=57., beta=-1.1, monopole=10., smoothing_scale=4.), g["amps"])
    assert np.array_equal(fg.realise_foreground_amp(amp=57., beta=-1.1, monopole=10.), g["amps_nosmooth"])
    assert np.array_equal(fg.realise_spectral_index(mean_spec_idx=-2.07, std_spec_idx=0.2, smoothing_scale=15.),
                          g["alpha"])
    sig = fb.noise.NoiseModel(box).radiometer_rms(18., 2.5, 1., 64)
    assert sig.shape == (int(g["N"]),) and np.all(sig > 0)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "fastbox_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in txt.replace("oracle/", "oracle/") or "import oracle" not in txt
                assert "from oracle" not in txt and "import oracle" not in txt


def test_bin_thresholds_property_random_edges_and_boxes():
    """Property test (hypothesis): for arbitrary increasing edges and box lengths, the sqrt-free threshold search
    gives np.digitize's index for every mode of a small grid and for values placed on and next to every edge."""
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=60, deadline=None)
    @given(st.lists(st.floats(min_value=1e-4, max_value=50.0, allow_nan=False), min_size=2, max_size=24, unique=True),
           st.tuples(st.floats(10.0, 5e3), st.floats(10.0, 5e3), st.floats(10.0, 5e3)), st.booleans())
    def check(edges, L, with_zero):
        edges = np.sort(np.array(edges))
        if with_zero:
            edges = np.concatenate([[0.0], edges])
        thr = ks.bin_thresholds(edges)
        N = 8
        s = (ks.axis_sq(N, L[0])[:, None, None] + ks.axis_sq(N, L[1])[None, :, None]) + ks.axis_sq(N, L[2])[None, None, :]
        k = 2. * np.pi * np.sqrt(s)
        assert np.array_equal(np.searchsorted(thr, s.ravel(), side="right"), np.digitize(k.ravel(), edges))
        # values whose k lands exactly on an edge, one ulp below and one above
        se = (edges / (2. * np.pi)) ** 2
        for cand in (se, np.nextafter(se, 0.0), np.nextafter(se, np.inf)):
            assert np.array_equal(np.searchsorted(thr, cand, side="right"), np.digitize(ks.k_of_s(cand), edges))
    check()
