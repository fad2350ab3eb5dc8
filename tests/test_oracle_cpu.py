"""
CPU tests (no GPU): pin oracle/restate.py against (a) the golden vectors produced by the
UNMODIFIED reference (tests/golden, oracle/make_golden.py) and (b) the live reference modules
when /root/reference is present.  Also the plain-C oracle vs the NumPy restatement.
"""
import ctypes
import os
import subprocess
import warnings

import numpy as np
import pytest

from oracle import ref_loader
from oracle import restate as R

from _util import draw_noise, load_golden, pk_function, transfer_fn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CASES = ["n16_cubic", "n16_cuboid", "n32_gpc"]


def _L(g):
    return tuple(float(x) for x in g["L"])


@pytest.mark.parametrize("name", CASES)
def test_port_reproduces_reference_golden(name):
    g = load_golden(name)
    N, L = int(g["N"]), _L(g)
    scale = tuple(g["scale"]) if g["scale"].size == 3 else float(g["scale"][0])
    assert R.box_lengths(scale, N) == L
    assert R.boxfactor(N, *L) == float(g["boxfactor"])
    assert R.kmin_kmax(N, *L) == (float(g["kmin"]), float(g["kmax"]))
    re, im = draw_noise(int(g["seed"]), N)
    _, pkf = pk_function(float(g["redshift"]))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        dx, dk = R.realise_density_port(re, im, pkf, N, *L)
        assert np.array_equal(dx, g["delta_x"])                       # bitwise
        assert np.array_equal(dk[:N // 2 + 1], g["delta_k_half"])
        for nb in (20, 50):
            kc, pk, err, cnt, idx = R.binned_power_spectrum_port(dk, N, *L, nbins=nb, return_raw=True)
            assert np.array_equal(kc, g["pk%d_k" % nb])
            assert np.array_equal(pk, g["pk%d_p" % nb], equal_nan=True)
            assert np.array_equal(err, g["pk%d_e" % nb], equal_nan=True)
            assert np.array_equal(np.bincount(idx, minlength=nb + 1), g["pk%d_counts" % nb])
        assert np.array_equal(R.apply_transfer_fn_port(dk, transfer_fn, N, *L).real, g["transfer"])
        assert np.array_equal(R.lognormal(dx * float(g["bias_HI"])), g["lognormal"])
        rs = R.redshift_space_density(g["lognormal"], g["vel_z"], g["z_grid"], float(g["Hz"]))
        assert np.allclose(rs, g["rsd0"], rtol=1e-13, atol=1e-13)
        np.random.seed(int(g["seed"]) + 100)
        vnl = 120. * np.random.normal(0., 1., (N, N, N))
        rs = R.redshift_space_density(g["lognormal"], g["vel_z"], g["z_grid"], float(g["Hz"]), vnl)
        assert np.allclose(rs, g["rsd120"], rtol=1e-13, atol=1e-13)
        # method='nearest' (box.py:403-405): a selection of input values, so bit-for-bit
        gm = load_golden("rsd_methods")
        rs = R.redshift_space_density(g["lognormal"], g["vel_z"], g["z_grid"], float(g["Hz"]), method="nearest")
        assert np.array_equal(rs, gm[name + "_nearest0"])
        rs = R.redshift_space_density(g["lognormal"], g["vel_z"], g["z_grid"], float(g["Hz"]), vnl, method="nearest")
        assert np.array_equal(rs, gm[name + "_nearest120"])
        from _util import DEFAULT_COSMO  # noqa: F401
        bc = R.convolve_fft(__import__("oracle.make_golden", fromlist=["beam_cube"]).beam_cube(N), g["rsd0"])
        assert np.allclose(bc, g["beam_conv"], rtol=1e-12, atol=1e-14)
        assert np.allclose(R.halo_mean_count(dx, 1e-3, 1.2, *L, lognormal_tf=True), g["halo_mean_ln"], rtol=1e-14)
        assert np.allclose(R.halo_mean_count(dx, np.linspace(1e-3, 2e-3, N), 1.2, *L), g["halo_mean_lin"], rtol=1e-14)


@pytest.mark.parametrize("name", CASES)
def test_lean_restatement_matches_port(name):
    g = load_golden(name)
    N, L = int(g["N"]), _L(g)
    re, im = draw_noise(int(g["seed"]), N)
    _, pkf = pk_function(float(g["redshift"]))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        dx, half = R.realise_density_lean(re, im, pkf, N, *L)
        assert np.linalg.norm(dx - g["delta_x"]) / np.linalg.norm(g["delta_x"]) < 1e-14
        assert np.abs(half - g["delta_k_half"]).max() / np.abs(g["delta_k_half"]).max() < 1e-14
        assert np.abs(R.expand_half_axis0(half)[N // 2 + 1:] - np.fft.fftn(g["delta_x"])[N // 2 + 1:]).max() \
            / np.abs(half).max() < 1e-13
        for nb in (20, 50):
            kc, pk, err, cnt = R.binned_power_spectrum_lean(half, N, *L, nbins=nb)
            assert np.array_equal(cnt[:nb], g["pk%d_counts" % nb][:nb])
            ref = g["pk%d_p" % nb]
            m = ~np.isnan(ref)
            assert np.array_equal(np.isnan(pk), np.isnan(ref))
            assert np.all(np.abs(pk[m] - ref[m]) <= 1e-13 * np.abs(ref[m]))
        # Parseval by-product (box.py:944-946)
        s1 = np.sum(dx ** 2.) * N ** 3.
        assert abs(s1 / g["parseval"][1] - 1) < 1e-12


@pytest.mark.skipif(not ref_loader.available(), reason="/root/reference not present (GPU box)")
def test_restatement_against_live_reference():
    """Run the unmodified reference side by side on a configuration the fixtures do not hold."""
    box_m = ref_loader.load("box")
    N, scale, seed, z = 24 if False else 16, (3e2, 2e2, 5e2), 41, 0.4
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        np.random.seed(seed)
        b = box_m.CosmoBox(cosmo=box_m.default_cosmo, box_scale=scale, nsamp=N, redshift=z, realise_now=False)
        b.realise_density()
        re, im = draw_noise(seed, N)
        _, pkf = pk_function(z)
        L = R.box_lengths(scale, N)
        dx, dk = R.realise_density_port(re, im, pkf, N, *L)
        assert np.array_equal(dx, b.delta_x)
        kc, pk, err = b.binned_power_spectrum(nbins=30)
        kc2, pk2, err2 = R.binned_power_spectrum_port(dk, N, *L, nbins=30)
        assert np.array_equal(pk, pk2, equal_nan=True)
        assert np.array_equal(R.k_grid(N, *L), b.k)
        hp = lambda kperp, kpar: 1. - np.exp(-0.5 * (np.abs(kpar) / 0.009) ** 3.)     # example_endtoend.py:133
        assert np.array_equal(R.apply_transfer_fn_port(dk, hp, N, *L), b.apply_transfer_fn(b.delta_k, hp))
        assert np.allclose(R.smooth_field_port(dk, 8.0, 0.7, N, *L), b.smooth_field(b.delta_k, 8.0), atol=1e-15)


def test_c_oracle_matches_numpy_restatement():
    subprocess.check_call(["make", "-s"], cwd=os.path.join(ROOT, "oracle"))
    lib = ctypes.CDLL(os.path.join(ROOT, "oracle", "_build", "libfb_oracle.so"))
    rng = np.random.default_rng(3)
    lam = np.concatenate([rng.uniform(0, 5, 20000), rng.uniform(5, 120, 20000), [0.0, 1e-12, 650.0]])
    u = rng.random(lam.size)
    out = np.zeros(lam.size, np.int32)
    lib.fb_oracle_poisson(lam.ctypes.data_as(ctypes.c_void_p), u.ctypes.data_as(ctypes.c_void_p),
                          ctypes.c_long(lam.size), out.ctypes.data_as(ctypes.c_void_p))
    ref = R.poisson_from_uniform(lam, u)
    assert np.array_equal(out.astype(np.int64), ref)
    # the certain-zero shortcut of the header (u < (1 - lam) - 1e-14 for lam < 0.5) never changes a count: uniforms
    # placed on and around its threshold and around p_0 = exp(-lam), sparse-tracer means (the NumPy restatement
    # has no shortcut)
    lam_s = np.concatenate([rng.uniform(0, 0.5, 20000), 10.0 ** rng.uniform(-12, -1, 20000), [0.5, 0.49999999999999994]])
    thr = (1.0 - lam_s) - 1e-14
    p0 = R._exp_neg(lam_s)
    cases = [thr, np.nextafter(thr, 0.0), np.nextafter(thr, 2.0), p0, np.nextafter(p0, 0.0), np.nextafter(p0, 2.0),
             1.0 - lam_s, np.minimum(p0 + lam_s * rng.random(lam_s.size), np.nextafter(1.0, 0.0)), rng.random(lam_s.size)]
    for us in cases:
        us = np.ascontiguousarray(np.clip(us, 0.0, np.nextafter(1.0, 0.0)))
        outs = np.zeros(lam_s.size, np.int32)
        lib.fb_oracle_poisson(lam_s.ctypes.data_as(ctypes.c_void_p), us.ctypes.data_as(ctypes.c_void_p),
                              ctypes.c_long(lam_s.size), outs.ctypes.data_as(ctypes.c_void_p))
        assert np.array_equal(outs.astype(np.int64), R.poisson_from_uniform(lam_s, us))
    # distributional sanity vs np.random.poisson (the reference's sampler, halos.py:116)
    lam1 = np.full(200000, 7.5)
    k = R.poisson_from_uniform(lam1, rng.random(lam1.size))
    assert abs(k.mean() - 7.5) < 0.03 and abs(k.var() - 7.5) < 0.1
    # large means (exp(-lam) underflows from ~745 on): same answer from C and NumPy, Poisson mean and variance
    for lam_big in (740.0, 746.0, 1e3, 2e3):
        lamb = np.full(4000, lam_big)
        ub = rng.random(lamb.size)
        outb = np.zeros(lamb.size, np.int32)
        lib.fb_oracle_poisson(lamb.ctypes.data_as(ctypes.c_void_p), ub.ctypes.data_as(ctypes.c_void_p),
                              ctypes.c_long(lamb.size), outb.ctypes.data_as(ctypes.c_void_p))
        assert abs(outb.mean() / lam_big - 1) < 4 * np.sqrt(1.0 / lam_big / lamb.size) + 1e-4
        assert abs(outb.var() / lam_big - 1) < 0.15
        assert np.array_equal(outb[:200].astype(np.int64), R.poisson_from_uniform(lamb[:200], ub[:200]))
    # exp(-lam) restatement is accurate to 1 ulp-ish
    lib.fb_oracle_exp_neg.restype = ctypes.c_double
    lib.fb_oracle_exp_neg.argtypes = [ctypes.c_double]
    for x in (0.0, 1e-3, 0.5, 3.0, 30.0, 200.0, 650.0):
        assert lib.fb_oracle_exp_neg(x) == float(R._exp_neg(np.array(x)))
        assert abs(lib.fb_oracle_exp_neg(x) / np.exp(-x) - 1) < 5e-16
    # digitize restatement in C == np.digitize on the reference's k array
    N, L = 16, (1e2, 2e2, 1e3)
    edges = R.pk_bin_edges(N, *L, nbins=20)
    counts = np.zeros(21, np.int64)
    lib.fb_oracle_digitize_counts(ctypes.c_int(N), ctypes.c_double(L[0]), ctypes.c_double(L[1]),
                                  ctypes.c_double(L[2]), edges.ctypes.data_as(ctypes.c_void_p), ctypes.c_int(20),
                                  counts.ctypes.data_as(ctypes.c_void_p))
    assert np.array_equal(counts, np.bincount(R.digitize_modes(N, *L, edges), minlength=21))


def test_philox_known_answer():
    """Philox4x32-10 known-answer vectors (Random123 kat_vectors): counter/key -> output."""
    out = R.philox4x32(np.zeros((1, 4), np.uint32), np.zeros(2, np.uint32))
    assert [hex(int(x)) for x in out[0]] == ['0x6627e8d5', '0xe169c58d', '0xbc57ac4c', '0x9b00dbd8']
    ctr = np.full((1, 4), 0xffffffff, np.uint32)
    out = R.philox4x32(ctr, np.full(2, 0xffffffff, np.uint32))
    assert [hex(int(x)) for x in out[0]] == ['0x408f276d', '0x41c83b0e', '0xa20bc7c6', '0x6d5451fd']
    # Hermitian white noise drawn directly: Re, Im of variance 1/2, H0(-k) = conj H0(k), 8 real self-conjugate modes
    N = 32
    re, im = R.philox_normals(42, np.arange(N ** 3, dtype=np.uint64).reshape(N, N, N), N)
    assert abs(re.mean()) < 0.01 and abs(re.var() - 0.5) < 0.01 and abs(im.var() - 0.5) < 0.01
    assert abs(np.corrcoef(re.ravel(), im.ravel())[0, 1]) < 0.02
    W = re + 1j * im
    assert np.array_equal(W, np.conj(np.roll(W[::-1, ::-1, ::-1], (1, 1, 1), (0, 1, 2))))
    assert np.count_nonzero(im == 0.0) == 8
    assert np.array_equal(R.hermitian_half_from_noise(re, im, np.ones((N // 2 + 1, N, N))), W[:N // 2 + 1])
    real_field = np.fft.ifftn(W)
    assert np.abs(real_field.imag).max() < 1e-15 and abs(real_field.real.var() * N ** 3 - 1) < 0.02


@pytest.mark.parametrize("name", ["sparse", "dense", "mixed"])
def test_halo_catalogue_port_matches_reference_golden(name):
    """halos.py:120-176: both restatements are bit-identical to the unmodified reference, with and without scatter."""
    g = load_golden("halo_catalogue")
    counts, L = g[name + "_counts"], g[name + "_L"]
    for f in (R.halo_catalogue_port, R.halo_catalogue_lean):
        cat = f(counts, *L)
        assert cat.dtype == np.float64 and cat.shape == (counts.sum(), 3)
        assert np.array_equal(cat, g[name + "_cat"])
        np.random.seed(int(g[name + "_scatter_seed"]))
        u = np.random.uniform(0., 1. - 1e-8, cat.size).reshape(cat.shape)       # halos.py:166
        assert np.array_equal(f(counts, *L, uniforms=u), g[name + "_cat_scatter"])
    assert R.halo_catalogue_port(np.zeros((4, 4, 4), int), 1., 1., 1.).shape == (0, 3)


def test_fg_and_noise_cube_ports_match_reference_golden():
    """foregrounds.py:152-174 and noise.py:55-75 restated; bit-identical to the unmodified reference."""
    g = load_golden("fg_noise_cube")
    assert np.array_equal(R.fg_construct_cube_port(g["amps"], g["alpha"], g["freqs"]), g["fg_cube_map"])
    assert np.array_equal(R.fg_construct_cube_port(g["amps"], -2.7, g["freqs"]), g["fg_cube_scalar"])
    sig = R.radiometer_rms_port(g["freqs"], g["ang_x"], 18., 2.5, 1., 64)
    np.random.seed(78)
    n = np.random.normal(0., 1., (16, 16, 16))                                   # noise.py:73
    assert np.array_equal(R.radiometer_noise_port(sig, n), g["noise"])
    c = R.philox_noise_cube(5, 16)
    assert abs(c.mean()) < 0.05 and abs(c.std() - 1) < 0.05


def test_mean_spectrum_filter_port_matches_reference_golden():
    """filters.py:35-55 restated; bit-identical to the unmodified reference."""
    g = load_golden("fg_noise_cube")
    assert np.array_equal(R.mean_spectrum_filter_port(g["data_cube"]), g["mean_filtered"])


def test_pca_filter_port_matches_reference_golden():
    """filters.py:93-183 restated; bit-identical to the unmodified reference (same LAPACK underneath)."""
    import warnings
    g = load_golden("fg_noise_cube")
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for nm in (2, 4):
            assert np.array_equal(R.pca_filter_port(g["pca_cube"], nm), g["pca_clean%d" % nm])
        c, U, a = R.pca_filter_port(g["pca_cube"], 3, return_filter=True)
        assert np.array_equal(c, g["pca_clean3"]) and np.array_equal(np.real(U), g["pca_U3"])
        assert np.array_equal(np.real(a), g["pca_amps3"])
        assert np.array_equal(np.real(R.pca_filter_port(g["pca_cube"], 3, fit_powerlaw=True)), g["pca_clean3_pl"])
        assert np.array_equal(R.pca_filter_port(g["pca_cube_fg"], 3), g["pca_fg_clean3"])


def test_slabwise_restatement_matches_pinned_ports():
    """realise_slabwise (the 512^3 / 1024^3 checker) == the restatements pinned to the reference above."""
    N, L = 32, (1e3, 1.5e3, 2e3)
    re, im = draw_noise(3, N)
    _, pkf = pk_function(0.8)
    noise = lambda a: (re[a].astype(np.float32), im[a].astype(np.float32))
    re32, im32 = re.astype(np.float32).astype(np.float64), im.astype(np.float32).astype(np.float64)
    kperp, kpar = R.kperp_kpar(N, *L)
    amp = R.sqrt_pk_half(pkf, N, *L) * np.nan_to_num(transfer_fn(kperp[:N // 2 + 1], kpar))
    half = R.hermitian_half_from_noise(re32, im32, amp)
    ref = R.irfft3_axis0(half)
    kc, pk, err, cnt = R.binned_power_spectrum_lean(half, N, *L, nbins=20)
    out = R.realise_slabwise(noise, pkf, N, *L, transfer_fn=transfer_fn, nbins=20, workers=3,
                             kinds=(None, "vel_x", "vel_y", "vel_z", "potential"))
    assert np.abs(out["fields"][None] - ref).max() <= 1e-13 * np.abs(ref).max()
    assert np.array_equal(out["count"][:20], cnt[:20])
    with np.errstate(all="ignore"):
        mean = out["sum1"] / out["count"]
    ok = ~np.isnan(pk)
    assert np.allclose(mean[1:20][ok], pk[ok], rtol=1e-12)
    full = R.expand_half_axis0(half)
    vel = R.velocity_k_port(full, N, *L, 1.0)                   # box.py:251-285
    for i, kd in enumerate(("vel_x", "vel_y", "vel_z")):
        want = np.fft.ifftn(vel[i]).real
        assert np.abs(out["fields"][kd] - want).max() <= 1e-12 * np.abs(want).max()
    want = np.fft.ifftn(R.potential_k_port(full, N, *L)).real   # box.py:347-352
    assert np.abs(out["fields"]["potential"] - want).max() <= 1e-12 * np.abs(want).max()
    part = R.realise_slabwise(noise, pkf, N, *L, transfer_fn=transfer_fn, x_planes=[0, 7, 31], workers=2)
    assert np.abs(part["fields"][None] - ref[[0, 7, 31]]).max() <= 1e-13 * np.abs(ref).max()


def test_two_point_restatements_are_self_consistent():
    """
    P(k_perp, k_par) and xi(r) (parity unpinned w.r.t. nbodykit): collapsing the 2-D table over one axis gives the
    pinned 1-D estimator's populations; xi(r) equals a direct pair average on a small grid.
    """
    N, L = 16, (1e2, 1.5e2, 2e2)
    rng = np.random.default_rng(2)
    f = rng.standard_normal((N, N, N))
    half = R.rfft3_axis0(f)
    kp = np.concatenate([[0.0], np.linspace(0.05, 1.2, 9)])
    kl = np.concatenate([[0.0], np.linspace(0.04, 0.6, 7)])
    cp, cl, mean, err, cnt = R.binned_power_spectrum_2d_lean(half, N, *L, kp, kl)
    assert mean.shape == (kp.size - 1, kl.size - 1) and cnt.shape == (kp.size + 1, kl.size + 1)
    assert cnt.sum() == N ** 3                                   # every mode lands in exactly one cell
    m = R.mode_numbers(N).astype(np.float64)
    kpar_full = np.abs(2 * np.pi * m / L[2])
    assert np.array_equal(cnt.sum(axis=0), np.bincount(np.digitize(kpar_full, kl), minlength=kl.size + 1) * N * N)
    # total power is conserved by the binning
    tot = np.nansum(np.where(cnt > 0, 1.0, 0.0))
    assert tot > 10
    # xi(r): Wiener-Khinchin == direct average of f(x) f(x + lag) for three lags
    edges = np.array([0.0, 1e-9, 7.0, 11.0, 16.0, 25.0])
    cent, xi, xerr, c = R.correlation_function_port(f, N, *L, edges)
    r = R.lag_separations(N, *L)
    direct = np.zeros((N, N, N))
    for (i, j, l) in [(0, 0, 0), (1, 0, 0), (0, 1, 1), (N - 1, 0, 2)]:
        direct[i, j, l] = np.mean(f * np.roll(f, (-i, -j, -l), axis=(0, 1, 2)))
    A = np.fft.fftn(f)
    full = np.fft.ifftn(A * np.conj(A)).real / N ** 3
    for (i, j, l) in [(0, 0, 0), (1, 0, 0), (0, 1, 1), (N - 1, 0, 2)]:
        assert abs(full[i, j, l] - direct[i, j, l]) < 1e-12
    assert abs(xi[0] - np.mean(f * f)) < 1e-12 and c[1] == 1      # the zero lag alone sits in the first bin
    sel = (r >= 7.0) & (r < 11.0)
    assert abs(xi[2] - full[sel].mean()) < 1e-13 and c[3] == sel.sum()


def test_multipoles_and_cross_power_known_answers():
    """
    The multipole and cross-power restatements cannot be pinned to nbodykit (absent), so they are tied to
    published known answers instead.  (1) Kaiser (1987) / Hamilton (1992): a spectrum d_s(k) = (1 + beta mu^2) d(k)
    with mu = k_z / |k| has P_0 / P = 1 + 2 beta/3 + beta^2/5, P_2 / P = 4 beta/3 + 4 beta^2/7, P_4 / P = 8 beta^2/35 --
    this fixes the (2l+1) normalisation, the Legendre polynomials and the line of sight (z) of `pk_multipoles`.
    (2) cross(a, c a) = c auto(a) with the pinned auto estimator, cross(a, b) = cross(b, a), and the cross power of a
    field with a copy shifted by one cell along x carries the factor cos(k_x dx).
    """
    N, L = 128, (5e2, 5e2, 5e2)
    rng = np.random.default_rng(5)
    f = rng.standard_normal((N, N, N))
    half = R.rfft3_axis0(f)
    half = half / np.maximum(np.abs(half), 1e-300) * np.sqrt(R.boxfactor(N, *L))      # |d_k|^2 / boxfactor = 1 exactly
    m = R.mode_numbers(N).astype(np.float64)
    h = N // 2 + 1
    kk = np.sqrt(m[:h, None, None] ** 2 + m[None, :, None] ** 2 + m[None, None, :] ** 2)
    with np.errstate(all="ignore"):
        mu = np.where(kk > 0, m[None, None, :] / kk, 0.0)
    beta = 0.6
    knyq = np.pi * N / L[0]
    edges = np.linspace(0.3 * knyq, 0.98 * knyq, 5)            # complete, well-populated shells inside the Nyquist cube
    cent, poles = R.pk_multipoles(half * (1.0 + beta * mu ** 2), N, *L, kbins=edges)
    want = {0: 1 + 2 * beta / 3 + beta ** 2 / 5, 2: 4 * beta / 3 + 4 * beta ** 2 / 7, 4: 8 * beta ** 2 / 35}
    # tolerance = the quadrature error of the discrete shells (falls as 1/N^2: 1.6e-4 / 5.5e-3 / 3.7e-2 at N = 64)
    for ell, tol in ((0, 2e-4), (2, 3e-3), (4, 1.5e-2)):
        assert np.all(np.abs(poles[ell] - want[ell]) < tol), (ell, poles[ell], want[ell])
    # an isotropic spectrum has no quadrupole / hexadecapole beyond the discreteness of the shells
    _, iso = R.pk_multipoles(half, N, *L, kbins=edges)
    assert np.all(np.abs(iso[0] - 1.0) < 1e-12) and np.all(np.abs(iso[2]) < 3e-3) and np.all(np.abs(iso[4]) < 1.5e-2)

    N = 32
    m = R.mode_numbers(N).astype(np.float64)
    h = N // 2 + 1
    f = rng.standard_normal((N, N, N))
    g = rng.standard_normal((N, N, N))
    ha, hb = R.rfft3_axis0(f), R.rfft3_axis0(g)
    kc, auto, aerr, acnt = R.binned_power_spectrum_lean(ha, N, *L, nbins=20)
    kc2, cr, cerr, ccnt = R.binned_power_spectrum_lean(ha, N, *L, nbins=20, half_b=2.5 * ha)
    ok = ~np.isnan(auto)
    assert np.array_equal(acnt, ccnt) and np.allclose(cr[ok], 2.5 * auto[ok], rtol=1e-13)
    ab = R.binned_power_spectrum_lean(ha, N, *L, nbins=20, half_b=hb)[1]
    ba = R.binned_power_spectrum_lean(hb, N, *L, nbins=20, half_b=ha)[1]
    assert np.allclose(ab[ok], ba[ok], rtol=0, atol=1e-12 * np.abs(auto[ok]).max())
    shifted = R.rfft3_axis0(np.roll(f, 1, axis=0))             # f(x - dx): spectrum times exp(-i k_x dx)
    power = (ha * np.conj(ha)).real * np.cos(2 * np.pi * m[:h] / N)[:, None, None]
    bins = R.pk_bin_edges(N, *L, 20, None)
    idx = R.digitize_half(N, *L, bins)
    w = np.broadcast_to(R.half_weights(N)[:, None, None], power.shape)
    s = np.bincount(idx.ravel(), weights=(w * power).ravel(), minlength=bins.size + 1)
    c = np.bincount(idx.ravel(), weights=w.ravel(), minlength=bins.size + 1)
    with np.errstate(all="ignore"):
        want_x = (s / c)[1:bins.size] / R.boxfactor(N, *L)
    got_x = R.binned_power_spectrum_lean(ha, N, *L, nbins=20, half_b=shifted)[1]
    assert np.allclose(got_x[ok], want_x[ok], rtol=1e-10, atol=1e-12 * np.abs(auto[ok]).max())
