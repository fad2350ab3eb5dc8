"""
GPU parity of the fused pipelines against (a) golden vectors produced by the
UNMODIFIED reference (tests/golden, oracle/make_golden.py) and (b) the NumPy
restatement oracle/restate.py at sizes the reference fixtures do not cover.
All calls go through the C ABI (fastbox_b200._lib.Plan).
"""
import numpy as np
import pytest

from fastbox_b200 import _lib
from fastbox_b200 import kspace as ks
from oracle import restate as R

from _util import (TOL, assert_pk_close, deviation_report, draw_noise, load_golden, pk_function, rel_l2, setup_plan,
                   transfer_fn)

pytestmark = pytest.mark.gpu

F = _lib


def _case_L(g):
    return tuple(float(x) for x in g["L"])


@pytest.mark.parametrize("name", ["n16_cubic", "n16_cuboid", "n32_gpc"])
@pytest.mark.parametrize("nbins", [20, 50])
def test_realise_matches_reference_golden(gpu, name, nbins):
    g = load_golden(name)
    N, L = int(g["N"]), _case_L(g)
    re, im = draw_noise(int(g["seed"]), N)
    plan, edges = setup_plan(N, L, float(g["redshift"]), nbins=nbins)
    field = np.empty((N, N, N), np.float32)
    spec = np.empty((N // 2 + 1, N, N), np.complex64)
    res, sums = plan.realise(re.astype(np.float32), im.astype(np.float32), flags=F.F_SQRTPK, field_out=field,
                             spec_out=spec, want_pk=True)
    tol = TOL if L[0] == L[1] == L[2] else 5 * TOL     # cuboid: interpolated sqrt(P) table
    assert rel_l2(field, g["delta_x"]) < tol
    assert rel_l2(spec, g["delta_k_half"]) < tol
    # bin populations are integers: bit exact against np.digitize on the reference's k array
    assert np.array_equal(res["count"].astype(np.int64), g["pk%d_counts" % nbins].astype(np.int64))
    got = ks.moments_to_spectrum(edges, res["count"], res["sum1"], res["sum2"])
    assert_pk_close(got, (g["pk%d_k" % nbins], g["pk%d_p" % nbins], g["pk%d_e" % nbins]), tol=2 * tol)
    # Parseval by-products (box.py:944-946)
    assert abs(sums[1] * N ** 3 / g["parseval"][0] - 1) < 1e-4
    plan.close()


@pytest.mark.parametrize("name", ["n16_cubic", "n32_gpc"])
def test_velocity_transfer_lognormal_golden(gpu, name):
    g = load_golden(name)
    N, L = int(g["N"]), _case_L(g)
    re, im = draw_noise(int(g["seed"]), N)
    c, _ = pk_function(float(g["redshift"]))
    plan, _ = setup_plan(N, L, float(g["redshift"]), filt=transfer_fn)
    field = np.empty((N, N, N), np.float32)
    spec = plan.alloc((N // 2 + 1) * N * N * 8)
    plan.realise(re.astype(np.float32), im.astype(np.float32), field_out=field, spec_out=spec)
    # velocity (box.py:254-285): fac = 100 h E f a
    from fastbox_b200 import cosmology as cos
    a = 1. / (1. + float(g["redshift"]))
    fac = 100. * c['h'] * cos.h_over_h0(c, a) * cos.growth_rate(c, a) * a
    for kind, key in ((F.KIND_VEL_Z, "vel_z"), (F.KIND_VEL_X, "vel_x")):
        v = np.empty((N, N, N), np.float32)
        plan.spectrum_to_field(spec, v, kind=kind, scale=fac)
        assert rel_l2(v, g[key]) < TOL, key
    # transfer function (box.py:374-380)
    t = np.empty((N, N, N), np.float32)
    plan.spectrum_to_field(spec, t, flags=F.F_FILTER)
    assert rel_l2(t, g["transfer"]) < TOL
    # bias + log-normal (example_endtoend.py:33-36; box.py:457-459)
    e = np.empty((N, N, N), np.float32)
    s1, _ = plan.spectrum_to_field(spec, e, flags=F.F_EXP, scale=float(g["bias_HI"]))
    ln = e.astype(np.float64) / (s1 / N ** 3) - 1.0
    assert rel_l2(ln, g["lognormal"]) < TOL
    plan.close()


@pytest.mark.parametrize("N", [64, 128])
def test_realise_filter_pk_vs_oracle(gpu, N):
    """Headline pipeline (realise + filter + P(k)) at sizes beyond the fixtures."""
    L = (1e3, 1e3, 1e3)
    re, im = draw_noise(14, N)
    _, pkf = pk_function(0.8)
    plan, edges = setup_plan(N, L, 0.8, nbins=50, filt=transfer_fn)
    field = np.empty((N, N, N), np.float32)
    res, _ = plan.realise(re.astype(np.float32), im.astype(np.float32), flags=F.F_SQRTPK | F.F_FILTER,
                          field_out=field, want_pk=True, poles=True)
    amp = R.sqrt_pk_half(pkf, N, *L)
    kperp, kpar = R.kperp_kpar(N, *L)
    amp = amp * np.nan_to_num(transfer_fn(kperp[:N // 2 + 1], kpar))
    half = R.hermitian_half_from_noise(re, im, amp)
    ref = R.irfft3_axis0(half)
    assert rel_l2(field, ref) < TOL
    kc, pk, err, cnt = R.binned_power_spectrum_lean(half, N, *L, nbins=50)
    assert np.array_equal(res["count"][:50].astype(np.int64), cnt[:50])
    assert_pk_close(ks.moments_to_spectrum(edges, res["count"], res["sum1"], res["sum2"]), (kc, pk, err))
    # multipoles (parity unpinned w.r.t. nbodykit; restated in oracle/restate.py)
    _, poles = R.pk_multipoles(half, N, *L, nbins=50)
    cntf = res["count"][:50].astype(np.float64)
    with np.errstate(all="ignore"):
        p2 = 5.0 * res["sum_l2"][:50] / cntf
        p4 = 9.0 * res["sum_l4"][:50] / cntf
    m = ~np.isnan(poles[2])
    scale = np.abs(pk[m]) + 1e-30
    assert np.all(np.abs(p2[1:][m] - poles[2][m]) <= 20 * TOL * scale)
    assert np.all(np.abs(p4[1:][m] - poles[4][m]) <= 20 * TOL * scale)
    plan.close()


def test_interpolated_sqrtp_table_and_custom_bins(gpu):
    """Cubic box forced onto the log2(s) cubic-interpolated table (the N >= 512 default) + non-log bins."""
    N, L = 128, (1e3, 1e3, 1e3)
    re, im = draw_noise(3, N)
    _, pkf = pk_function(0.8)
    plan, _ = setup_plan(N, L, 0.8, exact_below=0)
    kbins = np.concatenate([[0.0], np.linspace(0.01, 0.9, 24)])        # arbitrary edges (box.py:745-746)
    plan.set_pk_bins(ks.bin_thresholds(kbins))
    field = np.empty((N, N, N), np.float32)
    res, _ = plan.realise(re.astype(np.float32), im.astype(np.float32), field_out=field, want_pk=True)
    ref, half = R.realise_density_lean(re, im, pkf, N, *L)
    assert rel_l2(field, ref) < TOL
    kc, pk, err, cnt = R.binned_power_spectrum_lean(half, N, *L, kbins=kbins)
    assert np.array_equal(res["count"][:kbins.size].astype(np.int64), cnt[:kbins.size])
    assert_pk_close(ks.moments_to_spectrum(kbins, res["count"], res["sum1"], res["sum2"]), (kc, pk, err))
    plan.close()


@pytest.mark.parametrize("N", [32, 64])
def test_forward_pk_and_cross(gpu, N):
    L = (5e2, 5e2, 5e2)
    rng = np.random.default_rng(3)
    fa = rng.standard_normal((N, N, N)).astype(np.float32)
    fb_ = (0.5 * fa + rng.standard_normal((N, N, N))).astype(np.float32)
    plan, edges = setup_plan(N, L, 0.0, nbins=20)
    res = plan.field_to_spectrum(fa, want_pk=True)
    ha = R.rfft3_axis0(fa.astype(np.float64))
    hb = R.rfft3_axis0(fb_.astype(np.float64))
    kc, pk, err, cnt = R.binned_power_spectrum_lean(ha, N, *L, nbins=20)
    assert np.array_equal(res["count"][:20].astype(np.int64), cnt[:20])
    assert_pk_close(ks.moments_to_spectrum(edges, res["count"], res["sum1"], res["sum2"]), (kc, pk, err))
    # the same through the reference-style port on the full cube (box.py:741-768)
    full = np.fft.fftn(fa.astype(np.float64))
    kc2, pk2, err2 = R.binned_power_spectrum_port(full, N, *L, nbins=20)
    assert_pk_close(ks.moments_to_spectrum(edges, res["count"], res["sum1"], res["sum2"]), (kc2, pk2, err2))
    # cross spectrum: store spectrum of b, then bin Re[a conj b]
    spec_b = plan.alloc((N // 2 + 1) * N * N * 8)
    plan.field_to_spectrum(fb_, spec_out=spec_b)
    resx = plan.field_to_spectrum(fa, cross=spec_b, want_pk=True)
    kcx, pkx, errx, cntx = R.binned_power_spectrum_lean(ha, N, *L, nbins=20, half_b=hb)
    got = ks.moments_to_spectrum(edges, resx["count"], resx["sum1"], resx["sum2"])
    m = ~np.isnan(pkx)
    assert np.all(np.abs(got[1][m] - pkx[m]) <= 10 * TOL * np.abs(pk[m]))
    # P(k) from a stored spectrum, half and full-cube modes
    spec_a = np.ascontiguousarray(ha.astype(np.complex64))
    r1 = plan.pk_from_spectrum(spec_a)
    assert np.array_equal(r1["count"], res["count"])
    r2 = plan.pk_from_spectrum(np.ascontiguousarray(full.astype(np.complex64)), full_cube=True)
    assert np.array_equal(r2["count"], res["count"])
    assert_pk_close(ks.moments_to_spectrum(edges, r2["count"], r2["sum1"], r2["sum2"]), (kc2, pk2, err2))
    plan.close()


def test_philox_noise_matches_oracle(gpu):
    N, L = 32, (1e3, 1e3, 1e3)
    _, pkf = pk_function(0.0)
    plan, _ = setup_plan(N, L, 0.0)
    field = np.empty((N, N, N), np.float32)
    plan.realise(None, None, seed=0x1234567890ABCDEF, field_out=field)
    idx = np.arange(N ** 3, dtype=np.uint64).reshape(N, N, N)
    re, im = R.philox_normals(0x1234567890ABCDEF, idx, N)
    ref, _ = R.realise_density_lean(re, im, pkf, N, *L)
    assert rel_l2(field, ref) < TOL
    plan.close()


def test_cube_to_field_general_complex(gpu):
    """apply_transfer_fn on an arbitrary (non-Hermitian) cube with a k_par-odd filter (box.py:378-380)."""
    N, L = 32, (2e2, 2e2, 2e2)
    rng = np.random.default_rng(5)
    cube = (rng.standard_normal((N, N, N)) + 1j * rng.standard_normal((N, N, N)))
    fn = lambda kp, kl: np.exp(-0.5 * (kp / 0.3) ** 2) * (1.0 + 0.5 * np.tanh(kl / 0.2))
    plan = _lib.Plan(N, *L)
    ft = ks.filter_tables(fn, N, *L)
    assert not ft.even
    plan.set_filter(ft.tperp, ft.tpar, ft.tdense)
    ref = R.apply_transfer_fn_port(cube, fn, N, *L)
    c64 = np.ascontiguousarray(cube.astype(np.complex64))
    out_r = np.empty((N, N, N), np.float32)
    out_i = np.empty((N, N, N), np.float32)
    plan.cube_to_field(c64, out_r, flags=F.F_FILTER, part=0)
    plan.cube_to_field(c64, out_i, flags=F.F_FILTER, part=1)
    assert rel_l2(out_r, ref.real) < TOL
    assert rel_l2(out_i, ref.imag) < TOL
    # dense-table route gives the same answer
    ftd = ks.filter_tables(fn, N, *L, force_dense=True)
    plan.set_filter(None, None, ftd.tdense)
    plan.cube_to_field(c64, out_r, flags=F.F_FILTER, part=0)
    assert rel_l2(out_r, ref.real) < TOL
    plan.close()


@pytest.mark.parametrize("name", ["n16_cubic", "n32_gpc"])
def test_rsd_remap_golden(gpu, name):
    g = load_golden(name)
    N, L = int(g["N"]), _case_L(g)
    plan = _lib.Plan(N, *L)
    out = np.empty((N, N, N), np.float32)
    d32 = g["lognormal"].astype(np.float32)
    v32 = g["vel_z"].astype(np.float32)
    plan.rsd_remap(d32, v32, None, g["z_grid"], float(g["Hz"]), out)
    # oracle on the same float32-rounded inputs (the reference result for float64 inputs is the golden)
    ref32 = R.redshift_space_density(d32.astype(np.float64), v32.astype(np.float64), g["z_grid"], float(g["Hz"]))
    assert rel_l2(out, ref32) < TOL
    # ... and against the unmodified reference's float64 output itself: rounding the inputs to float32 moves no
    # interpolation node in these boxes (the float64 oracle on the rounded inputs is within 3e-8 of the golden)
    assert rel_l2(ref32, g["rsd0"]) < 1e-7
    rel, nbad, worst = deviation_report(out, g["rsd0"], "rsd_remap %s vs reference" % name)
    assert rel < TOL
    np.random.seed(int(g["seed"]) + 100)
    vnl = (120. * np.random.normal(0., 1., (N, N, N))).astype(np.float32)
    plan.rsd_remap(d32, v32, vnl, g["z_grid"], float(g["Hz"]), out)
    ref32 = R.redshift_space_density(d32.astype(np.float64), v32.astype(np.float64), g["z_grid"], float(g["Hz"]),
                                     vnl.astype(np.float64))
    assert rel_l2(out, ref32) < TOL
    plan.close()


@pytest.mark.parametrize("name", ["n16_cubic", "n16_cuboid", "n32_gpc"])
def test_rsd_remap_nearest_golden(gpu, name):
    """method='nearest' (box.py:403-405, 433-437) picks input values: every output equals the unmodified
    reference's float64 result rounded to float32, except where rounding the velocities to float32 (or the
    float32 in-cell fraction) flips a near-tie between the two neighbouring samples -- counted below."""
    g, gm = load_golden(name), load_golden("rsd_methods")
    N, L = int(g["N"]), _case_L(g)
    plan = _lib.Plan(N, *L)
    out = np.empty((N, N, N), np.float32)
    d32 = g["lognormal"].astype(np.float32)
    v32 = g["vel_z"].astype(np.float32)
    plan.rsd_remap(d32, v32, None, g["z_grid"], float(g["Hz"]), out, method="nearest")
    ref32 = R.redshift_space_density(d32.astype(np.float64), v32.astype(np.float64), g["z_grid"], float(g["Hz"]),
                                     method="nearest")
    nflip = int(np.count_nonzero(out != ref32.astype(np.float32)))
    nflip_ref = int(np.count_nonzero(out != gm[name + "_nearest0"].astype(np.float32)))
    print("rsd nearest %s: %d / %d voxels differ from the oracle, %d from the reference" % (name, nflip, N ** 3, nflip_ref))
    assert nflip <= 2 and nflip_ref <= 2
    np.random.seed(int(g["seed"]) + 100)
    vnl = (120. * np.random.normal(0., 1., (N, N, N))).astype(np.float32)
    plan.rsd_remap(d32, v32, vnl, g["z_grid"], float(g["Hz"]), out, method="nearest")
    ref32 = R.redshift_space_density(d32.astype(np.float64), v32.astype(np.float64), g["z_grid"], float(g["Hz"]),
                                     vnl.astype(np.float64), method="nearest")
    assert int(np.count_nonzero(out != ref32.astype(np.float32))) <= 2
    with pytest.raises(_lib.FastBoxError):
        plan.rsd_remap(d32, v32, None, g["z_grid"], float(g["Hz"]), out, method="cubic")
    plan.close()


@pytest.mark.parametrize("lognormal", [False, True])
def test_halo_counts_bit_exact(gpu, lognormal):
    g = load_golden("n32_gpc")
    N, L = int(g["N"]), _case_L(g)
    plan = _lib.Plan(N, *L)
    d32 = g["delta_x"].astype(np.float32)
    rng = np.random.RandomState(7)
    u = rng.uniform(0., 1., (N, N, N))
    nbar = np.float32(1e-3) if lognormal else np.linspace(1e-3, 2e-3, N).astype(np.float32)
    bias = np.float32(1.2)
    d64 = d32.astype(np.float64)
    mean_exp = float(np.mean(np.exp(np.float64(bias) * d64))) if lognormal else 0.0
    lam = R.halo_mean_count(d64, np.float64(nbar) if lognormal else nbar.astype(np.float64), np.float64(bias), *L,
                            lognormal_tf=lognormal)
    ref = R.poisson_from_uniform(lam, u)
    counts = np.empty((N, N, N), np.int32)
    mean_out = np.empty((N, N, N), np.float32)
    plan.halo_counts(d32, np.atleast_1d(nbar), 0 if lognormal else 1, np.atleast_1d(bias), 0, lognormal, mean_exp, u,
                     counts, mean_out)
    assert rel_l2(mean_out, lam) < 1e-6
    assert np.array_equal(counts.astype(np.int64), ref)            # integer output: bit exact
    assert counts.sum() > 0
    # against the reference's own mean-count field (golden, float64 density)
    key = "halo_mean_ln" if lognormal else "halo_mean_lin"
    if lognormal:
        assert rel_l2(mean_out, g[key]) < TOL
    plan.close()


@pytest.mark.parametrize("name", ["n16_cubic", "n32_gpc"])
def test_beam_convolve_golden(gpu, name):
    """BeamModel.convolve_fft (beams.py:81-87) against the unmodified reference's output."""
    from oracle.make_golden import beam_cube
    g = load_golden(name)
    N, L = int(g["N"]), _case_L(g)
    plan = _lib.Plan(N, *L)
    out = np.empty((N, N, N), np.float32)
    plan.beam_convolve(beam_cube(N).astype(np.float32), g["rsd0"].astype(np.float32), out)
    assert rel_l2(out, g["beam_conv"]) < TOL
    plan.close()


@pytest.mark.parametrize("N", [8, 64, 128])
def test_beam_convolve_vs_oracle(gpu, N):
    rng = np.random.default_rng(N)
    field = rng.standard_normal((N, N, N)).astype(np.float32)
    x = np.arange(N) - N / 2. + 0.5
    s = 1.0 + 3.0 * np.arange(N) / N
    beam = (np.exp(-0.5 * (x[:, None, None] ** 2 + (x[None, :, None] - 0.7) ** 2) / s[None, None, :] ** 2)
            * (1 + 0.1 * np.cos(x[:, None, None]))).astype(np.float32)      # not symmetric: checks the crop offset
    plan = _lib.Plan(N, 1e3, 1e3, 1e3)
    out = np.empty((N, N, N), np.float32)
    plan.beam_convolve(beam, field, out)
    ref = R.convolve_fft(beam.astype(np.float64), field.astype(np.float64))
    assert rel_l2(out, ref) < TOL
    # and through scipy itself, the call the reference makes
    import scipy.signal
    ref2 = scipy.signal.fftconvolve(beam.astype(np.float64), field.astype(np.float64), mode='same', axes=[0, 1]) \
        / np.sum(beam.astype(np.float64).reshape(-1, N), axis=0)[None, None, :]
    assert rel_l2(out, ref2) < TOL
    plan.close()


def test_beam_convolve_large_box_and_cached_spectrum(gpu):
    """
    N = 512 (2N = 1024-point x transforms, 8-channel tiles): single channels against scipy; the cached beam
    spectrum (fb_beam_set once, beam = NULL afterwards) gives bit-identical results, and a second beam replaces it.
    """
    import scipy.signal
    N = 512
    rng = np.random.default_rng(5)
    field = rng.standard_normal((N, N, N), dtype=np.float32)
    x = np.arange(N) - N / 2. + 0.5
    s = 2.0 + 6.0 * np.arange(N) / N
    beam = (np.exp(-0.5 * (x[:, None, None] ** 2 + (x[None, :, None] - 0.7) ** 2) / s[None, None, :] ** 2)
            * (1 + 0.1 * np.cos(x[:, None, None]))).astype(np.float32)
    plan = _lib.Plan(N, 1e3, 1e3, 1e3)
    out = np.empty((N, N, N), np.float32)
    plan.beam_convolve(beam, field, out)
    for z in (0, 201, N - 1):
        b2, f2 = beam[:, :, z].astype(np.float64), field[:, :, z].astype(np.float64)
        ref = scipy.signal.fftconvolve(b2, f2, mode='same') / b2.sum()         # beams.py:81-87, one channel
        assert rel_l2(out[:, :, z], ref) < TOL
    out2 = np.empty((N, N, N), np.float32)
    plan.beam_convolve(None, field, out2)                                      # cached spectrum
    assert np.array_equal(out, out2)
    beam_b = np.ascontiguousarray(beam[::-1] * 3.0)
    plan.beam_set(beam_b)
    plan.beam_convolve(None, field, out2)
    z = 77
    b2, f2 = beam_b[:, :, z].astype(np.float64), field[:, :, z].astype(np.float64)
    assert rel_l2(out2[:, :, z], scipy.signal.fftconvolve(b2, f2, mode='same') / b2.sum()) < TOL
    plan.close()


def test_beam_convolve_needs_a_beam(gpu):
    plan = _lib.Plan(16, 1e2, 1e2, 1e2)
    f = np.zeros((16, 16, 16), np.float32)
    with pytest.raises(_lib.FastBoxError):
        plan.beam_convolve(None, f, np.empty_like(f))
    plan.close()


@pytest.mark.parametrize("N,vscale", [(64, 5.0), (128, 60.0), (128, 2000.0), (256, 300.0)])
def test_rsd_remap_large_displacements(gpu, N, vscale):
    """Windowed bracket search == sort-based restatement, incl. periodic wrap-around and shell crossing."""
    rng = np.random.default_rng(N)
    L = 400.0
    z = np.linspace(-0.5 * L, 0.5 * L, N)
    Hz = 100.0
    d = rng.standard_normal((N, N, N)).astype(np.float32)
    xx = np.arange(N)
    coherent = np.sin(2 * np.pi * xx / N * 3)[None, None, :] * (1 + 0.3 * rng.standard_normal((N, N, 1)))
    v = (vscale * Hz * L / N / 10.0 * (coherent + 0.5 * rng.standard_normal((N, N, N)))).astype(np.float32)
    plan = _lib.Plan(N, L, L, L)
    out = np.empty((N, N, N), np.float32)
    plan.rsd_remap(d, v, None, z, Hz, out)
    sub = slice(0, 8)                       # the Python restatement loops over lines: check a slab
    ref = R.redshift_space_density(d[sub].astype(np.float64), v[sub].astype(np.float64), z, Hz)
    assert rel_l2(out[sub], ref) < TOL
    plan.close()
