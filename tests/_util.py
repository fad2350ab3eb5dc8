"""Shared helpers for the parity tests (oracle side + plan set-up)."""
import os

import numpy as np

from fastbox_b200 import cosmology as cos
from fastbox_b200 import kspace as ks
from oracle import restate as R

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
DEFAULT_COSMO = dict(Omega_c=0.25, Omega_b=0.05, h=0.7, n_s=0.95, sigma8=0.8,
                     transfer_function='eisenstein_hu')          # box.py:18-20
TOL = 1e-5                                                       # north_star tolerance


def rel_l2(a, b):
    cplx = np.iscomplexobj(a) or np.iscomplexobj(b)
    a = np.asarray(a, dtype=np.complex128 if cplx else np.float64)
    b = np.asarray(b, dtype=np.complex128 if cplx else np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def load_golden(name):
    g = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    return {k: g[k] for k in g.files}


def pk_function(redshift):
    c = cos.Cosmology(**DEFAULT_COSMO)
    a = 1. / (1. + redshift)
    return c, (lambda k: cos.nonlin_matter_power(c, k, a))


def draw_noise(seed, N):
    """The reference's own draw order (box.py:174-175): re first, then im, C order."""
    np.random.seed(seed)
    re = np.random.normal(0.0, 1.0, (N, N, N))
    im = np.random.normal(0.0, 1.0, (N, N, N))
    return re, im


def transfer_fn(k_perp, k_par):                                  # reference tests/test_box.py:88-90
    return (1. - np.exp(-0.5 * (k_par / 0.001) ** 2.)) * np.exp(-0.5 * (k_perp / 0.1) ** 2.)


def setup_plan(N, L, redshift, nbins=None, filt=None, device=0, exact_below=512):
    """Plan with sqrt(P) table (+ optional bins / filter) for box lengths L."""
    from fastbox_b200 import _lib
    Lx, Ly, Lz = L
    plan = _lib.Plan(N, Lx, Ly, Lz, device)
    _, pkf = pk_function(redshift)
    bf = R.boxfactor(N, Lx, Ly, Lz)
    mode, tab, l0, dl = ks.choose_sqrt_pk_table(pkf, N, Lx, Ly, Lz, bf, exact_below=exact_below)
    plan.set_sqrt_pk(tab, mode, l0, dl)
    edges = None
    if nbins is not None:
        kmin, kmax = R.kmin_kmax(N, Lx, Ly, Lz)
        edges = ks.pk_bin_edges(kmin, kmax, nbins)
        plan.set_pk_bins(ks.bin_thresholds(edges))
    if filt is not None:
        ft = ks.filter_tables(filt, N, Lx, Ly, Lz)
        plan.set_filter(ft.tperp, ft.tpar, ft.tdense)
    return plan, edges


def assert_pk_close(got, ref, tol=TOL, floor_rel=0.0):
    """
    Per-bin comparison incl. identical NaN (empty-bin) pattern.  ``floor_rel``: absolute floor as a fraction of the
    largest bin -- a Gaussian k_perp filter drives the last bins of a big box 150 orders of magnitude below the
    first ones, far outside the range of float32 (|H|^2 underflows ~56 orders below the peak).
    """
    kc, pk, err = got
    kc_r, pk_r, err_r = ref
    assert np.allclose(kc, kc_r, rtol=1e-14, atol=0)
    assert np.array_equal(np.isnan(pk), np.isnan(pk_r))
    m = ~np.isnan(pk_r)
    floor = floor_rel * np.max(np.abs(pk_r[m])) if np.any(m) else 0.0
    assert np.all(np.abs(pk[m] - pk_r[m]) <= tol * np.abs(pk_r[m]) + floor)
    # error bar: the reference's np.std of a 2-element Hermitian pair is rounding noise, so
    # compare with an absolute floor tied to the bin's power
    assert np.all(np.abs(err[m] - err_r[m]) <= 10 * tol * np.abs(err_r[m]) + 1e-9 * np.abs(pk_r[m]) + floor)


def deviation_report(got, ref, label=""):
    """(rel-L2, number of cells off by more than 1e-5 rms, largest deviation / rms); printed for the test log."""
    got = np.asarray(got, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    rms = float(np.sqrt(np.mean(ref ** 2)))
    diff = np.abs(got - ref)
    nbad = int(np.count_nonzero(diff > 1e-5 * rms))
    rel = rel_l2(got, ref)
    print("%s rel-L2 %.3e, cells off by > 1e-5 rms: %d of %d, max |dev|/rms %.3e" % (label, rel, nbad, ref.size,
                                                                                     diff.max() / rms))
    return rel, nbad, float(diff.max() / rms)
