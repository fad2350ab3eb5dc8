"""
Host-side k-space tables consumed by the CUDA library (NumPy, float64).

Everything here is O(N^2) or smaller; the N^3 arrays ``Kx, Ky, Kz, k`` that the
reference allocates eagerly (``fastbox/box.py:110-127``) are never built on the
hot path.
"""
import numpy as np

TWO_PI = 2. * np.pi


def mode_numbers(N):
    """Signed integer FFT mode numbers, as box.py:119."""
    return (N * np.fft.fftfreq(N, 1.)).astype("i")


def axis_sq(N, L):
    """(K/L)**2. per axis index, float64, the per-axis term of box.py:125-127."""
    m = mode_numbers(N).astype(np.float64)
    return (m / L) ** 2.


def k_of_s(s):
    return TWO_PI * np.sqrt(s)


def bin_thresholds(edges):
    """
    For each P(k) bin edge e return the smallest float64 ``s >= 0`` such that
    ``2*pi*sqrt(s) >= e`` *in float64 arithmetic*.  Because s -> 2 pi sqrt(s) is
    monotone under IEEE rounding, ``np.digitize(2*pi*sqrt(s), edges)`` equals
    ``#{j : thr[j] <= s}`` exactly, so the device needs no sqrt to reproduce the
    reference's bin index (box.py:758) bit for bit.
    """
    edges = np.atleast_1d(np.asarray(edges, dtype=np.float64))
    lo = np.zeros(edges.shape, dtype=np.uint64)                       # bits of +0.0
    hi = np.full(edges.shape, np.float64(1e300).view(np.uint64), dtype=np.uint64)
    # invariant: g(lo) < e (or lo == 0), g(hi) >= e
    for _ in range(64):
        mid = lo + (hi - lo) // np.uint64(2)
        ok = k_of_s(mid.view(np.float64)) >= edges
        hi = np.where(ok, mid, hi)
        lo = np.where(ok, lo, mid)
    thr = hi.view(np.float64).copy()
    thr[k_of_s(np.zeros_like(edges)) >= edges] = 0.0                  # e <= 0
    return thr


def pk_bin_edges(kmin, kmax, nbins=20, kbins=None):
    """box.py:745-749 (``nbins`` is the number of EDGES)."""
    if kbins is not None:
        return np.asarray(kbins, dtype=np.float64)
    return np.logspace(np.log10(kmin), np.log10(kmax), nbins)


def bin_centres(bins):
    """box.py:750-751."""
    full = [0.0] + list(bins)
    return np.array([0.5 * (full[j + 1] + full[j]) for j in range(len(bins))])


def moments_to_spectrum(bins, count, sum1, sum2):
    """
    (centres, mean, stddev/sqrt(n)) exactly as box.py:761-768 forms them from the
    per-bin moments; empty bins give NaN like the reference (mean of empty slice).
    """
    nb = len(bins)
    cnt = np.asarray(count[:nb], dtype=np.float64)
    with np.errstate(all="ignore"):
        mean = np.asarray(sum1[:nb]) / cnt
        var = np.maximum(np.asarray(sum2[:nb]) / cnt - mean * mean, 0.0)
        err = np.sqrt(var) / np.sqrt(cnt)
    cent = bin_centres(bins)
    return cent[1:], mean[1:], err[1:]


def sqrt_pk_int_lut(pk_of_k, N, L, boxfactor):
    """
    Cubic box: every |k|^2 is (2 pi/L)^2 n with integer n = i^2+j^2+l^2 <= 3 (N/2)^2,
    so sqrt(P(k) boxfactor) (box.py:161-176) is an exact LUT indexed by n.
    """
    n = np.arange(3 * (N // 2) ** 2 + 1, dtype=np.float64)
    k = TWO_PI * np.sqrt(n) / L
    pk = np.nan_to_num(np.asarray(pk_of_k(k), dtype=np.float64))       # box.py:167
    return np.sqrt(pk * boxfactor).astype(np.float32)


def sqrt_pk_log_table(pk_of_k, N, Lx, Ly, Lz, boxfactor, npts=1 << 16):
    """Cuboid box: table uniform in log2(s), s = sum (m/L)^2 (k = 2 pi sqrt(s))."""
    smin = min(1.0 / Lx ** 2, 1.0 / Ly ** 2, 1.0 / Lz ** 2)
    smax = (N / 2.0) ** 2 * (1.0 / Lx ** 2 + 1.0 / Ly ** 2 + 1.0 / Lz ** 2)
    l0 = np.log2(smin) - 1e-3
    l1 = np.log2(smax) + 1e-3
    dl = (l1 - l0) / (npts - 1)
    s = 2.0 ** (l0 + dl * np.arange(npts))
    pk = np.nan_to_num(np.asarray(pk_of_k(TWO_PI * np.sqrt(s)), dtype=np.float64))
    return np.sqrt(pk * boxfactor).astype(np.float32), l0, dl


class FilterTables(object):
    """Transfer function T(k_perp, k_par) (box.py:374-378) sampled for the device."""

    def __init__(self, tperp=None, tpar=None, tdense=None, even=True):
        self.tperp, self.tpar, self.tdense, self.even = tperp, tpar, tdense, even


def filter_tables(transfer_fn, N, Lx, Ly, Lz, nprobe=4096, rtol=1e-10, force_dense=False):
    """
    Evaluate ``transfer_fn(k_perp, k_par)`` on the half grid.  If it factorises as
    A(k_perp) * B(k_par) (checked on random probe points) only N^2/2 + N values are
    computed; otherwise a dense (N/2+1, N, N) float32 table is built plane by plane.
    NaN -> 0 as box.py:379 does for the product.
    """
    m = mode_numbers(N).astype(np.float64)
    h = N // 2 + 1
    kperp = TWO_PI * np.sqrt((m[:h, None] / Lx) ** 2. + (m[None, :] / Ly) ** 2.)     # (h, N)
    kpar = TWO_PI * m / Lz                                                           # (N,)
    with np.errstate(all="ignore"):
        if not force_dense:
            # reference point with a non-zero value
            rng = np.random.RandomState(12345)
            ia, ib, ic = rng.randint(0, h, nprobe), rng.randint(0, N, nprobe), rng.randint(0, N, nprobe)
            probe = np.nan_to_num(np.asarray(transfer_fn(kperp[ia, ib], kpar[ic]), dtype=np.float64))
            j = int(np.argmax(np.abs(probe)))
            t00 = probe[j]
            if t00 != 0.0:
                a = np.nan_to_num(np.asarray(transfer_fn(kperp, np.full_like(kperp, kpar[ic[j]])), dtype=np.float64))
                b = np.nan_to_num(np.asarray(transfer_fn(np.full_like(kpar, kperp[ia[j], ib[j]]), kpar),
                                             dtype=np.float64))
                pred = a[ia, ib] * b[ic] / t00
                scale = np.max(np.abs(probe))
                if np.all(np.abs(pred - probe) <= rtol * scale):
                    tpar = b / t00
                    even = bool(np.all(np.abs(tpar - tpar[(-np.arange(N)) % N]) <= 1e-14 * np.max(np.abs(tpar))))
                    return FilterTables(tperp=a.astype(np.float32), tpar=tpar.astype(np.float32), even=even)
            elif np.all(probe == 0.0):
                # identically zero on the probes: treat as dense to stay exact
                pass
        dense = np.empty((h, N, N), dtype=np.float32)
        even = True
        for i in range(h):
            plane = np.nan_to_num(np.asarray(transfer_fn(kperp[i][:, None] + 0 * kpar[None, :],
                                                         kpar[None, :] + 0 * kperp[i][:, None]), dtype=np.float64))
            dense[i] = plane
            if even:
                mir = plane[:, (-np.arange(N)) % N]
                even = bool(np.all(np.abs(plane - mir) <= 1e-14 * max(np.max(np.abs(plane)), 1e-300)))
        return FilterTables(tdense=dense, even=even)
