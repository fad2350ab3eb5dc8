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


def sqrt_pk_log_table(pk_of_k, N, Lx, Ly, Lz, boxfactor, npts=4096):
    """
    Table of sqrt(P(k) boxfactor) uniform in log2(s), s = sum (m/L)^2 (k = 2 pi sqrt(s)), read
    on the device with Catmull-Rom cubic interpolation.  Two guard nodes on each side.
    Returns (table float32, log2s0, dlog2s).
    """
    s, l0, dl = log_table_nodes(N, Lx, Ly, Lz, npts)
    pk = np.nan_to_num(np.asarray(pk_of_k(TWO_PI * np.sqrt(s)), dtype=np.float64))
    return np.sqrt(pk * boxfactor).astype(np.float32), l0, dl


def log_table_nodes(N, Lx, Ly, Lz, npts=4096):
    """Nodes s_i = 2^(l0 + i dl) covering every |k|^2/(2 pi)^2 of the grid, 2 guard nodes per side."""
    smin = min(1.0 / Lx ** 2, 1.0 / Ly ** 2, 1.0 / Lz ** 2)
    smax = (N / 2.0) ** 2 * (1.0 / Lx ** 2 + 1.0 / Ly ** 2 + 1.0 / Lz ** 2)
    dl = (np.log2(smax) - np.log2(smin)) / (npts - 5)
    l0 = np.log2(smin) - 2.0 * dl
    return 2.0 ** (l0 + dl * np.arange(npts)), l0, dl


def eval_log_table(tab, l0, dl, s):
    """Host emulation of the device interpolation (fb_kspace.cuh:sqrtp_logtable), float64."""
    s = np.asarray(s, dtype=np.float64)
    out = np.zeros(s.shape)
    ok = s > 0
    x = np.clip((np.log2(s[ok]) - l0) / dl, 1.0, tab.size - 3 + 0.999)
    i = np.floor(x).astype(np.int64)
    f = x - i
    t = tab.astype(np.float64)
    p0, p1, p2, p3 = t[i - 1], t[i], t[i + 1], t[i + 2]
    out[ok] = p1 + 0.5 * f * (p2 - p0 + f * (2 * p0 - 5 * p1 + 4 * p2 - p3 + f * (3 * (p1 - p2) + p3 - p0)))
    return out


def sqrt_pk_bit_table(pk_of_k, N, Lx, Ly, Lz, boxfactor, M=9):
    """
    Table indexed by the leading bits (exponent + M mantissa bits) of float32(s): geometric node
    spacing, linear interpolation inside a segment (fb_kspace.cuh:sqrtp_bittable).
    Returns (table float32, base index, M).
    """
    smin = min(1.0 / Lx ** 2, 1.0 / Ly ** 2, 1.0 / Lz ** 2)
    smax = (N / 2.0) ** 2 * (1.0 / Lx ** 2 + 1.0 / Ly ** 2 + 1.0 / Lz ** 2)
    sh = np.uint32(23 - M)
    k0 = int(np.float32(smin * 0.999).view(np.uint32) >> sh) - 1
    k1 = int(np.float32(smax * 1.001).view(np.uint32) >> sh) + 2
    nodes = (np.arange(k0, k1 + 1, dtype=np.uint32) << sh).view(np.float32).astype(np.float64)
    pk = np.nan_to_num(np.asarray(pk_of_k(TWO_PI * np.sqrt(nodes)), dtype=np.float64))
    return np.sqrt(pk * boxfactor).astype(np.float32), k0, M


def eval_bit_table(tab, base, M, s):
    """Host emulation of fb_kspace.cuh:sqrtp_bittable (s is rounded to float32 like on the device)."""
    s32 = np.asarray(s, dtype=np.float32)
    out = np.zeros(s32.shape)
    ok = s32 > 0
    key = s32[ok].view(np.uint32)
    i = np.clip((key >> np.uint32(23 - M)).astype(np.int64) - base, 0, tab.size - 2)
    frac = (key & np.uint32((1 << (23 - M)) - 1)).astype(np.float64) * 2.0 ** -(23 - M)
    t = tab.astype(np.float64)
    out[ok] = t[i] + frac * (t[i + 1] - t[i])
    return out


def choose_sqrt_pk_table(pk_of_k, N, Lx, Ly, Lz, boxfactor, rtol=2e-6, exact_below=512):
    """
    Pick the sqrt(P) representation for the device.
      mode 1: exact integer LUT (cubic boxes).  Its gathers miss L1, so it is used for small grids
              and whenever the interpolated table cannot be validated;
      mode 3: table indexed by the leading bits of float32(s) + linear interpolation (preferred);
      mode 2: log2(s) table + cubic interpolation;
              both validated here against the exact values at every distinct |k| of a cubic box
              (or 2^20 random modes of a cuboid).
    Returns (mode, table, log2s0, dlog2s).
    """
    cubic = (Lx == Ly == Lz)
    if cubic and N < exact_below:
        return 1, sqrt_pk_int_lut(pk_of_k, N, Lx, boxfactor), 0.0, 0.0
    if cubic:
        n2 = np.arange(1, 3 * (N // 2) ** 2 + 1, dtype=np.float64)
        s = n2 / Lx ** 2
    else:
        rng = np.random.RandomState(2718)
        m = rng.randint(-N // 2, N // 2, size=(1 << 20, 3)).astype(np.float64)
        s = (m[:, 0] / Lx) ** 2 + (m[:, 1] / Ly) ** 2 + (m[:, 2] / Lz) ** 2
        s = s[s > 0]
    exact = np.sqrt(np.nan_to_num(np.asarray(pk_of_k(TWO_PI * np.sqrt(s)), dtype=np.float64)) * boxfactor)
    scale = np.max(np.abs(exact))
    floor = np.maximum(np.abs(exact), 1e-6 * scale)
    for M in (9, 10, 11):                                # float-bit table: cheapest on the device
        tab, base, M = sqrt_pk_bit_table(pk_of_k, N, Lx, Ly, Lz, boxfactor, M)
        if np.max(np.abs(eval_bit_table(tab, base, M, s) - exact) / floor) < rtol:
            return 3, tab, float(base), float(M)
    for npts in (4096, 16384, 65536):
        tab, l0, dl = sqrt_pk_log_table(pk_of_k, N, Lx, Ly, Lz, boxfactor, npts)
        if np.max(np.abs(eval_log_table(tab, l0, dl, s) - exact) / floor) < rtol:
            return 2, tab, l0, dl
    if cubic:
        return 1, sqrt_pk_int_lut(pk_of_k, N, Lx, boxfactor), 0.0, 0.0
    return 2, tab, l0, dl


def log_bin_model(thresholds):
    """
    If the thresholds are (numerically) uniform in log2, return (log2 t0, 1/dlog2) so that the
    device can guess a bin with one log2 and confirm it with two exact comparisons; else (0, 0)
    (binary search).  Exactness never depends on the guess.
    """
    t = np.asarray(thresholds, dtype=np.float64)
    if t.size < 3 or np.any(t <= 0):
        return 0.0, 0.0
    l = np.log2(t)
    d = (l[-1] - l[0]) / (t.size - 1)
    if d <= 0 or np.max(np.abs(l - (l[0] + d * np.arange(t.size)))) > 0.25 * d:
        return 0.0, 0.0
    return float(l[0]), float(1.0 / d)


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
        rng = np.random.RandomState(12345)
        ia, ib, ic = rng.randint(0, h, nprobe), rng.randint(0, N, nprobe), rng.randint(0, N, nprobe)
        raw = np.asarray(transfer_fn(kperp[ia, ib], kpar[ic]))
        if np.iscomplexobj(raw) and np.any(np.nan_to_num(raw.imag) != 0.0):
            raise ValueError("transfer_fn returned complex values: only real transfer functions are supported on "
                             "the GPU path (the tables are real multipliers)")
        if not force_dense:
            # reference point with a non-zero value
            probe = np.nan_to_num(np.asarray(raw.real if np.iscomplexobj(raw) else raw, dtype=np.float64))
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
