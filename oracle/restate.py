"""
TEST INFRASTRUCTURE ONLY -- CPU restatement (NumPy, float64) of the FastBox
field-generation hot path.  Nothing in ``fastbox_b200/`` imports this file; it
is the *checker* used by ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py``.

Every function names the reference lines it restates (paths relative to
``/root/reference``).  Parity status: PINNED -- ``tests/test_oracle_cpu.py``
checks each function against (a) the unmodified reference modules loaded by
``oracle/ref_loader.py`` when ``/root/reference`` is present and (b) the golden
vectors in ``tests/golden/`` that ``oracle/make_golden.py`` produced by running
those reference modules.  Exceptions, which the reference itself delegates to
packages that are neither vendored nor installed (nbodykit) or that cannot be
reproduced bit-for-bit (``np.random.poisson``): ``pk_multipoles``,
``cross_power`` and ``poisson_from_uniform`` are *parity unpinned* and say so.

Two flavours are provided where it matters:
  * ``*_port``  -- follows the reference step by step (full complex128 cube,
    ``numpy.fft`` c2c, masked per-bin loops).  This is what is timed as the CPU
    baseline.
  * lean versions -- half-spectrum / rfft based, same float64 arithmetic to
    ~1e-15, usable at 512^3..1024^3 where the port does not fit in RAM.
"""
import numpy as np

TWO_PI = 2.0 * np.pi


# --------------------------------------------------------------------------
# Grid and wavenumbers                      fastbox/box.py:76-101, 110-127
# --------------------------------------------------------------------------
def box_lengths(box_scale, nsamp):
    """(Lx, Ly, Lz) exactly as box.py:76-89 derives them from linspace ends."""
    if isinstance(box_scale, tuple):
        assert len(box_scale) == 3
        scales = box_scale
    else:
        scales = (box_scale,) * 3
    out = []
    for s in scales:
        g = np.linspace(-0.5 * s, 0.5 * s, nsamp)
        out.append(g[-1] - g[0])
    return tuple(out)


def grid_coords(scale, nsamp):
    return np.linspace(-0.5 * scale, 0.5 * scale, nsamp)        # box.py:79-88


def boxfactor(N, Lx, Ly, Lz):
    return (N ** 6.) / (Lx * Ly * Lz)                            # box.py:94


def kmin_kmax(N, Lx, Ly, Lz):
    kmin = 2. * np.pi / np.max([Lx, Ly, Lz])                     # box.py:100
    kmax = 2. * np.pi * np.sqrt(3.) * N / np.min([Lx, Ly, Lz])   # box.py:101
    return kmin, kmax


def mode_numbers(N):
    """Signed integer mode numbers, box.py:119 ((N*fftfreq).astype('i'))."""
    return (N * np.fft.fftfreq(N, 1.)).astype("i")


def k_grid(N, Lx, Ly, Lz):
    """|k| on the full N^3 grid, same operation order as box.py:125-127."""
    m = mode_numbers(N).astype(np.float64)
    Kx = m[:, None, None]
    Ky = m[None, :, None]
    Kz = m[None, None, :]
    return TWO_PI * np.sqrt((Kx / Lx) ** 2. + (Ky / Ly) ** 2. + (Kz / Lz) ** 2.)


def kperp_kpar(N, Lx, Ly, Lz):
    """box.py:374-375 (broadcastable shapes instead of N^3 arrays)."""
    m = mode_numbers(N).astype(np.float64)
    kperp = TWO_PI * np.sqrt((m[:, None, None] / Lx) ** 2. + (m[None, :, None] / Ly) ** 2.)
    kpar = TWO_PI * m[None, None, :] / Lz
    return kperp, kpar


# --------------------------------------------------------------------------
# realise_density                                   fastbox/box.py:130-194
# --------------------------------------------------------------------------
def realise_density_port(re, im, pk_of_k, N, Lx, Ly, Lz):
    """
    Step-by-step port.  ``pk_of_k(kflat)`` plays CCL's role (box.py:161-165).
    Returns (delta_x float64 N^3, delta_k complex128 N^3 = fftn(delta_x)).
    """
    k = k_grid(N, Lx, Ly, Lz)
    pk = np.reshape(pk_of_k(k.flatten()), k.shape)
    pk = np.nan_to_num(pk)                                       # box.py:167
    pk = pk * boxfactor(N, Lx, Ly, Lz)                           # box.py:171
    noise_k = (re + 1j * im) * np.sqrt(pk)                       # box.py:176
    delta_x = np.fft.ifftn(noise_k).real                         # box.py:187
    delta_k = np.fft.fftn(delta_x)                               # box.py:193
    return delta_x, delta_k


def hermitian_half_from_noise(re, im, amp):
    """
    H(k) = amp(k) * 1/2 [ W(k) + conj W(-k) ],  W = re + i im, on the planes
    kx in [0, N/2] (axis 0 halved).  Identity behind box.py:176-193:
    fftn(Re ifftn(W*amp)) = H when amp(k)=amp(-k).
    ``amp`` must broadcast against (N/2+1, N, N).
    """
    N = re.shape[0]
    h = N // 2 + 1

    def mirror(a):                     # a[(-i)%N, (-j)%N, (-l)%N] restricted to i<=N/2
        ai = np.concatenate([a[:1], a[:0:-1]], axis=0)[:h]        # (-i)%N for i=0..N/2
        ai = np.concatenate([ai[:, :1], ai[:, :0:-1]], axis=1)
        ai = np.concatenate([ai[:, :, :1], ai[:, :, :0:-1]], axis=2)
        return ai
    hr = 0.5 * (re[:h] + mirror(re))
    hi = 0.5 * (im[:h] - mirror(im))
    return (hr + 1j * hi) * amp


def irfft3_axis0(half):
    """Inverse of fftn for a real field given planes kx in [0, N/2] (numpy 1/N^3 norm)."""
    N = half.shape[1]
    import scipy.fft
    return scipy.fft.irfftn(half, s=(N, N, N), axes=(1, 2, 0), workers=-1)


def rfft3_axis0(field):
    import scipy.fft
    return scipy.fft.rfftn(field, axes=(1, 2, 0), workers=-1)


def expand_half_axis0(half):
    """Full N^3 spectrum of a real field from its kx in [0, N/2] planes."""
    N = half.shape[1]
    full = np.empty((N, N, N), dtype=half.dtype)
    h = N // 2 + 1
    full[:h] = half
    rest = half[1:N - h + 1]                                   # kx = 1 .. N/2-1
    m = np.conj(rest[::-1])                                    # kx = N-1 ... -> ordering N/2+1..N-1
    m = np.concatenate([m[:, :1], m[:, :0:-1]], axis=1)
    m = np.concatenate([m[:, :, :1], m[:, :, :0:-1]], axis=2)
    full[h:] = m
    return full


def sqrt_pk_half(pk_of_k, N, Lx, Ly, Lz):
    """sqrt(P(k) * boxfactor) on the kx in [0,N/2] half grid (box.py:161-171)."""
    m = mode_numbers(N).astype(np.float64)
    h = N // 2 + 1
    k = TWO_PI * np.sqrt((m[:h, None, None] / Lx) ** 2. + (m[None, :, None] / Ly) ** 2.
                         + (m[None, None, :] / Lz) ** 2.)
    pk = np.nan_to_num(np.reshape(pk_of_k(k.ravel()), k.shape))
    return np.sqrt(pk * boxfactor(N, Lx, Ly, Lz))


def realise_density_lean(re, im, pk_of_k, N, Lx, Ly, Lz):
    """Half-spectrum equivalent of ``realise_density_port`` (agrees to ~1e-15)."""
    amp = sqrt_pk_half(pk_of_k, N, Lx, Ly, Lz)
    half = hermitian_half_from_noise(re, im, amp)
    return irfft3_axis0(half), half


# --------------------------------------------------------------------------
# Slab-wise restatement for sizes the ports above cannot hold (512^3 .. 1024^3):
# same arithmetic as ``realise_density_lean`` + ``binned_power_spectrum_lean``
# (both pinned bit-for-bit to the reference at small N; this function is checked
# against them in tests/test_oracle_cpu.py), one kx plane at a time, float64.
# --------------------------------------------------------------------------
def k_plane(a, N, Lx, Ly, Lz):
    """|k| on the plane kx index ``a`` with the operation order of box.py:125-127."""
    m = mode_numbers(N).astype(np.float64)
    return TWO_PI * np.sqrt((m[a] / Lx) ** 2. + (m[:, None] / Ly) ** 2. + (m[None, :] / Lz) ** 2.)


def _mirror2(p):
    p = np.concatenate([p[:1], p[:0:-1]], axis=0)
    return np.concatenate([p[:, :1], p[:, :0:-1]], axis=1)


def hermitian_plane(re_a, im_a, re_m, im_m):
    """One kx plane of ``hermitian_half_from_noise`` (amp = 1): re_m, im_m are the noise planes (-a) % N."""
    return 0.5 * (re_a + _mirror2(re_m)) + 0.5j * (im_a - _mirror2(im_m))


def kspace_factor_plane(kind, a, N, Lx, Ly, Lz):
    """Per-mode factor of box.py:254-274 (velocity components, without `fac`) / box.py:347-348 (potential)."""
    m = mode_numbers(N).astype(np.float64)
    k2 = k_plane(a, N, Lx, Ly, Lz) ** 2.
    with np.errstate(divide="ignore", invalid="ignore"):
        if kind == "potential":
            f = 1.0 / k2
            if a == 0:
                f[0, 0] = 0.
            return f
        ax = {"vel_x": 0, "vel_y": 1, "vel_z": 2}[kind]
        K = [np.full((1, 1), m[a]), m[:, None], m[None, :]][ax]
        L = (Lx, Ly, Lz)[ax]
        f = np.nan_to_num(1.j * K * (TWO_PI / L) / k2)
    if ax == 0 and a == N // 2:                                  # box.py:268-274
        f[:] = 0.
    elif ax == 1:
        f[N // 2, :] = 0.
    elif ax == 2:
        f[:, N // 2] = 0.
    return f


def realise_slabwise(noise_plane_fn, pk_of_k, N, Lx, Ly, Lz, transfer_fn=None, kinds=(None,), nbins=None,
                     x_planes=None, workers=4):
    """
    box.py:161-193 (+ :374-380 filter, :254-274 / :347 factors, :741-764 moments) plane by plane.
    ``noise_plane_fn(a) -> (re, im)``: the (N, N) noise planes W[a] (box.py:174-175).
    ``kinds``: which fields to build from the one spectrum: None = density (Re ifftn), "vel_x" / "vel_y" /
    "vel_z" (without `fac`), "potential".
    Returns dict(fields={kind: array}, count, sum1, sum2, edges): each field is the full float64 (N,N,N) cube,
    or only the x planes listed in ``x_planes`` (shape (len, N, N)); moments (of the density spectrum incl.
    the filter, Hermitian multiplicities, index = np.digitize) only when ``nbins`` is given.
    """
    import scipy.fft
    from concurrent.futures import ThreadPoolExecutor
    h = N // 2 + 1
    kinds = list(kinds)
    bf = boxfactor(N, Lx, Ly, Lz)
    bins = pk_bin_edges(N, Lx, Ly, Lz, nbins) if nbins else None
    m = mode_numbers(N).astype(np.float64)
    wts = half_weights(N)
    mabs = np.abs(mode_numbers(N)).astype(np.int64)
    xs = None if x_planes is None else np.asarray(x_planes, dtype=np.int64)
    G = {kd: np.empty((h, N, N), dtype=np.complex128) for kd in kinds} if xs is None else None

    def one_block(a_list):
        part = None if xs is None else {kd: np.zeros((xs.size, N, N)) for kd in kinds}
        mom = None if bins is None else [np.zeros(bins.size + 1) for _ in range(3)]
        for a in a_list:
            re_a, im_a = noise_plane_fn(a)
            re_m, im_m = (re_a, im_a) if (N - a) % N == a else noise_plane_fn((N - a) % N)
            k = k_plane(a, N, Lx, Ly, Lz)
            # P(k) depends on (|m_y|, |m_z|) only and (-m/L)^2 == (m/L)^2 bitwise: evaluate one quadrant
            kq = k[:N // 2 + 1, :N // 2 + 1]
            ampq = np.sqrt(np.nan_to_num(np.reshape(pk_of_k(kq.ravel()), kq.shape)) * bf)     # box.py:161-171
            amp = ampq[mabs[:, None], mabs[None, :]]
            if transfer_fn is not None:                                                     # box.py:374-379
                kperp = TWO_PI * np.sqrt((m[a] / Lx) ** 2. + (m[:, None] / Ly) ** 2.)
                kpar = TWO_PI * m[None, :] / Lz
                amp = amp * np.nan_to_num(transfer_fn(np.broadcast_to(kperp, (N, N)), np.broadcast_to(kpar, (N, N))))
            H = hermitian_plane(np.asarray(re_a, np.float64), np.asarray(im_a, np.float64),
                                np.asarray(re_m, np.float64), np.asarray(im_m, np.float64)) * amp
            if bins is not None:
                idx = np.digitize(k.ravel(), bins)
                power = (H * np.conj(H)).real.ravel() / bf
                c, s1, s2 = pk_moments(power, idx, bins.size, np.full(power.size, wts[a]))
                mom[0] += c
                mom[1] += s1
                mom[2] += s2
            for kd in kinds:
                S = H if kd is None else H * kspace_factor_plane(kd, a, N, Lx, Ly, Lz)
                g = scipy.fft.ifft2(S)                           # y, z of ifftn (1/N^2)
                if xs is None:
                    G[kd][a] = g
                else:                                            # x of ifftn for the requested planes only
                    ph = TWO_PI * ((a * xs) % N) / N             # Re(g e^{i ph}) = g.re cos - g.im sin
                    coef = (wts[a] / N) * np.stack([np.cos(ph), -np.sin(ph)], axis=1)
                    part[kd] += (coef @ g.view(np.float64).reshape(N * N, 2).T).reshape(xs.size, N, N)
        return part, mom

    workers = max(1, int(workers))
    blocks = [list(range(w, h, workers)) for w in range(workers)]
    with ThreadPoolExecutor(workers) as ex:
        results = list(ex.map(one_block, blocks))
    out = {"fields": {}}
    for kd in kinds:
        if xs is None:
            out["fields"][kd] = scipy.fft.irfft(G[kd], n=N, axis=0, workers=-1)
            G[kd] = None
        else:
            out["fields"][kd] = np.sum([r[0][kd] for r in results], axis=0)
    if bins is not None:
        out["count"] = np.rint(sum(r[1][0] for r in results)).astype(np.int64)
        out["sum1"] = sum(r[1][1] for r in results)
        out["sum2"] = sum(r[1][2] for r in results)
        out["edges"] = bins
    return out


# --------------------------------------------------------------------------
# realise_velocity / realise_potential              fastbox/box.py:197-353
# --------------------------------------------------------------------------
def velocity_k_port(delta_k, N, Lx, Ly, Lz, fac):
    """box.py:251-285.  ``fac`` = 100 h E(a) f(a) a (box.py:280-281)."""
    m = mode_numbers(N).astype(np.float64)
    Kx, Ky, Kz = m[:, None, None], m[None, :, None], m[None, None, :]
    k2 = k_grid(N, Lx, Ly, Lz) ** 2.
    with np.errstate(divide="ignore", invalid="ignore"):
        comps = [np.nan_to_num(1.j * delta_k * K * (TWO_PI / L) / k2)
                 for K, L in ((Kx, Lx), (Ky, Ly), (Kz, Lz))]
    if N % 2 == 0:                                               # box.py:268-274
        comps[0][N // 2, :, :] = 0.
        comps[1][:, N // 2, :] = 0.
        comps[2][:, :, N // 2] = 0.
    else:
        raise NameError("reference fails for odd N (box.py:268-274)")
    return tuple(c * fac for c in comps)


def potential_k_port(delta_k, N, Lx, Ly, Lz):
    """box.py:347-348 (the prefactor computed at :343-345 is never applied)."""
    with np.errstate(divide="ignore", invalid="ignore"):
        phi = delta_k / k_grid(N, Lx, Ly, Lz) ** 2.
    phi[0, 0, 0] = 0.
    return phi


# --------------------------------------------------------------------------
# apply_transfer_fn / smooth_field             fastbox/box.py:356-381,615-655
# --------------------------------------------------------------------------
def apply_transfer_fn_port(field_k, transfer_fn, N, Lx, Ly, Lz):
    kperp, kpar = kperp_kpar(N, Lx, Ly, Lz)
    kperp = np.broadcast_to(kperp, (N, N, N))
    kpar = np.broadcast_to(kpar, (N, N, N))
    dk = np.nan_to_num(field_k * transfer_fn(kperp, kpar))       # box.py:378-379
    return np.fft.ifftn(dk)                                      # box.py:380 (complex)


def tophat_window(k, R):
    """box.py:631-633."""
    x = k * R
    with np.errstate(divide="ignore", invalid="ignore"):
        return (3. / x ** 3.) * (np.sin(x) - x * np.cos(x))


def smooth_field_port(field_k, R, h, N, Lx, Ly, Lz):
    dk = np.nan_to_num(field_k * tophat_window(k_grid(N, Lx, Ly, Lz), R / h))   # box.py:652-653
    return np.fft.ifftn(dk)


# --------------------------------------------------------------------------
# lognormal                                         fastbox/box.py:441-460
# --------------------------------------------------------------------------
def lognormal(delta_x):
    e = np.exp(delta_x)
    return e / np.mean(e) - 1.


# --------------------------------------------------------------------------
# binned_power_spectrum                             fastbox/box.py:696-768
# --------------------------------------------------------------------------
def pk_bin_edges(N, Lx, Ly, Lz, nbins=20, kbins=None):
    if kbins is not None:
        return np.asarray(kbins, dtype=np.float64)
    kmin, kmax = kmin_kmax(N, Lx, Ly, Lz)
    return np.logspace(np.log10(kmin), np.log10(kmax), nbins)    # box.py:749


def bin_centres(bins):
    full = [0.0] + list(bins)                                    # box.py:750-751
    return np.array([0.5 * (full[j + 1] + full[j]) for j in range(len(bins))])


def digitize_modes(N, Lx, Ly, Lz, bins):
    """Integer bin index of every mode of the full grid (box.py:758)."""
    return np.digitize(k_grid(N, Lx, Ly, Lz).flatten(), bins)


def binned_power_spectrum_port(delta_k, N, Lx, Ly, Lz, nbins=20, kbins=None,
                               return_raw=False):
    """box.py:741-768 verbatim in structure: per-bin masked mean / std."""
    pk = (delta_k * np.conj(delta_k)).real / boxfactor(N, Lx, Ly, Lz)
    bins = pk_bin_edges(N, Lx, Ly, Lz, nbins, kbins)
    cent = bin_centres(bins)
    idxs = digitize_modes(N, Lx, Ly, Lz, bins)
    flat = pk.flatten()
    vals = np.zeros(bins.size)
    err = np.zeros(bins.size)
    counts = np.zeros(bins.size, dtype=np.int64)
    with np.errstate(all="ignore"):
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            for i in range(bins.size):
                sel = flat[idxs == i]
                counts[i] = sel.size
                vals[i] = np.mean(sel)
                err[i] = np.std(sel) / np.sqrt(sel.size)
    if return_raw:
        return cent[1:], vals[1:], err[1:], counts, idxs
    return cent[1:], vals[1:], err[1:]


def pk_moments(power, idxs, nb, weights=None):
    """(count, sum w p, sum w p^2) for bin indices 0..nb (index nb = overflow)."""
    w = np.ones_like(power) if weights is None else weights
    cnt = np.bincount(idxs, weights=w, minlength=nb + 1)
    s1 = np.bincount(idxs, weights=w * power, minlength=nb + 1)
    s2 = np.bincount(idxs, weights=w * power * power, minlength=nb + 1)
    return cnt, s1, s2


def half_weights(N):
    """Multiplicity of each kx in [0,N/2] plane in the full spectrum."""
    w = np.full(N // 2 + 1, 2.0)
    w[0] = 1.0
    if N % 2 == 0:
        w[-1] = 1.0
    return w


def digitize_half(N, Lx, Ly, Lz, bins):
    m = mode_numbers(N).astype(np.float64)
    h = N // 2 + 1
    k = TWO_PI * np.sqrt((m[:h, None, None] / Lx) ** 2. + (m[None, :, None] / Ly) ** 2.
                         + (m[None, None, :] / Lz) ** 2.)
    return np.digitize(k.ravel(), bins).reshape(k.shape)


def binned_power_spectrum_lean(half_a, N, Lx, Ly, Lz, nbins=20, kbins=None, half_b=None):
    """
    Same estimator from kx in [0,N/2] planes with multiplicity weights
    (reproduces the port's mean *and* stddev to ~3e-15).  With ``half_b`` the
    cross spectrum Re[a conj b] is binned instead (PARITY UNPINNED: the
    reference delegates cross-power to nbodykit, example_halos.py:52-53).
    """
    bins = pk_bin_edges(N, Lx, Ly, Lz, nbins, kbins)
    idx = digitize_half(N, Lx, Ly, Lz, bins)
    other = half_a if half_b is None else half_b
    power = (half_a * np.conj(other)).real / boxfactor(N, Lx, Ly, Lz)
    w = np.broadcast_to(half_weights(N)[:, None, None], power.shape)
    cnt, s1, s2 = pk_moments(power.ravel(), idx.ravel(), bins.size, w.ravel())
    with np.errstate(all="ignore"):
        mean = s1 / cnt
        var = np.maximum(s2 / cnt - mean * mean, 0.0)
        err = np.sqrt(var) / np.sqrt(cnt)
    cent = bin_centres(bins)
    return cent[1:], mean[1:bins.size], err[1:bins.size], cnt.astype(np.int64)


def pk_multipoles(half, N, Lx, Ly, Lz, nbins=20, kbins=None, ells=(0, 2, 4)):
    """
    PARITY UNPINNED w.r.t. nbodykit (the reference uses its FFTPower, example_box.py:48-52; absent here).
    Tied instead to the published Kaiser / Hamilton multipoles of (1 + beta mu^2)^2 P(k) -- which fix the (2l+1)
    factor, the Legendre polynomials and the line of sight -- in
    tests/test_oracle_cpu.py::test_multipoles_and_cross_power_known_answers.
    P_l(k) = (2l+1) < |d_k|^2 L_l(mu) >_bin / boxfactor, mu = k_z/|k|, LOS = z,
    same bins / weights as ``binned_power_spectrum_lean``.
    """
    bins = pk_bin_edges(N, Lx, Ly, Lz, nbins, kbins)
    idx = digitize_half(N, Lx, Ly, Lz, bins)
    m = mode_numbers(N).astype(np.float64)
    h = N // 2 + 1
    kz = TWO_PI * m[None, None, :] / Lz
    k = TWO_PI * np.sqrt((m[:h, None, None] / Lx) ** 2. + (m[None, :, None] / Ly) ** 2.
                         + (m[None, None, :] / Lz) ** 2.)
    with np.errstate(all="ignore"):
        mu = np.where(k > 0, kz / k, 0.0)
    power = (half * np.conj(half)).real / boxfactor(N, Lx, Ly, Lz)
    w = np.broadcast_to(half_weights(N)[:, None, None], power.shape)
    out = {}
    cnt = np.bincount(idx.ravel(), weights=w.ravel(), minlength=bins.size + 1)
    for ell in ells:
        if ell == 0:
            leg = np.ones_like(mu)
        elif ell == 2:
            leg = 0.5 * (3 * mu ** 2 - 1)
        elif ell == 4:
            leg = (35 * mu ** 4 - 30 * mu ** 2 + 3) / 8.0
        else:
            raise ValueError(ell)
        s = np.bincount(idx.ravel(), weights=(w * power * leg).ravel(), minlength=bins.size + 1)
        with np.errstate(all="ignore"):
            out[ell] = (2 * ell + 1) * (s / cnt)[1:bins.size]
    return bin_centres(bins)[1:], out


# --------------------------------------------------------------------------
# P(k_perp, k_par) and xi(r)      (reference: nbodykit FFTPower mode='2d' / FFTCorr mode='1d',
# examples/example_endtoend.py:128-151 -- nbodykit is neither vendored nor pinned, so these
# definitions are PARITY UNPINNED w.r.t. nbodykit; they reuse the conventions of
# binned_power_spectrum: np.digitize bins, multiplicity weights, population stddev / sqrt(n))
# --------------------------------------------------------------------------
def binned_power_spectrum_2d_lean(half_a, N, Lx, Ly, Lz, kperp_bins, kpar_bins, half_b=None):
    """
    Moments of |d_k|^2 / boxfactor (or Re a conj b) binned in k_perp = 2 pi sqrt((Kx/Lx)^2 + (Ky/Ly)^2) and
    |k_par| = 2 pi |Kz| / Lz.  Returns (kperp centres[1:], kpar centres[1:], mean, err, count) with mean / err of
    shape (len(kperp_bins) - 1, len(kpar_bins) - 1) and count the full (nperp + 1, npar + 1) population table.
    """
    kperp_bins = np.asarray(kperp_bins, dtype=np.float64)
    kpar_bins = np.asarray(kpar_bins, dtype=np.float64)
    m = mode_numbers(N).astype(np.float64)
    h = N // 2 + 1
    kperp = TWO_PI * np.sqrt((m[:h, None] / Lx) ** 2. + (m[None, :] / Ly) ** 2.)
    kpar = np.abs(TWO_PI * m / Lz)
    ip = np.digitize(kperp.ravel(), kperp_bins).reshape(kperp.shape)
    il = np.digitize(kpar, kpar_bins)
    npar = kpar_bins.size
    idx = (ip[:, :, None] * (npar + 1) + il[None, None, :]).ravel()
    other = half_a if half_b is None else half_b
    power = (half_a * np.conj(other)).real.ravel() / boxfactor(N, Lx, Ly, Lz)
    w = np.broadcast_to(half_weights(N)[:, None, None], (h, N, N)).ravel()
    nb2 = (kperp_bins.size + 1) * (npar + 1)
    cnt, s1, s2 = pk_moments(power, idx, nb2 - 1, w)
    shape = (kperp_bins.size + 1, npar + 1)
    cnt, s1, s2 = cnt.reshape(shape), s1.reshape(shape), s2.reshape(shape)
    with np.errstate(all="ignore"):
        mean = s1 / cnt
        err = np.sqrt(np.maximum(s2 / cnt - mean * mean, 0.0)) / np.sqrt(cnt)
    sl = (slice(1, kperp_bins.size), slice(1, npar))
    return bin_centres(kperp_bins)[1:], bin_centres(kpar_bins)[1:], mean[sl], err[sl], np.rint(cnt).astype(np.int64)


def lag_separations(N, Lx, Ly, Lz):
    """|r| of every lag of the periodic N^3 grid, cell size L/N: sqrt(((dx hx)^2 + (dy hy)^2) + (dz hz)^2)."""
    d = np.minimum(np.arange(N), N - np.arange(N)).astype(np.float64)
    dx, dy, dz = d * (Lx / N), d * (Ly / N), d * (Lz / N)
    return np.sqrt((dx[:, None, None] * dx[:, None, None] + dy[None, :, None] * dy[None, :, None])
                   + dz[None, None, :] * dz[None, None, :])


def correlation_function_port(field_a, N, Lx, Ly, Lz, edges, field_b=None):
    """
    xi(r) = < a(x) b(x + r) > by the Wiener-Khinchin route, xi = ifftn(A conj B) / N^3, averaged over the lags
    whose separation falls in [edges[i-1], edges[i]).  Returns (centres, mean, stddev / sqrt(n), count[nedges+1]).
    """
    edges = np.asarray(edges, dtype=np.float64)
    A = np.fft.fftn(np.asarray(field_a, dtype=np.float64))
    B = A if field_b is None else np.fft.fftn(np.asarray(field_b, dtype=np.float64))
    xi = np.fft.ifftn(A * np.conj(B)).real / float(N) ** 3
    idx = np.digitize(lag_separations(N, Lx, Ly, Lz).ravel(), edges)
    cnt, s1, s2 = pk_moments(xi.ravel(), idx, edges.size)
    with np.errstate(all="ignore"):
        mean = s1 / cnt
        err = np.sqrt(np.maximum(s2 / cnt - mean * mean, 0.0)) / np.sqrt(cnt)
    cent = 0.5 * (edges[1:] + edges[:-1])
    return cent, mean[1:edges.size], err[1:edges.size], np.rint(cnt).astype(np.int64)


# --------------------------------------------------------------------------
# redshift_space_density                            fastbox/box.py:384-438
#   scipy griddata 1-D linear == argsort + interp1d(linear, fill_value)
#   (scipy/interpolate/_ndgriddata.py:315-330, _interpolate.py:491-518,592-593)
# --------------------------------------------------------------------------
def rsd_remap_line(z, dens, vel_total, Hz, method="linear"):
    s = z - vel_total / Hz                                        # box.py:422
    zmin = np.min(z)
    length = np.max(z) - zmin
    s = (s - zmin) % length + zmin                                # box.py:425-426
    fill = 0.5 * (dens[0] + dens[-1])                             # box.py:429
    order = np.argsort(s)
    xs = s[order]
    ys = dens[order]
    if method == "nearest":
        # griddata 1-D 'nearest' = interp1d(kind='nearest', fill_value='extrapolate'): mid-points
        # x/2 + x/2 and searchsorted(side='left') (scipy/interpolate/_interpolate.py: _call_nearest)
        half = xs / 2.0
        bds = half[1:] + half[:-1]
        return ys[np.searchsorted(bds, z, side="left").clip(0, xs.size - 1)]
    if method != "linear":
        raise ValueError("rsd_remap_line: method %r" % (method,))
    hi = np.searchsorted(xs, z).clip(1, xs.size - 1)
    lo = hi - 1
    with np.errstate(all="ignore"):
        slope = (ys[hi] - ys[lo]) / (xs[hi] - xs[lo])
        val = slope * (z - xs[lo]) + ys[lo]
    out = np.where((z < xs[0]) | (z > xs[-1]), fill, val)
    return out


def redshift_space_density(delta_x, velocity_z, z, Hz, vel_nl=None, method="linear"):
    """vel_nl: optional (N,N,N) array = sigma_nl * N(0,1) drawn line by line (box.py:418)."""
    out = np.empty_like(delta_x)
    for i in range(delta_x.shape[0]):
        for j in range(delta_x.shape[1]):
            v = velocity_z[i, j, :] + (0. if vel_nl is None else vel_nl[i, j, :])
            out[i, j, :] = rsd_remap_line(z, delta_x[i, j, :], v, Hz, method)
    return out


# --------------------------------------------------------------------------
# BeamModel.convolve_fft                           fastbox/beams.py:63-87
#   scipy.signal.fftconvolve(beam, field, 'same', axes=[0,1]) = zero-padded
#   linear convolution, cropped to the first argument's frame
#   (scipy/signal/_signaltools.py: _centered)
# --------------------------------------------------------------------------
def convolve_fft(beam, field):
    N = beam.shape[0]
    P = 2 * N                              # >= 2N-1 (next_fast_len(2N-1)=2N for N=2^m)
    fb = np.fft.rfft2(beam, s=(P, P), axes=(0, 1))
    ff = np.fft.rfft2(field, s=(P, P), axes=(0, 1))
    full = np.fft.irfft2(fb * ff, s=(P, P), axes=(0, 1))
    st = (N - 1) // 2
    sm = full[st:st + N, st:st + N, :]
    norm = np.sum(beam.reshape(-1, beam.shape[-1]), axis=0)       # beams.py:81
    return sm / norm[np.newaxis, np.newaxis, :]                   # beams.py:87


# --------------------------------------------------------------------------
# HaloDistribution.halo_count_field                fastbox/halos.py:91-117
# --------------------------------------------------------------------------
def halo_mean_count(delta_x, nbar, bias, Lx, Ly, Lz, lognormal_tf=False):
    """Expected count per voxel (everything in halos.py:91-113 before the Poisson draw)."""
    N = delta_x.shape[0]
    nbar = np.atleast_1d(nbar)
    bias = np.atleast_1d(bias)
    if nbar.ndim == 1:
        nbar = nbar[np.newaxis, np.newaxis, :]
    if bias.ndim == 1:
        bias = bias[np.newaxis, np.newaxis, :]
    vol = Lx * Ly * Lz / N ** 3.
    dh = bias * delta_x
    if lognormal_tf:
        dh = np.exp(dh)
        dh = dh / np.mean(dh)
        dh = dh - 1.
    mean = vol * nbar * (1. + dh)
    if not lognormal_tf:
        mean = np.where(mean < 0., 0., mean)
    return np.nan_to_num(mean)


def _exp_neg_scaled(lam):
    """exp(-lam) = pm * 2^e from IEEE + * / only (shared with include/fb_poisson.h: bit-identical)."""
    lam = np.asarray(lam, dtype=np.float64)
    n = np.floor(lam * 1.4426950408889634 + 0.5)
    r = (lam - n * 0.693147180369123816490e+00) - n * 1.90821492927058770002e-10
    r = -r
    # exp(r), |r| <= 0.35: Taylor to degree 13 in Horner form
    p = np.full_like(r, 1.0 / 6227020800.0)
    for c in (1.0 / 479001600.0, 1.0 / 39916800.0, 1.0 / 3628800.0, 1.0 / 362880.0,
              1.0 / 40320.0, 1.0 / 5040.0, 1.0 / 720.0, 1.0 / 120.0, 1.0 / 24.0,
              1.0 / 6.0, 0.5, 1.0, 1.0):
        p = p * r + c
    return p, (-n).astype(np.int64)


def _exp_neg(lam):
    p, e = _exp_neg_scaled(lam)
    return np.ldexp(p, e)


def poisson_from_uniform(lam, u, kmax=1 << 22):
    """
    PARITY UNPINNED w.r.t. ``np.random.poisson`` (halos.py:116; legacy MT19937
    stream consumes a variable number of uniforms per sample).  Definition
    shared bit-for-bit with the CUDA kernel: inversion by sequential search
        p0 = exp(-lam); count = #{ k : u > sum_{j<=k} p_j },  p_k = p_{k-1} lam / k
    in float64 using only + * / (no FMA contraction); the term is carried as mantissa * 2^e (rescaled by
    exact powers of two) so that lam of any size works although exp(-lam) underflows from ~745 on.
    """
    lam = np.asarray(lam, dtype=np.float64)
    u = np.asarray(u, dtype=np.float64)
    if lam.size and np.any(lam >= 700.0) and np.any(lam < 700.0):         # the header branches at 700
        out = np.empty(lam.shape, dtype=np.int64)
        hi = lam >= 700.0
        out[hi] = poisson_from_uniform(lam[hi], u[hi], kmax)
        out[~hi] = poisson_from_uniform(lam[~hi], u[~hi], kmax)
        return out
    if lam.size and np.all(lam < 700.0):                                   # plain recurrence
        p = _exp_neg(lam)
        cdf = p.copy()
        k = np.zeros(lam.shape, dtype=np.int64)
        active = u > cdf
        it = 0
        while np.any(active) and it < 100000:
            it += 1
            k = np.where(active, k + 1, k)
            p = np.where(active, p * lam / k.clip(1), p)
            cdf = np.where(active, cdf + p, cdf)
            active = active & (u > cdf) & (p > 0.0)
        return k
    big, small = 1.3407807929942597e+154, 7.458340731200207e-155          # 2^512, 2^-512
    with np.errstate(all="ignore"):
        pm, e = _exp_neg_scaled(np.where(lam > 0.0, lam, 1.0))
        p = np.ldexp(pm, e)
        cdf = p.copy()
        k = np.zeros(lam.shape, dtype=np.int64)
        active = (lam > 0.0) & (u > cdf) & ((p > 0.0) | (k < lam))
        it = 0
        while np.any(active) and it < kmax:
            it += 1
            k = np.where(active, k + 1, k)
            pm = np.where(active, pm * lam / k.clip(1), pm)
            up, down = active & (pm > big), active & (pm < small)
            pm = np.where(up, pm * small, np.where(down, pm * big, pm))
            e = e + 512 * up.astype(np.int64) - 512 * down.astype(np.int64)
            p = np.where(active, np.ldexp(pm, e), p)
            cdf = np.where(active, cdf + p, cdf)
            active = active & (u > cdf) & ((p > 0.0) | (k < lam))
    return k


def halo_catalogue_port(Nhalo, Lx, Ly, Lz, uniforms=None):
    """
    Step-by-step restatement of HaloDistribution.realise_halo_catalogue
    (halos.py:142-176): loop over the distinct count values (halos.py:146-147),
    voxels of each value in C order repeated ``count`` times (halos.py:152-155),
    float64 indices (halos.py:158-160), optional offsets in catalogue order
    (halos.py:163-166; the reference draws them with np.random.uniform), box scaling
    (halos.py:172-174).  Pinned against the unmodified reference in
    tests/golden/halo_catalogue.npz.
    """
    Nhalo = np.asarray(Nhalo)
    N = Nhalo.shape[0]
    cx, cy, cz = [], [], []
    for count in np.unique(Nhalo):
        if count < 1:
            continue
        ix, iy, iz = np.where(Nhalo == count)
        cx.append(np.repeat(ix, count))
        cy.append(np.repeat(iy, count))
        cz.append(np.repeat(iz, count))
    if not cx:
        return np.empty((0, 3), dtype=np.float64)
    cat = np.column_stack((np.concatenate(cx), np.concatenate(cy), np.concatenate(cz))).astype(np.float64)
    if uniforms is not None:
        cat += np.asarray(uniforms, dtype=np.float64).reshape(cat.shape)
    cat[:, 0] *= Lx / N
    cat[:, 1] *= Ly / N
    cat[:, 2] *= Lz / N
    return cat


def halo_catalogue_lean(Nhalo, Lx, Ly, Lz, uniforms=None):
    """Same catalogue as a stable sort of the occupied voxels by count (the device algorithm's view)."""
    Nhalo = np.asarray(Nhalo)
    N = Nhalo.shape[0]
    flat = Nhalo.ravel()
    vox = np.flatnonzero(flat > 0)
    order = np.argsort(flat[vox], kind="stable")
    vox = np.repeat(vox[order], flat[vox][order])
    cat = np.column_stack(np.unravel_index(vox, Nhalo.shape)).astype(np.float64).reshape(-1, 3)
    if uniforms is not None:
        cat += np.asarray(uniforms, dtype=np.float64).reshape(cat.shape)
    cat[:, 0] *= Lx / N
    cat[:, 1] *= Ly / N
    cat[:, 2] *= Lz / N
    return cat


def fg_construct_cube_port(amps, spectral_idx, freqs, freq_ref=130.):
    """ForegroundModel.construct_cube, foregrounds.py:167-174 (float64 power law per pixel and channel)."""
    freqs = np.asarray(freqs, dtype=np.float64)
    if isinstance(spectral_idx, float):
        ffac = ((freqs / freq_ref) ** spectral_idx)[np.newaxis, np.newaxis, :]
    else:
        ffac = (freqs / freq_ref)[np.newaxis, np.newaxis, :] ** np.asarray(spectral_idx)[:, :, np.newaxis]
    return np.asarray(amps)[:, :, np.newaxis] * ffac


def mean_spectrum_filter_port(field):
    """filters.mean_spectrum_filter, filters.py:49-55: subtract the pixel-averaged spectrum."""
    d = np.asarray(field).reshape((-1, field.shape[-1]))
    d_mean = np.mean(d, axis=0)[np.newaxis, :]
    return (d - d_mean).reshape(field.shape)


def pca_filter_port(field, nmodes, fit_powerlaw=False, return_filter=False):
    """
    filters.pca_filter, filters.py:138-183, step by step: mean spectrum (optionally its power-law fit,
    filters.py:146-155), np.cov of the mean-subtracted channels, np.linalg.eig, modes sorted by
    eigenvalue, amplitudes U^T x per line of sight, foreground model U a + mean subtracted.
    """
    d = field.reshape((-1, field.shape[-1])).T
    d_mean = np.mean(d, axis=-1)[:, np.newaxis]
    if fit_powerlaw:
        from scipy.optimize import curve_fit
        freqs = np.linspace(1., 10., d.shape[0])

        def fn(nu, amp, beta):
            return amp * (nu / nu[0]) ** beta
        pfit, _ = curve_fit(fn, freqs, d_mean.flatten(), p0=[d_mean[0][0], -2.7])
        d_mean = fn(freqs, pfit[0], pfit[1])[:, np.newaxis]
    x = d - d_mean
    cov = np.cov(x)
    eigvals, eigvecs = np.linalg.eig(cov)
    idxs = np.argsort(eigvals)[::-1]
    eigvecs = eigvecs[:, idxs]
    U_fg = eigvecs[:, :nmodes]
    fg_amps = np.dot(U_fg.T, x)
    fg_field = (np.dot(U_fg, fg_amps) + d_mean).T.reshape(field.shape)
    cleaned = field - fg_field
    return (cleaned, U_fg, fg_amps) if return_filter else cleaned


def radiometer_rms_port(freqs, ang_x, Tinst, tp, fov, Ndish):
    """NoiseModel.realise_radiometer_noise up to the rms per channel, noise.py:55-69."""
    freqs = np.asarray(freqs, dtype=np.float64)
    dnu = np.abs(freqs[1] - freqs[0])
    tp = tp * 3600.
    dtheta = ang_x[1] - ang_x[0]
    t_res = tp * dtheta ** 2. / fov
    Tsky = 60e3 * (freqs / 300.) ** (-2.5)
    Tsys = Tinst * 1e3 + Tsky
    return Tsys / np.sqrt(Ndish * t_res * (dnu * 1e6))


def radiometer_noise_port(sigma_rms, normals):
    """noise.py:72-75: unit white noise times the per-channel rms."""
    return np.asarray(normals, dtype=np.float64) * np.asarray(sigma_rms)[np.newaxis, np.newaxis, :]


def philox_noise_cube(seed, N):
    """
    PARITY UNPINNED (the reference draws np.random.normal): the device's counter-based unit normals
    of the noise cube.  Philox4x32-10 block b = cell index // 4 with counter (b_lo, b_hi, 'NOIS', 0) and
    key = seed; words (0,1) and (2,3) feed two Box-Muller transforms giving cells 4b .. 4b+3.
    """
    n4 = N ** 3 // 4
    b = np.arange(n4, dtype=np.uint64)
    ctr = np.zeros((n4, 4), dtype=np.uint32)
    ctr[:, 0] = (b & np.uint64(0xFFFFFFFF)).astype(np.uint32)
    ctr[:, 1] = (b >> np.uint64(32)).astype(np.uint32)
    ctr[:, 2] = np.uint32(0x4e4f4953)
    key = np.array([seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF], dtype=np.uint32)
    x = philox4x32(ctr, key).astype(np.float64)
    out = np.empty((n4, 4))
    for j in (0, 2):
        u1, u2 = (x[:, j] + 0.5) * 2.0 ** -32, (x[:, j + 1] + 0.5) * 2.0 ** -32
        r = np.sqrt(-2.0 * np.log(u1))
        out[:, j], out[:, j + 1] = r * np.cos(TWO_PI * u2), r * np.sin(TWO_PI * u2)
    return out.reshape(N, N, N)


# --------------------------------------------------------------------------
# Counter-based white noise (no reference equivalent: the reference draws
# from NumPy's global MT19937, box.py:174-175): PARITY UNPINNED by construction.
# Philox4x32-10 keyed by (seed, global cell index) + Box-Muller; the Hermitian part
# H0 of the white field is drawn directly (it has the law of the reference's
# 1/2 [W(k) + conj W(-k)]).  Used for throughput / multi-GPU runs.
# --------------------------------------------------------------------------
_PH_M0, _PH_M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
_PH_W0, _PH_W1 = np.uint32(0x9E3779B9), np.uint32(0xBB67AE85)


def philox4x32(ctr, key):
    """ctr: (...,4) uint32, key: (2,) uint32 -> (...,4) uint32; 10 rounds."""
    c = [ctr[..., i].astype(np.uint32) for i in range(4)]
    k0, k1 = np.uint32(key[0]), np.uint32(key[1])
    mask = np.uint64(0xFFFFFFFF)
    with np.errstate(over="ignore"):
        for _ in range(10):
            p0 = _PH_M0 * c[0].astype(np.uint64)
            p1 = _PH_M1 * c[2].astype(np.uint64)
            hi0, lo0 = (p0 >> np.uint64(32)).astype(np.uint32), (p0 & mask).astype(np.uint32)
            hi1, lo1 = (p1 >> np.uint64(32)).astype(np.uint32), (p1 & mask).astype(np.uint32)
            c = [hi1 ^ c[1] ^ k0, lo1, hi0 ^ c[3] ^ k1, lo0]
            k0 = np.uint32((int(k0) + int(_PH_W0)) & 0xFFFFFFFF)
            k1 = np.uint32((int(k1) + int(_PH_W1)) & 0xFFFFFFFF)
    return np.stack(c, axis=-1)


def philox_normals(seed, index, N):
    """
    (re, im) of the Hermitian white noise H0 = 1/2 [W(k) + conj W(-k)] (box.py:174-176, 187) at global linear
    cell ``index`` (uint64 array, index = (a N + b) N + c) of an N^3 grid, drawn directly: one Box-Muller pair
    (n1, n2) per conjugate pair of modes.  Canonical cell j = min(index(k), index(-k)); Philox block j >> 1,
    key = seed, words (0,1) for even j and (2,3) for odd j;
    u1 = (x0 + 0.5) 2^-32, u2 = (x1 + 0.5) 2^-32, r = sqrt(-2 ln u1), n1 = r cos(2 pi u2), n2 = r sin(2 pi u2).
    k != -k: H0(canonical) = (n1 + i n2)/sqrt2, H0(other) = its conjugate; k == -k: H0 = n1 (real).
    Feeding (re, im) to ``hermitian_half_from_noise`` returns H0 itself (it is already Hermitian).
    """
    index = np.asarray(index, dtype=np.uint64)
    n = np.uint64(N)
    a, b, c = index // (n * n), (index // n) % n, index % n
    mirror = (((n - a) % n) * n + (n - b) % n) * n + (n - c) % n
    j = np.minimum(index, mirror)
    blk = j >> np.uint64(1)
    odd = (j & np.uint64(1)) != 0
    ctr = np.zeros(index.shape + (4,), dtype=np.uint32)
    ctr[..., 0] = (blk & np.uint64(0xFFFFFFFF)).astype(np.uint32)
    ctr[..., 1] = (blk >> np.uint64(32)).astype(np.uint32)
    key = np.array([seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF], dtype=np.uint32)
    x = philox4x32(ctr, key)
    u1 = (np.where(odd, x[..., 2], x[..., 0]).astype(np.float64) + 0.5) * 2.0 ** -32
    u2 = (np.where(odd, x[..., 3], x[..., 1]).astype(np.float64) + 0.5) * 2.0 ** -32
    r = np.sqrt(-2.0 * np.log(u1))
    n1, n2 = r * np.cos(TWO_PI * u2), r * np.sin(TWO_PI * u2)
    selfc = index == mirror
    re = np.where(selfc, n1, n1 * np.sqrt(0.5))
    im = np.where(selfc, 0.0, np.where(index < mirror, n2, -n2) * np.sqrt(0.5))
    return re, im
