"""
``CosmoBox`` -- drop-in for ``fastbox.box.CosmoBox`` (reference ``fastbox/box.py:23``)
whose field-generation hot path runs on a B200 through ``libfastbox_b200.so``.

Same constructor, method names, keyword names/defaults, return types (NumPy
float64 / complex128) and error behaviour as the reference.  What differs is
where the work happens:

* the N^3 arrays ``Kx, Ky, Kz, k`` (box.py:116-127) are lazy properties;
* ``delta_x`` lives on the device in float32, ``delta_k`` as a complex64 half
  spectrum (planes kx = 0..N/2); NumPy copies are made only when asked for;
* ``delta_k`` / ``velocity_k[i]`` / ``phi_k`` are ``DeviceSpectrum`` proxies that
  expand to the reference's full complex N^3 array on demand (``np.asarray``),
  and are recognised (no copy) when handed back to methods of this class.

There is no CPU fallback: compute methods raise ``FastBoxError`` without a GPU.
"""
import numpy as np

from . import _lib
from . import kspace as ks
from .cosmology import get_backend

# Speed of light (m/s)
C = 299792458.

# Default cosmology (same parameters as the reference, box.py:18-20)
default_cosmo = dict(Omega_c=0.25, Omega_b=0.05, h=0.7, n_s=0.95, sigma8=0.8,
                     transfer_function='eisenstein_hu')

ccl = get_backend()


class DeviceSpectrum(object):
    """
    Fourier-space field kept on the device as a Hermitian half spectrum, times an
    optional k-space factor (velocity component / potential) that is applied
    lazily.  Behaves like the reference's complex N^3 array via ``__array__``.
    """

    def __init__(self, box, buf, kind=_lib.KIND_PLAIN, scale=1.0):
        self.box, self.buf, self.kind, self.scale = box, buf, kind, scale
        self._full = None
        N = box.N
        self.shape = (N, N, N)
        self.dtype = np.dtype(np.complex128)
        self.ndim = 3
        self.size = N ** 3

    def half(self):
        """Host copy of the stored planes kx = 0..N/2 (complex64), before the lazy factor."""
        N = self.box.N
        return self.box._plan.download(self.buf, (N // 2 + 1, N, N), np.complex64)

    def __array__(self, dtype=None, copy=None):
        if self._full is None:
            N = self.box.N
            half = self.half().astype(np.complex128)
            if self.kind != _lib.KIND_PLAIN:
                half = half * self.box._kind_factor_half(self.kind) * self.scale
            h = N // 2 + 1
            full = np.empty((N, N, N), dtype=np.complex128)
            full[:h] = half
            m = np.conj(half[1:N - h + 1][::-1])
            m = np.concatenate([m[:, :1], m[:, :0:-1]], axis=1)
            m = np.concatenate([m[:, :, :1], m[:, :, :0:-1]], axis=2)
            full[h:] = m
            full.setflags(write=False)      # the device copy is the master: an in-place edit would not reach it
            self._full = full
        return self._full if dtype is None else self._full.astype(dtype)

    def real_space(self):
        """ifftn(self).real computed on the device (float64 NumPy result)."""
        box = self.box
        N = box.N
        out = box._plan.alloc(N ** 3 * 4)
        box._plan.spectrum_to_field(self.buf, out, kind=self.kind, scale=self.scale)
        return box._plan.download_f64(out, (N, N, N))

    # minimal ndarray-like behaviour used by reference-style user code
    def __getitem__(self, idx):
        return np.asarray(self)[idx]

    def __mul__(self, o):
        return np.asarray(self) * o

    __rmul__ = __mul__

    def conj(self):
        return np.conj(np.asarray(self))

    @property
    def real(self):
        return np.asarray(self).real

    @property
    def imag(self):
        return np.asarray(self).imag


class CosmoBox(object):

    def __init__(self, cosmo, box_scale=1e3, nsamp=32, redshift=0.,
                 line_freq=1420.405752, realise_now=True, device=0):
        """
        Same parameters as the reference (box.py:25-60); ``device`` selects the GPU.
        """
        if isinstance(cosmo, dict):
            cosmo = ccl.Cosmology(**cosmo)
        if not isinstance(cosmo, ccl.Cosmology):
            raise TypeError("`cosmo` must be a CCL Cosmology object or dict.")
        self.cosmo = cosmo
        if int(nsamp) != nsamp or nsamp < 8 or nsamp > 2048 or (int(nsamp) & (int(nsamp) - 1)):
            raise ValueError("nsamp=%r: the GPU path needs a power of two in [8, 2048] (the reference accepts any "
                             "size; see INTEGRATION.md, limits)" % (nsamp,))
        self.backend_name = ccl.name            # "pyccl" or "builtin-eh98" (cosmology.get_backend)
        self.N = nsamp
        self.redshift = redshift
        self.scale_factor = 1. / (1. + redshift)
        self.line_freq = line_freq
        self.device = device

        # grid coordinates (Mpc); the Fourier period is x[-1]-x[0] (box.py:79-89)
        if isinstance(box_scale, tuple):
            assert len(box_scale) == 3, "Must specify scale of x, y, z dimensions"
            sx, sy, sz = box_scale
            self.x = np.linspace(-0.5 * sx, 0.5 * sx, nsamp)
            self.y = np.linspace(-0.5 * sy, 0.5 * sy, nsamp)
            self.z = np.linspace(-0.5 * sz, 0.5 * sz, nsamp)
            self.Lx = self.x[-1] - self.x[0]
            self.Ly = self.y[-1] - self.y[0]
            self.Lz = self.z[-1] - self.z[0]
        else:
            self.x = self.y = self.z = np.linspace(-0.5 * box_scale, 0.5 * box_scale, nsamp)
            self.Lx = self.Ly = self.Lz = self.x[-1] - self.x[0]

        self.boxfactor = (self.N ** 6.) / (self.Lx * self.Ly * self.Lz)          # box.py:94
        self.kmin = 2. * np.pi / np.max([self.Lx, self.Ly, self.Lz])               # box.py:100
        self.kmax = 2. * np.pi * np.sqrt(3.) * self.N / np.min([self.Lx, self.Ly, self.Lz])   # box.py:101

        self.__plan = None
        self._kcache = {}
        self._pk_key = None
        self._bins_key = None
        self._d_delta_x = None          # device float32 field matching self.delta_x
        self._delta_x_host = None
        self._delta_x_sig = None

        if realise_now:
            self.realise_density()
            self.realise_velocity()
            self.realise_potential()

    # ------------------------------------------------------------------ plan
    @property
    def _plan(self):
        if self.__plan is None:
            self.__plan = _lib.Plan(self.N, self.Lx, self.Ly, self.Lz, self.device)
        return self.__plan

    @property
    def cubic(self):
        return self.Lx == self.Ly == self.Lz

    # -------------------------------------------- lazy N^3 arrays (box.py:116-127)
    def set_fft_sample_spacing(self):
        """Kept for API compatibility; the arrays are built on first access."""
        self._kcache.clear()

    def _mode_grid(self, axis):
        key = "K%d" % axis
        if key not in self._kcache:
            m = ks.mode_numbers(self.N).astype(np.float64)
            shape = [1, 1, 1]
            shape[axis] = self.N
            self._kcache[key] = np.ascontiguousarray(np.broadcast_to(m.reshape(shape), (self.N,) * 3))
        return self._kcache[key]

    @property
    def Kx(self):
        return self._mode_grid(0)

    @property
    def Ky(self):
        return self._mode_grid(1)

    @property
    def Kz(self):
        return self._mode_grid(2)

    @property
    def k(self):
        if "k" not in self._kcache:
            m = ks.mode_numbers(self.N).astype(np.float64)
            self._kcache["k"] = 2. * np.pi * np.sqrt((m[:, None, None] / self.Lx) ** 2.
                                                     + (m[None, :, None] / self.Ly) ** 2.
                                                     + (m[None, None, :] / self.Lz) ** 2.)
        return self._kcache["k"]

    def _kind_factor_half(self, kind):
        """Host copy of the lazy k-space factor on the half grid (box.py:254-274, 347)."""
        N, h = self.N, self.N // 2 + 1
        m = ks.mode_numbers(N).astype(np.float64)
        K = [m[:h, None, None], m[None, :, None], m[None, None, :]]
        L = [self.Lx, self.Ly, self.Lz]
        k2 = (2. * np.pi) ** 2 * ((K[0] / L[0]) ** 2. + (K[1] / L[1]) ** 2. + (K[2] / L[2]) ** 2.)
        with np.errstate(all="ignore"):
            if kind == _lib.KIND_POTENTIAL:
                f = np.nan_to_num(1.0 / k2, posinf=0.0)
                return f
            ax = kind - 1
            f = np.nan_to_num(1.j * K[ax] * (2. * np.pi / L[ax]) / k2)
        f = np.array(np.broadcast_to(f, (h, N, N)))
        idx = [slice(None)] * 3
        idx[ax] = N // 2
        f[tuple(idx)] = 0.                                           # box.py:268-274
        return f

    # ------------------------------------------------------------------ tables
    def _load_power(self, linear, scale_factor):
        key = (bool(linear), float(scale_factor))
        if self._pk_key == key:
            return
        fn = ccl.linear_matter_power if linear else ccl.nonlin_matter_power
        pk = lambda kk: fn(self.cosmo, k=kk, a=scale_factor)                     # box.py:162-165
        with np.errstate(all="ignore"):
            mode, tab, l0, dl = ks.choose_sqrt_pk_table(pk, self.N, self.Lx, self.Ly, self.Lz, self.boxfactor)
        self._plan.set_sqrt_pk(tab, mode, l0, dl)
        self._pk_key = key

    def _load_isotropic(self, fn_of_k):
        """Install an isotropic multiplier W(|k|) in the LUT slot (used by smooth_field)."""
        self._pk_key = None
        with np.errstate(all="ignore"):
            if self.cubic:
                n = np.arange(3 * (self.N // 2) ** 2 + 1, dtype=np.float64)
                kk = 2. * np.pi * np.sqrt(n) / self.Lx
                self._plan.set_sqrt_pk(np.nan_to_num(fn_of_k(kk)).astype(np.float32), 1)
            else:
                s, l0, dl = ks.log_table_nodes(self.N, self.Lx, self.Ly, self.Lz, 1 << 16)
                tab = np.nan_to_num(fn_of_k(2. * np.pi * np.sqrt(s))).astype(np.float32)
                self._plan.set_sqrt_pk(tab, 2, l0, dl)

    def _load_bins(self, nbins, kbins):
        bins = ks.pk_bin_edges(self.kmin, self.kmax, nbins, kbins)
        key = bins.tobytes()
        if self._bins_key != key:
            if bins.size > _lib.MAX_EDGES:
                raise ValueError("at most %d k-bin edges are supported" % _lib.MAX_EDGES)
            self._plan.set_pk_bins(ks.bin_thresholds(bins))
            self._bins_key = key
        return bins

    # ------------------------------------------------------------ device helpers
    def _to_device_field(self, arr):
        """float32 device buffer for a real N^3 array; reuses the resident copy of delta_x."""
        if isinstance(arr, _lib.DeviceBuffer):
            return arr
        if arr is self._delta_x_host and self._d_delta_x is not None:
            # zero-copy reuse only while the host array still holds what was downloaded: an in-place edit
            # (box.delta_x *= b, box.delta_x[mask] = 0) must reach the device like it reaches the reference
            if self._field_signature(arr) == self._delta_x_sig:
                return self._d_delta_x
            self._d_delta_x = None
        a = np.asarray(arr)
        if np.iscomplexobj(a):
            a = a.real
        assert a.shape == (self.N,) * 3, "field must have shape (N, N, N)"
        return self._plan.upload_f32(a)

    @staticmethod
    def _field_signature(a):
        """Cheap content check of a host field (two BLAS-speed reductions, ~10x cheaper than re-uploading)."""
        flat = a.reshape(-1)
        return (float(flat.sum()), float(np.dot(flat, flat)), float(flat[0]), float(flat[-1]))

    def _spectrum_arg(self, delta_x, delta_k):
        """Resolve the (delta_x, delta_k, self.delta_k) convention of box.py:241-248."""
        if delta_x is not None and delta_k is not None:
            raise ValueError("delta_x and delta_k specified; can only specify one")
        N = self.N
        if delta_x is not None:
            spec = self._plan.alloc((N // 2 + 1) * N * N * 8)
            self._plan.field_to_spectrum(self._to_device_field(delta_x), spec_out=spec)
            return DeviceSpectrum(self, spec)
        if delta_k is None:
            delta_k = self.delta_k
        if isinstance(delta_k, DeviceSpectrum):
            return delta_k
        # user-supplied full complex cube (the spectrum of a real field in every reference call site,
        # box.py:246,337): its Hermitian projection 1/2 [A(k) + conj A(-k)], planes kx = 0..N/2
        full = np.asarray(delta_k)
        if full.shape != (N, N, N):
            raise ValueError("delta_k must have shape (N, N, N)")
        h = N // 2 + 1
        mir = np.concatenate([full[:1], full[:0:-1]], axis=0)[:h]                  # A(-k) on kx = 0..N/2
        mir = np.concatenate([mir[:, :1], mir[:, :0:-1]], axis=1)
        mir = np.concatenate([mir[:, :, :1], mir[:, :, :0:-1]], axis=2)
        a = np.ascontiguousarray((0.5 * (full[:h] + np.conj(mir))).astype(np.complex64))
        return DeviceSpectrum(self, self._plan.upload(a))

    # ------------------------------------------------------------- realisations
    def realise_density(self, linear=False, redshift=None, inplace=True, seed=None):
        """
        Gaussian realisation of P(k) (box.py:130-194).  White noise is drawn from
        NumPy's global generator in the reference's order (re, then im) unless
        ``seed`` is given, in which case counter-based Philox noise is generated
        on the device (different stream, no host RNG cost).
        Returns ``delta_x`` (float64 NumPy array).
        """
        if redshift is None:
            redshift = self.redshift
        scale_factor = 1. / (1. + redshift)
        N = self.N
        self._load_power(linear, scale_factor)
        plan = self._plan
        if seed is None:
            re = np.random.normal(0.0, 1.0, (N, N, N))                            # box.py:174
            im = np.random.normal(0.0, 1.0, (N, N, N))                            # box.py:175
            d_re, d_im = plan.upload_f32(re), plan.upload_f32(im)
        else:
            d_re = d_im = None
        field = plan.alloc(N ** 3 * 4)
        spec = plan.alloc((N // 2 + 1) * N * N * 8)
        plan.realise(d_re, d_im, seed=0 if seed is None else seed, flags=_lib.F_SQRTPK, field_out=field,
                     spec_out=spec)
        delta_x = plan.download_f64(field, (N, N, N))
        if inplace:
            if redshift != self.redshift:
                print("Warning: Storing density field into self.delta_x with a "
                      "different redshift than self.redshift.")
            self.delta_x = delta_x
            self._delta_x_host = delta_x
            self._delta_x_sig = self._field_signature(delta_x)
            self._d_delta_x = field
            self.delta_k = DeviceSpectrum(self, spec)
        return delta_x

    def realise_velocity(self, delta_x=None, delta_k=None, redshift=None, inplace=True):
        """
        v(k) = i [f H a] delta_k k_vec / k^2 (box.py:197-290).  Returns a 3-tuple of
        lazy ``DeviceSpectrum`` objects; ``np.fft.ifftn(v[i])`` works as in the
        reference, ``v[i].real_space()`` does the same transform on the GPU.
        """
        if redshift is None:
            redshift = self.redshift
        scale_factor = 1. / (1. + redshift)
        spec = self._spectrum_arg(delta_x, delta_k)
        fac = 100. * self.cosmo['h'] * ccl.h_over_h0(self.cosmo, a=scale_factor) \
            * ccl.growth_rate(self.cosmo, a=scale_factor) * scale_factor         # box.py:280-281
        if self.N % 2 != 0:
            raise NameError("name 'mx' is not defined")                           # box.py:268-274 for odd N
        velocity_k = tuple(DeviceSpectrum(self, spec.buf, kind=kind, scale=fac)
                           for kind in (_lib.KIND_VEL_X, _lib.KIND_VEL_Y, _lib.KIND_VEL_Z))
        if inplace:
            self.velocity_k = velocity_k
        return velocity_k

    def realise_potential(self, delta_x=None, delta_k=None, redshift=None, inplace=True):
        """phi_k = delta_k / k^2 with the monopole zeroed (box.py:293-353)."""
        spec = self._spectrum_arg(delta_x, delta_k)
        phi_k = DeviceSpectrum(self, spec.buf, kind=_lib.KIND_POTENTIAL, scale=1.0)
        if inplace:
            self.phi_k = phi_k
        return phi_k

    def velocity_field(self, axis=2, delta_x=None, delta_k=None, redshift=None):
        """Real-space velocity component (km/s) = ifftn(velocity_k[axis]).real, on the GPU."""
        return self.realise_velocity(delta_x, delta_k, redshift, inplace=False)[axis].real_space()

    # ------------------------------------------------- filters (box.py:356-381)
    def apply_transfer_fn(self, field_k, transfer_fn):
        """
        ``ifftn(field_k * transfer_fn(k_perp, k_par))`` -- complex128 array like the
        reference.  ``field_k`` may be a ``DeviceSpectrum`` (e.g. ``box.delta_k``,
        no copy) or any complex (N,N,N) array.
        """
        N = self.N
        ft = ks.filter_tables(transfer_fn, N, self.Lx, self.Ly, self.Lz)
        plan = self._plan
        plan.set_filter(ft.tperp, ft.tpar, ft.tdense)
        out = plan.alloc(N ** 3 * 4)
        if isinstance(field_k, DeviceSpectrum) and field_k.kind == _lib.KIND_PLAIN and ft.even:
            plan.spectrum_to_field(field_k.buf, out, flags=_lib.F_FILTER)
            return plan.download_f64(out, (N, N, N)).astype(np.complex128)
        cube = plan.upload(np.ascontiguousarray(np.asarray(field_k), dtype=np.complex64))
        plan.cube_to_field(cube, out, flags=_lib.F_FILTER, part=0)
        res = plan.download_f64(out, (N, N, N)).astype(np.complex128)
        plan.cube_to_field(cube, out, flags=_lib.F_FILTER, part=1)
        res.imag = plan.download_f64(out, (N, N, N))
        return res

    def window(self, k, R):
        """Top-hat window squared (box.py:595-613)."""
        return self.window1(k, R) ** 2.

    def window1(self, k, R):
        """Top-hat window (box.py:615-633)."""
        x = k * R
        return (3. / x ** 3.) * (np.sin(x) - x * np.cos(x))

    def smooth_field(self, field_k, R):
        """Top-hat smoothing of radius R Mpc/h, complex result (box.py:635-655)."""
        N = self.N
        plan = self._plan
        Rm = R / self.cosmo['h']
        self._load_isotropic(lambda kk: self.window1(kk, Rm))
        out = plan.alloc(N ** 3 * 4)
        if isinstance(field_k, DeviceSpectrum) and field_k.kind == _lib.KIND_PLAIN:
            plan.spectrum_to_field(field_k.buf, out, flags=_lib.F_SQRTPK)
            return plan.download_f64(out, (N, N, N)).astype(np.complex128)
        cube = plan.upload(np.ascontiguousarray(np.asarray(field_k), dtype=np.complex64))
        plan.cube_to_field(cube, out, flags=_lib.F_SQRTPK, part=0)
        res = plan.download_f64(out, (N, N, N)).astype(np.complex128)
        plan.cube_to_field(cube, out, flags=_lib.F_SQRTPK, part=1)
        res.imag = plan.download_f64(out, (N, N, N))
        return res

    # --------------------------------------- redshift space (box.py:384-438)
    def redshift_space_density(self, delta_x=None, velocity_z=None, sigma_nl=0., method='linear'):
        """
        Remap the density along z by the peculiar velocity (box.py:384-438).  ``method`` is the keyword the
        reference forwards to ``scipy.interpolate.griddata`` (box.py:433-437): 'linear' (default) and 'nearest'
        run on the device.  'cubic' is rejected: scipy's 1-D cubic is a global not-a-knot spline through the
        sorted samples and fails on the duplicate sample that the periodic wrap of an unmoved end point
        creates, so it has no well-defined reference result to match -- and this package has no CPU path.
        """
        if method not in ('linear', 'nearest'):
            raise NotImplementedError("redshift_space_density(method=%r): 'linear' and 'nearest' are implemented "
                                      "on the GPU path (there is no CPU fallback)" % (method,))
        N = self.N
        Hz = 100. * self.cosmo['h'] * ccl.h_over_h0(self.cosmo, self.scale_factor)  # box.py:406
        plan = self._plan
        d = self._to_device_field(delta_x)
        v = self._to_device_field(velocity_z)
        vnl = None
        if sigma_nl > 0.:
            # the reference draws N normals per line of sight in (i, j) order (box.py:412-418),
            # i.e. one C-ordered (N, N, N) draw from the global generator
            vnl = plan.upload_f32(sigma_nl * np.random.normal(0., 1., (N, N, N)))
        out = plan.alloc(N ** 3 * 4)
        plan.rsd_remap(d, v, vnl, self.z, Hz, out, method=method)
        return plan.download_f64(out, (N, N, N))

    # ------------------------------------------------- log-normal (box.py:441-460)
    def lognormal(self, delta_x):
        N = self.N
        plan = self._plan
        d = self._to_device_field(delta_x)
        out = plan.alloc(N ** 3 * 4)
        total = plan.exp_sum(d, out, N ** 3, 1.0)                                 # box.py:457
        plan.affine(out, N ** 3, 1.0 / (total / N ** 3), -1.0)                    # box.py:458-459
        return plan.download_f64(out, (N, N, N))

    # --------------------------------------------- power spectrum (box.py:696-768)
    def binned_power_spectrum(self, delta_x=None, delta_k=None, nbins=20, kbins=None):
        if delta_x is not None and delta_k is not None:
            raise ValueError("delta_x and delta_k specified; can only specify one")
        bins = self._load_bins(nbins, kbins)
        plan = self._plan
        N = self.N
        if delta_x is not None:
            a = np.asarray(delta_x) if not isinstance(delta_x, _lib.DeviceBuffer) else None
            if a is not None and np.iscomplexobj(a) and np.any(a.imag != 0):
                # complex "field": P(k) of the full complex cube, as fftn would give (box.py:736)
                return self.binned_power_spectrum(delta_k=np.fft.fftn(a), nbins=nbins, kbins=kbins)
            res = plan.field_to_spectrum(self._to_device_field(delta_x), want_pk=True)
        else:
            if delta_k is None:
                delta_k = self.delta_k
            if isinstance(delta_k, DeviceSpectrum) and delta_k.kind == _lib.KIND_PLAIN:
                res = plan.pk_from_spectrum(delta_k.buf)
            else:
                cube = np.ascontiguousarray(np.asarray(delta_k), dtype=np.complex64)
                res = plan.pk_from_spectrum(cube, full_cube=True)
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            return ks.moments_to_spectrum(bins, res["count"], res["sum1"], res["sum2"])

    def binned_power_spectrum_2d(self, delta_x=None, delta_k=None, nbins=(20, 20), kperp_bins=None, kpar_bins=None):
        """
        EXTENSION (the reference delegates 2-D spectra to nbodykit's FFTPower): P(k_perp, |k_par|) with the line of
        sight along z, estimator and conventions of ``binned_power_spectrum`` (np.digitize bins, the first edge
        only bounds the first bin from below, population stddev / sqrt(n), NaN for empty bins).  Default edges:
        ``nbins`` log-spaced edges per axis between the fundamental and the Nyquist wavenumber of that axis.
        Returns (kperp_centres, kpar_centres, pk2d, stddev2d), pk2d of shape (len(kperp_bins)-1, len(kpar_bins)-1).
        """
        if delta_x is not None and delta_k is not None:
            raise ValueError("delta_x and delta_k specified; can only specify one")
        N = self.N
        if kperp_bins is None:
            kperp_bins = np.logspace(np.log10(2. * np.pi / max(self.Lx, self.Ly)),
                                     np.log10(np.sqrt(2.) * np.pi * N / min(self.Lx, self.Ly)), nbins[0])
        if kpar_bins is None:
            kpar_bins = np.logspace(np.log10(2. * np.pi / self.Lz), np.log10(np.pi * N / self.Lz), nbins[1])
        kperp_bins = np.asarray(kperp_bins, dtype=np.float64)
        kpar_bins = np.asarray(kpar_bins, dtype=np.float64)
        if max(kperp_bins.size, kpar_bins.size) > 64:
            raise ValueError("at most 64 edges per axis are supported")
        m = ks.mode_numbers(N).astype(np.float64)
        ipar = np.digitize(np.abs(2. * np.pi * m / self.Lz), kpar_bins)
        thr = ks.bin_thresholds(kperp_bins)
        plan = self._plan
        if delta_x is not None:
            spec = self._spectrum_arg(delta_x, None)
        else:
            spec = self._spectrum_arg(None, delta_k)
        if spec.kind != _lib.KIND_PLAIN:
            raise ValueError("binned_power_spectrum_2d needs a density-like spectrum")
        res = plan.pk2d_from_spectrum(spec.buf, thr, ipar, kpar_bins.size)
        cnt = res["count"].astype(np.float64)
        with np.errstate(all="ignore"):
            mean = res["sum1"] / cnt
            err = np.sqrt(np.maximum(res["sum2"] / cnt - mean * mean, 0.0)) / np.sqrt(cnt)
        sl = (slice(1, kperp_bins.size), slice(1, kpar_bins.size))
        return ks.bin_centres(kperp_bins)[1:], ks.bin_centres(kpar_bins)[1:], mean[sl], err[sl]

    def correlation_function(self, delta_x=None, delta_b=None, dr=2., rmin=20., rmax=200., rbins=None):
        """
        EXTENSION (nbodykit ``FFTCorr(first=mesh, mode='1d', dr=2., rmin=20., rmax=200.)`` in
        examples/example_endtoend.py:128-151): the correlation function xi(r) = < d(x) d(x + r) > of a real field
        (cross-correlation with ``delta_b`` if given), from xi = ifftn(|fftn(d)|^2) / N^3 on the device and a
        radial average over the periodic lags (cell size L/N) in the bins ``rbins`` (default: edges
        rmin, rmin + dr, ... up to rmax).  Returns (r_centres, xi, stddev / sqrt(n)); empty bins are NaN.
        """
        if rbins is None:
            rbins = np.arange(rmin, rmax + 0.5 * dr, dr)
        rbins = np.asarray(rbins, dtype=np.float64)
        if rbins.size < 2 or rbins.size > _lib.MAX_EDGES:
            raise ValueError("need between 2 and %d r-bin edges" % _lib.MAX_EDGES)
        if delta_x is None:
            delta_x = self.delta_x
        a = self._to_device_field(delta_x)
        b = None if delta_b is None else self._to_device_field(delta_b)
        res = self._plan.correlation_function(a, rbins, field_b=b)
        nb = rbins.size
        cnt = res["count"][:nb].astype(np.float64)
        with np.errstate(all="ignore"):
            mean = res["sum1"][:nb] / cnt
            err = np.sqrt(np.maximum(res["sum2"][:nb] / cnt - mean * mean, 0.0)) / np.sqrt(cnt)
        return 0.5 * (rbins[1:] + rbins[:-1]), mean[1:], err[1:]

    def power_multipoles(self, delta_x, nbins=20, kbins=None):
        """
        EXTENSION (the reference delegates this to nbodykit, example_box.py:48-52):
        P_l(k), l = 0, 2, 4, with the line of sight along z and the same bins as
        ``binned_power_spectrum``.
        """
        bins = self._load_bins(nbins, kbins)
        res = self._plan.field_to_spectrum(self._to_device_field(delta_x), want_pk=True, poles=True)
        nb = bins.size
        cnt = res["count"][:nb].astype(np.float64)
        with np.errstate(all="ignore"):
            p0 = res["sum1"][:nb] / cnt
            p2 = 5. * res["sum_l2"][:nb] / cnt
            p4 = 9. * res["sum_l4"][:nb] / cnt
        return ks.bin_centres(bins)[1:], p0[1:], p2[1:], p4[1:]

    def cross_power_spectrum(self, delta_a, delta_b, nbins=20, kbins=None):
        """EXTENSION (nbodykit FFTPower(first, second) in example_halos.py:52): Re[a b*] binned."""
        bins = self._load_bins(nbins, kbins)
        N = self.N
        plan = self._plan
        spec_b = plan.alloc((N // 2 + 1) * N * N * 8)
        plan.field_to_spectrum(self._to_device_field(delta_b), spec_out=spec_b)
        res = plan.field_to_spectrum(self._to_device_field(delta_a), cross=spec_b, want_pk=True)
        return ks.moments_to_spectrum(bins, res["count"], res["sum1"], res["sum2"])

    def sigmaR(self, R):
        """RMS in top-hat spheres of radius R Mpc/h from the binned P(k) (box.py:657-683)."""
        import scipy.integrate
        k, pk, stddev = self.binned_power_spectrum()
        good = ~np.isnan(pk)
        pk, k = pk[good], k[good]
        y = k ** 2. * pk * self.window(k, R / self.cosmo['h'])
        simps = getattr(scipy.integrate, "simps", None) or scipy.integrate.simpson
        I = simps(y, x=k)
        return np.sqrt(I / (2. * np.pi ** 2.))

    def sigma8(self):
        return self.sigmaR(8.0)

    def theoretical_power_spectrum(self):
        """box.py:770-782."""
        k = np.logspace(-3.5, 1., int(1e3))
        pk = ccl.nonlin_matter_power(self.cosmo, k=k, a=self.scale_factor)
        return k, pk

    # --------------------------------------------- coordinates (box.py:789-864)
    def freq_array(self, redshift=None):
        if redshift is None:
            redshift = self.redshift
        a = 1. / (1. + redshift)
        freq_centre = a * self.line_freq
        dx = self.Lz / self.N
        Hz = 100. * self.cosmo['h'] * ccl.h_over_h0(self.cosmo, a)
        df = dx * self.line_freq * (a ** 2. * Hz) / (C / 1e3)
        freqs = freq_centre + df * (np.arange(self.N) - 0.5 * (self.N - 1.))
        return freqs[::-1]

    def pixel_array(self, redshift=None):
        if redshift is None:
            redshift = self.redshift
        scale_factor = 1. / (1. + redshift)
        r = ccl.comoving_angular_distance(self.cosmo, scale_factor)
        x_px = self.x[1] - self.x[0]
        y_px = self.y[1] - self.y[0]
        ang_x = (180. / np.pi) * (x_px / r)
        ang_y = (180. / np.pi) * (y_px / r)
        grid = np.arange(self.N) - 0.5 * (self.N - 1.)
        return ang_x * grid, ang_y * grid

    # ------------------------------------------- consistency checks (box.py:871-948)
    def test_parseval(self):
        """sum(delta_x^2) N^3 vs sum |delta_k|^2, both reduced on the device (box.py:931-948)."""
        N = self.N
        _, sq = self._plan.field_moments(self._to_device_field(self.delta_x), N ** 3)
        s1 = sq * N ** 3.
        self._load_bins(2, None)
        res = self._plan.pk_from_spectrum(self.delta_k.buf)
        s2 = float(np.sum(res["sum1"])) * self.boxfactor
        print("Parseval test:", s1 / s2, "(should be 1.0)")
        return s1, s2

    def test_sampling_error(self):
        """Report sigma8 from the box vs theory (box.py:871-928)."""
        import scipy.integrate
        simps = getattr(scipy.integrate, "simps", None) or scipy.integrate.simpson
        s8_real = self.sigma8()
        _k = np.linspace(self.kmin, self.kmax, int(5e3))
        _pk = ccl.nonlin_matter_power(self.cosmo, k=_k, a=self.scale_factor)
        _y = np.nan_to_num(_k ** 2. * _pk * self.window(_k, 8.0 / self.cosmo['h']))
        s8_th_win = np.sqrt(simps(_y, x=_k) / (2. * np.pi ** 2.))
        _k2 = np.logspace(-5, 2, int(5e4))
        _pk2 = ccl.nonlin_matter_power(self.cosmo, k=_k2, a=self.scale_factor)
        _y2 = np.nan_to_num(_k2 ** 2. * _pk2 * self.window(_k2, 8.0 / self.cosmo['h']))
        s8_th_full = np.sqrt(simps(_y2, x=_k2) / (2. * np.pi ** 2.))
        s8_realspace = np.std(self.smooth_field(self.delta_k, 8.0))
        s20_realspace = np.std(self.smooth_field(self.delta_k, 20.0))
        s20_real = self.sigmaR(20.)
        print("")
        print("sigma8 (real.): \t", s8_real)
        print("sigma8 (th.win.):\t", s8_th_win)
        print("sigma8 (th.full):\t", s8_th_full)
        print("sigma8 (realsp.):\t", s8_realspace)
        print("ratio =", 1. / (s8_real / s8_realspace))
        print("")
        print("sigma20 (real.): \t", s20_real)
        print("sigma20 (realsp.):\t", s20_realspace)
        print("ratio =", 1. / (s20_real / s20_realspace))
        print("var(delta) =", np.std(self.delta_x))
