"""
ctypes binding of ``libfastbox_b200.so`` (C ABI declared in ``include/fastbox_b200.h``).

There is deliberately no CPU fallback: if the shared library is missing or no
CUDA device is present, every compute entry point raises ``FastBoxError``.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("FB_LIB") or os.path.join(_HERE, "libfastbox_b200.so")     # FB_LIB: A/B builds

# flags / kinds (mirror include/fastbox_b200.h)
KIND_PLAIN, KIND_VEL_X, KIND_VEL_Y, KIND_VEL_Z, KIND_POTENTIAL = 0, 1, 2, 3, 4
F_SQRTPK, F_FILTER, F_EXP, F_ANTIHERM, F_PK, F_POLES = 1, 2, 4, 8, 16, 32
RSD_METHODS = {"linear": 0, "nearest": 1}      # FB_RSD_LINEAR / FB_RSD_NEAREST
MAX_EDGES = 128


class FastBoxError(RuntimeError):
    pass


class PkResult(C.Structure):
    _fields_ = [("count", C.POINTER(C.c_uint64)), ("sum1", C.POINTER(C.c_double)),
                ("sum2", C.POINTER(C.c_double)), ("sum_l2", C.POINTER(C.c_double)),
                ("sum_l4", C.POINTER(C.c_double))]


_vp, _i, _d, _f, _sz, _u64, _l = C.c_void_p, C.c_int, C.c_double, C.c_float, C.c_size_t, C.c_uint64, C.c_long

# name -> (restype, argtypes); every symbol of include/fastbox_b200.h
SIGNATURES = {
    "fb_plan_create": (_i, [C.POINTER(_vp), _i, _d, _d, _d, _i]),
    "fb_plan_destroy": (_i, [_vp]),
    "fb_sync": (_i, [_vp]),
    "fb_last_error": (C.c_char_p, []),
    "fb_version": (C.c_char_p, []),
    "fb_launch_count": (_u64, []),
    "fb_plan_set_slab": (_i, [_vp, _i, _i, _i, _i]),
    "fb_dev_alloc": (_i, [C.POINTER(_vp), _sz]),
    "fb_dev_alloc_on": (_i, [_i, C.POINTER(_vp), _sz]),
    "fb_dev_free": (_i, [_vp]),
    "fb_host_alloc": (_i, [C.POINTER(_vp), _sz]),
    "fb_host_free": (_i, [_vp]),
    "fb_copy": (_i, [_vp, _vp, _vp, _sz]),
    "fb_convert_f64_to_f32": (_i, [_vp, _vp, _vp, _sz]),
    "fb_convert_f32_to_f64": (_i, [_vp, _vp, _vp, _sz]),
    "fb_device_info": (_i, [_i, C.c_char_p, _i, C.POINTER(_i), C.POINTER(_sz)]),
    "fb_set_sqrt_pk": (_i, [_vp, _vp, _l, _i, _d, _d]),
    "fb_set_filter": (_i, [_vp, _vp, _vp, _vp]),
    "fb_set_pk_bins": (_i, [_vp, _vp, _i]),
    "fb_realise": (_i, [_vp, _vp, _vp, _u64, _i, _f, _vp, _vp, C.POINTER(PkResult), C.POINTER(_d)]),
    "fb_spectrum_to_field": (_i, [_vp, _vp, _i, _i, _f, _vp, C.POINTER(_d)]),
    "fb_cube_to_field": (_i, [_vp, _vp, _i, _i, _f, _vp]),
    "fb_field_to_spectrum": (_i, [_vp, _vp, _vp, _vp, _i, C.POINTER(PkResult)]),
    "fb_pk_from_spectrum": (_i, [_vp, _vp, _vp, _i, _i, C.POINTER(PkResult)]),
    "fb_pk2d_from_spectrum": (_i, [_vp, _vp, _vp, _i, _vp, _i, _vp, _i, _vp, _vp, _vp]),
    "fb_correlation_function": (_i, [_vp, _vp, _vp, _vp, _i, _vp, _vp, _vp, _vp]),
    "fb_affine": (_i, [_vp, _vp, _sz, _f, _f]),
    "fb_exp_sum": (_i, [_vp, _vp, _vp, _sz, _f, C.POINTER(_d)]),
    "fb_field_moments": (_i, [_vp, _vp, _sz, C.POINTER(_d), C.POINTER(_d)]),
    "fb_rsd_remap": (_i, [_vp, _vp, _vp, _vp, _vp, _d, _vp]),
    "fb_rsd_remap_method": (_i, [_vp, _vp, _vp, _vp, _vp, _d, _i, _vp]),
    "fb_beam_set": (_i, [_vp, _vp]),
    "fb_beam_convolve": (_i, [_vp, _vp, _vp, _vp]),
    "fb_halo_counts": (_i, [_vp, _vp, _vp, _i, _vp, _i, _i, _d, _vp, _vp, _vp]),
    "fb_counts_to_field": (_i, [_vp, _vp, _sz, _f, _f, _vp]),
    "fb_halo_catalogue": (_i, [_vp, _vp, _vp, _vp, _u64, C.POINTER(_u64)]),
    "fb_fg_cube": (_i, [_vp, _vp, _vp, _i, _vp, _vp, _i]),
    "fb_radiometer_noise": (_i, [_vp, _vp, _vp, _u64, _vp, _i]),
    "fb_mean_spectrum_filter": (_i, [_vp, _vp, _vp, _vp]),
    "fb_pca_covariance": (_i, [_vp, _vp, _vp, _vp]),
    "fb_pca_project": (_i, [_vp, _vp, _vp, _vp, _i, _vp, _vp]),
    "fb_bench_fp64": (_i, [_vp, _i, C.POINTER(_d)]),
    "fb_fft_pass_c2c": (_i, [_vp, _vp, _i, _i, _i]),
    "fb_fft_pass_x_c2r": (_i, [_vp, _vp, _vp, _l, _i, _f, C.POINTER(_d)]),
    "fb_fft_pass_x_c2r_gather": (_i, [_vp, _vp, _vp, _vp, _l, _i, _f, C.POINTER(_d)]),
    "fb_fft_pass_x_r2c": (_i, [_vp, _vp, _vp, _l]),
    "fb_realise_local_kspace": (_i, [_vp, _u64, _i, _vp, _vp, _i, C.POINTER(PkResult)]),
    "fb_forward_local_kspace": (_i, [_vp, _vp, _vp, _i, _vp, _i, C.POINTER(PkResult)]),
    "fb_dist_init": (_i, [_vp, _i, _i, _i]),
    "fb_dist_get_handle": (_i, [_vp, _vp]),
    "fb_dist_connect": (_i, [_vp, _vp]),
    "fb_dist_info": (_i, [_vp, C.POINTER(_i), C.POINTER(_i), C.POINTER(_i), C.POINTER(_i), C.POINTER(_i),
                          C.POINTER(_i), C.POINTER(_sz)]),
    "fb_dist_barrier": (_i, [_vp]),
    "fb_dist_realise": (_i, [_vp, _u64, _i, _f, _i, _i, _vp, C.POINTER(PkResult), C.POINTER(_d)]),
    "fb_dist_bench_exchange": (_i, [_vp, _i, C.POINTER(_f)]),
    "fb_dist_set_option": (_i, [_vp, C.c_char_p, _i]),
    "fb_dist_power_spectrum": (_i, [_vp, _vp, _i, _i, C.POINTER(PkResult)]),
    "fb_bench_strided_copy": (_i, [_vp, _sz, _i, _i, C.POINTER(_d)]),
    "fb_last_timings": (_i, [_vp, C.POINTER(_f), _i]),
    "fb_timer_start": (_i, [_vp]),
    "fb_timer_stop": (_i, [_vp, C.POINTER(_f)]),
}

_lib = None


def load():
    """Load the shared library (once).  Raises FastBoxError if it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise FastBoxError("%s not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(there is no CPU fallback)" % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        raise FastBoxError("libfastbox_b200 error %d: %s" % (rc, load().fb_last_error().decode()))


def _ptr(x):
    """void* of a numpy array / DeviceBuffer / torch tensor / None."""
    if x is None:
        return None
    if isinstance(x, DeviceBuffer):
        return x.ptr
    if isinstance(x, np.ndarray):
        assert x.flags["C_CONTIGUOUS"], "array must be C contiguous"
        return x.ctypes.data
    if hasattr(x, "data_ptr"):          # torch tensor (CUDA or CPU), used in place
        return x.data_ptr()
    if isinstance(x, int):
        return x
    raise TypeError("cannot take a pointer of %r" % type(x))


class DeviceBuffer(object):
    """Owning handle of device memory allocated through the C ABI."""

    def __init__(self, nbytes, device=None):
        p = C.c_void_p()
        if device is None:
            check(load().fb_dev_alloc(C.byref(p), max(int(nbytes), 16)))
        else:
            check(load().fb_dev_alloc_on(int(device), C.byref(p), max(int(nbytes), 16)))
        self.ptr = p.value
        self.nbytes = int(nbytes)
        self.device = device

    def free(self):
        if self.ptr:
            load().fb_dev_free(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Plan(object):
    """Thin object wrapper over ``fb_plan`` (one per box / device)."""

    def __init__(self, N, Lx, Ly, Lz, device=0):
        self.lib = load()
        h = C.c_void_p()
        check(self.lib.fb_plan_create(C.byref(h), int(N), float(Lx), float(Ly), float(Lz), int(device)))
        self.h = h
        self.N = int(N)
        self.L = (float(Lx), float(Ly), float(Lz))
        self.device = int(device)
        self.nedges = 0
        self._keep = []

    def close(self):
        for hp in getattr(self, "_keep", []):
            try:
                self.lib.fb_host_free(hp)
            except Exception:
                pass
        self._keep = []
        if getattr(self, "h", None):
            self.lib.fb_plan_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- memory -------------------------------------------------------------
    def alloc(self, nbytes):
        return DeviceBuffer(nbytes, self.device)

    def upload(self, arr, dtype=None):
        a = np.ascontiguousarray(arr, dtype=dtype)
        buf = DeviceBuffer(a.nbytes, self.device)
        check(self.lib.fb_copy(self.h, buf.ptr, a.ctypes.data, a.nbytes))
        return buf

    def upload_f32(self, arr):
        """Upload any real array as float32 (float64 input is converted on the device)."""
        a = np.ascontiguousarray(arr)
        if a.dtype == np.float32:
            return self.upload(a)
        a = np.ascontiguousarray(a, dtype=np.float64)
        buf = DeviceBuffer(a.size * 4, self.device)
        check(self.lib.fb_convert_f64_to_f32(self.h, a.ctypes.data, buf.ptr, a.size))
        return buf

    def download(self, buf, shape, dtype):
        out = np.empty(shape, dtype=dtype)
        check(self.lib.fb_copy(self.h, out.ctypes.data, _ptr(buf), out.nbytes))
        return out

    def download_f64(self, buf, shape):
        out = np.empty(shape, dtype=np.float64)
        check(self.lib.fb_convert_f32_to_f64(self.h, _ptr(buf), out.ctypes.data, out.size))
        return out

    def sync(self):
        check(self.lib.fb_sync(self.h))

    # -- tables -------------------------------------------------------------
    def set_sqrt_pk(self, table, mode, log2s0=0.0, dlog2s=0.0):
        t = np.ascontiguousarray(table, dtype=np.float32)
        check(self.lib.fb_set_sqrt_pk(self.h, t.ctypes.data, t.size, int(mode), float(log2s0), float(dlog2s)))

    def set_filter(self, tperp=None, tpar=None, tdense=None):
        arrs = [None if t is None else np.ascontiguousarray(t, dtype=np.float32) for t in (tperp, tpar, tdense)]
        check(self.lib.fb_set_filter(self.h, *[None if a is None else a.ctypes.data for a in arrs]))

    def set_pk_bins(self, thresholds):
        t = np.ascontiguousarray(thresholds, dtype=np.float64)
        check(self.lib.fb_set_pk_bins(self.h, t.ctypes.data, t.size))
        self.nedges = t.size

    def set_slab(self, a0, na, y0, ny):
        check(self.lib.fb_plan_set_slab(self.h, a0, na, y0, ny))

    # -- helpers ------------------------------------------------------------
    def _pk_struct(self, poles=False):
        n = self.nedges + 1
        res = dict(count=np.zeros(n, np.uint64), sum1=np.zeros(n), sum2=np.zeros(n))
        st = PkResult()
        st.count = res["count"].ctypes.data_as(C.POINTER(C.c_uint64))
        st.sum1 = res["sum1"].ctypes.data_as(C.POINTER(C.c_double))
        st.sum2 = res["sum2"].ctypes.data_as(C.POINTER(C.c_double))
        if poles:
            res["sum_l2"] = np.zeros(n)
            res["sum_l4"] = np.zeros(n)
            st.sum_l2 = res["sum_l2"].ctypes.data_as(C.POINTER(C.c_double))
            st.sum_l4 = res["sum_l4"].ctypes.data_as(C.POINTER(C.c_double))
        return st, res

    # -- pipelines ----------------------------------------------------------
    def realise(self, re=None, im=None, seed=0, flags=F_SQRTPK, scale=1.0, field_out=None, spec_out=None,
                want_pk=False, poles=False, want_sums=True):
        """fb_realise.  Buffers may be numpy (host), DeviceBuffer or torch tensors."""
        st, res = (self._pk_struct(poles) if want_pk else (None, None))
        if want_pk and poles:
            flags |= F_POLES
        sums = (C.c_double * 2)()
        check(self.lib.fb_realise(self.h, _ptr(re), _ptr(im), int(seed), int(flags), float(scale), _ptr(field_out),
                                  _ptr(spec_out), C.byref(st) if st is not None else None,
                                  sums if want_sums else None))
        return res, (sums[0], sums[1])

    def spectrum_to_field(self, spec, field_out, flags=0, kind=KIND_PLAIN, scale=1.0):
        sums = (C.c_double * 2)()
        check(self.lib.fb_spectrum_to_field(self.h, _ptr(spec), int(flags), int(kind), float(scale),
                                            _ptr(field_out), sums))
        return sums[0], sums[1]

    def cube_to_field(self, cube, field_out, flags=0, part=0, scale=1.0):
        check(self.lib.fb_cube_to_field(self.h, _ptr(cube), int(flags), int(part), float(scale), _ptr(field_out)))

    def field_to_spectrum(self, field, spec_out=None, cross=None, want_pk=False, poles=False):
        st, res = (self._pk_struct(poles) if want_pk else (None, None))
        flags = (F_PK if want_pk else 0) | (F_POLES if poles else 0)
        check(self.lib.fb_field_to_spectrum(self.h, _ptr(field), _ptr(spec_out), _ptr(cross), flags,
                                            C.byref(st) if st is not None else None))
        return res

    def pk_from_spectrum(self, spec, cross=None, full_cube=False, poles=False):
        st, res = self._pk_struct(poles)
        check(self.lib.fb_pk_from_spectrum(self.h, _ptr(spec), _ptr(cross), int(bool(full_cube)),
                                           F_POLES if poles else 0, C.byref(st)))
        return res

    def pk2d_from_spectrum(self, spec, thr_perp, ipar, npar, cross=None, full_cube=False):
        """Moments on the (k_perp, |k_par|) grid: dict(count, sum1, sum2) of shape (nperp + 1, npar + 1)."""
        thr = np.ascontiguousarray(thr_perp, dtype=np.float64)
        ip = np.ascontiguousarray(ipar, dtype=np.int32)
        shape = (thr.size + 1, int(npar) + 1)
        out = dict(count=np.zeros(shape, np.uint64), sum1=np.zeros(shape), sum2=np.zeros(shape))
        check(self.lib.fb_pk2d_from_spectrum(self.h, _ptr(spec), _ptr(cross), int(bool(full_cube)), thr.ctypes.data,
                                             thr.size, ip.ctypes.data, int(npar), out["count"].ctypes.data,
                                             out["sum1"].ctypes.data, out["sum2"].ctypes.data))
        return out

    def correlation_function(self, field, edges, field_b=None, xi_out=None):
        """Radial moments of xi = ifftn(A conj B) / N^3: dict(count, sum1, sum2), index = np.digitize(r, edges)."""
        e = np.ascontiguousarray(edges, dtype=np.float64)
        out = dict(count=np.zeros(e.size + 1, np.uint64), sum1=np.zeros(e.size + 1), sum2=np.zeros(e.size + 1))
        check(self.lib.fb_correlation_function(self.h, _ptr(field), _ptr(field_b), e.ctypes.data, e.size,
                                               out["count"].ctypes.data, out["sum1"].ctypes.data,
                                               out["sum2"].ctypes.data, _ptr(xi_out)))
        return out

    def affine(self, field, n, mul, add):
        check(self.lib.fb_affine(self.h, _ptr(field), int(n), float(mul), float(add)))

    def exp_sum(self, src, dst, n, scale=1.0):
        s = C.c_double()
        check(self.lib.fb_exp_sum(self.h, _ptr(src), _ptr(dst), int(n), float(scale), C.byref(s)))
        return s.value

    def field_moments(self, field, n):
        a, b = C.c_double(), C.c_double()
        check(self.lib.fb_field_moments(self.h, _ptr(field), int(n), C.byref(a), C.byref(b)))
        return a.value, b.value

    def rsd_remap(self, delta, vel_z, vel_nl, zgrid, Hz, out, method="linear"):
        z = np.ascontiguousarray(zgrid, dtype=np.float64)
        if method not in RSD_METHODS:
            raise FastBoxError("rsd_remap: method must be one of %s" % (sorted(RSD_METHODS),))
        check(self.lib.fb_rsd_remap_method(self.h, _ptr(delta), _ptr(vel_z), _ptr(vel_nl), z.ctypes.data, float(Hz),
                                           RSD_METHODS[method], _ptr(out)))

    def beam_set(self, beam):
        """Transform and keep the beam cube (float32 [N][N][N]); later beam_convolve(None, ...) calls reuse it."""
        check(self.lib.fb_beam_set(self.h, _ptr(beam)))

    def beam_convolve(self, beam, field, out):
        """beam = None: convolve with the beam of the last beam_set / beam_convolve call."""
        check(self.lib.fb_beam_convolve(self.h, _ptr(beam), _ptr(field), _ptr(out)))

    def halo_counts(self, delta, nbar, nbar_kind, bias, bias_kind, lognormal, mean_exp, uniforms, counts_out,
                    mean_out=None):
        check(self.lib.fb_halo_counts(self.h, _ptr(delta), _ptr(nbar), int(nbar_kind), _ptr(bias), int(bias_kind),
                                      int(bool(lognormal)), float(mean_exp), _ptr(uniforms), _ptr(counts_out),
                                      _ptr(mean_out)))

    def counts_to_field(self, counts, n, mul, add, out):
        check(self.lib.fb_counts_to_field(self.h, _ptr(counts), int(n), float(mul), float(add), _ptr(out)))

    def halo_catalogue(self, counts, uniforms=None, cat_out=None, capacity=0):
        """Rows of the catalogue of `counts` (int32 [N^3]); with cat_out (float64 [capacity][3]) also fills it."""
        nh = _u64(0)
        check(self.lib.fb_halo_catalogue(self.h, _ptr(counts), _ptr(uniforms), _ptr(cat_out), int(capacity),
                                         C.byref(nh)))
        return int(nh.value)

    def fg_cube(self, amps, spectral_idx, log2_freq_ratio, out, accumulate=False):
        """out (+)= amps[x,y] * 2^(idx[x,y] * log2_freq_ratio[z]) (foregrounds.py:152-174); idx: (N,N) map or scalar."""
        idx = np.ascontiguousarray(np.atleast_1d(spectral_idx), dtype=np.float32)
        a = np.ascontiguousarray(amps, dtype=np.float32)
        l2 = np.ascontiguousarray(log2_freq_ratio, dtype=np.float32)
        check(self.lib.fb_fg_cube(self.h, _ptr(a), _ptr(idx), int(idx.size > 1), _ptr(l2), _ptr(out),
                                  int(bool(accumulate))))

    def radiometer_noise(self, sigma_z, out, normals=None, seed=0, accumulate=False):
        """out (+)= sigma_z[z] * n (noise.py:71-75); normals None -> device Philox keyed by seed."""
        sg = np.ascontiguousarray(sigma_z, dtype=np.float32)
        check(self.lib.fb_radiometer_noise(self.h, _ptr(sg), _ptr(normals), int(seed), _ptr(out),
                                           int(bool(accumulate))))

    def mean_spectrum_filter(self, field, out=None):
        """out = field - per-channel mean over (x, y) (filters.py:35-55); returns the means (float64 [N])."""
        mean = np.empty(self.N, dtype=np.float64)
        check(self.lib.fb_mean_spectrum_filter(self.h, _ptr(field), _ptr(out), mean.ctypes.data))
        return mean

    def pca_covariance(self, cube):
        """(mean [N], cov [N,N]) of a DEVICE float64 cube [N*N pixels][N channels] (filters.py:142,158-159)."""
        mean = np.empty(self.N, dtype=np.float64)
        cov = np.empty((self.N, self.N), dtype=np.float64)
        check(self.lib.fb_pca_covariance(self.h, _ptr(cube), mean.ctypes.data, cov.ctypes.data))
        return mean, cov

    def pca_project(self, cube, mean, U, cleaned, amps=None):
        """cleaned = cube - (U U^T (cube - mean) + mean) on the device (filters.py:173-178); U (N, nmodes)."""
        U = np.ascontiguousarray(U, dtype=np.float64)
        mean = np.ascontiguousarray(mean, dtype=np.float64)
        check(self.lib.fb_pca_project(self.h, _ptr(cube), mean.ctypes.data, U.ctypes.data, int(U.shape[1]),
                                      _ptr(cleaned), _ptr(amps)))

    def bench_fp64(self, mode=0):
        """Measured FP64 throughput in TFLOP/s (mode 0: FMA chains, 1: FP64 MMA)."""
        t = _d(0.0)
        check(self.lib.fb_bench_fp64(self.h, int(mode), C.byref(t)))
        return t.value

    def fft_pass_c2c(self, data, nplanes, axis_pass, sign):
        check(self.lib.fb_fft_pass_c2c(self.h, _ptr(data), int(nplanes), int(axis_pass), int(sign)))

    def fft_pass_x_c2r(self, spec, field, ncols, flags=0, scale=1.0):
        sums = (C.c_double * 2)()
        check(self.lib.fb_fft_pass_x_c2r(self.h, _ptr(spec), _ptr(field), int(ncols), int(flags), float(scale), sums))
        return sums[0], sums[1]

    def fft_pass_x_c2r_gather(self, spec, plane_off, field, ncols, flags=0, scale=1.0):
        sums = (C.c_double * 2)()
        check(self.lib.fb_fft_pass_x_c2r_gather(self.h, _ptr(spec), _ptr(plane_off), _ptr(field), int(ncols),
                                                int(flags), float(scale), sums))
        return sums[0], sums[1]

    def fft_pass_x_r2c(self, field, spec, ncols):
        check(self.lib.fb_fft_pass_x_r2c(self.h, _ptr(field), _ptr(spec), int(ncols)))

    def realise_local_kspace(self, seed, flags, work, send=None, ny=0, want_pk=False):
        st, res = (self._pk_struct(False) if want_pk else (None, None))
        check(self.lib.fb_realise_local_kspace(self.h, int(seed), int(flags), _ptr(work), _ptr(send), int(ny),
                                               C.byref(st) if st is not None else None))
        return res

    def forward_local_kspace(self, recv, work, ny, spec_out=None, want_pk=False, poles=False):
        st, res = (self._pk_struct(poles) if want_pk else (None, None))
        check(self.lib.fb_forward_local_kspace(self.h, _ptr(recv), _ptr(work), int(ny), _ptr(spec_out),
                                               F_POLES if poles else 0, C.byref(st) if st is not None else None))
        return res

    # -- multi-GPU (exchange inside the library) ---------------------------------
    DIST_HANDLE_BYTES = 128

    def dist_init(self, rank, world, with_forward=True):
        check(self.lib.fb_dist_init(self.h, int(rank), int(world), int(bool(with_forward))))

    def dist_handle(self):
        buf = np.zeros(self.DIST_HANDLE_BYTES, dtype=np.uint8)
        check(self.lib.fb_dist_get_handle(self.h, buf.ctypes.data))
        return buf

    def dist_connect(self, handles):
        h = np.ascontiguousarray(handles, dtype=np.uint8)
        check(self.lib.fb_dist_connect(self.h, h.ctypes.data))

    def dist_info(self):
        v = [C.c_int() for _ in range(6)]
        nb = C.c_size_t()
        check(self.lib.fb_dist_info(self.h, *[C.byref(x) for x in v], C.byref(nb)))
        return dict(zip(("rank", "world", "a0", "na", "y0", "ny"), [x.value for x in v]), block_bytes=nb.value)

    def dist_barrier(self):
        check(self.lib.fb_dist_barrier(self.h))

    def dist_realise(self, seed, flags, field_out, scale=1.0, chunks=4, phase=0, want_pk=False, want_sums=False):
        st, res = (self._pk_struct(False) if want_pk else (None, None))
        sums = (C.c_double * 2)()
        check(self.lib.fb_dist_realise(self.h, int(seed), int(flags), float(scale), int(chunks), int(phase),
                                       _ptr(field_out), C.byref(st) if st is not None else None,
                                       sums if want_sums else None))
        return res, (sums[0], sums[1])

    def dist_set_option(self, key, value):
        check(self.lib.fb_dist_set_option(self.h, key.encode(), int(value)))

    def dist_bench_exchange(self, iters=3):
        ms = C.c_float()
        check(self.lib.fb_dist_bench_exchange(self.h, int(iters), C.byref(ms)))
        return float(ms.value)

    def dist_power_spectrum(self, field, poles=False, phase=0, res=None):
        """``res`` (from an earlier phase of the same step) is filled when the step completes."""
        if res is None:
            st, out = self._pk_struct(poles)
            res = dict(out, _st=st)
        check(self.lib.fb_dist_power_spectrum(self.h, _ptr(field), F_POLES if poles else 0, int(phase),
                                              C.byref(res["_st"])))
        return res

    def bench_strided_copy(self, total_bytes, chunk_bytes, iters=5):
        g = C.c_double()
        check(self.lib.fb_bench_strided_copy(self.h, int(total_bytes), int(chunk_bytes), int(iters), C.byref(g)))
        return g.value

    def timer_start(self):
        check(self.lib.fb_timer_start(self.h))

    def timer_stop(self):
        ms = C.c_float()
        check(self.lib.fb_timer_stop(self.h, C.byref(ms)))
        return float(ms.value)

    def host_alloc(self, shape, dtype):
        """Pinned host array (freed with the plan)."""
        n = int(np.prod(shape)) * np.dtype(dtype).itemsize
        p = C.c_void_p()
        check(self.lib.fb_host_alloc(C.byref(p), n))
        self._keep.append(p.value)
        buf = (C.c_char * n).from_address(p.value)
        return np.frombuffer(buf, dtype=dtype).reshape(shape)

    def last_timings(self, n=3):
        ms = (C.c_float * n)()
        check(self.lib.fb_last_timings(self.h, ms, n))
        return [float(x) for x in ms]


def launch_count():
    return int(load().fb_launch_count())


def device_info(device=0):
    name = C.create_string_buffer(256)
    sm = C.c_int()
    mem = C.c_size_t()
    check(load().fb_device_info(device, name, 256, C.byref(sm), C.byref(mem)))
    return name.value.decode(), sm.value, mem.value
