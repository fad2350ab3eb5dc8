/*
 * fastbox_b200.h -- C ABI of libfastbox_b200.so (hand-written sm_100a CUDA).
 *
 * The reference (philbull/FastBox) is pure Python: its "FFI" for this path is
 * the set of NumPy/SciPy calls made by fastbox/box.py, fastbox/beams.py and
 * fastbox/halos.py.  Every entry point below names the reference lines whose
 * work it replaces.  The Python shim (fastbox_b200/_lib.py) binds these with
 * ctypes; INTEGRATION.md shows the stub a FastBox maintainer would add.
 *
 * Conventions
 *   - return 0 on success, negative on error; message via fb_last_error()
 *     (thread local).
 *   - "buffer" arguments may be HOST pointers (pageable or pinned) or DEVICE
 *     pointers; the library detects which (cudaPointerGetAttributes) and stages
 *     host buffers through plan-owned device memory.  The caller owns every
 *     buffer it passes; the library owns only plan workspaces.
 *   - arrays are C-order [x][y][z], z = line of sight (fastbox/beams.py:70-71).
 *   - "half spectrum": complex64 planes kx = 0..N/2 of the 3-D DFT of a real
 *     field, shape [N/2+1][N][N] (axis 0 halved; these are the first N/2+1
 *     planes of the reference's full `delta_k`, fastbox/box.py:193).
 *   - all work is enqueued on the plan's stream; results written to host
 *     buffers are complete when the call returns, results in device buffers
 *     after fb_sync().
 *   - a plan is not thread safe; different plans are independent.
 */
#ifndef FASTBOX_B200_H
#define FASTBOX_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct fb_plan fb_plan;

/* ---- k-space multiplier applied in the first pass of an inverse transform -- */
enum {
    FB_KIND_PLAIN = 0,      /* field_k * amp                                          */
    FB_KIND_VEL_X = 1,      /* i k_x / k^2, Nyquist plane of x zeroed  box.py:254,268-274 */
    FB_KIND_VEL_Y = 2,      /* i k_y / k^2                             box.py:255       */
    FB_KIND_VEL_Z = 3,      /* i k_z / k^2                             box.py:256       */
    FB_KIND_POTENTIAL = 4   /* 1 / k^2, DC = 0                         box.py:347-348   */
};

/* flags for fb_realise / fb_spectrum_to_field */
enum {
    FB_F_SQRTPK = 1,        /* multiply by the sqrt(P(k) boxfactor) table   box.py:161-176 */
    FB_F_FILTER = 2,        /* multiply by the transfer-function table      box.py:374-379 */
    FB_F_EXP = 4,           /* output exp(scale*field), accumulate its sum  box.py:457     */
    FB_F_ANTIHERM = 8,      /* take the anti-Hermitian part (gives Im of ifftn)            */
    FB_F_PK = 16,           /* bin |H|^2/boxfactor of the spectrum being transformed       */
    FB_F_POLES = 32         /* with FB_F_PK: also l=2,4 Legendre sums                      */
};

/* Binned-moment output, nb = number of edges; index i follows np.digitize
 * (fastbox/box.py:758): i = #{edges <= k}; i = 0..nb.  All arrays have nb+1
 * entries.  mean = sum1/count etc. are formed by the caller exactly like
 * box.py:761-768.                                                            */
typedef struct fb_pk_result {
    uint64_t* count;        /* multiplicity-weighted number of modes            */
    double* sum1;           /* sum w * p,  p = |d_k|^2 / boxfactor (or Re a b*)  */
    double* sum2;           /* sum w * p^2                                      */
    double* sum_l2;         /* sum w * p * L_2(mu)   (nullable)                 */
    double* sum_l4;         /* sum w * p * L_4(mu)   (nullable)                 */
} fb_pk_result;

/* ---- plan ---------------------------------------------------------------- */
/* Replaces CosmoBox.__init__ grid set-up, box.py:76-101,110-127 (the N^3
 * Kx,Ky,Kz,k arrays are never materialised).  N must be a power of two in
 * [8, 2048].                                                                 */
int fb_plan_create(fb_plan** plan, int N, double Lx, double Ly, double Lz, int device);
int fb_plan_destroy(fb_plan* plan);
int fb_sync(fb_plan* plan);
const char* fb_last_error(void);
const char* fb_version(void);
/* number of kernels launched by this library since load (process-wide)       */
uint64_t fb_launch_count(void);

/* ---- slab (multi-GPU) geometry ------------------------------------------- */
/* Restrict the plan to spectrum planes kx in [a0, a0+na) and real-space rows
 * y in [y0, y0+ny) (slab decomposition).  Default: everything. */
int fb_plan_set_slab(fb_plan* plan, int a0, int na, int y0, int ny);

/* ---- memory helpers (the shim keeps fields device resident between calls) - */
int fb_dev_alloc(void** ptr, size_t bytes);            /* on the current device */
int fb_dev_alloc_on(int device, void** ptr, size_t bytes);   /* on a named device (a plan's) */
int fb_dev_free(void* ptr);
int fb_host_alloc(void** ptr, size_t bytes);        /* pinned */
int fb_host_free(void* ptr);
int fb_copy(fb_plan* plan, void* dst, const void* src, size_t bytes);   /* any direction, async on the plan stream + sync */
int fb_convert_f64_to_f32(fb_plan* plan, const double* src, float* dst, size_t n);
int fb_convert_f32_to_f64(fb_plan* plan, const float* src, double* dst, size_t n);
int fb_device_info(int device, char* name, int name_len, int* sm_count, size_t* total_mem);

/* ---- tables ---------------------------------------------------------------- */
/* sqrt(P(k) * boxfactor), box.py:161-171.  mode 1: cubic box, exact LUT indexed
 * by the integer i^2+j^2+l^2 (n entries);  mode 2: table uniform in log2(s),
 * s = (i/Lx)^2+(j/Ly)^2+(l/Lz)^2, with origin log2s0 and spacing dlog2s (cubic interpolation);
 * mode 3: table indexed by the leading bits of float32(s): entry i is the value at the float whose
 * bit pattern is (i + base) << (23 - M); pass base in `log2s0` and M (mantissa bits) in `dlog2s`. */
int fb_set_sqrt_pk(fb_plan* plan, const float* table, long n, int mode, double log2s0, double dlog2s);
/* transfer function T(k_perp, k_par), box.py:374-378.  Separable form
 * tperp[(N/2+1)*N] (index a*N+b) times tpar[N]; or dense [(N/2+1)*N*N].
 * Pass NULLs to clear.                                                        */
int fb_set_filter(fb_plan* plan, const float* tperp, const float* tpar, const float* tdense);
/* Bin edges for P(k), box.py:745-758: `thresholds[j]` is the smallest double s
 * with 2*pi*sqrt(s) >= edge[j] (computed by the shim with NumPy, so that the
 * device bin index is bit-identical to np.digitize without a device sqrt).    */
int fb_set_pk_bins(fb_plan* plan, const double* thresholds, int nedges);

/* ---- realise: box.py:174-193 (+ :378, :457, tracers bias) ------------------ */
/* White noise (re, im) [N][N][N] float32 as drawn at box.py:174-175, or NULL
 * for on-device Philox4x32-10 noise keyed by (seed, cell index).  Computes
 *   H = amp * 1/2 [W(k) + conj W(-k)]        (= reference delta_k, box.py:193)
 *   field = Re ifftn(W * amp) * scale         (box.py:187; scale = bias)
 * optionally exp() of it with the sum returned in *sum_out (box.py:457-458),
 * optionally H stored to spec_out (half spectrum) and binned into pk.         */
int fb_realise(fb_plan* plan, const float* re, const float* im, uint64_t seed, int flags, float scale,
               float* field_out, void* spec_out, fb_pk_result* pk, double* sum_out);

/* ---- inverse transform of a stored half spectrum: box.py:380, 254-285, 347 -- */
int fb_spectrum_to_field(fb_plan* plan, const void* spec_half, int flags, int kind, float scale,
                         float* field_out, double* sum_out);
/* Same from a FULL complex cube [N][N][N] (complex64) that need not be
 * Hermitian: part 0 -> Re ifftn(cube*T), part 1 -> Im ifftn(cube*T).          */
int fb_cube_to_field(fb_plan* plan, const void* cube, int flags, int part, float scale, float* field_out);

/* ---- forward transform + P(k): box.py:736-764 ------------------------------- */
/* field [N][N][N] float32 -> optional half spectrum, optional binned moments
 * (auto, or cross with `cross_spec` half spectrum; FB_F_POLES adds l=2,4).    */
int fb_field_to_spectrum(fb_plan* plan, const float* field, void* spec_out, const void* cross_spec, int flags,
                         fb_pk_result* pk);
/* binned moments straight from a stored spectrum.  full_cube = 0: half
 * spectrum with Hermitian multiplicities; 1: full [N][N][N] cube, weight 1.   */
int fb_pk_from_spectrum(fb_plan* plan, const void* spec, const void* cross_spec, int full_cube, int flags,
                        fb_pk_result* pk);

/* ---- P(k_perp, |k_par|) and xi(r) (the reference delegates both to nbodykit,
 * examples/example_endtoend.py:128-151: parity is pinned to oracle/restate.py, unpinned w.r.t. nbodykit) ------ */
/* thr_perp[nperp]: thresholds on s = (Kx/Lx)^2 + (Ky/Ly)^2 as in fb_set_pk_bins; ipar[N]: np.digitize bin of
 * |2 pi Kz / Lz| for every z mode (host, int32); outputs host arrays [(nperp+1)*(npar+1)], row = k_perp bin.
 * spec: half spectrum (full_cube = 0) or full [N][N][N] cube; cross_spec nullable.                            */
int fb_pk2d_from_spectrum(fb_plan* plan, const void* spec, const void* cross_spec, int full_cube,
                          const double* thr_perp, int nperp, const int32_t* ipar, int npar, uint64_t* count,
                          double* sum1, double* sum2);
/* xi(r) = ifftn(fftn(a) conj fftn(b)) / N^3 binned over the lag separations (periodic, cell size L/N);
 * field_b nullable (auto-correlation); edges[nedges] host float64 ascending; outputs host [nedges+1], index =
 * np.digitize(r, edges); xi_out: nullable DEVICE float32 [N^3] receiving the lag cube.                        */
int fb_correlation_function(fb_plan* plan, const float* field, const float* field_b, const double* edges, int nedges,
                            uint64_t* count, double* sum1, double* sum2, float* xi_out);

/* ---- elementwise ------------------------------------------------------------- */
/* field = field * mul + add  (log-normal normalise box.py:458-459, Tb(1+d))   */
int fb_affine(fb_plan* plan, float* field, size_t n, float mul, float add);
/* out = exp(scale * in), *sum_out = sum(out)  (box.py:457; halos.py:106)      */
int fb_exp_sum(fb_plan* plan, const float* in, float* out, size_t n, float scale, double* sum_out);
/* sum and sum of squares of a field (box.py:944 Parseval)                     */
int fb_field_moments(fb_plan* plan, const float* field, size_t n, double* sum, double* sumsq);

/* ---- redshift-space remap: box.py:412-437 ------------------------------------ */
/* z[N] grid coordinates (host, float64); vel_nl nullable (sigma_nl * N(0,1)).  */
int fb_rsd_remap(fb_plan* plan, const float* delta, const float* vel_z, const float* vel_nl, const double* zgrid,
                 double Hz, float* out);
/* the `method` keyword the reference forwards to scipy.interpolate.griddata (box.py:403-405, 433-437):
 * FB_RSD_LINEAR = 'linear' (fb_rsd_remap), FB_RSD_NEAREST = 'nearest' (nearest sample, half-way points go to
 * the lower sample, end samples beyond the range).  'cubic' is not offered: scipy's 1-D cubic is a global
 * not-a-knot spline through the sorted samples, which rejects the duplicate sample that the periodic wrap of
 * an unmoved end point creates (s_{N-1} -> z_0), so the reference itself fails on it for small velocities. */
#define FB_RSD_LINEAR 0
#define FB_RSD_NEAREST 1
int fb_rsd_remap_method(fb_plan* plan, const float* delta, const float* vel_z, const float* vel_nl,
                        const double* zgrid, double Hz, int method, float* out);

/* ---- beam convolution: beams.py:81-87 ----------------------------------------- */
/* out = fftconvolve(beam, field, mode='same', axes=[0,1]) / sum_xy beam  per channel z.
 * fb_beam_set transforms the beam cube once (float64 channel sums, zero-padded 2-D spectrum divided by
 * (2N)^2 norm[z], 16 B/cell of device memory held by the plan until the next fb_beam_set / fb_plan_destroy);
 * fb_beam_convolve with beam = NULL reuses it, with a beam cube it calls fb_beam_set first (the reference's
 * call: the beam is re-derived on every convolve_fft, beams.py:79).  N <= 1024.                                */
int fb_beam_set(fb_plan* plan, const float* beam);
int fb_beam_convolve(fb_plan* plan, const float* beam, const float* field, float* out);

/* ---- halo counts: halos.py:91-117 ---------------------------------------------- */
/* nbar/bias kinds: 0 scalar (pointer to 1 float), 1 per-z [N], 2 per-voxel.
 * mean_exp: mean of exp(bias*delta) when lognormal != 0 (from fb_exp_sum).
 * uniforms: one float64 U[0,1) per voxel; counts by inversion (fb_poisson.h). */
int fb_halo_counts(fb_plan* plan, const float* delta, const float* nbar, int nbar_kind, const float* bias,
                   int bias_kind, int lognormal, double mean_exp, const double* uniforms, int32_t* counts_out,
                   float* mean_out);

/* out = counts * mul + add (float32): the halo overdensity N_h / N_bar - 1 fed to the cross spectrum,
 * examples/example_halos.py:46-53.  Device buffers.                                                */
int fb_counts_to_field(fb_plan* plan, const int32_t* counts, size_t n, float mul, float add, float* out);

/* ---- halo catalogue: halos.py:120-176 ------------------------------------------ */
/* counts[N^3] (int32, >= 0, < 1024) -> cat[nhalo][3] float64 comoving positions in the reference's order
 * (ascending count value, then C-order voxel index, each voxel repeated `count` times).
 * uniforms: NULL (no scatter) or [nhalo][3] U[0,1) offsets in catalogue order (halos.py:163-166).
 * cat_out == NULL: only *nhalo_out is set (size query).  capacity = rows cat_out can hold. */
int fb_halo_catalogue(fb_plan* plan, const int32_t* counts, const double* uniforms, double* cat_out,
                      uint64_t capacity, uint64_t* nhalo_out);

/* ---- foreground cube: foregrounds.py:152-174 ------------------------------------- */
/* out[x,y,z] (+)= amps[x,y] * 2^(idx[x,y] * log2_freq_ratio[z]);  log2_freq_ratio[z] = log2(freqs[z]/freq_ref)
 * (float32 [N], evaluated in float64 by the caller); idx_is_map = 0: spectral_idx points to one float.
 * accumulate != 0 adds to the cube already in `out` (signal + foregrounds in one pass). */
int fb_fg_cube(fb_plan* plan, const float* amps, const float* spectral_idx, int idx_is_map,
               const float* log2_freq_ratio, float* out, int accumulate);

/* ---- radiometer noise: noise.py:55-75 --------------------------------------------- */
/* out[x,y,z] (+)= sigma_z[z] * n[x,y,z]; normals [N^3] float32 or NULL = Philox4x32-10 N(0,1) keyed by
 * (seed, cell index) on the device. */
int fb_radiometer_noise(fb_plan* plan, const float* sigma_z, const float* normals, uint64_t seed, float* out,
                        int accumulate);

/* ---- mean_spectrum_filter: filters.py:35-55 ---------------------------------------- */
/* out[x,y,z] = field[x,y,z] - mean_xy(field[:,:,z]); mean_out [N] float64 (nullable); out nullable
 * (mean only).  The per-channel sums are accumulated in float64. */
int fb_mean_spectrum_filter(fb_plan* plan, const float* field, float* out, double* mean_out);

/* ---- PCA foreground filter, device steps in float64: filters.py:93-183 ------------ */
/* cube: DEVICE float64 [N*N pixels][N channels] (= field.reshape(-1, Nf) of the reference).
 * fb_pca_covariance: mean_out[N] = per-channel mean (filters.py:142), cov_out[N*N] = np.cov of the
 * channels over the pixels (filters.py:158-159, normalised by Npix-1); host or device outputs.
 * fb_pca_project: amps[nmodes][Npix] = U^T (cube - mean) (nullable, device), cleaned = cube - (U amps + mean)
 * (filters.py:173-178); U [N][nmodes] float64 row-major, 1 <= nmodes <= 32; cleaned: device.
 * The eigen-decomposition of cov between the two calls is the caller's (N x N, host). */
int fb_pca_covariance(fb_plan* plan, const double* cube, double* mean_out, double* cov_out);
int fb_pca_project(fb_plan* plan, const double* cube, const double* mean, const double* U, int nmodes,
                   double* cleaned, double* amps);

/* FP64 throughput probe (no memory traffic): mode 0 = FMA chains, 1 = FP64 MMA m8n8k4; TFLOP/s. */
int fb_bench_fp64(fb_plan* plan, int mode, double* tflops);

/* ---- building blocks exposed for tests / multi-GPU orchestration ---------------- */
/* pass = 0: rows (z, contiguous) c2c; 1: columns (y) c2c; sign = -1 fwd / +1 inv;
 * data: [nplanes][N][N] complex64, in place.                                   */
int fb_fft_pass_c2c(fb_plan* plan, void* data, int nplanes, int pass, int sign);
/* x pass, real <-> half complex over planes; ncols = columns per plane.        */
int fb_fft_pass_x_c2r(fb_plan* plan, const void* spec, float* field, long ncols, int flags, float scale,
                      double* sum_out);
/* same, with the kx planes gathered through plane_off[k] (device array of N/2+1 element offsets
 * relative to `spec`): lets the all-to-all be issued in chunks that overlap the k-space passes. */
int fb_fft_pass_x_c2r_gather(fb_plan* plan, const void* spec, const long* plane_off, float* field, long ncols,
                             int flags, float scale, double* sum_out);
int fb_fft_pass_x_r2c(fb_plan* plan, const float* field, void* spec, long ncols);
/* Slab-decomposed (multi-GPU) halves of the pipelines; the all-to-all between them is done by
 * the caller (torch.distributed / NCCL, fastbox_b200/dist.py).  `ny` = y rows per rank.
 * realise: Philox noise -> Hermitian spectrum on the local kx planes -> z rows, y columns;
 * the result is written to `send` as [dest rank][local plane][ny][N] (contiguous per peer).  */
int fb_realise_local_kspace(fb_plan* plan, uint64_t seed, int flags, void* work, void* send, int ny,
                            fb_pk_result* pk);
/* forward: `recv` = [src rank -> local plane... ] i.e. [d][local plane][ny][N] as received;
 * y columns + z rows on the local kx planes, optional spectrum store and binned moments.      */
int fb_forward_local_kspace(fb_plan* plan, const void* recv, void* work, int ny, void* spec_out, int flags,
                            fb_pk_result* pk);
/* ---- multi-GPU pipelines with the exchange inside the library (fb_dist.cu) ------------------------
 * One process per GPU.  Every rank calls fb_dist_init (sets the slab of `rank` out of `world`: kx planes
 * [rank*N/2/world, ...) -- the last rank also owns the Nyquist plane -- and y rows [rank*N/world, ...);
 * allocates the exchange block), publishes the FB_DIST_HANDLE_BYTES blob of fb_dist_get_handle to all ranks
 * by any means (fastbox_b200/dist.py: one torch.distributed all_gather at set-up) and passes the `world`
 * blobs in rank order to fb_dist_connect, which maps the peers' blocks (CUDA IPC over NVLink).  From then on
 * the transposes of the 3-D transforms are the STORES of the FFT pass before them, straight into the peer's
 * receive buffer; P(k) moments are summed through per-rank slots; ranks meet at device-side epoch flags.
 * No NCCL call is on the data path.  All calls are collective (same order on every rank).              */
#define FB_DIST_HANDLE_BYTES 128
int fb_dist_init(fb_plan* plan, int rank, int world, int with_forward);
int fb_dist_get_handle(fb_plan* plan, void* handle_out);
int fb_dist_connect(fb_plan* plan, const void* handles);
int fb_dist_info(fb_plan* plan, int* rank, int* world, int* a0, int* na, int* y0, int* ny, size_t* block_bytes);
int fb_dist_barrier(fb_plan* plan);
/* realise (+ filter + P(k)) of box.py:161-193 on this rank's slab, Philox noise keyed by the global cell
 * index (any GPU count gives the same field).  field_out: DEVICE float32 [N][ny][N]; pk: moments summed over
 * all ranks (nullable); sums_out[2]: local sum / sum of squares (nullable).  `chunks` first-pass chunks
 * overlap the peer stores of the y pass.  phase 0 = whole step; 1 = up to the signal; 2 = from the wait.  */
int fb_dist_realise(fb_plan* plan, uint64_t seed, int flags, float scale, int chunks, int phase, float* field_out,
                    fb_pk_result* pk, double* sums_out);
/* the exchange alone (y pass storing into the peers + barrier), average ms per iteration: NVLink roofline */
int fb_dist_bench_exchange(fb_plan* plan, int iters, float* ms_out);
/* tuning knobs of the exchange: "xmode" (0 peer stores of the y pass, 1 copy engines, 2 copy kernel), "push_ctas"
 * (CTAs per peer of the copy kernel), "cz_cols" (columns per tile of the exchanging y pass)                     */
int fb_dist_set_option(fb_plan* plan, const char* key, int value);
/* binned P(k) of the sharded real field (box.py:736-764); field: DEVICE float32 [N][ny][N].
 * phase 0 = all; 1 = x pass + signal; 2 = wait + k-space passes + signal; 3 = wait + sum.               */
int fb_dist_power_spectrum(fb_plan* plan, const float* field, int flags, int phase, fb_pk_result* pk);

/* strided HBM copy micro-benchmark: rows of `chunk_bytes`, returns GB/s        */
int fb_bench_strided_copy(fb_plan* plan, size_t total_bytes, int chunk_bytes, int iters, double* gbs);
/* CUDA-event stopwatch on the plan's stream (bench.py times the step loop with it) */
int fb_timer_start(fb_plan* plan);
int fb_timer_stop(fb_plan* plan, float* ms);
/* time (ms, CUDA events on the plan stream) of the last pipeline call's kernels */
int fb_last_timings(fb_plan* plan, float* ms, int n);

#ifdef __cplusplus
}
#endif
#endif
