// fb_misc.cu -- streaming kernels around the FFT pipeline: affine / exp / moments,
// dtype conversion, redshift-space remap (box.py:412-437), halo counts (halos.py:91-117),
// and a strided-copy probe used to pick tile widths.
#include "fb_launch.h"
#include "../../include/fb_poisson.h"

namespace fb {

static inline unsigned grid_for(size_t n, int per_block, int sm_count) {
    size_t b = (n + per_block - 1) / per_block;
    const size_t cap = (size_t)sm_count * 16;
    return (unsigned)(b < cap ? (b ? b : 1) : cap);
}

__global__ void __launch_bounds__(256) k_affine(float* __restrict__ x, size_t n, float mul, float add) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t n4 = n / 4;
    float4* x4 = reinterpret_cast<float4*>(x);
    for (; i < n4; i += stride) {
        float4 v = x4[i];
        v.x = fmaf(v.x, mul, add);
        v.y = fmaf(v.y, mul, add);
        v.z = fmaf(v.z, mul, add);
        v.w = fmaf(v.w, mul, add);
        x4[i] = v;
    }
    for (size_t j = n4 * 4 + (size_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += stride)
        x[j] = fmaf(x[j], mul, add);
}

__device__ __forceinline__ void block_sum2(double a, double b, double* out) {
    __shared__ double red[2][32];
    a = warp_sum(a);
    b = warp_sum(b);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) {
        red[0][warp] = a;
        red[1][warp] = b;
    }
    __syncthreads();
    if (warp == 0) {
        const int nw = (blockDim.x + 31) >> 5;
        a = lane < nw ? red[0][lane] : 0.0;
        b = lane < nw ? red[1][lane] : 0.0;
        a = warp_sum(a);
        b = warp_sum(b);
        if (lane == 0) {
            atomicAdd(&out[0], a);
            atomicAdd(&out[1], b);
        }
    }
}

// out = exp(scale*in); sums[0] += sum(out)
__global__ void __launch_bounds__(256) k_exp_sum(const float* __restrict__ in, float* __restrict__ out, size_t n,
                                                  float scale, double* sums) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    double acc = 0.0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const float e = expf(in[i] * scale);
        out[i] = e;
        acc += (double)e;
    }
    block_sum2(acc, 0.0, sums);
}

__global__ void __launch_bounds__(256) k_moments(const float* __restrict__ in, size_t n, double* sums) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    double a = 0.0, b = 0.0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const double v = (double)in[i];
        a += v;
        b += v * v;
    }
    block_sum2(a, b, sums);
}

__global__ void __launch_bounds__(256) k_f64_to_f32(const double* __restrict__ src, float* __restrict__ dst, size_t n) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) dst[i] = (float)src[i];
}
__global__ void __launch_bounds__(256) k_f32_to_f64(const float* __restrict__ src, double* __restrict__ dst, size_t n) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) dst[i] = (double)src[i];
}

// ---------------------------------------------------------------------------
// Redshift-space remap (box.py:412-437), several lines of sight per CTA, software pipelined.
//   s_l = z_l - (v_z + v_nl)_l / H, wrapped periodically into [z_0, z_{N-1});   out_l = linear
//   re-grid of the scattered samples (s, delta) onto z  (scipy griddata 1-D = argsort +
//   interp1d(linear, fill_value)): out(z) = y_lo + (y_hi-y_lo)/(x_hi-x_lo) (z-x_lo) with
//   x_lo = largest sample < z, x_hi = smallest sample >= z; fill = (delta_0+delta_{N-1})/2 outside
//   [min s, max s].
// No sort is needed.  Every sample is dropped into the grid cell that contains it, keeping per cell the
// largest and the smallest sample (native 32-bit shared-memory atomic max / min on a key = position inside the
// cell | sample index).  The bracket of output z_l is then the max of the nearest non-empty cell below l and
// the min of the nearest non-empty cell at or above l -- the pair the sorted search returns -- found in O(1)
// expected steps.
// Positions are carried in CELL UNITS: one float64 multiply-add, floor and multiply give the wrapped position
// p = frac((z_l - z_0)/length - v/(H length)) (N-1) to 1e-16 of the box; it is then split into the integer
// cell and a float32 fraction in [0, 1) (6e-8 of a cell; a fraction that rounds to 1 moves to the next cell,
// so a sample that sits on a grid point -- v = 0 -- is found there exactly).  The interpolation weight
// ((l - c_lo) - f_lo) / ((c_hi - c_lo) + (f_hi - f_lo)) needs float32 only.  This replaced float64 positions
// with ~2.5x the instructions (the kernel is issue bound, not HBM bound).
// ---------------------------------------------------------------------------
#define FB_RSD_IDX_BITS 12
#define FB_RSD_LPC 8            // lines of sight per CTA (software pipelined)
template <int N>
struct RsdGeom {
    static constexpr int NT = N >= 256 ? 256 : (N < 32 ? 32 : N);
    static constexpr int E = (N + NT - 1) / NT;                      // elements per thread
    static constexpr size_t SMEM = (size_t)N * (8 + 2 * (4 + 4 + 4 + 4));   // t0 + 2 x (frac, y, cmax, cmin)
};

// METHOD 0 = 'linear' (above); 1 = 'nearest': scipy griddata 1-D forwards to interp1d(kind='nearest',
// fill_value='extrapolate'): the sample nearest to z_l, the half-way point going to the lower sample
// (searchsorted(side='left') on the mid-points), the end samples outside [min s, max s] -- i.e. a choice
// between the same two bracket samples by their distances (l - c_lo) - f_lo and (c_hi - l) + f_hi.
template <int N, int METHOD>
__global__ void __launch_bounds__(RsdGeom<N>::NT) k_rsd_remap(const float* __restrict__ delta,
                                                              const float* __restrict__ vel,
                                                              const float* __restrict__ vnl,
                                                              const double* __restrict__ zgrid, double Hz,
                                                              float* __restrict__ out, long nlines) {
    constexpr int NT = RsdGeom<N>::NT, E = RsdGeom<N>::E;
    extern __shared__ __align__(16) unsigned char rsd_smem[];
    double* t0 = reinterpret_cast<double*>(rsd_smem);                 // (z_l - z_0) / length  [N]
    float* pf2 = reinterpret_cast<float*>(t0 + N);                    // fraction inside the cell [2][N]
    float* y2 = pf2 + 2 * N;                                          // sample values [2][N]
    unsigned* cmax2 = reinterpret_cast<unsigned*>(y2 + 2 * N);        // per cell: largest sample (32-bit key)
    unsigned* cmin2 = cmax2 + 2 * N;                                  // per cell: smallest sample
    const int tid = threadIdx.x;
    const double zmin = zgrid[0], zmax = zgrid[N - 1];       // increasing grid (linspace, box.py:79-88)
    const double length = zmax - zmin;
    const double inv_len = 1.0 / length, c1 = inv_len / Hz;
    // key = position inside the cell (20 bits) | sample index + 1 (12 bits): native 32-bit smem atomics
    constexpr unsigned LOW = (1u << FB_RSD_IDX_BITS) - 1u;
    for (int l = tid; l < N; l += NT) {
        t0[l] = (zgrid[l] - zmin) * inv_len;
        cmax2[l] = 0u;
        cmin2[l] = ~0u;
    }
    const long line0 = (long)blockIdx.x * FB_RSD_LPC;
    const long line1 = line0 + FB_RSD_LPC < nlines ? line0 + FB_RSD_LPC : nlines;
    float dcur[E], vcur[E];
    auto fetch = [&](long ln, float (&d)[E], float (&v)[E]) {
#pragma unroll
        for (int e = 0; e < E; ++e) {
            const int l = tid + e * NT;
            if (l < N) {
                d[e] = delta[(size_t)ln * N + l];
                v[e] = vel[(size_t)ln * N + l] + (vnl ? vnl[(size_t)ln * N + l] : 0.f);
            }
        }
    };
    if (line0 < line1) fetch(line0, dcur, vcur);
    __syncthreads();
    int cur = 0;
    for (long ln = line0; ln < line1; ++ln, cur ^= 1) {
        float* pf = pf2 + cur * N;
        float* y = y2 + cur * N;
        unsigned *cmax = cmax2 + cur * N, *cmin = cmin2 + cur * N;
        float dnext[E], vnext[E];
        if (ln + 1 < line1) fetch(ln + 1, dnext, vnext);            // in flight while this line is processed
        // ---- phase 1: wrapped positions in cell units, per-cell extremes
#pragma unroll
        for (int e = 0; e < E; ++e) {
            const int l = tid + e * NT;
            if (l < N) {
                // (s - zmin) % length with Python's sign convention (box.py:422-426), as a fraction of the box
                double t = fma(-(double)vcur[e], c1, t0[l]);
                t -= floor(t);
                const double p = t * (double)(N - 1);
                int c = (int)p;
                float f = (float)(p - (double)c);
                if (f >= 1.f) {                                  // rounds onto the next grid point
                    c += 1;
                    f = 0.f;
                }
                c = min(c, N - 1);
                pf[l] = f;
                y[l] = dcur[e];
                const unsigned qpos = min((unsigned)(f * 1048576.f), 1048575u);
                const unsigned key = (qpos << FB_RSD_IDX_BITS) | (unsigned)(l + 1);
                atomicMax(&cmax[c], key);
                atomicMin(&cmin[c], key);
            }
        }
        __syncthreads();
        // ---- phase 2: the bracket of every grid point
        const float fill = 0.5f * (y[0] + y[N - 1]);           // box.py:429
#pragma unroll
        for (int e = 0; e < E; ++e) {
            const int l = tid + e * NT;
            if (l < N) {
                int il = -1, ih = -1, clo = 0, chi = 0;
                for (int c = l - 1; c >= 0; --c) {             // largest sample < z_l
                    const unsigned k = cmax[c];
                    if (k) { il = (int)(k & LOW) - 1; clo = c; break; }
                }
                for (int c = l; c < N; ++c) {                  // smallest sample >= z_l
                    const unsigned k = cmin[c];
                    if (k != ~0u) { ih = (int)(k & LOW) - 1; chi = c; break; }
                }
                float r = fill;                                // outside [min s, max s]
                if (METHOD == 1) {
                    if (ih >= 0 && il >= 0) {
                        const float dlo = (float)(l - clo) - pf[il];
                        const float dhi = (float)(chi - l) + pf[ih];
                        r = dhi < dlo ? y[ih] : y[il];
                    } else {
                        r = ih >= 0 ? y[ih] : y[il];           // extrapolation = the end sample (a line has N >= 8 samples)
                    }
                } else if (ih >= 0) {
                    const float fh = pf[ih], yh = y[ih];
                    if (il >= 0) {
                        const float fl = pf[il], yl = y[il];
                        const float num = (float)(l - clo) - fl;
                        const float den = (float)(chi - clo) + (fh - fl);
                        const float wgt = den > 0.f ? __fdividef(num, den) : 0.f;
                        r = fmaf(yh - yl, wgt, yl);
                    } else if (chi == l && fh == 0.f) {
                        r = yh;                                // z_l is the smallest sample itself
                    }
                }
                out[(size_t)ln * N + l] = r;
                cmax2[(cur ^ 1) * N + l] = 0u;                 // cells of the next line (idle during this one)
                cmin2[(cur ^ 1) * N + l] = ~0u;
            }
        }
#pragma unroll
        for (int e = 0; e < E; ++e) {
            dcur[e] = dnext[e];
            vcur[e] = vnext[e];
        }
        __syncthreads();                                       // resets visible; this line's readers are done
    }
}

// ---------------------------------------------------------------------------
// halo counts: mean count lambda in float64 from the float32 density, then
// Poisson inversion from the supplied uniform (halos.py:104-117).
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_exp_sum_f64(const float* __restrict__ delta, const float* __restrict__ bias,
                                                      int bias_kind, int N, size_t n, double* sums) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    double acc = 0.0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const double b = (double)(bias_kind == 0 ? bias[0] : (bias_kind == 1 ? bias[i % N] : bias[i]));
        acc += exp(b * (double)delta[i]);
    }
    block_sum2(acc, 0.0, sums);
}

// Four consecutive voxels per thread: 16-byte loads / stores and four independent float64 chains
// (the inversion is a long dependent sequence of correctly rounded operations; one voxel per
// thread left the kernel waiting on its own latency).  n % 4 == 0 and N % 4 == 0 (N is a power of two >= 4).
__global__ void __launch_bounds__(256) k_halo_counts(const float* __restrict__ delta, const float* __restrict__ nbar,
                                                      int nbar_kind, const float* __restrict__ bias, int bias_kind,
                                                      int lognormal, double mean_exp, double voxel_vol,
                                                      const double* __restrict__ uniforms, int N, size_t n,
                                                      int32_t* __restrict__ counts, float* __restrict__ mean_out) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    const size_t n4 = n / 4;
    for (size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x; q < n4; q += stride) {
        const size_t i = 4 * q;
        const float4 d4 = __ldg(reinterpret_cast<const float4*>(delta) + q);
        const float d[4] = {d4.x, d4.y, d4.z, d4.w};
        float bf[4], nf[4];
        if (bias_kind == 0) {
            bf[0] = bf[1] = bf[2] = bf[3] = bias[0];
        } else {
            const float4 v = __ldg(reinterpret_cast<const float4*>(bias_kind == 1 ? bias + (i % N) : bias + i));
            bf[0] = v.x; bf[1] = v.y; bf[2] = v.z; bf[3] = v.w;
        }
        if (nbar_kind == 0) {
            nf[0] = nf[1] = nf[2] = nf[3] = nbar[0];
        } else {
            const float4 v = __ldg(reinterpret_cast<const float4*>(nbar_kind == 1 ? nbar + (i % N) : nbar + i));
            nf[0] = v.x; nf[1] = v.y; nf[2] = v.z; nf[3] = v.w;
        }
        double lam[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            double dh = __dmul_rn((double)bf[e], (double)d[e]);                 // halos.py:104
            if (lognormal) dh = __dadd_rn(__ddiv_rn(exp(dh), mean_exp), -1.0);  // halos.py:106-108
            double l = __dmul_rn(__dmul_rn(voxel_vol, (double)nf[e]), __dadd_rn(1.0, dh));   // halos.py:111
            if (!lognormal && l < 0.0) l = 0.0;                                 // halos.py:112-113
            if (l != l) l = 0.0;                                                // nan_to_num, halos.py:116
            lam[e] = l;
        }
        if (mean_out)
            reinterpret_cast<float4*>(mean_out)[q] = make_float4((float)lam[0], (float)lam[1], (float)lam[2], (float)lam[3]);
        if (counts) {
            const double2 u01 = __ldg(reinterpret_cast<const double2*>(uniforms) + 2 * q);
            const double2 u23 = __ldg(reinterpret_cast<const double2*>(uniforms) + 2 * q + 1);
            const double u[4] = {u01.x, u01.y, u23.x, u23.y};
            int32_t k[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) k[e] = fb_poisson_inv(lam[e], u[e]);
            reinterpret_cast<int4*>(counts)[q] = make_int4(k[0], k[1], k[2], k[3]);
        }
    }
}

// ---------------------------------------------------------------------------
// strided copy probe: tiles of `rows` x `chunk` bytes, rows `row_stride` apart
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_strided_copy(const uint4* __restrict__ src, uint4* __restrict__ dst, int rows,
                                                       int chunk16, size_t row_stride16, int chunks_per_row,
                                                       size_t plane_stride16) {
    const size_t tile = blockIdx.x;
    const size_t plane = tile / chunks_per_row, ch = tile % chunks_per_row;
    const size_t base = plane * plane_stride16 + ch * chunk16;
    const int total = rows * chunk16;
    for (int i0 = threadIdx.x; i0 < total; i0 += blockDim.x * 4) {
        uint4 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int i = i0 + u * blockDim.x;
            if (i < total) v[u] = src[base + (size_t)(i / chunk16) * row_stride16 + (i % chunk16)];
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int i = i0 + u * blockDim.x;
            if (i < total) dst[base + (size_t)(i / chunk16) * row_stride16 + (i % chunk16)] = v[u];
        }
    }
}

}  // namespace fb

using namespace fb;

extern "C" {

int fb_affine(fb_plan* p, float* field, size_t n, float mul, float add) {
    FB_CUDA(cudaSetDevice(p->device));
    FB_CHECK(is_device_ptr(field), "fb_affine: field must be device memory");
    k_affine<<<grid_for(n / 4 + 1, 256, p->sm_count), 256, 0, p->stream>>>(field, n, mul, add);
    FB_LAUNCH_CHECK();
    return 0;
}

int fb_exp_sum(fb_plan* p, const float* in, float* out, size_t n, float scale, double* sum_out) {
    FB_CUDA(cudaSetDevice(p->device));
    FB_CHECK(is_device_ptr(in) && is_device_ptr(out), "fb_exp_sum: buffers must be device memory");
    FB_CUDA(cudaMemsetAsync(p->scal, 0, 8 * sizeof(double), p->stream));
    k_exp_sum<<<grid_for(n, 256, p->sm_count), 256, 0, p->stream>>>(in, out, n, scale, p->scal);
    FB_LAUNCH_CHECK();
    FB_CUDA(cudaMemcpyAsync(p->scal_host, p->scal, 2 * sizeof(double), cudaMemcpyDeviceToHost, p->stream));
    FB_CUDA(cudaStreamSynchronize(p->stream));
    if (sum_out) *sum_out = p->scal_host[0];
    return 0;
}

int fb_field_moments(fb_plan* p, const float* field, size_t n, double* sum, double* sumsq) {
    FB_CUDA(cudaSetDevice(p->device));
    const void* d = nullptr;
    if (stage_in(p, 2, field, n * sizeof(float), &d)) return -2;
    FB_CUDA(cudaMemsetAsync(p->scal, 0, 8 * sizeof(double), p->stream));
    k_moments<<<grid_for(n, 256, p->sm_count), 256, 0, p->stream>>>((const float*)d, n, p->scal);
    FB_LAUNCH_CHECK();
    FB_CUDA(cudaMemcpyAsync(p->scal_host, p->scal, 2 * sizeof(double), cudaMemcpyDeviceToHost, p->stream));
    FB_CUDA(cudaStreamSynchronize(p->stream));
    if (sum) *sum = p->scal_host[0];
    if (sumsq) *sumsq = p->scal_host[1];
    return 0;
}

int fb_convert_f64_to_f32(fb_plan* p, const double* src, float* dst, size_t n) {
    FB_CUDA(cudaSetDevice(p->device));
    FB_CHECK(is_device_ptr(dst), "fb_convert_f64_to_f32: dst must be device memory");
    const void* d = nullptr;
    if (stage_in(p, 5, src, n * sizeof(double), &d)) return -2;
    k_f64_to_f32<<<grid_for(n, 256, p->sm_count), 256, 0, p->stream>>>((const double*)d, dst, n);
    FB_LAUNCH_CHECK();
    FB_CUDA(cudaStreamSynchronize(p->stream));
    return 0;
}

int fb_convert_f32_to_f64(fb_plan* p, const float* src, double* dst, size_t n) {
    FB_CUDA(cudaSetDevice(p->device));
    FB_CHECK(is_device_ptr(src), "fb_convert_f32_to_f64: src must be device memory");
    void* d = nullptr;
    if (stage_out_begin(p, 5, dst, n * sizeof(double), &d)) return -2;
    k_f32_to_f64<<<grid_for(n, 256, p->sm_count), 256, 0, p->stream>>>(src, (double*)d, n);
    FB_LAUNCH_CHECK();
    if (stage_out_end(p, 5, dst, n * sizeof(double))) return -2;
    FB_CUDA(cudaStreamSynchronize(p->stream));
    return 0;
}

// out = counts * mul + add  (halo overdensity N_h / N_bar - 1 for the cross spectrum, example_halos.py:46-53)
__global__ void __launch_bounds__(256) k_counts_to_field(const int32_t* __restrict__ c, float* __restrict__ out, size_t n,
                                                          float mul, float add) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    const size_t n4 = n / 4;
    for (size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x; q < n4; q += stride) {
        const int4 v = __ldg(reinterpret_cast<const int4*>(c) + q);
        reinterpret_cast<float4*>(out)[q] = make_float4(fmaf((float)v.x, mul, add), fmaf((float)v.y, mul, add),
                                                        fmaf((float)v.z, mul, add), fmaf((float)v.w, mul, add));
    }
    for (size_t j = n4 * 4 + (size_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += stride)
        out[j] = fmaf((float)c[j], mul, add);
}

int fb_counts_to_field(fb_plan* p, const int32_t* counts, size_t n, float mul, float add, float* out) {
    FB_CUDA(cudaSetDevice(p->device));
    FB_CHECK(is_device_ptr(counts) && is_device_ptr(out), "fb_counts_to_field: buffers must be device memory");
    k_counts_to_field<<<grid_for(n / 4 + 1, 256, p->sm_count), 256, 0, p->stream>>>(counts, out, n, mul, add);
    FB_LAUNCH_CHECK();
    return 0;
}

int fb_rsd_remap(fb_plan* p, const float* delta, const float* vel_z, const float* vel_nl, const double* zgrid,
                 double Hz, float* out) {
    return fb_rsd_remap_method(p, delta, vel_z, vel_nl, zgrid, Hz, FB_RSD_LINEAR, out);
}

int fb_rsd_remap_method(fb_plan* p, const float* delta, const float* vel_z, const float* vel_nl,
                        const double* zgrid, double Hz, int method, float* out) {
    FB_CUDA(cudaSetDevice(p->device));
    FB_CHECK(method == FB_RSD_LINEAR || method == FB_RSD_NEAREST,
             "fb_rsd_remap: method must be FB_RSD_LINEAR or FB_RSD_NEAREST");
    const int N = p->N;
    const size_t n3 = (size_t)N * N * N;
    FB_CHECK(delta && vel_z && zgrid && out, "fb_rsd_remap: NULL buffer");
    FB_CHECK(Hz > 0, "fb_rsd_remap: Hz must be positive");
    const void *dd = nullptr, *dv = nullptr, *dn = nullptr;
    void* dout = nullptr;
    if (stage_in(p, 0, delta, n3 * sizeof(float), &dd)) return -2;
    if (stage_in(p, 1, vel_z, n3 * sizeof(float), &dv)) return -2;
    if (stage_in(p, 4, vel_nl, n3 * sizeof(float), &dn)) return -2;
    if (stage_out_begin(p, 2, out, n3 * sizeof(float), &dout)) return -2;
    if (ensure_aux(p, (size_t)N * sizeof(double))) return -2;
    FB_CUDA(cudaMemcpyAsync(p->aux, zgrid, (size_t)N * sizeof(double), cudaMemcpyDefault, p->stream));
    const long nlines = (long)N * N;
    const unsigned ctas = (unsigned)((nlines + FB_RSD_LPC - 1) / FB_RSD_LPC);
#define FB_RSD(N_)                                                                                         \
    {                                                                                                      \
        auto kern = method == FB_RSD_NEAREST ? k_rsd_remap<N_, 1> : k_rsd_remap<N_, 0>;                    \
        if (set_smem(kern, RsdGeom<N_>::SMEM)) return -2;                                                  \
        kern<<<ctas, RsdGeom<N_>::NT, RsdGeom<N_>::SMEM, p->stream>>>(                                     \
            (const float*)dd, (const float*)dv, (const float*)dn, (const double*)p->aux, Hz, (float*)dout, \
            nlines);                                                                                       \
    }
    FB_DISPATCH_N(N, FB_RSD);
#undef FB_RSD
    FB_LAUNCH_CHECK();
    if (stage_out_end(p, 2, out, n3 * sizeof(float))) return -2;
    return 0;
}

int fb_halo_counts(fb_plan* p, const float* delta, const float* nbar, int nbar_kind, const float* bias, int bias_kind,
                   int lognormal, double mean_exp, const double* uniforms, int32_t* counts_out, float* mean_out) {
    FB_CUDA(cudaSetDevice(p->device));
    const int N = p->N;
    const size_t n3 = (size_t)N * N * N;
    FB_CHECK(delta && nbar && bias, "fb_halo_counts: NULL buffer");
    FB_CHECK(nbar_kind >= 0 && nbar_kind <= 2 && bias_kind >= 0 && bias_kind <= 2, "fb_halo_counts: bad kind");
    FB_CHECK((counts_out == nullptr) || (uniforms != nullptr), "fb_halo_counts: counts need uniforms");
    const size_t sz[3] = {1, (size_t)N, n3};
    const void *dd = nullptr, *dnb = nullptr, *dbi = nullptr, *du = nullptr;
    void *dc = nullptr, *dm = nullptr;
    if (stage_in(p, 0, delta, n3 * sizeof(float), &dd)) return -2;
    const size_t nb_slot = (sz[nbar_kind] + 3) & ~(size_t)3;                 // keep both tables 16-byte aligned
    if (ensure_aux(p, (nb_slot + sz[bias_kind]) * sizeof(float))) return -2;
    float* a_nb = (float*)p->aux;
    float* a_bi = a_nb + nb_slot;
    if (is_device_ptr(nbar)) dnb = nbar; else {
        FB_CUDA(cudaMemcpyAsync(a_nb, nbar, sz[nbar_kind] * sizeof(float), cudaMemcpyHostToDevice, p->stream));
        dnb = a_nb;
    }
    if (is_device_ptr(bias)) dbi = bias; else {
        FB_CUDA(cudaMemcpyAsync(a_bi, bias, sz[bias_kind] * sizeof(float), cudaMemcpyHostToDevice, p->stream));
        dbi = a_bi;
    }
    if (stage_in(p, 5, uniforms, n3 * sizeof(double), &du)) return -2;
    if (stage_out_begin(p, 1, counts_out, n3 * sizeof(int32_t), &dc)) return -2;
    if (stage_out_begin(p, 2, mean_out, n3 * sizeof(float), &dm)) return -2;
    auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
    FB_CHECK(al16(dd) && al16(du) && al16(dc) && al16(dm) && (nbar_kind == 0 || al16(dnb)) && (bias_kind == 0 || al16(dbi)),
             "fb_halo_counts: device buffers must be 16-byte aligned");
    if (lognormal && !(mean_exp > 0.0)) {
        FB_CUDA(cudaMemsetAsync(p->scal, 0, 8 * sizeof(double), p->stream));
        k_exp_sum_f64<<<grid_for(n3, 256, p->sm_count), 256, 0, p->stream>>>((const float*)dd, (const float*)dbi,
                                                                            bias_kind, N, n3, p->scal);
        FB_LAUNCH_CHECK();
        FB_CUDA(cudaMemcpyAsync(p->scal_host, p->scal, sizeof(double), cudaMemcpyDeviceToHost, p->stream));
        FB_CUDA(cudaStreamSynchronize(p->stream));
        mean_exp = p->scal_host[0] / (double)n3;
    }
    const double vol = p->Lx * p->Ly * p->Lz / pow((double)N, 3.0);        // halos.py:101
    k_halo_counts<<<grid_for(n3 / 4, 256, p->sm_count), 256, 0, p->stream>>>(
        (const float*)dd, (const float*)dnb, nbar_kind, (const float*)dbi, bias_kind, lognormal, mean_exp, vol,
        (const double*)du, N, n3, (int32_t*)dc, (float*)dm);
    FB_LAUNCH_CHECK();
    if (stage_out_end(p, 1, counts_out, n3 * sizeof(int32_t))) return -2;
    if (stage_out_end(p, 2, mean_out, n3 * sizeof(float))) return -2;
    return 0;
}

int fb_bench_strided_copy(fb_plan* p, size_t total_bytes, int chunk_bytes, int iters, double* gbs) {
    FB_CUDA(cudaSetDevice(p->device));
    FB_CHECK(chunk_bytes >= 16 && (chunk_bytes % 16) == 0, "chunk must be a multiple of 16 bytes");
    const int rows = 1024;
    const size_t row_bytes = 8192;                       // one z row of a 1024^3 complex64 cube
    const size_t plane_bytes = rows * row_bytes;
    const size_t planes = total_bytes / plane_bytes;
    FB_CHECK(planes >= 1, "total_bytes too small");
    const int chunks_per_row = (int)(row_bytes / chunk_bytes);
    void *src = nullptr, *dst = nullptr;
    FB_CUDA(cudaMalloc(&src, planes * plane_bytes));
    FB_CUDA(cudaMalloc(&dst, planes * plane_bytes));
    FB_CUDA(cudaMemsetAsync(src, 1, planes * plane_bytes, p->stream));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const unsigned grid = (unsigned)(planes * chunks_per_row);
    for (int it = 0; it < iters + 1; ++it) {
        if (it == 1) cudaEventRecord(e0, p->stream);
        k_strided_copy<<<grid, 256, 0, p->stream>>>((const uint4*)src, (uint4*)dst, rows, chunk_bytes / 16,
                                                    row_bytes / 16, chunks_per_row, plane_bytes / 16);
    }
    cudaEventRecord(e1, p->stream);
    FB_LAUNCH_CHECK();
    FB_CUDA(cudaStreamSynchronize(p->stream));
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    *gbs = 2.0 * (double)(planes * plane_bytes) * iters / (ms * 1e-3) / 1e9;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(src);
    cudaFree(dst);
    return 0;
}

}  // extern "C"
