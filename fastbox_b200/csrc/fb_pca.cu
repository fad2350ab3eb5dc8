// fb_pca.cu -- the device steps of filters.pca_filter (fastbox/filters.py:93-183), in FLOAT64:
//     d_mean = mean over pixels per channel                      (filters.py:142)
//     cov    = np.cov(d - d_mean)            [Nf x Nf]           (filters.py:158-159)
//     (eigen-decomposition of cov: host, Nf x Nf)                (filters.py:162-170)
//     fg_amps = U^T (d - d_mean), cleaned = field - (U fg_amps + d_mean)   (filters.py:173-178)
// Why float64: with foregrounds 1e4-1e5 x the signal a float32 cube perturbs the data by
// ~1e-2 of the signal, and an error eps in the covariance leaks a foreground mode as eps*lambda_1/sqrt(lambda_i).
// The cube is [pixel][channel] (channel contiguous), exactly field.reshape(-1, Nf) of the reference.
//
// Covariance = tall-skinny X^T X, 2 * Nf^2/2 * Npix flops: k_pca_cov_mma (from 256 channels on: 128 x 128 tiles of the
// upper triangle x pixel slices on the FP64 tensor path, see there) and k_pca_cov (fewer or an odd number of channels:
// 64 x 64 tiles, 256 threads with 4 x 4 float64 accumulators each, panels of 16 pixels x 64 channels through shared
// memory, mean subtracted on the way in); partial tiles are added to the global matrix with float64 reductions.
// k_pca_project_warp: one warp per line of sight, row in registers, operator in shared memory (nmodes <= 8);
// k_pca_project: general fallback, one CTA per line of sight with block reductions.
#include "fb_launch.h"

namespace fb {

constexpr int PCA_TILE = 64, PCA_KT = 16;


__global__ void __launch_bounds__(256) k_pca_sums(const double* __restrict__ x, int nf, size_t npix,
                                                   double* __restrict__ sums) {
    const size_t rows_per_cta = (npix + gridDim.x - 1) / gridDim.x;
    const size_t r0 = (size_t)blockIdx.x * rows_per_cta;
    const size_t r1 = r0 + rows_per_cta < npix ? r0 + rows_per_cta : npix;
    for (int z = threadIdx.x; z < nf; z += 256) {
        double a0 = 0.0, a1 = 0.0;
        size_t r = r0;
        for (; r + 1 < r1; r += 2) {
            a0 += __ldg(&x[r * nf + z]);
            a1 += __ldg(&x[(r + 1) * nf + z]);
        }
        if (r < r1) a0 += __ldg(&x[r * nf + z]);
        if (r1 > r0) atomicAdd(&sums[z], a0 + a1);
    }
}

__global__ void k_pca_scale(double* __restrict__ v, int n, double s) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) v[i] *= s;
}

// grid = (upper-triangle tiles, pixel slices)
__global__ void __launch_bounds__(256) k_pca_cov(const double* __restrict__ x, const double* __restrict__ mean, int nf,
                                                  size_t npix, int ntile, double* __restrict__ cov) {
    __shared__ double sa[PCA_KT][PCA_TILE + 2], sb[PCA_KT][PCA_TILE + 2];
    // tile index -> (ti <= tj)
    int ti = 0, rem = blockIdx.x;
    while (rem >= ntile - ti) {
        rem -= ntile - ti;
        ++ti;
    }
    const int tj = ti + rem;
    const int f0 = ti * PCA_TILE, g0 = tj * PCA_TILE;
    const size_t per_slice = (npix + gridDim.y - 1) / gridDim.y;
    const size_t p0 = (size_t)blockIdx.y * per_slice;
    const size_t p1 = p0 + per_slice < npix ? p0 + per_slice : npix;
    const int tx = threadIdx.x % 16, ty = threadIdx.x / 16;           // 16 x 16 threads, 4 x 4 outputs each
    const int lc = threadIdx.x % PCA_TILE, lr = threadIdx.x / PCA_TILE;   // loader: 64 channels x 4 pixel rows
    double acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;
    const double ma = (f0 + lc) < nf ? __ldg(&mean[f0 + lc]) : 0.0;
    const double mb = (g0 + lc) < nf ? __ldg(&mean[g0 + lc]) : 0.0;
    for (size_t pb = p0; pb < p1; pb += PCA_KT) {
#pragma unroll
        for (int k = lr; k < PCA_KT; k += 4) {
            const size_t p = pb + k;
            double va = 0.0, vb = 0.0;
            if (p < p1) {
                if (f0 + lc < nf) va = __ldg(&x[p * nf + f0 + lc]) - ma;
                if (g0 + lc < nf) vb = __ldg(&x[p * nf + g0 + lc]) - mb;
            }
            sa[k][lc] = va;
            sb[k][lc] = vb;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < PCA_KT; ++k) {
            double a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                a[i] = sa[k][ty + 16 * i];
                b[i] = sb[k][tx + 16 * i];
            }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fma(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int f = f0 + ty + 16 * i, g = g0 + tx + 16 * j;
            if (f < nf && g < nf) atomicAdd(&cov[(size_t)f * nf + g], acc[i][j]);
        }
}

// D(8 x 8) += A(8 x 4) B(4 x 8) on the FP64 tensor path.  tcgen05 has no float64 kind, so mma.sync m8n8k4 is the only
// tensor-core route for this contraction.  Lane l holds A[l >> 2][l & 3], B[l & 3][l >> 2] and D[l >> 2][2 (l & 3) + {0, 1}].
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

// k_pca_cov_mma: X^T X on 128 x 128 tiles of the upper triangle x pixel slices, on the FP64 tensor path (used from
// 256 channels on).  Eight warps in a 2 (f) x 4 (g) grid own 64 x 32 outputs each = 8 x 4 MMA tiles (64 accumulator
// registers per lane); per group of four pixels a warp reads 8 + 4 operand fragments (one float64 per lane each) for
// 32 MMAs = 8192 FMAs.  A fragment A[f][p] = X[p][f] is the element (pixel 4 kg + (l & 3), channel f0 + (l >> 2)) of
// the panel, B the same of the g panel: rows are padded to 136 float64 so that the four pixel rows of a fragment
// fall on the two halves of the banks (a 256-byte warp request in its minimum of two wavefronts).
// Panels of KT pixels x 128 channels are double buffered in shared memory; the next panel is fetched into registers
// while the current one is consumed and its channel means are subtracted only when it is parked, so that nothing
// ahead of the MMA block waits on a load in flight (an earlier build subtracted right after the loads: every warp then
// stalled on its prefetch first, 75 ms instead of 50); one barrier per panel; the grid is nine whole waves of one CTA
// per SM; diagonal tiles skip their lower-left 64 x 64 quadrant (never read by k_pca_cov_finish).  nf must be even
// (16-byte loads).  History (profiles/README.md): the same tile with 8 x 8 SIMT accumulators per thread read four
// times the shared-memory words per FMA and kept that pipe about as busy as the FP64 pipe:
// 50.5 ms at 1024^3 against 43.3 ms here; the 64 x 64 kernel above: 77 ms.
template <int KT>
__global__ void __launch_bounds__(256, 1) k_pca_cov_mma(const double* __restrict__ x, const double* __restrict__ mean,
                                                         int nf, size_t npix, int ntile, double* __restrict__ cov) {
    constexpr int TL = 128, LD = 136;
    constexpr int RPT = KT / 4;
    extern __shared__ __align__(16) unsigned char pca_cov_smem[];
    double (*sa)[KT][LD] = reinterpret_cast<double (*)[KT][LD]>(pca_cov_smem);                 // [2][KT][LD]
    double (*sb)[KT][LD] = reinterpret_cast<double (*)[KT][LD]>(pca_cov_smem + 2 * sizeof(double) * KT * LD);
    int ti = 0, rem = blockIdx.x;
    while (rem >= ntile - ti) {
        rem -= ntile - ti;
        ++ti;
    }
    const int tj = ti + rem;
    const int f0 = ti * TL, g0 = tj * TL;
    const size_t per_slice = ((npix + gridDim.y - 1) / gridDim.y + KT - 1) / KT * KT;
    const size_t p0 = (size_t)blockIdx.y * per_slice;
    const size_t p1 = p0 + per_slice < npix ? p0 + per_slice : npix;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int wf = warp & 1, wg = warp >> 1;                         // warp tile: f in [64 wf, +64), g in [32 wg, +32)
    const int fr = lane >> 2, kc = lane & 3;
    const bool active = !(ti == tj && wf == 1 && wg < 2);            // lower-left quadrant of a diagonal tile
    const int lc = 2 * (threadIdx.x & 63), lr = threadIdx.x >> 6;    // loader: channel pair, row 0..3
    double acc[8][4][2];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
    const bool oka = f0 + lc < nf, okb = g0 + lc < nf;
    double2 ma = make_double2(0.0, 0.0), mb = ma;
    if (oka) ma = *reinterpret_cast<const double2*>(mean + f0 + lc);
    if (okb) mb = *reinterpret_cast<const double2*>(mean + g0 + lc);
    double2 ra[RPT], rb[RPT];
    auto fetch = [&](size_t pb) {
#pragma unroll
        for (int r = 0; r < RPT; ++r) {
            const size_t p = pb + lr + 4 * r;
            ra[r] = ma;
            rb[r] = mb;
            if (p < p1) {
                const double* row = x + p * (size_t)nf;
                if (oka) ra[r] = __ldg(reinterpret_cast<const double2*>(row + f0 + lc));
                if (okb) rb[r] = __ldg(reinterpret_cast<const double2*>(row + g0 + lc));
            }
        }
    };
    auto park = [&](int b) {                                         // means subtracted here, not in fetch()
#pragma unroll
        for (int r = 0; r < RPT; ++r) {
            *reinterpret_cast<double2*>(&sa[b][lr + 4 * r][lc]) = make_double2(ra[r].x - ma.x, ra[r].y - ma.y);
            *reinterpret_cast<double2*>(&sb[b][lr + 4 * r][lc]) = make_double2(rb[r].x - mb.x, rb[r].y - mb.y);
        }
    };
    if (p0 < p1) {
        fetch(p0);
        park(0);
    }
    __syncthreads();
    int cur = 0;
    for (size_t pb = p0; pb < p1; pb += KT, cur ^= 1) {
        const bool more = pb + KT < p1;
        if (more) fetch(pb + KT);
        if (active) {
#pragma unroll
            for (int kg = 0; kg < KT / 4; ++kg) {
                const double* pa = &sa[cur][4 * kg + kc][64 * wf + fr];
                const double* pb2 = &sb[cur][4 * kg + kc][32 * wg + fr];
                double a[8], b[4];
#pragma unroll
                for (int i = 0; i < 8; ++i) a[i] = pa[8 * i];
#pragma unroll
                for (int j = 0; j < 4; ++j) b[j] = pb2[8 * j];
#pragma unroll
                for (int i = 0; i < 8; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
            }
        }
        if (more) park(cur ^ 1);
        __syncthreads();
    }
    if (active) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j)
#pragma unroll
                for (int v = 0; v < 2; ++v) {
                    const int f = f0 + 64 * wf + 8 * i + fr, g = g0 + 32 * wg + 8 * j + 2 * kc + v;
                    if (f < nf && g < nf) atomicAdd(&cov[(size_t)f * nf + g], acc[i][j][v]);
                }
    }
}

// mirror the upper-triangle tiles into the lower triangle and apply the 1/(npix-1) of np.cov
__global__ void k_pca_cov_finish(double* __restrict__ cov, int nf, double norm) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x, f = blockIdx.y;
    if (g >= nf) return;
    if (f <= g) {                                        // element-wise: the result is exactly symmetric even though
        const double v = cov[(size_t)f * nf + g] * norm;  // the slices' reductions arrive in any order
        cov[(size_t)f * nf + g] = v;
        if (f < g) cov[(size_t)g * nf + f] = v;
    }
}

// one CTA (256 threads) per pixel row: amplitudes by block reduction, then the residual
template <int MAXM>
__global__ void __launch_bounds__(256) k_pca_project(const double* __restrict__ x, const double* __restrict__ mean,
                                                      const double* __restrict__ U, int nf, int nmodes, size_t npix,
                                                      double* __restrict__ cleaned, double* __restrict__ amps) {
    __shared__ double red[8][MAXM];
    __shared__ double a_sh[MAXM];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (size_t p = blockIdx.x; p < npix; p += gridDim.x) {
        double part[MAXM];
#pragma unroll
        for (int m = 0; m < MAXM; ++m) part[m] = 0.0;
        for (int f = tid; f < nf; f += 256) {
            const double d = __ldg(&x[p * nf + f]) - __ldg(&mean[f]);
#pragma unroll
            for (int m = 0; m < MAXM; ++m)
                if (m < nmodes) part[m] = fma(__ldg(&U[(size_t)f * nmodes + m]), d, part[m]);
        }
#pragma unroll
        for (int m = 0; m < MAXM; ++m) {
            if (m < nmodes) {
                double v = part[m];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                if (lane == 0) red[warp][m] = v;
            }
        }
        __syncthreads();
        if (tid < nmodes) {
            double v = 0.0;
#pragma unroll
            for (int w = 0; w < 8; ++w) v += red[w][tid];
            a_sh[tid] = v;
            if (amps) amps[(size_t)tid * npix + p] = v;              // (Nmodes, Npix), filters.py:173
        }
        __syncthreads();
        for (int f = tid; f < nf; f += 256) {
            double fg = 0.0;
#pragma unroll
            for (int m = 0; m < MAXM; ++m)
                if (m < nmodes) fg = fma(__ldg(&U[(size_t)f * nmodes + m]), a_sh[m], fg);
            // cleaned = field - (U amps + mean), filters.py:176-178
            cleaned[p * nf + f] = __ldg(&x[p * nf + f]) - (fg + __ldg(&mean[f]));
        }
        __syncthreads();
    }
}

// Streaming variant for nmodes <= 8 and nf >= 32: one WARP per line of sight.  The row stays in
// registers (nf/32 float64 per lane) between the amplitude sums and the residual, the amplitudes need only
// warp shuffles, and the filter operator sits in shared memory as [mode][channel] (conflict-free).
template <int NF>
__global__ void __launch_bounds__(256) k_pca_project_warp(const double* __restrict__ x, const double* __restrict__ mean,
                                                           const double* __restrict__ U, int nmodes, size_t npix,
                                                           double* __restrict__ cleaned, double* __restrict__ amps) {
    constexpr int E = NF / 32;
    extern __shared__ __align__(16) unsigned char pca_smem[];
    double* us = reinterpret_cast<double*>(pca_smem);            // [nmodes][NF]
    double* ms = us + (size_t)nmodes * NF;                       // [NF]
    for (int i = threadIdx.x; i < nmodes * NF; i += blockDim.x) {
        const int m = i / NF, f = i - m * NF;
        us[i] = __ldg(&U[(size_t)f * nmodes + m]);
    }
    for (int f = threadIdx.x; f < NF; f += blockDim.x) ms[f] = __ldg(&mean[f]);
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const size_t warp0 = (size_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const size_t nwarps = (size_t)gridDim.x * (blockDim.x >> 5);
    for (size_t p = warp0; p < npix; p += nwarps) {
        double v[E];
#pragma unroll
        for (int i = 0; i < E; ++i) v[i] = __ldg(&x[p * NF + lane + 32 * i]);
        double a[8];
#pragma unroll
        for (int m = 0; m < 8; ++m) {
            a[m] = 0.0;
            if (m < nmodes) {
                double s = 0.0;
#pragma unroll
                for (int i = 0; i < E; ++i) s = fma(us[m * NF + lane + 32 * i], v[i] - ms[lane + 32 * i], s);
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
                a[m] = s;
                if (amps && lane == 0) amps[(size_t)m * npix + p] = s;      // (Nmodes, Npix), filters.py:173
            }
        }
#pragma unroll
        for (int i = 0; i < E; ++i) {
            double fg = 0.0;
#pragma unroll
            for (int m = 0; m < 8; ++m)
                if (m < nmodes) fg = fma(us[m * NF + lane + 32 * i], a[m], fg);
            cleaned[p * NF + lane + 32 * i] = v[i] - (fg + ms[lane + 32 * i]);   // filters.py:176-178
        }
    }
}

template <int NF>
static int launch_project_warp(fb_plan* p, const double* cube, const double* d_mean, const double* d_U, int nmodes,
                               size_t npix, double* cleaned, double* amps) {
    auto kern = k_pca_project_warp<NF>;
    const size_t smem = ((size_t)nmodes * NF + NF) * sizeof(double);
    if (set_smem(kern, smem)) return -2;
    const size_t want = (npix + 7) / 8, cap = (size_t)p->sm_count * 4;
    kern<<<(unsigned)(want < cap ? want : cap), 256, smem, p->stream>>>(cube, d_mean, d_U, nmodes, npix, cleaned, amps);
    return 0;
}

// FP64 throughput probe (the denominator for the covariance kernel's roofline): independent FMA chains,
// no memory traffic.  mode 0: DFMA, mode 1: DMMA m8n8k4.
__global__ void __launch_bounds__(256) k_fp64_probe(double* out, int iters, int mode) {
    double a = 1.0 + 1e-9 * threadIdx.x, b = 1.0 - 1e-9 * threadIdx.x;
    double c[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) c[i] = (double)i;
    if (mode == 0) {
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < 16; ++i) c[i] = fma(a, c[i], b);
        }
    } else {
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < 16; i += 2) dmma884(c[i], c[i + 1], a, b);
        }
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += c[i];
    if (s == 123.456) out[0] = s;                        // keep the chains alive
}

}  // namespace fb

using namespace fb;

extern "C" {

int fb_pca_covariance(fb_plan* p, const double* cube, double* mean_out, double* cov_out) {
    FB_CUDA(cudaSetDevice(p->device));
    const int nf = p->N;
    const size_t npix = (size_t)nf * nf;
    FB_CHECK(cube && cov_out, "fb_pca_covariance: NULL buffer");
    FB_CHECK(is_device_ptr(cube), "fb_pca_covariance: the float64 cube must be device memory");
    if (ensure_aux(p, ((size_t)nf * nf + nf) * sizeof(double))) return -2;
    double* d_cov = (double*)p->aux;
    double* d_mean = d_cov + (size_t)nf * nf;
    {   // np.cov centres every channel on its own mean, whatever was subtracted before (filters.py:158-159)
        FB_CUDA(cudaMemsetAsync(d_mean, 0, nf * sizeof(double), p->stream));
        const unsigned grid = (unsigned)((size_t)p->sm_count * 8 < npix ? (size_t)p->sm_count * 8 : npix);
        k_pca_sums<<<grid, 256, 0, p->stream>>>(cube, nf, npix, d_mean);
        FB_LAUNCH_CHECK();
        k_pca_scale<<<(nf + 255) / 256, 256, 0, p->stream>>>(d_mean, nf, 1.0 / (double)npix);
        FB_LAUNCH_CHECK();
    }
    FB_CUDA(cudaMemsetAsync(d_cov, 0, (size_t)nf * nf * sizeof(double), p->stream));
    // 128 x 128 tiles on the FP64 tensor path (k_pca_cov_mma) from 256 channels on; FB_PCA_TILE=64 keeps the 64 x 64 SIMT
    // kernel, FB_PCA_KT sets the panel depth
    const int tile_opt = env_int("FB_PCA_TILE", nf >= 256 && nf % 2 == 0 ? 128 : 64);
    if (tile_opt == 128 && nf % 2 == 0) {
        const int kt = env_int("FB_PCA_KT", 16);
        FB_CHECK(kt == 8 || kt == 16, "FB_PCA_KT must be 8 or 16");
        const int ntile = (nf + 127) / 128;
        const int ntri = ntile * (ntile + 1) / 2;
        // one resident CTA per SM: a whole number of waves (the diagonal tiles finish a quarter earlier)
        int waves = env_int("FB_PCA_WAVES", 9);
        int slices = (p->sm_count * waves + ntri / 2) / ntri;
        const size_t max_slices = (npix + 4 * kt - 1) / (4 * kt);
        if ((size_t)slices > max_slices) slices = (int)max_slices;
        if (slices < 1) slices = 1;
        const size_t smem = 4 * sizeof(double) * kt * 136;
        if (kt == 8) {
            auto kern = k_pca_cov_mma<8>;
            if (set_smem(kern, smem)) return -2;
            kern<<<dim3(ntri, slices), 256, smem, p->stream>>>(cube, d_mean, nf, npix, ntile, d_cov);
        } else {
            auto kern = k_pca_cov_mma<16>;
            if (set_smem(kern, smem)) return -2;
            kern<<<dim3(ntri, slices), 256, smem, p->stream>>>(cube, d_mean, nf, npix, ntile, d_cov);
        }
        FB_LAUNCH_CHECK();
    } else {
        const int ntile = (nf + PCA_TILE - 1) / PCA_TILE;
        const int ntri = ntile * (ntile + 1) / 2;
        // ~4 full waves at 3 resident CTAs per SM (a grid of 1.5 waves left the second one half empty)
        int slices = (p->sm_count * 12 + ntri / 2) / ntri;
        const size_t max_slices = (npix + PCA_KT - 1) / PCA_KT;
        if ((size_t)slices > max_slices) slices = (int)max_slices;
        if (slices < 1) slices = 1;
        k_pca_cov<<<dim3(ntri, slices), 256, 0, p->stream>>>(cube, d_mean, nf, npix, ntile, d_cov);
        FB_LAUNCH_CHECK();
    }
    k_pca_cov_finish<<<dim3((nf + 255) / 256, nf), 256, 0, p->stream>>>(d_cov, nf, 1.0 / (double)(npix - 1));
    FB_LAUNCH_CHECK();
    FB_CUDA(cudaMemcpyAsync(cov_out, d_cov, (size_t)nf * nf * sizeof(double), cudaMemcpyDefault, p->stream));
    if (mean_out) FB_CUDA(cudaMemcpyAsync(mean_out, d_mean, nf * sizeof(double), cudaMemcpyDefault, p->stream));
    FB_CUDA(cudaStreamSynchronize(p->stream));
    return 0;
}

int fb_bench_fp64(fb_plan* p, int mode, double* tflops) {
    FB_CUDA(cudaSetDevice(p->device));
    FB_CHECK(tflops, "fb_bench_fp64: NULL output");
    if (ensure_aux(p, 64)) return -2;
    const int iters = 20000, ctas = p->sm_count * 8;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    k_fp64_probe<<<ctas, 256, 0, p->stream>>>((double*)p->aux, 100, mode);
    cudaEventRecord(e0, p->stream);
    k_fp64_probe<<<ctas, 256, 0, p->stream>>>((double*)p->aux, iters, mode);
    cudaEventRecord(e1, p->stream);
    FB_LAUNCH_CHECK();
    FB_CUDA(cudaStreamSynchronize(p->stream));
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    // mode 0: 16 FMA per thread and iteration; mode 1: 8 MMAs of 8x8x4 = 256 FMA per warp -> 64 FMA per thread
    const double fma_per_thread = mode == 0 ? 16.0 : 64.0;
    *tflops = 2.0 * fma_per_thread * iters * 256.0 * ctas / (ms * 1e-3) / 1e12;
    return 0;
}

int fb_pca_project(fb_plan* p, const double* cube, const double* mean, const double* U, int nmodes, double* cleaned,
                   double* amps) {
    FB_CUDA(cudaSetDevice(p->device));
    const int nf = p->N;
    const size_t npix = (size_t)nf * nf;
    FB_CHECK(cube && mean && U && cleaned, "fb_pca_project: NULL buffer");
    FB_CHECK(is_device_ptr(cube) && is_device_ptr(cleaned) && (!amps || is_device_ptr(amps)),
             "fb_pca_project: cube / cleaned / amps must be device memory");
    FB_CHECK(nmodes >= 1 && nmodes <= 32, "fb_pca_project: nmodes=%d out of range [1,32]", nmodes);
    if (ensure_aux(p, ((size_t)nf * nmodes + nf) * sizeof(double))) return -2;
    double* d_U = (double*)p->aux;
    double* d_mean = d_U + (size_t)nf * nmodes;
    FB_CUDA(cudaMemcpyAsync(d_U, U, (size_t)nf * nmodes * sizeof(double), cudaMemcpyDefault, p->stream));
    FB_CUDA(cudaMemcpyAsync(d_mean, mean, nf * sizeof(double), cudaMemcpyDefault, p->stream));
    const unsigned grid = (unsigned)((size_t)p->sm_count * 8 < npix ? (size_t)p->sm_count * 8 : npix);
    if (nmodes <= 8 && nf >= 32 && nf <= 1024) {
        int rc = 0;
        switch (nf) {
            case 32: rc = launch_project_warp<32>(p, cube, d_mean, d_U, nmodes, npix, cleaned, amps); break;
            case 64: rc = launch_project_warp<64>(p, cube, d_mean, d_U, nmodes, npix, cleaned, amps); break;
            case 128: rc = launch_project_warp<128>(p, cube, d_mean, d_U, nmodes, npix, cleaned, amps); break;
            case 256: rc = launch_project_warp<256>(p, cube, d_mean, d_U, nmodes, npix, cleaned, amps); break;
            case 512: rc = launch_project_warp<512>(p, cube, d_mean, d_U, nmodes, npix, cleaned, amps); break;
            default: rc = launch_project_warp<1024>(p, cube, d_mean, d_U, nmodes, npix, cleaned, amps); break;
        }
        if (rc) return rc;
    } else if (nmodes <= 8)
        k_pca_project<8><<<grid, 256, 0, p->stream>>>(cube, d_mean, d_U, nf, nmodes, npix, cleaned, amps);
    else
        k_pca_project<32><<<grid, 256, 0, p->stream>>>(cube, d_mean, d_U, nf, nmodes, npix, cleaned, amps);
    FB_LAUNCH_CHECK();
    FB_CUDA(cudaStreamSynchronize(p->stream));
    return 0;
}

}  // extern "C"
