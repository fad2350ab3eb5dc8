// fb_cube.cu -- the N^3 element-wise steps either side of the beam convolution in the end-to-end
// data cube (examples/example_endtoend.py:58-87):
//   ForegroundModel.construct_cube   (fastbox/foregrounds.py:152-174)
//       cube[x,y,z] = amps[x,y] * (freqs[z] / freq_ref) ** spectral_idx[x,y]
//   NoiseModel.realise_radiometer_noise   (fastbox/noise.py:55-75)
//       noise[x,y,z] = sigma_rms[z] * n[x,y,z],   n ~ N(0,1)
// Both are streaming writes of the float32 cube (4 B/cell; +4 when accumulating into an existing cube,
// +4 when the unit normals are supplied instead of drawn by Philox).  The 2-D maps (amplitude, spectral
// index: N^2 values) and the per-channel tables (N values) are prepared by the host shim.
#include "fb_launch.h"

namespace fb {

static inline unsigned grid_rows(size_t nrows, int rows_per_cta, int sm_count) {
    const size_t want = (nrows + rows_per_cta - 1) / rows_per_cta;
    const size_t cap = (size_t)sm_count * 32;
    return (unsigned)(want < cap ? (want ? want : 1) : cap);
}

// one (x,y) row of N channels per thread group; 4 channels per thread (16-byte accesses).
// pow(f, s) = exp2(s * log2 f): log2(freqs/freq_ref) comes as a float32 table computed in float64 on
// the host; |s log2 f| ~ 10 keeps the relative error near 1e-6 (the cube is float32 anyway).
template <bool ACC>
__global__ void __launch_bounds__(256) k_fg_cube(const float* __restrict__ amps, const float* __restrict__ idx,
                                                  int idx_is_map, const float* __restrict__ log2f, int N, size_t nrows,
                                                  float* __restrict__ out) {
    const int per_row = N / 4;                           // threads per row
    const int rows_per_cta = 256 / per_row > 0 ? 256 / per_row : 1;
    const int tid = threadIdx.x;
    for (size_t row0 = (size_t)blockIdx.x * rows_per_cta; row0 < nrows; row0 += (size_t)gridDim.x * rows_per_cta) {
        for (int w = tid; w < rows_per_cta * per_row; w += 256) {
            const size_t row = row0 + w / per_row;
            const int z4 = w % per_row;
            if (row >= nrows) continue;
            const float a = __ldg(&amps[row]);
            const float s = idx_is_map ? __ldg(&idx[row]) : __ldg(&idx[0]);
            const float4 l = __ldg(reinterpret_cast<const float4*>(log2f) + z4);
            float4 r = make_float4(a * exp2f(s * l.x), a * exp2f(s * l.y), a * exp2f(s * l.z), a * exp2f(s * l.w));
            float4* dst = reinterpret_cast<float4*>(out + row * N) + z4;
            if (ACC) {
                const float4 o = *dst;
                r = make_float4(o.x + r.x, o.y + r.y, o.z + r.z, o.w + r.w);
            }
            *dst = r;
        }
    }
}

// Four consecutive channels per thread.  Philox stream of the noise cube: block counter =
// (cell index / 4, stream tag), key = seed; the block's two Box-Muller pairs are the four normals.
template <bool ACC>
__global__ void __launch_bounds__(256) k_radiometer_noise(const float* __restrict__ sigma,
                                                           const float* __restrict__ normals, uint64_t seed, int N,
                                                           size_t n4, float* __restrict__ out) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    const int per_row = N / 4;
    for (size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x; q < n4; q += stride) {
        const float4 sg = __ldg(reinterpret_cast<const float4*>(sigma) + (q % per_row));
        float4 nv;
        if (normals) {
            nv = __ldg(reinterpret_cast<const float4*>(normals) + q);
        } else {
            uint32_t c[4] = {(uint32_t)q, (uint32_t)(q >> 32), 0x4e4f4953u, 0u};       // tag "NOIS"
            uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
            for (int r = 0; r < 10; ++r) {
                philox_round(c, k0, k1);
                k0 += 0x9E3779B9u;
                k1 += 0xBB67AE85u;
            }
            const float2 g0 = box_muller(c[0], c[1]), g1 = box_muller(c[2], c[3]);
            nv = make_float4(g0.x, g0.y, g1.x, g1.y);
        }
        float4 r = make_float4(sg.x * nv.x, sg.y * nv.y, sg.z * nv.z, sg.w * nv.w);
        float4* dst = reinterpret_cast<float4*>(out) + q;
        if (ACC) {
            const float4 o = *dst;
            r = make_float4(o.x + r.x, o.y + r.y, o.z + r.z, o.w + r.w);
        }
        *dst = r;
    }
}

// ---- mean_spectrum_filter (fastbox/filters.py:35-55): per-channel mean over the N^2 pixels, then
// subtract.  Sums in float64 (one float64 reduction per channel and CTA); 4 B/cell + 8 B/cell.
__global__ void __launch_bounds__(256) k_channel_sums(const float* __restrict__ field, int N, size_t nrows,
                                                       double* __restrict__ sums) {
    // thread -> channel z = tid (+256 j); a CTA walks a contiguous block of pixel rows
    const size_t rows_per_cta = (nrows + gridDim.x - 1) / gridDim.x;
    const size_t r0 = (size_t)blockIdx.x * rows_per_cta;
    const size_t r1 = r0 + rows_per_cta < nrows ? r0 + rows_per_cta : nrows;
    for (int z = threadIdx.x; z < N; z += 256) {
        double acc0 = 0.0, acc1 = 0.0;                    // float64 throughout: a large monopole must not cost digits
        size_t r = r0;
        for (; r + 1 < r1; r += 2) {
            acc0 += (double)__ldg(&field[r * N + z]);
            acc1 += (double)__ldg(&field[(r + 1) * N + z]);
        }
        if (r < r1) acc0 += (double)__ldg(&field[r * N + z]);
        if (r1 > r0) atomicAdd(&sums[z], acc0 + acc1);
    }
}

// the subtraction is done in float64 (the residual of a cube with a large monopole is then the correctly
// rounded float32 of the exact difference); the kernel stays bound by its 8 B/cell of traffic
__global__ void __launch_bounds__(256) k_sub_channel(const float* __restrict__ field, const double* __restrict__ mean,
                                                      int N, size_t n4, float* __restrict__ out) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    const int per_row = N / 4;
    for (size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x; q < n4; q += stride) {
        const double2* m2 = reinterpret_cast<const double2*>(mean) + 2 * (q % per_row);
        const double2 ma = __ldg(m2), mb = __ldg(m2 + 1);
        const float4 v = __ldg(reinterpret_cast<const float4*>(field) + q);
        reinterpret_cast<float4*>(out)[q] = make_float4((float)((double)v.x - ma.x), (float)((double)v.y - ma.y),
                                                        (float)((double)v.z - mb.x), (float)((double)v.w - mb.y));
    }
}

__global__ void k_mean_from_sums(const double* __restrict__ sums, double inv_n, int N, double* __restrict__ mean64,
                                 float* __restrict__ mean32) {
    const int z = blockIdx.x * blockDim.x + threadIdx.x;
    if (z < N) {
        const double m = sums[z] * inv_n;
        mean64[z] = m;
        mean32[z] = (float)m;
    }
}

}  // namespace fb

using namespace fb;

extern "C" {

int fb_fg_cube(fb_plan* p, const float* amps, const float* spectral_idx, int idx_is_map, const float* log2_freq_ratio,
               float* out, int accumulate) {
    FB_CUDA(cudaSetDevice(p->device));
    const int N = p->N;
    FB_CHECK(amps && spectral_idx && log2_freq_ratio && out, "fb_fg_cube: NULL buffer");
    const size_t n2 = (size_t)N * N, n3 = n2 * N;
    // small inputs live in the aux buffer: amps [N^2], idx [N^2 or 1], log2 table [N]; 16-byte aligned slots
    const size_t idx_n = idx_is_map ? n2 : 4;
    if (ensure_aux(p, (2 * n2 + 4 + N) * sizeof(float))) return -2;
    float* d_amps = (float*)p->aux;
    float* d_idx = d_amps + n2;
    float* d_l2 = d_idx + (idx_is_map ? n2 : 4);
    FB_CUDA(cudaMemcpyAsync(d_amps, amps, n2 * sizeof(float), cudaMemcpyDefault, p->stream));
    FB_CUDA(cudaMemcpyAsync(d_idx, spectral_idx, (idx_is_map ? idx_n : 1) * sizeof(float), cudaMemcpyDefault, p->stream));
    FB_CUDA(cudaMemcpyAsync(d_l2, log2_freq_ratio, (size_t)N * sizeof(float), cudaMemcpyDefault, p->stream));
    void* dout = nullptr;
    if (accumulate) {
        const void* din = nullptr;
        if (stage_in(p, 2, out, n3 * sizeof(float), &din)) return -2;      // host cube: upload, add, download
        dout = const_cast<void*>(din);
    } else if (stage_out_begin(p, 2, out, n3 * sizeof(float), &dout)) {
        return -2;
    }
    const int per_row = N / 4, rows_per_cta = 256 / per_row > 0 ? 256 / per_row : 1;
    const unsigned grid = grid_rows(n2, rows_per_cta, p->sm_count);
    if (accumulate)
        k_fg_cube<true><<<grid, 256, 0, p->stream>>>(d_amps, d_idx, idx_is_map, d_l2, N, n2, (float*)dout);
    else
        k_fg_cube<false><<<grid, 256, 0, p->stream>>>(d_amps, d_idx, idx_is_map, d_l2, N, n2, (float*)dout);
    FB_LAUNCH_CHECK();
    if (stage_out_end(p, 2, out, n3 * sizeof(float))) return -2;
    return 0;
}

int fb_radiometer_noise(fb_plan* p, const float* sigma_z, const float* normals, uint64_t seed, float* out,
                        int accumulate) {
    FB_CUDA(cudaSetDevice(p->device));
    const int N = p->N;
    FB_CHECK(sigma_z && out, "fb_radiometer_noise: NULL buffer");
    const size_t n3 = (size_t)N * N * N;
    if (ensure_aux(p, (size_t)N * sizeof(float))) return -2;
    float* d_sigma = (float*)p->aux;
    FB_CUDA(cudaMemcpyAsync(d_sigma, sigma_z, (size_t)N * sizeof(float), cudaMemcpyDefault, p->stream));
    const void* dn = nullptr;
    if (stage_in(p, 0, normals, n3 * sizeof(float), &dn)) return -2;
    void* dout = nullptr;
    if (accumulate) {
        const void* din = nullptr;
        if (stage_in(p, 2, out, n3 * sizeof(float), &din)) return -2;
        dout = const_cast<void*>(din);
    } else if (stage_out_begin(p, 2, out, n3 * sizeof(float), &dout)) {
        return -2;
    }
    const size_t n4 = n3 / 4;
    const size_t want = (n4 + 255) / 256, cap = (size_t)p->sm_count * 32;
    const unsigned grid = (unsigned)(want < cap ? want : cap);
    if (accumulate)
        k_radiometer_noise<true><<<grid, 256, 0, p->stream>>>(d_sigma, (const float*)dn, seed, N, n4, (float*)dout);
    else
        k_radiometer_noise<false><<<grid, 256, 0, p->stream>>>(d_sigma, (const float*)dn, seed, N, n4, (float*)dout);
    FB_LAUNCH_CHECK();
    if (stage_out_end(p, 2, out, n3 * sizeof(float))) return -2;
    return 0;
}

int fb_mean_spectrum_filter(fb_plan* p, const float* field, float* out, double* mean_out) {
    FB_CUDA(cudaSetDevice(p->device));
    const int N = p->N;
    FB_CHECK(field, "fb_mean_spectrum_filter: NULL field");
    const size_t n2 = (size_t)N * N, n3 = n2 * N;
    const void* din = nullptr;
    void* dout = nullptr;
    if (stage_in(p, 0, field, n3 * sizeof(float), &din)) return -2;
    if (stage_out_begin(p, 2, out, n3 * sizeof(float), &dout)) return -2;
    // aux: sums [N] f64, mean [N] f64, mean [N] f32
    if (ensure_aux(p, (size_t)N * (2 * sizeof(double) + sizeof(float)))) return -2;
    double* d_sums = (double*)p->aux;
    double* d_mean64 = d_sums + N;
    float* d_mean32 = (float*)(d_mean64 + N);
    FB_CUDA(cudaMemsetAsync(d_sums, 0, (size_t)N * sizeof(double), p->stream));
    const unsigned grid = (unsigned)((size_t)p->sm_count * 8 < n2 ? (size_t)p->sm_count * 8 : n2);
    k_channel_sums<<<grid, 256, 0, p->stream>>>((const float*)din, N, n2, d_sums);
    FB_LAUNCH_CHECK();
    k_mean_from_sums<<<(N + 255) / 256, 256, 0, p->stream>>>(d_sums, 1.0 / (double)n2, N, d_mean64, d_mean32);
    FB_LAUNCH_CHECK();
    if (dout) {
        const size_t n4 = n3 / 4, want = (n4 + 255) / 256, cap = (size_t)p->sm_count * 32;
        k_sub_channel<<<(unsigned)(want < cap ? want : cap), 256, 0, p->stream>>>((const float*)din, d_mean64, N, n4,
                                                                                (float*)dout);
        FB_LAUNCH_CHECK();
    }
    if (mean_out) {
        FB_CUDA(cudaMemcpyAsync(mean_out, d_mean64, (size_t)N * sizeof(double), cudaMemcpyDefault, p->stream));
        FB_CUDA(cudaStreamSynchronize(p->stream));
    }
    if (stage_out_end(p, 2, out, n3 * sizeof(float))) return -2;
    return 0;
}

}  // extern "C"
