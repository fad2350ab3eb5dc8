// fb_beam.cu -- per-channel zero-padded 2-D FFT beam convolution, BeamModel.convolve_fft
// (fastbox/beams.py:81-87):
//     norm[z]  = sum_xy beam[x,y,z]
//     out      = scipy.signal.fftconvolve(beam, field, mode='same', axes=[0,1]) / norm[z]
// i.e. a LINEAR (zero padded to 2N >= 2N-1) 2-D convolution per frequency channel z, cropped to the
// first argument's frame: out[x,y] = full[x+st, y+st], st = (N-1)//2.  z is the contiguous batch axis.
//
// Three passes over HBM (60 B/cell):
//   A  y axis, real (N rows, zero padded to 2N) -> N+1 complex, for field and beam      4->8, 4->8
//   B  x axis: zero padded c2c(2N) of field and beam, product, inverse c2c(2N), crop     8+8->8
//   C  y axis, N+1 complex -> 2N real, crop, * 1/((2N)^2 norm[z])                        8->4
// The beam's channel sum norm[z] is the DC bin of its 2-D transform and falls out of pass B.
#include "fb_launch.h"

namespace fb {

// ---- pass A: y-axis r2c with zero padding.  grid = (N/CZ, N, 2), block = CZ*T(N)
template <int N, int CZ>
__global__ void __launch_bounds__(CZ * FftCfg<N>::T) k_beam_y_r2c(const float* __restrict__ field,
                                                                 const float* __restrict__ beam,
                                                                 float2* __restrict__ yf, float2* __restrict__ yb,
                                                                 const float2* __restrict__ tw) {
    constexpr int M = N, NF = 2 * N;                 // M complex points represent 2N reals
    using C = FftCfg<M>;
    constexpr int P = C::P, T = C::T;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2* sm = reinterpret_cast<float2*>(smem_raw);
    const int col = threadIdx.x % CZ, t = threadIdx.x / CZ;
    const int x = blockIdx.y;
    const size_t zc = (size_t)blockIdx.x * CZ + col;
    const float* src = (blockIdx.z ? beam : field) + (size_t)x * N * N + zc;
    float2* dst = (blockIdx.z ? yb : yf) + (size_t)x * (M + 1) * N + zc;
    float2 v[P];
#pragma unroll
    for (int q = 0; q < P; ++q) {
        const int m = t + T * q;                     // rows 2m, 2m+1; zero for rows >= N
        v[q] = (2 * m + 1 < N) ? make_float2(src[(size_t)(2 * m) * N], src[(size_t)(2 * m + 1) * N])
                               : make_float2(0.f, 0.f);
    }
    ColLayout<CZ> sl{col};
    fft_regs<M, P, C::R1, C::R2, C::R3, -1>(v, t, sm, sl, tw);
    __syncthreads();
#pragma unroll
    for (int q = 0; q < P; ++q) sm[sl(t + T * q)] = v[q];
    __syncthreads();
#pragma unroll
    for (int q = 0; q < P; ++q) {
        const int k = t + T * q;
        const float2 zk = v[q];
        const float2 zm = cconj(sm[sl((M - k) & (M - 1))]);
        const float2 w = FB_TW(tw, NF, k);      // e^{-2 pi i k / 2N}
        const float2 sp = cadd(zk, zm), df = cmul(csub(zk, zm), w);
        dst[(size_t)k * N] = make_float2(0.5f * (sp.x + df.y), 0.5f * (sp.y - df.x));
        if (k == 0) dst[(size_t)M * N] = make_float2(zk.x - zk.y, 0.f);
    }
}

// ---- pass B: x axis.  grid = (N/CZ, N+1), block = CZ*T(2N)
template <int N, int CZ>
__global__ void __launch_bounds__(CZ * FftCfg<2 * N>::T, 1) k_beam_x(float2* __restrict__ yf,
                                                                    const float2* __restrict__ yb,
                                                                    double* __restrict__ norm,
                                                                    const float2* __restrict__ tw) {
    constexpr int NF = 2 * N;
    using C = FftCfg<NF>;
    constexpr int P = C::P, T = C::T;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2* sm = reinterpret_cast<float2*>(smem_raw);
    const int col = threadIdx.x % CZ, t = threadIdx.x / CZ;
    const int ky = blockIdx.y;
    const size_t zc = (size_t)blockIdx.x * CZ + col;
    const size_t plane = (size_t)(N + 1) * N;
    float2* f = yf + (size_t)ky * N + zc;
    const float2* b = yb + (size_t)ky * N + zc;
    float2 vf[P], vb[P];
#pragma unroll
    for (int q = 0; q < P; ++q) {
        const int x = t + T * q;                     // zero padding for x >= N
        vf[q] = x < N ? f[(size_t)x * plane] : make_float2(0.f, 0.f);
        vb[q] = x < N ? b[(size_t)x * plane] : make_float2(0.f, 0.f);
    }
    ColLayout<CZ> sl{col};
    fft_regs<NF, P, C::R1, C::R2, C::R3, -1>(vf, t, sm, sl, tw);
    __syncthreads();
    fft_regs<NF, P, C::R1, C::R2, C::R3, -1>(vb, t, sm, sl, tw);
    __syncthreads();
    if (ky == 0 && t == 0) norm[zc] = (double)vb[0].x;           // DC bin = sum_xy beam (beams.py:81)
#pragma unroll
    for (int q = 0; q < P; ++q) vf[q] = cmul(vf[q], vb[q]);
    fft_regs<NF, P, C::R1, C::R2, C::R3, +1>(vf, t, sm, sl, tw);
    constexpr int st = (N - 1) / 2;                  // 'same' crop w.r.t. the first argument
#pragma unroll
    for (int q = 0; q < P; ++q) {
        const int xo = t + T * q - st;
        if (xo >= 0 && xo < N) f[(size_t)xo * plane] = vf[q];
    }
}

// ---- pass B, role-split variant for large boxes.  grid = (N/CZ, N+1), block = 2*CZ*T(2N).
// The first CZ*T threads transform the field columns, the other CZ*T the beam columns (whole warps per
// role), so a thread holds 16 points instead of 32 and a 1024-thread CTA fits an SM (32 warps instead
// of 16 for the two forward transforms).  The beam spectrum crosses to the field threads through the
// exchange buffer; the beam warps then only keep the barriers of the inverse transform company.
// Measured SLOWER than k_beam_x at 1024^3 (one 1024-thread CTA per SM serialises its load / transform /
// store phases); kept as an opt-in experiment (FB_BEAM_SPLIT=1) and covered by the parity tests.
template <int N, int CZ>
__global__ void __launch_bounds__(2 * CZ * FftCfg<2 * N>::T, 1) k_beam_x_split(float2* __restrict__ yf,
                                                                              const float2* __restrict__ yb,
                                                                              double* __restrict__ norm,
                                                                              const float2* __restrict__ tw) {
    constexpr int NF = 2 * N;
    using C = FftCfg<NF>;
    constexpr int P = C::P, T = C::T;
    static_assert((CZ * T) % 32 == 0, "roles must own whole warps");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2* sm = reinterpret_cast<float2*>(smem_raw);
    const int role = threadIdx.x / (CZ * T);             // 0 field, 1 beam
    const int r = threadIdx.x - role * (CZ * T);
    const int col = r % CZ, t = r / CZ;
    const int ky = blockIdx.y;
    const size_t zc = (size_t)blockIdx.x * CZ + col;
    const size_t plane = (size_t)(N + 1) * N;
    float2* f = yf + (size_t)ky * N + zc;
    const float2* src = (role ? yb : yf) + (size_t)ky * N + zc;
    float2 v[P];
#pragma unroll
    for (int q = 0; q < P; ++q) {
        const int x = t + T * q;                         // zero padding for x >= N
        v[q] = x < N ? src[(size_t)x * plane] : make_float2(0.f, 0.f);
    }
    ColLayout<2 * CZ> sl{role * CZ + col};
    fft_regs<NF, P, C::R1, C::R2, C::R3, -1>(v, t, sm, sl, tw);
    __syncthreads();
    if (role) {
        if (ky == 0 && t == 0) norm[zc] = (double)v[0].x;            // DC bin = sum_xy beam (beams.py:81)
        fft_store_natural<NF, P>(v, t, sm, sl);
    }
    __syncthreads();
    if (!role) {
        ColLayout<2 * CZ> sb{CZ + col};
        float2 vb[P];
        fft_exchange_read<NF, P>(vb, t, sm, sb);
#pragma unroll
        for (int q = 0; q < P; ++q) v[q] = cmul(v[q], vb[q]);
    }
    __syncthreads();
    if (role) {
        fft_regs_barriers_only<C::R2, C::R3>();
        return;
    }
    fft_regs<NF, P, C::R1, C::R2, C::R3, +1>(v, t, sm, sl, tw);
    constexpr int st = (N - 1) / 2;                      // 'same' crop w.r.t. the first argument
#pragma unroll
    for (int q = 0; q < P; ++q) {
        const int xo = t + T * q - st;
        if (xo >= 0 && xo < N) f[(size_t)xo * plane] = v[q];
    }
}

// ---- pass C: y axis c2r with crop and normalisation.  grid = (N/CZ, N), block = CZ*T(N)
template <int N, int CZ>
__global__ void __launch_bounds__(CZ * FftCfg<N>::T) k_beam_y_c2r(const float2* __restrict__ yf,
                                                                 const double* __restrict__ norm,
                                                                 float* __restrict__ out,
                                                                 const float2* __restrict__ tw) {
    constexpr int M = N, NF = 2 * N;
    using C = FftCfg<M>;
    constexpr int P = C::P, T = C::T;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2* sm = reinterpret_cast<float2*>(smem_raw);
    const int col = threadIdx.x % CZ, t = threadIdx.x / CZ;
    const int x = blockIdx.y;
    const size_t zc = (size_t)blockIdx.x * CZ + col;
    const float2* src = yf + (size_t)x * (M + 1) * N + zc;
    ColLayout<CZ> sl{col};
    float2 v[P];
#pragma unroll
    for (int q = 0; q < P; ++q) {
        v[q] = src[(size_t)(t + T * q) * N];
        sm[sl(t + T * q)] = v[q];
    }
    float2 xnyq = make_float2(0.f, 0.f);
    if (t == 0) xnyq = src[(size_t)M * N];
    __syncthreads();
#pragma unroll
    for (int q = 0; q < P; ++q) {
        const int k = t + T * q;
        const float2 xk = v[q];
        const float2 xm = cconj(k == 0 ? xnyq : sm[sl(M - k)]);
        float2 w = FB_TW(tw, NF, k);
        w.y = -w.y;
        const float2 sp = cadd(xk, xm), df = cmul(csub(xk, xm), w);
        v[q] = make_float2(sp.x - df.y, sp.y + df.x);
    }
    if constexpr (C::R2 > 1) __syncthreads();
    fft_regs<M, P, C::R1, C::R2, C::R3, +1>(v, t, sm, sl, tw);
    const float scale = (float)(1.0 / ((double)NF * (double)NF * norm[zc]));      // beams.py:87
    constexpr int st = (N - 1) / 2;
    float* dst = out + (size_t)x * N * N + zc;
#pragma unroll
    for (int q = 0; q < P; ++q) {
        const int m = t + T * q;
        const int y0 = 2 * m - st, y1 = 2 * m + 1 - st;
        if (y0 >= 0 && y0 < N) dst[(size_t)y0 * N] = v[q].x * scale;
        if (y1 >= 0 && y1 < N) dst[(size_t)y1 * N] = v[q].y * scale;
    }
}

template <int N, int CZA, int CZB>
static int beam_run_cz(fb_plan* p, const float* beam, const float* field, float* out, float2* yf, float2* yb,
                       double* norm) {
    {
        auto kern = k_beam_y_r2c<N, CZA>;
        const size_t smem = (size_t)(N + N / 16) * CZA * sizeof(float2);
        if (set_smem(kern, smem)) return -2;
        kern<<<dim3(N / CZA, N, 2), CZA * FftCfg<N>::T, smem, p->stream>>>(field, beam, yf, yb, p->tw);
        FB_LAUNCH_CHECK();
    }
    if constexpr (N >= 512) {
        if (env_int("FB_BEAM_SPLIT", 0)) {          // measured slower at 1024^3 (32.7 vs 28.8 ms): opt-in
            auto kern = k_beam_x_split<N, CZB>;
            const size_t smem = (size_t)(2 * N + 2 * N / 16) * 2 * CZB * sizeof(float2);
            if (set_smem(kern, smem)) return -2;
            kern<<<dim3(N / CZB, N + 1), 2 * CZB * FftCfg<2 * N>::T, smem, p->stream>>>(yf, yb, norm, p->tw);
            FB_LAUNCH_CHECK();
        } else {
            auto kern = k_beam_x<N, CZB>;
            const size_t smem = (size_t)(2 * N + 2 * N / 16) * CZB * sizeof(float2);
            if (set_smem(kern, smem)) return -2;
            kern<<<dim3(N / CZB, N + 1), CZB * FftCfg<2 * N>::T, smem, p->stream>>>(yf, yb, norm, p->tw);
            FB_LAUNCH_CHECK();
        }
    } else {
        auto kern = k_beam_x<N, CZB>;
        const size_t smem = (size_t)(2 * N + 2 * N / 16) * CZB * sizeof(float2);
        if (set_smem(kern, smem)) return -2;
        kern<<<dim3(N / CZB, N + 1), CZB * FftCfg<2 * N>::T, smem, p->stream>>>(yf, yb, norm, p->tw);
        FB_LAUNCH_CHECK();
    }
    {
        auto kern = k_beam_y_c2r<N, CZA>;
        const size_t smem = (size_t)(N + N / 16) * CZA * sizeof(float2);
        if (set_smem(kern, smem)) return -2;
        kern<<<dim3(N / CZA, N), CZA * FftCfg<N>::T, smem, p->stream>>>(yf, norm, out, p->tw);
        FB_LAUNCH_CHECK();
    }
    return 0;
}

// columns (z) per CTA: y passes CZA, x pass CZB.  Large boxes take narrower tiles so that two or
// three CTAs share an SM and their load / transform / store phases overlap.
template <int N>
static int beam_run(fb_plan* p, const float* beam, const float* field, float* out, float2* yf, float2* yb,
                    double* norm) {
    if constexpr (N >= 512) {
        const int cza = env_int("FB_BEAM_CZA", 8), czb = env_int("FB_BEAM_CZB", 4);
        if (cza == 8 && czb == 2) return beam_run_cz<N, 8, 2>(p, beam, field, out, yf, yb, norm);
        if (cza == 8) return beam_run_cz<N, 8, 4>(p, beam, field, out, yf, yb, norm);
        if (czb == 2) return beam_run_cz<N, 16, 2>(p, beam, field, out, yf, yb, norm);
        return beam_run_cz<N, 16, 4>(p, beam, field, out, yf, yb, norm);
    } else {
        constexpr int CZA = N >= 16 ? 16 : N;
        constexpr int CZB = N >= 4 ? 4 : N;
        return beam_run_cz<N, CZA, CZB>(p, beam, field, out, yf, yb, norm);
    }
}

}  // namespace fb

using namespace fb;

extern "C" int fb_beam_convolve(fb_plan* p, const float* beam, const float* field, float* out) {
    FB_CUDA(cudaSetDevice(p->device));
    const int N = p->N;
    FB_CHECK(beam && field && out, "fb_beam_convolve: NULL buffer");
    FB_CHECK(N >= 8 && N <= 1024, "fb_beam_convolve: N=%d not supported (8..1024; padded transform is 2N <= 2048)", N);
    const size_t n3 = (size_t)N * N * N, ny = (size_t)N * (N + 1) * N;
    const void *db = nullptr, *df = nullptr;
    void* dout = nullptr;
    if (stage_in(p, 0, beam, n3 * sizeof(float), &db)) return -2;
    if (stage_in(p, 1, field, n3 * sizeof(float), &df)) return -2;
    if (stage_out_begin(p, 2, out, n3 * sizeof(float), &dout)) return -2;
    if (ensure_aux(p, 2 * ny * sizeof(float2) + (size_t)N * sizeof(double))) return -2;
    float2* yf = (float2*)p->aux;
    float2* yb = yf + ny;
    double* norm = (double*)(yb + ny);
    int rc = 0;
    switch (N) {
        case 8: rc = beam_run<8>(p, (const float*)db, (const float*)df, (float*)dout, yf, yb, norm); break;
        case 16: rc = beam_run<16>(p, (const float*)db, (const float*)df, (float*)dout, yf, yb, norm); break;
        case 32: rc = beam_run<32>(p, (const float*)db, (const float*)df, (float*)dout, yf, yb, norm); break;
        case 64: rc = beam_run<64>(p, (const float*)db, (const float*)df, (float*)dout, yf, yb, norm); break;
        case 128: rc = beam_run<128>(p, (const float*)db, (const float*)df, (float*)dout, yf, yb, norm); break;
        case 256: rc = beam_run<256>(p, (const float*)db, (const float*)df, (float*)dout, yf, yb, norm); break;
        case 512: rc = beam_run<512>(p, (const float*)db, (const float*)df, (float*)dout, yf, yb, norm); break;
        case 1024: rc = beam_run<1024>(p, (const float*)db, (const float*)df, (float*)dout, yf, yb, norm); break;
        default: set_error("fb_beam_convolve: unsupported N=%d", N); return -1;
    }
    if (rc) return rc;
    if (stage_out_end(p, 2, out, n3 * sizeof(float))) return -2;
    return 0;
}
