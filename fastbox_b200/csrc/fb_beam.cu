// fb_beam.cu -- per-channel zero-padded 2-D FFT beam convolution, BeamModel.convolve_fft
// (fastbox/beams.py:81-87):
//     norm[z]  = sum_xy beam[x,y,z]
//     out      = scipy.signal.fftconvolve(beam, field, mode='same', axes=[0,1]) / norm[z]
// i.e. a LINEAR (zero padded to 2N >= 2N-1) 2-D convolution per frequency channel z, cropped to the
// first argument's frame: out[x,y] = full[x+st, y+st], st = (N-1)//2.  z is the contiguous batch axis.
//
// The beam cube of a survey is the same for every field that is convolved with it, so its transform is
// computed ONCE per plan (fb_beam_set) and kept in HBM, already divided by (2N)^2 norm[z]:
//     BS[ky][zt][kx][c]   kx = 0..2N-1, ky = 0..N, channel z = 8 zt + c      (16 B/cell, complex64)
// A convolution is then three passes over HBM (56 B/cell; the x pass does two 2N-point transforms per column
// instead of three and holds one register array instead of two):
//   A  y axis, real field rows (zero padded to 2N) -> N+1 complex                        4 -> 8
//   B  x axis: zero padded c2c(2N), * BS, inverse c2c(2N), crop, in place            8 + 16 -> 8
//   C  y axis, N+1 complex -> 2N real, crop                                              8 -> 4
// The intermediate Y lives in the same tiled layout Y[ky][zt][x][c] (c = 0..7): the tile a CTA of pass B
// transforms is one contiguous 64 N-byte block, and passes A / C touch it in 64-byte rows.
// norm[z] is a float64 sum over the beam cube (np.sum(beam, axis=(0,1)), beams.py:81).
#include "fb_launch.h"

namespace fb {

#define FB_BEAM_CT 8                                  // channels per tile of Y / BS

template <int N>
struct BeamGeom {
    static constexpr int CT = N < FB_BEAM_CT ? N : FB_BEAM_CT;
    static constexpr int NZT = N / CT;
    static constexpr int CZB = CT < 4 ? CT : 4;        // columns per CTA of the x pass
};

__device__ __forceinline__ size_t beam_tile(int N, int nzt, int ct, int k, int zt, int rows) {
    return (((size_t)k * nzt + zt) * rows) * ct;       // first element of tile (k, zt): `rows` rows of ct channels
}

// ---- norm[z] = sum over the N^2 pixels of beam[.,.,z] in float64; inv[z] = 1 / ((2N)^2 norm[z]) --------------
__global__ void __launch_bounds__(256) k_beam_norm(const float* __restrict__ beam, int N, size_t nrows,
                                                    double* __restrict__ norm) {
    const size_t per = (nrows + gridDim.x - 1) / gridDim.x;
    const size_t r0 = (size_t)blockIdx.x * per, r1 = r0 + per < nrows ? r0 + per : nrows;
    for (int z = threadIdx.x; z < N; z += blockDim.x) {
        double a0 = 0.0, a1 = 0.0;
        size_t r = r0;
        for (; r + 1 < r1; r += 2) {
            a0 += (double)__ldg(&beam[r * N + z]);
            a1 += (double)__ldg(&beam[(r + 1) * N + z]);
        }
        if (r < r1) a0 += (double)__ldg(&beam[r * N + z]);
        if (r1 > r0) atomicAdd(&norm[z], a0 + a1);
    }
}
__global__ void k_beam_inv(const double* __restrict__ norm, int N, float* __restrict__ inv) {
    const int z = blockIdx.x * blockDim.x + threadIdx.x;
    if (z < N) inv[z] = (float)(1.0 / (4.0 * (double)N * (double)N * norm[z]));       // beams.py:87
}

// ---- pass A: y-axis r2c with zero padding.  grid = (NZT, N), block = CT*T(N) ------------------------------------
template <int N>
__global__ void __launch_bounds__(BeamGeom<N>::CT* FftCfg<N>::T) k_beam_y_r2c(const float* __restrict__ src,
                                                                              float2* __restrict__ Y,
                                                                              const float2* __restrict__ tw) {
    constexpr int M = N, NF = 2 * N;                 // M complex points represent 2N reals
    using C = FftCfg<M>;
    using G = BeamGeom<N>;
    constexpr int P = C::P, T = C::T, CT = G::CT;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2* sm = reinterpret_cast<float2*>(smem_raw);
    const int col = threadIdx.x % CT, t = threadIdx.x / CT;
    const int x = blockIdx.y, zt = blockIdx.x;
    const float* in = src + (size_t)x * N * N + (size_t)zt * CT + col;
    float2 v[P];
#pragma unroll
    for (int q = 0; q < P; ++q) {
        const int m = t + T * q;                     // rows 2m, 2m+1; zero for rows >= N
        v[q] = (2 * m + 1 < N) ? make_float2(in[(size_t)(2 * m) * N], in[(size_t)(2 * m + 1) * N])
                               : make_float2(0.f, 0.f);
    }
    ColLayout<CT> sl{col};
    fft_regs<M, P, C::R1, C::R2, C::R3, -1>(v, t, sm, sl, tw);
    __syncthreads();
#pragma unroll
    for (int q = 0; q < P; ++q) sm[sl(t + T * q)] = v[q];
    __syncthreads();
    const size_t kstride = (size_t)G::NZT * N * CT;  // from tile (k, zt) to tile (k+1, zt)
    float2* dst = Y + beam_tile(N, G::NZT, CT, 0, zt, N) + (size_t)x * CT + col;
    const float2 w0 = FB_TW(tw, NF, t);               // one table load, compile-time rotations (see k_x_c2r)
#pragma unroll
    for (int q = 0; q < P; ++q) {
        const int k = t + T * q;
        const float2 zk = v[q];
        const float2 zm = cconj(sm[sl((M - k) & (M - 1))]);
        const float2 w = q == 0 ? w0 : rot_pi16(w0, q * (16 / P), -1);      // e^{-2 pi i k / 2N}
        const float2 sp = cadd(zk, zm), df = cmul(csub(zk, zm), w);
        dst[(size_t)k * kstride] = make_float2(0.5f * (sp.x + df.y), 0.5f * (sp.y - df.x));
        if (k == 0) dst[(size_t)M * kstride] = make_float2(zk.x - zk.y, 0.f);
    }
}

// ---- pass B: x axis.  grid = (NZT * CT/CZB, N+1), block = CZB*T(2N) ---------------------------------------------
// SETUP: Y holds the y-transformed BEAM; the 2N-point transform, scaled by inv[z], is stored to BS.
// else : Y holds the y-transformed field; transform, multiply by BS, inverse transform, crop, store in place.
template <int N, bool SETUP>
__global__ void __launch_bounds__(BeamGeom<N>::CZB* FftCfg<2 * N>::T, (BeamGeom<N>::CZB * FftCfg<2 * N>::T <= 512) ? 2 : 1)
    k_beam_x(float2* __restrict__ Y, float2* __restrict__ BS, const float* __restrict__ inv,
             const float2* __restrict__ tw, int pf_dist) {
    constexpr int NF = 2 * N;
    using C = FftCfg<NF>;
    using G = BeamGeom<N>;
    constexpr int P = C::P, T = C::T, CT = G::CT, CZ = G::CZB;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2* sm = reinterpret_cast<float2*>(smem_raw);
    const int col = threadIdx.x % CZ, t = threadIdx.x / CZ;
    const int ky = blockIdx.y;
    const int zt = blockIdx.x / (CT / CZ), c = (blockIdx.x % (CT / CZ)) * CZ + col;
    float2* y = Y + beam_tile(N, G::NZT, CT, ky, zt, N) + c;
    float2* bs = BS + beam_tile(N, G::NZT, CT, ky, zt, NF) + c;
    if constexpr (!SETUP && N >= 256) {
        // L2 prefetch of the (contiguous) Y and BS tiles of the CTA pf_dist blocks ahead, once per tile, in 32 KB pieces
        constexpr int PER = CT / CZ;                                 // CTAs per tile
        constexpr unsigned PIECE = (size_t)N * CT * sizeof(float2) >= 32768 ? 32768u : (unsigned)((size_t)N * CT * sizeof(float2));
        constexpr int NY = (int)((size_t)N * CT * sizeof(float2) / PIECE), NB = 2 * NY;
        if (pf_dist > 0 && (int)threadIdx.x < NY + NB) {
            const unsigned id = blockIdx.y * gridDim.x + blockIdx.x + pf_dist;
            const unsigned pk = id / gridDim.x, pbx = id - pk * gridDim.x;
            if (pk <= (unsigned)N && pbx % PER == 0) {
                const int pzt = pbx / PER;
                const unsigned char* base = threadIdx.x < NY
                    ? reinterpret_cast<const unsigned char*>(Y + beam_tile(N, G::NZT, CT, pk, pzt, N)) + threadIdx.x * PIECE
                    : reinterpret_cast<const unsigned char*>(BS + beam_tile(N, G::NZT, CT, pk, pzt, NF)) + (threadIdx.x - NY) * PIECE;
                l2_prefetch(base, PIECE);
            }
        }
    }
    float2 v[P];
#pragma unroll
    for (int q = 0; q < P; ++q) {
        const int x = t + T * q;                     // zero padding for x >= N
        v[q] = x < N ? y[(size_t)x * CT] : make_float2(0.f, 0.f);
    }
    ColLayout<CZ> sl{col};
    fft_regs<NF, P, C::R1, C::R2, C::R3, -1>(v, t, sm, sl, tw);
    if constexpr (SETUP) {
        const float s = __ldg(&inv[zt * CT + c]);
#pragma unroll
        for (int q = 0; q < P; ++q) bs[(size_t)(t + T * q) * CT] = make_float2(v[q].x * s, v[q].y * s);
    } else {
#pragma unroll
        for (int q = 0; q < P; ++q) v[q] = cmul(v[q], __ldg(&bs[(size_t)(t + T * q) * CT]));
        __syncthreads();                             // the exchange buffer is reused
        fft_regs<NF, P, C::R1, C::R2, C::R3, +1>(v, t, sm, sl, tw);
        constexpr int st = (N - 1) / 2;              // 'same' crop w.r.t. the first argument
#pragma unroll
        for (int q = 0; q < P; ++q) {
            const int xo = t + T * q - st;
            if (xo >= 0 && xo < N) y[(size_t)xo * CT] = v[q];
        }
    }
}

// ---- pass C: y axis c2r with crop (the normalisation sits in BS).  grid = (NZT, N), block = CT*T(N) -------------
template <int N>
__global__ void __launch_bounds__(BeamGeom<N>::CT* FftCfg<N>::T) k_beam_y_c2r(const float2* __restrict__ Y,
                                                                              float* __restrict__ out,
                                                                              const float2* __restrict__ tw) {
    constexpr int M = N, NF = 2 * N;
    using C = FftCfg<M>;
    using G = BeamGeom<N>;
    constexpr int P = C::P, T = C::T, CT = G::CT;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2* sm = reinterpret_cast<float2*>(smem_raw);
    const int col = threadIdx.x % CT, t = threadIdx.x / CT;
    const int x = blockIdx.y, zt = blockIdx.x;
    const size_t kstride = (size_t)G::NZT * N * CT;
    const float2* src = Y + beam_tile(N, G::NZT, CT, 0, zt, N) + (size_t)x * CT + col;
    ColLayout<CT> sl{col};
    float2 v[P];
#pragma unroll
    for (int q = 0; q < P; ++q) {
        v[q] = src[(size_t)(t + T * q) * kstride];
        sm[sl(t + T * q)] = v[q];
    }
    float2 xnyq = make_float2(0.f, 0.f);
    if (t == 0) xnyq = src[(size_t)M * kstride];
    __syncthreads();
    const float2 w0 = FB_TW(tw, NF, t);
#pragma unroll
    for (int q = 0; q < P; ++q) {
        const int k = t + T * q;
        const float2 xk = v[q];
        const float2 xm = cconj(k == 0 ? xnyq : sm[sl(M - k)]);
        float2 w = q == 0 ? w0 : rot_pi16(w0, q * (16 / P), -1);
        w.y = -w.y;
        const float2 sp = cadd(xk, xm), df = cmul(csub(xk, xm), w);
        v[q] = make_float2(sp.x - df.y, sp.y + df.x);
    }
    if constexpr (C::R2 > 1) __syncthreads();
    fft_regs<M, P, C::R1, C::R2, C::R3, +1>(v, t, sm, sl, tw);
    constexpr int st = (N - 1) / 2;
    float* dst = out + (size_t)x * N * N + (size_t)zt * CT + col;
#pragma unroll
    for (int q = 0; q < P; ++q) {
        const int m = t + T * q;
        const int y0 = 2 * m - st, y1 = 2 * m + 1 - st;
        if (y0 >= 0 && y0 < N) dst[(size_t)y0 * N] = v[q].x;
        if (y1 >= 0 && y1 < N) dst[(size_t)y1 * N] = v[q].y;
    }
}

template <int N>
static int beam_pass_a(fb_plan* p, const float* src, float2* Y) {
    using G = BeamGeom<N>;
    auto kern = k_beam_y_r2c<N>;
    const size_t smem = (size_t)(N + N / 16) * G::CT * sizeof(float2);
    if (set_smem(kern, smem)) return -2;
    kern<<<dim3(G::NZT, N), G::CT * FftCfg<N>::T, smem, p->stream>>>(src, Y, p->tw);
    FB_LAUNCH_CHECK();
    return 0;
}

template <int N, bool SETUP>
static int beam_pass_b(fb_plan* p, float2* Y, float2* BS, const float* inv) {
    using G = BeamGeom<N>;
    auto kern = k_beam_x<N, SETUP>;
    const size_t smem = (size_t)(2 * N + 2 * N / 16) * G::CZB * sizeof(float2);
    if (set_smem(kern, smem)) return -2;
    kern<<<dim3(G::NZT * (G::CT / G::CZB), N + 1), G::CZB * FftCfg<2 * N>::T, smem, p->stream>>>(
        Y, BS, inv, p->tw, env_int("FB_BEAM_PF", FB_BEAM_PF_DEFAULT));
    FB_LAUNCH_CHECK();
    return 0;
}

template <int N>
static int beam_set(fb_plan* p, const float* beam, float2* Y, float2* BS, double* norm, float* inv) {
    FB_CUDA(cudaMemsetAsync(norm, 0, (size_t)N * sizeof(double), p->stream));
    const size_t nrows = (size_t)N * N;
    const unsigned grid = (unsigned)((size_t)p->sm_count * 8 < nrows ? (size_t)p->sm_count * 8 : nrows);
    k_beam_norm<<<grid, 256, 0, p->stream>>>(beam, N, nrows, norm);
    FB_LAUNCH_CHECK();
    k_beam_inv<<<(N + 255) / 256, 256, 0, p->stream>>>(norm, N, inv);
    FB_LAUNCH_CHECK();
    if (int rc = beam_pass_a<N>(p, beam, Y)) return rc;
    return beam_pass_b<N, true>(p, Y, BS, inv);
}

template <int N>
static int beam_run(fb_plan* p, const float* field, float* out, float2* Y, float2* BS) {
    if (int rc = beam_pass_a<N>(p, field, Y)) return rc;
    if (int rc = beam_pass_b<N, false>(p, Y, BS, nullptr)) return rc;
    using G = BeamGeom<N>;
    auto kern = k_beam_y_c2r<N>;
    const size_t smem = (size_t)(N + N / 16) * G::CT * sizeof(float2);
    if (set_smem(kern, smem)) return -2;
    kern<<<dim3(G::NZT, N), G::CT * FftCfg<N>::T, smem, p->stream>>>(Y, out, p->tw);
    FB_LAUNCH_CHECK();
    return 0;
}

static size_t beam_y_elems(int N) { return (size_t)N * (N + 1) * N; }

}  // namespace fb

using namespace fb;

#define FB_BEAM_DISPATCH(N_, CALL)                     \
    switch (N_) {                                      \
        case 8: { constexpr int NN = 8; rc = CALL; } break;       \
        case 16: { constexpr int NN = 16; rc = CALL; } break;     \
        case 32: { constexpr int NN = 32; rc = CALL; } break;     \
        case 64: { constexpr int NN = 64; rc = CALL; } break;     \
        case 128: { constexpr int NN = 128; rc = CALL; } break;   \
        case 256: { constexpr int NN = 256; rc = CALL; } break;   \
        case 512: { constexpr int NN = 512; rc = CALL; } break;   \
        case 1024: { constexpr int NN = 1024; rc = CALL; } break; \
        default: set_error("beam convolution: unsupported N=%d", N_); return -1; \
    }

extern "C" int fb_beam_set(fb_plan* p, const float* beam) {
    FB_CUDA(cudaSetDevice(p->device));
    const int N = p->N;
    FB_CHECK(beam != nullptr, "fb_beam_set: NULL beam");
    FB_CHECK(N >= 8 && N <= 1024, "fb_beam_set: N=%d not supported (8..1024; padded transform is 2N <= 2048)", N);
    const size_t n3 = (size_t)N * N * N, ny = beam_y_elems(N);
    const void* db = nullptr;
    if (stage_in(p, 0, beam, n3 * sizeof(float), &db)) return -2;
    const size_t bs_bytes = 2 * ny * sizeof(float2) + (size_t)N * (sizeof(double) + sizeof(float));
    p->beam_ready = 0;
    if (p->beam_spec_bytes < bs_bytes) {
        if (p->beam_spec) FB_CUDA(cudaFree(p->beam_spec));
        p->beam_spec = nullptr;
        p->beam_spec_bytes = 0;
        FB_CUDA(cudaMalloc(&p->beam_spec, bs_bytes));
        p->beam_spec_bytes = bs_bytes;
    }
    if (ensure_aux(p, ny * sizeof(float2))) return -2;
    float2* BS = (float2*)p->beam_spec;
    double* norm = (double*)(BS + 2 * ny);
    float* inv = (float*)(norm + N);
    int rc = 0;
    FB_BEAM_DISPATCH(N, (beam_set<NN>(p, (const float*)db, (float2*)p->aux, BS, norm, inv)));
    if (rc) return rc;
    p->beam_ready = 1;
    return 0;
}

extern "C" int fb_beam_convolve(fb_plan* p, const float* beam, const float* field, float* out) {
    FB_CUDA(cudaSetDevice(p->device));
    const int N = p->N;
    FB_CHECK(field && out, "fb_beam_convolve: NULL buffer");
    FB_CHECK(N >= 8 && N <= 1024, "fb_beam_convolve: N=%d not supported (8..1024; padded transform is 2N <= 2048)", N);
    if (beam != nullptr) {
        if (int rc = fb_beam_set(p, beam)) return rc;
    }
    FB_CHECK(p->beam_ready, "fb_beam_convolve: no beam (pass a beam cube or call fb_beam_set first)");
    const size_t n3 = (size_t)N * N * N, ny = beam_y_elems(N);
    const void* df = nullptr;
    void* dout = nullptr;
    if (stage_in(p, 1, field, n3 * sizeof(float), &df)) return -2;
    if (stage_out_begin(p, 2, out, n3 * sizeof(float), &dout)) return -2;
    if (ensure_aux(p, ny * sizeof(float2))) return -2;
    int rc = 0;
    FB_BEAM_DISPATCH(N, (beam_run<NN>(p, (const float*)df, (float*)dout, (float2*)p->aux, (float2*)p->beam_spec)));
    if (rc) return rc;
    if (stage_out_end(p, 2, out, n3 * sizeof(float))) return -2;
    return 0;
}
