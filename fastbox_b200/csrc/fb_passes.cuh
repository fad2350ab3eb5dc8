// fb_passes.cuh -- the FFT pass kernels of the 3-D real<->half-complex transform.
//
// Layout: half spectrum S[a][b][c], a = kx in [0,N/2] (axis 0 halved),
// b = ky, c = kz, complex64, C order.  Inverse (realise) order:
//     rows  (c -> z, contiguous; prologue builds the Hermitian spectrum)
//     cols  (b -> y, stride N)
//     xc2r  (a -> x, stride = plane size; half-length complex FFT + real pre-processing)
// Forward (P(k)) order is the mirror image: xr2c, cols, rows (+ histogram epilogue).
#pragma once
#include "fb_kspace.cuh"

namespace fb {

enum { SRC_NOISE = 0, SRC_PHILOX = 1, SRC_SPEC = 2, SRC_CUBE = 3 };

template <int N>
struct RowGeom {
    using C = FftCfg<N>;
    static constexpr int T = C::T;
    static constexpr int RB = 256 / T;          // rows per CTA (rows are flat over (plane, b))
    static constexpr int THREADS = 256;
    static constexpr size_t SMEM = (size_t)RB * RowLayout<N>::ROW * sizeof(float2);
};

struct RowsArgs {
    const float* re;            // SRC_NOISE: full [N][N][N] float32 noise cubes (box.py:174-175)
    const float* im;
    const float2* src;          // SRC_SPEC: half spectrum (local planes); SRC_CUBE: full complex cube
    const float2* cross;        // rows_fwd: second spectrum for cross power (local planes) or NULL
    uint64_t seed;              // SRC_PHILOX
    float2* work;               // rows_inv: out [na][N][N];  rows_fwd: in/out
    float2* spec_out;           // optional copy of the spectrum (local planes)
    const float2* tw;
    long nrows;                 // na * N
    int flags;
    int kind;
    KSpace K;
    PkDev pk;
};

// ---------------------------------------------------------------------------
// rows, inverse.  grid = ceil(na*N / RB), block = 256
// ---------------------------------------------------------------------------
template <int N, int SRC>
__global__ void __launch_bounds__(RowGeom<N>::THREADS) k_rows_inv(const RowsArgs A) {
    using G = RowGeom<N>;
    using C = FftCfg<N>;
    constexpr int P = C::P, T = C::T;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2* sm = reinterpret_cast<float2*>(smem_raw);
    __shared__ PkShared pks;

    const int rl = threadIdx.x / T, t = threadIdx.x % T;
    const long row_raw = (long)blockIdx.x * G::RB + rl;
    const bool rvalid = row_raw < A.nrows;
    const long row_id = rvalid ? row_raw : A.nrows - 1;
    const int al = (int)(row_id / N);
    const int a = A.K.a0 + al;
    const int b = (int)(row_id % N);
    const bool do_pk = (A.flags & FB_F_PK) != 0;
    const bool poles = (A.flags & FB_F_POLES) != 0;
    if (do_pk) {
        pk_shared_init(pks, A.K);
        __syncthreads();
    }
    const size_t row_local = ((size_t)al * N + b) * N;
    const int am = (N - a) & (N - 1), bm = (N - b) & (N - 1);
    const size_t row_g = ((size_t)a * N + b) * N, row_m = ((size_t)am * N + bm) * N;
    const float wmult = (a == 0 || a == N / 2) ? 1.f : 2.f;
    const double sab = do_pk ? __dadd_rn(A.K.ax[a], A.K.ay[b]) : 0.0;
    const bool velocity = (A.kind >= FB_KIND_VEL_X && A.kind <= FB_KIND_VEL_Z);

    float2 v[P];
#pragma unroll
    for (int q = 0; q < P; ++q) {
        const int c = t + T * q;
        const int cm = (N - c) & (N - 1);
        float2 h;
        if constexpr (SRC == SRC_SPEC) {
            h = A.src[row_local + c];
            const float amp = k_amp(A.K, A.flags, A.kind, a, b, c, c);
            h.x *= amp;
            h.y *= amp;
        } else {
            float2 g, gm;
            if constexpr (SRC == SRC_NOISE) {
                g = make_float2(__ldg(&A.re[row_g + c]), __ldg(&A.im[row_g + c]));
                gm = make_float2(__ldg(&A.re[row_m + cm]), __ldg(&A.im[row_m + cm]));
            } else if constexpr (SRC == SRC_PHILOX) {
                g = philox_normal_pair(A.seed, row_g + c);
                gm = philox_normal_pair(A.seed, row_m + cm);
            } else {
                g = __ldg(&A.src[row_g + c]);
                gm = __ldg(&A.src[row_m + cm]);
            }
            const float amp = k_amp(A.K, A.flags, A.kind, a, b, c, c);
            float ampm = amp;
            if constexpr (SRC == SRC_CUBE) {
                if (A.flags & FB_F_FILTER) ampm = k_amp(A.K, A.flags, A.kind, a, b, c, cm);
            }
            g.x *= amp; g.y *= amp;
            gm.x *= ampm; gm.y *= ampm;
            if (A.flags & FB_F_ANTIHERM)        // (G(k) - conj G(-k)) / (2i)
                h = make_float2(0.5f * (g.y + gm.y), -0.5f * (g.x - gm.x));
            else                                // (G(k) + conj G(-k)) / 2
                h = make_float2(0.5f * (g.x + gm.x), 0.5f * (g.y - gm.y));
        }
        if (velocity) h = make_float2(-h.y, h.x);          // * i, box.py:254-256
        if (A.spec_out && rvalid) A.spec_out[row_local + c] = h;
        if (do_pk) {
            const double azc = A.K.az[c];
            const double s = __dadd_rn(sab, azc);
            const int bin = pk_bin(pks, A.K.nedges, s);
            const float p = (h.x * h.x + h.y * h.y) * (float)A.K.inv_boxfactor;
            const float mu2 = (poles && s > 0.0) ? (float)azc / (float)s : 0.f;
            pk_accumulate(pks, bin, wmult, p, mu2, poles, rvalid);
        }
        v[q] = h;
    }
    RowLayout<N> sl{rl * RowLayout<N>::ROW};
    fft_regs<N, P, C::R1, C::R2, C::R3, +1>(v, t, sm, sl, A.tw);
#pragma unroll
    for (int q = 0; q < P; ++q)
        if (rvalid) A.work[row_local + t + T * q] = v[q];
    if (do_pk) pk_shared_flush(pks, A.K, A.pk, poles);
}

// ---------------------------------------------------------------------------
// rows, forward: c <- z on work[a][b][:], then optional store + P(k) binning
// (auto |S|^2 or cross Re S conj(X)).   grid = ceil(na*N / RB)
// ---------------------------------------------------------------------------
template <int N>
__global__ void __launch_bounds__(RowGeom<N>::THREADS) k_rows_fwd(const RowsArgs A) {
    using G = RowGeom<N>;
    using C = FftCfg<N>;
    constexpr int P = C::P, T = C::T;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2* sm = reinterpret_cast<float2*>(smem_raw);
    __shared__ PkShared pks;

    const int rl = threadIdx.x / T, t = threadIdx.x % T;
    const long row_raw = (long)blockIdx.x * G::RB + rl;
    const bool rvalid = row_raw < A.nrows;
    const long row_id = rvalid ? row_raw : A.nrows - 1;
    const int al = (int)(row_id / N);
    const int a = A.K.a0 + al;
    const int b = (int)(row_id % N);
    const bool do_pk = (A.flags & FB_F_PK) != 0;
    const bool poles = (A.flags & FB_F_POLES) != 0;
    if (do_pk) {
        pk_shared_init(pks, A.K);
        __syncthreads();
    }
    const size_t row_local = ((size_t)al * N + b) * N;
    float2 v[P];
#pragma unroll
    for (int q = 0; q < P; ++q) v[q] = A.work[row_local + t + T * q];
    RowLayout<N> sl{rl * RowLayout<N>::ROW};
    fft_regs<N, P, C::R1, C::R2, C::R3, -1>(v, t, sm, sl, A.tw);
    const float wmult = (a == 0 || a == N / 2) ? 1.f : 2.f;
    const double sab = do_pk ? __dadd_rn(A.K.ax[a], A.K.ay[b]) : 0.0;
#pragma unroll
    for (int q = 0; q < P; ++q) {
        const int c = t + T * q;
        const float2 h = v[q];
        if (A.spec_out && rvalid) A.spec_out[row_local + c] = h;
        if (do_pk) {
            const double azc = A.K.az[c];
            const double s = __dadd_rn(sab, azc);
            const int bin = pk_bin(pks, A.K.nedges, s);
            float p;
            if (A.cross) {
                const float2 x = A.cross[row_local + c];
                p = (h.x * x.x + h.y * x.y) * (float)A.K.inv_boxfactor;
            } else {
                p = (h.x * h.x + h.y * h.y) * (float)A.K.inv_boxfactor;
            }
            const float mu2 = (poles && s > 0.0) ? (float)azc / (float)s : 0.f;
            pk_accumulate(pks, bin, wmult, p, mu2, poles, rvalid);
        }
    }
    if (do_pk) pk_shared_flush(pks, A.K, A.pk, poles);
}

// ---------------------------------------------------------------------------
// columns c2c (y axis): data[plane][b][z], FFT over b (stride N), CZ columns per CTA.
// grid = (N/CZ, nplanes), block = CZ*T, dyn smem = N*CZ*8
// ---------------------------------------------------------------------------
template <int N, int CZ>
struct ColGeom {
    using C = FftCfg<N>;
    static constexpr int THREADS = CZ * C::T;
    static constexpr size_t SMEM = (size_t)N * CZ * sizeof(float2);
};

template <int N, int CZ, int S>
__global__ void __launch_bounds__(ColGeom<N, CZ>::THREADS) k_cols_c2c(float2* __restrict__ data,
                                                                     const float2* __restrict__ tw) {
    using C = FftCfg<N>;
    constexpr int P = C::P, T = C::T;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2* sm = reinterpret_cast<float2*>(smem_raw);
    const int col = threadIdx.x % CZ, t = threadIdx.x / CZ;
    float2* base = data + (size_t)blockIdx.y * N * N + (size_t)blockIdx.x * CZ + col;
    float2 v[P];
#pragma unroll
    for (int q = 0; q < P; ++q) v[q] = base[(size_t)(t + T * q) * N];
    ColLayout<CZ> sl{col};
    fft_regs<N, P, C::R1, C::R2, C::R3, S>(v, t, sm, sl, tw);
#pragma unroll
    for (int q = 0; q < P; ++q) base[(size_t)(t + T * q) * N] = v[q];
}

// ---------------------------------------------------------------------------
// x axis, half complex -> real (last pass of the inverse transform).
// spec[a][g], a in [0,N/2], g in [0,ncols) (ncols = plane size = stride);
// out[x][g], x in [0,N).  M = N/2 point complex FFT:
//   Z[k] = (X[k] + conj X[M-k]) + i e^{+2 pi i k/N} (X[k] - conj X[M-k]),  z = IFFT_M(Z)
//   out[2m] = Re z[m], out[2m+1] = Im z[m]
// epilogue: * scale, optional exp (log-normal, box.py:457) and sum / sum of squares.
// grid = ncols/CZ, block = CZ*T(M)
// ---------------------------------------------------------------------------
struct XArgs {
    const float2* spec;
    float* field;
    const float* field_in;
    float2* spec_out;
    const float2* tw;
    size_t ncols;
    int flags;
    float scale;
    double* sums;           // [0] sum, [1] sum of squares
};

template <int N, int CZ>
struct XGeom {
    static constexpr int M = N / 2;
    using C = FftCfg<M>;
    static constexpr int THREADS = CZ * C::T;
    static constexpr size_t SMEM = (size_t)M * CZ * sizeof(float2);
};

template <int N, int CZ>
__global__ void __launch_bounds__(XGeom<N, CZ>::THREADS) k_x_c2r(const XArgs A) {
    constexpr int M = N / 2;
    using C = FftCfg<M>;
    constexpr int P = C::P, T = C::T;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2* sm = reinterpret_cast<float2*>(smem_raw);
    __shared__ double red[2][32];
    const int col = threadIdx.x % CZ, t = threadIdx.x / CZ;
    const size_t g = (size_t)blockIdx.x * CZ + col;
    const float2* src = A.spec + g;
    float2 v[P];
#pragma unroll
    for (int q = 0; q < P; ++q) {
        const int k = t + T * q;
        const float2 xk = src[(size_t)k * A.ncols];
        const float2 xm = cconj(src[(size_t)(M - k) * A.ncols]);
        float2 w = __ldg(&A.tw[k * (FB_NMAX_TW / N)]);
        w.y = -w.y;                                    // e^{+2 pi i k / N}
        const float2 sp = cadd(xk, xm), df = cmul(csub(xk, xm), w);
        v[q] = make_float2(sp.x - df.y, sp.y + df.x);  // sp + i*df
    }
    ColLayout<CZ> sl{col};
    fft_regs<M, P, C::R1, C::R2, C::R3, +1>(v, t, sm, sl, A.tw);
    float* dst = A.field + g;
    const bool do_exp = (A.flags & FB_F_EXP) != 0;
    float acc = 0.f, acc2 = 0.f;
#pragma unroll
    for (int q = 0; q < P; ++q) {
        const int m = t + T * q;
        float r0 = v[q].x * A.scale, r1 = v[q].y * A.scale;
        if (do_exp) {
            r0 = expf(r0);
            r1 = expf(r1);
        }
        acc += r0 + r1;
        acc2 = fmaf(r0, r0, fmaf(r1, r1, acc2));
        dst[(size_t)(2 * m) * A.ncols] = r0;
        dst[(size_t)(2 * m + 1) * A.ncols] = r1;
    }
    if (A.sums) {
        double s1 = warp_sum((double)acc), s2 = warp_sum((double)acc2);
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
        if (lane == 0) {
            red[0][warp] = s1;
            red[1][warp] = s2;
        }
        __syncthreads();
        if (warp == 0) {
            const int nw = (blockDim.x + 31) >> 5;
            s1 = lane < nw ? red[0][lane] : 0.0;
            s2 = lane < nw ? red[1][lane] : 0.0;
            s1 = warp_sum(s1);
            s2 = warp_sum(s2);
            if (lane == 0) {
                atomicAdd(&A.sums[0], s1);
                atomicAdd(&A.sums[1], s2);
            }
        }
    }
}

// ---------------------------------------------------------------------------
// x axis, real -> half complex (first pass of the forward transform).
//   z[m] = in[2m] + i in[2m+1],  Z = FFT_M(z),
//   X[k] = 1/2 (Z[k] + conj Z[M-k]) - i/2 e^{-2 pi i k/N} (Z[k] - conj Z[M-k]),  k = 0..M
// ---------------------------------------------------------------------------
template <int N, int CZ>
__global__ void __launch_bounds__(XGeom<N, CZ>::THREADS) k_x_r2c(const XArgs A) {
    constexpr int M = N / 2;
    using C = FftCfg<M>;
    constexpr int P = C::P, T = C::T;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2* sm = reinterpret_cast<float2*>(smem_raw);
    const int col = threadIdx.x % CZ, t = threadIdx.x / CZ;
    const size_t g = (size_t)blockIdx.x * CZ + col;
    const float* src = A.field_in + g;
    float2 v[P];
#pragma unroll
    for (int q = 0; q < P; ++q) {
        const int m = t + T * q;
        v[q] = make_float2(src[(size_t)(2 * m) * A.ncols], src[(size_t)(2 * m + 1) * A.ncols]);
    }
    ColLayout<CZ> sl{col};
    fft_regs<M, P, C::R1, C::R2, C::R3, -1>(v, t, sm, sl, A.tw);
    __syncthreads();
#pragma unroll
    for (int q = 0; q < P; ++q) sm[sl(t + T * q)] = v[q];
    __syncthreads();
    float2* dst = A.spec_out + g;
#pragma unroll
    for (int q = 0; q < P; ++q) {
        const int k = t + T * q;
        const float2 zk = v[q];
        const float2 zm = cconj(sm[sl((M - k) & (M - 1))]);
        const float2 w = __ldg(&A.tw[k * (FB_NMAX_TW / N)]);      // e^{-2 pi i k / N}
        const float2 sp = cadd(zk, zm), df = cmul(csub(zk, zm), w);
        // 1/2 (sp - i df)
        dst[(size_t)k * A.ncols] = make_float2(0.5f * (sp.x + df.y), 0.5f * (sp.y - df.x));
        if (k == 0) dst[(size_t)M * A.ncols] = make_float2(zk.x - zk.y, 0.f);
    }
}

}  // namespace fb
