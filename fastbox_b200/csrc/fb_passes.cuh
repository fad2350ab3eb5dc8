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
// Row kernels.  The k-space prologue / epilogue works in NATURAL order: thread t of a
// row owns the P consecutive modes c = P*t .. P*t+P-1 (vectorised 16-byte global
// accesses; the P(k) bin changes at most once or twice along such a run, so moments
// are accumulated in registers and flushed once per thread).  The FFT itself works in
// Stockham order (thread t owns t + T*q); one shared-memory transpose connects the two.
// ---------------------------------------------------------------------------
template <int P>
__device__ __forceinline__ void load_run(const float* __restrict__ p, float (&out)[P]) {
    const float4* p4 = reinterpret_cast<const float4*>(p);
#pragma unroll
    for (int i = 0; i < P / 4; ++i) {
        const float4 v = __ldg(p4 + i);
        out[4 * i] = v.x; out[4 * i + 1] = v.y; out[4 * i + 2] = v.z; out[4 * i + 3] = v.w;
    }
}
template <int P>
__device__ __forceinline__ void load_run(const float2* __restrict__ p, float2 (&out)[P]) {
    const float4* p4 = reinterpret_cast<const float4*>(p);
#pragma unroll
    for (int i = 0; i < P / 2; ++i) {
        const float4 v = __ldg(p4 + i);
        out[2 * i] = make_float2(v.x, v.y);
        out[2 * i + 1] = make_float2(v.z, v.w);
    }
}
template <int P>
__device__ __forceinline__ void store_run(float2* __restrict__ p, const float2 (&in)[P]) {
    float4* p4 = reinterpret_cast<float4*>(p);
#pragma unroll
    for (int i = 0; i < P / 2; ++i) p4[i] = make_float4(in[2 * i].x, in[2 * i].y, in[2 * i + 1].x, in[2 * i + 1].y);
}

// running P(k) moments of one thread while it walks along c
struct PkWalk {
    int bin;
    unsigned cnt;
    double s1, s2, l2, l4;      // float64: a bin holding only a Hermitian pair gets stddev == 0 exactly
};

__device__ __forceinline__ void pk_walk_flush_thread(PkShared& sh, PkWalk& w, bool poles) {
    if (w.cnt) {
        atomicAdd(&sh.cnt[w.bin], (unsigned long long)w.cnt);
        atomicAdd(&sh.s1[w.bin], w.s1);
        atomicAdd(&sh.s2[w.bin], w.s2);
        if (poles) {
            atomicAdd(&sh.l2[w.bin], w.l2);
            atomicAdd(&sh.l4[w.bin], w.l4);
        }
    }
    w.cnt = 0u;
    w.s1 = w.s2 = w.l2 = w.l4 = 0.0;
}

// add mode (s, p) with multiplicity wf to the walk; s is monotone along a run so the bin moves by
// single steps (both directions are handled: |m_c| decreases for c >= N/2)
__device__ __forceinline__ void pk_walk_add(PkShared& sh, PkWalk& w, int nedges, double s, float p, float wf,
                                            float mu2, bool poles) {
    int nb = w.bin;
    while (nb < nedges && s >= sh.thr[nb]) ++nb;
    while (nb > 0 && s < sh.thr[nb - 1]) --nb;
    if (nb != w.bin) {
        pk_walk_flush_thread(sh, w, poles);
        w.bin = nb;
    }
    const double pd = (double)p, wp = (double)wf * pd;
    w.cnt += (unsigned)(wf + 0.5f);
    w.s1 += wp;
    w.s2 += wp * pd;
    if (poles) {
        const double m2 = (double)mu2;
        w.l2 += wp * (1.5 * m2 - 0.5);
        w.l4 += wp * ((35.0 * m2 * m2 - 30.0 * m2 + 3.0) * 0.125);
    }
}

// end of the run: lanes of a warp holding the same bin are combined before touching smem
__device__ __forceinline__ void pk_walk_flush_warp(PkShared& sh, const PkWalk& w, bool poles, bool valid) {
    const unsigned full = 0xffffffffu;
    const bool have = valid && w.cnt > 0u;
    unsigned todo = __ballot_sync(full, have);
    const int lane = threadIdx.x & 31;
    while (todo) {
        const int leader = __ffs(todo) - 1;
        const int lb = __shfl_sync(full, w.bin, leader);
        const bool mine = have && (w.bin == lb);
        const unsigned grp = __ballot_sync(full, mine);
        const unsigned cnt = __reduce_add_sync(full, mine ? w.cnt : 0u);
        const double a1 = warp_sum(mine ? w.s1 : 0.0);
        const double a2 = warp_sum(mine ? w.s2 : 0.0);
        double b2 = 0.0, b4 = 0.0;
        if (poles) {
            b2 = warp_sum(mine ? w.l2 : 0.0);
            b4 = warp_sum(mine ? w.l4 : 0.0);
        }
        if (lane == leader) {
            atomicAdd(&sh.cnt[lb], (unsigned long long)cnt);
            atomicAdd(&sh.s1[lb], a1);
            atomicAdd(&sh.s2[lb], a2);
            if (poles) {
                atomicAdd(&sh.l2[lb], b2);
                atomicAdd(&sh.l4[lb], b4);
            }
        }
        todo &= ~grp;
    }
}

// Per-run multiplier amp[e] for modes c0+e (sqrt(P) LUT, filter, velocity / potential factor).
// Each option is ONE warp-uniform branch around an unrolled loop (no per-element flag tests).
template <int N, int P>
__device__ __forceinline__ void run_amp(const KSpace& K, int flags, int kind, int a, int b, int c0, float base,
                                        float (&amp)[P]) {
    const int ma = mode_number(a, N), mb = mode_number(b, N);
#pragma unroll
    for (int e = 0; e < P; ++e) amp[e] = base;
    if (flags & FB_F_SQRTPK) {
        if (K.sqrtp_mode == 1) {
            const float* lut = K.sqrtp + (ma * ma + mb * mb);
#pragma unroll
            for (int e = 0; e < P; ++e) {
                const int mc = mode_number(c0 + e, N);
                amp[e] *= __ldg(lut + mc * mc);
            }
        } else {
            const float sab = (float)(ma * ma) * K.inv_lx2 + (float)(mb * mb) * K.inv_ly2;
#pragma unroll
            for (int e = 0; e < P; ++e) {
                const int mc = mode_number(c0 + e, N);
                const float s = sab + (float)(mc * mc) * K.inv_lz2;
                float val = 0.f;                                   // nan_to_num(P(0)) = 0, box.py:167
                if (s > 0.f) {
                    float x = (log2f(s) - K.log2s0) * K.inv_dlog2s;
                    x = fminf(fmaxf(x, 0.f), (float)(K.sqrtp_n - 1) - 1e-3f);
                    const int i0 = (int)x;
                    const float y0 = __ldg(&K.sqrtp[i0]), y1 = __ldg(&K.sqrtp[i0 + 1]);
                    val = fmaf(x - (float)i0, y1 - y0, y0);
                }
                amp[e] *= val;
            }
        }
    }
    if (flags & FB_F_FILTER) {
        if (K.tdense) {
            float tf[P];
            load_run<P>(K.tdense + ((size_t)a * N + b) * N + c0, tf);
#pragma unroll
            for (int e = 0; e < P; ++e) amp[e] *= tf[e];
        } else {
            float tf[P];
            load_run<P>(K.tpar + c0, tf);
            const float tp = __ldg(&K.tperp[a * N + b]);
#pragma unroll
            for (int e = 0; e < P; ++e) amp[e] *= tp * tf[e];
        }
    }
    if (kind != FB_KIND_PLAIN) {
        const float sab = (float)(ma * ma) * K.inv_lx2 + (float)(mb * mb) * K.inv_ly2;
#pragma unroll
        for (int e = 0; e < P; ++e) {
            const int mc = mode_number(c0 + e, N);
            const float k2 = 39.478417604357434f * (sab + (float)(mc * mc) * K.inv_lz2);    // (2 pi)^2 s
            const float ik2 = k2 > 0.f ? 1.f / k2 : 0.f;          // nan_to_num at k = 0, box.py:257-259
            float comp = 1.f;
            int m = 1;
            if (kind == FB_KIND_VEL_X) { comp = (float)ma * K.two_pi_over_lx; m = ma; }
            else if (kind == FB_KIND_VEL_Y) { comp = (float)mb * K.two_pi_over_ly; m = mb; }
            else if (kind == FB_KIND_VEL_Z) { comp = (float)mc * K.two_pi_over_lz; m = mc; }
            if (kind != FB_KIND_POTENTIAL && m == -N / 2) comp = 0.f;                       // box.py:268-274
            amp[e] *= comp * ik2;
        }
    }
}

// P(k) moments of one natural-order run h[0..P) (modes c0..c0+P-1 of row (a,b)), optionally
// crossed with x[].  Fast path: first and last mode of the run fall in the same bin (s is monotone
// along a run when T > 1), so no per-mode bin search is needed.
template <int N, int P, bool MONOTONE>
__device__ __forceinline__ void run_pk(PkShared& pks, const KSpace& K, int a, int b, int c0, const float2 (&h)[P],
                                       const float2* x, bool poles, bool rvalid) {
    const float wmult = (a == 0 || a == N / 2) ? 1.f : 2.f;
    const double sab = __dadd_rn(K.ax[a], K.ay[b]);
    const float invb = (float)K.inv_boxfactor;
    PkWalk walk;
    walk.bin = 0;
    walk.cnt = 0u;
    walk.s1 = walk.s2 = walk.l2 = walk.l4 = 0.0;
    const double s_first = __dadd_rn(sab, K.az[c0]), s_last = __dadd_rn(sab, K.az[c0 + P - 1]);
    const int b0 = pk_bin(pks, K.nedges, s_first), b1 = pk_bin(pks, K.nedges, s_last);
    if (MONOTONE && b0 == b1 && !poles) {
        double s1 = 0.0, s2 = 0.0;
#pragma unroll
        for (int e = 0; e < P; ++e) {
            const float p = x ? (h[e].x * x[e].x + h[e].y * x[e].y) * invb : (h[e].x * h[e].x + h[e].y * h[e].y) * invb;
            const double pd = (double)p;
            s1 += pd;
            s2 = fma(pd, pd, s2);
        }
        if (rvalid) {
            walk.bin = b0;
            walk.cnt = (unsigned)P * (unsigned)(wmult + 0.5f);
            walk.s1 = (double)wmult * s1;
            walk.s2 = (double)wmult * s2;
        }
    } else {
        walk.bin = b0;
#pragma unroll
        for (int e = 0; e < P; ++e) {
            const double azc = K.az[c0 + e];
            const double s = __dadd_rn(sab, azc);
            const float p = x ? (h[e].x * x[e].x + h[e].y * x[e].y) * invb : (h[e].x * h[e].x + h[e].y * h[e].y) * invb;
            const float mu2 = (poles && s > 0.0) ? (float)azc / (float)s : 0.f;
            if (rvalid) pk_walk_add(pks, walk, K.nedges, s, p, wmult, mu2, poles);
        }
    }
    pk_walk_flush_warp(pks, walk, poles, rvalid);
}

// ---------------------------------------------------------------------------
// rows, inverse.  grid = ceil(na*N / RB), block = 256
// ---------------------------------------------------------------------------
template <int N, int SRC>
__global__ void __launch_bounds__(RowGeom<N>::THREADS, 3) k_rows_inv(const RowsArgs A) {
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
    const int c0 = P * t;                          // natural-order run of this thread
    const int mstart = N - c0 - P;                 // mirror block: cells mstart .. mstart+P-1  (c -> N-c)
    const int cm0 = (N - c0) & (N - 1);            // mirror of the first cell of the run
    const bool antiherm = (A.flags & FB_F_ANTIHERM) != 0;

    // ---- 1. gather: h = G(k) +/- conj G(-k)   (the factor 1/2 is folded into amp)
    float2 h[P];
    if constexpr (SRC == SRC_SPEC) {
        load_run<P>(A.src + row_local + c0, h);
    } else {
        float2 gm0;
        if constexpr (SRC == SRC_NOISE) {
            float r[P], i[P], rm[P], im[P];
            load_run<P>(A.re + row_g + c0, r);
            load_run<P>(A.im + row_g + c0, i);
            load_run<P>(A.re + row_m + mstart, rm);
            load_run<P>(A.im + row_m + mstart, im);
            gm0 = make_float2(__ldg(&A.re[row_m + cm0]), __ldg(&A.im[row_m + cm0]));
#pragma unroll
            for (int e = 0; e < P; ++e) {
                const float2 g = make_float2(r[e], i[e]);
                const float2 gm = (e == 0) ? gm0 : make_float2(rm[P - e], im[P - e]);
                h[e] = antiherm ? make_float2(g.y + gm.y, gm.x - g.x) : make_float2(g.x + gm.x, g.y - gm.y);
            }
        } else if constexpr (SRC == SRC_PHILOX) {
            float2 mm[P];
#pragma unroll
            for (int e = 0; e < P; e += 2) philox_normal_quad(A.seed, row_g + c0 + e, h[e], h[e + 1]);
#pragma unroll
            for (int e = 0; e < P; e += 2) philox_normal_quad(A.seed, row_m + mstart + e, mm[e], mm[e + 1]);
            float2 q0, q1;
            philox_normal_quad(A.seed, (row_m + cm0) & ~(size_t)1, q0, q1);
            gm0 = ((row_m + cm0) & 1) ? q1 : q0;
#pragma unroll
            for (int e = 0; e < P; ++e) {
                const float2 g = h[e];
                const float2 gm = (e == 0) ? gm0 : mm[P - e];
                h[e] = antiherm ? make_float2(g.y + gm.y, gm.x - g.x) : make_float2(g.x + gm.x, g.y - gm.y);
            }
        } else {
            // full complex cube; a k_par-odd filter multiplies G(k) and G(-k) differently (box.py:378)
            float2 mm[P];
            load_run<P>(A.src + row_g + c0, h);
            load_run<P>(A.src + row_m + mstart, mm);
            gm0 = __ldg(&A.src[row_m + cm0]);
            const bool filt = (A.flags & FB_F_FILTER) != 0;
#pragma unroll
            for (int e = 0; e < P; ++e) {
                float2 g = h[e];
                float2 gm = (e == 0) ? gm0 : mm[P - e];
                if (filt) {
                    const int c = c0 + e, cm = (N - c) & (N - 1);
                    const float f = k_amp(A.K, FB_F_FILTER, FB_KIND_PLAIN, a, b, c, c);
                    const float fm = k_amp(A.K, FB_F_FILTER, FB_KIND_PLAIN, a, b, c, cm);
                    g.x *= f; g.y *= f;
                    gm.x *= fm; gm.y *= fm;
                }
                h[e] = antiherm ? make_float2(g.y + gm.y, gm.x - g.x) : make_float2(g.x + gm.x, g.y - gm.y);
            }
        }
    }
    // ---- 2. k-space multiplier
    {
        float amp[P];
        const int amp_flags = (SRC == SRC_CUBE) ? (A.flags & ~FB_F_FILTER) : A.flags;
        run_amp<N, P>(A.K, amp_flags, A.kind, a, b, c0, SRC == SRC_SPEC ? 1.f : 0.5f, amp);
        if (A.kind >= FB_KIND_VEL_X && A.kind <= FB_KIND_VEL_Z) {          // * i, box.py:254-256
#pragma unroll
            for (int e = 0; e < P; ++e) h[e] = make_float2(-h[e].y * amp[e], h[e].x * amp[e]);
        } else {
#pragma unroll
            for (int e = 0; e < P; ++e) h[e] = make_float2(h[e].x * amp[e], h[e].y * amp[e]);
        }
    }
    if (A.spec_out && rvalid) store_run<P>(A.spec_out + row_local + c0, h);
    // ---- 3. binned moments of |H|^2 (box.py:741-764)
    if (do_pk) run_pk<N, P, (T > 1)>(pks, A.K, a, b, c0, h, nullptr, poles, rvalid);

    // ---- 4. natural -> Stockham order, transform c -> z, store
    RowLayout<N> sl{rl * RowLayout<N>::ROW};
    float2 v[P];
    if constexpr (T > 1) {
#pragma unroll
        for (int e = 0; e < P; ++e) sm[sl(c0 + e)] = h[e];
        __syncthreads();
        fft_exchange_read<N, P>(v, t, sm, sl);
        __syncthreads();
    } else {
#pragma unroll
        for (int e = 0; e < P; ++e) v[e] = h[e];
    }
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
__global__ void __launch_bounds__(RowGeom<N>::THREADS, 3) k_rows_fwd(const RowsArgs A) {
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
    // back to natural order: thread owns c0 .. c0+P-1
    const int c0 = P * t;
    float2 h[P];
    if constexpr (T > 1) {
        __syncthreads();
#pragma unroll
        for (int q = 0; q < P; ++q) sm[sl(t + T * q)] = v[q];
        __syncthreads();
#pragma unroll
        for (int e = 0; e < P; ++e) h[e] = sm[sl(c0 + e)];
    } else {
#pragma unroll
        for (int e = 0; e < P; ++e) h[e] = v[e];
    }
    if (A.spec_out && rvalid) store_run<P>(A.spec_out + row_local + c0, h);
    if (do_pk) {
        if (A.cross) {
            float2 x[P];
            load_run<P>(A.cross + row_local + c0, x);
            run_pk<N, P, (T > 1)>(pks, A.K, a, b, c0, h, x, poles, rvalid);
        } else {
            run_pk<N, P, (T > 1)>(pks, A.K, a, b, c0, h, nullptr, poles, rvalid);
        }
        pk_shared_flush(pks, A.K, A.pk, poles);
    }
}

// ---------------------------------------------------------------------------
// columns c2c (y axis): data[plane][b][z], FFT over b (stride N), CZ columns per CTA.
// grid = (N/CZ, nplanes), block = CZ*T, dyn smem = N*CZ*8
// ---------------------------------------------------------------------------
template <int N, int CZ>
struct ColGeom {
    using C = FftCfg<N>;
    static constexpr int THREADS = CZ * C::T;
    static constexpr size_t SMEM = (size_t)(N + N / 16) * CZ * sizeof(float2);
};

// Slab-decomposed runs exchange the y axis between ranks right after (inverse) / before
// (forward) this pass.  With ny > 0 the y index is split as y = d*ny + y' and the element lives
// at [d][plane][y'][z]: the block for destination rank d is contiguous, so the all-to-all needs
// no pack / unpack pass over HBM.  ny = 0 is the plain [plane][y][z] layout.
__device__ __forceinline__ size_t col_index(int plane, int nplanes, int y, int ny, int N) {
    if (ny == 0) return ((size_t)plane * N + y) * N;
    const int d = y / ny, yy = y - d * ny;
    return (((size_t)d * nplanes + plane) * ny + yy) * N;
}

template <int N, int CZ, int S>
__global__ void __launch_bounds__(ColGeom<N, CZ>::THREADS) k_cols_c2c(const float2* __restrict__ in,
                                                                     float2* __restrict__ out, int in_ny, int out_ny,
                                                                     const float2* __restrict__ tw) {
    using C = FftCfg<N>;
    constexpr int P = C::P, T = C::T;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2* sm = reinterpret_cast<float2*>(smem_raw);
    const int col = threadIdx.x % CZ, t = threadIdx.x / CZ;
    const int plane = blockIdx.y, nplanes = gridDim.y;
    const size_t zc = (size_t)blockIdx.x * CZ + col;
    float2 v[P];
#pragma unroll
    for (int q = 0; q < P; ++q) v[q] = in[col_index(plane, nplanes, t + T * q, in_ny, N) + zc];
    ColLayout<CZ> sl{col};
    fft_regs<N, P, C::R1, C::R2, C::R3, S>(v, t, sm, sl, tw);
#pragma unroll
    for (int q = 0; q < P; ++q) out[col_index(plane, nplanes, t + T * q, out_ny, N) + zc] = v[q];
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
    static constexpr size_t SMEM = (size_t)(M + M / 16) * CZ * sizeof(float2);
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
