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
#define FB_F_FULLCUBE_INTERNAL 1024   // P(k) of a full (non-Hermitian) cube: every plane has weight 1

#ifndef FB_ROWS_MINB
#define FB_ROWS_MINB 3      // CTAs per SM the row kernels are compiled for (register cap 65536/(256*MINB))
#endif
#ifndef FB_ROWS_INV_MINB
#define FB_ROWS_INV_MINB FB_ROWS_MINB
#endif

template <int N>
struct RowGeom {
    using C = FftCfg<N>;
    static constexpr int T = C::T;
    static constexpr int RB = 256 / T;          // rows per CTA (rows are flat over (plane, b))
    static constexpr int THREADS = 256;
    static constexpr size_t FFT_SMEM = (size_t)RB * RowLayout<N>::ROW * sizeof(float2);
    static constexpr size_t PK_SMEM = (size_t)(N + N / 16 + FB_MAX_EDGES) * sizeof(double);           // P(k) tables
    static constexpr int AMP_ROW = N / 2 + 4;   // floats: sqrt(P) amplitude of modes c = 0..N/2 of one row
    static constexpr size_t SMEM = FFT_SMEM + PK_SMEM;
    static constexpr size_t SMEM_INV = SMEM + (T > 1 ? (size_t)RB * AMP_ROW * sizeof(float) : 0);     // rows_inv only
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
    int pf_dist;                // rows_inv, noise source: CTA b asks L2 for the rows of CTA b + pf_dist (0 = off)
    KSpace K;
    PkDev pk;
};

// L2 prefetch of a contiguous block (16-byte aligned, size a multiple of 16): one instruction, no register or
// shared-memory cost.  The first pass of the inverse transform loads its whole input in the prologue and then
// computes for a long time, so the loads of one CTA are poorly overlapped by the two other CTAs of the SM; asking
// L2 for the rows of a CTA that will start ~one wave later turns its HBM latency into an L2 hit.
__device__ __forceinline__ void l2_prefetch(const void* p, unsigned bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}
// one 128-byte line (strided tiles: every thread asks for one row chunk of the tile a later CTA will load)
__device__ __forceinline__ void l2_prefetch_line(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// ---------------------------------------------------------------------------
// Row kernels.  Three thread -> mode mappings are used inside one row (see k_rows_inv): quad order
// for coalesced 16-byte global accesses, run order (P consecutive modes per thread; the P(k) bin
// changes at most once along such a run, so moments are accumulated in registers and flushed once
// per thread) and Stockham order (thread t owns t + T*q) for the FFT itself and for evaluating the
// sqrt(P) amplitudes; shared memory connects them.
// ---------------------------------------------------------------------------
// bulk read-once data (noise cubes).  ld.global.cs was tried here and measured no better (4.30 vs 4.27 ms).
template <int P>
__device__ __forceinline__ void load_run_stream(const float* __restrict__ p, float (&out)[P]) {
    const float4* p4 = reinterpret_cast<const float4*>(p);
#pragma unroll
    for (int i = 0; i < P / 4; ++i) {
        const float4 v = __ldg(p4 + i);
        out[4 * i] = v.x; out[4 * i + 1] = v.y; out[4 * i + 2] = v.z; out[4 * i + 3] = v.w;
    }
}
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

// ---- P(k) moments, accumulated straight into the global (L2-resident) histogram with
// fire-and-forget float64 / uint64 reductions (RED.ADD): no shared-memory histogram, no CTA
// level init / flush barriers.  Thresholds are read through L1 (a few hundred bytes).
__device__ __forceinline__ int pk_bin_g(const KSpace& K, const double* __restrict__ thr, double s) {   // #{ j : thr[j] <= s }
    int g;
    if (K.bin_inv_d > 0.f) {
        // log-spaced edges (box.py:749): one log2 gives the bin, two exact comparisons confirm it
        const float sf = (float)s;
        g = sf > 0.f ? (int)floorf((__log2f(sf) - K.bin_l0) * K.bin_inv_d) + 1 : 0;
        g = max(0, min(g, K.nedges));
    } else {
        int lo = 0, hi = K.nedges;
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (thr[mid] <= s) lo = mid + 1; else hi = mid;
        }
        g = lo;
    }
    while (g < K.nedges && s >= thr[g]) ++g;
    while (g > 0 && s < thr[g - 1]) --g;
    return g;
}
// same, starting from a nearby bin (the other end of a short run)
__device__ __forceinline__ int pk_bin_near(const KSpace& K, const double* __restrict__ thr, double s, int g) {
    while (g < K.nedges && s >= thr[g]) ++g;
    while (g > 0 && s < thr[g - 1]) --g;
    return g;
}

// per-CTA shared-memory copies of the float64 tables the binning needs.  (Per-lane strided reads
// of these through L1 cost 32 sector look-ups per warp request and dominated the kernel.)
struct PkTables {
    const double* az;       // (m_c/Lz)^2 at padded index c + (c >> 4)
    const double* thr;      // bin thresholds on s
};
template <int N>
struct PkSmem {
    static constexpr int AZ = N + N / 16;
    static constexpr size_t BYTES = (size_t)(AZ + FB_MAX_EDGES) * sizeof(double);
};
template <int N>
__device__ __forceinline__ PkTables pk_stage_tables(const KSpace& K, double* sm) {
    const double2* az2 = reinterpret_cast<const double2*>(K.az);
    for (int c = 2 * threadIdx.x; c < N; c += 2 * blockDim.x) {       // pairs never straddle a pad
        const double2 v = __ldg(az2 + (c >> 1));
        double* d = sm + c + (c >> 4);
        d[0] = v.x;
        d[1] = v.y;
    }
    double* thr = sm + PkSmem<N>::AZ;
    for (int j = threadIdx.x; j < K.nedges; j += blockDim.x) thr[j] = __ldg(&K.thr[j]);
    PkTables t;
    t.az = sm;
    t.thr = thr;
    return t;
}

// Split form of pk_stage_tables for kernels that want the table loads in flight while they issue their
// own bulk loads: load into registers first, store to shared memory later (before the CTA barrier).
template <int N>
struct PkStageRegs {
    static constexpr int NI = (N / 2 + 255) / 256;       // double2 loads per thread (256 threads)
    double2 az[NI];
    double thr;
};
template <int N>
__device__ __forceinline__ void pk_stage_load(const KSpace& K, PkStageRegs<N>& r) {
    const double2* az2 = reinterpret_cast<const double2*>(K.az);
#pragma unroll
    for (int i = 0; i < PkStageRegs<N>::NI; ++i) {
        const int c = 2 * ((int)threadIdx.x + i * 256);
        r.az[i] = c < N ? __ldg(az2 + (c >> 1)) : make_double2(0.0, 0.0);
    }
    r.thr = (int)threadIdx.x < K.nedges ? __ldg(&K.thr[threadIdx.x]) : 0.0;
}
template <int N>
__device__ __forceinline__ PkTables pk_stage_store(const KSpace& K, const PkStageRegs<N>& r, double* sm) {
#pragma unroll
    for (int i = 0; i < PkStageRegs<N>::NI; ++i) {
        const int c = 2 * ((int)threadIdx.x + i * 256);
        if (c < N) {
            double* d = sm + c + (c >> 4);
            d[0] = r.az[i].x;
            d[1] = r.az[i].y;
        }
    }
    double* thr = sm + PkSmem<N>::AZ;
    if ((int)threadIdx.x < K.nedges) thr[threadIdx.x] = r.thr;
    PkTables t;
    t.az = sm;
    t.thr = thr;
    return t;
}

struct PkAcc {
    int bin;
    unsigned cnt;
    double s1, s2, l2, l4;      // float64: a bin holding only a Hermitian pair gets stddev == 0 exactly
};

__device__ __forceinline__ void pk_red(const PkDev& out, const PkAcc& a, bool poles) {
    if (a.cnt) {
        // replica chosen by CTA: keeps the per-address reduction rate at L2 far below its limit
        const int i = a.bin + (int)(blockIdx.x & (FB_PK_COPIES - 1)) * (FB_MAX_EDGES + 1);
        atomicAdd(&out.count[i], (unsigned long long)a.cnt);
        atomicAdd(&out.sum1[i], a.s1);
        atomicAdd(&out.sum2[i], a.s2);
        if (poles) {
            atomicAdd(&out.l2[i], a.l2);
            atomicAdd(&out.l4[i], a.l4);
        }
    }
}

// Lanes of a warp hold consecutive runs of one row, so equal bins sit in contiguous lane
// segments: a segmented inclusive scan (5 shuffle steps, independent of the number of bins)
// leaves each segment's total in its last lane, which issues the reductions.
template <bool POLES>
__device__ __forceinline__ void pk_seg_flush2(const PkDev& out, PkAcc a, PkAcc b, bool valid) {
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const int ka = (valid && a.cnt) ? a.bin : (-1 - lane);
    const int kb = (valid && b.cnt) ? b.bin : (-1 - lane);
    // segment = maximal run of consecutive lanes with the same bin (equal bins in non-adjacent
    // lanes, e.g. from different rows sharing a warp, are separate segments)
    const unsigned below = 0xffffffffu >> (31 - lane);
    const int pa = __shfl_up_sync(full, ka, 1), pb = __shfl_up_sync(full, kb, 1);   // all lanes take part
    const int sa = 31 - __clz(__ballot_sync(full, lane == 0 || pa != ka) & below);
    const int sb = 31 - __clz(__ballot_sync(full, lane == 0 || pb != kb) & below);
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const unsigned ca = __shfl_up_sync(full, a.cnt, d), cb = __shfl_up_sync(full, b.cnt, d);
        const double a1 = __shfl_up_sync(full, a.s1, d), a2 = __shfl_up_sync(full, a.s2, d);
        const double b1 = __shfl_up_sync(full, b.s1, d), b2 = __shfl_up_sync(full, b.s2, d);
        double a3 = 0.0, a4 = 0.0, b3 = 0.0, b4 = 0.0;
        if constexpr (POLES) {
            a3 = __shfl_up_sync(full, a.l2, d); a4 = __shfl_up_sync(full, a.l4, d);
            b3 = __shfl_up_sync(full, b.l2, d); b4 = __shfl_up_sync(full, b.l4, d);
        }
        if (lane - d >= sa) {
            a.cnt += ca;
            a.s1 += a1;
            a.s2 += a2;
            if constexpr (POLES) { a.l2 += a3; a.l4 += a4; }
        }
        if (lane - d >= sb) {
            b.cnt += cb;
            b.s1 += b1;
            b.s2 += b2;
            if constexpr (POLES) { b.l2 += b3; b.l4 += b4; }
        }
    }
    const int na = __shfl_down_sync(full, ka, 1), nb = __shfl_down_sync(full, kb, 1);
    if (ka >= 0 && (lane == 31 || na != ka)) pk_red(out, a, POLES);
    if (kb >= 0 && (lane == 31 || nb != kb)) pk_red(out, b, POLES);
}

// Per-run multiplier amp[e] for modes c0+e (sqrt(P) LUT, filter, velocity / potential factor).
// Each option is ONE warp-uniform branch around an unrolled loop (no per-element flag tests).
template <int N, int P>
__device__ __forceinline__ void run_amp(const KSpace& K, int flags, int kind, int a, int b, int c0, float base,
                                        float (&amp)[P], int half = -1) {
    // half: 0 = run lies in c < N/2 (m_c = c), 1 = in c >= N/2 (m_c = c - N), -1 = unknown
    const int ma = mode_number(a, N), mb = mode_number(b, N);
    const int moff = half == 0 ? 0 : (half == 1 ? -N : 0);
#pragma unroll
    for (int e = 0; e < P; ++e) amp[e] = base;
    if (flags & FB_F_SQRTPK) {
        if (K.sqrtp_mode == 1) {
            const float* lut = K.sqrtp + (ma * ma + mb * mb);
#pragma unroll
            for (int e = 0; e < P; ++e) {
                const int mc = half < 0 ? mode_number(c0 + e, N) : c0 + e + moff;
                amp[e] *= __ldg(lut + mc * mc);
            }
        } else {
            const float sab = (float)(ma * ma) * K.inv_lx2 + (float)(mb * mb) * K.inv_ly2;
            float sv[P];
#pragma unroll
            for (int e = 0; e < P; ++e) {
                const int mc = half < 0 ? mode_number(c0 + e, N) : c0 + e + moff;
                sv[e] = sab + (float)(mc * mc) * K.inv_lz2;
            }
            if (K.sqrtp_mode == 3) {                     // branch-free inner loops; k = 0 is fixed up below
#pragma unroll
                for (int e = 0; e < P; ++e) amp[e] *= sqrtp_bittable_nz(K, sv[e]);
            } else {
#pragma unroll
                for (int e = 0; e < P; ++e) amp[e] *= sqrtp_logtable(K, sv[e]);
            }
            if (ma == 0 && mb == 0) {
#pragma unroll
                for (int e = 0; e < P; ++e)
                    if (!(sv[e] > 0.f)) amp[e] = 0.f;    // nan_to_num(P(0)) = 0, box.py:167
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
// crossed with x[].  s = |k|^2/(2 pi)^2 is monotone along a run (T > 1), so a run touches the
// bins between those of its first and last mode.  Common case (warp-uniform test): every run of
// the warp spans at most two adjacent bins -> branch-free two-accumulator path.  Otherwise
// (very low |k|, or multipoles requested) each lane walks its run mode by mode.
template <int N, int P, bool MONOTONE>
__device__ __forceinline__ void run_pk(const KSpace& K, const PkTables& tb, const PkDev& out, int a, int b, int c0,
                                       const float2 (&h)[P], const float2* x, bool poles, bool rvalid,
                                       bool full_cube = false) {
    const unsigned full = 0xffffffffu;
    const float wmult = (full_cube || a == 0 || a == N / 2) ? 1.f : 2.f;
    const unsigned wi = (unsigned)(wmult + 0.5f);
    const double sab = __dadd_rn(K.ax[a], K.ay[b]);
    const float invb = (float)K.inv_boxfactor;
    const double* az = tb.az + c0 + (c0 >> 4);           // (m_c/Lz)^2, shared memory, padded every 16
    const double* thr = tb.thr;
    const double s_first = __dadd_rn(sab, az[0]), s_last = __dadd_rn(sab, az[P - 1]);
    const int b0 = pk_bin_g(K, thr, s_first), b1 = MONOTONE ? pk_bin_near(K, thr, s_last, b0) : pk_bin_g(K, thr, s_last);
    const int lo = min(b0, b1), hi = max(b0, b1);
    const bool simple = MONOTONE && (hi - lo <= 1);
    if (__all_sync(full, simple || !rvalid)) {
        // s is monotone along the run, so the run splits at one index: the first kx modes (in the
        // direction of increasing s) fall in bin lo, the rest in bin hi.  kx comes from a 4-step
        // binary search on the shared-memory table; accumulation is then predicated on constants.
        const bool rising = s_last >= s_first;
        int kx = P;                                      // number of modes with s < cut
        if (hi > lo) {
            const double cut = thr[lo];                  // upper edge of bin lo
            int lo_i = 0, hi_i = P;                      // first position (in rising order) with s >= cut
            while (lo_i < hi_i) {
                const int mid = (lo_i + hi_i) >> 1;
                const double sm_ = __dadd_rn(sab, az[rising ? mid : P - 1 - mid]);
                if (sm_ < cut) lo_i = mid + 1; else hi_i = mid;
            }
            kx = lo_i;
        }
        // position of element e in rising order: e (rising) or P-1-e (falling); "low" iff position < kx
        const int e_lo = rising ? 0 : P - kx, e_hi = rising ? kx : P;       // elements [e_lo, e_hi) are in bin lo
        PkAcc A, B;
        A.bin = lo; B.bin = hi;
        A.s1 = A.s2 = B.s1 = B.s2 = 0.0;
        A.l2 = A.l4 = B.l2 = B.l4 = 0.0;
#pragma unroll
        for (int e = 0; e < P; ++e) {
            const float p = x ? (h[e].x * x[e].x + h[e].y * x[e].y) * invb : (h[e].x * h[e].x + h[e].y * h[e].y) * invb;
            const double pd = (double)p;
            if (e >= e_lo && e < e_hi) {
                A.s1 += pd;
                A.s2 = fma(pd, pd, A.s2);
            } else {
                B.s1 += pd;
                B.s2 = fma(pd, pd, B.s2);
            }
        }
        if (poles) {
            // Legendre weights from mu^2 = k_par^2 / k^2 in float32 (relative error 1e-7; the walk path
            // below keeps the float64 form for the few low-k runs)
            const int ma = mode_number(a, N), mb = mode_number(b, N);
            const float sab_f = (float)(ma * ma) * K.inv_lx2 + (float)(mb * mb) * K.inv_ly2;
#pragma unroll
            for (int e = 0; e < P; ++e) {
                const float p = x ? (h[e].x * x[e].x + h[e].y * x[e].y) * invb : (h[e].x * h[e].x + h[e].y * h[e].y) * invb;
                const int mc = mode_number(c0 + e, N);
                const float kz2 = (float)(mc * mc) * K.inv_lz2, sf = sab_f + kz2;
                const float m2 = sf > 0.f ? __fdividef(kz2, sf) : 0.f;
                const double w2 = (double)(p * (1.5f * m2 - 0.5f));
                const double w4 = (double)(p * ((35.f * m2 * m2 - 30.f * m2 + 3.f) * 0.125f));
                if (e >= e_lo && e < e_hi) {
                    A.l2 += w2;
                    A.l4 += w4;
                } else {
                    B.l2 += w2;
                    B.l4 += w4;
                }
            }
        }
        A.cnt = (unsigned)kx * wi;
        B.cnt = (unsigned)(P - kx) * wi;
        A.s1 *= (double)wmult; A.s2 *= (double)wmult;
        B.s1 *= (double)wmult; B.s2 *= (double)wmult;
        if (poles) {
            A.l2 *= (double)wmult; A.l4 *= (double)wmult;
            B.l2 *= (double)wmult; B.l4 *= (double)wmult;
            pk_seg_flush2<true>(out, A, B, rvalid);
        } else {
            pk_seg_flush2<false>(out, A, B, rvalid);
        }
    } else if (rvalid) {
        PkAcc w;
        w.bin = b0;
        w.cnt = 0u;
        w.s1 = w.s2 = w.l2 = w.l4 = 0.0;
#pragma unroll
        for (int e = 0; e < P; ++e) {
            const float p = x ? (h[e].x * x[e].x + h[e].y * x[e].y) * invb : (h[e].x * h[e].x + h[e].y * h[e].y) * invb;
            const double se = __dadd_rn(sab, az[e]);
            int nb = w.bin;
            while (nb < K.nedges && se >= thr[nb]) ++nb;
            while (nb > 0 && se < thr[nb - 1]) --nb;
            if (nb != w.bin) {
                pk_red(out, w, poles);
                w.bin = nb;
                w.cnt = 0u;
                w.s1 = w.s2 = w.l2 = w.l4 = 0.0;
            }
            const double pd = (double)p, wp = (double)wmult * pd;
            w.cnt += wi;
            w.s1 += wp;
            w.s2 = fma(wp, pd, w.s2);
            if (poles) {
                const double m2 = se > 0.0 ? az[e] / se : 0.0;      // (k_par/k)^2
                w.l2 += wp * (1.5 * m2 - 0.5);
                w.l4 += wp * ((35.0 * m2 * m2 - 30.0 * m2 + 3.0) * 0.125);
            }
        }
        pk_red(out, w, poles);
    }
}

// ---------------------------------------------------------------------------
// rows, inverse.  grid = ceil(na*N / RB), block = 256
//
// Three thread->mode mappings are used inside one row (T threads, P modes each):
//   quad     : c = 4t + 4T*j + e      global loads/stores: a warp touches 128 consecutive modes
//                                     with 16-byte lanes (fully coalesced LDG/STG.128)
//   run      : c = P*t + e            P(k) moments: one bin per run almost always
//   stockham : c = t + T*q            the FFT itself
// The Hermitian spectrum is built in quad order, parked in shared memory in natural order,
// and read back in run order (moments) and Stockham order (transform).
// ---------------------------------------------------------------------------
template <int N, int SRC>
__global__ void __launch_bounds__(RowGeom<N>::THREADS, FB_ROWS_INV_MINB) k_rows_inv(const RowsArgs A) {
    using G = RowGeom<N>;
    using C = FftCfg<N>;
    constexpr int P = C::P, T = C::T;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2* sm = reinterpret_cast<float2*>(smem_raw);

    const int rl = threadIdx.x / T, t = threadIdx.x % T;
    const long row_raw = (long)blockIdx.x * G::RB + rl;
    if constexpr (SRC == SRC_NOISE && G::RB * N * sizeof(float) % 16 == 0 && G::RB <= N) {
        // rows of the CTA pf_dist blocks ahead: RB consecutive rows (a, b0 .. b0+RB-1) of re / im and their mirror
        // rows (N-a, N-b), a contiguous block as well (b0 = 0: row 0 and rows N-RB+1 .. N-1)
        const long prow = ((long)blockIdx.x + A.pf_dist) * G::RB;
        if (A.pf_dist > 0 && threadIdx.x < 4 && prow < A.nrows) {
            const int pa = A.K.a0 + (int)(prow / N), pb = (int)(prow % N);
            const int pam = (N - pa) & (N - 1);
            const float* base = (threadIdx.x & 1) ? A.im : A.re;
            constexpr unsigned ROWB = N * sizeof(float);
            if (threadIdx.x < 2) {
                l2_prefetch(base + ((size_t)pa * N + pb) * N, G::RB * ROWB);
            } else if (pb > 0) {
                l2_prefetch(base + ((size_t)pam * N + (N - pb - G::RB + 1)) * N, G::RB * ROWB);
            } else {
                l2_prefetch(base + (size_t)pam * N * N, ROWB);
                if (G::RB > 1) l2_prefetch(base + ((size_t)pam * N + (N - G::RB + 1)) * N, (G::RB - 1) * ROWB);
            }
        }
    }
    if constexpr (SRC == SRC_SPEC && (G::RB * N * sizeof(float2)) % 16 == 0) {
        const long prow = ((long)blockIdx.x + A.pf_dist) * G::RB;       // stored spectrum: RB contiguous rows
        if (A.pf_dist > 0 && threadIdx.x == 0 && prow + G::RB <= A.nrows)
            l2_prefetch(A.src + (size_t)prow * N, G::RB * N * sizeof(float2));
    }
    const bool rvalid = row_raw < A.nrows;
    const long row_id = rvalid ? row_raw : A.nrows - 1;
    const int al = (int)(row_id / N);
    const int a = A.K.a0 + al;
    const int b = (int)(row_id % N);
    const bool do_pk = (A.flags & FB_F_PK) != 0;
    const bool poles = (A.flags & FB_F_POLES) != 0;
    const size_t row_local = ((size_t)al * N + b) * N;
    const int am = (N - a) & (N - 1), bm = (N - b) & (N - 1);
    const size_t row_g = ((size_t)a * N + b) * N, row_m = ((size_t)am * N + bm) * N;
    const bool antiherm = (A.flags & FB_F_ANTIHERM) != 0;
    const bool velocity = (A.kind >= FB_KIND_VEL_X && A.kind <= FB_KIND_VEL_Z);
    const int amp_flags = (SRC == SRC_CUBE) ? (A.flags & ~FB_F_FILTER) : A.flags;
    RowLayout<N> sl{rl * RowLayout<N>::ROW};
    float2 v[P];
    PkTables tb;
    PkStageRegs<N> stage;                                 // table loads in flight during the prologue
    if (do_pk) pk_stage_load<N>(A.K, stage);
    // fast path of the prologue (warp-uniform): noise / Philox source, float-bit sqrt(P) table or
    // none, separable filter or none, plain density field
    const bool fast = T > 1 && SRC != SRC_SPEC && SRC != SRC_CUBE && A.kind == FB_KIND_PLAIN && !antiherm &&
                      (!(A.flags & FB_F_SQRTPK) || A.K.sqrtp_mode == 3) && (!(A.flags & FB_F_FILTER) || !A.K.tdense);
    const int ma_ = mode_number(a, N), mb_ = mode_number(b, N);
    const float sab_f = (float)(ma_ * ma_) * A.K.inv_lx2 + (float)(mb_ * mb_) * A.K.inv_ly2;
    const bool dc_row = ma_ == 0 && mb_ == 0;
    // Philox draws sqrt2 * H0 and the combine doubles it: 1/2 * 1/sqrt2
    float fast_base = (SRC == SRC_PHILOX) ? 0.35355339059327376f : 0.5f;
    if (fast && (A.flags & FB_F_FILTER)) fast_base *= __ldg(&A.K.tperp[a * N + b]);

    bool fast_done = false;
    if constexpr (T > 1 && (SRC == SRC_NOISE || SRC == SRC_PHILOX)) {
        if (fast) {
            // Common configuration (bit-table sqrt(P) or none, separable filter or none, plain
            // field).  The loop body is free of branches (options are predicated loads / selects),
            // so the loads of later quads are scheduled above the arithmetic of earlier ones.
            fast_done = true;
            const bool has_f = (A.flags & FB_F_FILTER) != 0, has_s = (A.flags & FB_F_SQRTPK) != 0;
            const bool do_store = A.spec_out != nullptr && rvalid;
            const float4* tpar4 = reinterpret_cast<const float4*>(A.K.tpar);
            // sqrt(P) depends on |m_c| only: evaluate it once for c = 0..N/2 with consecutive lanes on
            // consecutive modes (their table entries share a few 32-byte sectors; in quad order every
            // lane hit its own sector and the L1 data path became the limiter), park it in shared
            // memory, and read it back in quad order, mirrored for c > N/2.
            float* amp_row = reinterpret_cast<float*>(smem_raw + G::SMEM) + rl * G::AMP_ROW;
            if (has_s) {
#pragma unroll
                for (int i = 0; i <= P / 2; ++i) {
                    const int c = t + T * i;
                    if (i < P / 2 || t == 0) {
                        float am = sqrtp_bittable_nz(A.K, sab_f + (float)(c * c) * A.K.inv_lz2);
                        if (i == 0 && dc_row && t == 0) am = 0.f;                 // nan_to_num(P(0)) = 0, box.py:167
                        amp_row[c] = am;
                    }
                }
                __syncthreads();
            }
            const unsigned full = 0xffffffffu;
            const bool lane_first = (threadIdx.x & 31) == 0 || t == 0;
#pragma unroll
            for (int j = 0; j < P / 4; ++j) {
                const int cq = 4 * t + 4 * T * j;
                const int mq = N - cq - 4, cm0 = (N - cq) & (N - 1);
                float2 g[4], mm[4], gm0;
                if constexpr (SRC == SRC_NOISE) {
                    float r[4], i[4], rm[4], im[4];
                    load_run_stream<4>(A.re + row_g + cq, r);
                    load_run_stream<4>(A.im + row_g + cq, i);
                    load_run_stream<4>(A.re + row_m + mq, rm);
                    load_run_stream<4>(A.im + row_m + mq, im);
                    // the mirror of mode cq is cell N - cq = first cell of the previous lane's mirror block
                    gm0.x = __shfl_up_sync(full, rm[0], 1);
                    gm0.y = __shfl_up_sync(full, im[0], 1);
                    if (lane_first) gm0 = make_float2(__ldg(&A.re[row_m + cm0]), __ldg(&A.im[row_m + cm0]));
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        g[e] = make_float2(r[e], i[e]);
                        mm[e] = make_float2(rm[e], im[e]);
                    }
                } else {
                    // sqrt2 * H0(k) drawn directly (fb_common.cuh); the 1/sqrt2 sits in fast_base.  Rows with
                    // row_g < row_m (every plane except a = 0, N/2) hold canonical cells only: two Philox
                    // blocks per quad.  g = H, y = conj H makes the shared combine below return 2 H.
                    if (row_g < row_m) {
                        philox_h0_sqrt2_quad(A.seed, row_g + cq, g);
                    } else {
#pragma unroll
                        for (int e = 0; e < 4; ++e)
                            g[e] = philox_h0_sqrt2(A.seed, row_g + cq + e, row_m + ((N - cq - e) & (N - 1)));
                    }
                    gm0 = g[0];
                    gm0.y = -gm0.y;
#pragma unroll
                    for (int e = 1; e < 4; ++e) mm[4 - e] = make_float2(g[e].x, -g[e].y);
                }
                const float4 tq = has_f ? __ldg(tpar4 + (cq >> 2)) : make_float4(1.f, 1.f, 1.f, 1.f);
                float am[4] = {tq.x, tq.y, tq.z, tq.w};
                if (has_s) {
                    if (j < P / 8) {                     // c < N/2: amplitudes cq..cq+3
                        const float4 aq = *reinterpret_cast<const float4*>(amp_row + cq);
                        am[0] *= aq.x; am[1] *= aq.y; am[2] *= aq.z; am[3] *= aq.w;
                    } else {                             // c >= N/2: |m_c| = N - c, i.e. amplitudes N-cq, .., N-cq-3
                        const float4 aq = *reinterpret_cast<const float4*>(amp_row + mq);
                        am[0] *= amp_row[mq + 4]; am[1] *= aq.w; am[2] *= aq.z; am[3] *= aq.y;
                    }
                }
                float2 h[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const float f = fast_base * am[e];
                    const float2 y = (e == 0) ? gm0 : mm[4 - e];
                    h[e] = make_float2((g[e].x + y.x) * f, (g[e].y - y.y) * f);
                }
                if (do_store) store_run<4>(A.spec_out + row_local + cq, h);
                float2* pq = sm + sl(cq);
#pragma unroll
                for (int e = 0; e < 4; ++e) pq[e] = h[e];
            }
        }
    }
    if (!fast_done) {
#pragma unroll
    for (int j = 0; j < P / 4; ++j) {
        const int cq = 4 * t + 4 * T * j;              // first mode of this quad
        float2 h[4];
        // ---- 1. gather: h = G(k) +/- conj G(-k)   (the factor 1/2 is folded into amp)
        if constexpr (SRC == SRC_SPEC) {
            load_run<4>(A.src + row_local + cq, h);
        } else {
            const int mq = N - cq - 4;                 // aligned block with the mirrors of e = 1..3
            const int cm0 = (N - cq) & (N - 1);        // mirror of e = 0
            float2 g[4], mm[4], gm0;
            if constexpr (SRC == SRC_NOISE) {
                float r[4], i[4], rm[4], im[4];
                load_run_stream<4>(A.re + row_g + cq, r);
                load_run_stream<4>(A.im + row_g + cq, i);
                load_run_stream<4>(A.re + row_m + mq, rm);
                load_run_stream<4>(A.im + row_m + mq, im);
                gm0 = make_float2(__ldg(&A.re[row_m + cm0]), __ldg(&A.im[row_m + cm0]));
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    g[e] = make_float2(r[e], i[e]);
                    mm[e] = make_float2(rm[e], im[e]);
                }
            } else if constexpr (SRC == SRC_PHILOX) {
#pragma unroll
                for (int e = 0; e < 4; ++e) {                // H0(k) drawn directly; W = H0 is Hermitian
                    const float2 hs = philox_h0_sqrt2(A.seed, row_g + cq + e, row_m + ((N - cq - e) & (N - 1)));
                    g[e] = make_float2(0.70710678118654752f * hs.x, 0.70710678118654752f * hs.y);
                    if (e > 0) mm[4 - e] = cconj(g[e]);
                }
                gm0 = cconj(g[0]);
                (void)mq; (void)cm0;
            } else {
                load_run<4>(A.src + row_g + cq, g);
                load_run<4>(A.src + row_m + mq, mm);
                gm0 = __ldg(&A.src[row_m + cm0]);
            }
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                float2 x = g[e];
                float2 y = (e == 0) ? gm0 : mm[4 - e];                     // cell N - cq - e
                if constexpr (SRC == SRC_CUBE) {
                    // a k_par-odd filter multiplies G(k) and G(-k) differently (box.py:378)
                    if (A.flags & FB_F_FILTER) {
                        const int c = cq + e, cm = (N - c) & (N - 1);
                        const float f = k_amp(A.K, FB_F_FILTER, FB_KIND_PLAIN, a, b, c, c);
                        const float fm = k_amp(A.K, FB_F_FILTER, FB_KIND_PLAIN, a, b, c, cm);
                        x.x *= f; x.y *= f;
                        y.x *= fm; y.y *= fm;
                    }
                }
                h[e] = antiherm ? make_float2(x.y + y.y, y.x - x.x) : make_float2(x.x + y.x, x.y - y.y);
            }
        }
        // ---- 2. k-space multiplier
        float amp[4];
        run_amp<N, 4>(A.K, amp_flags, A.kind, a, b, cq, SRC == SRC_SPEC ? 1.f : 0.5f, amp,
                      (T > 1) ? (j >= P / 8 ? 1 : 0) : -1);
#pragma unroll
        for (int e = 0; e < 4; ++e)
            h[e] = velocity ? make_float2(-h[e].y * amp[e], h[e].x * amp[e])      // * i, box.py:254-256
                            : make_float2(h[e].x * amp[e], h[e].y * amp[e]);
        if (A.spec_out && rvalid) store_run<4>(A.spec_out + row_local + cq, h);
        if constexpr (T > 1) {
            float2* pq = sm + sl(cq);                    // a quad never straddles a pad
#pragma unroll
            for (int e = 0; e < 4; ++e) pq[e] = h[e];
        } else {
#pragma unroll
            for (int e = 0; e < 4; ++e) v[4 * j + e] = h[e];
        }
    }
    }
    if (do_pk) tb = pk_stage_store<N>(A.K, stage, reinterpret_cast<double*>(smem_raw + G::FFT_SMEM));
    if (T > 1 || do_pk) __syncthreads();
    // ---- 3. binned moments of |H|^2 (box.py:741-764), run order
    if (do_pk) {
        const int c0 = P * t;
        float2 hr[P];
        const float2* pr = sm + sl(c0);                  // P consecutive modes, c0 % 16 == 0 (P == 16 when T > 1)
#pragma unroll
        for (int e = 0; e < P; ++e) hr[e] = (T > 1) ? pr[e] : v[e];
        run_pk<N, P, (T > 1)>(A.K, tb, A.pk, a, b, c0, hr, nullptr, poles, rvalid);
    }
    // ---- 4. Stockham order, transform c -> z, store
    if constexpr (T > 1) {
        fft_exchange_read<N, P>(v, t, sm, sl);
        __syncthreads();
    }
    fft_regs<N, P, C::R1, C::R2, C::R3, +1>(v, t, sm, sl, A.tw);
#pragma unroll
    for (int q = 0; q < P; ++q)
        if (rvalid) A.work[row_local + t + T * q] = v[q];
}

// ---------------------------------------------------------------------------
// rows, forward: c <- z on work[a][b][:], then optional store + P(k) binning
// (auto |S|^2 or cross Re S conj(X)).   grid = ceil(na*N / RB)
// ---------------------------------------------------------------------------
template <int N, bool DOFFT>
__global__ void __launch_bounds__(RowGeom<N>::THREADS, FB_ROWS_MINB) k_rows_fwd(const RowsArgs A) {
    using G = RowGeom<N>;
    using C = FftCfg<N>;
    constexpr int P = C::P, T = C::T;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2* sm = reinterpret_cast<float2*>(smem_raw);

    const int rl = threadIdx.x / T, t = threadIdx.x % T;
    const long row_raw = (long)blockIdx.x * G::RB + rl;
    const bool rvalid = row_raw < A.nrows;
    const long row_id = rvalid ? row_raw : A.nrows - 1;
    const int al = (int)(row_id / N);
    const int a = A.K.a0 + al;
    const int b = (int)(row_id % N);
    const bool do_pk = (A.flags & FB_F_PK) != 0;
    const bool poles = (A.flags & FB_F_POLES) != 0;
    const size_t row_local = ((size_t)al * N + b) * N;
    if constexpr ((G::RB * N * sizeof(float2)) % 16 == 0) {
        const long prow = ((long)blockIdx.x + A.pf_dist) * G::RB;       // RB contiguous rows of the block pf_dist ahead
        if (A.pf_dist > 0 && threadIdx.x == 0 && prow + G::RB <= A.nrows) {
            l2_prefetch(A.work + (size_t)prow * N, G::RB * N * sizeof(float2));
            if (A.cross) l2_prefetch(A.cross + (size_t)prow * N, G::RB * N * sizeof(float2));
        }
    }
    float2 v[P];
#pragma unroll
    for (int q = 0; q < P; ++q) v[q] = A.work[row_local + t + T * q];
    PkTables tb;
    if (do_pk) tb = pk_stage_tables<N>(A.K, reinterpret_cast<double*>(smem_raw + G::FFT_SMEM));
    RowLayout<N> sl{rl * RowLayout<N>::ROW};
    if constexpr (DOFFT) fft_regs<N, P, C::R1, C::R2, C::R3, -1>(v, t, sm, sl, A.tw);   // else: binning only
    const bool full_cube = (A.flags & FB_F_FULLCUBE_INTERNAL) != 0;
    if (T == 1 && do_pk) __syncthreads();
    if constexpr (T > 1) {
        __syncthreads();
        fft_store_natural<N, P>(v, t, sm, sl);                       // natural order in smem
        __syncthreads();
    }
    if (A.spec_out && rvalid) {                                      // quad order: coalesced STG.128
#pragma unroll
        for (int j = 0; j < P / 4; ++j) {
            const int cq = 4 * t + 4 * T * j;
            float2 h[4];
            const float2* pq = sm + sl(cq);
#pragma unroll
            for (int e = 0; e < 4; ++e) h[e] = (T > 1) ? pq[e] : v[4 * j + e];
            store_run<4>(A.spec_out + row_local + cq, h);
        }
    }
    if (do_pk) {                                                     // run order
        const int c0 = P * t;
        float2 hr[P];
        const float2* pr = sm + sl(c0);
#pragma unroll
        for (int e = 0; e < P; ++e) hr[e] = (T > 1) ? pr[e] : v[e];
        if (A.cross) {
            float2 x[P];
            load_run<P>(A.cross + row_local + c0, x);
            run_pk<N, P, (T > 1)>(A.K, tb, A.pk, a, b, c0, hr, x, poles, rvalid, full_cube);
        } else {
            run_pk<N, P, (T > 1)>(A.K, tb, A.pk, a, b, c0, hr, nullptr, poles, rvalid, full_cube);
        }
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
// (forward) this pass.  A SlabView splits the y index as y = d*ny + y' and places element
// (plane, y, z) at base[d] + (plane*ny + y')*N + z: one base pointer per rank d.  The bases may be
//   * blocks of one local buffer laid out [d][plane][y'][z] (contiguous per peer: an NCCL all-to-all
//     needs no pack / unpack pass), or
//   * the receive buffers of the PEER GPUs themselves, mapped over NVLink (fb_dist.cu): the stores of
//     the y pass are then the exchange -- compute and "collective" are one kernel.
// ny = 0 is the plain local [plane][y][z] layout at base[0].
#define FB_MAX_RANKS 8
struct SlabView {
    float2* base[FB_MAX_RANKS];
    int ny, ny_shift;
};
__device__ __forceinline__ float2* slab_ptr(const SlabView& v, int plane, int y, int N) {
    if (v.ny == 0) return v.base[0] + ((size_t)plane * N + y) * N;
    const int d = y >> v.ny_shift, yy = y & (v.ny - 1);
    return v.base[d] + ((size_t)plane * v.ny + yy) * N;
}

template <int N, int CZ, int S, bool SLAB>
__global__ void __launch_bounds__(ColGeom<N, CZ>::THREADS) k_cols_c2c(const float2* __restrict__ in,
                                                                     float2* __restrict__ out, const SlabView vin,
                                                                     const SlabView vout,
                                                                     const float2* __restrict__ tw, int pf_dist) {
    using C = FftCfg<N>;
    constexpr int P = C::P, T = C::T;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2* sm = reinterpret_cast<float2*>(smem_raw);
    const int col = threadIdx.x % CZ, t = threadIdx.x / CZ;
    const int plane = blockIdx.y;
    const size_t zc = (size_t)blockIdx.x * CZ + col;
    if constexpr (!SLAB) {
        if (pf_dist > 0) {                             // the tile of the CTA pf_dist blocks ahead -> L2
            const unsigned id = blockIdx.y * gridDim.x + blockIdx.x + pf_dist;
            const unsigned pp = id / gridDim.x, pz = id - pp * gridDim.x;
            if (pp < gridDim.y)
                for (int r = threadIdx.x; r < N; r += ColGeom<N, CZ>::THREADS)
                    l2_prefetch_line(in + ((size_t)pp * N + r) * N + (size_t)pz * CZ);
        }
    }
    float2 v[P];
    if constexpr (SLAB) {
#pragma unroll
        for (int q = 0; q < P; ++q) v[q] = slab_ptr(vin, plane, t + T * q, N)[zc];
    } else {
        const float2* p = in + ((size_t)plane * N + t) * N + zc;     // one 64-bit base, 32-bit constant offsets
#pragma unroll
        for (int q = 0; q < P; ++q) v[q] = p[(unsigned)(T * q) * (unsigned)N];
    }
    ColLayout<CZ> sl{col};
    fft_regs<N, P, C::R1, C::R2, C::R3, S>(v, t, sm, sl, tw);
    if constexpr (SLAB) {
#pragma unroll
        for (int q = 0; q < P; ++q) slab_ptr(vout, plane, t + T * q, N)[zc] = v[q];
    } else {
        float2* p = out + ((size_t)plane * N + t) * N + zc;
#pragma unroll
        for (int q = 0; q < P; ++q) p[(unsigned)(T * q) * (unsigned)N] = v[q];
    }
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
    const long* plane_off;  // optional: element offset of each kx plane (chunked exchange buffers)
    const float2* spec;
    float* field;
    const float* field_in;
    float2* spec_out;
    const float2* tw;
    size_t ncols;
    int flags;
    float scale;
    double* sums;           // [0] sum, [1] sum of squares
    // x r2c of a slab-decomposed run: plane k of the result belongs to rank d = min(k >> per_shift, nranks-1)
    // and is stored at peer_out[d] + (k - (d << per_shift)) * ncols (the peer's receive buffer, mapped over
    // NVLink); nranks = 0: everything goes to spec_out
    float2* peer_out[FB_MAX_RANKS];
    int nranks, per_shift;
};

template <int N, int CZ>
struct XGeom {
    static constexpr int M = N / 2;
    using C = FftCfg<M>;
    static constexpr int THREADS = CZ * C::T;
    static constexpr int MINB = THREADS >= 1024 ? 1 : (1024 / THREADS > 4 ? 4 : 1024 / THREADS);   // keep <= 64 regs
    static constexpr size_t SMEM = (size_t)(M + M / 16) * CZ * sizeof(float2);
};

template <int N, int CZ>
__global__ void __launch_bounds__(XGeom<N, CZ>::THREADS, XGeom<N, CZ>::MINB) k_x_c2r(const XArgs A) {
    constexpr int M = N / 2;
    using C = FftCfg<M>;
    constexpr int P = C::P, T = C::T;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2* sm = reinterpret_cast<float2*>(smem_raw);
    __shared__ double red[2][32];
    const int col = threadIdx.x % CZ, t = threadIdx.x / CZ;
    const size_t g = (size_t)blockIdx.x * CZ + col;
    const float2* src = A.spec + g;
    ColLayout<CZ> sl{col};
    float2 v[P];
    // each plane element is read from HBM once; the mirrored partner X[M-k] comes from the tile
    // parked in shared memory (plane M, the partner of k = 0, is outside the tile)
    if (A.plane_off) {                                 // planes gathered from several receive buffers
#pragma unroll
        for (int q = 0; q < P; ++q) v[q] = src[__ldg(&A.plane_off[t + T * q])];
    } else {
        const float2* pl = src + (size_t)t * A.ncols;
        const size_t lstride = (size_t)T * A.ncols;
#pragma unroll
        for (int q = 0; q < P; ++q, pl += lstride) v[q] = *pl;
    }
    fft_store_natural<M, P>(v, t, sm, sl);
    float2 xnyq = make_float2(0.f, 0.f);
    if (t == 0) xnyq = A.plane_off ? src[__ldg(&A.plane_off[M])] : src[(size_t)M * A.ncols];
    __syncthreads();
    // e^{-2 pi i k/N}, k = t + T q, from ONE table load: w_N^t times the compile-time rotation e^{-i pi q/P}
    const float2 w0 = FB_TW(A.tw, N, t);
#pragma unroll
    for (int q = 0; q < P; ++q) {
        const int k = t + T * q;
        const float2 xk = v[q];
        const float2 xm = cconj(k == 0 ? xnyq : sm[sl(M - k)]);
        float2 w = q == 0 ? w0 : rot_pi16(w0, q * (16 / P), -1);
        w.y = -w.y;                                    // e^{+2 pi i k / N}
        const float2 sp = cadd(xk, xm), df = cmul(csub(xk, xm), w);
        v[q] = make_float2(sp.x - df.y, sp.y + df.x);  // sp + i*df
    }
    if constexpr (C::R2 > 1) __syncthreads();          // the exchanges reuse the buffer
    fft_regs<M, P, C::R1, C::R2, C::R3, +1>(v, t, sm, sl, A.tw);
    float* dst = A.field + g + (size_t)(2 * t) * A.ncols;       // rows 2m, 2m+1 of m = t + T*q
    const size_t sstride = (size_t)(2 * T) * A.ncols;
    const bool do_exp = (A.flags & FB_F_EXP) != 0;
    float acc = 0.f, acc2 = 0.f;
    if (A.sums || do_exp) {
#pragma unroll
        for (int q = 0; q < P; ++q, dst += sstride) {
            float r0 = v[q].x * A.scale, r1 = v[q].y * A.scale;
            if (do_exp) {
                r0 = expf(r0);
                r1 = expf(r1);
            }
            acc += r0 + r1;
            acc2 = fmaf(r0, r0, fmaf(r1, r1, acc2));
            dst[0] = r0;
            dst[A.ncols] = r1;
        }
    } else {                                           // plain field: scale and store
#pragma unroll
        for (int q = 0; q < P; ++q, dst += sstride) {
            dst[0] = v[q].x * A.scale;
            dst[A.ncols] = v[q].y * A.scale;
        }
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
__global__ void __launch_bounds__(XGeom<N, CZ>::THREADS, XGeom<N, CZ>::MINB) k_x_r2c(const XArgs A) {
    constexpr int M = N / 2;
    using C = FftCfg<M>;
    constexpr int P = C::P, T = C::T;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2* sm = reinterpret_cast<float2*>(smem_raw);
    const int col = threadIdx.x % CZ, t = threadIdx.x / CZ;
    const size_t g = (size_t)blockIdx.x * CZ + col;
    const float* src = A.field_in + g + (size_t)(2 * t) * A.ncols;
    const size_t lstride = (size_t)(2 * T) * A.ncols;
    float2 v[P];
#pragma unroll
    for (int q = 0; q < P; ++q, src += lstride) v[q] = make_float2(src[0], src[A.ncols]);
    ColLayout<CZ> sl{col};
    fft_regs<M, P, C::R1, C::R2, C::R3, -1>(v, t, sm, sl, A.tw);
    __syncthreads();
    fft_store_natural<M, P>(v, t, sm, sl);
    __syncthreads();
    auto plane_ptr = [&](int k) -> float2* {
        if (A.nranks == 0) return A.spec_out + (size_t)k * A.ncols + g;
        const int d = min(k >> A.per_shift, A.nranks - 1);
        return A.peer_out[d] + (size_t)(k - (d << A.per_shift)) * A.ncols + g;
    };
    const float2 w0 = FB_TW(A.tw, N, t);               // one table load, compile-time rotations (see k_x_c2r)
#pragma unroll
    for (int q = 0; q < P; ++q) {
        const int k = t + T * q;
        const float2 zk = v[q];
        const float2 zm = cconj(sm[sl((M - k) & (M - 1))]);
        const float2 w = q == 0 ? w0 : rot_pi16(w0, q * (16 / P), -1);      // e^{-2 pi i k / N}
        const float2 sp = cadd(zk, zm), df = cmul(csub(zk, zm), w);
        // 1/2 (sp - i df)
        *plane_ptr(k) = make_float2(0.5f * (sp.x + df.y), 0.5f * (sp.y - df.x));
        if (k == 0) *plane_ptr(M) = make_float2(zk.x - zk.y, 0.f);
    }
}

}  // namespace fb
