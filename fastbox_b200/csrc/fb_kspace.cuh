// fb_kspace.cuh -- per-mode k-space arithmetic fused into the first / last FFT pass:
// sqrt(P) lookup (box.py:161-176), transfer function (box.py:374-378), velocity /
// potential factors (box.py:254-274, 347), and the P(k) bin index + binned moments
// (box.py:741-764).
#pragma once
#include "fb_common.cuh"

namespace fb {

// sqrt(P) from the table uniform in log2(s): Catmull-Rom cubic through 4 neighbouring nodes
// (the table is a few tens of KB and stays in L1, unlike the exact integer LUT whose gathers go
// to L2).  The shim validates the interpolation error against the exact values (kspace.py).
__device__ __forceinline__ float sqrtp_logtable(const KSpace& K, float s) {
    if (!(s > 0.f)) return 0.f;                          // nan_to_num(P(0)) = 0, box.py:167
    float x = (__log2f(s) - K.log2s0) * K.inv_dlog2s;
    x = fminf(fmaxf(x, 1.f), (float)(K.sqrtp_n - 3) + 0.999f);
    const int i = (int)x;
    const float f = x - (float)i;
    const float p0 = __ldg(&K.sqrtp[i - 1]), p1 = __ldg(&K.sqrtp[i]), p2 = __ldg(&K.sqrtp[i + 1]),
                p3 = __ldg(&K.sqrtp[i + 2]);
    const float c3 = 3.f * (p1 - p2) + p3 - p0;
    const float c2 = 2.f * p0 - 5.f * p1 + 4.f * p2 - p3;
    return fmaf(0.5f * f, fmaf(f, fmaf(f, c3, c2), p2 - p0), p1);
}

// real multiplier for mode (a,b,c) (global indices); `cf` = index used for the k_par /
// dense filter lookup (c itself, or (N-c)%N when the factor at -k is wanted).
__device__ __forceinline__ float k_amp(const KSpace& K, int flags, int kind, int a, int b, int c, int cf) {
    const int N = K.N;
    const int ma = mode_number(a, N), mb = mode_number(b, N), mc = mode_number(c, N);
    float amp = 1.f;
    float s = 0.f;
    if ((flags & FB_F_SQRTPK) && K.sqrtp_mode == 1) {
        amp = __ldg(&K.sqrtp[ma * ma + mb * mb + mc * mc]);
    }
    if (kind != FB_KIND_PLAIN || ((flags & FB_F_SQRTPK) && K.sqrtp_mode == 2)) {
        s = (float)(ma * ma) * K.inv_lx2 + (float)(mb * mb) * K.inv_ly2 + (float)(mc * mc) * K.inv_lz2;
    }
    if ((flags & FB_F_SQRTPK) && K.sqrtp_mode == 2) amp = sqrtp_logtable(K, s);
    if (flags & FB_F_FILTER) {
        if (K.tdense)
            amp *= __ldg(&K.tdense[((size_t)a * N + b) * N + cf]);
        else
            amp *= __ldg(&K.tperp[a * N + b]) * __ldg(&K.tpar[cf]);
    }
    if (kind != FB_KIND_PLAIN) {
        const float k2 = 39.478417604357434f * s;       // (2 pi)^2 s
        const float ik2 = k2 > 0.f ? 1.f / k2 : 0.f;    // nan_to_num at k = 0, box.py:257-259
        float comp = 1.f;
        int m = 1;
        if (kind == FB_KIND_VEL_X) { comp = (float)ma * K.two_pi_over_lx; m = ma; }
        else if (kind == FB_KIND_VEL_Y) { comp = (float)mb * K.two_pi_over_ly; m = mb; }
        else if (kind == FB_KIND_VEL_Z) { comp = (float)mc * K.two_pi_over_lz; m = mc; }
        if (kind != FB_KIND_POTENTIAL && m == -N / 2) comp = 0.f;   // box.py:268-274
        amp *= comp * ik2;
    }
    return amp;
}

// ---- P(k) ------------------------------------------------------------------
struct PkShared {
    double thr[FB_MAX_EDGES];
    double s1[FB_MAX_EDGES + 1], s2[FB_MAX_EDGES + 1], l2[FB_MAX_EDGES + 1], l4[FB_MAX_EDGES + 1];
    unsigned long long cnt[FB_MAX_EDGES + 1];
};

__device__ __forceinline__ void pk_shared_init(PkShared& sh, const KSpace& K) {
    for (int i = threadIdx.x; i <= K.nedges; i += blockDim.x) {
        if (i < K.nedges) sh.thr[i] = K.thr[i];
        sh.s1[i] = 0.0; sh.s2[i] = 0.0; sh.l2[i] = 0.0; sh.l4[i] = 0.0;
        sh.cnt[i] = 0ull;
    }
}

// np.digitize(k, edges) (right=False) evaluated on s = |k|^2/(2 pi)^2:
// index = #{ j : thr[j] <= s }.   s is formed exactly like box.py:125-127,
// ((Kx/Lx)^2 + (Ky/Ly)^2) + (Kz/Lz)^2 in float64 (per-axis squares precomputed by NumPy).
__device__ __forceinline__ int pk_bin(const PkShared& sh, int nedges, double s) {
    int lo = 0, hi = nedges;                 // first j with thr[j] > s
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (sh.thr[mid] <= s) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// One mode per lane; lanes of a warp are aggregated by bin before touching smem.
// w = multiplicity of the lane's mode (1 or 2), p = power, mu2 = (k_par/k)^2.
__device__ __forceinline__ void pk_accumulate(PkShared& sh, int bin, float w, float p, float mu2, bool poles,
                                              bool valid) {
    const unsigned full = 0xffffffffu;
    unsigned todo = __ballot_sync(full, valid);
    const int lane = threadIdx.x & 31;
    const unsigned wi = (unsigned)(w + 0.5f);
    while (todo) {
        const int leader = __ffs(todo) - 1;
        const int lb = __shfl_sync(full, bin, leader);
        const bool mine = valid && (bin == lb);
        const unsigned grp = __ballot_sync(full, mine);
        const unsigned cnt = __reduce_add_sync(full, mine ? wi : 0u);
        const double pd = mine ? (double)p : 0.0;
        const double wp = (double)w * pd;
        double a1 = warp_sum(wp);
        double a2 = warp_sum(wp * pd);
        double b2 = 0.0, b4 = 0.0;
        if (poles) {
            const double m2 = (double)mu2;
            b2 = warp_sum(wp * (1.5 * m2 - 0.5));
            b4 = warp_sum(wp * ((35.0 * m2 * m2 - 30.0 * m2 + 3.0) * 0.125));
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

__device__ __forceinline__ void pk_shared_flush(PkShared& sh, const KSpace& K, const PkDev& out, bool poles) {
    __syncthreads();
    for (int i = threadIdx.x; i <= K.nedges; i += blockDim.x) {
        if (sh.cnt[i]) {
            atomicAdd(&out.count[i], sh.cnt[i]);
            atomicAdd(&out.sum1[i], sh.s1[i]);
            atomicAdd(&out.sum2[i], sh.s2[i]);
            if (poles) {
                atomicAdd(&out.l2[i], sh.l2[i]);
                atomicAdd(&out.l4[i], sh.l4[i]);
            }
        }
    }
}

}  // namespace fb
