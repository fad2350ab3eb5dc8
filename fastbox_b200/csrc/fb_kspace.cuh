// fb_kspace.cuh -- per-mode k-space multipliers fused into the first FFT pass: sqrt(P) lookup
// (box.py:161-176), transfer function (box.py:374-378), velocity / potential factors
// (box.py:254-274, 347).  The P(k) binning (box.py:741-764) lives in fb_passes.cuh.
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

// sqrt(P) from a table indexed by the leading bits of float(s) (exponent + M mantissa bits) with
// linear interpolation inside the segment: geometric node spacing without a log2.  The device
// table holds (node value, slope per mantissa step) pairs, so a lookup is one 8-byte load (L1
// resident) and one FFMA.  Validated on the host like the log2 table.  The _nz form does not
// special-case s = 0 (callers zero the k = 0 mode themselves).
__device__ __forceinline__ float sqrtp_bittable_nz(const KSpace& K, float s) {
    const unsigned key = __float_as_uint(s);
    int i = (int)(key >> K.bt_shift) - K.bt_base;
    i = max(0, min(i, K.sqrtp_n - 2));
    const float2 tp = __ldg(&K.sqrtp_pairs[i]);
    return fmaf((float)(key & K.bt_mask), tp.y, tp.x);
}
__device__ __forceinline__ float sqrtp_bittable(const KSpace& K, float s) {
    return s > 0.f ? sqrtp_bittable_nz(K, s) : 0.f;      // nan_to_num(P(0)) = 0, box.py:167
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
    if (kind != FB_KIND_PLAIN || ((flags & FB_F_SQRTPK) && K.sqrtp_mode >= 2)) {
        s = (float)(ma * ma) * K.inv_lx2 + (float)(mb * mb) * K.inv_ly2 + (float)(mc * mc) * K.inv_lz2;
    }
    if ((flags & FB_F_SQRTPK) && K.sqrtp_mode == 2) amp = sqrtp_logtable(K, s);
    if ((flags & FB_F_SQRTPK) && K.sqrtp_mode == 3) amp = sqrtp_bittable(K, s);
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

}  // namespace fb
