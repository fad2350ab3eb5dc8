/*
 * fb_poisson.h -- Poisson sampling by inversion from ONE supplied uniform per voxel.
 *
 * Replaces np.random.poisson at fastbox/halos.py:116.  NumPy's legacy sampler
 * consumes a variable number of MT19937 draws per sample, so bit parity with it
 * cannot be defined; instead the count is a pure function of (lambda, u):
 *
 *     p_0 = exp(-lambda);  count = #{ k >= 0 : u > sum_{j<=k} p_j },
 *     p_k = p_{k-1} * lambda / k
 *
 * evaluated in IEEE float64 with round-to-nearest +, *, / only (no FMA
 * contraction, own exp).  The same header is compiled for the host (gcc
 * -ffp-contract=off, oracle/fb_oracle.c) and the device (explicit __d*_rn
 * intrinsics), and restated in NumPy (oracle/restate.py:poisson_from_uniform),
 * so all three agree bit for bit.  Valid for any lambda >= 0 (counts up to FB_POISSON_KMAX).
 */
#ifndef FB_POISSON_H
#define FB_POISSON_H

#include <math.h>
#include <stdint.h>

#ifdef __CUDA_ARCH__
#define FB_HD __host__ __device__ __forceinline__
#define FB_MUL(a, b) __dmul_rn((a), (b))
#define FB_ADD(a, b) __dadd_rn((a), (b))
#define FB_DIV(a, b) __ddiv_rn((a), (b))
#else
#ifdef __CUDACC__
#define FB_HD __host__ __device__ __forceinline__
#else
#define FB_HD static inline
#endif
#define FB_MUL(a, b) ((a) * (b))
#define FB_ADD(a, b) ((a) + (b))
#define FB_DIV(a, b) ((a) / (b))
#endif

/* exp(-lam) = pm * 2^(*e), lam >= 0: n = round(lam/ln2), r = -(lam - n ln2) in two pieces, degree-13 Taylor
 * for pm in [0.7, 1.42], *e = -n.  Keeping the exponent apart makes the inversion below valid for any lam
 * (exp(-lam) itself underflows from lam ~ 745 on). */
FB_HD double fb_exp_neg_scaled(double lam, int* e) {
    const double n = floor(FB_ADD(FB_MUL(lam, 1.4426950408889634), 0.5));
    double r = FB_ADD(FB_ADD(lam, -FB_MUL(n, 0.693147180369123816490e+00)), -FB_MUL(n, 1.90821492927058770002e-10));
    r = -r;
    const double c[13] = {1.0 / 479001600.0, 1.0 / 39916800.0, 1.0 / 3628800.0, 1.0 / 362880.0, 1.0 / 40320.0,
                          1.0 / 5040.0,      1.0 / 720.0,      1.0 / 120.0,     1.0 / 24.0,     1.0 / 6.0,
                          0.5,               1.0,              1.0};
    double p = 1.0 / 6227020800.0;
    for (int i = 0; i < 13; ++i) p = FB_ADD(FB_MUL(p, r), c[i]);
    *e = -(int)n;
    return p;
}
FB_HD double fb_exp_neg(double lam) {
    int e;
    const double p = fb_exp_neg_scaled(lam, &e);
    return ldexp(p, e);
}

#define FB_POISSON_KMAX (1 << 22)

/* lam >= 700 (p_0 underflows from ~745 on): the term p_k = pm * 2^e is carried as (mantissa, exponent), the
 * mantissa rescaled by exact powers of two, and the walk passes through terms that are still zero in double
 * precision until they count.  Sequential from k = 0: O(lam) steps, meant for the occasional dense voxel. */
FB_HD int32_t fb_poisson_inv(double lam, double u) {
    if (!(lam > 0.0)) return 0;
    /* Certain zero without evaluating exp: exp(-lam) > 1 - lam, and p_0 as computed below is within a few ulp
     * (< 1e-15) of exp(-lam), so u < (1 - lam) - 1e-14 implies u < p_0, for which the recurrence returns 0.  The
     * result is the one the full evaluation gives; sparse tracers (lam ~ 1e-2 per voxel) leave here 99 % of
     * the time, which takes the kernel from FP64-bound to HBM-bound. */
    if (lam < 0.5 && u < FB_ADD(FB_ADD(1.0, -lam), -1e-14)) return 0;
    if (lam < 700.0) {                               /* common case: the plain recurrence (no exponent bookkeeping) */
        double q = fb_exp_neg(lam);
        double c = q;
        int32_t j = 0;
        while (u > c && q > 0.0 && j < 100000) {
            ++j;
            q = FB_DIV(FB_MUL(q, lam), (double)j);
            c = FB_ADD(c, q);
        }
        return j;
    }
    int e;
    double pm = fb_exp_neg_scaled(lam, &e);
    double p = ldexp(pm, e);
    double cdf = p;
    int32_t k = 0;
    while (u > cdf && (p > 0.0 || (double)k < lam) && k < FB_POISSON_KMAX) {
        ++k;
        pm = FB_DIV(FB_MUL(pm, lam), (double)k);
        if (pm > 1.3407807929942597e+154) {          /* 2^512 */
            pm = FB_MUL(pm, 7.458340731200207e-155); /* 2^-512, exact */
            e += 512;
        } else if (pm < 7.458340731200207e-155) {
            pm = FB_MUL(pm, 1.3407807929942597e+154);
            e -= 512;
        }
        p = ldexp(pm, e);
        cdf = FB_ADD(cdf, p);
    }
    return k;
}

#endif
