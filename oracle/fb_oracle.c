/*
 * TEST INFRASTRUCTURE ONLY -- plain-C restatement of the integer-valued pieces of the
 * hot path, compiled with gcc -ffp-contract=off.  Nothing under fastbox_b200/ links it.
 *
 *   fb_oracle_poisson : Poisson inversion from a supplied uniform (include/fb_poisson.h,
 *                       the same header the CUDA kernel compiles) -- stands in for
 *                       np.random.poisson at fastbox/halos.py:116 (parity unpinned w.r.t.
 *                       NumPy's stream; pinned against oracle/restate.py bit for bit).
 *   fb_oracle_digitize: bin index of every mode, k = 2 pi sqrt(((Kx/Lx)^2+(Ky/Ly)^2)+(Kz/Lz)^2)
 *                       then np.digitize(k, edges) (fastbox/box.py:119-127, 758).
 */
#include <math.h>
#include <stdint.h>
#include "../include/fb_poisson.h"

void fb_oracle_poisson(const double* lam, const double* u, long n, int32_t* out) {
    for (long i = 0; i < n; ++i) out[i] = fb_poisson_inv(lam[i], u[i]);
}

double fb_oracle_exp_neg(double lam) { return fb_exp_neg(lam); }

static int digitize1(double k, const double* edges, int ne) {
    int lo = 0, hi = ne;                 /* first j with edges[j] > k */
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (edges[mid] <= k) lo = mid + 1; else hi = mid;
    }
    return lo;
}

/* counts[ne+1] += 1 for every mode of the full N^3 grid */
void fb_oracle_digitize_counts(int N, double Lx, double Ly, double Lz, const double* edges, int ne, int64_t* counts) {
    const double two_pi = 2. * 3.14159265358979323846;
    for (int i = 0; i < N; ++i) {
        const double kx = (double)(i < N / 2 ? i : i - N) / Lx;
        for (int j = 0; j < N; ++j) {
            const double ky = (double)(j < N / 2 ? j : j - N) / Ly;
            const double sxy = kx * kx + ky * ky;
            for (int l = 0; l < N; ++l) {
                const double kz = (double)(l < N / 2 ? l : l - N) / Lz;
                const double k = two_pi * sqrt(sxy + kz * kz);
                counts[digitize1(k, edges, ne)] += 1;
            }
        }
    }
}
