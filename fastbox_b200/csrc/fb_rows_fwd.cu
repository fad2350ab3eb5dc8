// fb_rows_fwd.cu -- last forward pass (z rows) with the P(k) epilogue, and the
// stand-alone P(k) moment kernel (box.py:741-764).
#include "fb_launch.h"

namespace fb {

template <int N>
static int launch_n(fb_plan* p, const RowsArgs& a) {
    using G = RowGeom<N>;
    auto kern = k_rows_fwd<N>;
    if (set_smem(kern, G::SMEM)) return -2;
    const long blocks = (a.nrows + G::RB - 1) / G::RB;
    kern<<<(unsigned)blocks, G::THREADS, G::SMEM, p->stream>>>(a);
    FB_LAUNCH_CHECK();
    return 0;
}

int launch_rows_fwd(fb_plan* p, const RowsArgs& a) {
#define FB_CASE(N_) return launch_n<N_>(p, a)
    FB_DISPATCH_N(p->N, FB_CASE);
#undef FB_CASE
    return 0;
}

// ---------------------------------------------------------------------------
// P(k) moments from a stored spectrum, no transform.  One warp per row.
// full_cube: planes a in [0,N), weight 1 (arbitrary, possibly non-Hermitian cube).
// grid = (N/8, nplanes), block = 256
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_pk_spectrum(const float2* __restrict__ spec, const float2* __restrict__ cross,
                                                      int full_cube, int flags, KSpace K, PkDev out) {
    __shared__ PkShared pks;
    const int N = K.N;
    const bool poles = (flags & FB_F_POLES) != 0;
    pk_shared_init(pks, K);
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int al = blockIdx.y, a = K.a0 + al;
    const int b = blockIdx.x * 8 + warp;
    if (b < N) {
        const float wmult = (full_cube || a == 0 || a == N / 2) ? 1.f : 2.f;
        const double sab = __dadd_rn(K.ax[a], K.ay[b]);
        const size_t row = ((size_t)al * N + b) * N;
        for (int c0 = 0; c0 < N; c0 += 32) {
            const int c = c0 + lane;
            const bool valid = c < N;
            float p = 0.f, mu2 = 0.f;
            int bin = 0;
            if (valid) {
                const float2 h = spec[row + c];
                const double azc = K.az[c];
                const double s = __dadd_rn(sab, azc);
                bin = pk_bin(pks, K.nedges, s);
                if (cross) {
                    const float2 x = cross[row + c];
                    p = (h.x * x.x + h.y * x.y) * (float)K.inv_boxfactor;
                } else {
                    p = (h.x * h.x + h.y * h.y) * (float)K.inv_boxfactor;
                }
                mu2 = (poles && s > 0.0) ? (float)azc / (float)s : 0.f;
            }
            pk_accumulate(pks, bin, wmult, p, mu2, poles, valid);
        }
    }
    pk_shared_flush(pks, K, out, poles);
}


int launch_pk_spectrum(fb_plan* p, const float2* spec, const float2* cross, int nplanes, int full_cube, int flags) {
    KSpace K = p->kspace();
    dim3 grid((p->N + 7) / 8, nplanes);
    k_pk_spectrum<<<grid, 256, 0, p->stream>>>(spec, cross, full_cube, flags, K, p->pkdev());
    FB_LAUNCH_CHECK();
    return 0;
}

}  // namespace fb
