// fb_rows_inv.cu -- instantiations of the first inverse pass for one source kind.
// Compiled once per FB_SRC (0 noise, 1 philox, 2 spectrum, 3 cube) to keep builds parallel.
#include "fb_launch.h"

#ifndef FB_SRC
#error "compile with -DFB_SRC=0..3"
#endif


namespace fb {

#if FB_SRC == 0
#define FB_FN launch_rows_inv_noise
#elif FB_SRC == 1
#define FB_FN launch_rows_inv_philox
#elif FB_SRC == 2
#define FB_FN launch_rows_inv_spec
#else
#define FB_FN launch_rows_inv_cube
#endif

template <int N>
static int launch_n(fb_plan* p, const RowsArgs& a) {
    using G = RowGeom<N>;
    auto kern = k_rows_inv<N, FB_SRC>;
    if (set_smem(kern, G::SMEM_INV)) return -2;
    const long blocks = (a.nrows + G::RB - 1) / G::RB;
    RowsArgs b = a;
    b.pf_dist = env_int("FB_ROWS_PF", FB_ROWS_PF_DEFAULT);     // L2 prefetch distance in CTAs (noise and stored-spectrum sources)
    kern<<<(unsigned)blocks, G::THREADS, G::SMEM_INV, p->stream>>>(b);
    FB_LAUNCH_CHECK();
    return 0;
}

int FB_FN(fb_plan* p, const RowsArgs& a) {
#define FB_CASE(N_) return launch_n<N_>(p, a)
    FB_DISPATCH_N(p->N, FB_CASE);
#undef FB_CASE
    return 0;
}

}  // namespace fb
