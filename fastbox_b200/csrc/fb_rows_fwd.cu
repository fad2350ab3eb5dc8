// fb_rows_fwd.cu -- last forward pass (z rows) with the P(k) epilogue, and the
// stand-alone P(k) moment kernel (box.py:741-764).
#include "fb_launch.h"

namespace fb {

template <int N, bool DOFFT>
static int launch_n(fb_plan* p, const RowsArgs& a) {
    using G = RowGeom<N>;
    auto kern = k_rows_fwd<N, DOFFT>;
    if (set_smem(kern, G::SMEM)) return -2;
    const long blocks = (a.nrows + G::RB - 1) / G::RB;
    RowsArgs b = a;
    b.pf_dist = env_int("FB_ROWS_PF", FB_ROWS_PF_DEFAULT);     // L2 prefetch distance in CTAs
    kern<<<(unsigned)blocks, G::THREADS, G::SMEM, p->stream>>>(b);
    FB_LAUNCH_CHECK();
    return 0;
}

int launch_rows_fwd(fb_plan* p, const RowsArgs& a) {
#define FB_CASE(N_) return launch_n<N_, true>(p, a)
    FB_DISPATCH_N(p->N, FB_CASE);
#undef FB_CASE
    return 0;
}

// binning only (no transform): the stored spectrum is read in place through the same epilogue
int launch_rows_pk_only(fb_plan* p, const RowsArgs& a) {
#define FB_CASE(N_) return launch_n<N_, false>(p, a)
    FB_DISPATCH_N(p->N, FB_CASE);
#undef FB_CASE
    return 0;
}

int launch_pk_spectrum(fb_plan* p, const float2* spec, const float2* cross, int nplanes, int full_cube, int flags) {
    RowsArgs ra;
    memset(&ra, 0, sizeof(ra));
    ra.work = const_cast<float2*>(spec);          // read only (no spec_out, no transform)
    ra.cross = cross;
    ra.tw = p->tw;
    ra.nrows = (long)nplanes * p->N;
    ra.flags = (flags & FB_F_POLES) | FB_F_PK | (full_cube ? FB_F_FULLCUBE_INTERNAL : 0);
    ra.K = p->kspace();
    ra.pk = p->pkdev();
    return launch_rows_pk_only(p, ra);
}

}  // namespace fb
