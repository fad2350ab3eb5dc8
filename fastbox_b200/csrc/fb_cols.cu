// fb_cols.cu -- strided passes: y columns (c2c) and the x axis real<->half-complex passes.
#include <stdlib.h>
#include "fb_launch.h"

namespace fb {

int env_int(const char* name, int dflt) {
    const char* s = getenv(name);
    return s && *s ? atoi(s) : dflt;
}

template <int N, int CZ>
static int launch_cols_t(fb_plan* p, const float2* in, float2* out, int in_ny, int out_ny, int nplanes, int sign) {
    using G = ColGeom<N, CZ>;
    dim3 grid(N / CZ, nplanes);
    const bool slab = in_ny != 0 || out_ny != 0;
    if constexpr (N >= 256) {
        // persistent cp.async-pipelined kernel (two tile buffers).  Measured SLOWER than the plain kernel on
        // B200 (2.25 ms vs 1.76 ms at 1024^3: 8-byte LDGSTS issue rate + one resident CTA), so it is
        // opt-in (FB_COLS_PIPE=1) and kept as the starting point for a TMA-tiled version.
        if (!slab && 2 * G::SMEM <= 220 * 1024 && env_int("FB_COLS_PIPE", 0)) {
            const int ntiles = (N / CZ) * nplanes;
            int per_sm = (int)((220 * 1024) / (2 * G::SMEM + 1024));
            per_sm = per_sm < 1 ? 1 : (per_sm > 4 ? 4 : per_sm);
            per_sm = env_int("FB_PIPE_CTAS", per_sm);
            const int want = p->sm_count * per_sm;
            const int ctas = want < ntiles ? want : ntiles;
            if (sign < 0) {
                auto kern = k_cols_c2c_pipe<N, CZ, -1>;
                if (set_smem(kern, 2 * G::SMEM)) return -2;
                kern<<<ctas, G::THREADS, 2 * G::SMEM, p->stream>>>(in, out, ntiles, p->tw);
            } else {
                auto kern = k_cols_c2c_pipe<N, CZ, +1>;
                if (set_smem(kern, 2 * G::SMEM)) return -2;
                kern<<<ctas, G::THREADS, 2 * G::SMEM, p->stream>>>(in, out, ntiles, p->tw);
            }
            FB_LAUNCH_CHECK();
            return 0;
        }
    }
#define FB_COLS_LAUNCH(SIGN_, SLAB_)                                                         \
    {                                                                                        \
        auto kern = k_cols_c2c<N, CZ, SIGN_, SLAB_>;                                         \
        if (set_smem(kern, G::SMEM)) return -2;                                              \
        kern<<<grid, G::THREADS, G::SMEM, p->stream>>>(in, out, in_ny, out_ny, p->tw);       \
    }
    if (sign < 0) {
        if (slab) FB_COLS_LAUNCH(-1, true) else FB_COLS_LAUNCH(-1, false)
    } else {
        if (slab) FB_COLS_LAUNCH(+1, true) else FB_COLS_LAUNCH(+1, false)
    }
#undef FB_COLS_LAUNCH
    FB_LAUNCH_CHECK();
    return 0;
}

// tile width (columns per CTA).  HBM3e on B200 sustains full bandwidth down to 32-byte row
// chunks (tools/probe.py), so narrow tiles are preferred: more CTAs per SM overlap load /
// exchange / store phases.  FB_CZ_COLS / FB_CZ_X override for tuning.
template <int N>
static int launch_cols_n(fb_plan* p, const float2* in, float2* out, int in_ny, int out_ny, int nplanes, int sign, int dflt) {
    const int cz = env_int("FB_CZ_COLS", dflt);
    if (cz == 4) return launch_cols_t<N, 4>(p, in, out, in_ny, out_ny, nplanes, sign);
    if (cz == 8) return launch_cols_t<N, 8>(p, in, out, in_ny, out_ny, nplanes, sign);
    if constexpr (N <= 1024) {
        if (cz == 16) return launch_cols_t<N, 16>(p, in, out, in_ny, out_ny, nplanes, sign);
    }
    set_error("FB_CZ_COLS=%d not available for N=%d", cz, N);
    return -1;
}

int launch_cols(fb_plan* p, float2* data, int nplanes, int sign) { return launch_cols_ex(p, data, data, 0, 0, nplanes, sign); }

int launch_cols_ex(fb_plan* p, const float2* in, float2* out, int in_ny, int out_ny, int nplanes, int sign) {
    switch (p->N) {
        case 8: return launch_cols_t<8, 8>(p, in, out, in_ny, out_ny, nplanes, sign);
        case 16: return launch_cols_t<16, 16>(p, in, out, in_ny, out_ny, nplanes, sign);
        case 32: return launch_cols_t<32, 16>(p, in, out, in_ny, out_ny, nplanes, sign);
        case 64: return launch_cols_t<64, 16>(p, in, out, in_ny, out_ny, nplanes, sign);
        case 128: return launch_cols_t<128, 16>(p, in, out, in_ny, out_ny, nplanes, sign);
        case 256: return launch_cols_n<256>(p, in, out, in_ny, out_ny, nplanes, sign, 16);
        case 512: return launch_cols_n<512>(p, in, out, in_ny, out_ny, nplanes, sign, 8);
        case 1024: return launch_cols_n<1024>(p, in, out, in_ny, out_ny, nplanes, sign, 8);
        case 2048: return launch_cols_n<2048>(p, in, out, in_ny, out_ny, nplanes, sign, 4);
        default: set_error("unsupported N=%d", p->N); return -1;
    }
}

template <int N, int CZ>
static int launch_x_t(fb_plan* p, const XArgs& a, bool inverse) {
    using G = XGeom<N, CZ>;
    if (a.ncols % CZ) {
        set_error("x pass: ncols=%zu not a multiple of %d", a.ncols, CZ);
        return -1;
    }
    const unsigned grid = (unsigned)(a.ncols / CZ);
    if (inverse) {
        auto kern = k_x_c2r<N, CZ>;
        if (set_smem(kern, G::SMEM)) return -2;
        kern<<<grid, G::THREADS, G::SMEM, p->stream>>>(a);
    } else {
        auto kern = k_x_r2c<N, CZ>;
        if (set_smem(kern, G::SMEM)) return -2;
        kern<<<grid, G::THREADS, G::SMEM, p->stream>>>(a);
    }
    FB_LAUNCH_CHECK();
    return 0;
}

template <int N>
static int launch_x_n(fb_plan* p, const XArgs& a, bool inv, int dflt) {
    const int cz = env_int("FB_CZ_X", dflt);
    if (cz == 8) return launch_x_t<N, 8>(p, a, inv);
    if (cz == 16) return launch_x_t<N, 16>(p, a, inv);
    if constexpr (N <= 1024) {
        if (cz == 32) return launch_x_t<N, 32>(p, a, inv);
    }
    set_error("FB_CZ_X=%d not available for N=%d", cz, N);
    return -1;
}

static int launch_x(fb_plan* p, const XArgs& a, bool inv) {
    switch (p->N) {
        case 8: return launch_x_t<8, 32>(p, a, inv);
        case 16: return launch_x_t<16, 32>(p, a, inv);
        case 32: return launch_x_t<32, 32>(p, a, inv);
        case 64: return launch_x_t<64, 32>(p, a, inv);
        case 128: return launch_x_t<128, 32>(p, a, inv);
        case 256: return launch_x_n<256>(p, a, inv, 32);
        case 512: return launch_x_n<512>(p, a, inv, 16);
        case 1024: return launch_x_n<1024>(p, a, inv, 16);
        case 2048: return launch_x_n<2048>(p, a, inv, 16);
        default: set_error("unsupported N=%d", p->N); return -1;
    }
}

int launch_x_c2r(fb_plan* p, const XArgs& a) { return launch_x(p, a, true); }
int launch_x_r2c(fb_plan* p, const XArgs& a) { return launch_x(p, a, false); }

}  // namespace fb
