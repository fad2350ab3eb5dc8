// fb_cols.cu -- strided passes: y columns (c2c) and the x axis real<->half-complex passes.
#include <stdlib.h>
#include "fb_launch.h"

namespace fb {

int env_int(const char* name, int dflt) {
    const char* s = getenv(name);
    return s && *s ? atoi(s) : dflt;
}

template <int N, int CZ>
static int launch_cols_t(fb_plan* p, float2* data, int nplanes, int sign) {
    using G = ColGeom<N, CZ>;
    dim3 grid(N / CZ, nplanes);
    if (sign < 0) {
        auto kern = k_cols_c2c<N, CZ, -1>;
        if (set_smem(kern, G::SMEM)) return -2;
        kern<<<grid, G::THREADS, G::SMEM, p->stream>>>(data, p->tw);
    } else {
        auto kern = k_cols_c2c<N, CZ, +1>;
        if (set_smem(kern, G::SMEM)) return -2;
        kern<<<grid, G::THREADS, G::SMEM, p->stream>>>(data, p->tw);
    }
    FB_LAUNCH_CHECK();
    return 0;
}

int launch_cols(fb_plan* p, float2* data, int nplanes, int sign) {
    switch (p->N) {
        case 8: return launch_cols_t<8, 8>(p, data, nplanes, sign);
        case 16: return launch_cols_t<16, 16>(p, data, nplanes, sign);
        case 32: return launch_cols_t<32, 16>(p, data, nplanes, sign);
        case 64: return launch_cols_t<64, 16>(p, data, nplanes, sign);
        case 128: return launch_cols_t<128, 16>(p, data, nplanes, sign);
        case 256: return launch_cols_t<256, 16>(p, data, nplanes, sign);
        case 512: return launch_cols_t<512, 16>(p, data, nplanes, sign);
        case 1024:
            if (env_int("FB_CZ_COLS", 16) == 8) return launch_cols_t<1024, 8>(p, data, nplanes, sign);
            return launch_cols_t<1024, 16>(p, data, nplanes, sign);
        case 2048: return launch_cols_t<2048, 8>(p, data, nplanes, sign);
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

static int launch_x(fb_plan* p, const XArgs& a, bool inv) {
    switch (p->N) {
        case 8: return launch_x_t<8, 32>(p, a, inv);
        case 16: return launch_x_t<16, 32>(p, a, inv);
        case 32: return launch_x_t<32, 32>(p, a, inv);
        case 64: return launch_x_t<64, 32>(p, a, inv);
        case 128: return launch_x_t<128, 32>(p, a, inv);
        case 256: return launch_x_t<256, 32>(p, a, inv);
        case 512: return launch_x_t<512, 32>(p, a, inv);
        case 1024:
            if (env_int("FB_CZ_X", 32) == 16) return launch_x_t<1024, 16>(p, a, inv);
            return launch_x_t<1024, 32>(p, a, inv);
        case 2048: return launch_x_t<2048, 16>(p, a, inv);
        default: set_error("unsupported N=%d", p->N); return -1;
    }
}

int launch_x_c2r(fb_plan* p, const XArgs& a) { return launch_x(p, a, true); }
int launch_x_r2c(fb_plan* p, const XArgs& a) { return launch_x(p, a, false); }

}  // namespace fb
