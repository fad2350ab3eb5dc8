// fb_cols.cu -- strided passes: y columns (c2c) and the x axis real<->half-complex passes.
#include <stdlib.h>
#include "fb_launch.h"

namespace fb {

int env_int(const char* name, int dflt) {
    const char* s = getenv(name);
    return s && *s ? atoi(s) : dflt;
}

SlabView plain_view(float2* base) {
    SlabView v;
    memset(&v, 0, sizeof(v));
    v.base[0] = base;
    return v;
}

// [d][plane][y'][z] blocks of one local buffer (what an NCCL all-to-all sends / receives)
SlabView block_view(float2* base, int ny, int nplanes, int N) {
    SlabView v = plain_view(base);
    if (ny > 0) {
        v.ny = ny;
        v.ny_shift = 0;
        while ((1 << v.ny_shift) < ny) ++v.ny_shift;
        for (int d = 0; d < N / ny && d < FB_MAX_RANKS; ++d) v.base[d] = base + (size_t)d * nplanes * ny * N;
    }
    return v;
}

template <int N, int CZ>
static int launch_cols_t(fb_plan* p, const SlabView& vin, const SlabView& vout, int nplanes, int sign, cudaStream_t st) {
    using G = ColGeom<N, CZ>;
    dim3 grid(N / CZ, nplanes);
    const bool slab = vin.ny != 0 || vout.ny != 0;
    const int pf = env_int("FB_COLS_PF", FB_COLS_PF_DEFAULT);
#define FB_COLS_LAUNCH(SIGN_, SLAB_)                                                         \
    {                                                                                        \
        auto kern = k_cols_c2c<N, CZ, SIGN_, SLAB_>;                                         \
        if (set_smem(kern, G::SMEM)) return -2;                                              \
        kern<<<grid, G::THREADS, G::SMEM, st>>>(vin.base[0], vout.base[0], vin, vout, p->tw, pf); \
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
// exchange / store phases.  FB_CZ_COLS overrides for tuning; `cz_hint` > 0 is the caller's choice
// (the peer-store exchange prefers 64-byte rows on NVLink).
template <int N>
static int launch_cols_n(fb_plan* p, const SlabView& vin, const SlabView& vout, int nplanes, int sign, cudaStream_t st,
                         int dflt, int cz_hint) {
    const int cz = cz_hint > 0 ? cz_hint : env_int("FB_CZ_COLS", dflt);
    if (cz == 4) return launch_cols_t<N, 4>(p, vin, vout, nplanes, sign, st);
    if (cz == 8) return launch_cols_t<N, 8>(p, vin, vout, nplanes, sign, st);
    if constexpr (N <= 1024) {
        if (cz == 16) return launch_cols_t<N, 16>(p, vin, vout, nplanes, sign, st);
    }
    set_error("y pass: %d columns per CTA not available for N=%d", cz, N);
    return -1;
}

int launch_cols_views(fb_plan* p, const SlabView& vin, const SlabView& vout, int nplanes, int sign, cudaStream_t st,
                      int cz_hint) {
    if (nplanes <= 0) return 0;
    // plain layout on large grids: persistent TMA-pipelined kernel (fb_cols_tma.cu); FB_COLS_TMA=0 keeps the
    // per-thread LDG/STG kernel below
    if (vin.ny == 0 && vout.ny == 0 && cz_hint == 0 && env_int("FB_COLS_TMA", FB_COLS_TMA_DEFAULT)) {
        const int cz = env_int("FB_CZ_TMA", p->N == 512 ? 16 : (p->N == 1024 ? 8 : 4));
        if (cols_tma_available(p->N, cz)) return launch_cols_tma(p, vin.base[0], vout.base[0], nplanes, sign, cz, st);
    }
    switch (p->N) {
        case 8: return launch_cols_t<8, 8>(p, vin, vout, nplanes, sign, st);
        case 16: return launch_cols_t<16, 16>(p, vin, vout, nplanes, sign, st);
        case 32: return launch_cols_t<32, 16>(p, vin, vout, nplanes, sign, st);
        case 64: return launch_cols_t<64, 16>(p, vin, vout, nplanes, sign, st);
        case 128: return launch_cols_t<128, 16>(p, vin, vout, nplanes, sign, st);
        case 256: return launch_cols_n<256>(p, vin, vout, nplanes, sign, st, 16, cz_hint);
        case 512: return launch_cols_n<512>(p, vin, vout, nplanes, sign, st, 8, cz_hint);
        case 1024: return launch_cols_n<1024>(p, vin, vout, nplanes, sign, st, 8, cz_hint);
        case 2048: return launch_cols_n<2048>(p, vin, vout, nplanes, sign, st, 4, cz_hint);
        default: set_error("unsupported N=%d", p->N); return -1;
    }
}

int launch_cols(fb_plan* p, float2* data, int nplanes, int sign) { return launch_cols_ex(p, data, data, 0, 0, nplanes, sign); }

int launch_cols_ex(fb_plan* p, const float2* in, float2* out, int in_ny, int out_ny, int nplanes, int sign) {
    return launch_cols_views(p, block_view(const_cast<float2*>(in), in_ny, nplanes, p->N),
                             block_view(out, out_ny, nplanes, p->N), nplanes, sign, p->stream, 0);
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
