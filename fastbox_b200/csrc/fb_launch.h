// fb_launch.h -- launcher entry points shared between translation units.
#pragma once
#include "fb_passes.cuh"

namespace fb {

int launch_rows_inv_noise(fb_plan* p, const RowsArgs& a);
int launch_rows_inv_philox(fb_plan* p, const RowsArgs& a);
int launch_rows_inv_spec(fb_plan* p, const RowsArgs& a);
int launch_rows_inv_cube(fb_plan* p, const RowsArgs& a);
int launch_rows_fwd(fb_plan* p, const RowsArgs& a);
int launch_pk_spectrum(fb_plan* p, const float2* spec, const float2* cross, int nplanes, int full_cube, int flags);
int launch_cols(fb_plan* p, float2* data, int nplanes, int sign);
int launch_cols_ex(fb_plan* p, const float2* in, float2* out, int in_ny, int out_ny, int nplanes, int sign);
SlabView plain_view(float2* base);
SlabView block_view(float2* base, int ny, int nplanes, int N);
int launch_cols_views(fb_plan* p, const SlabView& vin, const SlabView& vout, int nplanes, int sign, cudaStream_t st,
                      int cz_hint);
bool cols_tma_available(int N, int cz);
bool tma_available();          // the driver exposes cuTensorMapEncodeTiled
int launch_cols_tma(fb_plan* p, const float2* in, float2* out, int nplanes, int sign, int cz, cudaStream_t st);
#ifndef FB_COLS_TMA_DEFAULT
#define FB_COLS_TMA_DEFAULT 1
#endif
// L2 prefetch distance in CTAs (0 = off): a CTA asks L2 for the input of the CTA that many blocks ahead.
// Measured at 1024^3 (profiles/README.md): first pass 3.06 -> 2.79 ms for any distance in 37..185 (worse beyond
// 300: the lines are evicted before use), per-thread y pass 1.68 -> 1.54 ms at 74; the same trick on the x
// passes (rows a whole plane apart: per-thread line prefetches as well as one bulk-tensor prefetch per CTA)
// cost 0.5-0.6 ms at every distance and was removed.
#ifndef FB_ROWS_PF_DEFAULT
#define FB_ROWS_PF_DEFAULT 74
#endif
#ifndef FB_BEAM_PF_DEFAULT
#define FB_BEAM_PF_DEFAULT 32      // beam x pass: contiguous Y / BS tiles of the CTA that many blocks ahead (17.2 -> 16.5 ms)
#endif
#ifndef FB_COLS_PF_DEFAULT
#define FB_COLS_PF_DEFAULT 74
#endif
int launch_x_c2r(fb_plan* p, const XArgs& a);
int launch_x_r2c(fb_plan* p, const XArgs& a);

int env_int(const char* name, int dflt);

#define FB_DISPATCH_N(N_, MACRO)                                 \
    switch (N_) {                                                \
        case 8: MACRO(8); break;                                 \
        case 16: MACRO(16); break;                               \
        case 32: MACRO(32); break;                               \
        case 64: MACRO(64); break;                               \
        case 128: MACRO(128); break;                             \
        case 256: MACRO(256); break;                             \
        case 512: MACRO(512); break;                             \
        case 1024: MACRO(1024); break;                           \
        case 2048: MACRO(2048); break;                           \
        default:                                                 \
            fb::set_error("unsupported grid size N=%d (power of two in [8,2048])", N_); \
            return -1;                                           \
    }

template <class Kern>
inline int set_smem(Kern kern, size_t smem) {
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) {
            set_error("cudaFuncSetAttribute(smem=%zu) failed: %s", smem, cudaGetErrorString(e));
            return -2;
        }
    }
    return 0;
}

}  // namespace fb
