// fb_x_tma.cu -- the x passes (stride = plane size: the last pass of the inverse transform and the first pass
// of the forward transform) as persistent, TMA-pipelined kernels.
//
// One CTA per SM walks over tiles of CZ adjacent columns.  Inverse (half complex -> real), per tile:
//     TMA loads (cp.async.bulk.tensor.2d: <= 256 planes per box, plus one 1-row box for the Nyquist plane)
//         -> dense tile S[kx][column] in shared memory, completion on an mbarrier
//     registers <- S;  the mirrored partner X[M-k] of the real pre-processing is read from the same tile
//     prefetch of the NEXT tile into the other buffer (it lands during the transform)
//     register-resident Stockham FFT of length M = N/2, exchanges through the padded buffer X
//     epilogue (scale, exp, sums) -> dense real tile F[x][column] in the buffer the spectrum tile came in
//     fence.proxy.async, TMA stores (bulk group) -> HBM
// The forward pass (real -> half complex) is the mirror image.  No global access of the tile data goes through
// the LSU / L1 data pipe (the limiter of k_x_c2r / k_x_r2c according to ncu), rows of the output tile leave as
// 64..128-byte bursts regardless of which thread produced them, and the load of tile i+1 / store of tile i-1
// overlap the transform of tile i.  Same arithmetic as the per-thread kernels (fb_passes.cuh): bit-identical.
// Plain single-GPU layouts only; chunked receive buffers (plane_off) and peer stores keep the per-thread kernels.
#include "fb_launch.h"
#include "fb_tma.cuh"

namespace fb {

template <int N, int CZ>
struct XTmaGeom {
    static constexpr int M = N / 2;
    using C = FftCfg<M>;
    static constexpr int THREADS = CZ * C::T;
    static constexpr int BOX_ROWS = 256;
    static constexpr int NBOX_SPEC = M / BOX_ROWS;                                   // + one 1-row box (Nyquist plane)
    static constexpr int NBOX_FIELD = N / BOX_ROWS;
    static constexpr size_t TILE_SPEC = (size_t)(M + 1) * CZ * sizeof(float2);
    static constexpr size_t TILE_FIELD = (size_t)N * CZ * sizeof(float);             // fits inside TILE_SPEC
    static constexpr size_t TILE_PAD = (TILE_SPEC + 1023) / 1024 * 1024;
    static constexpr size_t XBUF = (size_t)(M + M / 16) * CZ * sizeof(float2);       // padded exchange buffer
    static constexpr size_t SMEM = 2 * TILE_PAD + XBUF + 64;
    static_assert(M % BOX_ROWS == 0, "TMA x pass needs N >= 512");
};

// ---- inverse: spec[a][g] (a = 0..M) -> field[x][g] --------------------------------------------------------------
template <int N, int CZ>
__global__ void __launch_bounds__(XTmaGeom<N, CZ>::THREADS, 1)
    k_x_c2r_tma(const __grid_constant__ CUtensorMap map_spec, const __grid_constant__ CUtensorMap map_nyq,
                const __grid_constant__ CUtensorMap map_field, int ntiles, const float2* __restrict__ tw, float scale,
                int flags, double* __restrict__ sums) {
    using G = XTmaGeom<N, CZ>;
    constexpr int M = N / 2;
    using C = FftCfg<M>;
    constexpr int P = C::P, T = C::T;
    extern __shared__ __align__(1024) unsigned char smem_tma[];
    __shared__ double red[2][32];
    float2* X = reinterpret_cast<float2*>(smem_tma + 2 * G::TILE_PAD);
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem_tma + 2 * G::TILE_PAD + G::XBUF);
    const int tid = threadIdx.x;
    const int col = tid % CZ, t = tid / CZ;
    if (tid == 0) {
        mbar_init(&bar[0], 1);
        mbar_init(&bar[1], 1);
        mbar_fence_init();
    }
    __syncthreads();
    auto issue_load = [&](int tile, int b) {                 // one thread
        float2* dst = reinterpret_cast<float2*>(smem_tma + (size_t)b * G::TILE_PAD);
        mbar_expect_tx(&bar[b], (uint32_t)G::TILE_SPEC);
#pragma unroll
        for (int j = 0; j < G::NBOX_SPEC; ++j)
            tma_load_2d(dst + (size_t)j * G::BOX_ROWS * CZ, &map_spec, tile * CZ * 2, j * G::BOX_ROWS, &bar[b]);
        tma_load_2d(dst + (size_t)M * CZ, &map_nyq, tile * CZ * 2, M, &bar[b]);
    };
    int tile = blockIdx.x;
    if (tid == 0 && tile < ntiles) issue_load(tile, 0);
    ColLayout<CZ> sl{col};
    const bool do_exp = (flags & FB_F_EXP) != 0;
    const bool want_sums = sums != nullptr;
    // a warp of a 16-column tile covers two values of t: the odd one stores its row pair in the opposite order so
    // that the two halves of the warp hit different shared-memory banks
    const int sw = (CZ < 32) ? (t & 1) : 0;
    float acc = 0.f, acc2 = 0.f;
    for (int it = 0; tile < ntiles; tile += gridDim.x, ++it) {
        const int b = it & 1;
        const float2* L = reinterpret_cast<const float2*>(smem_tma + (size_t)b * G::TILE_PAD);
        mbar_wait(&bar[b], (it >> 1) & 1);
        float2 v[P];
        {
            const float2* p = L + t * CZ + col;              // dense [plane][column]
#pragma unroll
            for (int q = 0; q < P; ++q) v[q] = p[q * T * CZ];
        }
        // Z[k] = (X[k] + conj X[M-k]) + i e^{+2 pi i k/N} (X[k] - conj X[M-k]); the partner of k = 0 is the Nyquist plane
#pragma unroll
        for (int q = 0; q < P; ++q) {
            const int k = t + T * q;
            const float2 xk = v[q];
            const float2 xm = cconj(L[(M - k) * CZ + col]);
            float2 w = FB_TW(tw, N, k);
            w.y = -w.y;
            const float2 sp = cadd(xk, xm), df = cmul(csub(xk, xm), w);
            v[q] = make_float2(sp.x - df.y, sp.y + df.x);
        }
        if (tid == 0) {
            const int next = tile + gridDim.x;
            if (next < ntiles) {
                tma_wait_read<0>();                          // the store of the previous tile has left the other buffer
                issue_load(next, b ^ 1);
            }
        }
        fft_regs<M, P, C::R1, C::R2, C::R3, +1>(v, t, X, sl, tw);   // >= 1 barrier: every thread is done reading L
        float* O = reinterpret_cast<float*>(smem_tma + (size_t)b * G::TILE_PAD);
        {
            float* o = O + (2 * t) * CZ + col;               // rows 2m, 2m+1 of m = t + T*q
#pragma unroll
            for (int q = 0; q < P; ++q) {
                float r0 = v[q].x * scale, r1 = v[q].y * scale;
                if (do_exp) {
                    r0 = expf(r0);
                    r1 = expf(r1);
                }
                if (want_sums) {
                    acc += r0 + r1;
                    acc2 = fmaf(r0, r0, fmaf(r1, r1, acc2));
                }
                float* oq = o + q * (2 * T) * CZ;
                oq[sw * CZ] = sw ? r1 : r0;
                oq[(1 - sw) * CZ] = sw ? r0 : r1;
            }
        }
        fence_proxy_async();
        __syncthreads();
        if (tid == 0) {
#pragma unroll
            for (int j = 0; j < G::NBOX_FIELD; ++j)
                tma_store_2d(&map_field, tile * CZ, j * G::BOX_ROWS, O + (size_t)j * G::BOX_ROWS * CZ);
            tma_commit();
        }
    }
    if (tid == 0) tma_wait_all<0>();                         // shared memory must outlive the last store
    if (want_sums) {
        double s1 = warp_sum((double)acc), s2 = warp_sum((double)acc2);
        const int warp = tid >> 5, lane = tid & 31;
        if (lane == 0) {
            red[0][warp] = s1;
            red[1][warp] = s2;
        }
        __syncthreads();
        if (warp == 0) {
            const int nw = (blockDim.x + 31) >> 5;
            s1 = lane < nw ? red[0][lane] : 0.0;
            s2 = lane < nw ? red[1][lane] : 0.0;
            s1 = warp_sum(s1);
            s2 = warp_sum(s2);
            if (lane == 0) {
                atomicAdd(&sums[0], s1);
                atomicAdd(&sums[1], s2);
            }
        }
    }
}

// ---- forward: field[x][g] -> spec[a][g] (a = 0..M) --------------------------------------------------------------
template <int N, int CZ>
__global__ void __launch_bounds__(XTmaGeom<N, CZ>::THREADS, 1)
    k_x_r2c_tma(const __grid_constant__ CUtensorMap map_field, const __grid_constant__ CUtensorMap map_spec,
                const __grid_constant__ CUtensorMap map_nyq, int ntiles, const float2* __restrict__ tw) {
    using G = XTmaGeom<N, CZ>;
    constexpr int M = N / 2;
    using C = FftCfg<M>;
    constexpr int P = C::P, T = C::T;
    extern __shared__ __align__(1024) unsigned char smem_tma[];
    float2* X = reinterpret_cast<float2*>(smem_tma + 2 * G::TILE_PAD);
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem_tma + 2 * G::TILE_PAD + G::XBUF);
    const int tid = threadIdx.x;
    const int col = tid % CZ, t = tid / CZ;
    if (tid == 0) {
        mbar_init(&bar[0], 1);
        mbar_init(&bar[1], 1);
        mbar_fence_init();
    }
    __syncthreads();
    auto issue_load = [&](int tile, int b) {                 // one thread
        float* dst = reinterpret_cast<float*>(smem_tma + (size_t)b * G::TILE_PAD);
        mbar_expect_tx(&bar[b], (uint32_t)G::TILE_FIELD);
#pragma unroll
        for (int j = 0; j < G::NBOX_FIELD; ++j)
            tma_load_2d(dst + (size_t)j * G::BOX_ROWS * CZ, &map_field, tile * CZ, j * G::BOX_ROWS, &bar[b]);
    };
    int tile = blockIdx.x;
    if (tid == 0 && tile < ntiles) issue_load(tile, 0);
    ColLayout<CZ> sl{col};
    const int sw = (CZ < 32) ? (t & 1) : 0;
    for (int it = 0; tile < ntiles; tile += gridDim.x, ++it) {
        const int b = it & 1;
        mbar_wait(&bar[b], (it >> 1) & 1);
        float2 v[P];
        {
            const float* in = reinterpret_cast<const float*>(smem_tma + (size_t)b * G::TILE_PAD) + (2 * t) * CZ + col;
#pragma unroll
            for (int q = 0; q < P; ++q) {                    // z[m] = in[2m] + i in[2m+1], m = t + T*q
                const float* iq = in + q * (2 * T) * CZ;
                const float f0 = iq[sw * CZ], f1 = iq[(1 - sw) * CZ];
                v[q] = sw ? make_float2(f1, f0) : make_float2(f0, f1);
            }
        }
        if (tid == 0) {
            const int next = tile + gridDim.x;
            if (next < ntiles) {
                tma_wait_read<0>();
                issue_load(next, b ^ 1);
            }
        }
        fft_regs<M, P, C::R1, C::R2, C::R3, -1>(v, t, X, sl, tw);   // >= 1 barrier: every thread is done reading the tile
        float2* Z = reinterpret_cast<float2*>(smem_tma + (size_t)b * G::TILE_PAD);
        {
            float2* p = Z + t * CZ + col;
#pragma unroll
            for (int q = 0; q < P; ++q) p[q * T * CZ] = v[q];
        }
        __syncthreads();
        float2 zm[P];
#pragma unroll
        for (int q = 0; q < P; ++q) zm[q] = cconj(Z[((M - (t + T * q)) & (M - 1)) * CZ + col]);
        __syncthreads();
        // X[k] = 1/2 (Z[k] + conj Z[M-k]) - i/2 e^{-2 pi i k/N} (Z[k] - conj Z[M-k]),  X[M] = Re Z[0] - Im Z[0]
#pragma unroll
        for (int q = 0; q < P; ++q) {
            const int k = t + T * q;
            const float2 zk = v[q];
            const float2 w = FB_TW(tw, N, k);
            const float2 sp = cadd(zk, zm[q]), df = cmul(csub(zk, zm[q]), w);
            Z[k * CZ + col] = make_float2(0.5f * (sp.x + df.y), 0.5f * (sp.y - df.x));
            if (k == 0) Z[M * CZ + col] = make_float2(zk.x - zk.y, 0.f);
        }
        fence_proxy_async();
        __syncthreads();
        if (tid == 0) {
#pragma unroll
            for (int j = 0; j < G::NBOX_SPEC; ++j)
                tma_store_2d(&map_spec, tile * CZ * 2, j * G::BOX_ROWS, Z + (size_t)j * G::BOX_ROWS * CZ);
            tma_store_2d(&map_nyq, tile * CZ * 2, M, Z + (size_t)M * CZ);
            tma_commit();
        }
    }
    if (tid == 0) tma_wait_all<0>();
}

template <int N, int CZ>
static int launch_t(fb_plan* p, const XArgs& a, bool inverse) {
    using G = XTmaGeom<N, CZ>;
    constexpr int M = N / 2;
    const float2* spec = inverse ? a.spec : a.spec_out;
    const float* field = inverse ? a.field : a.field_in;
    CUtensorMap mspec, mnyq, mfield;
    if (make_tensor_map_2d(&mspec, spec, 2ull * a.ncols, M + 1, a.ncols * sizeof(float2), 2 * CZ, G::BOX_ROWS)) return -2;
    if (make_tensor_map_2d(&mnyq, spec, 2ull * a.ncols, M + 1, a.ncols * sizeof(float2), 2 * CZ, 1)) return -2;
    if (make_tensor_map_2d(&mfield, field, a.ncols, N, a.ncols * sizeof(float), CZ, G::BOX_ROWS)) return -2;
    const int ntiles = (int)(a.ncols / CZ);
    const int ctas = p->sm_count < ntiles ? p->sm_count : ntiles;
    if (inverse) {
        auto kern = k_x_c2r_tma<N, CZ>;
        if (set_smem(kern, G::SMEM)) return -2;
        kern<<<ctas, G::THREADS, G::SMEM, p->stream>>>(mspec, mnyq, mfield, ntiles, a.tw, a.scale, a.flags, a.sums);
    } else {
        auto kern = k_x_r2c_tma<N, CZ>;
        if (set_smem(kern, G::SMEM)) return -2;
        kern<<<ctas, G::THREADS, G::SMEM, p->stream>>>(mfield, mspec, mnyq, ntiles, a.tw);
    }
    FB_LAUNCH_CHECK();
    return 0;
}

// the configurations the TMA x passes are instantiated for
bool x_tma_available(const fb_plan* p, const XArgs& a, bool inverse) {
    if (!cols_tma_available(1024, 8)) return false;                 // driver entry point for tensor maps
    if (a.plane_off != nullptr || a.nranks != 0) return false;      // chunked receive buffers / peer stores
    if (p->N == 1024) return a.ncols % 16 == 0;
    if (p->N == 512) return a.ncols % 32 == 0;
    (void)inverse;
    return false;
}

int launch_x_tma(fb_plan* p, const XArgs& a, bool inverse) {
    switch (p->N) {
        case 512: return launch_t<512, 32>(p, a, inverse);
        case 1024: return launch_t<1024, 16>(p, a, inverse);
        default: break;
    }
    set_error("TMA x pass: N=%d is not instantiated", p->N);
    return -1;
}

}  // namespace fb
