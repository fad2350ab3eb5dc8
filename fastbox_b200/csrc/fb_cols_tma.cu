// fb_cols_tma.cu -- the y pass (columns, stride N) as a persistent, TMA-pipelined kernel.
//
// One CTA per SM walks over (plane, z-tile) tiles of N lines x CZ columns.  Per tile:
//     TMA load (cp.async.bulk.tensor.2d, <= 256 lines per box) -> dense tile L[b] in shared memory, mbarrier
//     registers <- L[b];  prefetch of the NEXT tile into L[b^1] is issued now (it lands during the transform)
//     register-resident Stockham FFT, exchanges through the padded buffer X (fb_fft.cuh)
//     registers -> L[b] (natural order), fence.proxy.async, TMA store (bulk group) -> HBM
// so that the load of tile i+1 and the store of tile i-1 are in flight while tile i is transformed: HBM never
// waits for the compute phases of the CTA, and no global access goes through the LSU / L1 data pipe.
// Plain [plane][y][z] layout only (the slab-decomposed layouts keep the per-thread kernel, whose stores are
// the NVLink exchange).  Same arithmetic as k_cols_c2c: results are bit-identical.
#include "fb_launch.h"
#include "fb_tma.cuh"

namespace fb {

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled encode_fn() {
    static PFN_encodeTiled fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* f = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (PFN_encodeTiled)f;
        else
            cudaGetLastError();
    }
    return fn;
}

int make_tensor_map_2d(CUtensorMap* map, const void* base, uint64_t inner, uint64_t rows, uint64_t row_stride_bytes,
                       uint32_t box_inner, uint32_t box_rows) {
    PFN_encodeTiled fn = encode_fn();
    if (!fn) {
        set_error("cuTensorMapEncodeTiled is not available from this driver");
        return -2;
    }
    const cuuint64_t dims[2] = {inner, rows};
    const cuuint64_t strides[1] = {row_stride_bytes};
    const cuuint32_t box[2] = {box_inner, box_rows};
    const cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed (%d): inner=%llu rows=%llu stride=%llu box=%ux%u", (int)r,
                  (unsigned long long)inner, (unsigned long long)rows, (unsigned long long)row_stride_bytes, box_inner,
                  box_rows);
        return -2;
    }
    return 0;
}

template <int N, int CZ>
struct ColTmaGeom {
    using C = FftCfg<N>;
    static constexpr int THREADS = CZ * C::T;
    static constexpr int BOX_ROWS = N < 256 ? N : 256;
    static constexpr int NBOX = N / BOX_ROWS;
    static constexpr size_t TILE = (size_t)N * CZ * sizeof(float2);                    // dense
    static constexpr size_t XBUF = (size_t)(N + N / 16) * CZ * sizeof(float2);         // padded exchange buffer
    static constexpr size_t SMEM = 2 * TILE + XBUF + 64;
};

template <int N, int CZ, int S>
__global__ void __launch_bounds__(ColTmaGeom<N, CZ>::THREADS, 1)
    k_cols_tma(const __grid_constant__ CUtensorMap map_in, const __grid_constant__ CUtensorMap map_out, int ntiles,
               const float2* __restrict__ tw) {
    using G = ColTmaGeom<N, CZ>;
    using C = FftCfg<N>;
    constexpr int P = C::P, T = C::T;
    constexpr int TPP = N / CZ;                              // tiles per plane
    extern __shared__ __align__(1024) unsigned char smem_tma[];
    float2* L0 = reinterpret_cast<float2*>(smem_tma);
    float2* L1 = reinterpret_cast<float2*>(smem_tma + G::TILE);
    float2* X = reinterpret_cast<float2*>(smem_tma + 2 * G::TILE);
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem_tma + 2 * G::TILE + G::XBUF);
    const int tid = threadIdx.x;
    const int col = tid % CZ, t = tid / CZ;
    if (tid == 0) {
        mbar_init(&bar[0], 1);
        mbar_init(&bar[1], 1);
        mbar_fence_init();
    }
    __syncthreads();
    auto issue_load = [&](int tile, int b) {                 // one thread
        const int plane = tile / TPP, zt = tile - plane * TPP;
        float2* dst = b ? L1 : L0;
        mbar_expect_tx(&bar[b], (uint32_t)G::TILE);
#pragma unroll
        for (int j = 0; j < G::NBOX; ++j)
            tma_load_2d(dst + (size_t)j * G::BOX_ROWS * CZ, &map_in, zt * CZ * 2, plane * N + j * G::BOX_ROWS, &bar[b]);
    };
    int tile = blockIdx.x;
    if (tid == 0 && tile < ntiles) issue_load(tile, 0);
    ColLayout<CZ> sl{col};
    for (int it = 0; tile < ntiles; tile += gridDim.x, ++it) {
        const int b = it & 1;
        float2* L = b ? L1 : L0;
        mbar_wait(&bar[b], (it >> 1) & 1);
        float2 v[P];
        {
            const float2* p = L + t * CZ + col;              // dense [line][column]: a warp reads 256 contiguous bytes
#pragma unroll
            for (int q = 0; q < P; ++q) v[q] = p[q * T * CZ];
        }
        if (tid == 0) {
            const int next = tile + gridDim.x;
            if (next < ntiles) {
                tma_wait_read<0>();                          // the store of the previous tile has left L[b^1]
                issue_load(next, b ^ 1);
            }
        }
        fft_regs<N, P, C::R1, C::R2, C::R3, S>(v, t, X, sl, tw);
        {
            float2* p = L + t * CZ + col;                    // everyone passed a barrier since reading L
#pragma unroll
            for (int q = 0; q < P; ++q) p[q * T * CZ] = v[q];
        }
        fence_proxy_async();
        __syncthreads();
        if (tid == 0) {
            const int plane = tile / TPP, zt = tile - plane * TPP;
#pragma unroll
            for (int j = 0; j < G::NBOX; ++j)
                tma_store_2d(&map_out, zt * CZ * 2, plane * N + j * G::BOX_ROWS, L + (size_t)j * G::BOX_ROWS * CZ);
            tma_commit();
        }
    }
    if (tid == 0) tma_wait_all<0>();                         // shared memory must outlive the last store
}

template <int N, int CZ>
static int launch_t(fb_plan* p, const float2* in, float2* out, int nplanes, int sign, cudaStream_t st) {
    using G = ColTmaGeom<N, CZ>;
    CUtensorMap min, mout;
    const uint64_t rows = (uint64_t)nplanes * N;
    if (make_tensor_map_2d(&min, in, 2ull * N, rows, (uint64_t)N * sizeof(float2), 2 * CZ, G::BOX_ROWS)) return -2;
    if (make_tensor_map_2d(&mout, out, 2ull * N, rows, (uint64_t)N * sizeof(float2), 2 * CZ, G::BOX_ROWS)) return -2;
    const int ntiles = nplanes * (N / CZ);
    const int per_sm = (int)((227 * 1024) / (G::SMEM + 1024)) > 0 ? (int)((227 * 1024) / (G::SMEM + 1024)) : 1;
    const int want = p->sm_count * per_sm;
    const int ctas = want < ntiles ? want : ntiles;
    if (sign < 0) {
        auto kern = k_cols_tma<N, CZ, -1>;
        if (set_smem(kern, G::SMEM)) return -2;
        kern<<<ctas, G::THREADS, G::SMEM, st>>>(min, mout, ntiles, p->tw);
    } else {
        auto kern = k_cols_tma<N, CZ, +1>;
        if (set_smem(kern, G::SMEM)) return -2;
        kern<<<ctas, G::THREADS, G::SMEM, st>>>(min, mout, ntiles, p->tw);
    }
    FB_LAUNCH_CHECK();
    return 0;
}

bool tma_available() { return encode_fn() != nullptr; }

bool cols_tma_available(int N, int cz) {
    if (!encode_fn()) return false;
    if (N == 512) return cz == 8 || cz == 16;
    if (N == 1024) return cz == 4 || cz == 8;
    if (N == 2048) return cz == 4;
    return false;
}

// plain layout only; returns -1 with an error set for unsupported (N, cz)
int launch_cols_tma(fb_plan* p, const float2* in, float2* out, int nplanes, int sign, int cz, cudaStream_t st) {
    if (nplanes <= 0) return 0;
    switch (p->N) {
        case 512:
            if (cz == 8) return launch_t<512, 8>(p, in, out, nplanes, sign, st);
            if (cz == 16) return launch_t<512, 16>(p, in, out, nplanes, sign, st);
            break;
        case 1024:
            if (cz == 4) return launch_t<1024, 4>(p, in, out, nplanes, sign, st);
            if (cz == 8) return launch_t<1024, 8>(p, in, out, nplanes, sign, st);
            break;
        case 2048:
            if (cz == 4) return launch_t<2048, 4>(p, in, out, nplanes, sign, st);
            break;
        default: break;
    }
    set_error("TMA y pass: N=%d with %d columns per tile is not instantiated", p->N, cz);
    return -1;
}

}  // namespace fb
