// fb_catalogue.cu -- halo counts per voxel -> catalogue of comoving positions,
// HaloDistribution.realise_halo_catalogue (fastbox/halos.py:120-176).
//
// Reference order (halos.py:145-160): for each distinct count value c >= 1 in ascending order, the
// voxels with Nhalo == c in C order (np.where), each repeated c times; positions are the voxel
// indices as float64, plus an optional uniform offset per coordinate (halos.py:163-166, drawn in
// catalogue order, x y z interleaved), times L/N per axis (halos.py:172-174).
//
// That order is a stable counting sort of the voxels by count value.  Device algorithm, three
// streaming passes over the counts (4 B/voxel each) and one scattered write of 24 B/halo:
//   1  k_cat_max      largest (and smallest) count -> K = max + 1 keys
//   2  k_cat_tiles<0> per tile of 4096 voxels: number of voxels per key -> hist[key][tile]
//      k_scan_*       exclusive scan of key * hist over (key major, tile minor) = first catalogue
//                     row of every (key, tile) group; grand total = number of halos
//   3  k_cat_tiles<1> per tile: stable rank of every voxel among its key inside the tile (warp
//                     match_any + running per-warp counters), then write its `key` rows
// Integer work throughout: the catalogue is bit-identical to the reference's for the same uniforms.
#include <cstdint>
#include "fb_launch.h"

namespace fb {

static inline unsigned grid_for(size_t n, int per_block, int sm_count) {
    const size_t want = (n + per_block - 1) / per_block;
    const size_t cap = (size_t)sm_count * 16;
    return (unsigned)(want < cap ? (want ? want : 1) : cap);
}

constexpr int CAT_TILE = 4096;      // voxels per CTA
constexpr int CAT_THREADS = 256;
constexpr int CAT_WARPS = CAT_THREADS / 32;
constexpr int CAT_KMAX = 1024;      // count values 0..1023 per voxel
constexpr int SCAN_BLOCK = 4096;    // scan elements per CTA

__global__ void __launch_bounds__(256) k_cat_max(const int32_t* __restrict__ counts, size_t n, int* __restrict__ mx,
                                                  int* __restrict__ mn) {
    int hi = 0, lo = 0;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    const int4* c4 = reinterpret_cast<const int4*>(counts);          // n = N^3 is a multiple of 4
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n / 4; i += stride) {
        const int4 c = __ldg(&c4[i]);
        hi = max(max(hi, c.x), max(c.y, max(c.z, c.w)));
        lo = min(min(lo, c.x), min(c.y, min(c.z, c.w)));
    }
    hi = __reduce_max_sync(0xffffffffu, hi);
    lo = __reduce_min_sync(0xffffffffu, lo);
    if ((threadIdx.x & 31) == 0) {
        if (hi > 0) atomicMax(mx, hi);
        if (lo < 0) atomicMin(mn, lo);
    }
}

// PHASE 0: hist[(key-1)*ntiles + tile] = voxels of this tile with that count.
// PHASE 1: write the catalogue rows of this tile (off = first row of each (key, tile) group).
// dynamic smem: wcount[CAT_WARPS][K] u32, then (PHASE 1) wbase[CAT_WARPS][K] u64
template <int PHASE>
__global__ void __launch_bounds__(CAT_THREADS) k_cat_tiles(const int32_t* __restrict__ counts, size_t n, int K,
                                                            size_t ntiles, uint32_t* __restrict__ hist,
                                                            const unsigned long long* __restrict__ off,
                                                            const double* __restrict__ uniforms,
                                                            double* __restrict__ cat, int log2n, double sx, double sy,
                                                            double sz) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint32_t* wcount = reinterpret_cast<uint32_t*>(smem_raw);
    unsigned long long* wbase = reinterpret_cast<unsigned long long*>(smem_raw + (size_t)CAT_WARPS * K * sizeof(uint32_t));
    const unsigned full = 0xffffffffu;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const size_t tile = blockIdx.x;
    const size_t v0 = tile * CAT_TILE + (size_t)warp * (CAT_TILE / CAT_WARPS);
    for (int i = threadIdx.x; i < CAT_WARPS * K; i += CAT_THREADS) wcount[i] = 0u;
    __syncthreads();
    uint32_t* mine = wcount + (size_t)warp * K;
    // pass A: voxels per key in this warp's 512 consecutive voxels (32 at a time, lane = voxel order)
    int keys[CAT_TILE / CAT_THREADS];
#pragma unroll
    for (int c = 0; c < CAT_TILE / CAT_THREADS; ++c) {
        const size_t v = v0 + (size_t)c * 32 + lane;
        keys[c] = v < n ? __ldg(&counts[v]) : 0;
    }
#pragma unroll
    for (int c = 0; c < CAT_TILE / CAT_THREADS; ++c) {
        const int key = keys[c];
        if (!__any_sync(full, key > 0)) continue;        // empty chunk (the common case for rare tracers)
        const unsigned m = __match_any_sync(full, key);
        if (key > 0 && lane == __ffs(m) - 1) mine[key] += __popc(m);
        __syncwarp();
    }
    __syncthreads();
    if constexpr (PHASE == 0) {
        for (int k = threadIdx.x + 1; k < K; k += CAT_THREADS) {
            uint32_t tot = 0;
#pragma unroll
            for (int w = 0; w < CAT_WARPS; ++w) tot += wcount[(size_t)w * K + k];
            hist[(size_t)(k - 1) * ntiles + tile] = tot;
        }
    } else {
        for (int k = threadIdx.x + 1; k < K; k += CAT_THREADS) {
            unsigned long long base = __ldg(&off[(size_t)(k - 1) * ntiles + tile]);
#pragma unroll
            for (int w = 0; w < CAT_WARPS; ++w) {
                const uint32_t c = wcount[(size_t)w * K + k];
                wbase[(size_t)w * K + k] = base;
                base += (unsigned long long)k * c;
                wcount[(size_t)w * K + k] = 0u;          // becomes the running counter of pass B
            }
        }
        __syncthreads();
        const unsigned long long* mybase = wbase + (size_t)warp * K;
        const unsigned lt = (1u << lane) - 1u;
        const int nmask = (1 << log2n) - 1;
        // second walk: the tile is re-read (L1 / L2 hot) instead of holding 16 keys in registers, and the
        // loop stays rolled -- unrolled it needed 128 registers and halved the streaming rate
#pragma unroll 1
        for (int c = 0; c < CAT_TILE / CAT_THREADS; ++c) {
            const size_t vv = v0 + (size_t)c * 32 + lane;
            const int key = vv < n ? __ldg(&counts[vv]) : 0;
            if (!__any_sync(full, key > 0)) continue;
            const unsigned m = __match_any_sync(full, key);
            uint32_t run = 0;
            if (key > 0) run = mine[key];
            __syncwarp();
            if (key > 0 && lane == __ffs(m) - 1) mine[key] = run + __popc(m);
            __syncwarp();
            if (key > 0) {
                const size_t v = v0 + (size_t)c * 32 + lane;
                const double ix = (double)(v >> (2 * log2n)), iy = (double)((v >> log2n) & nmask),
                             iz = (double)(v & nmask);
                unsigned long long row = mybase[key] + (unsigned long long)key * (run + __popc(m & lt));
                for (int j = 0; j < key; ++j, ++row) {
                    double x = ix, y = iy, z = iz;
                    if (uniforms) {                       // halos.py:166
                        x = __dadd_rn(x, __ldg(&uniforms[3 * row]));
                        y = __dadd_rn(y, __ldg(&uniforms[3 * row + 1]));
                        z = __dadd_rn(z, __ldg(&uniforms[3 * row + 2]));
                    }
                    cat[3 * row] = __dmul_rn(x, sx);      // halos.py:172-174
                    cat[3 * row + 1] = __dmul_rn(y, sy);
                    cat[3 * row + 2] = __dmul_rn(z, sz);
                }
            }
        }
    }
}

// ---- exclusive scan of w[e] = key(e) * hist[e], key(e) = e / ntiles + 1, in three kernels
__device__ __forceinline__ unsigned long long block_exclusive_scan(unsigned long long x, unsigned long long* total) {
    __shared__ unsigned long long wsum[32];
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    unsigned long long inc = x;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const unsigned long long y = __shfl_up_sync(full, inc, d);
        if (lane >= d) inc += y;
    }
    if (lane == 31) wsum[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        unsigned long long s = lane < nw ? wsum[lane] : 0ull, si = s;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const unsigned long long y = __shfl_up_sync(full, si, d);
            if (lane >= d) si += y;
        }
        wsum[lane] = si - s;                              // exclusive warp offsets
        if (lane == 31 && total) *total = si;
    }
    __syncthreads();
    const unsigned long long r = wsum[warp] + inc - x;
    __syncthreads();
    return r;
}

constexpr int SCAN_PER_THREAD = SCAN_BLOCK / 256;

__global__ void __launch_bounds__(256) k_scan_sums(const uint32_t* __restrict__ hist, size_t E, size_t ntiles,
                                                    unsigned long long* __restrict__ sums) {
    __shared__ unsigned long long tot;
    const size_t e0 = (size_t)blockIdx.x * SCAN_BLOCK + (size_t)threadIdx.x * SCAN_PER_THREAD;
    unsigned long long s = 0;
#pragma unroll
    for (int i = 0; i < SCAN_PER_THREAD; ++i) {
        const size_t e = e0 + i;
        if (e < E) s += (unsigned long long)(e / ntiles + 1) * hist[e];
    }
    block_exclusive_scan(s, &tot);
    if (threadIdx.x == 0) sums[blockIdx.x] = tot;
}

__global__ void __launch_bounds__(1024) k_scan_top(unsigned long long* __restrict__ sums, size_t nb,
                                                    unsigned long long* __restrict__ total) {
    __shared__ unsigned long long tot;
    unsigned long long carry = 0;
    for (size_t b0 = 0; b0 < nb; b0 += blockDim.x) {
        const size_t b = b0 + threadIdx.x;
        const unsigned long long x = b < nb ? sums[b] : 0ull;
        const unsigned long long ex = block_exclusive_scan(x, &tot);
        if (b < nb) sums[b] = carry + ex;
        carry += tot;
        __syncthreads();
    }
    if (threadIdx.x == 0) *total = carry;
}

__global__ void __launch_bounds__(256) k_scan_apply(const uint32_t* __restrict__ hist, size_t E, size_t ntiles,
                                                     const unsigned long long* __restrict__ sums,
                                                     unsigned long long* __restrict__ off) {
    const size_t e0 = (size_t)blockIdx.x * SCAN_BLOCK + (size_t)threadIdx.x * SCAN_PER_THREAD;
    unsigned long long w[SCAN_PER_THREAD], s = 0;
#pragma unroll
    for (int i = 0; i < SCAN_PER_THREAD; ++i) {
        const size_t e = e0 + i;
        w[i] = e < E ? (unsigned long long)(e / ntiles + 1) * hist[e] : 0ull;
        s += w[i];
    }
    unsigned long long run = sums[blockIdx.x] + block_exclusive_scan(s, nullptr);
#pragma unroll
    for (int i = 0; i < SCAN_PER_THREAD; ++i) {
        const size_t e = e0 + i;
        if (e < E) off[e] = run;
        run += w[i];
    }
}

}  // namespace fb

using namespace fb;

extern "C" int fb_halo_catalogue(fb_plan* p, const int32_t* counts, const double* uniforms, double* cat_out,
                                 uint64_t capacity, uint64_t* nhalo_out) {
    FB_CUDA(cudaSetDevice(p->device));
    FB_CHECK(counts && nhalo_out, "fb_halo_catalogue: NULL counts / nhalo_out");
    const int N = p->N;
    int log2n = 0;
    while ((1 << log2n) < N) ++log2n;
    const size_t n = (size_t)N * N * N;
    const void* dcounts = nullptr;
    if (stage_in(p, 0, counts, n * sizeof(int32_t), &dcounts)) return -2;
    const int32_t* dc = (const int32_t*)dcounts;

    // ---- 1. largest count
    int* d_mm = (int*)p->scal;                           // 64-byte device scratch: max, min, total
    FB_CUDA(cudaMemsetAsync(d_mm, 0, 2 * sizeof(int) + sizeof(unsigned long long), p->stream));
    k_cat_max<<<grid_for(n, 256, p->sm_count), 256, 0, p->stream>>>(dc, n, d_mm, d_mm + 1);
    FB_LAUNCH_CHECK();
    int h_mm[2];
    FB_CUDA(cudaMemcpyAsync(h_mm, d_mm, sizeof(h_mm), cudaMemcpyDeviceToHost, p->stream));
    FB_CUDA(cudaStreamSynchronize(p->stream));
    FB_CHECK(h_mm[1] >= 0, "fb_halo_catalogue: negative count %d", h_mm[1]);
    const int K = h_mm[0] + 1;
    FB_CHECK(K <= CAT_KMAX, "fb_halo_catalogue: count %d in one voxel exceeds the supported maximum %d", h_mm[0],
             CAT_KMAX - 1);
    if (K == 1) {
        *nhalo_out = 0;
        return 0;
    }

    // ---- 2. per-tile histograms and their scan
    const size_t ntiles = (n + CAT_TILE - 1) / CAT_TILE;
    const size_t E = (size_t)(K - 1) * ntiles;
    const size_t nb = (E + SCAN_BLOCK - 1) / SCAN_BLOCK;
    const size_t hist_bytes = (E * sizeof(uint32_t) + 15) & ~(size_t)15;
    if (ensure_aux(p, hist_bytes + (E + nb) * sizeof(unsigned long long))) return -2;
    uint32_t* hist = (uint32_t*)p->aux;
    unsigned long long* off = (unsigned long long*)((char*)p->aux + hist_bytes);
    unsigned long long* sums = off + E;
    unsigned long long* d_total = (unsigned long long*)(d_mm + 2);
    const size_t smem0 = (size_t)CAT_WARPS * K * sizeof(uint32_t);
    const size_t smem1 = smem0 + (size_t)CAT_WARPS * K * sizeof(unsigned long long);
    if (set_smem(k_cat_tiles<0>, smem0) || set_smem(k_cat_tiles<1>, smem1)) return -2;
    k_cat_tiles<0><<<(unsigned)ntiles, CAT_THREADS, smem0, p->stream>>>(dc, n, K, ntiles, hist, nullptr, nullptr, nullptr,
                                                                        log2n, 0.0, 0.0, 0.0);
    FB_LAUNCH_CHECK();
    k_scan_sums<<<(unsigned)nb, 256, 0, p->stream>>>(hist, E, ntiles, sums);
    FB_LAUNCH_CHECK();
    k_scan_top<<<1, 1024, 0, p->stream>>>(sums, nb, d_total);
    FB_LAUNCH_CHECK();
    k_scan_apply<<<(unsigned)nb, 256, 0, p->stream>>>(hist, E, ntiles, sums, off);
    FB_LAUNCH_CHECK();
    unsigned long long h_total = 0;
    FB_CUDA(cudaMemcpyAsync(&h_total, d_total, sizeof(h_total), cudaMemcpyDeviceToHost, p->stream));
    FB_CUDA(cudaStreamSynchronize(p->stream));
    *nhalo_out = h_total;
    if (!cat_out) return 0;                              // size query
    FB_CHECK(capacity >= h_total, "fb_halo_catalogue: catalogue has %llu rows, buffer holds %llu", h_total,
             (unsigned long long)capacity);

    // ---- 3. scatter
    const void* du = nullptr;
    void* dcat = nullptr;
    if (stage_in(p, 1, uniforms, (size_t)h_total * 3 * sizeof(double), &du)) return -2;
    if (stage_out_begin(p, 2, cat_out, (size_t)h_total * 3 * sizeof(double), &dcat)) return -2;
    k_cat_tiles<1><<<(unsigned)ntiles, CAT_THREADS, smem1, p->stream>>>(
        dc, n, K, ntiles, nullptr, off, (const double*)du, (double*)dcat, log2n, p->Lx / (double)N, p->Ly / (double)N,
        p->Lz / (double)N);
    FB_LAUNCH_CHECK();
    if (stage_out_end(p, 2, cat_out, (size_t)h_total * 3 * sizeof(double))) return -2;
    return 0;
}
