// fb_dist.cu -- slab-decomposed (multi-GPU) pipelines with the exchange INSIDE the library: one process per
// GPU, every rank maps the exchange block of every other rank (CUDA IPC over NVLink / NVSwitch) and the FFT
// pass that precedes the transpose stores its output straight into the receive buffer of the rank that owns
// it.  The "all-to-all" is therefore the store phase of the y pass (inverse) or of the x r2c pass (forward):
// compute and collective are one kernel, tile by tile, and no NCCL call sits on the data path.  The P(k)
// moments travel the same way (each rank writes its folded histogram into a slot of every peer, then every
// rank sums the slots in rank order: deterministic, identical on all ranks).
//
// Ordering across GPUs: after its stores a rank raises an epoch flag in every peer's block (k_dist_signal,
// preceded by a system-scope fence); before consuming, a rank waits until all flags reach the epoch
// (k_dist_wait, bounded spin: a peer that never arrives sets an error word instead of hanging the GPU).
// Receive buffers are double buffered per step, so one barrier per transform is enough.
//
// The reference (fastbox/box.py) is single process; SURVEY section 8(e) defines this decomposition:
// spectrum planes kx split over the ranks, real space split along y, z (line of sight) never split.
#include <unistd.h>

#include "fb_launch.h"
#include "fb_tma.cuh"

#define FB_DIST_MAX_CHUNKS 16
#ifndef FB_DIST_XMODE_DEFAULT
#define FB_DIST_XMODE_DEFAULT 2
#endif
#define FB_DIST_HANDLE_BYTES 128

struct fb_dist_state {
    int rank, world;
    int per, per_shift;             // kx planes per rank (N/2/world), log2
    int with_forward;
    unsigned char* block;           // local exchange block (cudaMalloc, exported through CUDA IPC)
    size_t block_bytes;
    size_t off_flags, off_pkx, off_recv[2], off_fwd;
    size_t recv_bytes, fwd_bytes, pkx_slot;
    unsigned char* peer[FB_MAX_RANKS];      // mapped base of every rank's block (peer[rank] == block)
    bool ipc_opened[FB_MAX_RANKS];
    bool connected;
    unsigned long long epoch;       // barrier counter (identical on all ranks: calls are collective)
    unsigned long long inv_step, fwd_step;
    cudaStream_t aux;               // first-pass chunks run here, overlapping the peer stores of the y pass
    cudaEvent_t ev_rows[FB_DIST_MAX_CHUNKS], ev_start, ev_ydone;
    int* err_dev;                   // device error word (barrier time-out)
    int* err_host;                  // pinned
    int cz_cols;                    // columns per CTA of the exchanging y pass (64-byte rows on NVLink)
    double timeout_s;
    // exchange mechanism of the inverse transform:
    //   0 = the y pass stores straight into the peers' receive buffers (one kernel computes and exchanges);
    //   1 = the y pass writes per-destination blocks locally and the COPY ENGINES push them to the peers
    //       (cudaMemcpyAsync on one stream per peer): no SM is held by NVLink traffic, so the first pass of the
    //       next chunk overlaps the transfer of the previous one.
    int xmode;
    float2* send;                   // xmode 1: [chunk][dest][plane][y'][z] staging, same size as the work array
    size_t send_bytes;
    //   2 = as 1, but a small high-priority copy kernel (k_dist_push: a few CTAs per peer, 16-byte loads, 16-byte
    //       peer stores in fully contiguous 512-byte warp requests) pushes the blocks: NVLink sees large
    //       requests instead of the 64-byte rows of the y pass tiles, and the k-space kernels keep the SMs.
    //   3 = as 2, the copy kernel drives the bulk copy engine (cp.async.bulk) instead of the LSU.
    cudaStream_t cps[FB_MAX_RANKS];
    cudaEvent_t ev_y[FB_DIST_MAX_CHUNKS], ev_cp[FB_MAX_RANKS];
    cudaStream_t push;              // xmode 2: highest-priority stream of the copy kernel
    int push_ctas;                  // CTAs per peer
};

namespace fb {

struct HandleBlob {                 // FB_DIST_HANDLE_BYTES, opaque to the caller
    long long pid;
    unsigned long long ptr;
    int device;
    int pad;
    cudaIpcMemHandle_t ipc;
};
static_assert(sizeof(HandleBlob) <= FB_DIST_HANDLE_BYTES, "handle blob too large");

struct PeerBlocks {
    unsigned char* base[FB_MAX_RANKS];
};

static size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// xmode 2: contiguous blocks local -> peer, blockIdx.y = peer slot.  Eight independent 16-byte loads per thread
// in flight, then eight 16-byte stores; a warp request covers 512 contiguous bytes on both sides.
struct PushArgs {
    const uint4* src[FB_MAX_RANKS];
    uint4* dst[FB_MAX_RANKS];
    size_t n16;
};
__global__ void __launch_bounds__(512) k_dist_push(const PushArgs a) {
    const uint4* __restrict__ s = a.src[blockIdx.y];
    uint4* __restrict__ d = a.dst[blockIdx.y];
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    constexpr int UN = 8;
    for (; i + (UN - 1) * stride < a.n16; i += UN * stride) {
        uint4 v[UN];
#pragma unroll
        for (int u = 0; u < UN; ++u) v[u] = __ldcs(s + i + u * stride);
#pragma unroll
        for (int u = 0; u < UN; ++u) d[i + u * stride] = v[u];
    }
    for (; i < a.n16; i += stride) d[i] = __ldcs(s + i);
}

// xmode 3: the same transfer driven by ONE thread per CTA through the bulk copy engine of its SM
// (cp.async.bulk: local HBM -> shared memory ring -> peer HBM over NVLink).  No LSU instruction, no register
// traffic: the k-space kernels that share the SM keep their issue slots; the cost is 64 KB of shared memory.
#define FB_PUSH_STAGES 4
#define FB_PUSH_CHUNK 16384
__global__ void __launch_bounds__(32) k_dist_push_tma(const PushArgs a) {
    extern __shared__ __align__(128) unsigned char sm_push[];
    if (threadIdx.x != 0) return;
    uint64_t* bar = reinterpret_cast<uint64_t*>(sm_push + FB_PUSH_STAGES * FB_PUSH_CHUNK);
    const unsigned char* src = reinterpret_cast<const unsigned char*>(a.src[blockIdx.y]);
    unsigned char* dst = reinterpret_cast<unsigned char*>(a.dst[blockIdx.y]);
    const size_t total = a.n16 * sizeof(uint4);
    const size_t nchunks = (total + FB_PUSH_CHUNK - 1) / FB_PUSH_CHUNK;
    // this CTA moves chunks blockIdx.x, blockIdx.x + gridDim.x, ...
    const long nk = nchunks > blockIdx.x ? (long)((nchunks - blockIdx.x + gridDim.x - 1) / gridDim.x) : 0;
#pragma unroll
    for (int s = 0; s < FB_PUSH_STAGES; ++s) mbar_init(&bar[s], 1);
    mbar_fence_init();
    auto off_of = [&](long k) { return ((size_t)blockIdx.x + (size_t)k * gridDim.x) * FB_PUSH_CHUNK; };
    auto len_of = [&](long k) {
        const size_t off = off_of(k);
        return (uint32_t)(total - off < (size_t)FB_PUSH_CHUNK ? total - off : (size_t)FB_PUSH_CHUNK);
    };
    auto issue_load = [&](long k) {
        const int s = (int)(k % FB_PUSH_STAGES);
        const uint32_t len = len_of(k);
        mbar_expect_tx(&bar[s], len);
        bulk_load_1d(sm_push + (size_t)s * FB_PUSH_CHUNK, src + off_of(k), len, &bar[s]);
    };
    for (long k = 0; k < nk && k < FB_PUSH_STAGES; ++k) issue_load(k);
    for (long k = 0; k < nk; ++k) {
        const int s = (int)(k % FB_PUSH_STAGES);
        mbar_wait(&bar[s], (uint32_t)((k / FB_PUSH_STAGES) & 1));
        bulk_store_1d(dst + off_of(k), sm_push + (size_t)s * FB_PUSH_CHUNK, len_of(k));
        tma_commit();
        if (k >= 1 && k - 1 + FB_PUSH_STAGES < nk) {
            tma_wait_read<1>();                              // the store of chunk k-1 has read its stage
            issue_load(k - 1 + FB_PUSH_STAGES);
        }
    }
    tma_wait_all<0>();
}
static int launch_push(fb_plan* p, fb_dist_state* d, const PushArgs& pa, cudaStream_t st) {
    if (d->world <= 1) return 0;
    if (d->xmode == 3) {
        const size_t smem = (size_t)FB_PUSH_STAGES * FB_PUSH_CHUNK + 64;
        if (set_smem(k_dist_push_tma, smem)) return -2;
        k_dist_push_tma<<<dim3(d->push_ctas, d->world - 1), 32, smem, st>>>(pa);
    } else {
        k_dist_push<<<dim3(d->push_ctas, d->world - 1), 512, 0, st>>>(pa);
    }
    FB_LAUNCH_CHECK();
    return 0;
}

// flags[d][rank] = epoch on every peer d: "everything this rank stores for step `epoch` is visible"
__global__ void k_dist_signal(PeerBlocks peers, size_t off_flags, int rank, int world, unsigned long long epoch) {
    const int d = threadIdx.x;
    if (d < world) {
        __threadfence_system();
        volatile unsigned long long* f = reinterpret_cast<volatile unsigned long long*>(peers.base[d] + off_flags);
        f[rank] = epoch;
        __threadfence_system();
    }
}

// wait until every rank's flag in the LOCAL block has reached `epoch`; bounded (sets *err on time-out)
__global__ void k_dist_wait(const unsigned char* block, size_t off_flags, int world, unsigned long long epoch,
                            unsigned long long timeout_ns, int* err) {
    const int s = threadIdx.x;
    if (s < world) {
        const volatile unsigned long long* f = reinterpret_cast<const volatile unsigned long long*>(block + off_flags);
        unsigned long long t0, t1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
        while (f[s] < epoch) {
            __nanosleep(100);
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
            if (t1 - t0 > timeout_ns) {
                atomicExch(err, 1 + s);
                break;
            }
        }
        __threadfence_system();
    }
}

// fold the replicated local histogram (like k_pk_fold) and store the result into slot `rank` of every peer
__global__ void __launch_bounds__(32) k_pk_share(const unsigned long long* __restrict__ cnt,
                                                  const double* __restrict__ sums, PeerBlocks peers, size_t off_slot0,
                                                  size_t slot_bytes, int rank, int world) {
    const int i = blockIdx.x, lane = threadIdx.x;
    unsigned long long c = 0;
    double s[4] = {0.0, 0.0, 0.0, 0.0};
    for (int k = lane; k < FB_PK_COPIES; k += 32) {
        c += cnt[(size_t)k * (FB_MAX_EDGES + 1) + i];
#pragma unroll
        for (int j = 0; j < 4; ++j) s[j] += sums[((size_t)j * FB_PK_COPIES + k) * (FB_MAX_EDGES + 1) + i];
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        c += __shfl_down_sync(0xffffffffu, c, d);
#pragma unroll
        for (int j = 0; j < 4; ++j) s[j] += __shfl_down_sync(0xffffffffu, s[j], d);
    }
    c = __shfl_sync(0xffffffffu, c, 0);
#pragma unroll
    for (int j = 0; j < 4; ++j) s[j] = __shfl_sync(0xffffffffu, s[j], 0);
    if (lane < world) {                                   // lane d writes this rank's bin i into peer d
        unsigned char* slot = peers.base[lane] + off_slot0 + (size_t)rank * slot_bytes;
        reinterpret_cast<unsigned long long*>(slot)[i] = c;
        double* ds = reinterpret_cast<double*>(slot + (FB_MAX_EDGES + 1) * 8);
#pragma unroll
        for (int j = 0; j < 4; ++j) ds[(size_t)j * (FB_MAX_EDGES + 1) + i] = s[j];
    }
}

// sum the slots of all ranks in rank order -> folded layout of fb_plan::pk_fold
__global__ void k_pk_gather(const unsigned char* block, size_t off_slot0, size_t slot_bytes, int world, int n,
                            unsigned long long* cnt_out, double* sums_out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    unsigned long long c = 0;
    double s[4] = {0.0, 0.0, 0.0, 0.0};
    for (int r = 0; r < world; ++r) {
        const unsigned char* slot = block + off_slot0 + (size_t)r * slot_bytes;
        c += reinterpret_cast<const unsigned long long*>(slot)[i];
        const double* ds = reinterpret_cast<const double*>(slot + (FB_MAX_EDGES + 1) * 8);
#pragma unroll
        for (int j = 0; j < 4; ++j) s[j] += ds[(size_t)j * (FB_MAX_EDGES + 1) + i];
    }
    cnt_out[i] = c;
#pragma unroll
    for (int j = 0; j < 4; ++j) sums_out[(size_t)j * (FB_MAX_EDGES + 1) + i] = s[j];
}

void dist_destroy(fb_plan* p) {
    fb_dist_state* d = p->dist;
    if (!d) return;
    for (int r = 0; r < d->world; ++r)
        if (d->ipc_opened[r]) cudaIpcCloseMemHandle(d->peer[r]);
    if (d->block) cudaFree(d->block);
    if (d->aux) cudaStreamDestroy(d->aux);
    if (d->push) cudaStreamDestroy(d->push);
    for (int c = 0; c < FB_DIST_MAX_CHUNKS; ++c)
        if (d->ev_rows[c]) cudaEventDestroy(d->ev_rows[c]);
    if (d->ev_start) cudaEventDestroy(d->ev_start);
    if (d->ev_ydone) cudaEventDestroy(d->ev_ydone);
    if (d->err_dev) cudaFree(d->err_dev);
    if (d->err_host) cudaFreeHost(d->err_host);
    if (d->send) cudaFree(d->send);
    for (int r = 0; r < FB_MAX_RANKS; ++r) {
        if (d->cps[r]) cudaStreamDestroy(d->cps[r]);
        if (d->ev_cp[r]) cudaEventDestroy(d->ev_cp[r]);
    }
    for (int c = 0; c < FB_DIST_MAX_CHUNKS; ++c)
        if (d->ev_y[c]) cudaEventDestroy(d->ev_y[c]);
    free(d);
    p->dist = nullptr;
}

static PeerBlocks peer_blocks(const fb_dist_state* d) {
    PeerBlocks pb;
    memset(&pb, 0, sizeof(pb));
    for (int r = 0; r < d->world; ++r) pb.base[r] = d->peer[r];
    return pb;
}

static int dist_signal(fb_plan* p) {
    fb_dist_state* d = p->dist;
    ++d->epoch;
    k_dist_signal<<<1, 32, 0, p->stream>>>(peer_blocks(d), d->off_flags, d->rank, d->world, d->epoch);
    FB_LAUNCH_CHECK();
    return 0;
}
static int dist_wait(fb_plan* p) {
    fb_dist_state* d = p->dist;
    k_dist_wait<<<1, 32, 0, p->stream>>>(d->block, d->off_flags, d->world, d->epoch,
                                         (unsigned long long)(d->timeout_s * 1e9), d->err_dev);
    FB_LAUNCH_CHECK();
    return 0;
}
// after a host synchronisation: did a barrier time out?
static int dist_check_error(fb_plan* p) {
    fb_dist_state* d = p->dist;
    FB_CUDA(cudaMemcpyAsync(d->err_host, d->err_dev, sizeof(int), cudaMemcpyDeviceToHost, p->stream));
    FB_CUDA(cudaStreamSynchronize(p->stream));
    if (*d->err_host) {
        set_error("multi-GPU barrier timed out after %.1f s waiting for rank %d (epoch %llu)", d->timeout_s,
                  *d->err_host - 1, d->epoch);
        return -5;
    }
    return 0;
}

struct StreamScope {                                     // launchers use p->stream
    fb_plan* p;
    cudaStream_t saved;
    StreamScope(fb_plan* p_, cudaStream_t s) : p(p_), saved(p_->stream) { p->stream = s; }
    ~StreamScope() { p->stream = saved; }
};

}  // namespace fb

using namespace fb;

extern "C" {

int fb_dist_init(fb_plan* p, int rank, int world, int with_forward) {
    FB_CUDA(cudaSetDevice(p->device));
    FB_CHECK(p->dist == nullptr, "fb_dist_init: already initialised");
    FB_CHECK(world >= 1 && world <= FB_MAX_RANKS && (world & (world - 1)) == 0,
             "fb_dist_init: world=%d must be a power of two <= %d", world, FB_MAX_RANKS);
    FB_CHECK(rank >= 0 && rank < world, "fb_dist_init: rank %d out of range", rank);
    const int N = p->N;
    FB_CHECK((N / 2) % world == 0 && N / world >= 1, "fb_dist_init: world=%d must divide N/2=%d", world, N / 2);
    fb_dist_state* d = (fb_dist_state*)calloc(1, sizeof(fb_dist_state));
    FB_CHECK(d != nullptr, "out of host memory");
    d->rank = rank;
    d->world = world;
    d->per = (N / 2) / world;
    while ((1 << d->per_shift) < d->per) ++d->per_shift;
    d->with_forward = with_forward;
    const int ny = N / world;
    const int na = d->per + (rank == world - 1 ? 1 : 0);
    if (fb_plan_set_slab(p, rank * d->per, na, rank * ny, ny)) {
        free(d);
        return -1;
    }
    d->pkx_slot = align_up(5 * (size_t)(FB_MAX_EDGES + 1) * 8, 256);
    d->recv_bytes = align_up((size_t)(N / 2 + 1) * ny * N * sizeof(float2), 512);
    d->fwd_bytes = with_forward ? align_up((size_t)(d->per + 1) * N * N * sizeof(float2), 512) : 0;
    size_t off = 0;
    d->off_flags = off;
    off += 512;
    d->off_pkx = off;
    off += 4 * (size_t)world * d->pkx_slot;               // [0,1] inverse steps, [2,3] forward steps
    off = align_up(off, 512);
    for (int b = 0; b < 2; ++b) {
        d->off_recv[b] = off;
        off += d->recv_bytes;
    }
    d->off_fwd = off;
    off += d->fwd_bytes;
    d->block_bytes = off;
    p->dist = d;
    cudaError_t e = cudaMalloc((void**)&d->block, d->block_bytes);
    if (e != cudaSuccess) {
        set_error("fb_dist_init: cudaMalloc of the %.2f GB exchange block failed: %s", d->block_bytes / 1e9,
                  cudaGetErrorString(e));
        dist_destroy(p);
        return -2;
    }
    FB_CUDA(cudaMemset(d->block, 0, d->off_recv[0]));
    FB_CUDA(cudaStreamCreateWithFlags(&d->aux, cudaStreamNonBlocking));
    for (int c = 0; c < FB_DIST_MAX_CHUNKS; ++c) FB_CUDA(cudaEventCreateWithFlags(&d->ev_rows[c], cudaEventDisableTiming));
    FB_CUDA(cudaEventCreateWithFlags(&d->ev_start, cudaEventDisableTiming));
    FB_CUDA(cudaEventCreateWithFlags(&d->ev_ydone, cudaEventDisableTiming));
    FB_CUDA(cudaMalloc((void**)&d->err_dev, sizeof(int)));
    FB_CUDA(cudaMemset(d->err_dev, 0, sizeof(int)));
    FB_CUDA(cudaMallocHost((void**)&d->err_host, sizeof(int)));
    *d->err_host = 0;
    d->xmode = env_int("FB_DIST_XMODE", FB_DIST_XMODE_DEFAULT);
    for (int r = 0; r < world; ++r) {
        FB_CUDA(cudaStreamCreateWithFlags(&d->cps[r], cudaStreamNonBlocking));
        FB_CUDA(cudaEventCreateWithFlags(&d->ev_cp[r], cudaEventDisableTiming));
    }
    for (int c = 0; c < FB_DIST_MAX_CHUNKS; ++c) FB_CUDA(cudaEventCreateWithFlags(&d->ev_y[c], cudaEventDisableTiming));
    {
        int lo = 0, hi = 0;
        FB_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
        FB_CUDA(cudaStreamCreateWithPriority(&d->push, cudaStreamNonBlocking, hi));
        const int dflt = world > 1 ? (32 / (world - 1) > 4 ? 32 / (world - 1) : 4) : 1;   // 8 GPUs: 4 per peer (measured best)
        d->push_ctas = env_int("FB_DIST_PUSH_CTAS", dflt);
    }
    d->cz_cols = env_int("FB_DIST_CZ", N >= 2048 ? 8 : 0);
    d->timeout_s = (double)env_int("FB_DIST_TIMEOUT_S", 20);
    d->peer[rank] = d->block;
    d->connected = (world == 1);
    return 0;
}

int fb_dist_get_handle(fb_plan* p, void* handle_out) {
    FB_CHECK(p->dist && handle_out, "fb_dist_get_handle: call fb_dist_init first");
    FB_CUDA(cudaSetDevice(p->device));
    HandleBlob hb;
    memset(&hb, 0, sizeof(hb));
    hb.pid = (long long)getpid();
    hb.ptr = (unsigned long long)(uintptr_t)p->dist->block;
    hb.device = p->device;
    FB_CUDA(cudaIpcGetMemHandle(&hb.ipc, p->dist->block));
    memset(handle_out, 0, FB_DIST_HANDLE_BYTES);
    memcpy(handle_out, &hb, sizeof(hb));
    return 0;
}

int fb_dist_connect(fb_plan* p, const void* handles) {
    fb_dist_state* d = p->dist;
    FB_CHECK(d && handles, "fb_dist_connect: call fb_dist_init first");
    FB_CUDA(cudaSetDevice(p->device));
    for (int r = 0; r < d->world; ++r) {
        if (r == d->rank) continue;
        HandleBlob hb;
        memcpy(&hb, (const unsigned char*)handles + (size_t)r * FB_DIST_HANDLE_BYTES, sizeof(hb));
        if (hb.pid == (long long)getpid()) {
            // same process (several plans, e.g. tests): the pointer is valid as is; other device -> peer access
            if (hb.device != p->device) {
                cudaError_t e = cudaDeviceEnablePeerAccess(hb.device, 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) {
                    set_error("fb_dist_connect: peer access %d -> %d: %s", p->device, hb.device, cudaGetErrorString(e));
                    return -2;
                }
                cudaGetLastError();
            }
            d->peer[r] = (unsigned char*)(uintptr_t)hb.ptr;
        } else {
            void* mapped = nullptr;
            cudaError_t e = cudaIpcOpenMemHandle(&mapped, hb.ipc, cudaIpcMemLazyEnablePeerAccess);
            if (e != cudaSuccess) {
                set_error("fb_dist_connect: cudaIpcOpenMemHandle(rank %d) failed: %s", r, cudaGetErrorString(e));
                cudaGetLastError();
                return -2;
            }
            d->peer[r] = (unsigned char*)mapped;
            d->ipc_opened[r] = true;
        }
    }
    d->connected = true;
    return 0;
}

int fb_dist_info(fb_plan* p, int* rank, int* world, int* a0, int* na, int* y0, int* ny, size_t* block_bytes) {
    FB_CHECK(p->dist, "fb_dist_info: call fb_dist_init first");
    if (rank) *rank = p->dist->rank;
    if (world) *world = p->dist->world;
    if (a0) *a0 = p->a0;
    if (na) *na = p->na;
    if (y0) *y0 = p->y0;
    if (ny) *ny = p->ny;
    if (block_bytes) *block_bytes = p->dist->block_bytes;
    return 0;
}

// device-side barrier over all ranks + host synchronisation (set-up / tear-down / tests)
int fb_dist_barrier(fb_plan* p) {
    fb_dist_state* d = p->dist;
    FB_CHECK(d && d->connected, "fb_dist_barrier: not connected");
    FB_CUDA(cudaSetDevice(p->device));
    if (dist_signal(p)) return -3;
    if (dist_wait(p)) return -3;
    return dist_check_error(p);
}

// Realise (+ filter + P(k)) on the slab of this rank.  phase 0: whole step; 1: k-space passes, peer stores and
// the signal; 2: wait, x pass, reduced P(k).  (A caller driving several ranks from one thread -- tests on one
// box -- runs phase 1 on every rank before phase 2.)  `chunks` first-pass chunks overlap the peer stores.
//   field_out : DEVICE float32 [N][ny][N] (x, local y, z)
//   pk        : moments summed over ALL ranks (nullable)        sums_out: local sum / sum of squares (nullable)
int fb_dist_realise(fb_plan* p, uint64_t seed, int flags, float scale, int chunks, int phase, float* field_out,
                    fb_pk_result* pk, double* sums_out) {
    fb_dist_state* d = p->dist;
    FB_CHECK(d && d->connected, "fb_dist_realise: call fb_dist_init / fb_dist_connect first");
    FB_CHECK(phase >= 0 && phase <= 2, "fb_dist_realise: phase must be 0, 1 or 2");
    FB_CUDA(cudaSetDevice(p->device));
    const int N = p->N, ny = p->ny, na = p->na;
    if (pk) flags |= FB_F_PK; else flags &= ~(FB_F_PK | FB_F_POLES);
    if (phase != 2) {
        if (check_flags(p, flags)) return -1;
        if (ensure_work(p)) return -2;
        if (d->xmode != 0 && d->send_bytes < p->work_bytes) {
            if (d->send) FB_CUDA(cudaFree(d->send));
            d->send = nullptr;
            d->send_bytes = 0;
            FB_CUDA(cudaMalloc((void**)&d->send, p->work_bytes));
            d->send_bytes = p->work_bytes;
        }
        if (chunks < 1) chunks = 1;
        if (chunks > FB_DIST_MAX_CHUNKS) chunks = FB_DIST_MAX_CHUNKS;
        while (chunks > 1 && d->per % chunks) --chunks;
        const int nc = d->per / chunks;
        const int buf = (int)(d->inv_step & 1);
        if (pk && pk_clear(p)) return -2;
        cudaEventRecord(p->ev[0], p->stream);
        FB_CUDA(cudaEventRecord(d->ev_start, p->stream));
        FB_CUDA(cudaStreamWaitEvent(d->aux, d->ev_start, 0));
        // peers' receive buffers [kx][y'][z], shifted to this rank's first plane
        SlabView vout;
        memset(&vout, 0, sizeof(vout));
        vout.ny = ny;
        while ((1 << vout.ny_shift) < ny) ++vout.ny_shift;
        for (int c = 0; c < chunks; ++c) {
            const int pl0 = c * nc, npl = (c == chunks - 1) ? na - pl0 : nc;
            RowsArgs ra;
            memset(&ra, 0, sizeof(ra));
            ra.seed = seed;
            ra.work = p->work + (size_t)pl0 * N * N;
            ra.tw = p->tw;
            ra.nrows = (long)npl * N;
            ra.flags = flags;
            ra.kind = FB_KIND_PLAIN;
            ra.K = p->kspace();
            ra.K.a0 = p->a0 + pl0;
            ra.pk = p->pkdev();
            {
                StreamScope sc(p, d->aux);
                if (launch_rows_inv_philox(p, ra)) return -3;
            }
            FB_CUDA(cudaEventRecord(d->ev_rows[c], d->aux));
            FB_CUDA(cudaStreamWaitEvent(p->stream, d->ev_rows[c], 0));
            if (d->xmode == 0) {
                for (int r = 0; r < d->world; ++r)
                    vout.base[r] =
                        reinterpret_cast<float2*>(d->peer[r] + d->off_recv[buf]) + (size_t)(p->a0 + pl0) * ny * N;
                if (launch_cols_views(p, plain_view(p->work + (size_t)pl0 * N * N), vout, npl, +1, p->stream, d->cz_cols))
                    return -3;
            } else if (d->xmode >= 2) {
                // y pass into local per-destination blocks (this rank's own block goes straight to its receive
                // buffer), then the copy kernel pushes the other blocks on the high-priority stream
                float2* send_c = d->send + (size_t)pl0 * N * N;
                SlabView vsend = block_view(send_c, ny, npl, N);
                vsend.base[d->rank] =
                    reinterpret_cast<float2*>(d->block + d->off_recv[buf]) + (size_t)(p->a0 + pl0) * ny * N;
                if (launch_cols_views(p, plain_view(p->work + (size_t)pl0 * N * N), vsend, npl, +1, p->stream, 0)) return -3;
                FB_CUDA(cudaEventRecord(d->ev_y[c], p->stream));
                if (d->world > 1) {
                    const size_t blk = (size_t)npl * ny * N;
                    PushArgs pa;
                    memset(&pa, 0, sizeof(pa));
                    pa.n16 = blk * sizeof(float2) / sizeof(uint4);
                    for (int k = 0; k < d->world - 1; ++k) {
                        const int r = (d->rank + 1 + k) % d->world;
                        pa.src[k] = reinterpret_cast<const uint4*>(send_c + (size_t)r * blk);
                        pa.dst[k] = reinterpret_cast<uint4*>(reinterpret_cast<float2*>(d->peer[r] + d->off_recv[buf]) +
                                                             (size_t)(p->a0 + pl0) * ny * N);
                    }
                    FB_CUDA(cudaStreamWaitEvent(d->push, d->ev_y[c], 0));
                    if (launch_push(p, d, pa, d->push)) return -3;
                }
            } else {
                // y pass into local per-destination blocks, then one copy-engine transfer per peer
                float2* send_c = d->send + (size_t)pl0 * N * N;
                if (launch_cols_views(p, plain_view(p->work + (size_t)pl0 * N * N), block_view(send_c, ny, npl, N), npl,
                                      +1, p->stream, 0))
                    return -3;
                FB_CUDA(cudaEventRecord(d->ev_y[c], p->stream));
                const size_t blk = (size_t)npl * ny * N;
                for (int k = 0; k < d->world; ++k) {
                    const int r = (d->rank + 1 + k) % d->world;      // start with the neighbour: spreads the links
                    FB_CUDA(cudaStreamWaitEvent(d->cps[r], d->ev_y[c], 0));
                    float2* dst = reinterpret_cast<float2*>(d->peer[r] + d->off_recv[buf]) + (size_t)(p->a0 + pl0) * ny * N;
                    FB_CUDA(cudaMemcpyAsync(dst, send_c + (size_t)r * blk, blk * sizeof(float2), cudaMemcpyDeviceToDevice,
                                            d->cps[r]));
                }
            }
        }
        if (d->xmode == 1) {
            for (int r = 0; r < d->world; ++r) {
                FB_CUDA(cudaEventRecord(d->ev_cp[r], d->cps[r]));
                FB_CUDA(cudaStreamWaitEvent(p->stream, d->ev_cp[r], 0));
            }
        } else if (d->xmode >= 2) {
            FB_CUDA(cudaEventRecord(d->ev_cp[0], d->push));
            FB_CUDA(cudaStreamWaitEvent(p->stream, d->ev_cp[0], 0));
        }
        if (pk) {
            k_pk_share<<<p->nedges + 1, 32, 0, p->stream>>>(p->h_count, p->h_sums, peer_blocks(d),
                                                            d->off_pkx + (size_t)buf * d->world * d->pkx_slot,
                                                            d->pkx_slot, d->rank, d->world);
            FB_LAUNCH_CHECK();
        }
        if (dist_signal(p)) return -3;
        cudaEventRecord(p->ev[1], p->stream);
    }
    if (phase != 1) {
        FB_CHECK(field_out != nullptr && is_device_ptr(field_out), "fb_dist_realise: field_out must be device memory");
        const int buf = (int)(d->inv_step & 1);
        if (dist_wait(p)) return -3;
        cudaEventRecord(p->ev[2], p->stream);
        XArgs xa;
        memset(&xa, 0, sizeof(xa));
        xa.spec = reinterpret_cast<const float2*>(d->block + d->off_recv[buf]);
        xa.field = field_out;
        xa.tw = p->tw;
        xa.ncols = (size_t)ny * N;
        xa.flags = flags & FB_F_EXP;
        xa.scale = (float)((double)scale / ((double)N * N * N));
        xa.sums = sums_out ? p->scal : nullptr;
        if (sums_out && scal_clear(p)) return -2;
        if (launch_x_c2r(p, xa)) return -3;
        cudaEventRecord(p->ev[3], p->stream);
        p->n_last = 3;
        if (pk) {
            unsigned long long* dc = reinterpret_cast<unsigned long long*>(p->pk_fold);
            double* ds = reinterpret_cast<double*>(dc + (FB_MAX_EDGES + 1));
            const int n = p->nedges + 1;
            k_pk_gather<<<(n + 127) / 128, 128, 0, p->stream>>>(d->block, d->off_pkx + (size_t)buf * d->world * d->pkx_slot,
                                                               d->pkx_slot, d->world, n, dc, ds);
            FB_LAUNCH_CHECK();
            if (pk_fetch_folded(p, pk)) return -2;
        }
        if (scal_fetch(p, sums_out, 2)) return -2;
        ++d->inv_step;
        // rows of the next step overwrite `work`: they are ordered behind this step by ev_start
        return dist_check_error(p);
    }
    return 0;
}

int fb_dist_set_option(fb_plan* p, const char* key, int value) {
    fb_dist_state* d = p->dist;
    FB_CHECK(d && key, "fb_dist_set_option: call fb_dist_init first");
    if (!strcmp(key, "push_ctas")) {
        FB_CHECK(value >= 1 && value <= 64, "push_ctas must be in 1..64");
        d->push_ctas = value;
    } else if (!strcmp(key, "xmode")) {
        FB_CHECK(value >= 0 && value <= 3, "xmode must be 0..3");
        d->xmode = value;
    } else if (!strcmp(key, "cz_cols")) {
        d->cz_cols = value;
    } else {
        set_error("fb_dist_set_option: unknown key '%s'", key);
        return -1;
    }
    return 0;
}

// The exchange alone, for the NVLink roofline: y pass over the local planes (whatever `work` holds) storing into
// the peers' receive buffers, then the barrier; average milliseconds per iteration (CUDA events on the plan stream).
int fb_dist_bench_exchange(fb_plan* p, int iters, float* ms_out) {
    fb_dist_state* d = p->dist;
    FB_CHECK(d && d->connected && ms_out && iters >= 1, "fb_dist_bench_exchange: bad arguments / not connected");
    FB_CUDA(cudaSetDevice(p->device));
    if (ensure_work(p)) return -2;
    const int N = p->N, ny = p->ny, na = p->na;
    const int buf = (int)(d->inv_step & 1);
    SlabView vout;
    memset(&vout, 0, sizeof(vout));
    vout.ny = ny;
    while ((1 << vout.ny_shift) < ny) ++vout.ny_shift;
    for (int r = 0; r < d->world; ++r)
        vout.base[r] = reinterpret_cast<float2*>(d->peer[r] + d->off_recv[buf]) + (size_t)p->a0 * ny * N;
    // xmode 2: the copy kernel alone over the whole per-destination staging array (what the pipeline pushes per step)
    PushArgs pa;
    memset(&pa, 0, sizeof(pa));
    if (d->xmode >= 2) {
        FB_CHECK(d->send != nullptr, "fb_dist_bench_exchange: run fb_dist_realise once first (staging array)");
        // blocks as laid out by a one-chunk step: [dest][plane][y'][z], na planes each
        const size_t blk = (size_t)na * ny * N;
        pa.n16 = blk * sizeof(float2) / sizeof(uint4);
        for (int k = 0; k < d->world - 1; ++k) {
            const int r = (d->rank + 1 + k) % d->world;
            pa.src[k] = reinterpret_cast<const uint4*>(d->send + (size_t)r * blk);
            pa.dst[k] = reinterpret_cast<uint4*>(reinterpret_cast<float2*>(d->peer[r] + d->off_recv[buf]) +
                                                 (size_t)p->a0 * ny * N);
        }
    }
    float total = 0.f;
    for (int it = 0; it < iters + 1; ++it) {                 // first iteration = warm-up, aligns the ranks
        FB_CUDA(cudaEventRecord(p->ev[4], p->stream));
        if (d->xmode >= 2) {
            if (launch_push(p, d, pa, p->stream)) return -3;
        } else if (launch_cols_views(p, plain_view(p->work), vout, na, +1, p->stream, d->cz_cols)) {
            return -3;
        }
        if (dist_signal(p)) return -3;
        if (dist_wait(p)) return -3;
        FB_CUDA(cudaEventRecord(p->ev[5], p->stream));
        if (dist_check_error(p)) return -5;
        float ms = 0.f;
        FB_CUDA(cudaEventElapsedTime(&ms, p->ev[4], p->ev[5]));
        if (it > 0) total += ms;
    }
    *ms_out = total / (float)iters;
    return 0;
}

// Binned P(k) moments of the sharded real field (box.py:736-764): local x r2c storing each plane into the
// receive buffer of the rank that owns it -> barrier -> local y and z passes with the histogram epilogue ->
// moments through the peers' slots -> barrier -> sum.  phase as in fb_dist_realise.
int fb_dist_power_spectrum(fb_plan* p, const float* field, int flags, int phase, fb_pk_result* pk) {
    fb_dist_state* d = p->dist;
    FB_CHECK(d && d->connected && d->with_forward, "fb_dist_power_spectrum: fb_dist_init(with_forward=1) first");
    FB_CHECK(phase >= 0 && phase <= 3, "fb_dist_power_spectrum: bad phase");
    FB_CHECK(pk != nullptr, "fb_dist_power_spectrum: pk is NULL");
    FB_CUDA(cudaSetDevice(p->device));
    const int N = p->N, ny = p->ny, na = p->na;
    flags = (flags & FB_F_POLES) | FB_F_PK;
    const int buf = 2 + (int)(d->fwd_step & 1);
    // phases (for single-thread multi-rank drivers): 1 = x pass + signal, 2 = wait + k-space + share + signal,
    // 3 = wait + gather; 0 = all
    if (phase == 0 || phase == 1) {
        FB_CHECK(field != nullptr && is_device_ptr(field), "fb_dist_power_spectrum: field must be device memory");
        if (check_flags(p, flags)) return -1;
        if (ensure_work(p)) return -2;
        XArgs xa;
        memset(&xa, 0, sizeof(xa));
        xa.field_in = field;
        xa.tw = p->tw;
        xa.ncols = (size_t)ny * N;
        xa.nranks = d->world;
        xa.per_shift = d->per_shift;
        for (int r = 0; r < d->world; ++r) {
            const int na_r = d->per + (r == d->world - 1 ? 1 : 0);
            // rank r's forward buffer is [src][local plane][y'][z]; this rank is src = d->rank
            xa.peer_out[r] = reinterpret_cast<float2*>(d->peer[r] + d->off_fwd) + (size_t)d->rank * na_r * ny * N;
        }
        xa.spec_out = xa.peer_out[0];
        cudaEventRecord(p->ev[0], p->stream);
        if (launch_x_r2c(p, xa)) return -3;
        if (dist_signal(p)) return -3;
        cudaEventRecord(p->ev[1], p->stream);
    }
    if (phase == 0 || phase == 2) {
        if (dist_wait(p)) return -3;
        if (pk_clear(p)) return -2;
        float2* fwd = reinterpret_cast<float2*>(d->block + d->off_fwd);
        if (launch_cols_views(p, block_view(fwd, ny, na, N), plain_view(p->work), na, -1, p->stream, 0)) return -3;
        RowsArgs ra;
        memset(&ra, 0, sizeof(ra));
        ra.work = p->work;
        ra.tw = p->tw;
        ra.nrows = (long)na * N;
        ra.flags = flags;
        ra.K = p->kspace();
        ra.pk = p->pkdev();
        if (launch_rows_fwd(p, ra)) return -3;
        k_pk_share<<<p->nedges + 1, 32, 0, p->stream>>>(p->h_count, p->h_sums, peer_blocks(d),
                                                        d->off_pkx + (size_t)buf * d->world * d->pkx_slot, d->pkx_slot,
                                                        d->rank, d->world);
        FB_LAUNCH_CHECK();
        if (dist_signal(p)) return -3;
        cudaEventRecord(p->ev[2], p->stream);
    }
    if (phase == 0 || phase == 3) {
        if (dist_wait(p)) return -3;
        unsigned long long* dc = reinterpret_cast<unsigned long long*>(p->pk_fold);
        double* ds = reinterpret_cast<double*>(dc + (FB_MAX_EDGES + 1));
        const int n = p->nedges + 1;
        k_pk_gather<<<(n + 127) / 128, 128, 0, p->stream>>>(d->block, d->off_pkx + (size_t)buf * d->world * d->pkx_slot,
                                                           d->pkx_slot, d->world, n, dc, ds);
        FB_LAUNCH_CHECK();
        cudaEventRecord(p->ev[3], p->stream);
        p->n_last = 3;
        if (pk_fetch_folded(p, pk)) return -2;
        ++d->fwd_step;
        return dist_check_error(p);
    }
    return 0;
}

}  // extern "C"
