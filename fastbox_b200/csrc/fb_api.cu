// fb_api.cu -- C ABI: plan management, host staging, the realise / spectrum pipelines.
#include <math.h>
#include <stdarg.h>
#include <stdlib.h>
#include <vector>

#include "fb_launch.h"

namespace fb {

static thread_local char g_err[1024] = "";
std::atomic<uint64_t> g_launches{0};

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

bool is_device_ptr(const void* ptr) {
    if (!ptr) return false;
    cudaPointerAttributes at;
    cudaError_t e = cudaPointerGetAttributes(&at, ptr);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged;
}

static int ensure_slot(fb_plan* p, int slot, size_t bytes) {
    if (p->stage_bytes[slot] >= bytes) return 0;
    if (p->stage[slot]) FB_CUDA(cudaFree(p->stage[slot]));
    p->stage[slot] = nullptr;
    p->stage_bytes[slot] = 0;
    FB_CUDA(cudaMalloc(&p->stage[slot], bytes));
    p->stage_bytes[slot] = bytes;
    return 0;
}

int stage_in(fb_plan* p, int slot, const void* ptr, size_t bytes, const void** dev) {
    if (!ptr) {
        *dev = nullptr;
        return 0;
    }
    if (is_device_ptr(ptr)) {
        *dev = ptr;
        return 0;
    }
    if (ensure_slot(p, slot, bytes)) return -2;
    FB_CUDA(cudaMemcpyAsync(p->stage[slot], ptr, bytes, cudaMemcpyHostToDevice, p->stream));
    // The caller owns its buffer again when the call returns.  A copy from pageable memory has left the buffer by
    // now; one from pinned memory is truly asynchronous and an entry point whose outputs all stay on the device
    // would return with it in flight, so wait for it here (the kernels that follow depend on it anyway).
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, ptr) != cudaSuccess) {
        cudaGetLastError();
    } else if (at.type == cudaMemoryTypeHost) {
        FB_CUDA(cudaStreamSynchronize(p->stream));
    }
    *dev = p->stage[slot];
    return 0;
}

int stage_out_begin(fb_plan* p, int slot, void* ptr, size_t bytes, void** dev) {
    if (!ptr) {
        *dev = nullptr;
        return 0;
    }
    if (is_device_ptr(ptr)) {
        *dev = ptr;
        return 0;
    }
    if (ensure_slot(p, slot, bytes)) return -2;
    *dev = p->stage[slot];
    return 0;
}

int stage_out_end(fb_plan* p, int slot, void* ptr, size_t bytes) {
    if (!ptr || is_device_ptr(ptr)) return 0;
    FB_CUDA(cudaMemcpyAsync(ptr, p->stage[slot], bytes, cudaMemcpyDeviceToHost, p->stream));
    FB_CUDA(cudaStreamSynchronize(p->stream));
    return 0;
}

int ensure_work(fb_plan* p) {
    const size_t need = (size_t)p->na * p->N * p->N * sizeof(float2);
    if (p->work_bytes >= need) return 0;
    if (p->work) FB_CUDA(cudaFree(p->work));
    p->work = nullptr;
    p->work_bytes = 0;
    FB_CUDA(cudaMalloc((void**)&p->work, need));
    p->work_bytes = need;
    return 0;
}

int ensure_aux(fb_plan* p, size_t bytes) {
    if (p->aux_bytes >= bytes) return 0;
    if (p->aux) FB_CUDA(cudaFree(p->aux));
    p->aux = nullptr;
    p->aux_bytes = 0;
    FB_CUDA(cudaMalloc(&p->aux, bytes));
    p->aux_bytes = bytes;
    return 0;
}

int pk_clear(fb_plan* p) {
    FB_CUDA(cudaMemsetAsync(p->h_count, 0, FB_PK_COPIES * (FB_MAX_EDGES + 1) * sizeof(unsigned long long), p->stream));
    FB_CUDA(cudaMemsetAsync(p->h_sums, 0, 4 * FB_PK_COPIES * (FB_MAX_EDGES + 1) * sizeof(double), p->stream));
    return 0;
}

// the device histogram is replicated FB_PK_COPIES times (CTAs spread their reductions over the
// copies so that no single L2 address serialises them); fold the copies on the device, then one
// small pinned D2H copy
__global__ void __launch_bounds__(32) k_pk_fold(const unsigned long long* __restrict__ cnt,
                                                const double* __restrict__ sums,
                                                unsigned long long* __restrict__ cnt_out,
                                                double* __restrict__ sums_out) {
    // one warp per bin: lane l folds replicas l, l+32, ..., then a shuffle tree (fixed order: deterministic)
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
    if (lane == 0) {
        cnt_out[i] = c;
#pragma unroll
        for (int j = 0; j < 4; ++j) sums_out[(size_t)j * (FB_MAX_EDGES + 1) + i] = s[j];
    }
}

int pk_fetch(fb_plan* p, fb_pk_result* out) {
    const int n = p->nedges + 1;
    unsigned long long* dc = reinterpret_cast<unsigned long long*>(p->pk_fold);
    double* ds = reinterpret_cast<double*>(dc + (FB_MAX_EDGES + 1));
    k_pk_fold<<<n, 32, 0, p->stream>>>(p->h_count, p->h_sums, dc, ds);
    FB_LAUNCH_CHECK();
    return pk_fetch_folded(p, out);
}

int pk_fetch_folded(fb_plan* p, fb_pk_result* out) {
    const int n = p->nedges + 1;
    const size_t bytes = 5 * (size_t)(FB_MAX_EDGES + 1) * 8;
    FB_CUDA(cudaMemcpyAsync(p->pk_host, p->pk_fold, bytes, cudaMemcpyDeviceToHost, p->stream));
    FB_CUDA(cudaStreamSynchronize(p->stream));
    const unsigned long long* hc = reinterpret_cast<const unsigned long long*>(p->pk_host);
    const double* hs = reinterpret_cast<const double*>(hc + (FB_MAX_EDGES + 1));
    double* dst[4] = {out->sum1, out->sum2, out->sum_l2, out->sum_l4};
    for (int i = 0; i < n; ++i) {
        if (out->count) out->count[i] = hc[i];
        for (int j = 0; j < 4; ++j)
            if (dst[j]) dst[j][i] = hs[(size_t)j * (FB_MAX_EDGES + 1) + i];
    }
    return 0;
}

int scal_clear(fb_plan* p) {
    FB_CUDA(cudaMemsetAsync(p->scal, 0, 8 * sizeof(double), p->stream));
    return 0;
}
int scal_fetch(fb_plan* p, double* out, int n) {
    if (!out) return 0;
    FB_CUDA(cudaMemcpyAsync(p->scal_host, p->scal, n * sizeof(double), cudaMemcpyDeviceToHost, p->stream));
    FB_CUDA(cudaStreamSynchronize(p->stream));
    for (int i = 0; i < n; ++i) out[i] = p->scal_host[i];
    return 0;
}

}  // namespace fb

using namespace fb;

fb::KSpace fb_plan::kspace() const {
    fb::KSpace K;
    K.N = N;
    K.a0 = a0;
    K.inv_lx2 = (float)(1.0 / (Lx * Lx));
    K.inv_ly2 = (float)(1.0 / (Ly * Ly));
    K.inv_lz2 = (float)(1.0 / (Lz * Lz));
    K.two_pi_over_lx = (float)(2.0 * M_PI / Lx);
    K.two_pi_over_ly = (float)(2.0 * M_PI / Ly);
    K.two_pi_over_lz = (float)(2.0 * M_PI / Lz);
    K.sqrtp = sqrtp;
    K.sqrtp_mode = sqrtp_mode;
    K.sqrtp_n = (int)sqrtp_n;
    K.log2s0 = (float)log2s0;
    K.inv_dlog2s = dlog2s > 0 ? (float)(1.0 / dlog2s) : 0.f;
    K.bt_shift = 23 - (int)dlog2s;                       // mode 3: dlog2s carries M, log2s0 carries the base index
    K.bt_base = (int)log2s0;
    K.bt_scale = 1.0f / (float)(1u << (K.bt_shift > 0 && K.bt_shift < 24 ? K.bt_shift : 1));
    K.bt_mask = (1u << (K.bt_shift > 0 && K.bt_shift < 24 ? K.bt_shift : 1)) - 1u;
    K.sqrtp_pairs = sqrtp_pairs;
    K.tperp = tperp;
    K.tpar = tpar;
    K.tdense = tdense;
    K.ax = ax;
    K.ay = ay;
    K.az = az;
    K.thr = thr;
    K.nedges = nedges;
    K.bin_l0 = (float)bin_l0;
    K.bin_inv_d = (float)bin_inv_d;
    K.inv_boxfactor = (Lx * Ly * Lz) / pow((double)N, 6.0);        // 1 / box.py:94
    return K;
}

fb::PkDev fb_plan::pkdev() const {
    fb::PkDev d;
    d.count = h_count;
    d.sum1 = h_sums;
    d.sum2 = h_sums + (size_t)FB_PK_COPIES * (FB_MAX_EDGES + 1);
    d.l2 = h_sums + 2 * (size_t)FB_PK_COPIES * (FB_MAX_EDGES + 1);
    d.l4 = h_sums + 3 * (size_t)FB_PK_COPIES * (FB_MAX_EDGES + 1);
    return d;
}

int fb::check_flags(fb_plan* p, int flags) {
    if (flags & FB_F_SQRTPK) FB_CHECK(p->sqrtp != nullptr, "FB_F_SQRTPK set but no sqrt(P) table (fb_set_sqrt_pk)");
    if (flags & FB_F_FILTER)
        FB_CHECK(p->tdense != nullptr || (p->tperp != nullptr && p->tpar != nullptr),
                 "FB_F_FILTER set but no filter table (fb_set_filter)");
    if (flags & FB_F_PK) FB_CHECK(p->nedges > 0, "FB_F_PK set but no bins (fb_set_pk_bins)");
    return 0;
}

extern "C" {

const char* fb_last_error(void) { return fb::g_err; }
const char* fb_version(void) { return "fastbox_b200 0.1 (sm_100a)"; }
uint64_t fb_launch_count(void) { return fb::g_launches.load(); }

static int plan_init(fb_plan* p, int N, double Lx, double Ly, double Lz, int device);

int fb_plan_create(fb_plan** out, int N, double Lx, double Ly, double Lz, int device) {
    FB_CHECK(out != nullptr, "fb_plan_create: null output pointer");
    FB_CHECK(N >= 8 && N <= 2048 && (N & (N - 1)) == 0, "fb_plan_create: N=%d must be a power of two in [8,2048]", N);
    FB_CHECK(Lx > 0 && Ly > 0 && Lz > 0, "fb_plan_create: box lengths must be positive");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        set_error("fb_plan_create: no CUDA device available (%s); this library has no CPU fallback",
                  e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
        return -4;
    }
    FB_CHECK(device >= 0 && device < ndev, "fb_plan_create: device %d out of range (have %d)", device, ndev);
    FB_CUDA(cudaSetDevice(device));
    fb_plan* p = (fb_plan*)calloc(1, sizeof(fb_plan));
    FB_CHECK(p != nullptr, "out of host memory");
    const int rc = plan_init(p, N, Lx, Ly, Lz, device);
    if (rc) {                                            // nothing of a half-built plan is left behind
        fb_plan_destroy(p);
        return rc;
    }
    *out = p;
    return 0;
}

static int plan_init(fb_plan* p, int N, double Lx, double Ly, double Lz, int device) {
    p->N = N;
    p->Lx = Lx;
    p->Ly = Ly;
    p->Lz = Lz;
    p->device = device;
    p->a0 = 0;
    p->na = N / 2 + 1;
    p->y0 = 0;
    p->ny = N;
    cudaDeviceProp prop;
    FB_CUDA(cudaGetDeviceProperties(&prop, device));
    p->sm_count = prop.multiProcessorCount;
    FB_CUDA(cudaStreamCreateWithFlags(&p->stream, cudaStreamNonBlocking));
    // twiddles exp(-2 pi i m / n) for every power-of-two n <= FB_NMAX_TW at [n, 2n), evaluated in double
    {
        std::vector<float2> tw(FB_TW_ENTRIES, make_float2(1.f, 0.f));
        for (int n = 1; n <= FB_NMAX_TW; n *= 2)
        for (int m = 0; m < n; ++m) {
            const double ang = -2.0 * M_PI * (double)m / (double)n;
            tw[n + m] = make_float2((float)cos(ang), (float)sin(ang));
        }
        FB_CUDA(cudaMalloc((void**)&p->tw, FB_TW_ENTRIES * sizeof(float2)));
        FB_CUDA(cudaMemcpy(p->tw, tw.data(), FB_TW_ENTRIES * sizeof(float2), cudaMemcpyHostToDevice));
    }
    // per-axis (m/L)^2 in float64: same two roundings as NumPy's (K/L)**2. (box.py:125-127)
    {
        std::vector<double> t(3 * (size_t)N);
        const double L[3] = {Lx, Ly, Lz};
        for (int ax = 0; ax < 3; ++ax)
            for (int i = 0; i < N; ++i) {
                const double m = (double)(i < N / 2 ? i : i - N);
                volatile double r = m / L[ax];
                volatile double sq = r * r;
                t[(size_t)ax * N + i] = sq;
            }
        FB_CUDA(cudaMalloc((void**)&p->ax, 3 * (size_t)N * sizeof(double)));
        FB_CUDA(cudaMemcpy(p->ax, t.data(), 3 * (size_t)N * sizeof(double), cudaMemcpyHostToDevice));
        p->ay = p->ax + N;
        p->az = p->ax + 2 * (size_t)N;
    }
    FB_CUDA(cudaMalloc((void**)&p->thr, FB_MAX_EDGES * sizeof(double)));
    FB_CUDA(cudaMalloc((void**)&p->h_count, FB_PK_COPIES * (FB_MAX_EDGES + 1) * sizeof(unsigned long long)));
    FB_CUDA(cudaMalloc((void**)&p->h_sums, 4 * FB_PK_COPIES * (FB_MAX_EDGES + 1) * sizeof(double)));
    FB_CUDA(cudaMalloc(&p->pk_fold, 5 * (FB_MAX_EDGES + 1) * 8));
    FB_CUDA(cudaMallocHost(&p->pk_host, 5 * (FB_MAX_EDGES + 1) * 8));
    FB_CUDA(cudaMalloc((void**)&p->scal, 8 * sizeof(double)));
    FB_CUDA(cudaMallocHost((void**)&p->scal_host, 8 * sizeof(double)));
    for (int i = 0; i < 8; ++i) FB_CUDA(cudaEventCreate(&p->ev[i]));
    return 0;
}

int fb_plan_destroy(fb_plan* p) {
    if (!p) return 0;
    cudaSetDevice(p->device);
    if (p->stream) cudaStreamSynchronize(p->stream);
    fb::dist_destroy(p);
    cudaFree(p->tw);
    cudaFree(p->ax);
    cudaFree(p->thr);
    cudaFree(p->h_count);
    cudaFree(p->h_sums);
    cudaFree(p->scal);
    cudaFree(p->pk_fold);
    cudaFreeHost(p->pk_host);
    cudaFreeHost(p->scal_host);
    cudaFree(p->sqrtp);
    cudaFree(p->sqrtp_pairs);
    cudaFree(p->tperp);
    cudaFree(p->tpar);
    cudaFree(p->tdense);
    cudaFree(p->work);
    cudaFree(p->aux);
    cudaFree(p->beam_spec);
    for (int i = 0; i < 6; ++i) cudaFree(p->stage[i]);
    for (int i = 0; i < 8; ++i)
        if (p->ev[i]) cudaEventDestroy(p->ev[i]);
    if (p->stream) cudaStreamDestroy(p->stream);
    cudaGetLastError();
    free(p);
    return 0;
}

int fb_sync(fb_plan* p) {
    FB_CUDA(cudaStreamSynchronize(p->stream));
    return 0;
}

int fb_plan_set_slab(fb_plan* p, int a0, int na, int y0, int ny) {
    FB_CHECK(a0 >= 0 && na >= 1 && a0 + na <= p->N / 2 + 1, "fb_plan_set_slab: bad kx range [%d,%d)", a0, a0 + na);
    FB_CHECK(y0 >= 0 && ny >= 1 && y0 + ny <= p->N, "fb_plan_set_slab: bad y range [%d,%d)", y0, y0 + ny);
    p->a0 = a0;
    p->na = na;
    p->y0 = y0;
    p->ny = ny;
    return 0;
}

int fb_dev_alloc(void** ptr, size_t bytes) {
    FB_CUDA(cudaMalloc(ptr, bytes));
    return 0;
}
int fb_dev_alloc_on(int device, void** ptr, size_t bytes) {
    FB_CUDA(cudaSetDevice(device));
    FB_CUDA(cudaMalloc(ptr, bytes));
    return 0;
}
int fb_dev_free(void* ptr) {
    FB_CUDA(cudaFree(ptr));
    return 0;
}
int fb_host_alloc(void** ptr, size_t bytes) {
    FB_CUDA(cudaMallocHost(ptr, bytes));
    return 0;
}
int fb_host_free(void* ptr) {
    FB_CUDA(cudaFreeHost(ptr));
    return 0;
}
int fb_copy(fb_plan* p, void* dst, const void* src, size_t bytes) {
    FB_CUDA(cudaSetDevice(p->device));
    FB_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, p->stream));
    FB_CUDA(cudaStreamSynchronize(p->stream));
    return 0;
}
int fb_device_info(int device, char* name, int name_len, int* sm_count, size_t* total_mem) {
    cudaDeviceProp prop;
    FB_CUDA(cudaGetDeviceProperties(&prop, device));
    if (name && name_len > 0) {
        strncpy(name, prop.name, name_len - 1);
        name[name_len - 1] = 0;
    }
    if (sm_count) *sm_count = prop.multiProcessorCount;
    if (total_mem) *total_mem = prop.totalGlobalMem;
    return 0;
}

__global__ void k_sqrtp_pairs(const float* __restrict__ t, float2* __restrict__ out, long n, float scale) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float t0 = t[i], t1 = t[i + 1 < n ? i + 1 : i];
    out[i] = make_float2(t0, (t1 - t0) * scale);         // scale is a power of two: the product stays exact
}

static int upload_table(float** slot, size_t* cap, const float* host, size_t n) {
    if (!host) {
        if (*slot) FB_CUDA(cudaFree(*slot));
        *slot = nullptr;
        *cap = 0;
        return 0;
    }
    if (*cap < n) {                                  // reuse the allocation when the table fits
        if (*slot) FB_CUDA(cudaFree(*slot));
        *slot = nullptr;
        *cap = 0;
        FB_CUDA(cudaMalloc((void**)slot, n * sizeof(float)));
        *cap = n;
    }
    FB_CUDA(cudaMemcpy(*slot, host, n * sizeof(float), cudaMemcpyDefault));
    return 0;
}

int fb_set_sqrt_pk(fb_plan* p, const float* table, long n, int mode, double log2s0, double dlog2s) {
    FB_CUDA(cudaStreamSynchronize(p->stream));
    FB_CHECK(mode >= 1 && mode <= 3, "fb_set_sqrt_pk: mode must be 1 (integer LUT), 2 (log2 table) or 3 (float-bit table)");
    if (mode == 3) FB_CHECK(dlog2s >= 4 && dlog2s <= 16 && n >= 2, "fb_set_sqrt_pk: mode 3 needs 4 <= M <= 16 mantissa bits");
    if (mode == 1) {
        const long need = 3L * (p->N / 2) * (p->N / 2) + 1;
        FB_CHECK(n >= need, "fb_set_sqrt_pk: integer LUT needs %ld entries, got %ld", need, n);
    } else if (mode == 2) {
        FB_CHECK(n >= 2 && dlog2s > 0, "fb_set_sqrt_pk: log table needs n>=2 and dlog2s>0");
    }
    if (upload_table(&p->sqrtp, &p->tab_cap[0], table, (size_t)n)) return -2;
    p->sqrtp_mode = mode;
    p->sqrtp_n = n;
    p->log2s0 = log2s0;
    p->dlog2s = dlog2s;
    if (mode == 3) {
        if (p->sqrtp_pairs_cap < (size_t)n) {
            if (p->sqrtp_pairs) FB_CUDA(cudaFree(p->sqrtp_pairs));
            p->sqrtp_pairs = nullptr;
            p->sqrtp_pairs_cap = 0;
            FB_CUDA(cudaMalloc((void**)&p->sqrtp_pairs, (size_t)n * sizeof(float2)));
            p->sqrtp_pairs_cap = (size_t)n;
        }
        const int shift = 23 - (int)dlog2s;
        k_sqrtp_pairs<<<(unsigned)((n + 255) / 256), 256, 0, p->stream>>>(p->sqrtp, p->sqrtp_pairs, n,
                                                                          1.0f / (float)(1u << shift));
        FB_LAUNCH_CHECK();
        FB_CUDA(cudaStreamSynchronize(p->stream));
    }
    return 0;
}

int fb_set_filter(fb_plan* p, const float* tperp, const float* tpar, const float* tdense) {
    FB_CUDA(cudaStreamSynchronize(p->stream));
    const size_t N = p->N, H = p->N / 2 + 1;
    FB_CHECK((tperp == nullptr) == (tpar == nullptr), "fb_set_filter: tperp and tpar go together");
    if (upload_table(&p->tperp, &p->tab_cap[1], tperp, H * N)) return -2;
    if (upload_table(&p->tpar, &p->tab_cap[2], tpar, N)) return -2;
    if (upload_table(&p->tdense, &p->tab_cap[3], tdense, H * N * N)) return -2;
    return 0;
}

int fb_set_pk_bins(fb_plan* p, const double* thresholds, int nedges) {
    FB_CHECK(nedges >= 1 && nedges <= FB_MAX_EDGES, "fb_set_pk_bins: nedges=%d out of range [1,%d]", nedges,
             FB_MAX_EDGES);
    for (int i = 1; i < nedges; ++i)
        FB_CHECK(thresholds[i] >= thresholds[i - 1], "fb_set_pk_bins: thresholds must be non-decreasing");
    FB_CUDA(cudaStreamSynchronize(p->stream));
    FB_CUDA(cudaMemcpy(p->thr, thresholds, nedges * sizeof(double), cudaMemcpyHostToDevice));
    p->nedges = nedges;
    // log-spaced edges (box.py:749)?  then a bin can be guessed from one log2 on the device
    p->bin_l0 = 0.0;
    p->bin_inv_d = 0.0;
    if (nedges >= 3 && thresholds[0] > 0.0) {
        const double l0 = log2(thresholds[0]), d = (log2(thresholds[nedges - 1]) - l0) / (nedges - 1);
        bool ok = d > 0.0;
        for (int i = 0; ok && i < nedges; ++i) ok = fabs(log2(thresholds[i]) - (l0 + d * i)) <= 0.25 * d;
        if (ok) {
            p->bin_l0 = l0;
            p->bin_inv_d = 1.0 / d;
        }
    }
    return 0;
}

static void mark(fb_plan* p, int i) { cudaEventRecord(p->ev[i], p->stream); }

// inverse pipeline common tail: y columns, x c2r
static int inverse_tail(fb_plan* p, int flags, float scale, float* field_dev, double* sums_dev) {
    if (launch_cols(p, p->work, p->na, +1)) return -3;
    mark(p, 2);
    XArgs xa;
    memset(&xa, 0, sizeof(xa));
    xa.spec = p->work;
    xa.field = field_dev;
    xa.tw = p->tw;
    xa.ncols = (size_t)p->N * p->N;
    xa.flags = flags;
    const double n3 = (double)p->N * p->N * p->N;
    xa.scale = (float)((double)scale / n3);              // numpy ifftn normalisation
    xa.sums = sums_dev;
    if (launch_x_c2r(p, xa)) return -3;
    mark(p, 3);
    p->n_last = 3;
    return 0;
}

int fb_realise(fb_plan* p, const float* re, const float* im, uint64_t seed, int flags, float scale, float* field_out,
               void* spec_out, fb_pk_result* pk, double* sum_out) {
    FB_CUDA(cudaSetDevice(p->device));
    FB_CHECK(p->a0 == 0 && p->na == p->N / 2 + 1, "fb_realise needs the full grid (use the slab building blocks)");
    FB_CHECK((re == nullptr) == (im == nullptr), "fb_realise: re and im must both be given or both NULL");
    FB_CHECK(field_out != nullptr, "fb_realise: field_out is NULL");
    if (pk) flags |= FB_F_PK; else flags &= ~(FB_F_PK | FB_F_POLES);
    if (check_flags(p, flags)) return -1;
    if (ensure_work(p)) return -2;
    const size_t n3 = (size_t)p->N * p->N * p->N, nh = (size_t)p->na * p->N * p->N;
    RowsArgs ra;
    memset(&ra, 0, sizeof(ra));
    const void *dre = nullptr, *dim = nullptr;
    if (re) {
        if (stage_in(p, 0, re, n3 * sizeof(float), &dre)) return -2;
        if (stage_in(p, 1, im, n3 * sizeof(float), &dim)) return -2;
    }
    void *dfield = nullptr, *dspec = nullptr;
    if (stage_out_begin(p, 2, field_out, n3 * sizeof(float), &dfield)) return -2;
    if (stage_out_begin(p, 3, spec_out, nh * sizeof(float2), &dspec)) return -2;
    ra.re = (const float*)dre;
    ra.im = (const float*)dim;
    ra.seed = seed;
    ra.work = p->work;
    ra.spec_out = (float2*)dspec;
    ra.tw = p->tw;
    ra.nrows = (long)p->na * p->N;
    ra.flags = flags;
    ra.kind = FB_KIND_PLAIN;
    ra.K = p->kspace();
    ra.pk = p->pkdev();
    if (pk && pk_clear(p)) return -2;
    if (scal_clear(p)) return -2;
    mark(p, 0);
    if (re ? launch_rows_inv_noise(p, ra) : launch_rows_inv_philox(p, ra)) return -3;
    mark(p, 1);
    if (inverse_tail(p, flags, scale, (float*)dfield, sum_out ? p->scal : nullptr)) return -3;   // sums only on request
    if (stage_out_end(p, 2, field_out, n3 * sizeof(float))) return -2;
    if (stage_out_end(p, 3, spec_out, nh * sizeof(float2))) return -2;
    if (pk && pk_fetch(p, pk)) return -2;
    if (scal_fetch(p, sum_out, 2)) return -2;
    return 0;
}

int fb_spectrum_to_field(fb_plan* p, const void* spec_half, int flags, int kind, float scale, float* field_out,
                         double* sum_out) {
    FB_CUDA(cudaSetDevice(p->device));
    FB_CHECK(p->a0 == 0 && p->na == p->N / 2 + 1, "fb_spectrum_to_field needs the full grid");
    FB_CHECK(spec_half && field_out, "fb_spectrum_to_field: NULL buffer");
    FB_CHECK(kind >= FB_KIND_PLAIN && kind <= FB_KIND_POTENTIAL, "bad kind %d", kind);
    flags &= ~(FB_F_PK | FB_F_POLES | FB_F_ANTIHERM);
    if (check_flags(p, flags)) return -1;
    if (ensure_work(p)) return -2;
    const size_t n3 = (size_t)p->N * p->N * p->N, nh = (size_t)p->na * p->N * p->N;
    const void* dspec = nullptr;
    void* dfield = nullptr;
    if (stage_in(p, 3, spec_half, nh * sizeof(float2), &dspec)) return -2;
    if (stage_out_begin(p, 2, field_out, n3 * sizeof(float), &dfield)) return -2;
    RowsArgs ra;
    memset(&ra, 0, sizeof(ra));
    ra.src = (const float2*)dspec;
    ra.work = p->work;
    ra.tw = p->tw;
    ra.nrows = (long)p->na * p->N;
    ra.flags = flags;
    ra.kind = kind;
    ra.K = p->kspace();
    ra.pk = p->pkdev();
    if (scal_clear(p)) return -2;
    mark(p, 0);
    if (launch_rows_inv_spec(p, ra)) return -3;
    mark(p, 1);
    if (inverse_tail(p, flags, scale, (float*)dfield, sum_out ? p->scal : nullptr)) return -3;
    if (stage_out_end(p, 2, field_out, n3 * sizeof(float))) return -2;
    if (scal_fetch(p, sum_out, 2)) return -2;
    return 0;
}

int fb_cube_to_field(fb_plan* p, const void* cube, int flags, int part, float scale, float* field_out) {
    FB_CUDA(cudaSetDevice(p->device));
    FB_CHECK(p->a0 == 0 && p->na == p->N / 2 + 1, "fb_cube_to_field needs the full grid");
    FB_CHECK(cube && field_out, "fb_cube_to_field: NULL buffer");
    flags &= ~(FB_F_PK | FB_F_POLES | FB_F_ANTIHERM | FB_F_EXP);
    if (part) flags |= FB_F_ANTIHERM;
    if (check_flags(p, flags)) return -1;
    if (ensure_work(p)) return -2;
    const size_t n3 = (size_t)p->N * p->N * p->N;
    const void* dcube = nullptr;
    void* dfield = nullptr;
    if (stage_in(p, 0, cube, n3 * sizeof(float2), &dcube)) return -2;
    if (stage_out_begin(p, 2, field_out, n3 * sizeof(float), &dfield)) return -2;
    RowsArgs ra;
    memset(&ra, 0, sizeof(ra));
    ra.src = (const float2*)dcube;
    ra.work = p->work;
    ra.tw = p->tw;
    ra.nrows = (long)p->na * p->N;
    ra.flags = flags;
    ra.kind = FB_KIND_PLAIN;
    ra.K = p->kspace();
    ra.pk = p->pkdev();
    mark(p, 0);
    if (launch_rows_inv_cube(p, ra)) return -3;
    mark(p, 1);
    if (inverse_tail(p, flags, scale, (float*)dfield, nullptr)) return -3;
    if (stage_out_end(p, 2, field_out, n3 * sizeof(float))) return -2;
    return 0;
}

int fb_field_to_spectrum(fb_plan* p, const float* field, void* spec_out, const void* cross_spec, int flags,
                         fb_pk_result* pk) {
    FB_CUDA(cudaSetDevice(p->device));
    FB_CHECK(p->a0 == 0 && p->na == p->N / 2 + 1, "fb_field_to_spectrum needs the full grid");
    FB_CHECK(field != nullptr, "fb_field_to_spectrum: field is NULL");
    if (pk) flags |= FB_F_PK; else flags &= ~(FB_F_PK | FB_F_POLES);
    flags &= (FB_F_PK | FB_F_POLES);
    if (check_flags(p, flags)) return -1;
    if (ensure_work(p)) return -2;
    const size_t n3 = (size_t)p->N * p->N * p->N, nh = (size_t)p->na * p->N * p->N;
    const void *dfield = nullptr, *dcross = nullptr;
    void* dspec = nullptr;
    if (stage_in(p, 2, field, n3 * sizeof(float), &dfield)) return -2;
    if (stage_in(p, 4, cross_spec, nh * sizeof(float2), &dcross)) return -2;
    if (stage_out_begin(p, 3, spec_out, nh * sizeof(float2), &dspec)) return -2;
    if (pk && pk_clear(p)) return -2;
    XArgs xa;
    memset(&xa, 0, sizeof(xa));
    xa.field_in = (const float*)dfield;
    xa.spec_out = p->work;
    xa.tw = p->tw;
    xa.ncols = (size_t)p->N * p->N;
    mark(p, 0);
    if (launch_x_r2c(p, xa)) return -3;
    mark(p, 1);
    if (launch_cols(p, p->work, p->na, -1)) return -3;
    mark(p, 2);
    RowsArgs ra;
    memset(&ra, 0, sizeof(ra));
    ra.work = p->work;
    ra.cross = (const float2*)dcross;
    ra.spec_out = (float2*)dspec;
    ra.tw = p->tw;
    ra.nrows = (long)p->na * p->N;
    ra.flags = flags;
    ra.K = p->kspace();
    ra.pk = p->pkdev();
    if (launch_rows_fwd(p, ra)) return -3;
    mark(p, 3);
    p->n_last = 3;
    if (stage_out_end(p, 3, spec_out, nh * sizeof(float2))) return -2;
    if (pk && pk_fetch(p, pk)) return -2;
    return 0;
}

int fb_pk_from_spectrum(fb_plan* p, const void* spec, const void* cross_spec, int full_cube, int flags,
                        fb_pk_result* pk) {
    FB_CUDA(cudaSetDevice(p->device));
    FB_CHECK(spec && pk, "fb_pk_from_spectrum: NULL argument");
    flags = (flags & FB_F_POLES) | FB_F_PK;
    if (check_flags(p, flags)) return -1;
    const int nplanes = full_cube ? p->N : p->na;
    const size_t n = (size_t)nplanes * p->N * p->N;
    const void *dspec = nullptr, *dcross = nullptr;
    if (stage_in(p, 3, spec, n * sizeof(float2), &dspec)) return -2;
    if (stage_in(p, 4, cross_spec, n * sizeof(float2), &dcross)) return -2;
    if (pk_clear(p)) return -2;
    const int a0_saved = p->a0;
    if (full_cube) p->a0 = 0;
    int rc = launch_pk_spectrum(p, (const float2*)dspec, (const float2*)dcross, nplanes, full_cube, flags);
    p->a0 = a0_saved;
    if (rc) return -3;
    return pk_fetch(p, pk);
}

// ---- building blocks -----------------------------------------------------------
int fb_fft_pass_c2c(fb_plan* p, void* data, int nplanes, int pass, int sign) {
    FB_CUDA(cudaSetDevice(p->device));
    FB_CHECK(is_device_ptr(data), "fb_fft_pass_c2c: data must be device memory");
    FB_CHECK(sign == 1 || sign == -1, "sign must be +1/-1");
    if (pass == 1) return launch_cols(p, (float2*)data, nplanes, sign);
    FB_CHECK(pass == 0, "pass must be 0 (rows) or 1 (columns)");
    // rows: inverse via the SRC_SPEC kernel with no multiplier, forward via rows_fwd (in place)
    RowsArgs ra;
    memset(&ra, 0, sizeof(ra));
    ra.tw = p->tw;
    ra.nrows = (long)nplanes * p->N;
    ra.K = p->kspace();
    ra.pk = p->pkdev();
    if (sign > 0) {
        ra.src = (const float2*)data;
        ra.work = (float2*)data;
        return launch_rows_inv_spec(p, ra);
    }
    ra.work = (float2*)data;
    ra.spec_out = (float2*)data;
    return launch_rows_fwd(p, ra);
}

static int x_c2r_impl(fb_plan* p, const void* spec, const long* plane_off, float* field, long ncols, int flags,
                      float scale, double* sum_out) {
    FB_CUDA(cudaSetDevice(p->device));
    FB_CHECK(is_device_ptr(spec) && is_device_ptr(field), "fb_fft_pass_x_c2r: buffers must be device memory");
    XArgs xa;
    memset(&xa, 0, sizeof(xa));
    xa.plane_off = plane_off;
    xa.spec = (const float2*)spec;
    xa.field = field;
    xa.tw = p->tw;
    xa.ncols = (size_t)ncols;
    xa.flags = flags & FB_F_EXP;
    xa.scale = scale;
    xa.sums = p->scal;
    if (scal_clear(p)) return -2;
    if (launch_x_c2r(p, xa)) return -3;
    return scal_fetch(p, sum_out, 2);
}

int fb_fft_pass_x_c2r(fb_plan* p, const void* spec, float* field, long ncols, int flags, float scale, double* sum_out) {
    return x_c2r_impl(p, spec, nullptr, field, ncols, flags, scale, sum_out);
}

int fb_fft_pass_x_c2r_gather(fb_plan* p, const void* spec, const long* plane_off, float* field, long ncols, int flags,
                             float scale, double* sum_out) {
    FB_CHECK(plane_off != nullptr && is_device_ptr(plane_off), "fb_fft_pass_x_c2r_gather: plane_off must be device memory");
    return x_c2r_impl(p, spec, plane_off, field, ncols, flags, scale, sum_out);
}

int fb_fft_pass_x_r2c(fb_plan* p, const float* field, void* spec, long ncols) {
    FB_CUDA(cudaSetDevice(p->device));
    FB_CHECK(is_device_ptr(spec) && is_device_ptr(field), "fb_fft_pass_x_r2c: buffers must be device memory");
    XArgs xa;
    memset(&xa, 0, sizeof(xa));
    xa.field_in = field;
    xa.spec_out = (float2*)spec;
    xa.tw = p->tw;
    xa.ncols = (size_t)ncols;
    return launch_x_r2c(p, xa);
}

int fb_realise_local_kspace(fb_plan* p, uint64_t seed, int flags, void* work, void* send, int ny, fb_pk_result* pk) {
    FB_CUDA(cudaSetDevice(p->device));
    FB_CHECK(is_device_ptr(work), "fb_realise_local_kspace: work must be device memory");
    FB_CHECK(ny == 0 || (send != nullptr && is_device_ptr(send) && send != work && p->N % ny == 0),
             "fb_realise_local_kspace: ny must divide N and send must be a distinct device buffer");
    if (pk) flags |= FB_F_PK; else flags &= ~(FB_F_PK | FB_F_POLES);
    if (check_flags(p, flags)) return -1;
    RowsArgs ra;
    memset(&ra, 0, sizeof(ra));
    ra.seed = seed;
    ra.work = (float2*)work;
    ra.tw = p->tw;
    ra.nrows = (long)p->na * p->N;
    ra.flags = flags;
    ra.kind = FB_KIND_PLAIN;
    ra.K = p->kspace();
    ra.pk = p->pkdev();
    if (pk && pk_clear(p)) return -2;
    mark(p, 0);
    if (launch_rows_inv_philox(p, ra)) return -3;
    mark(p, 1);
    if (ny == 0) {
        if (launch_cols(p, (float2*)work, p->na, +1)) return -3;
    } else {
        if (launch_cols_ex(p, (const float2*)work, (float2*)send, 0, ny, p->na, +1)) return -3;
    }
    mark(p, 2);
    p->n_last = 2;
    if (pk && pk_fetch(p, pk)) return -2;
    return 0;
}

int fb_forward_local_kspace(fb_plan* p, const void* recv, void* work, int ny, void* spec_out, int flags,
                            fb_pk_result* pk) {
    FB_CUDA(cudaSetDevice(p->device));
    FB_CHECK(is_device_ptr(recv) && is_device_ptr(work) && recv != work, "fb_forward_local_kspace: need two device buffers");
    FB_CHECK(ny > 0 && p->N % ny == 0, "fb_forward_local_kspace: ny must divide N");
    FB_CHECK(spec_out == nullptr || is_device_ptr(spec_out), "fb_forward_local_kspace: spec_out must be device memory");
    if (pk) flags |= FB_F_PK; else flags &= ~(FB_F_PK | FB_F_POLES);
    flags &= (FB_F_PK | FB_F_POLES);
    if (check_flags(p, flags)) return -1;
    if (pk && pk_clear(p)) return -2;
    mark(p, 0);
    if (launch_cols_ex(p, (const float2*)recv, (float2*)work, ny, 0, p->na, -1)) return -3;
    mark(p, 1);
    RowsArgs ra;
    memset(&ra, 0, sizeof(ra));
    ra.work = (float2*)work;
    ra.spec_out = (float2*)spec_out;
    ra.tw = p->tw;
    ra.nrows = (long)p->na * p->N;
    ra.flags = flags;
    ra.K = p->kspace();
    ra.pk = p->pkdev();
    if (launch_rows_fwd(p, ra)) return -3;
    mark(p, 2);
    p->n_last = 2;
    if (pk && pk_fetch(p, pk)) return -2;
    return 0;
}

int fb_timer_start(fb_plan* p) {
    FB_CUDA(cudaSetDevice(p->device));
    FB_CUDA(cudaEventRecord(p->ev[6], p->stream));
    return 0;
}
int fb_timer_stop(fb_plan* p, float* ms) {
    FB_CUDA(cudaEventRecord(p->ev[7], p->stream));
    FB_CUDA(cudaEventSynchronize(p->ev[7]));
    FB_CUDA(cudaEventElapsedTime(ms, p->ev[6], p->ev[7]));
    return 0;
}

int fb_last_timings(fb_plan* p, float* ms, int n) {
    FB_CUDA(cudaStreamSynchronize(p->stream));
    for (int i = 0; i < n; ++i) {
        ms[i] = 0.f;
        if (i < p->n_last) FB_CUDA(cudaEventElapsedTime(&ms[i], p->ev[i], p->ev[i + 1]));
    }
    return 0;
}

}  // extern "C"
