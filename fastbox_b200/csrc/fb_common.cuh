// fb_common.cuh -- plan object, error handling, small device helpers.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <atomic>
#include <string>
#include <vector>

#include "../../include/fastbox_b200.h"
#include "fb_fft.cuh"

namespace fb {

#define FB_MAX_EDGES 128
#define FB_PK_COPIES 64      // replicas of the global P(k) histogram (power of two)

void set_error(const char* fmt, ...);
extern std::atomic<uint64_t> g_launches;

#define FB_CUDA(expr)                                                                              \
    do {                                                                                           \
        cudaError_t _e = (expr);                                                                   \
        if (_e != cudaSuccess) {                                                                   \
            fb::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
            return -2;                                                                             \
        }                                                                                          \
    } while (0)

#define FB_CHECK(cond, ...)          \
    do {                             \
        if (!(cond)) {               \
            fb::set_error(__VA_ARGS__); \
            return -1;               \
        }                            \
    } while (0)

#define FB_LAUNCH_CHECK()                                                                       \
    do {                                                                                        \
        fb::g_launches.fetch_add(1);                                                            \
        cudaError_t _e = cudaGetLastError();                                                    \
        if (_e != cudaSuccess) {                                                                \
            fb::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), __FILE__, __LINE__); \
            return -3;                                                                          \
        }                                                                                       \
    } while (0)

// Everything a first/last pass needs to know about k-space.
struct KSpace {
    int N;
    int a0;                 // global kx index of local plane 0 (slab decomposition)
    float inv_lx2, inv_ly2, inv_lz2;     // 1/L^2
    float two_pi_over_lx, two_pi_over_ly, two_pi_over_lz;
    // sqrt(P) table
    const float* sqrtp;
    const float2* sqrtp_pairs;  // mode 3: (t[i], (t[i+1]-t[i]) * 2^-shift)
    int sqrtp_mode;         // 0 none, 1 integer |m|^2 LUT, 2 log2(s) table
    int sqrtp_n;
    float log2s0, inv_dlog2s;
    int bt_shift, bt_base;  // mode 3: index = (float_bits(s) >> bt_shift) - bt_base
    float bt_scale;         //         fraction = low bits * bt_scale
    unsigned bt_mask;       //         low bits = float_bits(s) & bt_mask
    // filter
    const float* tperp;     // [(N/2+1)*N]
    const float* tpar;      // [N]
    const float* tdense;    // [(N/2+1)*N*N]
    // P(k) binning
    const double* ax;       // (m/Lx)^2 per axis index, float64, bitwise as NumPy
    const double* ay;
    const double* az;
    const double* thr;      // thresholds on s
    int nedges;
    float bin_l0, bin_inv_d; // log2 model of the thresholds (0 = none, use binary search)
    double inv_boxfactor;
};

struct PkDev {              // device histogram, nb+1 entries each
    unsigned long long* count;
    double* sum1;
    double* sum2;
    double* l2;
    double* l4;
};

}  // namespace fb

struct fb_plan {
    int N;
    double Lx, Ly, Lz;
    int device;
    int sm_count;
    cudaStream_t stream;
    // slab
    int a0, na, y0, ny;
    // tables
    float2* tw;
    double *ax, *ay, *az, *thr;
    int nedges;
    double bin_l0, bin_inv_d;
    float* sqrtp;
    float2* sqrtp_pairs;
    size_t sqrtp_pairs_cap;
    int sqrtp_mode;
    long sqrtp_n;
    double log2s0, dlog2s;
    float *tperp, *tpar, *tdense;
    size_t tab_cap[4];
    // workspaces
    float2* work;           // [(na)][N][N] complex
    size_t work_bytes;
    void* stage[6];         // host-staging device buffers
    size_t stage_bytes[6];
    void* pinned;           // pinned bounce buffer for pageable host memory
    size_t pinned_bytes;
    unsigned long long* h_count;
    void* pk_fold;          // device: folded histogram (count + 4 sums)
    void* pk_host;          // pinned copy
    double* h_sums;         // 4 arrays of FB_MAX_EDGES+1
    double* scal;           // device scalars [8]
    double* scal_host;      // pinned
    // beam / misc workspaces
    void* aux;
    size_t aux_bytes;
    void* beam_spec;        // fb_beam_set: normalised 2-D beam spectrum BS[ky][zt][kx][c] + norm[z] + inv[z]
    size_t beam_spec_bytes;
    int beam_ready;
    cudaEvent_t ev[8];
    float last_ms[8];
    int n_last;
    struct fb_dist_state* dist;     // multi-GPU exchange state (fb_dist.cu), NULL until fb_dist_init
    fb::KSpace kspace() const;
    fb::PkDev pkdev() const;
};

namespace fb {

int ensure_work(fb_plan* p);
int ensure_aux(fb_plan* p, size_t bytes);
// returns device pointer for an input buffer (stages host memory into slot)
int stage_in(fb_plan* p, int slot, const void* ptr, size_t bytes, const void** dev);
// returns device pointer to write results to; if `ptr` is host memory a staging slot is used
int stage_out_begin(fb_plan* p, int slot, void* ptr, size_t bytes, void** dev);
int stage_out_end(fb_plan* p, int slot, void* ptr, size_t bytes);
bool is_device_ptr(const void* ptr);
int pk_clear(fb_plan* p);
int pk_fetch(fb_plan* p, fb_pk_result* out);
int pk_fetch_folded(fb_plan* p, fb_pk_result* out);     // D2H of p->pk_fold only (already folded / reduced)
int scal_clear(fb_plan* p);
int scal_fetch(fb_plan* p, double* out, int n);
int check_flags(fb_plan* p, int flags);
void dist_destroy(fb_plan* p);

__device__ __forceinline__ int mode_number(int i, int N) { return i < N / 2 ? i : i - N; }   // box.py:119

// warp sum of a double
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ---- Philox4x32-10 (counter = cell index, key = seed) + Box-Muller ----------
__device__ __forceinline__ void philox_round(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
    const uint32_t n0 = hi1 ^ c[1] ^ k0, n2 = hi0 ^ c[3] ^ k1;
    c[0] = n0;
    c[1] = lo1;
    c[2] = n2;
    c[3] = lo0;
}
__device__ __forceinline__ float2 box_muller(uint32_t x0, uint32_t x1) {
    // u in (0,1): fl((x0 + 0.5) 2^-32) built from two exact pieces; fast intrinsics (MUFU): the
    // absolute error of a normal is ~5e-7, far inside the 1e-5 field tolerance
    const float u = (float)(x0 >> 8) * 5.9604644775390625e-08f + ((float)(x0 & 0xffu) + 0.5f) * 2.3283064365386963e-10f;
    const float r = sqrtf(-2.0f * __logf(u));
    float s, co;
    __sincosf(6.283185307179586f * (((float)x1 + 0.5f) * 2.3283064365386963e-10f) - 3.141592653589793f, &s, &co);
    return make_float2(-r * co, -r * s);             // cos(t) = -cos(t - pi), sin(t) = -sin(t - pi)
}
// Hermitian white noise H0(k) = 1/2 [W(k) + conj W(-k)] of a unit complex white field W (box.py:174-176,
// 187) drawn DIRECTLY, one Box-Muller pair per conjugate pair of modes: for k != -k, H0(k) = (n1 + i n2)/sqrt2
// and H0(-k) = conj H0(k) (Re and Im of variance 1/2, exactly the law of the Hermitianised reference noise);
// a self-conjugate mode gets H0 = n1 (real, variance 1).  The pair is identified by its canonical cell index
// j = min(index(k), index(-k)); Philox4x32-10 block j >> 1 (key = seed) serves two consecutive pairs: words
// (0,1) for even j, (2,3) for odd j.  The stream depends on the global cell index only (any GPU count).
__device__ __forceinline__ void philox_block(uint64_t seed, uint64_t ctr, uint32_t (&c)[4]) {
    c[0] = (uint32_t)ctr;
    c[1] = (uint32_t)(ctr >> 32);
    c[2] = 0u;
    c[3] = 0u;
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        philox_round(c, k0, k1);
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
}
// sqrt2 * H0(k) for one mode (generic form: any row, any mode); callers fold the 1/sqrt2 into their multiplier
__device__ __forceinline__ float2 philox_h0_sqrt2(uint64_t seed, uint64_t idx_k, uint64_t idx_mk) {
    const uint64_t j = idx_k < idx_mk ? idx_k : idx_mk;
    uint32_t c[4];
    philox_block(seed, j >> 1, c);
    const bool odd = (j & 1u) != 0;
    const float2 n = box_muller(odd ? c[2] : c[0], odd ? c[3] : c[1]);
    if (idx_k == idx_mk) return make_float2(1.41421356237309505f * n.x, 0.f);
    return make_float2(n.x, idx_k < idx_mk ? n.y : -n.y);
}
// four consecutive canonical modes j0 .. j0+3 (j0 % 4 == 0, none self-conjugate, none mirrored): two blocks
__device__ __forceinline__ void philox_h0_sqrt2_quad(uint64_t seed, uint64_t j0, float2 (&h)[4]) {
    uint32_t c[4];
    philox_block(seed, j0 >> 1, c);
    h[0] = box_muller(c[0], c[1]);
    h[1] = box_muller(c[2], c[3]);
    philox_block(seed, (j0 >> 1) + 1, c);
    h[2] = box_muller(c[0], c[1]);
    h[3] = box_muller(c[2], c[3]);
}

}  // namespace fb
