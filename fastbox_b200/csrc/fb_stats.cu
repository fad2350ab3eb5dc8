// fb_stats.cu -- two-point statistics next to the 1-D binned P(k):
//   * P(k_perp, |k_par|): moments of |d_k|^2 / boxfactor (or Re a conj b) on a 2-D grid of bins;
//   * xi(r): the correlation function by the Wiener-Khinchin route  xi = ifftn(|fftn(delta)|^2) / N^3  and a
//     radial binning of the lag cube.
// The reference hands both to nbodykit (FFTPower mode='2d', FFTCorr mode='1d': examples/example_endtoend.py:
// 128-151, examples/example_box.py:48-52), which is neither vendored nor pinned: these definitions are restated
// in oracle/restate.py (parity unpinned w.r.t. nbodykit, pinned w.r.t. that restatement).
// Bin index conventions are those of binned_power_spectrum (box.py:758): np.digitize on float64 coordinates.
#include "fb_launch.h"

namespace fb {

#define FB_STAT_MAX_EDGES 64        // per axis for the 2-D spectrum

struct Pk2dArgs {
    const float2* spec;
    const float2* cross;
    long nrows;                     // nplanes * N
    int N, a0, full_cube;
    const double *ax, *ay;          // (m/L)^2 per axis
    const double* thr_perp;         // thresholds on s_perp = ax[a] + ay[b] (np.digitize without a sqrt)
    int nperp;
    const int* ipar;                // [N] bin of |k_par| for every z mode (host digitize)
    int npar;
    float inv_boxfactor;
    unsigned long long* count;      // [(nperp+1)*(npar+1)]
    double* sum1;
    double* sum2;
};

struct StatCell {
    unsigned int cnt;
    double s1, s2;
};

__device__ __forceinline__ void smem_add(StatCell* h, unsigned cnt, double s1, double s2) {
    if (!cnt) return;
    atomicAdd(&h->cnt, cnt);
    atomicAdd(&h->s1, s1);
    atomicAdd(&h->s2, s2);
}

// one warp per row (a, b): the k_perp bin is a property of the row, the k_par bin of the column; every lane
// walks a contiguous piece of the row and flushes its partial sums when the k_par bin changes
__global__ void __launch_bounds__(256) k_pk2d(const Pk2dArgs A) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    StatCell* hist = reinterpret_cast<StatCell*>(smem_raw);
    const int nb2 = (A.nperp + 1) * (A.npar + 1);
    for (int i = threadIdx.x; i < nb2; i += blockDim.x) {
        hist[i].cnt = 0u;
        hist[i].s1 = 0.0;
        hist[i].s2 = 0.0;
    }
    __syncthreads();
    const int N = A.N, lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    const int per = (N + 31) / 32;                       // modes per lane
    for (long row = (long)blockIdx.x * nwarp + warp; row < A.nrows; row += (long)gridDim.x * nwarp) {
        const int a = A.a0 + (int)(row / N), b = (int)(row % N);
        const double sp = __dadd_rn(A.ax[a], A.ay[b]);
        int lo = 0, hi = A.nperp;                        // #{ j : thr[j] <= sp }
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (A.thr_perp[mid] <= sp) lo = mid + 1; else hi = mid;
        }
        StatCell* hrow = hist + (size_t)lo * (A.npar + 1);
        const float w = (A.full_cube || a == 0 || a == N / 2) ? 1.f : 2.f;
        const float2* src = A.spec + (size_t)row * N;
        const float2* crs = A.cross ? A.cross + (size_t)row * N : nullptr;
        const int c0 = lane * per, c1 = min(N, c0 + per);
        int bin = -1;
        unsigned cnt = 0u;
        double s1 = 0.0, s2 = 0.0;
        for (int c = c0; c < c1; ++c) {
            const int jb = __ldg(&A.ipar[c]);
            if (jb != bin) {
                if (bin >= 0) smem_add(&hrow[bin], cnt, s1, s2);
                bin = jb;
                cnt = 0u;
                s1 = s2 = 0.0;
            }
            const float2 h = __ldg(&src[c]);
            float p;
            if (crs) {
                const float2 x = __ldg(&crs[c]);
                p = (h.x * x.x + h.y * x.y) * A.inv_boxfactor;
            } else {
                p = (h.x * h.x + h.y * h.y) * A.inv_boxfactor;
            }
            const double pd = (double)p;
            cnt += (unsigned)w;
            s1 += (double)w * pd;
            s2 = fma((double)w * pd, pd, s2);
        }
        if (bin >= 0) smem_add(&hrow[bin], cnt, s1, s2);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < nb2; i += blockDim.x) {
        if (hist[i].cnt) {
            atomicAdd(&A.count[i], (unsigned long long)hist[i].cnt);
            atomicAdd(&A.sum1[i], hist[i].s1);
            atomicAdd(&A.sum2[i], hist[i].s2);
        }
    }
}

// S <- A conj(B) / N^3 (B = A for the auto-correlation): the power cube whose inverse transform is xi
__global__ void __launch_bounds__(256) k_spec_power(float2* __restrict__ a, const float2* __restrict__ b, size_t n,
                                                     float inv_n3) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const float2 x = a[i];
        const float2 y = b ? b[i] : x;
        a[i] = make_float2((x.x * y.x + x.y * y.y) * inv_n3, (x.y * y.x - x.x * y.y) * inv_n3);
    }
}

struct XiArgs {
    const float* xi;                // lag cube [x][y][z]
    int N;
    double hx, hy, hz;              // cell sizes L/N
    const double* edges;            // r bin edges
    int nedges;
    unsigned long long* count;      // [nedges+1]
    double* sum1;
    double* sum2;
};

// separation of lag (i, j, l) with periodic wrap: r = sqrt(((dx hx)^2 + (dy hy)^2) + (dz hz)^2), float64,
// index = np.digitize(r, edges); lags outside [edges[0], edges[-1]) are dropped (most of the cube)
__global__ void __launch_bounds__(256) k_xi_bin(const XiArgs A) {
    __shared__ StatCell hist[FB_MAX_EDGES + 1];
    for (int i = threadIdx.x; i <= A.nedges; i += blockDim.x) {
        hist[i].cnt = 0u;
        hist[i].s1 = 0.0;
        hist[i].s2 = 0.0;
    }
    __syncthreads();
    const int N = A.N;
    const size_t nlines = (size_t)N * N;
    const double rmin = A.edges[0], rmax = A.edges[A.nedges - 1];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = blockDim.x >> 5;
    for (size_t line = (size_t)blockIdx.x * nwarp + warp; line < nlines; line += (size_t)gridDim.x * nwarp) {
        const int i = (int)(line / N), j = (int)(line % N);
        const double dx = __dmul_rn((double)min(i, N - i), A.hx), dy = __dmul_rn((double)min(j, N - j), A.hy);
        const double rxy2 = __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
        if (rxy2 >= rmax * rmax * 1.0000001) continue;   // the whole line is beyond the last edge (warp uniform)
        const float* src = A.xi + line * N;
        for (int l = lane; l < N; l += 32) {
            const double dz = __dmul_rn((double)min(l, N - l), A.hz);
            const double r = sqrt(__dadd_rn(rxy2, __dmul_rn(dz, dz)));
            if (r < rmin || r >= rmax) continue;
            int lo = 0, hi = A.nedges;                   // #{ e : edges[e] <= r }
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if (A.edges[mid] <= r) lo = mid + 1; else hi = mid;
            }
            const double v = (double)__ldg(&src[l]);
            smem_add(&hist[lo], 1u, v, v * v);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i <= A.nedges; i += blockDim.x) {
        if (hist[i].cnt) {
            atomicAdd(&A.count[i], (unsigned long long)hist[i].cnt);
            atomicAdd(&A.sum1[i], hist[i].s1);
            atomicAdd(&A.sum2[i], hist[i].s2);
        }
    }
}

}  // namespace fb

using namespace fb;

extern "C" {

int fb_pk2d_from_spectrum(fb_plan* p, const void* spec, const void* cross_spec, int full_cube, const double* thr_perp,
                          int nperp, const int32_t* ipar, int npar, uint64_t* count, double* sum1, double* sum2) {
    FB_CUDA(cudaSetDevice(p->device));
    FB_CHECK(spec && thr_perp && ipar && count && sum1 && sum2, "fb_pk2d_from_spectrum: NULL argument");
    FB_CHECK(nperp >= 1 && nperp <= FB_STAT_MAX_EDGES && npar >= 1 && npar <= FB_STAT_MAX_EDGES,
             "fb_pk2d_from_spectrum: 1..%d edges per axis", FB_STAT_MAX_EDGES);
    const int N = p->N;
    const int nplanes = full_cube ? N : p->na;
    const size_t n = (size_t)nplanes * N * N;
    const int nb2 = (nperp + 1) * (npar + 1);
    const void *dspec = nullptr, *dcross = nullptr;
    if (stage_in(p, 3, spec, n * sizeof(float2), &dspec)) return -2;
    if (stage_in(p, 4, cross_spec, n * sizeof(float2), &dcross)) return -2;
    // small tables + results in the aux workspace
    const size_t off_thr = 0, off_ipar = off_thr + FB_STAT_MAX_EDGES * 8, off_cnt = off_ipar + (size_t)N * 4 + 64,
                 off_s1 = (off_cnt + (size_t)nb2 * 8 + 63) / 64 * 64, off_s2 = off_s1 + (size_t)nb2 * 8,
                 total = off_s2 + (size_t)nb2 * 8;
    if (ensure_aux(p, total)) return -2;
    unsigned char* aux = (unsigned char*)p->aux;
    FB_CUDA(cudaMemcpyAsync(aux + off_thr, thr_perp, nperp * sizeof(double), cudaMemcpyHostToDevice, p->stream));
    FB_CUDA(cudaMemcpyAsync(aux + off_ipar, ipar, (size_t)N * sizeof(int32_t), cudaMemcpyHostToDevice, p->stream));
    FB_CUDA(cudaMemsetAsync(aux + off_cnt, 0, total - off_cnt, p->stream));
    Pk2dArgs a;
    memset(&a, 0, sizeof(a));
    a.spec = (const float2*)dspec;
    a.cross = (const float2*)dcross;
    a.nrows = (long)nplanes * N;
    a.N = N;
    a.a0 = full_cube ? 0 : p->a0;
    a.full_cube = full_cube;
    a.ax = p->ax;
    a.ay = p->ay;
    a.thr_perp = (const double*)(aux + off_thr);
    a.nperp = nperp;
    a.ipar = (const int*)(aux + off_ipar);
    a.npar = npar;
    a.inv_boxfactor = (float)((p->Lx * p->Ly * p->Lz) / pow((double)N, 6.0));
    a.count = (unsigned long long*)(aux + off_cnt);
    a.sum1 = (double*)(aux + off_s1);
    a.sum2 = (double*)(aux + off_s2);
    const size_t smem = (size_t)nb2 * sizeof(StatCell);
    if (set_smem(k_pk2d, smem)) return -2;
    k_pk2d<<<p->sm_count * 2, 256, smem, p->stream>>>(a);
    FB_LAUNCH_CHECK();
    FB_CUDA(cudaMemcpyAsync(count, aux + off_cnt, (size_t)nb2 * 8, cudaMemcpyDeviceToHost, p->stream));
    FB_CUDA(cudaMemcpyAsync(sum1, aux + off_s1, (size_t)nb2 * 8, cudaMemcpyDeviceToHost, p->stream));
    FB_CUDA(cudaMemcpyAsync(sum2, aux + off_s2, (size_t)nb2 * 8, cudaMemcpyDeviceToHost, p->stream));
    FB_CUDA(cudaStreamSynchronize(p->stream));
    return 0;
}

// xi(r) of a real field (cross-correlation with field_b if given): forward transform, power cube, inverse
// transform, radial moments.  edges: r bin edges (host, ascending); outputs [nedges+1] (host), index = np.digitize.
// xi_out (nullable, DEVICE float32 [N^3]): the lag cube itself.
int fb_correlation_function(fb_plan* p, const float* field, const float* field_b, const double* edges, int nedges,
                            uint64_t* count, double* sum1, double* sum2, float* xi_out) {
    FB_CUDA(cudaSetDevice(p->device));
    FB_CHECK(p->a0 == 0 && p->na == p->N / 2 + 1, "fb_correlation_function needs the full grid");
    FB_CHECK(field && edges && count && sum1 && sum2, "fb_correlation_function: NULL argument");
    FB_CHECK(nedges >= 2 && nedges <= FB_MAX_EDGES, "fb_correlation_function: 2..%d edges", FB_MAX_EDGES);
    FB_CHECK(xi_out == nullptr || is_device_ptr(xi_out), "fb_correlation_function: xi_out must be device memory");
    for (int i = 1; i < nedges; ++i) FB_CHECK(edges[i] >= edges[i - 1], "fb_correlation_function: edges must ascend");
    if (ensure_work(p)) return -2;
    const int N = p->N;
    const size_t n3 = (size_t)N * N * N, nh = (size_t)p->na * N * N;
    const void *dfa = nullptr, *dfb = nullptr;
    if (stage_in(p, 2, field, n3 * sizeof(float), &dfa)) return -2;
    if (stage_in(p, 0, field_b, n3 * sizeof(float), &dfb)) return -2;
    // aux: [spectrum of b (cross only)] [lag cube unless xi_out] [edges, results]
    const size_t off_sb = 0, off_xi = dfb ? nh * sizeof(float2) : 0, off_tab = off_xi + (xi_out ? 0 : n3 * sizeof(float)),
                 off_cnt = off_tab + FB_MAX_EDGES * 8, off_s1 = off_cnt + (FB_MAX_EDGES + 1) * 8,
                 off_s2 = off_s1 + (FB_MAX_EDGES + 1) * 8, total = off_s2 + (FB_MAX_EDGES + 1) * 8;
    if (ensure_aux(p, total)) return -2;
    unsigned char* aux = (unsigned char*)p->aux;
    float2* sb = dfb ? (float2*)(aux + off_sb) : nullptr;
    float* xi = xi_out ? xi_out : (float*)(aux + off_xi);
    FB_CUDA(cudaMemcpyAsync(aux + off_tab, edges, nedges * sizeof(double), cudaMemcpyHostToDevice, p->stream));
    FB_CUDA(cudaMemsetAsync(aux + off_cnt, 0, total - off_cnt, p->stream));
    auto forward = [&](const float* f, float2* spec_store) -> int {     // f -> p->work (spectrum, in place)
        XArgs xa;
        memset(&xa, 0, sizeof(xa));
        xa.field_in = f;
        xa.spec_out = p->work;
        xa.tw = p->tw;
        xa.ncols = (size_t)N * N;
        if (launch_x_r2c(p, xa)) return -3;
        if (launch_cols(p, p->work, p->na, -1)) return -3;
        RowsArgs ra;
        memset(&ra, 0, sizeof(ra));
        ra.work = p->work;
        ra.spec_out = spec_store ? spec_store : p->work;
        ra.tw = p->tw;
        ra.nrows = (long)p->na * N;
        ra.K = p->kspace();
        ra.pk = p->pkdev();
        return launch_rows_fwd(p, ra);
    };
    if (dfb) {
        if (forward((const float*)dfb, sb)) return -3;
    }
    if (forward((const float*)dfa, nullptr)) return -3;
    const unsigned grid = (unsigned)(p->sm_count * 8);
    k_spec_power<<<grid, 256, 0, p->stream>>>(p->work, sb, nh, (float)(1.0 / ((double)N * N * N)));
    FB_LAUNCH_CHECK();
    {                                                                    // inverse of the power cube (in p->work)
        RowsArgs ra;
        memset(&ra, 0, sizeof(ra));
        ra.src = p->work;
        ra.work = p->work;
        ra.tw = p->tw;
        ra.nrows = (long)p->na * N;
        ra.kind = FB_KIND_PLAIN;
        ra.K = p->kspace();
        ra.pk = p->pkdev();
        if (launch_rows_inv_spec(p, ra)) return -3;
        if (launch_cols(p, p->work, p->na, +1)) return -3;
        XArgs xa;
        memset(&xa, 0, sizeof(xa));
        xa.spec = p->work;
        xa.field = xi;
        xa.tw = p->tw;
        xa.ncols = (size_t)N * N;
        xa.scale = (float)(1.0 / ((double)N * N * N));                   // numpy ifftn normalisation
        if (launch_x_c2r(p, xa)) return -3;
    }
    XiArgs xg;
    memset(&xg, 0, sizeof(xg));
    xg.xi = xi;
    xg.N = N;
    xg.hx = p->Lx / N;
    xg.hy = p->Ly / N;
    xg.hz = p->Lz / N;
    xg.edges = (const double*)(aux + off_tab);
    xg.nedges = nedges;
    xg.count = (unsigned long long*)(aux + off_cnt);
    xg.sum1 = (double*)(aux + off_s1);
    xg.sum2 = (double*)(aux + off_s2);
    k_xi_bin<<<p->sm_count * 4, 256, 0, p->stream>>>(xg);
    FB_LAUNCH_CHECK();
    FB_CUDA(cudaMemcpyAsync(count, aux + off_cnt, (size_t)(nedges + 1) * 8, cudaMemcpyDeviceToHost, p->stream));
    FB_CUDA(cudaMemcpyAsync(sum1, aux + off_s1, (size_t)(nedges + 1) * 8, cudaMemcpyDeviceToHost, p->stream));
    FB_CUDA(cudaMemcpyAsync(sum2, aux + off_s2, (size_t)(nedges + 1) * 8, cudaMemcpyDeviceToHost, p->stream));
    FB_CUDA(cudaStreamSynchronize(p->stream));
    return 0;
}

}  // extern "C"
