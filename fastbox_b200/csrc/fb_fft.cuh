// fb_fft.cuh -- register-resident Stockham FFT building blocks (sm_100a).
//
// One FFT of length n is computed by T = n/P threads; every thread keeps P
// complex points in registers.  Element ownership is the same at the input of
// every stage and at the final output:  thread t holds elements  t + T*q,
// q = 0..P-1  (register q).  A stage of radix R (R | P) lets each thread do
// P/R butterflies in registers; between stages the points are exchanged
// through shared memory (write at the Stockham position j0 + r*Ns, read back
// at t + T*q).  This replaces numpy.fft / pocketfft used by the reference at
// fastbox/box.py:187,193,246,337,380,654,736.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace fb {

#define FB_NMAX_TW 4096      // longest transform the twiddle tables cover
// The plan owns 2*FB_NMAX_TW entries (forward sign): the table of every power-of-two length n sits at
// [n, 2n), so w_n^k = tw[n + k].  Lanes with consecutive k read consecutive entries (a subsampled
// master table made each lane hit its own 32-byte sector, which saturated the L1 data path).
#define FB_TW_ENTRIES (2 * FB_NMAX_TW)
#define FB_TW(tw, n, k) FB_LDG(&(tw)[(n) + (k)])

// FB_DEV functions also compile for the host so the CPU unit test (tests/host_fft_check.cu)
// can run the exact index logic thread by thread.
#define FB_DEV __host__ __device__ __forceinline__
#ifdef __CUDA_ARCH__
#define FB_LDG(p) __ldg(p)
#define FB_SYNC() __syncthreads()
#else
#define FB_LDG(p) (*(p))
#define FB_SYNC() ((void)0)
#endif

// Complex arithmetic on the packed 2 x fp32 instructions of sm_100 (FADD2 / FMUL2 / FFMA2): one issue slot per
// complex add, two per complex multiply (the rotated operand (-y, x) and a scalar broadcast are operand modifiers
// in SASS, not instructions).  The FFT kernels are issue-slot bound next to HBM, and ~3/4 of their instructions
// were scalar FADD / FMUL / FFMA.  FB_PACKED_F32=0 (and the host build) keeps the scalar forms.
#ifndef FB_PACKED_F32
#define FB_PACKED_F32 1
#endif
#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ >= 1000) && FB_PACKED_F32
#define FB_PK2 1
#else
#define FB_PK2 0
#endif
FB_DEV float2 cadd(float2 a, float2 b) {
#if FB_PK2
    return __fadd2_rn(a, b);
#else
    return make_float2(a.x + b.x, a.y + b.y);
#endif
}
FB_DEV float2 csub(float2 a, float2 b) {
#if FB_PK2
    return __ffma2_rn(b, make_float2(-1.f, -1.f), a);
#else
    return make_float2(a.x - b.x, a.y - b.y);
#endif
}
FB_DEV float2 cmul(float2 a, float2 b) {
#if FB_PK2
    return __ffma2_rn(make_float2(a.y, a.y), make_float2(-b.y, b.x), __fmul2_rn(make_float2(a.x, a.x), b));
#else
    return make_float2(fmaf(a.x, b.x, -a.y * b.y), fmaf(a.x, b.y, a.y * b.x));
#endif
}
// a * (c + i s) with compile-time-known c, s
FB_DEV float2 cmul_const(float2 a, float c, float s) {
#if FB_PK2
    return __ffma2_rn(make_float2(-a.y, a.x), make_float2(s, s), __fmul2_rn(a, make_float2(c, c)));
#else
    return make_float2(fmaf(a.x, c, -a.y * s), fmaf(a.x, s, a.y * c));
#endif
}
FB_DEV float2 cscale(float2 a, float h) {
#if FB_PK2
    return __fmul2_rn(a, make_float2(h, h));
#else
    return make_float2(a.x * h, a.y * h);
#endif
}
FB_DEV float2 cconj(float2 a) { return make_float2(a.x, -a.y); }

// cos(pi*m/16), m = 0..16
__host__ __device__ constexpr float cos_pi16(int m) {
    constexpr double C[17] = {1.0, 0.98078528040323044913, 0.92387953251128675613, 0.83146961230254523708,
                              0.70710678118654752440, 0.55557023301960222474, 0.38268343236508977173,
                              0.19509032201612826785, 0.0, -0.19509032201612826785, -0.38268343236508977173,
                              -0.55557023301960222474, -0.70710678118654752440, -0.83146961230254523708,
                              -0.92387953251128675613, -0.98078528040323044913, -1.0};
    return (float)C[m];
}
__host__ __device__ constexpr float sin_pi16(int m) {   // sin(pi*m/16) = cos(pi*(8-m)/16), m = 0..16
    return m <= 8 ? cos_pi16(8 - m) : cos_pi16(m - 8);
}

// multiply by exp(S * i * pi * M16 / 16), M16 in [0,16), compile-time
template <int M16, int S>
FB_DEV float2 ctwiddle(float2 a) {
    if constexpr (M16 == 0) {
        return a;
    } else if constexpr (M16 == 8) {          // * (S i)
        return S > 0 ? make_float2(-a.y, a.x) : make_float2(a.y, -a.x);
    } else if constexpr (M16 == 4) {          // * (1 + S i)/sqrt2 = (a + S i a) / sqrt2
        constexpr float h = 0.70710678118654752440f;
        return cscale(cadd(a, S > 0 ? make_float2(-a.y, a.x) : make_float2(a.y, -a.x)), h);
    } else if constexpr (M16 == 12) {         // * (-1 + S i)/sqrt2 = -(a - S i a) / sqrt2
        constexpr float h = 0.70710678118654752440f;
        return cscale(cadd(a, S > 0 ? make_float2(a.y, -a.x) : make_float2(-a.y, a.x)), -h);
    } else {
        constexpr float c = cos_pi16(M16);
        constexpr float s = (S > 0 ? 1.f : -1.f) * sin_pi16(M16);
        return cmul_const(a, c, s);
    }
}

// Natural-order in / natural-order out DFT of R points held in registers.
template <int R, int S>
struct Dft;

template <int S>
struct Dft<1, S> {
    FB_DEV static void run(float2 (&)[1]) {}
};
template <int S>
struct Dft<2, S> {
    FB_DEV static void run(float2 (&a)[2]) {
        float2 t = a[0];
        a[0] = cadd(t, a[1]);
        a[1] = csub(t, a[1]);
    }
};
template <int S>
struct Dft<4, S> {
    FB_DEV static void run(float2 (&a)[4]) {
        float2 b0 = cadd(a[0], a[2]), b1 = csub(a[0], a[2]);
        float2 b2 = cadd(a[1], a[3]), b3 = ctwiddle<8, S>(csub(a[1], a[3]));
        a[0] = cadd(b0, b2);
        a[2] = csub(b0, b2);
        a[1] = cadd(b1, b3);
        a[3] = csub(b1, b3);
    }
};
template <int R, int S>
struct Dft {
    FB_DEV static void run(float2 (&a)[R]) {
        constexpr int H = R / 2;
        float2 e[H], o[H];
#pragma unroll
        for (int i = 0; i < H; ++i) {
            e[i] = a[2 * i];
            o[i] = a[2 * i + 1];
        }
        Dft<H, S>::run(e);
        Dft<H, S>::run(o);
        combine<0>(a, e, o);
    }
    template <int K>
    FB_DEV static void combine(float2 (&a)[R], float2 (&e)[R / 2], float2 (&o)[R / 2]) {
        if constexpr (K < R / 2) {
            float2 w = ctwiddle<(K * 32) / R, S>(o[K]);
            a[K] = cadd(e[K], w);
            a[K + R / 2] = csub(e[K], w);
            combine<K + 1>(a, e, o);
        }
    }
};

// ---------------------------------------------------------------------------
// Shared-memory layouts for the exchange buffer.
//   RowLayout : one FFT per smem row, padded one float2 every 16 (bank spread
//               for the stride-R Stockham writes).
//   ColLayout : CZ FFTs interleaved (element-major), one pad row every 16 elements so
//               that the stride-16 Stockham writes of narrow tiles (CZ = 4, 8) spread over banks.
// ---------------------------------------------------------------------------
// Both layouts are linear in the logical index except for the pad every 16 elements, so
//   phys(i0 + r*S) = phys(i0) + r*pstride(S)   when S % 16 == 0, or S == 1 with i0 % 16 == 0, r < 16.
// The exchange code uses this to address a whole butterfly from one base (no shifts per access).
template <int n>
struct RowLayout {
    static constexpr int ROW = n + n / 16;
    int base;
    FB_DEV int operator()(int i) const { return base + i + (i >> 4); }
    __host__ __device__ static constexpr int pstride(int S) { return S % 16 == 0 ? S + S / 16 : S; }
};
template <int CZ>
struct ColLayout {
    int col;
    FB_DEV int operator()(int i) const { return (i + (i >> 4)) * CZ + col; }   // one padded row per 16
    __host__ __device__ static constexpr int pstride(int S) { return (S % 16 == 0 ? S + S / 16 : S) * CZ; }
};

// a * exp(S i pi m / 16), m a compile-time constant after unrolling (the table look-up folds away)
FB_DEV float2 rot_pi16(float2 a, int m, int S) {
    return cmul_const(a, cos_pi16(m), (S > 0 ? 1.f : -1.f) * sin_pi16(m));
}

template <int n, int P, int R, int Ns, int S>
FB_DEV void fft_stage(float2 (&v)[P], int t, const float2* __restrict__ tw) {
    constexpr int T = n / P;
    constexpr int B = P / R;
    // Last stage (Ns * R == n): butterfly u of thread t has j = t + u T without wrap, so its base twiddle is
    // w_n^t * w_P^u -- ONE table load per thread and a compile-time rotation per butterfly instead of B loads
    // (memory-instruction slots, not arithmetic, are what these kernels run out of).
    constexpr bool ONE_LOAD = (Ns > 1) && (Ns * R == n) && (P == 16) && (B > 1);
    float2 wbase = make_float2(1.f, 0.f);
    if constexpr (ONE_LOAD) {
        wbase = FB_TW(tw, n, t);
        if (S > 0) wbase.y = -wbase.y;
    }
#pragma unroll
    for (int u = 0; u < B; ++u) {
        float2 a[R];
#pragma unroll
        for (int r = 0; r < R; ++r) a[r] = v[u + r * B];
        if constexpr (Ns > 1) {
            const int jm = (t + u * T) & (Ns - 1);
            float2 w1;
            if constexpr (ONE_LOAD) {
                w1 = u == 0 ? wbase : rot_pi16(wbase, 2 * u, S);
            } else {
                w1 = FB_TW(tw, Ns * R, jm);
                if (S > 0) w1.y = -w1.y;
            }
            if constexpr (R >= 4) {
                // twiddles w^r, r = 1..R-1, from ONE table load: powers by a product tree of depth
                // <= 4 (each lane's R-1 twiddles are distinct, so loading them all costs ~R sector
                // look-ups per lane in L1 -- far more than the data itself)
                float2 w[R];
                w[1] = w1;
#pragma unroll
                for (int r = 2; r < R; ++r) w[r] = cmul(w[r / 2], w[r - r / 2]);
#pragma unroll
                for (int r = 1; r < R; ++r) a[r] = cmul(a[r], w[r]);
            } else if constexpr (R == 2) {
                a[1] = cmul(a[1], w1);
            } else {
#pragma unroll
                for (int r = 1; r < R; ++r) {
                    float2 w = FB_TW(tw, Ns * R, r * jm);
                    if (S > 0) w.y = -w.y;
                    a[r] = cmul(a[r], w);
                }
            }
        }
        Dft<R, S>::run(a);
#pragma unroll
        for (int r = 0; r < R; ++r) v[u + r * B] = a[r];
    }
}

// write the outputs of a radix-R stage (previous product Ns) to smem ...
template <int n, int P, int R, int Ns, class SL>
FB_DEV void fft_exchange_write(const float2 (&v)[P], int t, float2* sm, const SL& sl) {
    constexpr int T = n / P;
    constexpr int B = P / R;
    static_assert(Ns % 16 == 0 || (Ns == 1 && R == 16), "exchange strides assume radix-16 first stages");
    constexpr int PS = SL::pstride(Ns);
#pragma unroll
    for (int u = 0; u < B; ++u) {
        const int j = t + u * T;
        const int j0 = (j / Ns) * (Ns * R) + (j & (Ns - 1));
        float2* p = sm + sl(j0);
#pragma unroll
        for (int r = 0; r < R; ++r) p[r * PS] = v[u + r * B];
    }
}
// ... and read back the inputs of the next stage (thread t owns elements t + T*q).
template <int n, int P, class SL>
FB_DEV void fft_exchange_read(float2 (&v)[P], int t, const float2* sm, const SL& sl) {
    constexpr int T = n / P;
    if constexpr (T % 16 == 0) {
        const float2* p = sm + sl(t);
        constexpr int PS = SL::pstride(T);
#pragma unroll
        for (int q = 0; q < P; ++q) v[q] = p[q * PS];
    } else {
#pragma unroll
        for (int q = 0; q < P; ++q) v[q] = sm[sl(t + T * q)];
    }
}
// store thread-owned elements t + T*q at their natural positions
template <int n, int P, class SL>
FB_DEV void fft_store_natural(const float2 (&v)[P], int t, float2* sm, const SL& sl) {
    constexpr int T = n / P;
    if constexpr (T % 16 == 0) {
        float2* p = sm + sl(t);
        constexpr int PS = SL::pstride(T);
#pragma unroll
        for (int q = 0; q < P; ++q) p[q * PS] = v[q];
    } else {
#pragma unroll
        for (int q = 0; q < P; ++q) sm[sl(t + T * q)] = v[q];
    }
}
// Two barriers: the buffer is reused in place.
template <int n, int P, int R, int Ns, class SL>
FB_DEV void fft_exchange(float2 (&v)[P], int t, float2* sm, const SL& sl, bool trailing_sync) {
    fft_exchange_write<n, P, R, Ns, SL>(v, t, sm, sl);
    FB_SYNC();
    fft_exchange_read<n, P, SL>(v, t, sm, sl);
    if (trailing_sync) FB_SYNC();
}

// Full length-n transform of the register-resident points.
template <int n, int P, int R1, int R2, int R3, int S, class SL>
FB_DEV void fft_regs(float2 (&v)[P], int t, float2* sm, const SL& sl,
                                         const float2* __restrict__ tw) {
    static_assert(R1 * R2 * R3 == n, "radix product");
    fft_stage<n, P, R1, 1, S>(v, t, tw);
    if constexpr (R2 > 1) {
        fft_exchange<n, P, R1, 1, SL>(v, t, sm, sl, R3 > 1);
        fft_stage<n, P, R2, R1, S>(v, t, tw);
    }
    if constexpr (R3 > 1) {
        fft_exchange<n, P, R2, R1, SL>(v, t, sm, sl, false);
        fft_stage<n, P, R3, R1 * R2, S>(v, t, tw);
    }
}

// The barriers of fft_regs, for warps of a CTA that sit a transform out (they must arrive at the
// same __syncthreads as the warps that run it).
template <int R2, int R3>
FB_DEV void fft_regs_barriers_only() {
    if constexpr (R2 > 1) {
        FB_SYNC();
        if constexpr (R3 > 1) FB_SYNC();
    }
    if constexpr (R3 > 1) FB_SYNC();
}

// compile-time FFT configuration for each supported length
template <int n>
struct FftCfg;
#define FB_CFG(N_, P_, A_, B_, C_)                                              \
    template <>                                                                 \
    struct FftCfg<N_> {                                                         \
        static constexpr int n = N_, P = P_, R1 = A_, R2 = B_, R3 = C_, T = N_ / P_; \
    };
FB_CFG(4, 4, 4, 1, 1)
FB_CFG(8, 8, 8, 1, 1)
FB_CFG(16, 16, 16, 1, 1)
FB_CFG(32, 16, 16, 2, 1)
FB_CFG(64, 16, 16, 4, 1)
FB_CFG(128, 16, 16, 8, 1)
FB_CFG(256, 16, 16, 16, 1)
FB_CFG(512, 16, 16, 16, 2)
FB_CFG(1024, 16, 16, 16, 4)
FB_CFG(2048, 16, 16, 16, 8)
#undef FB_CFG

}  // namespace fb
