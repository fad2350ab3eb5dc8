// CPU unit test of the register-FFT building blocks (fb_fft.cuh compiled for the host):
// runs the per-thread stage / exchange logic thread by thread and compares with a
// direct O(n^2) DFT in double.  Built and run by tests/test_host_fft.py.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>
#include "../../fastbox_b200/csrc/fb_fft.cuh"

using namespace fb;

static std::vector<float2> g_tw;

template <int n, int S, class SLF>
static double run_case(SLF make_layout, int smem_len) {
    using C = FftCfg<n>;
    constexpr int P = C::P, T = C::T;
    std::vector<float2> x(n), sm(smem_len);
    for (int i = 0; i < n; ++i) x[i] = make_float2((float)drand48() - 0.5f, (float)drand48() - 0.5f);
    std::vector<std::vector<float2>> regs(T, std::vector<float2>(P));
    auto as_arr = [&](int t) -> float2(&)[P] { return *reinterpret_cast<float2(*)[P]>(regs[t].data()); };
    for (int t = 0; t < T; ++t)
        for (int q = 0; q < P; ++q) regs[t][q] = x[t + T * q];
    auto sl = make_layout();
    for (int t = 0; t < T; ++t) fft_stage<n, P, C::R1, 1, S>(as_arr(t), t, g_tw.data());
    if constexpr (C::R2 > 1) {
        for (int t = 0; t < T; ++t) fft_exchange_write<n, P, C::R1, 1>(as_arr(t), t, sm.data(), sl);
        for (int t = 0; t < T; ++t) fft_exchange_read<n, P>(as_arr(t), t, sm.data(), sl);
        for (int t = 0; t < T; ++t) fft_stage<n, P, C::R2, C::R1, S>(as_arr(t), t, g_tw.data());
    }
    if constexpr (C::R3 > 1) {
        for (int t = 0; t < T; ++t) fft_exchange_write<n, P, C::R2, C::R1>(as_arr(t), t, sm.data(), sl);
        for (int t = 0; t < T; ++t) fft_exchange_read<n, P>(as_arr(t), t, sm.data(), sl);
        for (int t = 0; t < T; ++t) fft_stage<n, P, C::R3, C::R1 * C::R2, S>(as_arr(t), t, g_tw.data());
    }
    double num = 0, den = 0;
    for (int k = 0; k < n; ++k) {
        double re = 0, im = 0;
        for (int i = 0; i < n; ++i) {
            const double ang = S * 2.0 * M_PI * (double)((long)i * k % n) / n;
            re += x[i].x * cos(ang) - x[i].y * sin(ang);
            im += x[i].x * sin(ang) + x[i].y * cos(ang);
        }
        const float2 g = regs[k % T][k / T];
        num += (g.x - re) * (g.x - re) + (g.y - im) * (g.y - im);
        den += re * re + im * im;
    }
    return sqrt(num / den);
}

template <int n>
static int check() {
    int bad = 0;
    const double e1 = run_case<n, -1>([] { return RowLayout<n>{0}; }, RowLayout<n>::ROW);
    const double e2 = run_case<n, +1>([] { return RowLayout<n>{0}; }, RowLayout<n>::ROW);
    const double e3 = run_case<n, -1>([] { return ColLayout<4>{1}; }, (n + n / 16) * 4);
    const double e4 = run_case<n, +1>([] { return ColLayout<4>{3}; }, (n + n / 16) * 4);
    printf("n=%4d  row fwd %.2e inv %.2e  col fwd %.2e inv %.2e\n", n, e1, e2, e3, e4);
    if (!(e1 < 2e-6 && e2 < 2e-6 && e3 < 2e-6 && e4 < 2e-6)) bad = 1;
    return bad;
}

int main() {
    g_tw.assign(FB_TW_ENTRIES, make_float2(1.f, 0.f));
    for (int len = 1; len <= FB_NMAX_TW; len *= 2)
    for (int m = 0; m < len; ++m) {
        const double ang = -2.0 * M_PI * m / len;
        g_tw[len + m] = make_float2((float)cos(ang), (float)sin(ang));
    }
    int bad = 0;
    bad |= check<4>();
    bad |= check<8>();
    bad |= check<16>();
    bad |= check<32>();
    bad |= check<64>();
    bad |= check<128>();
    bad |= check<256>();
    bad |= check<512>();
    bad |= check<1024>();
    bad |= check<2048>();
    printf(bad ? "FAIL\n" : "OK\n");
    return bad;
}
