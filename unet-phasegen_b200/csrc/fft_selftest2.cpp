// Host-side check of stft_core.cuh (no GPU needed): runs the per-thread frame programs of the STFT and
// ISTFT kernels on TG host threads with a std::barrier in place of __syncwarp / bar.sync, and compares with
// a naive double-precision real DFT.
//   g++ -O2 -std=c++20 -pthread -o fft_selftest2.bin fft_selftest2.cpp && ./fft_selftest2.bin
#include <barrier>
#include <cmath>
#include <complex>
#include <cstdio>
#include <cstdlib>
#include <thread>
#include <vector>
#define PG_HD inline
#include "stft_core.cuh"
using namespace pgfft;

template <int NC> double check_forward() {
    constexpr int TG = NC / 16, NF = 2 * NC;
    using R = Radix<NC, false>;
    std::vector<cpx> tw(NF), tabs(R::TOTAL), s(padded_len2(NC));
    for (int m = 0; m < NF; ++m) tw[m] = {(float)cos(-2 * M_PI * m / NF), (float)sin(-2 * M_PI * m / NF)};
    for (int e = 0; e < R::TOTAL; ++e) tabs[e] = table_entry<NC, false>(tw.data(), e);
    std::vector<double> x(NF);
    for (auto& v : x) v = drand48() - 0.5;
    std::vector<std::complex<double>> got(NC + 1);
    std::barrier bar(TG);
    auto sync = [&] { bar.arrive_and_wait(); };
    auto prog = [&](int t) {
        cpx v[16];
        for (int r = 0; r < 16; ++r) { int m = t + r * TG; v[r] = {(float)(0.5 * x[2 * m]), (float)(0.5 * x[2 * m + 1])}; }
        fwd_phase0<NC>(s.data(), t, v);
        sync();
        fwd_phase1<NC>(s.data(), t, tabs.data(), sync);
        sync();
        auto emit = [&](int bin, cpx X) { got[bin] = {X.x, X.y}; };
        if (NC == 512) {
            fwd_fused_last_512(s.data(), t, tabs.data(), emit);
        } else {
            fwd_phase2_unfused<NC>(s.data(), t, tabs.data(), sync);
            sync();
            fwd_post_generic<NC>(s.data(), t, tabs.data(), emit);
        }
    };
    std::vector<std::thread> th;
    for (int t = 0; t < TG; ++t) th.emplace_back(prog, t);
    for (auto& q : th) q.join();
    double err = 0, nrm = 0;
    for (int k = 1; k <= NC; ++k) {
        std::complex<double> ref = 0;
        for (int n = 0; n < NF; ++n) ref += x[n] * std::polar(1.0, -2 * M_PI * (double)((long)k * n % NF) / NF);
        err += std::norm(ref - got[k]); nrm += std::norm(ref);
    }
    return sqrt(err / nrm);
}

template <int NC> double check_inverse() {
    constexpr int TG = NC / 16, NF = 2 * NC;
    using R = Radix<NC, true>;
    std::vector<cpx> tw(NF), tabs(R::TOTAL), s(padded_len2(NC)), win(NC);
    for (int m = 0; m < NF; ++m) tw[m] = {(float)cos(-2 * M_PI * m / NF), (float)sin(-2 * M_PI * m / NF)};
    for (int e = 0; e < R::TOTAL; ++e) tabs[e] = table_entry<NC, true>(tw.data(), e);
    std::vector<double> wn(NF);
    for (int n = 0; n < NF; ++n) wn[n] = 0.5 - 0.5 * cos(2 * M_PI * n / NF);
    for (int m = 0; m < NC; ++m) win[m] = {(float)wn[2 * m], (float)wn[2 * m + 1]};
    std::vector<std::complex<double>> X(NC + 1);
    for (int k = 1; k <= NC; ++k) X[k] = {drand48() - 0.5, drand48() - 0.5};
    X[0] = 0;
    std::barrier bar(TG);
    auto sync = [&] { bar.arrive_and_wait(); };
    auto prog = [&](int t) {
        cpx xin[16];
        for (int j = 0; j < 16; ++j) { const int bin = inv_bin<NC>(t, j); xin[j] = {(float)X[bin].real(), (float)X[bin].imag()}; }
        if (NC == 512) inv_fused_first_512(s.data(), t, tabs.data(), 0.5f / NC, xin);
        else inv_pre_generic<NC>(s.data(), t, tabs.data(), 0.5f / NC, xin);
        sync();
        inv_passes<NC>(s.data(), t, tabs.data(), win.data(), sync);
    };
    std::vector<std::thread> th;
    for (int t = 0; t < TG; ++t) th.emplace_back(prog, t);
    for (auto& q : th) q.join();
    // reference: irfft of the Hermitian completion (imaginary part of the Nyquist bin ignored), times the window
    double err = 0, nrm = 0;
    for (int n = 0; n < NF; ++n) {
        double acc = X[NC].real() * ((n & 1) ? -1.0 : 1.0);
        for (int k = 1; k < NC; ++k) acc += 2.0 * (X[k] * std::polar(1.0, 2 * M_PI * (double)((long)k * n % NF) / NF)).real();
        const double ref = acc / NF * wn[n];
        const cpx z = s[pad2(n >> 1)];
        const double g = (n & 1) ? z.y : z.x;
        err += (ref - g) * (ref - g); nrm += ref * ref;
    }
    return sqrt(err / nrm);
}

int main() {
    double worst = 0, e;
#define RUN(NC) e = check_forward<NC>(); printf("NC=%d stft frame rel %.3e\n", NC, e); worst = fmax(worst, e); \
                e = check_inverse<NC>(); printf("NC=%d istft frame rel %.3e\n", NC, e); worst = fmax(worst, e);
    RUN(128) RUN(256) RUN(512) RUN(1024)
    if (!(worst < 2e-6)) { printf("FAIL\n"); return 1; }
    printf("OK\n");
    return 0;
}
