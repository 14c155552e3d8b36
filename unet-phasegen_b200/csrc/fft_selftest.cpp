// Host-side check of fft_core.cuh (no GPU needed): emulates the thread groups of a
// Stockham pass sequentially and compares with a naive double-precision DFT.
//   g++ -O2 -std=c++17 -o /tmp/fft_selftest fft_selftest.cpp && /tmp/fft_selftest
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>
#define PG_HD inline
#include "fft_core.cuh"
using namespace pgfft;

template <int NC, int R, int NS, bool INV>
void run_pass(std::vector<float>& re, std::vector<float>& im, const std::vector<cpx>& tw) {
    constexpr int TG = NC / 16;
    std::vector<Pass<NC, R, NS, INV>> th(TG);
    for (int t = 0; t < TG; ++t) th[t].load(re.data(), im.data(), t);
    for (int t = 0; t < TG; ++t) { th[t].twiddle_butterfly(tw.data(), t); th[t].store(re.data(), im.data(), t); }
}

template <int NC, bool INV> double check() {
    using P = Plan<NC>;
    std::vector<cpx> tw(2 * NC);
    for (int m = 0; m < 2 * NC; ++m) tw[m] = {(float)cos(-M_PI * m / NC), (float)sin(-M_PI * m / NC)};
    std::vector<float> re(padded_len(NC)), im(padded_len(NC));
    std::vector<double> xr(NC), xi(NC);
    for (int i = 0; i < NC; ++i) {
        xr[i] = drand48() - 0.5; xi[i] = drand48() - 0.5;
        re[pad(i)] = (float)xr[i]; im[pad(i)] = (float)xi[i];
    }
    run_pass<NC, P::R0, 1, INV>(re, im, tw);
    run_pass<NC, P::R1, P::R0, INV>(re, im, tw);
    if (P::R2 > 1) run_pass<NC, (P::R2 > 1 ? P::R2 : 2), P::R0 * P::R1, INV>(re, im, tw);
    double err = 0, nrm = 0;
    for (int k = 0; k < NC; ++k) {
        double sr = 0, si = 0;
        for (int n = 0; n < NC; ++n) {
            double a = (INV ? 2.0 : -2.0) * M_PI * (double)((long)k * n % NC) / NC;
            sr += xr[n] * cos(a) - xi[n] * sin(a);
            si += xr[n] * sin(a) + xi[n] * cos(a);
        }
        err += (sr - re[pad(k)]) * (sr - re[pad(k)]) + (si - im[pad(k)]) * (si - im[pad(k)]);
        nrm += sr * sr + si * si;
    }
    return sqrt(err / nrm);
}

int main() {
    double worst = 0, e;
#define RUN(NC) e = check<NC, false>(); printf("NC=%d fwd rel %.3e\n", NC, e); worst = fmax(worst, e); \
                e = check<NC, true>();  printf("NC=%d inv rel %.3e\n", NC, e); worst = fmax(worst, e);
    RUN(128) RUN(256) RUN(512) RUN(1024)
    if (!(worst < 2e-6)) { printf("FAIL\n"); return 1; }
    printf("OK\n");
    return 0;
}
