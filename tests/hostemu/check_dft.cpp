// Host check of the in-register DFT butterflies against a naive fp64 DFT.
// Build: g++ -O2 -I../../riffusion-hobby_b200/csrc check_dft.cpp -o check_dft
#include <cmath>
#include <complex>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "rf_dft.cuh"

template <int P, bool INV, typename F>
static double check(F fn, const char* name) {
    std::vector<rf_c32> v(P);
    std::vector<std::complex<double>> x(P), ref(P);
    srand(P);
    for (int i = 0; i < P; ++i) {
        v[i].x = (float)rand() / RAND_MAX - 0.5f;
        v[i].y = (float)rand() / RAND_MAX - 0.5f;
        x[i] = {v[i].x, v[i].y};
    }
    const double sg = INV ? 1.0 : -1.0;
    for (int k = 0; k < P; ++k) {
        std::complex<double> s = 0;
        for (int n = 0; n < P; ++n) s += x[n] * std::polar(1.0, sg * 2 * M_PI * n * k / P);
        ref[k] = s;
    }
    fn(v.data());
    double err = 0;
    for (int k = 0; k < P; ++k) {
        int pos = (P == 49) ? dft49_out_index(k) : k;
        err = std::fmax(err, std::abs(std::complex<double>(v[pos].x, v[pos].y) - ref[k]));
    }
    printf("%s P=%d inv=%d max_err=%.3e\n", name, P, (int)INV, err);
    return err;
}

int main() {
    double e = 0;
    e = std::fmax(e, check<3, false>([](rf_c32* v) { dft3<false>(v[0], v[1], v[2]); }, "dft3"));
    e = std::fmax(e, check<3, true>([](rf_c32* v) { dft3<true>(v[0], v[1], v[2]); }, "dft3"));
    e = std::fmax(e, check<5, false>([](rf_c32* v) { dft5<false>(v[0], v[1], v[2], v[3], v[4]); }, "dft5"));
    e = std::fmax(e, check<5, true>([](rf_c32* v) { dft5<true>(v[0], v[1], v[2], v[3], v[4]); }, "dft5"));
    e = std::fmax(e, check<7, false>([](rf_c32* v) { dft7<false>(v[0], v[1], v[2], v[3], v[4], v[5], v[6]); }, "dft7"));
    e = std::fmax(e, check<7, true>([](rf_c32* v) { dft7<true>(v[0], v[1], v[2], v[3], v[4], v[5], v[6]); }, "dft7"));
    e = std::fmax(e, check<9, false>([](rf_c32* v) { dft9<false>(v); }, "dft9"));
    e = std::fmax(e, check<9, true>([](rf_c32* v) { dft9<true>(v); }, "dft9"));
    e = std::fmax(e, check<10, false>([](rf_c32* v) { dft10<false>(v); }, "dft10"));
    e = std::fmax(e, check<10, true>([](rf_c32* v) { dft10<true>(v); }, "dft10"));
    e = std::fmax(e, check<49, false>([](rf_c32* v) { dft49<false>(v); }, "dft49"));
    e = std::fmax(e, check<49, true>([](rf_c32* v) { dft49<true>(v); }, "dft49"));
    if (e > 2e-5) { printf("FAIL\n"); return 1; }
    printf("OK\n");
    return 0;
}
