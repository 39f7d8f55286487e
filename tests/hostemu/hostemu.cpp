// CPU emulation of the device control flow of the STFT / iSTFT / Griffin-Lim kernels.
//
// TEST INFRASTRUCTURE ONLY.  This file re-uses the exact phase functions the CUDA kernels
// call (riffusion-hobby_b200/csrc/rf_gl_phases.cuh, compiled for the host) and replaces
// "256 threads + __syncthreads()" by loops over tid, so that the index tables, the
// prime-factor FFT passes, the pair packing and the overlap-add chunking can be checked
// against torch on a box without a GPU.  It is never linked into librf_b200.so and the
// product has no CPU path.
//
// Build: g++ -O2 -shared -fPIC -I../../riffusion-hobby_b200/csrc hostemu.cpp \
//            ../../riffusion-hobby_b200/csrc/rf_plan.cpp -o librf_hostemu.so
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "rf_gl_phases.cuh"
#include "rf_plan.h"

namespace {
constexpr int NT = RF_NT;
std::string g_err;

template <typename F>
void phase(F f) {
    for (int tid = 0; tid < NT; ++tid) f(tid);
}

rf_gl_tables tables(const rf_plan_host& h) {
    rf_gl_tables tb;
    tb.wt_fwd = reinterpret_cast<const rf_c32*>(h.wt_fwd.data());
    tb.wt_inv = reinterpret_cast<const rf_c32*>(h.wt_inv.data());
    tb.pp = h.pp.data();
    tb.n_live = h.n_live;
    tb.n_even = h.n_even;
    tb.hop = h.H;
    return tb;
}

// emulates k_stft_pair over the whole grid for one clip
void emu_stft_clip(const rf_plan_host& h, const float* x, int L, int T, rf_c32* R) {
    const rf_gl_tables tb = tables(h);
    std::vector<rf_c32> V(2 * RF_PW);
    std::vector<float> xs(RF_PW + h.H);
    for (int pr = 0; 2 * pr < T; ++pr)
        for (int g = 0; g < 2; ++g) {
            const int t0 = 2 * pr;
            const bool has1 = t0 + 1 < T;
            phase([&](int tid) { rf_stage_x(tid, NT, xs.data(), x, L, t0, h.H); });
            phase([&](int tid) { rf_stft_pass_b(tid, NT, V.data(), xs.data(), tb, g, has1); });
            phase([&](int tid) { rf_pass_a<false>(tid, NT, V.data()); });
            phase([&](int tid) { rf_pass_c<false>(tid, NT, V.data()); });
            const int j0 = g ? tb.n_even : 0, j1 = g ? tb.n_live : tb.n_even;
            rf_c32* out0 = R + static_cast<size_t>(t0) * tb.n_live;
            phase([&](int tid) {
                rf_stft_post(tid, NT, V.data(), tb, j0, j1, out0, has1 ? out0 + tb.n_live : nullptr);
            });
        }
}

// emulates k_istft_chunk (+ k_ola_assemble) for one clip
void emu_istft_clip(const rf_plan_host& h, const float* S, const rf_c32* cur, const rf_c32* prev, int mode,
                    float momentum, int T, float* x) {
    const rf_gl_tables tb = tables(h);
    const int G = RF_CHUNK;
    const int nchunks = (T + G - 1) / G;
    const int PL = (G - 1) * h.H + h.W;
    std::vector<float> part(static_cast<size_t>(2) * nchunks * PL, 0.f);
    std::vector<rf_c32> V(2 * RF_PW);
    std::vector<float> ola(PL);
    for (int chunk = 0; chunk < nchunks; ++chunk)
        for (int g = 0; g < 2; ++g) {
            const int f0 = chunk * G;
            const int nf = std::min(G, T - f0);
            const int j0 = g ? tb.n_even : 0, j1 = g ? tb.n_live : tb.n_even;
            std::fill(ola.begin(), ola.end(), 0.f);
            for (int pr = 0; 2 * pr < nf; ++pr) {
                const int t0 = f0 + 2 * pr;
                const bool has1 = (2 * pr + 1) < nf;
                phase([&](int tid) { rf_istft_zero(tid, NT, V.data()); });
                rf_istft_in in;
                const size_t o0 = static_cast<size_t>(t0) * tb.n_live;
                in.S0 = S + o0;
                in.cur0 = cur + o0;
                in.prev0 = prev ? prev + o0 : nullptr;
                in.S1 = has1 ? S + o0 + tb.n_live : nullptr;
                in.cur1 = cur + o0 + tb.n_live;
                in.prev1 = prev ? prev + o0 + tb.n_live : nullptr;
                in.mode = mode;
                in.momentum = momentum;
                phase([&](int tid) { rf_istft_load(tid, NT, V.data(), tb, j0, j1, in); });
                phase([&](int tid) { rf_pass_c<true>(tid, NT, V.data()); });
                phase([&](int tid) { rf_pass_a<true>(tid, NT, V.data()); });
                // device: one call with which=2 (barrier between the real- and imaginary-part adds)
                phase([&](int tid) { rf_istft_pass_b(tid, NT, V.data(), ola.data() + 2 * pr * h.H, tb, g, has1, 0); });
                phase([&](int tid) { rf_istft_pass_b(tid, NT, V.data(), ola.data() + 2 * pr * h.H, tb, g, has1, 1); });
            }
            std::memcpy(&part[(static_cast<size_t>(g) * nchunks + chunk) * PL], ola.data(), PL * sizeof(float));
        }
    std::vector<float> win2(h.W);
    for (int i = 0; i < h.W; ++i) win2[i] = h.window[i] * h.window[i];
    const int L = h.H * (T - 1);
    for (int i = 0; i < L; ++i)
        x[i] = rf_ola_sample(i, part.data(), rf_envelope(i, win2.data(), T, h.H, h.W), T, G, PL, nchunks, h.H, h.W);
}
}  // namespace

extern "C" {

const char* emu_last_error() { return g_err.c_str(); }

void* emu_plan_create(const rf_plan_desc* d, const float* window, const float* fb) {
    rf_plan_host* h = new rf_plan_host();
    int code = 0;
    g_err = rf_plan_build_host(*d, window, fb, *h, code);
    if (code != RF_OK) {
        delete h;
        return nullptr;
    }
    return h;
}
void emu_plan_destroy(void* p) { delete static_cast<rf_plan_host*>(p); }
int emu_plan_n_live(void* p) { return static_cast<rf_plan_host*>(p)->n_live; }
void emu_plan_bins(void* p, int32_t* out) {
    auto* h = static_cast<rf_plan_host*>(p);
    std::memcpy(out, h->bins.data(), h->bins.size() * 4);
}

// STFT of one clip: x[L] -> spec[F][T] complex64 (torchaudio layout; dead bins = 0)
void emu_stft(void* p, const float* x, int L, float* spec) {
    auto* h = static_cast<rf_plan_host*>(p);
    const int T = 1 + L / h->H;
    std::vector<rf_c32> R(static_cast<size_t>(T) * h->n_live);
    emu_stft_clip(*h, x, L, T, R.data());
    std::memset(spec, 0, static_cast<size_t>(h->F) * T * 8);
    for (int t = 0; t < T; ++t)
        for (int j = 0; j < h->n_live; ++j) {
            const size_t o = (static_cast<size_t>(h->bins[j]) * T + t) * 2;
            spec[o] = R[static_cast<size_t>(t) * h->n_live + j].x;
            spec[o + 1] = R[static_cast<size_t>(t) * h->n_live + j].y;
        }
}

// Griffin-Lim of one clip, same buffer rotation as gl_loop() in rf_audio.cu.
// lin[F][T], angles[F][T] complex64 (or null), wave[hop*(T-1)]
void emu_griffinlim(void* p, const float* lin, const float* angles, int T, int n_iter, float momentum_in,
                    float* wave) {
    auto* h = static_cast<rf_plan_host*>(p);
    const size_t n = static_cast<size_t>(T) * h->n_live;
    std::vector<float> S(n);
    std::vector<rf_c32> R0(n), R1(n);
    rf_c32* R[2] = {R0.data(), R1.data()};
    for (int t = 0; t < T; ++t)
        for (int j = 0; j < h->n_live; ++j) {
            const size_t src = static_cast<size_t>(h->bins[j]) * T + t;
            S[static_cast<size_t>(t) * h->n_live + j] = lin[src];
            R1[static_cast<size_t>(t) * h->n_live + j] =
                angles ? c_make(angles[2 * src], angles[2 * src + 1]) : c_make(1.f, 0.f);
        }
    const float m = static_cast<float>(static_cast<double>(momentum_in) / (1.0 + static_cast<double>(momentum_in)));
    const int L = h->H * (T - 1);
    for (int it = 0; it <= n_iter; ++it) {
        const rf_c32* cur;
        const rf_c32* prev = nullptr;
        int mode;
        if (it == 0) {
            cur = R[1];
            mode = 0;
        } else {
            cur = R[(it - 1) & 1];
            mode = 1;
            if (it >= 2 && m != 0.f) prev = R[it & 1];
        }
        emu_istft_clip(*h, S.data(), cur, prev, mode, m, T, wave);
        if (it == n_iter) break;
        emu_stft_clip(*h, wave, L, T, R[it & 1]);
    }
}
}
