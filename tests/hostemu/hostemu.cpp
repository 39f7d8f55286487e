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
#include <cmath>
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

// other = true (NA = 5 only): the inverse tables of the other sample parity (rf_plan_host::t5e), as gl_loop() builds `tbo`
rf_gl_tables tables(const rf_plan_host& h, int NA, bool other = false) {
    rf_gl_tables tb;
    const rf_bin_tabs& t = NA == 10 ? h.t10 : (other ? h.t5e : h.t5);
    tb.wg_fwd = reinterpret_cast<const rf_f4*>(t.wg_fwd.data());
    tb.wg_inv = reinterpret_cast<const rf_f4*>(t.wg_inv.data());
    tb.items = t.items.data();
    tb.bt = t.bt.data();
    tb.ab_inv = reinterpret_cast<const rf_f4*>(t.ab_inv.data());
    tb.ab_fwd = reinterpret_cast<const rf_f4*>(t.ab_fwd.data());
    tb.zpos = t.zpos.data();
    tb.nz0 = t.nz[0];
    tb.nz1 = t.nz[1];
    static std::vector<float> zero_row;
    if (static_cast<int>(zero_row.size()) < h.n_live) zero_row.assign(h.n_live, 0.f);
    tb.zero_row = zero_row.data();
    tb.off1 = NA == 10 ? h.H : (other ? (h.H - 1) / 2 : (h.H + 1) / 2);
    tb.n_live = h.n_live;
    tb.n_even = h.n_even;
    return tb;
}

// emulates stft_pair_body<NA> for the frame pairs [pr_lo, pr_hi) of one clip.  NA = 10: x holds waveform samples
// [base, ...); NA = 5: x is the odd-sample waveform
template <int NA>
void emu_stft_pairs(const rf_plan_host& h, const float* x, int base, int L, int T, int pr_lo, int pr_hi, rf_c32* R) {
    const rf_gl_tables tb = tables(h, NA);
    constexpr int W = rf_geom<NA>::W;
    std::vector<rf_c32> V(2 * W);
    std::vector<float> xs(W + tb.off1);
    for (int pr = pr_lo; pr < pr_hi; ++pr)
        for (int g = 0; g < 2; ++g) {
            const int t0 = 2 * pr;
            const bool has1 = t0 + 1 < T;
            if (NA == 10) phase([&](int tid) { rf_stage_x(tid, NT, xs.data(), x, L, t0, h.H, base); });
            else phase([&](int tid) { rf_stage_x_d2(tid, NT, xs.data(), x, L, t0, h.H); });
            phase([&](int tid) { rf_stft_pass_b<NA>(tid, NT, V.data(), xs.data(), tb, g, has1); });
            phase([&](int tid) { rf_pass_a<false, NA>(tid, NT, V.data()); });
            phase([&](int tid) { rf_pass_c7<false, NA, 0>(tid, NT, V.data()); });
            phase([&](int tid) { rf_pass_c7<false, NA, 1>(tid, NT, V.data()); });
            const int j0 = g ? tb.n_even : 0, j1 = g ? tb.n_live : tb.n_even;
            rf_c32* out0 = R + static_cast<size_t>(t0) * tb.n_live;
            phase([&](int tid) {
                rf_stft_post<NA>(tid, NT, V.data(), tb, j0, j1, out0, has1 ? out0 + tb.n_live : nullptr);
            });
        }
}

template <int NA>
void emu_stft_clip(const rf_plan_host& h, const float* x, int L, int T, rf_c32* R) {
    emu_stft_pairs<NA>(h, x, 0, L, T, 0, (T + 1) / 2, R);
}

// emulates istft_chunk_body<NA> for one (chunk, group): dst[PL]
template <int NA>
void emu_istft_chunk(const rf_plan_host& h, const float* S, const rf_c32* cur, const rf_c32* prev, int mode,
                     float momentum, int T, int PL, int g, int chunk, float* dst, bool other = false) {
    const rf_gl_tables tb = tables(h, NA, other);
    constexpr int W = rf_geom<NA>::W;
    const int G = RF_CHUNK;
    const int pair_stride = NA == 10 ? 2 * h.H : h.H;
    std::vector<rf_c32> V(2 * W, c_make(NAN, NAN));   // shared memory starts out as garbage on the device
    std::vector<float> ola(PL, 0.f);
    const int f0 = chunk * G;
    const int nf = std::min(G, T - f0);
    const int j0 = g ? tb.n_even : 0, j1 = g ? tb.n_live : tb.n_even;
    for (int pr = 0; 2 * pr < nf; ++pr) {
        const int t0 = f0 + 2 * pr;
        const bool has1 = (2 * pr + 1) < nf;
        phase([&](int tid) { rf_istft_zero<NA>(tid, NT, V.data(), tb, g); });
        rf_istft_in in;
        const size_t o0 = static_cast<size_t>(t0) * tb.n_live;
        in.S0 = S + o0;
        in.cur0 = cur + o0;
        in.prev0 = prev ? prev + o0 : nullptr;
        in.S1 = has1 ? S + o0 + tb.n_live : tb.zero_row;
        in.cur1 = cur + o0 + (has1 ? tb.n_live : 0);
        in.prev1 = prev ? prev + o0 + (has1 ? tb.n_live : 0) : nullptr;
        in.mode = mode;
        in.momentum = momentum;
        phase([&](int tid) { rf_istft_load<NA>(tid, NT, V.data(), tb, j0, j1, in); });
        phase([&](int tid) { rf_pass_c7<true, NA, 0>(tid, NT, V.data()); });
        phase([&](int tid) { rf_pass_c7<true, NA, 1>(tid, NT, V.data()); });
        phase([&](int tid) { rf_pass_a<true, NA>(tid, NT, V.data()); });
        // device: one call with which=2 (barrier between the real- and imaginary-part adds)
        float* o = ola.data() + pr * pair_stride;
        phase([&](int tid) { rf_istft_pass_b<NA>(tid, NT, V.data(), o, tb, g, has1, 0); });
        phase([&](int tid) { rf_istft_pass_b<NA>(tid, NT, V.data(), o, tb, g, has1, 1); });
    }
    std::memcpy(dst, ola.data(), PL * sizeof(float));
}

std::vector<float> window_sq(const rf_plan_host& h) {
    std::vector<float> win2(h.W);
    for (int i = 0; i < h.W; ++i) win2[i] = h.window[i] * h.window[i];
    return win2;
}

// k_istft_chunk + k_ola_assemble (NA = 10 -> x[L]) or the half-rate part of k_istft_dec + k_ola_assemble_dec
// (NA = 5 -> x[(L-1)/2] odd samples) for one clip
template <int NA>
void emu_istft_clip(const rf_plan_host& h, const float* S, const rf_c32* cur, const rf_c32* prev, int mode,
                    float momentum, int T, float* x) {
    const int G = RF_CHUNK;
    const int nchunks = (T + G - 1) / G;
    const int PL = NA == 10 ? (G - 1) * h.H + h.W : ((G - 1) * h.H + h.W + 1) / 2;
    std::vector<float> part(static_cast<size_t>(2) * nchunks * PL, 0.f);
    for (int chunk = 0; chunk < nchunks; ++chunk)
        for (int g = 0; g < 2; ++g)
            emu_istft_chunk<NA>(h, S, cur, prev, mode, momentum, T, PL, g, chunk,
                                &part[(static_cast<size_t>(g) * nchunks + chunk) * PL]);
    const std::vector<float> win2 = window_sq(h);
    const int L = h.H * (T - 1);
    if (NA == 10) {
        for (int i = 0; i < L; ++i)
            x[i] = rf_ola_sample(i, part.data(), rf_envelope(i, win2.data(), T, h.H, h.W), T, G, PL, nchunks, h.H,
                                 h.W);
    } else {
        for (int v = 0; v < L / 2; ++v)      // every odd sample index 2v+1 < L (the device's xo holds (L-1)/2 of them; the
                                             // last one of an even L only exists in the tail strip)
            x[v] = rf_ola_sample_d2(v, part.data(), rf_envelope(2 * v + 1, win2.data(), T, h.H, h.W), G, PL, nchunks,
                                    h.H, h.W);
    }
}

// the edge strips of the hybrid loop as k_istft_half + k_ola_assemble_dec produce them: xe[2E] = head strip | tail strip.
// Odd sample indices i are ordinary half-rate samples (xo[(i-1)/2], from emu_istft_clip<5>), even ones come from the edge
// chunks evaluated on the other sample parity.
void emu_istft_edges(const rf_plan_host& h, const float* S, const rf_c32* cur, const rf_c32* prev, int mode,
                     float momentum, int T, const float* xo, float* xe) {
    const int G = RF_CHUNK;
    const rf_gl_dec_geom d = rf_dec_geom(T, G, h.H, h.W);
    const int PLh = ((G - 1) * h.H + h.W + 1) / 2;
    std::vector<float> part(static_cast<size_t>(2) * d.nslots * PLh, 0.f);
    for (int slot = 0; slot < d.nslots; ++slot)
        for (int g = 0; g < 2; ++g)
            emu_istft_chunk<5>(h, S, cur, prev, mode, momentum, T, PLh, g, slot == 0 ? 0 : d.c_tail + slot - 1,
                               &part[(static_cast<size_t>(g) * d.nslots + slot) * PLh], true);
    const std::vector<float> win2 = window_sq(h);
    const int L = h.H * (T - 1);
    for (int e = 0; e < 2 * d.E; ++e) {
        const int i = e < d.E ? e : L - 2 * d.E + e;
        if (i & 1)
            xe[e] = xo[(i - 1) >> 1];
        else
            xe[e] = rf_ola_sample_d2_slots(h.W / 2 + i, part.data(), rf_envelope(i, win2.data(), T, h.H, h.W), T, G, PLh,
                                           d.c_tail, d.nslots, h.H, h.W);
    }
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
    emu_stft_clip<10>(*h, x, L, T, R.data());
    std::memset(spec, 0, static_cast<size_t>(h->F) * T * 8);
    for (int t = 0; t < T; ++t)
        for (int j = 0; j < h->n_live; ++j) {
            const size_t o = (static_cast<size_t>(h->bins[j]) * T + t) * 2;
            spec[o] = R[static_cast<size_t>(t) * h->n_live + j].x;
            spec[o + 1] = R[static_cast<size_t>(t) * h->n_live + j].y;
        }
}

// debug/validation: half-rate STFT of the odd samples of x
void emu_stft_d2(void* p, const float* x, int L, float* spec) {
    auto* h = static_cast<rf_plan_host*>(p);
    const int T = 1 + L / h->H;
    std::vector<rf_c32> R(static_cast<size_t>(T) * h->n_live);
    std::vector<float> xo((L - 1) / 2 + 1);
    for (int v = 0; v < (L - 1) / 2; ++v) xo[v] = x[2 * v + 1];
    emu_stft_clip<5>(*h, xo.data(), L, T, R.data());
    std::memset(spec, 0, static_cast<size_t>(h->F) * T * 8);
    for (int t = 0; t < T; ++t)
        for (int j = 0; j < h->n_live; ++j) {
            const size_t o = (static_cast<size_t>(h->bins[j]) * T + t) * 2;
            spec[o] = R[static_cast<size_t>(t) * h->n_live + j].x;
            spec[o + 1] = R[static_cast<size_t>(t) * h->n_live + j].y;
        }
}

// debug/validation: one inverse STFT (mode 0: spec = S * cur) at full (NA=10 -> wave[L]) or half rate (-> wave[(L-1)/2])
// debug/validation: one inverse STFT (mode 0) of EVERY chunk on the other sample parity of the half-rate grid (tables t5e):
// even[v] = x[2v], v < (L + 1) / 2 — the samples the regular half-rate pass skips
void emu_istft_other_parity(void* p, const float* lin, const float* angles, int T, float* even) {
    auto* h = static_cast<rf_plan_host*>(p);
    const size_t n = static_cast<size_t>(T) * h->n_live;
    std::vector<float> S(n);
    std::vector<rf_c32> R1(n);
    for (int t = 0; t < T; ++t)
        for (int j = 0; j < h->n_live; ++j) {
            const size_t src = static_cast<size_t>(h->bins[j]) * T + t;
            S[static_cast<size_t>(t) * h->n_live + j] = lin[src];
            R1[static_cast<size_t>(t) * h->n_live + j] = c_make(angles[2 * src], angles[2 * src + 1]);
        }
    const int G = RF_CHUNK, nchunks = (T + G - 1) / G;
    const int PLh = ((G - 1) * h->H + h->W + 1) / 2;
    std::vector<float> part(static_cast<size_t>(2) * nchunks * PLh, 0.f);
    for (int c = 0; c < nchunks; ++c)
        for (int g = 0; g < 2; ++g)
            emu_istft_chunk<5>(*h, S.data(), R1.data(), nullptr, 0, 0.f, T, PLh, g, c,
                               &part[(static_cast<size_t>(g) * nchunks + c) * PLh], true);
    const std::vector<float> win2 = window_sq(*h);
    const int L = h->H * (T - 1);
    for (int i = 0; i < L; i += 2)      // slot s = chunk s: c_tail = 1 makes rf_ola_sample_d2_slots' slot map the identity
        even[i >> 1] = rf_ola_sample_d2_slots(h->W / 2 + i, part.data(), rf_envelope(i, win2.data(), T, h->H, h->W), T, G, PLh,
                                              1, nchunks, h->H, h->W);
}

void emu_istft(void* p, const float* lin, const float* angles, int T, int half, float* wave) {
    auto* h = static_cast<rf_plan_host*>(p);
    const size_t n = static_cast<size_t>(T) * h->n_live;
    std::vector<float> S(n);
    std::vector<rf_c32> R1(n);
    for (int t = 0; t < T; ++t)
        for (int j = 0; j < h->n_live; ++j) {
            const size_t src = static_cast<size_t>(h->bins[j]) * T + t;
            S[static_cast<size_t>(t) * h->n_live + j] = lin[src];
            R1[static_cast<size_t>(t) * h->n_live + j] = c_make(angles[2 * src], angles[2 * src + 1]);
        }
    if (half) {
        const int L = h->H * (T - 1);
        std::vector<float> xo(static_cast<size_t>(L) / 2 + 1);
        emu_istft_clip<5>(*h, S.data(), R1.data(), nullptr, 0, 0.f, T, xo.data());
        std::memcpy(wave, xo.data(), static_cast<size_t>((L - 1) / 2) * sizeof(float));
    }
    else emu_istft_clip<10>(*h, S.data(), R1.data(), nullptr, 0, 0.f, T, wave);
}

int emu_plan_decimate(void* p) { return static_cast<rf_plan_host*>(p)->decimate ? 1 : 0; }

// Griffin-Lim of one clip, same buffer rotation and decimation rule as gl_loop() in rf_audio.cu.
// lin[F][T], angles[F][T] complex64 (or null), wave[hop*(T-1)]; decimate != 0 runs the half-rate inner loop
void emu_griffinlim2(void* p, const float* lin, const float* angles, int T, int n_iter, float momentum_in,
                     int decimate, float* wave) {
    auto* h = static_cast<rf_plan_host*>(p);
    const bool dec = decimate && h->decimate && rf_dec_ok(T, RF_CHUNK);
    const rf_gl_dec_geom d = rf_dec_geom(T, RF_CHUNK, h->H, h->W);
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
    std::vector<float> xo(static_cast<size_t>(L) / 2 + 1), xe(2 * (h->W + h->H));
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
        const bool last = it == n_iter;
        if (dec && !last) {
            emu_istft_clip<5>(*h, S.data(), cur, prev, mode, m, T, xo.data());
            emu_istft_edges(*h, S.data(), cur, prev, mode, m, T, xo.data(), xe.data());
        } else {
            emu_istft_clip<10>(*h, S.data(), cur, prev, mode, m, T, wave);
        }
        if (last) break;
        if (dec) {   // k_stft_dec
            emu_stft_pairs<10>(*h, xe.data(), 0, L, T, 0, 3, R[it & 1]);
            emu_stft_pairs<10>(*h, xe.data() + d.E, L - d.E, L, T, d.pr_tail, (T + 1) / 2, R[it & 1]);
            emu_stft_pairs<5>(*h, xo.data(), 0, L, T, 3, d.pr_tail, R[it & 1]);
        } else {
            emu_stft_clip<10>(*h, wave, L, T, R[it & 1]);
        }
    }
}

void emu_griffinlim(void* p, const float* lin, const float* angles, int T, int n_iter, float momentum_in,
                    float* wave) {
    emu_griffinlim2(p, lin, angles, T, n_iter, momentum_in, 0, wave);
}
}
