// Phase functions of the STFT / iSTFT kernels.  Every phase is a function of (tid, nthreads)
// with no intra-phase cross-thread dependency, so the CUDA kernels call them with
// __syncthreads() in between and tests/hostemu runs them as plain loops over tid (a
// bit-for-bit CPU emulation of the device control flow — test infrastructure only).
//
// Algorithm: two consecutive frames (t, t+1) are packed as the real and
// imaginary part of one complex sequence z[n'] over the 4410 live samples of the
// 17640-sample frame.  Bin k = 4m + r of the 17640-point DFT is the m-th output of a
// 4410-point DFT of z[n'] * exp(-2 pi i r n'/17640); the 4410-point DFT is a twiddle-free
// prime-factor 10 x 9 x 49 three-dimensional DFT held in shared memory as V[a][b][c].
// One CTA handles one "group" (r in {0,2} or r in {1,3}) so that a bin and its Hermitian
// partner N-k live in the same CTA.
#pragma once
#include <math.h>
#include <stddef.h>
#include <stdint.h>

#include "rf_dft.cuh"

#define RF_PW 4410
#define RF_PN 17640

struct rf_gl_tables {
    const rf_c32* wt_fwd;  // [4][4410]
    const rf_c32* wt_inv;  // [4][4410]
    const uint32_t* pp;    // [n_live]
    int n_live;
    int n_even;
    int hop;
};

// phase factor exp(-2 pi i * 3k/8) (frame offset (N-W)/2 = 3N/8), indexed by k & 7
RF_HD rf_c32 rf_phase8(int k7) {
    const float h = 0.70710678118654752440f;
    switch (k7) {
        case 0: return c_make(1.f, 0.f);
        case 1: return c_make(-h, -h);
        case 2: return c_make(0.f, 1.f);
        case 3: return c_make(h, -h);
        case 4: return c_make(-1.f, 0.f);
        case 5: return c_make(h, h);
        case 6: return c_make(0.f, -1.f);
        default: return c_make(-h, h);
    }
}

// ------------------------------------------------------------------ shared passes
// radix-9 over b: items (s, a, c), s = sub-FFT 0/1
template <bool INV>
RF_HD void rf_pass_b(int tid, int nt, rf_c32* V) {
    for (int it = tid; it < 2 * 490; it += nt) {
        const int s = it / 490;
        const int rem = it - s * 490;
        const int a = rem / 49;
        const int c = rem - a * 49;
        rf_c32* p = V + s * RF_PW + a * 441 + c;
        rf_c32 v[9];
#pragma unroll
        for (int b = 0; b < 9; ++b) v[b] = p[49 * b];
        dft9<INV>(v);
#pragma unroll
        for (int b = 0; b < 9; ++b) p[49 * b] = v[b];
    }
}

// radix-49 over c: items (s, ab)
template <bool INV>
RF_HD void rf_pass_c(int tid, int nt, rf_c32* V) {
    for (int it = tid; it < 2 * 90; it += nt) {
        rf_c32* p = V + it * 49;  // s*4410 + ab*49 == it*49
        rf_c32 v[49];
#pragma unroll
        for (int c = 0; c < 49; ++c) v[c] = p[c];
        dft49<INV>(v);
#pragma unroll
        for (int c = 0; c < 49; ++c) p[c] = v[7 * (c % 7) + c / 7];
    }
}

// ------------------------------------------------------------------ forward (STFT)
// xs[0 .. W+hop): padded signal starting at the first live sample of frame t0; frame t0+1 is
// xs[hop + n'] (zero when the pair has no second frame).
RF_HD void rf_stft_pass_a(int tid, int nt, rf_c32* V, const float* xs, const rf_gl_tables& tb,
                          int g) {
    const int r0 = g ? 1 : 0, r1 = g ? 3 : 2;
    const rf_c32* w0 = tb.wt_fwd + r0 * RF_PW;
    const rf_c32* w1 = tb.wt_fwd + r1 * RF_PW;
    for (int tau = tid; tau < 441; tau += nt) {
        const int b = tau / 49;
        const int c = tau - b * 49;
        const int base = (490 * b + 90 * c) % RF_PW;
        rf_c32 u0[10], u1[10];
#pragma unroll
        for (int a = 0; a < 10; ++a) {
            int n = base + 441 * a;
            if (n >= RF_PW) n -= RF_PW;
            const rf_c32 z = c_make(xs[n], xs[tb.hop + n]);
            u0[a] = c_mul(z, w0[a * 441 + tau]);
            u1[a] = c_mul(z, w1[a * 441 + tau]);
        }
        dft10<false>(u0);
        dft10<false>(u1);
#pragma unroll
        for (int a = 0; a < 10; ++a) {
            V[a * 441 + tau] = u0[a];
            V[RF_PW + a * 441 + tau] = u1[a];
        }
    }
}

// Unpack the pair: X_t[k] and X_{t+1}[k] for the live bins j in [j0, j1) of this group.
// out0/out1: rows of the [T][n_live] spectrum for frames t0, t0+1 (out1 may be null).
RF_HD void rf_stft_post(int tid, int nt, const rf_c32* V, const rf_gl_tables& tb, int j0, int j1,
                        rf_c32* out0, rf_c32* out1) {
    for (int j = j0 + tid; j < j1; j += nt) {
        const uint32_t p = tb.pp[j];
        const int r = p & 3, idx = (p >> 2) & 8191, idx2 = (p >> 15) & 8191, k7 = p >> 28;
        const int s = r >> 1, s2 = ((4 - r) & 3) >> 1;
        const rf_c32 zk = V[s * RF_PW + idx];
        const rf_c32 zp = V[s2 * RF_PW + idx2];
        const rf_c32 ph = rf_phase8(k7);
        const rf_c32 g0 = c_make(0.5f * (zk.x + zp.x), 0.5f * (zk.y - zp.y));
        const rf_c32 g1 = c_make(0.5f * (zk.y + zp.y), -0.5f * (zk.x - zp.x));
        out0[j] = c_mul(ph, g0);
        if (out1) out1[j] = c_mul(ph, g1);
    }
}

// ------------------------------------------------------------------ inverse (iSTFT)
RF_HD void rf_istft_zero(int tid, int nt, rf_c32* V) {
    for (int i = tid; i < 2 * RF_PW; i += nt) V[i] = c_make(0.f, 0.f);
}

// Griffin-Lim phase update fused into the load:
//   mode 0: coefficient = S * A0            (A0 in cur: the caller's initial angles)
//   mode 1: A = R - m*Rprev ; A /= (|A| + 1e-16) ; coefficient = S * A
//           (TA/functional/functional.py:337-340; prev may be null: tprev = 0 on iteration 1)
RF_HD rf_c32 rf_gl_coef(int mode, float S, const rf_c32* cur, const rf_c32* prev, float momentum,
                        int j) {
    rf_c32 a = cur[j];
    if (mode) {
        if (prev) {
            const rf_c32 q = prev[j];
#if defined(__CUDA_ARCH__)
            a.x = __fsub_rn(a.x, __fmul_rn(q.x, momentum));
            a.y = __fsub_rn(a.y, __fmul_rn(q.y, momentum));
#else
            volatile float mx = q.x * momentum, my = q.y * momentum;
            a.x = a.x - mx;
            a.y = a.y - my;
#endif
        }
#if defined(__CUDA_ARCH__)
        const float d = __fadd_rn(__fsqrt_rn(__fadd_rn(__fmul_rn(a.x, a.x), __fmul_rn(a.y, a.y))),
                                  1e-16f);
        a.x = __fdiv_rn(a.x, d);
        a.y = __fdiv_rn(a.y, d);
#else
        volatile float xx = a.x * a.x, yy = a.y * a.y;
        const float d = sqrtf(xx + yy) + 1e-16f;
        a.x = a.x / d;
        a.y = a.y / d;
#endif
    }
    return c_make(S * a.x, S * a.y);
}

struct rf_istft_in {
    const float* S0;      // magnitudes, frame t0   [n_live]
    const float* S1;      // frame t0+1 or null
    const rf_c32* cur0;   // R (or A0) rows
    const rf_c32* cur1;
    const rf_c32* prev0;  // previous R rows or null
    const rf_c32* prev1;
    int mode;
    float momentum;
};

RF_HD void rf_istft_load(int tid, int nt, rf_c32* V, const rf_gl_tables& tb, int j0, int j1,
                         const rf_istft_in& in) {
    for (int j = j0 + tid; j < j1; j += nt) {
        const uint32_t p = tb.pp[j];
        const int r = p & 3, idx = (p >> 2) & 8191, idx2 = (p >> 15) & 8191, k7 = p >> 28;
        const int rp = (4 - r) & 3;
        const int s = r >> 1, s2 = rp >> 1;
        rf_c32 c0 = rf_gl_coef(in.mode, in.S0[j], in.cur0, in.prev0, in.momentum, j);
        rf_c32 c1 = c_make(0.f, 0.f);
        if (in.S1) c1 = rf_gl_coef(in.mode, in.S1[j], in.cur1, in.prev1, in.momentum, j);
        const bool selfp = (idx2 == idx) && (rp == r);
        if (selfp) {  // DC / Nyquist: irfft ignores the imaginary part
            c0.y = 0.f;
            c1.y = 0.f;
        }
        const rf_c32 ph = rf_phase8(k7);
        // Z[k] = conj(ph) * (C0 + i C1)
        V[s * RF_PW + idx] = c_mul(c_conj(ph), c_make(c0.x - c1.y, c0.y + c1.x));
        // Z[N-k] = ph * (conj(C0) + i conj(C1))
        if (!selfp) V[s2 * RF_PW + idx2] = c_mul(ph, c_make(c0.x + c1.y, c1.x - c0.y));
    }
}

// last inverse pass: radix-10 over a for both sub-FFTs, demodulate, window, overlap-add.
// ola points at the chunk accumulator position of frame t0's first live sample; frame t0+1
// lands hop samples later.  The two read-modify-writes that can alias (n'+hop of frame t0+1
// vs n' of frame t0) are always issued by the same thread (same (b,c), a+1), so no barrier
// is needed between them.
RF_HD void rf_istft_pass_a(int tid, int nt, const rf_c32* V, float* ola, const rf_gl_tables& tb,
                           int g, bool has1) {
    const int r0 = g ? 1 : 0, r1 = g ? 3 : 2;
    const rf_c32* w0 = tb.wt_inv + r0 * RF_PW;
    const rf_c32* w1 = tb.wt_inv + r1 * RF_PW;
    for (int tau = tid; tau < 441; tau += nt) {
        const int b = tau / 49;
        const int c = tau - b * 49;
        const int base = (490 * b + 90 * c) % RF_PW;
        rf_c32 u0[10], u1[10];
#pragma unroll
        for (int a = 0; a < 10; ++a) {
            u0[a] = V[a * 441 + tau];
            u1[a] = V[RF_PW + a * 441 + tau];
        }
        dft10<true>(u0);
        dft10<true>(u1);
        float re[10], im[10];
        int nn[10];
#pragma unroll
        for (int a = 0; a < 10; ++a) {
            int n = base + 441 * a;
            if (n >= RF_PW) n -= RF_PW;
            nn[a] = n;
            const rf_c32 z = c_add(c_mul(u0[a], w0[a * 441 + tau]), c_mul(u1[a], w1[a * 441 + tau]));
            re[a] = z.x;
            im[a] = z.y;
        }
#pragma unroll
        for (int a = 0; a < 10; ++a) ola[nn[a]] += re[a];
        if (has1) {
#pragma unroll
            for (int a = 0; a < 10; ++a) ola[tb.hop + nn[a]] += im[a];
        }
    }
}

// ------------------------------------------------------------------ staging / assembly
RF_HD int rf_reflect_index(int i, int L) {
    if (i < 0) i = -i;
    else if (i >= L) i = 2 * (L - 1) - i;
    return i;
}

// xs[i] = x_padded[t0*hop + i] for i in [0, W+hop), with torch.stft's reflect padding
// (center=True, pad_mode="reflect") applied on the fly to the un-padded signal x[0..L).
RF_HD void rf_stage_x(int tid, int nt, float* xs, const float* x, int L, int t0, int hop, bool has1) {
    const int q0 = t0 * hop - RF_PW / 2;
    for (int i = tid; i < RF_PW + hop; i += nt) {
        const int ii = rf_reflect_index(q0 + i, L);
        xs[i] = (ii >= 0 && ii < L && (has1 || i < RF_PW)) ? x[ii] : 0.f;
    }
}

// One output sample of torch.istft's overlap-add: sum of the chunk partial sums that cover
// sample i of the kept region, divided by the window envelope (sum of w^2 of covering frames).
// part layout: [2 groups][nchunks][PL]
RF_HD float rf_ola_sample(int i, const float* part, const float* win2, int T, int G, int PL,
                          int nchunks, int H, int W) {
    const int q = W / 2 + i;  // hop coordinates: frame t covers [tH, tH+W)
    const int t_hi = (T - 1 < q / H) ? T - 1 : q / H;
    int t_lo = (q - W + H) / H;  // ceil((q-W+1)/H) for q-W+1 > -H
    if (q - W + 1 <= 0) t_lo = 0;
    float env = 0.f;
    for (int t = t_lo; t <= t_hi; ++t) env += win2[q - t * H];
    const int c_lo = t_lo / G, c_hi = t_hi / G;
    float acc = 0.f;
    for (int g = 0; g < 2; ++g)
        for (int c = c_lo; c <= c_hi; ++c) {
            const int off = q - c * G * H;
            const int nf = (G < T - c * G) ? G : T - c * G;
            if (off >= 0 && off < (nf - 1) * H + W)
                acc += part[(static_cast<size_t>(g) * nchunks + c) * PL + off];
        }
    return acc / env;
}
