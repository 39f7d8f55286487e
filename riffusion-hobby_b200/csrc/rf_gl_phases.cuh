// Phase functions of the STFT / iSTFT kernels.  Every phase is a function of (tid, nthreads)
// with no intra-phase cross-thread dependency, so the CUDA kernels call them with
// __syncthreads() in between and tests/hostemu runs them as plain loops over tid (a
// CPU emulation of the device control flow — test infrastructure only).
//
// Algorithm: two consecutive frames (t, t+1) are packed as the real and
// imaginary part of one complex sequence z[n'] over the 4410 live samples of the
// 17640-sample frame.  Bin k = 4m + r of the 17640-point DFT is the m-th output of a
// 4410-point DFT of z[n'] * exp(-2 pi i r n'/17640); the 4410-point DFT is a twiddle-free
// prime-factor 10 x 9 x 49 three-dimensional DFT held in shared memory as V[a][b][c]
// (position a*441 + b*49 + c).  One CTA handles one "group" (r in {0,2} or r in {1,3}) so
// that a bin and its Hermitian partner N-k live in the same CTA.
//
// Pass order and thread mapping are chosen for conflict-free shared memory:
//   forward: radix-9 over b fused with gather/window/modulate (lanes run over a: sample stride
//            441 is odd) -> radix-10 over a (lanes over c) -> radix-49 over c (lane stride 49)
//   inverse: radix-49 -> radix-10 -> radix-9 fused with demodulate/window/overlap-add.
#pragma once
#include <math.h>
#include <stddef.h>
#include <stdint.h>

#include "rf_dft.cuh"

#define RF_PW 4410
#define RF_PN 17640
#define RF_NT 256                              // threads per CTA of the FFT kernels
#define RF_B_ITEMS 490                         // (a, c) items of the fused radix-9 pass
#define RF_B_ITERS ((RF_B_ITEMS + RF_NT - 1) / RF_NT)
#define RF_LOAD_UNROLL 4

struct rf_gl_tables {
    const rf_c32* wt_fwd;  // [4][9][49][10]  w[n'] * exp(-2 pi i r n'/N), index r*4410 + b*490 + c*10 + a
    const rf_c32* wt_inv;  // same layout,    w[n']/N * exp(+2 pi i r n'/N)
    const uint32_t* pp;    // [n_live] r | idx<<2 | idx2<<15 | (k&7)<<28
    int n_live;
    int n_even;
    int hop;
};

// phase factor exp(-2 pi i * 3k/8) (frame offset (N-W)/2 = 3N/8), from k & 7, branch-free
RF_HD rf_c32 rf_phase8(int k7) {
    const float h = 0.70710678118654752440f;
    const int q = (3 * k7) & 7;  // angle = -pi q / 4
    const float mc = (q & 1) ? h : ((q & 2) ? 0.f : 1.f);
    const float ms = (q & 1) ? h : ((q & 2) ? 1.f : 0.f);
    const bool neg_c = ((q - 3) & 7) < 3;  // q in {3,4,5}
    const bool neg_s = q >= 5;             // q in {5,6,7}
    return c_make(neg_c ? -mc : mc, neg_s ? ms : -ms);
}

// ------------------------------------------------------------------ shared passes
// radix-10 over a: items (s, b, c); lanes run over consecutive positions
template <bool INV>
RF_HD void rf_pass_a(int tid, int nt, rf_c32* V) {
    for (int it = tid; it < 2 * 441; it += nt) {
        const int s = it / 441;
        rf_c32* p = V + s * RF_PW + (it - s * 441);
        rf_c32 v[10];
#pragma unroll
        for (int a = 0; a < 10; ++a) v[a] = p[441 * a];
        dft10<INV>(v);
#pragma unroll
        for (int a = 0; a < 10; ++a) p[441 * a] = v[a];
    }
}

// radix-49 over c: items (s, ab); lane stride 49 elements (odd: conflict-free for 8-byte words)
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
// xs[hop + n'].  First pass: gather, window*modulate (two r values), radix-9 over b.
RF_HD void rf_stft_pass_b(int tid, int nt, rf_c32* V, const float* xs, const rf_gl_tables& tb, int g,
                          bool has1) {
    const int r0 = g ? 1 : 0, r1 = g ? 3 : 2;
    const rf_c32* w0 = tb.wt_fwd + r0 * RF_PW;
    const rf_c32* w1 = tb.wt_fwd + r1 * RF_PW;
    for (int tau = tid; tau < RF_B_ITEMS; tau += nt) {
        const int c = tau / 10;
        const int a = tau - c * 10;
        const int base = (441 * a + 90 * c) % RF_PW;
        rf_c32 u0[9], u1[9];
#pragma unroll
        for (int b = 0; b < 9; ++b) {
            int n = base + 490 * b;
            if (n >= RF_PW) n -= RF_PW;
            const rf_c32 z = c_make(xs[n], has1 ? xs[tb.hop + n] : 0.f);
            u0[b] = c_mul(z, w0[b * 490 + tau]);
            u1[b] = c_mul(z, w1[b * 490 + tau]);
        }
        dft9<false>(u0);
        dft9<false>(u1);
        rf_c32* p = V + a * 441 + c;
#pragma unroll
        for (int b = 0; b < 9; ++b) {
            p[49 * b] = u0[b];
            p[RF_PW + 49 * b] = u1[b];
        }
    }
}

// Unpack the pair: X_t[k] and X_{t+1}[k] for the live bins j in [j0, j1) of this group.
// out0/out1: rows of the [T][n_live] spectrum for frames t0, t0+1 (out1 may be null).
RF_HD void rf_stft_post(int tid, int nt, const rf_c32* V, const rf_gl_tables& tb, int j0, int j1,
                        rf_c32* out0, rf_c32* out1) {
    for (int jb = j0 + tid; jb < j1; jb += nt * RF_LOAD_UNROLL) {
        uint32_t pw[RF_LOAD_UNROLL];
#pragma unroll
        for (int u = 0; u < RF_LOAD_UNROLL; ++u) {
            const int j = jb + u * nt;
            pw[u] = (j < j1) ? tb.pp[j] : 0u;
        }
#pragma unroll
        for (int u = 0; u < RF_LOAD_UNROLL; ++u) {
            const int j = jb + u * nt;
            if (j >= j1) continue;
            const uint32_t p = pw[u];
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
}

// ------------------------------------------------------------------ inverse (iSTFT)
RF_HD void rf_istft_zero(int tid, int nt, rf_c32* V) {
    for (int i = tid; i < 2 * RF_PW; i += nt) V[i] = c_make(0.f, 0.f);
}

// Griffin-Lim phase update fused into the load:
//   mode 0: coefficient = S * A0            (A0 = cur: the caller's initial angles)
//   mode 1: A = R - m*Rprev ; A /= (|A| + 1e-16) ; coefficient = S * A
//           (TA/functional/functional.py:337-340; no momentum term on the first update)
RF_HD rf_c32 rf_gl_coef(int mode, bool use_prev, float S, rf_c32 a, rf_c32 q, float momentum) {
    if (mode) {
        if (use_prev) {
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

// All global loads of a batch of RF_LOAD_UNROLL bins are issued before any is used, so one
// DRAM latency is paid per batch instead of per bin.
RF_HD void rf_istft_load(int tid, int nt, rf_c32* V, const rf_gl_tables& tb, int j0, int j1,
                         const rf_istft_in& in) {
    const bool f1 = in.S1 != nullptr;
    const bool up = in.prev0 != nullptr;
    for (int jb = j0 + tid; jb < j1; jb += nt * RF_LOAD_UNROLL) {
        uint32_t pw[RF_LOAD_UNROLL];
        float s0[RF_LOAD_UNROLL], s1[RF_LOAD_UNROLL];
        rf_c32 a0[RF_LOAD_UNROLL], a1[RF_LOAD_UNROLL], q0[RF_LOAD_UNROLL], q1[RF_LOAD_UNROLL];
#pragma unroll
        for (int u = 0; u < RF_LOAD_UNROLL; ++u) {
            const int j = jb + u * nt;
            const bool ok = j < j1;
            pw[u] = ok ? tb.pp[j] : 0u;
            s0[u] = ok ? in.S0[j] : 0.f;
            a0[u] = ok ? in.cur0[j] : c_make(0.f, 0.f);
            q0[u] = (ok && up) ? in.prev0[j] : c_make(0.f, 0.f);
            s1[u] = (ok && f1) ? in.S1[j] : 0.f;
            a1[u] = (ok && f1) ? in.cur1[j] : c_make(0.f, 0.f);
            q1[u] = (ok && f1 && up) ? in.prev1[j] : c_make(0.f, 0.f);
        }
#pragma unroll
        for (int u = 0; u < RF_LOAD_UNROLL; ++u) {
            const int j = jb + u * nt;
            if (j >= j1) continue;
            const uint32_t p = pw[u];
            const int r = p & 3, idx = (p >> 2) & 8191, idx2 = (p >> 15) & 8191, k7 = p >> 28;
            const int rp = (4 - r) & 3;
            const int s = r >> 1, s2 = rp >> 1;
            rf_c32 c0 = rf_gl_coef(in.mode, up, s0[u], a0[u], q0[u], in.momentum);
            rf_c32 c1 = c_make(0.f, 0.f);
            if (f1) c1 = rf_gl_coef(in.mode, up, s1[u], a1[u], q1[u], in.momentum);
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
}

// Last inverse pass: radix-9 over b for both sub-FFTs, demodulate, window (x 1/N), overlap-add.
// ola points at the chunk accumulator position of frame t0's first live sample; frame t0+1
// lands hop samples later.  The add of frame t0+1 at n'+hop aliases the add of frame t0 at
// n'' = n'+hop done by ANOTHER thread, so the real-part adds and the imaginary-part adds are
// separated by a barrier:
//   which = 0: real-part adds only, 1: imaginary-part adds only (the host emulation runs the
//   function twice), 2: both with __syncthreads() in between (device).
// Every thread executes the same number of iterations so the barrier is convergent.
RF_HD void rf_istft_pass_b(int tid, int nt, const rf_c32* V, float* ola, const rf_gl_tables& tb, int g,
                           bool has1, int which) {
    const int r0 = g ? 1 : 0, r1 = g ? 3 : 2;
    const rf_c32* w0 = tb.wt_inv + r0 * RF_PW;
    const rf_c32* w1 = tb.wt_inv + r1 * RF_PW;
    float re[RF_B_ITERS][9], im[RF_B_ITERS][9];
    int base_n[RF_B_ITERS];
#pragma unroll
    for (int itn = 0; itn < RF_B_ITERS; ++itn) {
        const int tau = tid + itn * nt;
        if (tau < RF_B_ITEMS) {
            const int c = tau / 10;
            const int a = tau - c * 10;
            base_n[itn] = (441 * a + 90 * c) % RF_PW;
            const rf_c32* p = V + a * 441 + c;
            rf_c32 u0[9], u1[9];
#pragma unroll
            for (int b = 0; b < 9; ++b) {
                u0[b] = p[49 * b];
                u1[b] = p[RF_PW + 49 * b];
            }
            dft9<true>(u0);
            dft9<true>(u1);
#pragma unroll
            for (int b = 0; b < 9; ++b) {
                const rf_c32 z = c_add(c_mul(u0[b], w0[b * 490 + tau]), c_mul(u1[b], w1[b * 490 + tau]));
                re[itn][b] = z.x;
                im[itn][b] = z.y;
            }
        }
    }
    if (which != 1) {
#pragma unroll
        for (int itn = 0; itn < RF_B_ITERS; ++itn)
            if (tid + itn * nt < RF_B_ITEMS) {
#pragma unroll
                for (int b = 0; b < 9; ++b) {
                    int n = base_n[itn] + 490 * b;
                    if (n >= RF_PW) n -= RF_PW;
                    ola[n] += re[itn][b];
                }
            }
    }
#if defined(__CUDA_ARCH__)
    if (which == 2) __syncthreads();
#endif
    if (which != 0 && has1) {
#pragma unroll
        for (int itn = 0; itn < RF_B_ITERS; ++itn)
            if (tid + itn * nt < RF_B_ITEMS) {
#pragma unroll
                for (int b = 0; b < 9; ++b) {
                    int n = base_n[itn] + 490 * b;
                    if (n >= RF_PW) n -= RF_PW;
                    ola[tb.hop + n] += im[itn][b];
                }
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
RF_HD void rf_stage_x(int tid, int nt, float* xs, const float* x, int L, int t0, int hop) {
    const int q0 = t0 * hop - RF_PW / 2;
    for (int i = tid; i < RF_PW + hop; i += nt) {
        const int ii = rf_reflect_index(q0 + i, L);
        xs[i] = (ii >= 0 && ii < L) ? x[ii] : 0.f;
    }
}

// window envelope of torch.istft at sample i of the kept region: sum of w^2 of covering frames
RF_HD float rf_envelope(int i, const float* win2, int T, int H, int W) {
    const int q = W / 2 + i;  // hop coordinates: frame t covers [tH, tH+W)
    const int t_hi = (T - 1 < q / H) ? T - 1 : q / H;
    int t_lo = (q - W + H) / H;  // ceil((q-W+1)/H)
    if (q - W + 1 <= 0) t_lo = 0;
    float env = 0.f;
    for (int t = t_lo; t <= t_hi; ++t) env += win2[q - t * H];
    return env;
}

// One output sample of torch.istft's overlap-add: sum of the chunk partial sums that cover
// sample i of the kept region, divided by the window envelope.  part: [2 groups][nchunks][PL]
RF_HD float rf_ola_sample(int i, const float* part, float env, int T, int G, int PL, int nchunks,
                          int H, int W) {
    const int q = W / 2 + i;
    const int t_hi = (T - 1 < q / H) ? T - 1 : q / H;
    int t_lo = (q - W + H) / H;
    if (q - W + 1 <= 0) t_lo = 0;
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
