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
#define RF_LOAD_UNROLL 4

// The passes are templated on NA, the size of the "a" axis of the prime-factor grid:
//   NA = 10: full-rate transform, W = 4410 samples per frame (n' = 441a + 490b + 90c mod 4410)
//   NA = 5 : time-decimated transform used inside the Griffin-Lim loop when the live band allows it:
//            only every second sample of the waveform is carried (W = 2205, n'' = 441a + 245b + 45c mod 2205),
//            the bins k = 4m + r of the 17640-grid become outputs of 2205-point DFTs (period 8820 in k).
// Shared-memory position of (a, b, c) is a*441 + b*49 + c in both cases.
template <int NA> struct rf_geom {
    static constexpr int W = NA * 441;          // complex points per sub-transform
    static constexpr int SB = W / 9;            // index step of b in the time-side (Ruritanian) map
    static constexpr int SC = W / 49;           // index step of c
    static constexpr int B_ITEMS = 49 * NA;     // (a, c) items of the fused radix-9 pass
    static constexpr int B_ITERS = (B_ITEMS + RF_NT - 1) / RF_NT;
};

struct rf_gl_tables {
    const rf_c32* wt_fwd;  // [2 parities][4][9][49][NA]  window x modulation, index (par*4 + r)*W + b*49*NA + c*NA + a
    const rf_c32* wt_inv;  // same layout (parity 1 only exists for NA = 5)
    const uint32_t* pp;    // [n_live] r | idx<<2 | idx2<<15 | (k&7)<<28  (positions for this NA)
    const rf_c32* ph_odd;  // NA = 5: exp(-2 pi i k/N) per live bin (the odd-sample frame of a pair), else null
    int n_live;
    int n_even;
    int off1;              // offset of frame t0+1 in the staged sample buffer: hop (NA = 10) or (hop+1)/2 (NA = 5)
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
template <bool INV, int NA>
RF_HD void rf_pass_a(int tid, int nt, rf_c32* V) {
    constexpr int W = rf_geom<NA>::W;
    for (int it = tid; it < 2 * 441; it += nt) {
        const int s = it / 441;
        rf_c32* p = V + s * W + (it - s * 441);
        rf_c32 v[NA];
#pragma unroll
        for (int a = 0; a < NA; ++a) v[a] = p[441 * a];
        if (NA == 10) dft10<INV>(v);
        else dft5<INV>(v[0], v[1], v[2], v[3], v[4]);
#pragma unroll
        for (int a = 0; a < NA; ++a) p[441 * a] = v[a];
    }
}

// radix-49 over c, SEVEN threads per 49-point transform (work item w = j * transforms + transform), two phases with a barrier
// in between, both in place and free of cross-thread hazards:
//   column phase (thread j = c2): DFT7 over the stride-7 elements p[7 c1 + j], twiddle w49^(k1 j), back to p[7 k1 + j]
//   row phase    (thread j = k1): DFT7 over the contiguous row p[7 j + c2], back to p[7 j + k2] = X[j + 7 k2]
// forward = column phase then row phase: natural time order in, TRANSPOSED spectral order out (X[k] at 7 (k % 7) + k / 7,
// which is what the plan's position tables pp / pp2 point at: rf_pfa_spec_pos);  inverse = row phase (twiddle after the
// DFT) then column phase: transposed spectral order in, natural time order out.  One thread per transform (the previous
// form) kept 49 complex values = 98 registers live and left 166 of 256 threads idle for NA = 5.
#ifndef RF_GL_PASS7
#define RF_GL_PASS7 1       // 0: the one-thread-per-transform radix-49 pass (A/B builds, scratch/variants)
#endif
template <bool INV, int NA, int STEP>
RF_HD void rf_pass_c7(int tid, int nt, rf_c32* V) {
#if !RF_GL_PASS7
    if (STEP == 0) {
        for (int it = tid; it < 2 * 9 * NA; it += nt) {
            rf_c32* p = V + it * 49;
            rf_c32 v[49];
#pragma unroll
            for (int c = 0; c < 49; ++c) v[c] = p[c];
            dft49<INV>(v);
#pragma unroll
            for (int c = 0; c < 49; ++c) p[c] = v[7 * (c % 7) + c / 7];
        }
    }
    return;
#endif
    const float sg = INV ? 1.0f : -1.0f;
    constexpr bool COLUMN = (STEP == 0) != INV;
    constexpr int NTR = 2 * 9 * NA;         // 49-point transforms in V
    for (int w = tid; w < NTR * 7; w += nt) {
        // lanes run over the transforms (stride 49 elements = 1 mod 16 eight-byte banks: conflict-free in both phases) and
        // j is uniform over (most of) a warp, so the twiddle reads below are constant-cache broadcasts
        const int j = w / NTR, it = w - j * NTR;
        rf_c32* p = V + it * 49;  // s*W + ab*49 == it*49
        rf_c32 v[7];
        if (COLUMN) {
#pragma unroll
            for (int i = 0; i < 7; ++i) v[i] = p[7 * i + j];
        } else {
#pragma unroll
            for (int i = 0; i < 7; ++i) v[i] = p[7 * j + i];
        }
        dft7<INV>(v[0], v[1], v[2], v[3], v[4], v[5], v[6]);
        if (STEP == 0) {   // the twiddle sits between the two DFT7 stages: after the first phase in either direction
#pragma unroll
            for (int i = 1; i < 7; ++i) v[i] = c_mulk(v[i], rf_w49_cos(i * j), sg * rf_w49_sin(i * j));
        }
        if (COLUMN) {
#pragma unroll
            for (int i = 0; i < 7; ++i) p[7 * i + j] = v[i];
        } else {
#pragma unroll
            for (int i = 0; i < 7; ++i) p[7 * j + i] = v[i];
        }
    }
}

// ------------------------------------------------------------------ forward (STFT)
// xs[0 .. W+hop): padded signal starting at the first live sample of frame t0; frame t0+1 is
// xs[hop + n'].  First pass: gather, window*modulate (two r values), radix-9 over b.
template <int NA>
RF_HD void rf_stft_pass_b(int tid, int nt, rf_c32* V, const float* xs, const rf_gl_tables& tb, int g,
                          bool has1) {
    constexpr int W = rf_geom<NA>::W, SB = rf_geom<NA>::SB, SC = rf_geom<NA>::SC;
    const int r0 = g ? 1 : 0, r1 = g ? 3 : 2;
    // frame t0 uses the parity-0 tables, frame t0+1 the parity-1 tables (identical for NA = 10)
    const rf_c32* w00 = tb.wt_fwd + r0 * W;
    const rf_c32* w01 = tb.wt_fwd + r1 * W;
    const rf_c32* w10 = tb.wt_fwd + ((NA == 5 ? 4 : 0) + r0) * W;
    const rf_c32* w11 = tb.wt_fwd + ((NA == 5 ? 4 : 0) + r1) * W;
    for (int tau = tid; tau < rf_geom<NA>::B_ITEMS; tau += nt) {
        const int c = tau / NA;
        const int a = tau - c * NA;
        const int base = (441 * a + SC * c) % W;
        rf_c32 u0[9], u1[9];
#pragma unroll
        for (int b = 0; b < 9; ++b) {
            int n = base + SB * b;
            if (n >= W) n -= W;
            const float x0 = xs[n], x1 = has1 ? xs[tb.off1 + n] : 0.f;
            const int ti = b * (49 * NA) + tau;
            if (NA == 10) {
                const rf_c32 z = c_make(x0, x1);
                u0[b] = c_mul(z, w00[ti]);
                u1[b] = c_mul(z, w01[ti]);
            } else {   // the two frames of the pair sit on different sample parities: separate window tables
                const rf_c32 f00 = w00[ti], f01 = w01[ti], f10 = w10[ti], f11 = w11[ti];
                u0[b] = c_make(x0 * f00.x - x1 * f10.y, x0 * f00.y + x1 * f10.x);   // x0*f00 + i*x1*f10
                u1[b] = c_make(x0 * f01.x - x1 * f11.y, x0 * f01.y + x1 * f11.x);
            }
        }
        dft9<false>(u0);
        dft9<false>(u1);
        rf_c32* p = V + a * 441 + c;
#pragma unroll
        for (int b = 0; b < 9; ++b) {
            p[49 * b] = u0[b];
            p[W + 49 * b] = u1[b];
        }
    }
}

// Unpack the pair: X_t[k] and X_{t+1}[k] for the live bins j in [j0, j1) of this group.
// out0/out1: rows of the [T][n_live] spectrum for frames t0, t0+1 (out1 may be null).
template <int NA>
RF_HD void rf_stft_post(int tid, int nt, const rf_c32* V, const rf_gl_tables& tb, int j0, int j1,
                        rf_c32* out0, rf_c32* out1) {
    constexpr int W = rf_geom<NA>::W;
    for (int jb = j0 + tid; jb < j1; jb += nt * RF_LOAD_UNROLL) {
        uint32_t pw[RF_LOAD_UNROLL];
        rf_c32 po[RF_LOAD_UNROLL];
#pragma unroll
        for (int u = 0; u < RF_LOAD_UNROLL; ++u) {
            const int j = jb + u * nt;
            pw[u] = (j < j1) ? tb.pp[j] : 0u;
            if (NA == 5) po[u] = (j < j1) ? tb.ph_odd[j] : c_make(1.f, 0.f);
        }
#pragma unroll
        for (int u = 0; u < RF_LOAD_UNROLL; ++u) {
            const int j = jb + u * nt;
            if (j >= j1) continue;
            const uint32_t p = pw[u];
            const int r = p & 3, idx = (p >> 2) & 8191, idx2 = (p >> 15) & 8191, k7 = p >> 28;
            const int s = r >> 1, s2 = ((4 - r) & 3) >> 1;
            const rf_c32 zk = V[s * W + idx];
            const rf_c32 zp = V[s2 * W + idx2];
            const rf_c32 ph = rf_phase8(k7);
            const rf_c32 g0 = c_make(0.5f * (zk.x + zp.x), 0.5f * (zk.y - zp.y));
            const rf_c32 g1 = c_make(0.5f * (zk.y + zp.y), -0.5f * (zk.x - zp.x));
            out0[j] = c_mul(ph, g0);
            if (out1) out1[j] = (NA == 5) ? c_mul(c_mul(ph, po[u]), g1) : c_mul(ph, g1);
        }
    }
}

// ------------------------------------------------------------------ inverse (iSTFT)
template <int NA>
RF_HD void rf_istft_zero(int tid, int nt, rf_c32* V) {
    for (int i = tid; i < 2 * rf_geom<NA>::W; i += nt) V[i] = c_make(0.f, 0.f);
}

// Griffin-Lim phase update fused into the load:
//   mode 0: coefficient = S * A0            (A0 = cur: the caller's initial angles)
//   mode 1: A = R - m*Rprev ; A /= (|A| + 1e-16) ; coefficient = S * A
//           (TA/functional/functional.py:337-340; no momentum term on the first update)
// The normalisation is one reciprocal square root (MUFU.RSQ, ~1 ulp) instead of sqrt + two IEEE divisions: torch's
// own complex abs() is a hypot with a rounding of its own, so neither form is bit-identical to the reference, and a
// 1-ulp change of a unit phasor is the same size as the rounding differences between any two FFT implementations.
// |A| < 1e-15 (a bin with no energy) is clamped instead of adding 1e-16: the product with S is noise either way.
RF_HD rf_c32 rf_gl_coef(int mode, bool use_prev, float S, rf_c32 a, rf_c32 q, float momentum) {
    if (mode) {
        if (use_prev) {
            a.x = fmaf(-momentum, q.x, a.x);
            a.y = fmaf(-momentum, q.y, a.y);
        }
        const float n2 = fmaxf(fmaf(a.x, a.x, a.y * a.y), 1e-30f);
#if defined(__CUDA_ARCH__)
        S *= rsqrtf(n2);
#else
        S *= 1.0f / sqrtf(n2);
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
template <int NA>
RF_HD void rf_istft_load(int tid, int nt, rf_c32* V, const rf_gl_tables& tb, int j0, int j1,
                         const rf_istft_in& in) {
    constexpr int W = rf_geom<NA>::W;
    const bool f1 = in.S1 != nullptr;
    const bool up = in.prev0 != nullptr;
    for (int jb = j0 + tid; jb < j1; jb += nt * RF_LOAD_UNROLL) {
        uint32_t pw[RF_LOAD_UNROLL];
        float s0[RF_LOAD_UNROLL], s1[RF_LOAD_UNROLL];
        rf_c32 a0[RF_LOAD_UNROLL], a1[RF_LOAD_UNROLL], q0[RF_LOAD_UNROLL], q1[RF_LOAD_UNROLL], po[RF_LOAD_UNROLL];
#pragma unroll
        for (int u = 0; u < RF_LOAD_UNROLL; ++u) {
            const int j = jb + u * nt;
            const bool ok = j < j1;
            pw[u] = ok ? tb.pp[j] : 0u;
            if (NA == 5) po[u] = (ok && f1) ? tb.ph_odd[j] : c_make(1.f, 0.f);
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
            if (NA == 5) c1 = c_mul(c1, c_conj(po[u]));   // odd-sample frame: exp(+2 pi i k/N)
            const bool selfp = (idx2 == idx) && (rp == r);
            if (selfp) {  // DC / Nyquist: irfft ignores the imaginary part
                c0.y = 0.f;
                c1.y = 0.f;
            }
            const rf_c32 ph = rf_phase8(k7);
            // Z[k] = conj(ph) * (C0 + i C1)
            V[s * W + idx] = c_mul(c_conj(ph), c_make(c0.x - c1.y, c0.y + c1.x));
            // Z[N-k] = ph * (conj(C0) + i conj(C1))
            if (!selfp) V[s2 * W + idx2] = c_mul(ph, c_make(c0.x + c1.y, c1.x - c0.y));
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
template <int NA>
RF_HD void rf_istft_pass_b(int tid, int nt, const rf_c32* V, float* ola, const rf_gl_tables& tb, int g,
                           bool has1, int which) {
    constexpr int W = rf_geom<NA>::W, SB = rf_geom<NA>::SB, SC = rf_geom<NA>::SC;
    constexpr int ITERS = rf_geom<NA>::B_ITERS;
    const int r0 = g ? 1 : 0, r1 = g ? 3 : 2;
    const rf_c32* w00 = tb.wt_inv + r0 * W;
    const rf_c32* w01 = tb.wt_inv + r1 * W;
    const rf_c32* w10 = tb.wt_inv + ((NA == 5 ? 4 : 0) + r0) * W;
    const rf_c32* w11 = tb.wt_inv + ((NA == 5 ? 4 : 0) + r1) * W;
    float re[ITERS][9], im[ITERS][9];
    int base_n[ITERS];
#pragma unroll
    for (int itn = 0; itn < ITERS; ++itn) {
        const int tau = tid + itn * nt;
        if (tau < rf_geom<NA>::B_ITEMS) {
            const int c = tau / NA;
            const int a = tau - c * NA;
            base_n[itn] = (441 * a + SC * c) % W;
            const rf_c32* p = V + a * 441 + c;
            rf_c32 u0[9], u1[9];
#pragma unroll
            for (int b = 0; b < 9; ++b) {
                u0[b] = p[49 * b];
                u1[b] = p[W + 49 * b];
            }
            dft9<true>(u0);
            dft9<true>(u1);
#pragma unroll
            for (int b = 0; b < 9; ++b) {
                const int ti = b * (49 * NA) + tau;
                if (NA == 10) {
                    const rf_c32 z = c_add(c_mul(u0[b], w00[ti]), c_mul(u1[b], w01[ti]));
                    re[itn][b] = z.x;
                    im[itn][b] = z.y;
                } else {   // frame t0 = Re(z with parity-0 tables), frame t0+1 = Im(z with parity-1 tables)
                    const rf_c32 z0 = c_add(c_mul(u0[b], w00[ti]), c_mul(u1[b], w01[ti]));
                    const rf_c32 z1 = c_add(c_mul(u0[b], w10[ti]), c_mul(u1[b], w11[ti]));
                    re[itn][b] = z0.x;
                    im[itn][b] = z1.y;
                }
            }
        }
    }
    if (which != 1) {
#pragma unroll
        for (int itn = 0; itn < ITERS; ++itn)
            if (tid + itn * nt < rf_geom<NA>::B_ITEMS) {
#pragma unroll
                for (int b = 0; b < 9; ++b) {
                    int n = base_n[itn] + SB * b;
                    if (n >= W) n -= W;
                    ola[n] += re[itn][b];
                }
            }
    }
#if defined(__CUDA_ARCH__)
    if (which == 2) __syncthreads();
#endif
    if (which != 0 && has1) {
#pragma unroll
        for (int itn = 0; itn < ITERS; ++itn)
            if (tid + itn * nt < rf_geom<NA>::B_ITEMS) {
#pragma unroll
                for (int b = 0; b < 9; ++b) {
                    int n = base_n[itn] + SB * b;
                    if (n >= W) n -= W;
                    ola[tb.off1 + n] += im[itn][b];
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
// x points at waveform sample `base` (base = 0 for a whole waveform; the decimated loop keeps only the two full-rate
// edge strips [0, E) and [L-E, L), see rf_gl_dec_geom)
RF_HD void rf_stage_x(int tid, int nt, float* xs, const float* x, int L, int t0, int hop, int base = 0) {
    const int q0 = t0 * hop - RF_PW / 2;
    for (int i = tid; i < RF_PW + hop; i += nt) {
        const int ii = rf_reflect_index(q0 + i, L);
        xs[i] = (ii >= 0 && ii < L) ? x[ii - base] : 0.f;
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

// ------------------------------------------------------------------ time-decimated variants (NA = 5)
// The decimated loop carries the waveform only at odd sample indices i = 2v+1 (q = W/2 + i even in hop coordinates),
// stored as xo[v], v in [0, (L-1)/2).  Frame pairs start at even t0, so frame t0 uses the even live samples n' = 2u
// and frame t0+1 the odd ones n' = 2u+1.

// xs[v] = x_padded(q = t0*hop + 2v) for v in [0, W/2 + (hop+1)/2), reflect padding applied on the fly
RF_HD void rf_stage_x_d2(int tid, int nt, float* xs, const float* xo, int L, int t0, int hop) {
    const int nv = RF_PW / 2 + (hop + 1) / 2;
    const int nxo = (L - 1) / 2;
    for (int v = tid; v < nv; v += nt) {
        const int i = rf_reflect_index(t0 * hop + 2 * v - RF_PW / 2, L);   // odd, reflection keeps parity
        const int vo = (i - 1) >> 1;
        xs[v] = (i >= 1 && vo < nxo) ? xo[vo] : 0.f;
    }
}

// decimated overlap-add assembly: waveform sample i = 2v+1 from the half-rate chunk partial sums
// part: [2 groups][nchunks][PLh], chunk c starts at q = c*G*H (even), PLh = ((G-1)*H + W + 1) / 2
RF_HD float rf_ola_sample_d2(int v, const float* part, float env, int G, int PLh, int nchunks, int H, int W) {
    const int q = W / 2 + 2 * v + 1;
    const int cs = G * H;                        // chunk stride in q (even)
    int c_lo = (q - ((G - 1) * H + W) + cs) / cs;  // ceil((q - extent + 1) / cs)
    if (q - ((G - 1) * H + W) + 1 <= 0) c_lo = 0;
    int c_hi = q / cs;
    if (c_hi > nchunks - 1) c_hi = nchunks - 1;
    float acc = 0.f;
    for (int g = 0; g < 2; ++g)
        for (int c = c_lo; c <= c_hi; ++c) {
            const int off = (q - c * cs) >> 1;
            if (off >= 0 && off < PLh) acc += part[(static_cast<size_t>(g) * nchunks + c) * PLh + off];
        }
    return acc / env;
}

// ------------------------------------------------------------------ geometry of the hybrid decimated loop
// Half-rate processing aliases wherever the padded waveform is not band limited: torch.stft's reflect padding puts a
// kink at samples 0 and L-1, measured 3e-3 relative in the first/last five frames and < 1e-7 elsewhere.  The loop
// therefore keeps two full-rate edge strips [0, E) and [L-E, L), E = W + H:
//   * forward STFT: the first 3 frame pairs (frames 0..5) and the pairs from pr_tail = (T-6)/2 on read the strips at
//     full rate, every other pair reads the odd-sample waveform xo;
//   * inverse STFT: the frames that overlap the strips are chunk 0 (frames <= 15) and chunks >= c_tail = (T-17)/G;
//     those chunks are ALSO evaluated at full rate into `nslots` extra partial-sum slots (slot 0 = chunk 0, slot s =
//     chunk c_tail + s - 1).
struct rf_gl_dec_geom {
    int E, pr_tail, c_tail, nslots, n_edge_pairs, nxo;
};
RF_HD rf_gl_dec_geom rf_dec_geom(int T, int G, int H, int W) {
    rf_gl_dec_geom d;
    const int nchunks = (T + G - 1) / G;
    d.E = W + H;
    d.pr_tail = (T - 6) / 2;
    d.c_tail = (T - 17) / G;
    d.nslots = 1 + nchunks - d.c_tail;
    d.n_edge_pairs = 3 + (T + 1) / 2 - d.pr_tail;
    d.nxo = (H * (T - 1) - 1) / 2;
    return d;
}
RF_HD bool rf_dec_ok(int T, int G) { return T >= 4 * G; }

// full-rate overlap-add sample i inside an edge strip from the edge slots; part_e: [2 groups][nslots][PL]
RF_HD float rf_ola_sample_edge(int i, const float* part_e, float env, int T, int G, int PL, int c_tail, int nslots,
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
            const int slot = c == 0 ? 0 : c - c_tail + 1;
            if (off >= 0 && off < (nf - 1) * H + W)
                acc += part_e[(static_cast<size_t>(g) * nslots + slot) * PL + off];
        }
    return acc / env;
}

