// Generic STFT / iSTFT engine for geometries the prime-factor engine does not cover (sample rates other than 44.1 kHz,
// custom window / padding durations, hops that do not divide the window): torch.stft / torch.istft semantics
// (center=True, reflect padding, window zero-padded to n_fft, one-sided spectrum; TA/functional/functional.py:54-145,
// 255-353) for any even n_fft with n_fft/2 = 2^a 3^b 5^c 7^d <= 14000.
//
// One CTA per (frame, clip).  A frame of n_fft real samples is packed as n_fft/2 complex points z[n] = x[2n] + i x[2n+1],
// transformed by a mixed-radix Stockham autosort FFT in shared memory (two ping-pong buffers, no bit reversal, twiddles
// from one table of (n_fft/2)-th roots built in fp64), and un-packed with the n_fft-th roots:
//     X[k] = (Z[k] + conj Z[N2-k]) / 2  - i w^k (Z[k] - conj Z[N2-k]) / 2,   w = exp(-2 pi i / n_fft)
// The inverse runs the same steps backwards.  Spectra live in the [B][T][J] layout of the Griffin-Lim workspace (J live
// bins in natural order, bins[j] = k).  This path is about coverage, not speed: no pruning of the zero-padded samples,
// radix butterflies by direct summation.
#pragma once
#include "rf_dft.cuh"

struct rf_gen_tab {
    const rf_c32* roots2;   // [N2]
    const rf_c32* rootsN;   // [N2 + 1]
    const float* window;    // [W]
    const int32_t* bins;    // [J] j -> k
    int N, N2, W, H, lo;    // lo = (N - W) / 2: first live sample of a frame
    int J, nrad;
    int rad[16];
};

// in-place (ping-pong) FFT of N2 points; returns the buffer holding the result
template <bool INV>
__device__ rf_c32* rf_gen_fft(rf_c32* a, rf_c32* b, const rf_gen_tab& g) {
    const int N2 = g.N2;
    int Ns = 1;
    for (int s = 0; s < g.nrad; ++s) {
        const int R = g.rad[s], M = N2 / R, step = N2 / (Ns * R), stepR = N2 / R;
        for (int j = threadIdx.x; j < M; j += blockDim.x) {
            const int k = j % Ns;
            rf_c32 v[7];
#pragma unroll
            for (int t = 0; t < 7; ++t) {
                if (t < R) {
                    rf_c32 w = g.roots2[((k * t) % (Ns * R)) * step];
                    if (INV) w.y = -w.y;
                    v[t] = c_mul(a[j + t * M], w);
                }
            }
            const int j0 = (j - k) * R + k;
#pragma unroll
            for (int u = 0; u < 7; ++u) {
                if (u < R) {
                    rf_c32 acc = v[0];
#pragma unroll
                    for (int t = 1; t < 7; ++t) {
                        if (t < R) {
                            rf_c32 w = g.roots2[((u * t) % R) * stepR];
                            if (INV) w.y = -w.y;
                            acc = c_add(acc, c_mul(v[t], w));
                        }
                    }
                    b[j0 + u * Ns] = acc;
                }
            }
        }
        __syncthreads();
        rf_c32* t = a;
        a = b;
        b = t;
        Ns *= R;
    }
    return a;
}

// STFT of frame t of clip b -> out[(b*T + t)*J + j] for the live bins; x: [B][L] un-padded signal
__global__ void __launch_bounds__(256) k_gen_stft(rf_gen_tab g, const float* __restrict__ x, int L, int T,
                                                  rf_c32* __restrict__ out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    rf_c32* A = reinterpret_cast<rf_c32*>(smem_raw);
    rf_c32* Bf = A + g.N2 + 1;
    const int t = blockIdx.x, b = blockIdx.y;
    const float* xb = x + static_cast<size_t>(b) * L;
    const int base = t * g.H - g.N / 2;                 // signal index of frame sample 0
    for (int n = threadIdx.x; n < g.N2; n += blockDim.x) {
        float s[2];
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const int m = 2 * n + e;
            float val = 0.f;
            if (m >= g.lo && m < g.lo + g.W) {
                int idx = base + m;
                if (idx < 0) idx = -idx;                 // reflect padding (pad_mode="reflect")
                if (idx >= L) idx = 2 * (L - 1) - idx;
                val = xb[idx] * g.window[m - g.lo];
            }
            s[e] = val;
        }
        A[n] = c_make(s[0], s[1]);
    }
    __syncthreads();
    const rf_c32* Z = rf_gen_fft<false>(A, Bf, g);
    rf_c32* dst = out + (static_cast<size_t>(b) * T + t) * g.J;
    for (int j = threadIdx.x; j < g.J; j += blockDim.x) {
        const int k = g.bins[j];
        const rf_c32 zk = Z[k == g.N2 ? 0 : k];
        const rf_c32 zc = c_conj(Z[(g.N2 - k) % g.N2]);
        const rf_c32 e = c_make(0.5f * (zk.x + zc.x), 0.5f * (zk.y + zc.y));
        const rf_c32 o = c_make(0.5f * (zk.x - zc.x), 0.5f * (zk.y - zc.y));
        const rf_c32 wo = c_mul(g.rootsN[k], o);         // w^k * o ; X = e - i * (w^k o)
        dst[j] = c_make(e.x + wo.y, e.y - wo.x);
    }
}

// inverse STFT of frame t: C[j] = S * cur (mode 0) or S * normalise(cur - m prev) (mode 1: the Griffin-Lim phase update,
// TA/functional/functional.py:337-340), irfft, window; frames[(b*T + t)*W + i] = window[i] * frame sample lo + i
__global__ void __launch_bounds__(256) k_gen_istft(rf_gen_tab g, const float* __restrict__ S, const rf_c32* __restrict__ cur,
                                                   const rf_c32* __restrict__ prev, int mode, float m, int T,
                                                   float* __restrict__ frames) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    rf_c32* A = reinterpret_cast<rf_c32*>(smem_raw);
    rf_c32* Bf = A + g.N2 + 1;
    const int t = blockIdx.x, b = blockIdx.y;
    const size_t row = (static_cast<size_t>(b) * T + t) * g.J;
    for (int k = threadIdx.x; k <= g.N2; k += blockDim.x) Bf[k] = c_make(0.f, 0.f);
    __syncthreads();
    for (int j = threadIdx.x; j < g.J; j += blockDim.x) {
        rf_c32 a = cur[row + j];
        if (mode) {
            if (prev) {
                const rf_c32 pv = prev[row + j];
                a = c_make(a.x - m * pv.x, a.y - m * pv.y);
            }
            const float inv = 1.f / (sqrtf(a.x * a.x + a.y * a.y) + 1e-16f);
            a = c_make(a.x * inv, a.y * inv);
        }
        const float s = S[row + j];
        const int k = g.bins[j];
        Bf[k] = c_make(s * a.x, (k == 0 || k == g.N2) ? 0.f : s * a.y);      // c2r ignores the imaginary part of DC / Nyquist
    }
    __syncthreads();
    for (int k = threadIdx.x; k < g.N2; k += blockDim.x) {
        const rf_c32 xk = Bf[k], xc = c_conj(Bf[g.N2 - k]);
        const rf_c32 e = c_add(xk, xc), o = c_sub(xk, xc);
        const rf_c32 wo = c_mul(c_conj(g.rootsN[k]), o);  // w^-k * o ; Z = e + i * (w^-k o)
        A[k] = c_make(e.x - wo.y, e.y + wo.x);
    }
    __syncthreads();
    // A and Bf overlap by one element (Bf = A + N2 + 1 entries apart): the FFT ping-pongs between A and Bf[0..N2)
    const rf_c32* z = rf_gen_fft<true>(A, Bf, g);
    const float invN = 1.f / static_cast<float>(g.N);
    float* dst = frames + (static_cast<size_t>(b) * T + t) * g.W;
    for (int i = threadIdx.x; i < g.W; i += blockDim.x) {
        const int mm = g.lo + i;
        const rf_c32 zz = z[mm >> 1];
        dst[i] = ((mm & 1) ? zz.y : zz.x) * invN * g.window[i];
    }
}

// overlap-add of the windowed frames, divided by the window envelope, trimmed by n_fft/2 (torch.istft)
__global__ void k_gen_ola(const float* __restrict__ frames, const float* __restrict__ win2, int T, int H, int W, int c0, int L,
                          float* __restrict__ x) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y;
    if (i >= L) return;
    const int q = c0 + i;                                 // frame t covers q in [tH, tH + W)
    const int t_hi = min(T - 1, q / H);
    int t_lo = (q - W + H) / H;
    if (q - W + 1 <= 0) t_lo = 0;
    float acc = 0.f, env = 0.f;
    for (int t = t_lo; t <= t_hi; ++t) {
        acc += frames[(static_cast<size_t>(b) * T + t) * W + (q - t * H)];
        env += win2[q - t * H];
    }
    x[static_cast<size_t>(b) * L + i] = acc / env;
}

// mel[b][m][t] = sum_e w_e |R[b][t][j_e]| over the CSR column of mel filter m
__global__ void k_gen_mel_from_TJ(const rf_c32* __restrict__ R, int T, int J, int n_mels, const int32_t* __restrict__ col_ptr,
                                  const int32_t* __restrict__ col_j, const float* __restrict__ col_w, float* __restrict__ mel) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x, m = blockIdx.y, b = blockIdx.z;
    if (t >= T) return;
    const rf_c32* row = R + (static_cast<size_t>(b) * T + t) * J;
    float acc = 0.f;
    for (int e = col_ptr[m]; e < col_ptr[m + 1]; ++e) {
        const rf_c32 v = row[col_j[e]];
        acc += col_w[e] * sqrtf(v.x * v.x + v.y * v.y);
    }
    mel[(static_cast<size_t>(b) * n_mels + m) * T + t] = acc;
}
