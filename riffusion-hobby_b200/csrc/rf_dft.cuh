// Small in-register complex DFTs (3,5,7,9,10,49 points) used by the 4410-point
// prime-factor FFT (10 x 9 x 49, no inter-axis twiddles).
//
// All routines are in-place on a register array of rf_c32 and are compiled for
// both device and host (the host build is only used by tests/hostemu to check the
// butterflies against a naive DFT without a GPU).
//
// template<bool INV>: INV=false computes X[k] = sum x[n] e^{-2 pi i nk/P},
//                     INV=true  computes X[k] = sum x[n] e^{+2 pi i nk/P} (unnormalised).
#pragma once

#if defined(__CUDACC__)
#define RF_HD __host__ __device__ __forceinline__
#else
#define RF_HD inline
#endif

struct alignas(8) rf_c32 {
    float x, y;
};

RF_HD rf_c32 c_make(float x, float y) {
    rf_c32 r;
    r.x = x;
    r.y = y;
    return r;
}
RF_HD rf_c32 c_add(rf_c32 a, rf_c32 b) { return c_make(a.x + b.x, a.y + b.y); }
RF_HD rf_c32 c_sub(rf_c32 a, rf_c32 b) { return c_make(a.x - b.x, a.y - b.y); }
RF_HD rf_c32 c_mul(rf_c32 a, rf_c32 b) {
    return c_make(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
RF_HD rf_c32 c_conj(rf_c32 a) { return c_make(a.x, -a.y); }
// a * (cr + i ci) with compile-time-ish constants
RF_HD rf_c32 c_mulk(rf_c32 a, float cr, float ci) {
    return c_make(a.x * cr - a.y * ci, a.x * ci + a.y * cr);
}
// a + i*b  /  a - i*b
RF_HD rf_c32 c_add_i(rf_c32 a, rf_c32 b) { return c_make(a.x - b.y, a.y + b.x); }
RF_HD rf_c32 c_sub_i(rf_c32 a, rf_c32 b) { return c_make(a.x + b.y, a.y - b.x); }

// ---------------------------------------------------------------- radix 2
RF_HD void dft2(rf_c32& a, rf_c32& b) {
    rf_c32 t = c_sub(a, b);
    a = c_add(a, b);
    b = t;
}

// ---------------------------------------------------------------- radix 3
template <bool INV>
RF_HD void dft3(rf_c32& x0, rf_c32& x1, rf_c32& x2) {
    const float C = -0.5f;
    const float S = 0.86602540378443864676f;  // sin(2pi/3)
    rf_c32 s = c_add(x1, x2), d = c_sub(x1, x2);
    rf_c32 a = c_make(x0.x + C * s.x, x0.y + C * s.y);
    rf_c32 b = c_make(S * d.x, S * d.y);
    x0 = c_add(x0, s);
    if (INV) {
        x1 = c_add_i(a, b);
        x2 = c_sub_i(a, b);
    } else {
        x1 = c_sub_i(a, b);
        x2 = c_add_i(a, b);
    }
}

// ---------------------------------------------------------------- radix 5
template <bool INV>
RF_HD void dft5(rf_c32& x0, rf_c32& x1, rf_c32& x2, rf_c32& x3, rf_c32& x4) {
    const float C1 = 0.30901699437494742410f;   // cos(2pi/5)
    const float C2 = -0.80901699437494742410f;  // cos(4pi/5)
    const float S1 = 0.95105651629515357212f;   // sin(2pi/5)
    const float S2 = 0.58778525229247312917f;   // sin(4pi/5)
    rf_c32 s1 = c_add(x1, x4), d1 = c_sub(x1, x4);
    rf_c32 s2 = c_add(x2, x3), d2 = c_sub(x2, x3);
    rf_c32 a1 = c_make(x0.x + C1 * s1.x + C2 * s2.x, x0.y + C1 * s1.y + C2 * s2.y);
    rf_c32 a2 = c_make(x0.x + C2 * s1.x + C1 * s2.x, x0.y + C2 * s1.y + C1 * s2.y);
    rf_c32 b1 = c_make(S1 * d1.x + S2 * d2.x, S1 * d1.y + S2 * d2.y);
    rf_c32 b2 = c_make(S2 * d1.x - S1 * d2.x, S2 * d1.y - S1 * d2.y);
    x0 = c_make(x0.x + s1.x + s2.x, x0.y + s1.y + s2.y);
    if (INV) {
        x1 = c_add_i(a1, b1);
        x4 = c_sub_i(a1, b1);
        x2 = c_add_i(a2, b2);
        x3 = c_sub_i(a2, b2);
    } else {
        x1 = c_sub_i(a1, b1);
        x4 = c_add_i(a1, b1);
        x2 = c_sub_i(a2, b2);
        x3 = c_add_i(a2, b2);
    }
}

// ---------------------------------------------------------------- radix 7
template <bool INV>
RF_HD void dft7(rf_c32& x0, rf_c32& x1, rf_c32& x2, rf_c32& x3, rf_c32& x4, rf_c32& x5,
                rf_c32& x6) {
    const float C1 = 0.62348980185873353053f;   // cos(2pi/7)
    const float C2 = -0.22252093395631440429f;  // cos(4pi/7)
    const float C3 = -0.90096886790241912624f;  // cos(6pi/7)
    const float S1 = 0.78183148246802980871f;   // sin(2pi/7)
    const float S2 = 0.97492791218182360702f;   // sin(4pi/7)
    const float S3 = 0.43388373911755812048f;   // sin(6pi/7)
    rf_c32 s1 = c_add(x1, x6), d1 = c_sub(x1, x6);
    rf_c32 s2 = c_add(x2, x5), d2 = c_sub(x2, x5);
    rf_c32 s3 = c_add(x3, x4), d3 = c_sub(x3, x4);
    // k=1: angles j*1 -> (1,2,3); k=2: (2,4,6)->cos(C2,C3,C1) sin(S2,-S3,-S1)... derived below
    rf_c32 a1 = c_make(x0.x + C1 * s1.x + C2 * s2.x + C3 * s3.x,
                       x0.y + C1 * s1.y + C2 * s2.y + C3 * s3.y);
    rf_c32 a2 = c_make(x0.x + C2 * s1.x + C3 * s2.x + C1 * s3.x,
                       x0.y + C2 * s1.y + C3 * s2.y + C1 * s3.y);
    rf_c32 a3 = c_make(x0.x + C3 * s1.x + C1 * s2.x + C2 * s3.x,
                       x0.y + C3 * s1.y + C1 * s2.y + C2 * s3.y);
    // sin(2pi jk/7): k=1: S1,S2,S3 ; k=2: sin(4pi/7)=S2, sin(8pi/7)=-S3, sin(12pi/7)=-S1
    //                k=3: sin(6pi/7)=S3, sin(12pi/7)=-S1, sin(18pi/7)=sin(4pi/7)=S2
    rf_c32 b1 = c_make(S1 * d1.x + S2 * d2.x + S3 * d3.x, S1 * d1.y + S2 * d2.y + S3 * d3.y);
    rf_c32 b2 = c_make(S2 * d1.x - S3 * d2.x - S1 * d3.x, S2 * d1.y - S3 * d2.y - S1 * d3.y);
    rf_c32 b3 = c_make(S3 * d1.x - S1 * d2.x + S2 * d3.x, S3 * d1.y - S1 * d2.y + S2 * d3.y);
    x0 = c_make(x0.x + s1.x + s2.x + s3.x, x0.y + s1.y + s2.y + s3.y);
    if (INV) {
        x1 = c_add_i(a1, b1);
        x6 = c_sub_i(a1, b1);
        x2 = c_add_i(a2, b2);
        x5 = c_sub_i(a2, b2);
        x3 = c_add_i(a3, b3);
        x4 = c_sub_i(a3, b3);
    } else {
        x1 = c_sub_i(a1, b1);
        x6 = c_add_i(a1, b1);
        x2 = c_sub_i(a2, b2);
        x5 = c_add_i(a2, b2);
        x3 = c_sub_i(a3, b3);
        x4 = c_add_i(a3, b3);
    }
}

// ---------------------------------------------------------------- 9 = 3 x 3 (Cooley-Tukey)
// in: v[n], n = 3*n1 + n2 ; out: v[k], k = k1 + 3*k2 (natural order in and out)
template <bool INV>
RF_HD void dft9(rf_c32* v) {
    // twiddles W9^j = exp(-2 pi i j/9), j = 1,2,4
    const float W1R = 0.76604444311897803520f, W1I = 0.64278760968653932632f;   // cos/sin(2pi/9)
    const float W2R = 0.17364817766693034885f, W2I = 0.98480775301220805937f;   // cos/sin(4pi/9)
    const float W4R = -0.93969262078590838405f, W4I = 0.34202014332566873304f;  // cos/sin(8pi/9)
    const float sg = INV ? 1.0f : -1.0f;
    // step 1: for each n2, DFT3 over n1 (elements n2, 3+n2, 6+n2) -> t[k1][n2] stored at v[3*k1+n2]
    dft3<INV>(v[0], v[3], v[6]);
    dft3<INV>(v[1], v[4], v[7]);
    dft3<INV>(v[2], v[5], v[8]);
    // twiddle t[k1][n2] *= W9^{n2*k1}
    v[4] = c_mulk(v[4], W1R, sg * W1I);  // k1=1,n2=1
    v[5] = c_mulk(v[5], W2R, sg * W2I);  // k1=1,n2=2
    v[7] = c_mulk(v[7], W2R, sg * W2I);  // k1=2,n2=1
    v[8] = c_mulk(v[8], W4R, sg * W4I);  // k1=2,n2=2
    // step 2: for each k1, DFT3 over n2 (elements 3*k1 + {0,1,2}) -> X[k1 + 3*k2] at v[3*k1+k2]
    dft3<INV>(v[0], v[1], v[2]);
    dft3<INV>(v[3], v[4], v[5]);
    dft3<INV>(v[6], v[7], v[8]);
    // now v[3*k1 + k2] holds X[k1 + 3*k2]: transpose to natural order
    rf_c32 t;
    t = v[1]; v[1] = v[3]; v[3] = t;
    t = v[2]; v[2] = v[6]; v[6] = t;
    t = v[5]; v[5] = v[7]; v[7] = t;
}

// ---------------------------------------------------------------- 10 = 2 x 5 (prime factor)
// in: v[n] natural, out: v[k] natural
template <bool INV>
RF_HD void dft10(rf_c32* v) {
    // input map n = (5*n1 + 2*n2) mod 10 ; output k = (5*k1 + 6*k2) mod 10
    // step 1: DFT2 over n1 for each n2: elements (2*n2) and (2*n2+5) mod 10
    rf_c32 e0 = v[0], o0 = v[5];  // n2=0
    rf_c32 e1 = v[2], o1 = v[7];  // n2=1
    rf_c32 e2 = v[4], o2 = v[9];  // n2=2
    rf_c32 e3 = v[6], o3 = v[1];  // n2=3
    rf_c32 e4 = v[8], o4 = v[3];  // n2=4
    dft2(e0, o0);
    dft2(e1, o1);
    dft2(e2, o2);
    dft2(e3, o3);
    dft2(e4, o4);
    // step 2: DFT5 over n2 for k1 = 0 (e*) and k1 = 1 (o*)
    dft5<INV>(e0, e1, e2, e3, e4);
    dft5<INV>(o0, o1, o2, o3, o4);
    // k = (5*k1 + 6*k2) mod 10
    v[0] = e0; v[6] = e1; v[2] = e2; v[8] = e3; v[4] = e4;
    v[5] = o0; v[1] = o1; v[7] = o2; v[3] = o3; v[9] = o4;
}

// ---------------------------------------------------------------- 49 = 7 x 7 (Cooley-Tukey)
// W49^j = exp(-2 pi i j/49) for j = n2*k1, n2,k1 in 1..6 (j <= 36)
struct rf_w49 {
    float c[37], s[37];
};

#define RF_W49_COS \
    {1.0f, 0.9917899966239929f, 0.9672948718070984f, 0.926916778087616f,                         \
     0.8713186979293823f, 0.8014135956764221f, 0.7183493375778198f, 0.6234897971153259f,         \
     0.5183925628662109f, 0.40478333830833435f, 0.28452759981155396f, 0.1595999002456665f,       \
     0.03205157816410065f, -0.09602302312850952f, -0.22252093255519867f, -0.345365047454834f,    \
     -0.46253830194473267f, -0.5721166729927063f, -0.6723008751869202f, -0.761445939540863f,     \
     -0.8380880951881409f, -0.9009688496589661f, -0.9490557312965393f, -0.981559157371521f,      \
     -0.9979453682899475f, -0.9979453682899475f, -0.981559157371521f, -0.9490557312965393f,      \
     -0.9009688496589661f, -0.8380880951881409f, -0.761445939540863f, -0.6723008751869202f,      \
     -0.5721166729927063f, -0.46253830194473267f, -0.345365047454834f, -0.22252093255519867f,    \
     -0.09602302312850952f}
#define RF_W49_SIN \
    {0.0f, 0.1278771609067917f, 0.2536545693874359f, 0.37526699900627136f,                       \
     0.4907175600528717f, 0.598110556602478f, 0.6956825256347656f, 0.7818315029144287f,          \
     0.8551427721977234f, 0.9144126176834106f, 0.9586678743362427f, 0.9871817827224731f,         \
     0.9994862079620361f, 0.9953790903091431f, 0.9749279022216797f, 0.9384683966636658f,         \
     0.8865993022918701f, 0.8201722502708435f, 0.7402780055999756f, 0.6482284069061279f,         \
     0.5455349087715149f, 0.4338837265968323f, 0.31510820984840393f, 0.19115862250328064f,       \
     0.0640702173113823f, -0.0640702173113823f, -0.19115862250328064f, -0.31510820984840393f,    \
     -0.4338837265968323f, -0.5455349087715149f, -0.6482284069061279f, -0.7402780055999756f,     \
     -0.8201722502708435f, -0.8865993022918701f, -0.9384683966636658f, -0.9749279022216797f,     \
     -0.9953790903091431f}

// in: v[c], c = 7*c1 + c2 ; out: v[c'], c' = c1' + 7*c2'  (natural in / natural out)
// cos / sin(2 pi m / 49), m = 0..36, for code that indexes the twiddle at run time (the 7-thread radix-49 pass): a
// __constant__ table on the device (a local array would be copied to the thread's stack and read with LDL), a static one
// on the host
#if defined(__CUDACC__)
__device__ __constant__ float rf_w49_cos_dev[37] = RF_W49_COS;
__device__ __constant__ float rf_w49_sin_dev[37] = RF_W49_SIN;
#endif
static const float rf_w49_cos_host[37] = RF_W49_COS;
static const float rf_w49_sin_host[37] = RF_W49_SIN;
RF_HD float rf_w49_cos(int m) {
#if defined(__CUDA_ARCH__)
    return rf_w49_cos_dev[m];
#else
    return rf_w49_cos_host[m];
#endif
}
RF_HD float rf_w49_sin(int m) {
#if defined(__CUDA_ARCH__)
    return rf_w49_sin_dev[m];
#else
    return rf_w49_sin_host[m];
#endif
}

template <bool INV>
RF_HD void dft49(rf_c32* v) {
    const float WC[37] = RF_W49_COS;
    const float WS[37] = RF_W49_SIN;
    const float sg = INV ? 1.0f : -1.0f;
    // step 1: for each c2, DFT7 over c1 (elements c2 + 7*c1), result t[c1'][c2] at v[7*c1' + c2]
#pragma unroll
    for (int c2 = 0; c2 < 7; ++c2)
        dft7<INV>(v[c2], v[7 + c2], v[14 + c2], v[21 + c2], v[28 + c2], v[35 + c2], v[42 + c2]);
    // twiddle
#pragma unroll
    for (int k1 = 1; k1 < 7; ++k1)
#pragma unroll
        for (int c2 = 1; c2 < 7; ++c2)
            v[7 * k1 + c2] = c_mulk(v[7 * k1 + c2], WC[k1 * c2], sg * WS[k1 * c2]);
    // step 2: for each c1', DFT7 over c2 (elements 7*c1' + c2) -> X[c1' + 7*c2'] at v[7*c1' + c2']
#pragma unroll
    for (int k1 = 0; k1 < 7; ++k1)
        dft7<INV>(v[7 * k1], v[7 * k1 + 1], v[7 * k1 + 2], v[7 * k1 + 3], v[7 * k1 + 4],
                  v[7 * k1 + 5], v[7 * k1 + 6]);
    // v[7*k1 + k2] holds X[k1 + 7*k2]; the caller stores with the transposed index
    // (see dft49_out_index) so no register shuffle is needed.
}
// position in v[] (after dft49) that holds output X[cp]
RF_HD int dft49_out_index(int cp) { return 7 * (cp % 7) + cp / 7; }
