// In-CTA line FFTs for the hologram passes (sm_100a).
//
// A "line" is one row or one column of an H x W plane.  Each thread owns E points of a line
// in registers; a line of N = E * m * E points is transformed by
//   stage 1 : radix-E butterflies on register data            -> shared memory
//   middle  : one radix-m stage, shared -> registers -> shared (m in {1,2,3,4,8,16})
//   last    : radix-E butterflies, shared -> registers
// (Stockham autosort, so results come out in natural order and thread j again owns the
// points j + r*N/E, r = 0..E-1: exactly the mapping stage 1 consumes.  A fused
// inverse-FFT -> pointwise -> forward-FFT chain therefore never leaves registers between the
// two transforms, and global loads/stores are always unit-stride across threads.)
//
// Replaces scipy.fft.fft2/ifft2 (pocketfft/DUCC) at the reference call sites
// algorithms.py:27,31,34,84,88,155 and generate_hologram.py:25.  Convention follows scipy:
// forward unnormalised; the 1/(H*W) of the inverse is applied by the caller where it matters.
#pragma once
#include "cuda_compat.h"

namespace slm {

template <typename R> struct cpx;
template <> struct alignas(8) cpx<float> { float x, y; };
template <> struct alignas(16) cpx<double> { double x, y; };

template <typename R> struct vec2;
template <> struct vec2<float> { using type = float2; };
template <> struct vec2<double> { using type = double2; };

// Plane data is streamed (read once / written once per pass): keep it out of L1.
template <typename R> SLM_DEV cpx<R> ld_plane(const cpx<R>* p) {
    typename vec2<R>::type v = ld_cg(reinterpret_cast<const typename vec2<R>::type*>(p));
    cpx<R> r; r.x = v.x; r.y = v.y; return r;
}
template <typename R> SLM_DEV void st_plane(cpx<R>* p, cpx<R> v) {
    typename vec2<R>::type t; t.x = v.x; t.y = v.y;
    st_cg(reinterpret_cast<typename vec2<R>::type*>(p), t);
}
template <typename R> SLM_DEV cpx<R> ld_const(const cpx<R>* p) {
    typename vec2<R>::type v = ld_ro(reinterpret_cast<const typename vec2<R>::type*>(p));
    cpx<R> r; r.x = v.x; r.y = v.y; return r;
}

template <typename R> SLM_DEV cpx<R> cadd(cpx<R> a, cpx<R> b) { cpx<R> r; r.x = a.x + b.x; r.y = a.y + b.y; return r; }
template <typename R> SLM_DEV cpx<R> csub(cpx<R> a, cpx<R> b) { cpx<R> r; r.x = a.x - b.x; r.y = a.y - b.y; return r; }
template <typename R> SLM_DEV cpx<R> cmul(cpx<R> a, cpx<R> b) {
    cpx<R> r; r.x = a.x * b.x - a.y * b.y; r.y = a.x * b.y + a.y * b.x; return r;
}
template <typename R> SLM_DEV cpx<R> cconj(cpx<R> a) { a.y = -a.y; return a; }
template <typename R> SLM_DEV cpx<R> cscale(cpx<R> a, R s) { a.x *= s; a.y *= s; return a; }

// multiply by exp(DIR * i * pi/2): DIR=-1 (forward) -> -i ; DIR=+1 (inverse) -> +i
template <int DIR, typename R> SLM_DEV cpx<R> rot90(cpx<R> a) {
    cpx<R> r;
    if (DIR < 0) { r.x = a.y; r.y = -a.x; } else { r.x = -a.y; r.y = a.x; }
    return r;
}
// multiply by exp(DIR * 2*pi*i * P/16)
template <int DIR, int P, typename R> SLM_DEV cpx<R> mul_w16(cpx<R> a) {
    constexpr int p = ((P % 16) + 16) % 16;
    if constexpr (p == 0) return a;
    else if constexpr (p == 4) return rot90<DIR>(a);
    else if constexpr (p == 8) { a.x = -a.x; a.y = -a.y; return a; }
    else if constexpr (p == 12) return rot90<-DIR>(a);
    else {
        constexpr double C[16] = {1.0, 0.92387953251128673848, 0.70710678118654752440, 0.38268343236508977173,
                                  0.0, -0.38268343236508977173, -0.70710678118654752440, -0.92387953251128673848,
                                  -1.0, -0.92387953251128673848, -0.70710678118654752440, -0.38268343236508977173,
                                  0.0, 0.38268343236508977173, 0.70710678118654752440, 0.92387953251128673848};
        constexpr double S[16] = {0.0, 0.38268343236508977173, 0.70710678118654752440, 0.92387953251128673848,
                                  1.0, 0.92387953251128673848, 0.70710678118654752440, 0.38268343236508977173,
                                  0.0, -0.38268343236508977173, -0.70710678118654752440, -0.92387953251128673848,
                                  -1.0, -0.92387953251128673848, -0.70710678118654752440, -0.38268343236508977173};
        const R c = (R)C[p], s = (R)(DIR < 0 ? -S[p] : S[p]);
        cpx<R> r; r.x = a.x * c - a.y * s; r.y = a.x * s + a.y * c; return r;
    }
}

// ---- small DFTs on register arrays, natural order in and out -----------------------------
template <int DIR, typename R> SLM_DEV void dft2(cpx<R>& a, cpx<R>& b) {
    cpx<R> t = csub(a, b); a = cadd(a, b); b = t;
}
template <int DIR, typename R> SLM_DEV void dft3(cpx<R>& a, cpx<R>& b, cpx<R>& c) {
    const R h = (R)0.86602540378443864676;          // sin(2*pi/3)
    cpx<R> s = cadd(b, c), d = csub(b, c);
    cpx<R> m; m.x = a.x - (R)0.5 * s.x; m.y = a.y - (R)0.5 * s.y;
    cpx<R> q = rot90<DIR>(d); q.x *= h; q.y *= h;    // DIR * i * sin(2pi/3) * (b - c)
    a = cadd(a, s);
    b = cadd(m, q);
    c = csub(m, q);
}
template <int DIR, typename R> SLM_DEV void dft4(cpx<R>& a0, cpx<R>& a1, cpx<R>& a2, cpx<R>& a3) {
    cpx<R> t0 = cadd(a0, a2), t1 = csub(a0, a2), t2 = cadd(a1, a3), t3 = rot90<DIR>(csub(a1, a3));
    a0 = cadd(t0, t2); a1 = cadd(t1, t3); a2 = csub(t0, t2); a3 = csub(t1, t3);
}
template <int DIR, typename R> SLM_DEV void dft8(cpx<R>* v) {
    // 8 = 4 x 2 : DFT4 over n1 (stride 2), twiddle W8^(n2*k1), DFT2 over n2
    dft4<DIR>(v[0], v[2], v[4], v[6]);
    dft4<DIR>(v[1], v[3], v[5], v[7]);
    v[3] = mul_w16<DIR, 2>(v[3]);
    v[5] = rot90<DIR>(v[5]);
    v[7] = mul_w16<DIR, 6>(v[7]);
    dft2<DIR>(v[0], v[1]); dft2<DIR>(v[2], v[3]); dft2<DIR>(v[4], v[5]); dft2<DIR>(v[6], v[7]);
    // position 2*k1 + k2 holds X[k1 + 4*k2]
    cpx<R> x1 = v[2], x2 = v[4], x3 = v[6], x4 = v[1], x5 = v[3], x6 = v[5];
    v[1] = x1; v[2] = x2; v[3] = x3; v[4] = x4; v[5] = x5; v[6] = x6;
}
template <int DIR, typename R> SLM_DEV void dft16(cpx<R>* v) {
    // 16 = 4 x 4
#pragma unroll
    for (int n2 = 0; n2 < 4; ++n2) dft4<DIR>(v[n2], v[n2 + 4], v[n2 + 8], v[n2 + 12]);
    // position n2 + 4*k1 *= W16^(n2*k1)
    v[5] = mul_w16<DIR, 1>(v[5]);  v[9] = mul_w16<DIR, 2>(v[9]);   v[13] = mul_w16<DIR, 3>(v[13]);
    v[6] = mul_w16<DIR, 2>(v[6]);  v[10] = mul_w16<DIR, 4>(v[10]); v[14] = mul_w16<DIR, 6>(v[14]);
    v[7] = mul_w16<DIR, 3>(v[7]);  v[11] = mul_w16<DIR, 6>(v[11]); v[15] = mul_w16<DIR, 9>(v[15]);
#pragma unroll
    for (int k1 = 0; k1 < 4; ++k1) dft4<DIR>(v[4 * k1], v[4 * k1 + 1], v[4 * k1 + 2], v[4 * k1 + 3]);
    // position 4*k1 + k2 holds X[k1 + 4*k2] : transpose
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = a + 1; b < 4; ++b) { cpx<R> t = v[4 * a + b]; v[4 * a + b] = v[4 * b + a]; v[4 * b + a] = t; }
}
template <int RAD, int DIR, typename R> SLM_DEV void dft_small(cpx<R>* v) {
    if constexpr (RAD == 2) dft2<DIR>(v[0], v[1]);
    else if constexpr (RAD == 3) dft3<DIR>(v[0], v[1], v[2]);
    else if constexpr (RAD == 4) dft4<DIR>(v[0], v[1], v[2], v[3]);
    else if constexpr (RAD == 8) dft8<DIR>(v);
    else if constexpr (RAD == 16) dft16<DIR>(v);
    else static_assert(RAD == 2, "unsupported radix");
}

// ---- twiddles ----------------------------------------------------------------------------
// tw[q] = exp(-2*pi*i*q/N), q in [0,N), computed in extended precision on the host.
template <int DIR, typename R> SLM_DEV cpx<R> tw_load(const cpx<R>* tw, int q) {
    cpx<R> w = ld_const(tw + q);
    if (DIR > 0) w.y = -w.y;
    return w;
}
// v[r] *= w^r for r = 1..RAD-1, with w^r built by a depth-log2 product tree (error ~ log2(RAD) ulp)
template <int RAD, typename R> SLM_DEV void apply_twiddle_powers(cpx<R>* v, cpx<R> w1) {
    cpx<R> w[RAD > 1 ? RAD : 2];
    w[1] = w1;
#pragma unroll
    for (int r = 2; r < RAD; ++r) w[r] = cmul(w[r / 2], w[r - r / 2]);
#pragma unroll
    for (int r = 1; r < RAD; ++r) v[r] = cmul(v[r], w[r]);
}

// ---- line FFT ----------------------------------------------------------------------------
template <int N> struct FftPlan {
    static constexpr int E = (N >= 256) ? 16 : 8;      // points per thread
    static constexpr int M = N / E;                    // threads per line
    static constexpr int MID = N / (E * E);            // middle radix (1 = none)
    static constexpr int NP = N + N / E;               // padded line length in shared memory
    static_assert(E * E * MID == N, "line length must be E*E*m");
    static_assert(MID == 1 || MID == 2 || MID == 3 || MID == 4 || MID == 8 || MID == 16, "unsupported line length");
    static_assert(MID <= E || MID == 3, "middle radix must fit the per-thread register tile");
    SLM_HOSTDEV static constexpr int pad(int i) { return i + i / E; }
};

// Transform one line.  v[r] = x[j + r*M] on entry, X[j + r*M] on exit.  `line` points at the
// line's element 0 in shared memory; element i lives at line[pad(i) * STRIDE].  Every thread of
// the CTA must call this together (it contains __syncthreads()).
template <typename R, int N, int DIR, int STRIDE>
SLM_DEV void line_fft(cpx<R>* v, cpx<R>* line, int j, const cpx<R>* __restrict__ tw) {
    using P = FftPlan<N>;
    constexpr int E = P::E, M = P::M, MID = P::MID;

    // stage 1 (Ns = 1): no twiddles; butterfly output r goes to position j*E + r
    dft_small<E, DIR>(v);
#pragma unroll
    for (int r = 0; r < E; ++r) line[P::pad(j * E + r) * STRIDE] = v[r];
    sync_cta();

    if constexpr (MID > 1 && MID != 3) {
        constexpr int Q = E / MID;          // butterflies per thread
        constexpr int Ns = E;
#pragma unroll
        for (int e = 0; e < E; ++e) v[e] = line[P::pad(j + e * M) * STRIDE];
        sync_cta();
#pragma unroll
        for (int q = 0; q < Q; ++q) {
            const int b = j + q * M;
            const int k = b % Ns;
            cpx<R> a[MID];
#pragma unroll
            for (int r = 0; r < MID; ++r) a[r] = v[q + r * Q];
            apply_twiddle_powers<MID>(a, tw_load<DIR>(tw, k * (N / (Ns * MID))));
            dft_small<MID, DIR>(a);
            const int o = (b / Ns) * (Ns * MID) + k;
#pragma unroll
            for (int r = 0; r < MID; ++r) line[P::pad(o + r * Ns) * STRIDE] = a[r];
        }
        sync_cta();
    } else if constexpr (MID == 3) {
        constexpr int NB = N / 3, QMAX = (NB + M - 1) / M, Ns = E;
        cpx<R> a[QMAX][3];
#pragma unroll
        for (int q = 0; q < QMAX; ++q) {
            const int b = j + q * M;
            if (b < NB) {
#pragma unroll
                for (int r = 0; r < 3; ++r) a[q][r] = line[P::pad(b + r * NB) * STRIDE];
            }
        }
        sync_cta();
#pragma unroll
        for (int q = 0; q < QMAX; ++q) {
            const int b = j + q * M;
            if (b < NB) {
                const int k = b % Ns;
                apply_twiddle_powers<3>(a[q], tw_load<DIR>(tw, k * (N / (Ns * 3))));
                dft3<DIR>(a[q][0], a[q][1], a[q][2]);
                const int o = (b / Ns) * (Ns * 3) + k;
#pragma unroll
                for (int r = 0; r < 3; ++r) line[P::pad(o + r * Ns) * STRIDE] = a[q][r];
            }
        }
        sync_cta();
    }

    // last stage (Ns = M): twiddle W_N^(r*j), butterfly, result r is X[j + r*M]
#pragma unroll
    for (int r = 0; r < E; ++r) v[r] = line[P::pad(j + r * M) * STRIDE];
    sync_cta();                      // the tile may be overwritten by the next transform
    apply_twiddle_powers<E>(v, tw_load<DIR>(tw, j));
    dft_small<E, DIR>(v);
}

}  // namespace slm
