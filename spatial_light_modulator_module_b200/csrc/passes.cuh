// The pass kernels of the hologram loops (sm_100a).
//
// A 2-D transform is a row pass and a column pass.  Because the in-CTA line FFT returns every
// point to the thread that loaded it (fft_tile.cuh), two transforms that meet at a pointwise
// step are fused into ONE pass over HBM:
//
//   SLM-plane pass  (rows)    : finish ifft2  ->  phase-only projection (GS) / tangent-gradient
//                               update of x (GD)  ->  start fft2
//   Fourier-plane pass (cols) : finish fft2   ->  amplitude replacement + error sums (GS) /
//                               mask*F*(I-T) + error sum (GD)  ->  start ifft2
//
// so one GS iteration (algorithms.py:30-38) reads and writes the field twice instead of four
// times, and one GD iteration (algorithms.py:84-92) adds a read-only column pass for the global
// max that line 86 needs before the gradient can be formed.
#pragma once
#include "engine_types.h"
#include "fft_tile.cuh"

namespace slm {

// ---- pointwise pieces ---------------------------------------------------------------------------
// exp(1j*angle(z)) as written at algorithms.py:30,33: z/|z|, with angle(+0)=0 -> 1, angle(-0)=pi -> -1
template <typename R> SLM_DEV cpx<R> unit_phasor(cpx<R> z) {
    const R m2 = z.x * z.x + z.y * z.y;
    cpx<R> r;
    if (m2 == (R)0) { r.x = copysign((R)1, z.x); r.y = (R)0; return r; }
    const R inv = rsqrt_fast(m2);
    r.x = z.x * inv; r.y = z.y * inv;
    return r;
}

// dEdX_complex (algorithms.py:179-185) == (g - xh <xh, g>) / |x| with xh = x/|x|; returns the
// updated x (algorithms.py:91).
template <typename R> SLM_DEV cpx<R> tangent_step(cpx<R> x, cpx<R> g, R lr) {
    const R inv = rsqrt_fast(x.x * x.x + x.y * x.y);
    cpx<R> xh; xh.x = x.x * inv; xh.y = x.y * inv;
    const R dot = xh.x * g.x + xh.y * g.y;
    const R step = lr * inv;
    x.x -= step * (g.x - xh.x * dot);
    x.y -= step * (g.y - xh.y * dot);
    return x;
}

// ---- deterministic reductions ---------------------------------------------------------------------
SLM_DEV Partial combine(Partial p, Partial q) {
    p.mx = fmax(p.mx, q.mx); p.a += q.a; p.b += q.b; p.c += q.c; return p;
}
SLM_DEV Partial warp_reduce(Partial p) {
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) {
        Partial q;
        q.mx = shfl_xor(p.mx, m); q.a = shfl_xor(p.a, m); q.b = shfl_xor(p.b, m); q.c = shfl_xor(p.c, m);
        p = combine(p, q);
    }
    return p;
}
// Result valid in thread 0.  NT = threads per CTA (multiple of 32).
template <int NT> SLM_DEV Partial block_reduce(Partial p, int t) {
    static_assert(NT % 32 == 0 && NT <= 1024, "CTA size");
    SLM_STATIC_SMEM Partial red[32];
    p = warp_reduce(p);
    if (t % 32 == 0) red[t / 32] = p;
    sync_cta();
    if (t < 32) {
        Partial q; q.mx = 0; q.a = 0; q.b = 0; q.c = 0;
        if (t < NT / 32) q = red[t];
        p = warp_reduce(q);
    }
    return p;
}
SLM_DEV Partial ld_partial(const Partial* p) {
    Partial q; q.mx = ld_cg(&p->mx); q.a = ld_cg(&p->a); q.b = ld_cg(&p->b); q.c = ld_cg(&p->c); return q;
}
// Publish this tile's partial; returns (to every thread) whether this CTA is the last of its plane,
// in which case thread 0 of it receives the plane total in `total`.
template <int NT> SLM_DEV bool publish_and_collect(Partial mine, int t, Partial* plane_partials, int tile, int tiles,
                                                   unsigned* counter, Partial& total) {
    SLM_STATIC_SMEM int is_last;
    if (t == 0) {
        plane_partials[tile] = mine;
        fence_device();
        const unsigned ticket = atomic_inc_wrap(counter, (unsigned)tiles - 1);   // wraps to 0: reusable
        is_last = (ticket == (unsigned)tiles - 1);
    }
    sync_cta();
    const bool last = is_last != 0;
    if (!last) return false;
    fence_device();
    if (t < 32) {
        Partial q; q.mx = 0; q.a = 0; q.b = 0; q.c = 0;
        for (int i = t; i < tiles; i += 32) q = combine(q, ld_partial(plane_partials + i));
        total = warp_reduce(q);
    }
    return true;
}

// ---- geometry --------------------------------------------------------------------------------------
constexpr int floor_pow2(int v) { int p = 1; while (2 * p <= v) p *= 2; return p; }
template <typename R, int L> struct RowGeom {
    using P = FftPlan<L>;
    static constexpr int M = P::M;
    static constexpr int NR = floor_pow2(256 / M);           // rows per CTA (power of two)
    static constexpr int THREADS = NR * M;
    static constexpr size_t SMEM = (size_t)NR * P::NP * sizeof(cpx<R>);
};
template <typename R, int L> struct ColGeom {
    using P = FftPlan<L>;
    static constexpr int M = P::M;
    static constexpr int TMAX = sizeof(R) == 4 ? 512 : 256;
    static constexpr int TCMAX = sizeof(R) == 4 ? 8 : 4;
    static constexpr int TCRAW = (TMAX / M) < TCMAX ? (TMAX / M) : TCMAX;
    static constexpr int TC = TCRAW >= 8 ? 8 : TCRAW >= 4 ? 4 : TCRAW >= 2 ? 2 : 1;            // columns per CTA
    static constexpr int THREADS = TC * M;
    static constexpr size_t SMEM = (size_t)TC * P::NP * sizeof(cpx<R>);
    static_assert(THREADS % 32 == 0, "column CTA must be whole warps");
};

// ---- SLM-plane pass ------------------------------------------------------------------------------
template <typename R, int W, int ALG>
SLM_GLOBAL void SLM_LAUNCH_BOUNDS((RowGeom<R, W>::THREADS), 1) row_pass_kernel(RowArgs a) {
    using G = RowGeom<R, W>;
    using P = FftPlan<W>;
    constexpr int E = P::E, M = P::M;
    SLM_DYN_SMEM(raw);
    const int t = threadIdx.x, rr = t / M, j = t % M;
    const long long grow = (long long)blockIdx.x * G::NR + rr;
    const int b = (int)(grow / a.H), y = (int)(grow % a.H);
    const PlaneStats* st = a.stats + b;
    const int done = ld_cg(&st->done);
    if (!a.final_pass && done) return;                       // uniform: a CTA never straddles planes
    cpx<R>* line = reinterpret_cast<cpx<R>*>(raw) + (size_t)rr * P::NP;
    const cpx<R>* tw = static_cast<const cpx<R>*>(a.tw);
    const size_t base = ((size_t)b * a.H + y) * W + j;
    const size_t ibase = (size_t)y * W + j;
    const R* inc = static_cast<const R*>(a.inc);
    cpx<R> v[E];

    if (ALG == ALG_GS) {
        if (a.source == ROW_FROM_Y) {
            const cpx<R>* Y = static_cast<const cpx<R>*>(a.Y) + base;
#pragma unroll
            for (int r = 0; r < E; ++r) v[r] = ld_plane(Y + r * M);
            line_fft<R, W, +1, 1>(v, line, j, tw);           // A = ifft2(D) up to a positive scale
#pragma unroll
            for (int r = 0; r < E; ++r) if (!a.final_pass) v[r] = unit_phasor(v[r]);
        } else if (a.source == ROW_FROM_A32) {
            // first phasor in complex64, as the reference computes it (algorithms.py:27,30)
            const cpx<float>* A = static_cast<const cpx<float>*>(a.A32) + base;
#pragma unroll
            for (int r = 0; r < E; ++r) {
                cpx<float> z = ld_plane(A + r * M);
                if (!a.final_pass) z = unit_phasor(z);
                v[r].x = (R)z.x; v[r].y = (R)z.y;
            }
        } else {
            const cpx<R>* F = static_cast<const cpx<R>*>(a.field) + base;
#pragma unroll
            for (int r = 0; r < E; ++r) {
                v[r] = ld_plane(F + r * M);
                if (a.source == ROW_FROM_A && !a.final_pass) v[r] = unit_phasor(v[r]);
            }
        }
        if (a.final_pass) {                                   // hologram = angle(A), algorithms.py:48
            double* h = a.hologram + base;
#pragma unroll
            for (int r = 0; r < E; ++r) h[r * M] = atan2((double)v[r].y, (double)v[r].x);
            return;
        }
        if (inc && a.source != ROW_FROM_FIELD) {
#pragma unroll
            for (int r = 0; r < E; ++r) { const R s = ld_ro(inc + ibase + r * M); v[r].x *= s; v[r].y *= s; }
        }
    } else {
        cpx<R>* xp = static_cast<cpx<R>*>(a.x) + base;
        cpx<R> xx[E];
#pragma unroll
        for (int r = 0; r < E; ++r) xx[r] = ld_plane(xp + r * M);
        if (a.source == ROW_FROM_Y) {
            const cpx<R>* Y = static_cast<const cpx<R>*>(a.Y) + base;
#pragma unroll
            for (int r = 0; r < E; ++r) v[r] = ld_plane(Y + r * M);
            line_fft<R, W, +1, 1>(v, line, j, tw);
            // the update of the iteration whose Fourier-plane pass produced Y: lr of THAT iteration
            const R lr = (R)ld_ro(a.lr + (ld_cg(&st->iters) - 1));
            const R nrm = (R)a.inv_hw;
#pragma unroll
            for (int r = 0; r < E; ++r) {
                cpx<R> g;                                       // dEdF = ifft2(...) * inc_amp, algorithms.py:87-89
                g.x = v[r].x * nrm; g.y = v[r].y * nrm;
                if (inc) { const R s = ld_ro(inc + ibase + r * M); g.x *= s; g.y *= s; }
                xx[r] = tangent_step(xx[r], g, lr);             // algorithms.py:90-91
                st_plane(xp + r * M, xx[r]);
            }
        }
        if (a.final_pass) {                                   // hologram = angle(input), algorithms.py:111
            double* h = a.hologram + base;
#pragma unroll
            for (int r = 0; r < E; ++r) h[r * M] = atan2((double)xx[r].y, (double)xx[r].x);
            return;
        }
#pragma unroll
        for (int r = 0; r < E; ++r) {                            // input / abs(input) * inc_amp, algorithms.py:84
            const R inv = rsqrt_fast(xx[r].x * xx[r].x + xx[r].y * xx[r].y);
            v[r].x = xx[r].x * inv; v[r].y = xx[r].y * inv;
            if (inc) { const R s = ld_ro(inc + ibase + r * M); v[r].x *= s; v[r].y *= s; }
        }
    }
    line_fft<R, W, -1, 1>(v, line, j, tw);
    cpx<R>* X = static_cast<cpx<R>*>(a.X) + base;
#pragma unroll
    for (int r = 0; r < E; ++r) st_plane(X + r * M, v[r]);
}

// ---- Fourier-plane pass ---------------------------------------------------------------------------
template <typename R, int H, int ALG>
SLM_GLOBAL void SLM_LAUNCH_BOUNDS((ColGeom<R, H>::THREADS), 1) col_pass_kernel(ColArgs a) {
    using G = ColGeom<R, H>;
    using P = FftPlan<H>;
    constexpr int E = P::E, M = P::M, TC = G::TC;
    SLM_DYN_SMEM(raw);
    const int tiles = a.W / TC;
    const int b = blockIdx.x / tiles, tile = blockIdx.x % tiles;
    PlaneStats* st = a.stats + b;
    if (ld_cg(&st->done)) return;
    const int t = threadIdx.x, c = t % TC, j = t / TC;
    const size_t W = a.W;
    const size_t off = (size_t)b * H * W + (size_t)j * W + tile * TC + c;
    const cpx<R>* tw = static_cast<const cpx<R>*>(a.tw);
    cpx<R>* line = reinterpret_cast<cpx<R>*>(raw) + c;
    const double s0 = ld_cg(&st->scale), imax = ld_cg(&st->imax), norm = ld_ro(a.norm + b);

    cpx<R> v[E];
    const cpx<R>* X = static_cast<const cpx<R>*>(a.X) + off;
#pragma unroll
    for (int r = 0; r < E; ++r) v[r] = ld_plane(X + (size_t)r * M * W);
    // target grey level and its amplitude (GS) / weight (GD), fetched while the transform runs
    R tv[E], aux[E];
    if (a.T8) {
        const uint8_t* T = a.T8 + off;
        const R* lut = static_cast<const R*>(a.lut);
#pragma unroll
        for (int r = 0; r < E; ++r) { const int g = ld_ro(T + (size_t)r * M * W); tv[r] = (R)g; aux[r] = ld_ro(lut + g); }
    } else {
        const R* T = static_cast<const R*>(a.Treal) + off;
        const R* Q = static_cast<const R*>(a.plane2) + off;
#pragma unroll
        for (int r = 0; r < E; ++r) { tv[r] = ld_ro(T + (size_t)r * M * W); aux[r] = ld_ro(Q + (size_t)r * M * W); }
    }
    line_fft<R, H, -1, TC>(v, line, j, tw);                   // C = fft2(B)  /  med_output

    Partial p; p.mx = 0; p.a = 0; p.b = 0; p.c = 0;
#pragma unroll
    for (int r = 0; r < E; ++r) {
        const R m2 = v[r].x * v[r].x + v[r].y * v[r].y;       // |C|^2, algorithms.py:36 / :85
        if (ALG == ALG_GS) {
            // error of this iteration against the scale s0 of the previous one; the exact scale
            // s = norm/max is folded in by the last tile (see finish below).
            const double u = s0 * (double)m2, d = u - (double)tv[r];
            p.mx = fmax(p.mx, (double)m2); p.a += d * d; p.b += d * u; p.c += u * u;
            const cpx<R> ph = unit_phasor(v[r]);                // D = |amp| * exp(1j*angle(C)), :33
            v[r].x = aux[r] * ph.x; v[r].y = aux[r] * ph.y;
        } else {
            const double I = ((double)m2 * norm) / imax;        // output, algorithms.py:86
            const double d = I - (double)tv[r];
            p.a += d * d;
            const R dr = (R)d;                                   // mask * med_output * (output - T), :88
            v[r].x = (aux[r] * v[r].x) * dr; v[r].y = (aux[r] * v[r].y) * dr;
        }
    }
    line_fft<R, H, +1, TC>(v, line, j, tw);
    cpx<R>* Y = static_cast<cpx<R>*>(a.Y) + off;
#pragma unroll
    for (int r = 0; r < E; ++r) st_plane(Y + (size_t)r * M * W, v[r]);

    p = block_reduce<G::THREADS>(p, t);
    Partial tot;
    if (!publish_and_collect<G::THREADS>(p, t, a.partial + (size_t)b * tiles, tile, tiles, a.counter + b, tot)) return;
    if (t == 0) {
        const double hw = (double)H * (double)W;
        double err;
        if (ALG == ALG_GS) {
            const double s = norm / tot.mx;                      // algorithms.py:37
            const double dl = (s0 != 0.0) ? s / s0 - 1.0 : 0.0;
            err = (tot.a + 2.0 * dl * tot.b + dl * dl * tot.c) / hw;   // == sum((s*I - T)^2)/HW, :38,:162
            st->imax = tot.mx; st->scale = s;
        } else {
            err = tot.a / hw;                                    // algorithms.py:92
        }
        const int k = st->iters;
        a.err_curve[(size_t)b * a.max_loops + k] = err;
        st->err = err; st->iters = k + 1;
        st->done = !(err > a.tolerance);                         // loop condition, algorithms.py:29,83
    }
}

// ---- plain row transform (setup, preview, slm_fft2) --------------------------------------------------
template <typename R, int W>
SLM_GLOBAL void SLM_LAUNCH_BOUNDS((RowGeom<R, W>::THREADS), 1) row_plain_kernel(PlainRowArgs a) {
    using G = RowGeom<R, W>;
    using P = FftPlan<W>;
    constexpr int E = P::E, M = P::M;
    SLM_DYN_SMEM(raw);
    const int t = threadIdx.x, rr = t / M, j = t % M;
    const long long grow = (long long)blockIdx.x * G::NR + rr;
    cpx<R>* line = reinterpret_cast<cpx<R>*>(raw) + (size_t)rr * P::NP;
    const cpx<R>* tw = static_cast<const cpx<R>*>(a.tw);
    const size_t base = (size_t)grow * W + j;
    cpx<R> v[E];
    if (a.input == IN_COMPLEX) {
        const cpx<R>* in = static_cast<const cpx<R>*>(a.in) + base;
#pragma unroll
        for (int r = 0; r < E; ++r) v[r] = ld_plane(in + r * M);
    } else if (a.input == IN_LUT_U8) {
        const R* lut = static_cast<const R*>(a.lut);
#pragma unroll
        for (int r = 0; r < E; ++r) { v[r].x = ld_ro(lut + ld_ro(a.T8 + base + r * M)); v[r].y = 0; }
    } else if (a.input == IN_REAL) {
        const R* in = static_cast<const R*>(a.in) + base;
#pragma unroll
        for (int r = 0; r < E; ++r) { v[r].x = ld_ro(in + r * M); v[r].y = 0; }
    } else {
        const double* in = static_cast<const double*>(a.in) + base;
#pragma unroll
        for (int r = 0; r < E; ++r) { const double h = ld_ro(in + r * M); v[r].x = (R)cos(h); v[r].y = (R)sin(h); }
    }
    if (a.inverse) line_fft<R, W, +1, 1>(v, line, j, tw);
    else line_fft<R, W, -1, 1>(v, line, j, tw);
    cpx<R>* out = static_cast<cpx<R>*>(a.out) + base;
#pragma unroll
    for (int r = 0; r < E; ++r) st_plane(out + r * M, v[r]);
}

// ---- plain column transform: complex out, max-only, or normalised intensity out -------------------------
template <typename R, int H>
SLM_GLOBAL void SLM_LAUNCH_BOUNDS((ColGeom<R, H>::THREADS), 1) col_plain_kernel(PlainColArgs a) {
    using G = ColGeom<R, H>;
    using P = FftPlan<H>;
    constexpr int E = P::E, M = P::M, TC = G::TC;
    SLM_DYN_SMEM(raw);
    const int tiles = a.W / TC;
    const int b = blockIdx.x / tiles, tile = blockIdx.x % tiles;
    PlaneStats* st = a.stats ? a.stats + b : nullptr;
    if (a.output == OUT_STATS && ld_cg(&st->done)) return;    // in-loop use (GD): finished planes rest
    const int t = threadIdx.x, c = t % TC, j = t / TC;
    const size_t W = a.W;
    const size_t off = (size_t)b * H * W + (size_t)j * W + tile * TC + c;
    const cpx<R>* tw = static_cast<const cpx<R>*>(a.tw);
    cpx<R>* line = reinterpret_cast<cpx<R>*>(raw) + c;
    cpx<R> v[E];
    const cpx<R>* in = static_cast<const cpx<R>*>(a.in) + off;
#pragma unroll
    for (int r = 0; r < E; ++r) v[r] = ld_plane(in + (size_t)r * M * W);
    if (a.inverse) line_fft<R, H, +1, TC>(v, line, j, tw);
    else line_fft<R, H, -1, TC>(v, line, j, tw);

    if (a.output == OUT_COMPLEX) {
        cpx<R>* out = static_cast<cpx<R>*>(a.out) + off;
        const R s = (R)a.scale;
#pragma unroll
        for (int r = 0; r < E; ++r) { v[r].x *= s; v[r].y *= s; st_plane(out + (size_t)r * M * W, v[r]); }
    } else if (a.output == OUT_STATS) {
        Partial p; p.mx = 0; p.a = 0; p.b = 0; p.c = 0;
#pragma unroll
        for (int r = 0; r < E; ++r) p.mx = fmax(p.mx, (double)(v[r].x * v[r].x + v[r].y * v[r].y));
        p = block_reduce<G::THREADS>(p, t);
        Partial tot;
        if (!publish_and_collect<G::THREADS>(p, t, a.partial + (size_t)b * tiles, tile, tiles, a.counter + b, tot)) return;
        if (t == 0) { st->imax = tot.mx; st->scale = ld_ro(a.norm + b) / tot.mx; }
    } else {
        double* out = static_cast<double*>(a.out) + off;
        const double imax = ld_cg(&st->imax), scale = ld_cg(&st->scale), norm = ld_ro(a.norm + b);
#pragma unroll
        for (int r = 0; r < E; ++r) {
            const double m2 = (double)(v[r].x * v[r].x + v[r].y * v[r].y);
            double I;
            if (a.output == OUT_INTENSITY_GS) I = m2 * scale;
            else if (a.output == OUT_INTENSITY_GD) I = (m2 * norm) / imax;
            else I = (m2 / imax) * norm;
            out[(size_t)r * M * W] = I;
        }
    }
}

// ---- host-side launchers ----------------------------------------------------------------------------
template <typename R, int L> struct LineOps {
    using RG = RowGeom<R, L>;
    using CG = ColGeom<R, L>;
    static int check() { cudaError_t e = cudaGetLastError(); return e == cudaSuccess ? 0 : -(int)e - 1000; }
    static void prepare() {
        cudaFuncSetAttribute(row_pass_kernel<R, L, ALG_GS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)RG::SMEM);
        cudaFuncSetAttribute(row_pass_kernel<R, L, ALG_GD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)RG::SMEM);
        cudaFuncSetAttribute(row_plain_kernel<R, L>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)RG::SMEM);
        cudaFuncSetAttribute(col_pass_kernel<R, L, ALG_GS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CG::SMEM);
        cudaFuncSetAttribute(col_pass_kernel<R, L, ALG_GD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CG::SMEM);
        cudaFuncSetAttribute(col_plain_kernel<R, L>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CG::SMEM);
    }
    static int row_pass(int alg, const RowArgs& a, cudaStream_t s) {
        const dim3 grid((unsigned)((long long)a.B * a.H / RG::NR)), block(RG::THREADS);
        if (alg == ALG_GS) SLM_LAUNCH((row_pass_kernel<R, L, ALG_GS>), grid, block, RG::SMEM, s, a);
        else SLM_LAUNCH((row_pass_kernel<R, L, ALG_GD>), grid, block, RG::SMEM, s, a);
        return check();
    }
    static int row_plain(const PlainRowArgs& a, cudaStream_t s) {
        const dim3 grid((unsigned)((long long)a.B * a.H / RG::NR)), block(RG::THREADS);
        SLM_LAUNCH((row_plain_kernel<R, L>), grid, block, RG::SMEM, s, a);
        return check();
    }
    static int col_pass(int alg, const ColArgs& a, cudaStream_t s) {
        const dim3 grid((unsigned)((long long)a.B * (a.W / CG::TC))), block(CG::THREADS);
        if (alg == ALG_GS) SLM_LAUNCH((col_pass_kernel<R, L, ALG_GS>), grid, block, CG::SMEM, s, a);
        else SLM_LAUNCH((col_pass_kernel<R, L, ALG_GD>), grid, block, CG::SMEM, s, a);
        return check();
    }
    static int col_plain(const PlainColArgs& a, cudaStream_t s) {
        const dim3 grid((unsigned)((long long)a.B * (a.W / CG::TC))), block(CG::THREADS);
        SLM_LAUNCH((col_plain_kernel<R, L>), grid, block, CG::SMEM, s, a);
        return check();
    }
};

}  // namespace slm
