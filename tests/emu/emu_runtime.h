// TEST INFRASTRUCTURE ONLY -- host emulation of the CUDA execution model for the kernel sources
// in spatial_light_modulator_module_b200/csrc (compiled with g++ -DSLM_EMULATE).
//
// One CTA at a time; every CUDA thread of the CTA is a ucontext fibre on a single OS thread, so
// execution is deterministic.  __syncthreads() and warp shuffles are real rendezvous points
// (a fibre yields until its whole CTA / warp has arrived), which makes missing or misplaced
// barriers show up as wrong results or dead-locks ("emu: deadlock" abort) instead of passing
// silently.  Shared memory is poisoned with NaN bytes before each CTA.
//
// Not emulated: memory-model reordering, bank conflicts, occupancy.  This checks the LOGIC of
// the kernels (indexing, butterflies, reductions, the last-CTA-done protocol); speed and the
// hardware-facing parts are checked on the B200.
#pragma once
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <ucontext.h>

#include <cmath>
#include <vector>

struct dim3 {
    unsigned x, y, z;
    dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {}
};
typedef void* cudaStream_t;
typedef void* cudaEvent_t;
typedef int cudaError_t;
enum { cudaSuccess = 0 };
enum cudaMemcpyKind { cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyDefault };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize };

inline cudaError_t cudaMalloc(void** p, size_t n) { *p = aligned_alloc(256, (n + 255) / 256 * 256); memset(*p, 0xFF, (n + 255) / 256 * 256); return *p ? 0 : 2; }
inline cudaError_t cudaFree(void* p) { free(p); return 0; }
inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t) { memcpy(d, s, n); return 0; }
inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { memcpy(d, s, n); return 0; }
inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t) { memset(d, v, n); return 0; }
inline cudaError_t cudaMemset(void* d, int v, size_t n) { memset(d, v, n); return 0; }
inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return 0; }
inline cudaError_t cudaGetLastError() { return 0; }
inline cudaError_t cudaSetDevice(int) { return 0; }
inline cudaError_t cudaGetDevice(int* d) { *d = 0; return 0; }
inline const char* cudaGetErrorString(cudaError_t) { return "emulated"; }
template <class K> inline cudaError_t cudaFuncSetAttribute(K, cudaFuncAttribute, int) { return 0; }

namespace emu {

struct Fibre {
    ucontext_t ctx;
    void* stack;
    unsigned tid;
    bool done;
    int shfl_parity;
};
struct WarpSlot {
    double buf[2][32];
    unsigned count, gen;
};
struct State {
    std::vector<Fibre> fibres;
    std::vector<WarpSlot> warps;
    ucontext_t main_ctx;
    Fibre* cur = nullptr;
    unsigned n = 0, alive = 0;
    unsigned bar_count = 0, bar_gen = 0;
    unsigned named_count[16] = {0}, named_gen[16] = {0};
    dim3 block_idx, block_dim, grid_dim;
    unsigned char* smem = nullptr;
    void (*body)(void*) = nullptr;
    void* body_arg = nullptr;
    unsigned long progress = 0;      // barriers completed, mbarrier arrivals, global atomics, threads finished ...
    unsigned long weak = 0;          // warp rendezvous: a polling warp produces these for ever, so they only postpone the verdict
    std::vector<unsigned char> smem_store;
};
// The CTA whose fibres are running.  Ordinary launches run one CTA at a time on `solo`; a cooperative launch
// (launch_coop) keeps one State per CTA of the grid and switches this pointer as it goes round them.
inline State*& current() { static State solo; static State* cur = &solo; return cur; }
inline State& S() { return *current(); }

inline void yield() { State& s = S(); swapcontext(&s.cur->ctx, &s.main_ctx); }

inline void trampoline() {
    State& s = S();
    s.body(s.body_arg);
    s.cur->done = true;
    s.alive--;
    s.progress++;
    // a thread that exits no longer takes part in CTA barriers (Volta+ semantics)
    if (s.alive && s.bar_count == s.alive) { s.bar_count = 0; s.bar_gen++; }
    swapcontext(&s.cur->ctx, &s.main_ctx);
}

inline void sync_cta() {
    State& s = S();
    unsigned gen = s.bar_gen;
    s.progress++;
    if (++s.bar_count == s.alive) { s.bar_count = 0; s.bar_gen++; return; }
    while (s.bar_gen == gen) yield();
}
inline void sync_named(int id, int nthreads) {
    State& s = S();
    if (id < 1 || id > 15 || nthreads % 32) { fprintf(stderr, "emu: bad named barrier %d/%d\n", id, nthreads); abort(); }
    unsigned gen = s.named_gen[id];
    s.progress++;
    if (++s.named_count[id] == (unsigned)nthreads) { s.named_count[id] = 0; s.named_gen[id]++; return; }
    while (s.named_gen[id] == gen) yield();
}
inline unsigned warp_lanes(unsigned w) { State& s = S(); unsigned r = s.n - w * 32; return r < 32 ? r : 32; }
inline void sync_warp() {
    State& s = S();
    unsigned w = s.cur->tid / 32;
    WarpSlot& ws = s.warps[w];
    unsigned gen = ws.gen;
    s.weak++;
    if (++ws.count == warp_lanes(w)) { ws.count = 0; ws.gen++; return; }
    while (ws.gen == gen) yield();
}
template <class T> inline T shfl_xor(T v, int m) {
    static_assert(sizeof(T) <= sizeof(double), "shuffle payload");
    State& s = S();
    unsigned lane = s.cur->tid % 32;
    WarpSlot& ws = s.warps[s.cur->tid / 32];
    int p = s.cur->shfl_parity;
    s.cur->shfl_parity ^= 1;
    memcpy(&ws.buf[p][lane], &v, sizeof(T));
    sync_warp();
    T r;
    memcpy(&r, &ws.buf[p][(lane ^ (unsigned)m) % 32], sizeof(T));
    return r;
}

template <class T> inline T shfl_idx(T v, int src) {
    State& s = S();
    unsigned lane = s.cur->tid % 32;
    WarpSlot& ws = s.warps[s.cur->tid / 32];
    int p = s.cur->shfl_parity;
    s.cur->shfl_parity ^= 1;
    memcpy(&ws.buf[p][lane], &v, sizeof(T));
    sync_warp();
    T r;
    memcpy(&r, &ws.buf[p][src % 32], sizeof(T));
    return r;
}

template <class F> void setup_cta(State& s, dim3 bidx, dim3 grid, dim3 block, size_t smem_bytes, F& f) {
    const unsigned n = block.x * block.y * block.z;
    const size_t kStack = 256 * 1024;
    if (s.fibres.size() < n) {
        size_t old = s.fibres.size();
        s.fibres.resize(n);
        for (size_t i = old; i < n; ++i) s.fibres[i].stack = malloc(kStack);
    }
    s.warps.assign((n + 31) / 32, WarpSlot{});
    s.n = s.alive = n;
    s.bar_count = 0;
    for (int i = 0; i < 16; ++i) s.named_count[i] = 0;
    s.block_idx = bidx; s.block_dim = block; s.grid_dim = grid;
    s.smem_store.assign(smem_bytes + 2048, 0xFF);
    s.smem = (unsigned char*)(((uintptr_t)s.smem_store.data() + 1023) & ~(uintptr_t)1023);
    s.body = [](void* a) { (*static_cast<F*>(a))(); };
    s.body_arg = &f;
    for (unsigned i = 0; i < n; ++i) {
        Fibre& fb = s.fibres[i];
        fb.tid = i; fb.done = false; fb.shfl_parity = 0;
        getcontext(&fb.ctx);
        fb.ctx.uc_stack.ss_sp = fb.stack;
        fb.ctx.uc_stack.ss_size = kStack;
        fb.ctx.uc_link = nullptr;
        makecontext(&fb.ctx, (void (*)())trampoline, 0);
    }
}
// one turn of every live fibre of the CTA
inline void sweep_cta(State& s) {
    for (unsigned i = 0; i < s.n; ++i) {
        if (s.fibres[i].done) continue;
        s.cur = &s.fibres[i];
        swapcontext(&s.main_ctx, &s.fibres[i].ctx);
    }
    s.cur = nullptr;
}

template <class F> void run_cta(dim3 bidx, dim3 grid, dim3 block, size_t smem_bytes, F& f) {
    State& s = S();
    setup_cta(s, bidx, grid, block, smem_bytes, f);
    unsigned long idle = 0;
    while (s.alive) {
        unsigned long before = s.progress, weak_before = s.weak;
        sweep_cta(s);
        if (s.progress != before) { idle = 0; continue; }
        if (s.alive && (s.weak == weak_before || ++idle > 200000)) { fprintf(stderr, "emu: deadlock in CTA (%u,%u)\n", bidx.x, bidx.y); abort(); }
    }
}

template <class F> void launch(dim3 grid, dim3 block, size_t smem_bytes, F f) {
    for (unsigned z = 0; z < grid.z; ++z)
        for (unsigned y = 0; y < grid.y; ++y)
            for (unsigned x = 0; x < grid.x; ++x) run_cta(dim3(x, y, z), grid, block, smem_bytes, f);
}

// Cooperative launch: every CTA of the (1-D) grid is resident at once, so CTAs may wait for each other through
// global memory (spin loops must call spin_pause(), which yields here).  CTAs take turns, one sweep each.
template <class F> void launch_coop(dim3 grid, dim3 block, size_t smem_bytes, F f) {
    static std::vector<State*> pool;
    while (pool.size() < grid.x) pool.push_back(new State);
    State* const saved = current();
    for (unsigned x = 0; x < grid.x; ++x) { current() = pool[x]; setup_cta(*pool[x], dim3(x, 1, 1), grid, block, smem_bytes, f); }
    unsigned long idle = 0;
    for (;;) {
        unsigned long alive = 0, moved = 0, weak = 0;
        for (unsigned x = 0; x < grid.x; ++x) {
            State& s = *pool[x];
            if (!s.alive) continue;
            current() = &s;
            const unsigned long before = s.progress, weak_before = s.weak;
            sweep_cta(s);
            moved += s.progress - before;
            weak += s.weak - weak_before;
            alive += s.alive;
        }
        if (!alive) break;
        if (moved) { idle = 0; continue; }
        if (!weak || ++idle > 200000) { fprintf(stderr, "emu: deadlock in a cooperative grid of %u CTAs\n", grid.x); abort(); }
    }
    current() = saved;
}

struct Idx { unsigned x, y, z; };
inline Idx thread_idx() {
    State& s = S();
    unsigned t = s.cur->tid;
    return Idx{t % s.block_dim.x, (t / s.block_dim.x) % s.block_dim.y, t / (s.block_dim.x * s.block_dim.y)};
}
}  // namespace emu

#define threadIdx (emu::thread_idx())
#define blockIdx (emu::S().block_idx)
#define blockDim (emu::S().block_dim)
#define gridDim (emu::S().grid_dim)

#define SLM_HD inline
#define SLM_DEV inline
#define SLM_GLOBAL
#define SLM_HOSTDEV
#define SLM_LAUNCH_BOUNDS(t, b)
#define SLM_DYN_SMEM(name) unsigned char* name = emu::S().smem
#define SLM_STATIC_SMEM static
#define SLM_RESTRICT __restrict__
#define SLM_LAUNCH(kernel, grid, block, smem, stream, ...) \
    emu::launch(grid, block, smem, [=]() { kernel(__VA_ARGS__); })
#define SLM_LAUNCH_PDL SLM_LAUNCH
#define SLM_LAUNCH_COOP(kernel, grid, block, smem, stream, ...) \
    emu::launch_coop(grid, block, smem, [=]() { kernel(__VA_ARGS__); })

namespace slm {
template <typename T> inline T ld_ro(const T* p) { return *p; }
template <typename T> inline T ld_cg(const T* p) { return *p; }
template <typename T> inline void st_cg(T* p, T v) { *p = v; }
inline void griddep_wait() {}
inline void griddep_launch() {}
inline void fence_device() {}
inline void fence_block() {}
inline unsigned atomic_add_shared(unsigned* p, unsigned v) { unsigned o = *p; *p = o + v; return o; }
inline unsigned atomic_inc_wrap(unsigned* p, unsigned limit) { unsigned o = *p; *p = (o >= limit) ? 0 : o + 1; return o; }
// (global atomics change what OTHER CTAs see: they count as progress for the cooperative grid's dead-lock detection)
inline unsigned atomic_max_u32(unsigned* p, unsigned v) { emu::S().progress++; unsigned o = *p; if (v > o) *p = v; return o; }
inline unsigned atomic_add_u32(unsigned* p, unsigned v) { emu::S().progress++; unsigned o = *p; *p = o + v; return o; }
inline unsigned ld_acquire(const unsigned* p) { return *p; }
inline unsigned long long ld_acquire_u64(const unsigned long long* p) { return *p; }
inline long long clock_now() { return 0; }
inline void spin_pause() { emu::yield(); }          // a wait on another CTA's progress: let the other fibres / CTAs run
}  // namespace slm
inline float __uint_as_float(unsigned u) { float f; memcpy(&f, &u, 4); return f; }
inline unsigned __float_as_uint(float f) { unsigned u; memcpy(&u, &f, 4); return u; }
namespace slm {
inline float shfl_xor(float v, int m) { return emu::shfl_xor(v, m); }
inline double shfl_xor(double v, int m) { return emu::shfl_xor(v, m); }
inline unsigned shfl_idx(unsigned v, int src) { return emu::shfl_idx(v, src); }
inline void sync_warp() { (void)emu::shfl_idx(0u, 0); }
template <int N> inline void reg_alloc() {}
template <int N> inline void reg_dealloc() {}
inline void sync_cta() { emu::sync_cta(); }
inline void sync_named(int id, int nthreads) { emu::sync_named(id, nthreads); }

// compiled with -ffp-contract=off, so plain operators are single IEEE operations
inline double mul_rn(double a, double b) { return a * b; }
inline double add_rn(double a, double b) { return a + b; }
inline double sub_rn(double a, double b) { return a - b; }
inline double div_rn(double a, double b) { return a / b; }
inline double sqrt_rn(double a) { return std::sqrt(a); }
inline float rsqrt_fast(float a) { return 1.0f / std::sqrt(a); }
inline double rsqrt_fast(double a) { return 1.0 / std::sqrt(a); }
}  // namespace slm

struct float2 { float x, y; };
struct uint2 { unsigned x, y; };
struct alignas(16) double2 { double x, y; };
