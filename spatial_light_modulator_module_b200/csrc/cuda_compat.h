// Thin indirection between the kernel sources and the CUDA toolchain.
//
// Product build (nvcc, sm_100a): every macro maps 1:1 onto the CUDA construct it names.
// Host-emulation build (-DSLM_EMULATE, g++ only; tests/emu): the same kernel SOURCE is run
// on the CPU with one fibre per CUDA thread (tests/emu/emu_runtime.h) so index arithmetic,
// barriers and reductions can be checked without a GPU.  The emulation library is test
// infrastructure; the Python package never loads it.
#pragma once

#ifdef SLM_EMULATE
#include "emu_runtime.h"
#else
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#define SLM_HD __host__ __device__ __forceinline__
#define SLM_DEV __device__ __forceinline__
#define SLM_GLOBAL __global__
#define SLM_HOSTDEV __host__ __device__
#define SLM_LAUNCH_BOUNDS(t, b) __launch_bounds__(t, b)
// dynamic shared memory of the running CTA
#define SLM_DYN_SMEM(name) extern __shared__ __align__(1024) unsigned char name[]   // (swizzled TMA tiles and the XOR-addressed exchange assume 1 KB)
#define SLM_STATIC_SMEM __shared__
#define SLM_LAUNCH(kernel, grid, block, smem, stream, ...) kernel<<<grid, block, smem, stream>>>(__VA_ARGS__)
#define SLM_RESTRICT __restrict__
// Programmatic dependent launch: the kernel may start while its predecessor in the stream drains; it must
// call griddep_wait() before touching anything the predecessor wrote.
#define SLM_LAUNCH_PDL(kernel, grid, block, smem, stream, ...) slm::launch_pdl(kernel, grid, block, smem, stream, __VA_ARGS__)

namespace slm {
// Set by the engine per run (registry.cu): short, latency-bound passes (a single SLM-size plane) gain from the
// overlap; on large batches the early-resident dependents cost a few percent (measured), so they launch plainly.
extern thread_local bool tl_pdl;
template <class K, class... A> inline void launch_pdl(K kernel, dim3 grid, dim3 block, size_t smem, cudaStream_t stream, A... args) {
    if (!tl_pdl) { kernel<<<grid, block, smem, stream>>>(args...); return; }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    cudaLaunchKernelEx(&cfg, kernel, args...);
}
// Cooperative launch: the whole grid is resident at once (the launch fails otherwise), so CTAs may wait for one
// another through global memory.
#define SLM_LAUNCH_COOP(kernel, grid, block, smem, stream, ...) slm::launch_coop(kernel, grid, block, smem, stream, __VA_ARGS__)
template <class K, class... A> inline void launch_coop(K kernel, dim3 grid, dim3 block, size_t smem, cudaStream_t stream, A... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeCooperative;
    attr[0].val.cooperative = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    cudaLaunchKernelEx(&cfg, kernel, args...);
}
SLM_DEV void spin_pause() { __nanosleep(20); }       // inside a wait on another CTA's progress
SLM_DEV void griddep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
SLM_DEV void griddep_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
template <typename T> SLM_DEV T ld_ro(const T* p) { return __ldg(p); }     // immutable during the launch
template <typename T> SLM_DEV T ld_cg(const T* p) { return __ldcg(p); }    // streamed plane data (L2 only)
template <typename T> SLM_DEV void st_cg(T* p, T v) { __stcg(p, v); }
SLM_DEV void fence_device() { __threadfence(); }
SLM_DEV void fence_block() { __threadfence_block(); }
SLM_DEV unsigned atomic_add_shared(unsigned* p, unsigned v) { return atomicAdd(p, v); }
SLM_DEV unsigned atomic_inc_wrap(unsigned* p, unsigned limit) { return atomicInc(p, limit); }
SLM_DEV unsigned atomic_max_u32(unsigned* p, unsigned v) { return atomicMax(p, v); }
SLM_DEV unsigned atomic_add_u32(unsigned* p, unsigned v) { return atomicAdd(p, v); }
SLM_DEV long long clock_now() { return clock64(); }
SLM_DEV unsigned ld_acquire(const unsigned* p) { unsigned v; asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
SLM_DEV unsigned long long ld_acquire_u64(const unsigned long long* p) { unsigned long long v; asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory"); return v; }
SLM_DEV float shfl_xor(float v, int m) { return __shfl_xor_sync(0xffffffffu, v, m); }
SLM_DEV double shfl_xor(double v, int m) { return __shfl_xor_sync(0xffffffffu, v, m); }
SLM_DEV unsigned shfl_idx(unsigned v, int src) { return __shfl_sync(0xffffffffu, v, src); }
SLM_DEV void sync_cta() { __syncthreads(); }
SLM_DEV void sync_warp() { __syncwarp(); }
// warpgroup-wide register reallocation (all four warps of an aligned warpgroup execute the same one)
template <int N> SLM_DEV void reg_alloc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N> SLM_DEV void reg_dealloc() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }
// named barrier `id` (1..15) over `nthreads` threads (whole warps) of the CTA
SLM_DEV void sync_named(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }
// IEEE operations that must not be contracted into FMAs (bit parity with numpy)
SLM_DEV double mul_rn(double a, double b) { return __dmul_rn(a, b); }
SLM_DEV double add_rn(double a, double b) { return __dadd_rn(a, b); }
SLM_DEV double sub_rn(double a, double b) { return __dsub_rn(a, b); }
SLM_DEV double div_rn(double a, double b) { return __ddiv_rn(a, b); }
SLM_DEV double sqrt_rn(double a) { return __dsqrt_rn(a); }
SLM_DEV float rsqrt_fast(float a) { return rsqrtf(a); }
SLM_DEV double rsqrt_fast(double a) { return rsqrt(a); }
}  // namespace slm
#endif
