// TMA tile prefetch for the persistent column kernels (sm_100a): cp.async.bulk.tensor.2d loads of
// an [H rows][TC columns] tile of the field into dense shared memory, completion on an mbarrier.
// The emulation build replaces the three primitives by synchronous host code with the same
// visible semantics (tests/emu).
#pragma once
#include "cuda_compat.h"

namespace slm {

#ifdef SLM_EMULATE
struct alignas(64) TileMap {
    const unsigned char* base;     // element (row 0, column 0)
    size_t pitch_bytes;            // bytes between rows
    int col_bytes;                 // bytes of one complex column element
    int box_rows;
};
struct TileBarrier { volatile unsigned phase; unsigned pad; };
inline void tile_barrier_init(TileBarrier* b) { b->phase = 0; }
inline void tile_barrier_fence() {}
// copy rows [row0, row0+rows) x columns [col0, col0+cols) into dst (dense [rows][cols])
inline void tile_prefetch(const TileMap& tm, void* dst, TileBarrier* bar, long long row0, int rows, int col0, int cols, int col_bytes) {
    unsigned char* d = static_cast<unsigned char*>(dst);
    const size_t seg = (size_t)cols * col_bytes;
    for (int r = 0; r < rows; ++r)
        memcpy(d + (size_t)r * seg, tm.base + (size_t)(row0 + r) * tm.pitch_bytes + (size_t)col0 * col_bytes, seg);
    bar->phase = bar->phase + 1;
}
inline void tile_wait(TileBarrier* bar, unsigned uses_before) { while (bar->phase == uses_before) emu::yield(); }
#define SLM_GRID_CONSTANT
#else
#include <cuda.h>
struct alignas(64) TileMap { CUtensorMap map; int box_rows; };
typedef unsigned long long TileBarrier;
#define SLM_GRID_CONSTANT __grid_constant__

SLM_DEV unsigned smem_addr(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
SLM_DEV void tile_barrier_init(TileBarrier* b) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_addr(b)) : "memory");
}
SLM_DEV void tile_barrier_fence() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
// One thread: arm the barrier with the tile's byte count and issue one bulk tensor copy per box of rows.
SLM_DEV void tile_prefetch(const TileMap& tm, void* dst, TileBarrier* bar, long long row0, int rows, int col0, int cols, int col_bytes) {
    const unsigned bytes = (unsigned)rows * (unsigned)cols * (unsigned)col_bytes;
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
    const unsigned long long desc = reinterpret_cast<unsigned long long>(&tm.map);
    const int c0 = col0 * 2;                                            // inner coordinate in real elements (2 per complex)
    for (int r = 0; r < rows; r += tm.box_rows) {
        const unsigned d = smem_addr(static_cast<unsigned char*>(dst) + (size_t)r * cols * col_bytes);
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                     ::"r"(d), "l"(desc), "r"(smem_addr(bar)), "r"(c0), "r"((int)(row0 + r))
                     : "memory");
    }
}
// All threads: wait until the use number `uses_before` (0, 1, 2, ...) of this barrier has completed.
SLM_DEV void tile_wait(TileBarrier* bar, unsigned uses_before) {
    const unsigned parity = uses_before & 1u;
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "SLM_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra SLM_DONE;\n"
        "bra SLM_WAIT;\n"
        "SLM_DONE:\n"
        "}\n" ::"r"(smem_addr(bar)),
        "r"(parity)
        : "memory");
}
#endif

}  // namespace slm
