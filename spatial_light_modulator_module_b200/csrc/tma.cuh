// TMA + mbarrier primitives of the column kernels (sm_100a).
//
//   tile load  : cp.async.bulk.tensor.2d.shared::cluster.global  ([H rows][TC columns] of the field ->
//                shared memory, 32B/64B-swizzled so a warp can read ONE column without bank conflicts)
//   tile store : cp.async.bulk.tensor.2d.global.shared::cta      (same image back to the other field)
//   mbarrier   : completion of loads (transaction bytes) and hand-over between the producer warp and
//                the compute warps
//
// The emulation build (tests/emu) replaces every primitive by host code with the same visible
// semantics -- including the swizzle -- so the kernels' indexing and protocol run on the CPU.
#pragma once
#include "cuda_compat.h"

namespace slm {

// Swizzle of a byte offset inside a tile buffer (buffer 1024-byte aligned):
// SWIZZLE_32B: bit 4 ^= bit 7; SWIZZLE_64B: bits 4..5 ^= bits 7..8 (CUtensorMapSwizzle semantics).
template <int ROW_BYTES> SLM_HOSTDEV inline unsigned tile_swizzle(unsigned off) {
    if (ROW_BYTES == 64) return off ^ ((off >> 3) & 0x30u);
    if (ROW_BYTES == 32) return off ^ ((off >> 3) & 0x10u);
    return off;
}

#ifdef SLM_EMULATE
struct alignas(64) TileMap {
    unsigned char* base;           // element (row 0, column 0)
    size_t pitch_bytes;            // bytes between rows
    int box_rows;
    int row_bytes;                 // bytes of one tile row (swizzle span); 0 = dense
};
struct TileBarrier { unsigned init, pending, phase; long long tx; };
inline void mbar_init(TileBarrier* b, unsigned count) { b->init = b->pending = count; b->phase = 0; b->tx = 0; }
inline void mbar_fence_init() {}
inline void mbar_check(TileBarrier* b) { if (b->pending == 0 && b->tx == 0) { b->phase ^= 1u; b->pending = b->init; } }
inline void mbar_arrive(TileBarrier* b) { emu::S().progress++; b->pending--; mbar_check(b); }
inline void mbar_wait(TileBarrier* b, unsigned parity) { while ((b->phase & 1u) == (parity & 1u)) emu::yield(); }
inline bool mbar_test(TileBarrier* b, unsigned parity) { return (b->phase & 1u) != (parity & 1u); }
inline unsigned emu_swz(int row_bytes, unsigned off) {
    return row_bytes == 64 ? tile_swizzle<64>(off) : row_bytes == 32 ? tile_swizzle<32>(off) : off;
}
// rows [row0,row0+rows) x bytes [col_byte0, col_byte0+row_bytes) -> dst (tile image, swizzled per 16-byte chunk)
inline void tile_load(const TileMap& tm, void* dst, TileBarrier* bar, long long row0, int rows, long long col_byte0, int row_bytes) {
    emu::S().progress++;
    bar->tx += (long long)rows * row_bytes;
    bar->pending--;
    unsigned char* d = static_cast<unsigned char*>(dst);
    for (int r = 0; r < rows; ++r)
        for (int ch = 0; ch < row_bytes; ch += 16)
            memcpy(d + emu_swz(tm.row_bytes, (unsigned)(r * row_bytes + ch)),
                   tm.base + (size_t)(row0 + r) * tm.pitch_bytes + col_byte0 + ch, 16);
    bar->tx -= (long long)rows * row_bytes;
    mbar_check(bar);
}
inline void tile_store(const TileMap& tm, const void* src, long long row0, int rows, long long col_byte0, int row_bytes) {
    emu::S().progress++;
    const unsigned char* s = static_cast<const unsigned char*>(src);
    for (int r = 0; r < rows; ++r)
        for (int ch = 0; ch < row_bytes; ch += 16)
            memcpy(tm.base + (size_t)(row0 + r) * tm.pitch_bytes + col_byte0 + ch,
                   s + emu_swz(tm.row_bytes, (unsigned)(r * row_bytes + ch)), 16);
}
inline void tile_store_commit() {}
inline void tile_store_wait_read() {}
inline void tile_store_wait_all() {}
inline void fence_async_smem() {}
#define SLM_GRID_CONSTANT
#else
#include <cuda.h>
struct alignas(64) TileMap { CUtensorMap map; int box_rows; int row_bytes; };
typedef unsigned long long TileBarrier;
#define SLM_GRID_CONSTANT __grid_constant__

SLM_DEV unsigned smem_addr(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
SLM_DEV void mbar_init(TileBarrier* b, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(b)), "r"(count) : "memory");
}
SLM_DEV void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
SLM_DEV void mbar_arrive(TileBarrier* b) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_addr(b)) : "memory");
}
// wait until the phase with parity `parity` has completed
SLM_DEV void mbar_wait(TileBarrier* b, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "SLM_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra SLM_DONE;\n"
        "bra SLM_WAIT;\n"
        "SLM_DONE:\n"
        "}\n" ::"r"(smem_addr(b)),
        "r"(parity & 1u)
        : "memory");
}
// has the phase with parity `parity` completed?  (no waiting)
SLM_DEV bool mbar_test(TileBarrier* b, unsigned parity) {
    unsigned ok;
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "mbarrier.test_wait.parity.shared::cta.b64 P1, [%1], %2;\n"
        "selp.u32 %0, 1, 0, P1;\n"
        "}\n" : "=r"(ok) : "r"(smem_addr(b)), "r"(parity & 1u) : "memory");
    return ok != 0;
}
// One thread: arm `bar` with the tile's byte count (this is also its arrival) and issue one bulk
// tensor copy per box of rows.  col_byte0 / row_bytes in bytes; the map's elements are 4- or 8-byte reals.
SLM_DEV void tile_load(const TileMap& tm, void* dst, TileBarrier* bar, long long row0, int rows, long long col_byte0, int row_bytes,
                       int real_bytes) {
    const unsigned bytes = (unsigned)rows * (unsigned)row_bytes;
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
    const unsigned long long desc = reinterpret_cast<unsigned long long>(&tm.map);
    const int c0 = (int)(col_byte0 / real_bytes);
    for (int r = 0; r < rows; r += tm.box_rows) {
        const unsigned d = smem_addr(static_cast<unsigned char*>(dst) + (size_t)r * row_bytes);
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                     ::"r"(d), "l"(desc), "r"(smem_addr(bar)), "r"(c0), "r"((int)(row0 + r))
                     : "memory");
    }
}
// Ask L2 for a tile that will be loaded a little later (no shared-memory destination, nothing to wait for).
SLM_DEV void tile_prefetch(const TileMap& tm, long long row0, int rows, long long col_byte0, int real_bytes) {
    const unsigned long long desc = reinterpret_cast<unsigned long long>(&tm.map);
    const int c0 = (int)(col_byte0 / real_bytes);
    for (int r = 0; r < rows; r += tm.box_rows)
        asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];" ::"l"(desc), "r"(c0), "r"((int)(row0 + r)) : "memory");
}
SLM_DEV void tile_store(const TileMap& tm, const void* src, long long row0, int rows, long long col_byte0, int row_bytes, int real_bytes) {
    const unsigned long long desc = reinterpret_cast<unsigned long long>(&tm.map);
    const int c0 = (int)(col_byte0 / real_bytes);
    for (int r = 0; r < rows; r += tm.box_rows) {
        const unsigned s = smem_addr(static_cast<const unsigned char*>(src) + (size_t)r * row_bytes);
        asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                     ::"l"(desc), "r"(s), "r"(c0), "r"((int)(row0 + r))
                     : "memory");
    }
}
SLM_DEV void tile_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
SLM_DEV void tile_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
SLM_DEV void tile_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
// make this thread's shared-memory writes visible to the async proxy (TMA store engine)
SLM_DEV void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
#endif

#ifdef SLM_EMULATE
inline void tile_prefetch(const TileMap&, long long, int, long long, int) {}
inline void tile_load(const TileMap& tm, void* dst, TileBarrier* bar, long long row0, int rows, long long col_byte0, int row_bytes, int) {
    tile_load(tm, dst, bar, row0, rows, col_byte0, row_bytes);
}
inline void tile_store(const TileMap& tm, const void* src, long long row0, int rows, long long col_byte0, int row_bytes, int) {
    tile_store(tm, src, row0, rows, col_byte0, row_bytes);
}
#endif

}  // namespace slm
