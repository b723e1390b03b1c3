// One translation unit per (line length, precision): compiled with -DSLM_LINE_L=<L> -DSLM_LINE_PREC=<0|1>.
// Lines of 8192 points and more (32 points per thread) are built for the row kernels only: they serve
// the slab-decomposed transform of one very large plane, whose column steps run on rows after an exchange.
#include "line_ops.cuh"

namespace slm {
#if SLM_LINE_PREC == 0
using LineReal = float;
#else
using LineReal = double;
#endif
using Rows = RowLaunch<LineReal, SLM_LINE_L>;

#define SLM_CAT3(a, b, c) a##b##_##c
#define SLM_TABLE_NAME(l, p) SLM_CAT3(line_table_, l, p)

static int row_pass_(int alg, const RowArgs& a, cudaStream_t s) { return Rows::row_pass(alg, a, s); }
static int row_plain_(const PlainRowArgs& a, cudaStream_t s) { return Rows::row_plain(a, s); }
static int row_fourier_(const RowFourierArgs& a, cudaStream_t s) { return Rows::row_fourier(a, s); }

#if SLM_LINE_L < 8192
using Cols = ColLaunch<LineReal, SLM_LINE_L>;
static int col_pass_(int alg, const ColArgs& a, cudaStream_t s) { return Cols::col_pass(alg, a, s); }
static int col_plain_(const PlainColArgs& a, cudaStream_t s) { return Cols::col_plain(a, s); }
static int col_group_(int mode, const ColGroupArgs& ga, const void* in, const void* out, int ctas, cudaStream_t s) {
    return Cols::col_group(mode, ga, in, out, ctas, s);
}
static void prepare_() { Rows::prepare(); Cols::prepare(); }
extern const LineTable SLM_TABLE_NAME(SLM_LINE_L, SLM_LINE_PREC);
const LineTable SLM_TABLE_NAME(SLM_LINE_L, SLM_LINE_PREC) = {
    SLM_LINE_L, SLM_LINE_PREC,
    Rows::RG::NR, Rows::RG::THREADS, Rows::RG::SMEM,
    Cols::CG::TC, Cols::CG::THREADS, Cols::CG::SMEM,
    &prepare_, &row_pass_, &row_plain_, &col_pass_, &col_plain_,
    Cols::GG::OK ? 1 : 0, Cols::GG::ROWB, (Cols::GG::OK && ColWarpGeom<LineReal, SLM_LINE_L>::OK) ? 1 : 0, &col_group_,
    &row_fourier_, 0,
};
#else
static void prepare_() { Rows::prepare(); }
extern const LineTable SLM_TABLE_NAME(SLM_LINE_L, SLM_LINE_PREC);
const LineTable SLM_TABLE_NAME(SLM_LINE_L, SLM_LINE_PREC) = {
    SLM_LINE_L, SLM_LINE_PREC,
    Rows::RG::NR, Rows::RG::THREADS, Rows::RG::SMEM,
    0, 0, 0,
    &prepare_, &row_pass_, &row_plain_, nullptr, nullptr,
    0, 0, 0, nullptr,
    &row_fourier_, 1,
};
#endif
}  // namespace slm
