// One translation unit per (line length, precision): compiled with -DSLM_LINE_L=<L> -DSLM_LINE_PREC=<0|1>.
#include "line_ops.cuh"

namespace slm {
#if SLM_LINE_PREC == 0
using LineReal = float;
#else
using LineReal = double;
#endif
using Ops = LineOps<LineReal, SLM_LINE_L>;

#define SLM_CAT3(a, b, c) a##b##_##c
#define SLM_TABLE_NAME(l, p) SLM_CAT3(line_table_, l, p)

static int row_pass_(int alg, const RowArgs& a, cudaStream_t s) { return Ops::row_pass(alg, a, s); }
static int row_plain_(const PlainRowArgs& a, cudaStream_t s) { return Ops::row_plain(a, s); }
static int col_pass_(int alg, const ColArgs& a, cudaStream_t s) { return Ops::col_pass(alg, a, s); }
static int col_plain_(const PlainColArgs& a, cudaStream_t s) { return Ops::col_plain(a, s); }
static void prepare_() { Ops::prepare(); }
static int col_group_(int mode, const ColGroupArgs& ga, const void* in, const void* out, int ctas, cudaStream_t s) {
    return Ops::col_group(mode, ga, in, out, ctas, s);
}

extern const LineTable SLM_TABLE_NAME(SLM_LINE_L, SLM_LINE_PREC);
const LineTable SLM_TABLE_NAME(SLM_LINE_L, SLM_LINE_PREC) = {
    SLM_LINE_L, SLM_LINE_PREC,
    Ops::RG::NR, Ops::RG::THREADS, Ops::RG::SMEM,
    Ops::CG::TC, Ops::CG::THREADS, Ops::CG::SMEM,
    &prepare_, &row_pass_, &row_plain_, &col_pass_, &col_plain_,
    Ops::GG::OK ? 1 : 0, Ops::GG::ROWB, &col_group_,
};
}  // namespace slm
