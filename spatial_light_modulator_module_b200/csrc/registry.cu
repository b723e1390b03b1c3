// Table of the instantiated line lengths (line_inst.cu objects).
#include "engine_types.h"
#include "line_list.h"

namespace slm {
#ifndef SLM_EMULATE
thread_local bool tl_pdl = false;
#endif
#define SLM_DECL(l) extern const LineTable line_table_##l##_0; extern const LineTable line_table_##l##_1;
SLM_LINE_LENGTHS(SLM_DECL)
#undef SLM_DECL

#define SLM_ENTRY(l) &line_table_##l##_0, &line_table_##l##_1,
static const LineTable* const kTables[] = { SLM_LINE_LENGTHS(SLM_ENTRY) };
#undef SLM_ENTRY

const LineTable* find_line_table(int L, int prec) {
    for (const LineTable* t : kTables)
        if (t->L == L && t->prec == prec) return t;
    return nullptr;
}
int supported_lengths(int* out, int cap) {
    int n = 0;
    for (const LineTable* t : kTables)
        if (t->prec == 0) { if (n < cap) out[n] = t->L; ++n; }
    return n;
}
}  // namespace slm
