// ntt_big.cu -- tuned two-pass transforms for large rings (placeholder: not covered yet).
#include "internal.hpp"

int big_tables_create(b200_ctx *) { return 0; }
void big_tables_destroy(b200_ctx *) {}
int launch_ntt_big(b200_ctx *, const u64 *, u64 *, Span, bool, int) { return 1; }
int launch_polymul_big(b200_ctx *, const u64 *, const u64 *, u64 *, Span) { return 1; }
