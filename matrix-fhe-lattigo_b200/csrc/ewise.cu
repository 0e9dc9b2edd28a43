// ewise.cu -- the element-wise operator family (ring/vec_ops.go, ring/operations.go) as ONE templated
// streaming kernel: all batch x (level+1) limbs in a single launch, 128-bit coalesced loads/stores,
// per-limb constants fetched once per thread.  HBM-bound: 16-32 B/element, ~15-40 integer ops.
#include "internal.hpp"

namespace {

template <int OP>
__device__ __forceinline__ u64 apply(u64 x, u64 y, u64 z, u64 sc, const LimbConst &c) {
    const u64 q = c.q;
    switch (OP) {
    case B200_OP_ADD: return csub(x + y, q);                                   // vec_ops.go:7-29
    case B200_OP_ADD_LAZY: return x + y;                                       // :31-53
    case B200_OP_SUB: return csub((x + q) - y, q);                             // :55-77
    case B200_OP_SUB_LAZY: return x + q - y;                                   // :79-101
    case B200_OP_NEG: return q - x;                                            // :103-123
    case B200_OP_REDUCE: return bred_add(x, q, c.brc_hi);                      // :125-145
    case B200_OP_REDUCE_LAZY: return bred_add_lazy(x, q, c.brc_hi);            // :147-167
    case B200_OP_MFORM: return mform(x, q, c.brc_hi, c.brc_lo);                // :778-798
    case B200_OP_MFORM_LAZY: return mform_lazy(x, q, c.brc_hi, c.brc_lo);      // :800-820
    case B200_OP_IMFORM: return imform(x, q, c.qinv);                          // :822-845
    case B200_OP_MUL_BARRETT: return bred(x, y, q, c.brc_hi, c.brc_lo);        // :217-239
    case B200_OP_MUL_BARRETT_LAZY: return bred_lazy(x, y, q, c.brc_hi, c.brc_lo);          // :241-263
    case B200_OP_MUL_BARRETT_ADD: return csub(z + bred(x, y, q, c.brc_hi, c.brc_lo), q);   // :265-287
    case B200_OP_MUL_BARRETT_ADD_LAZY: return z + bred(x, y, q, c.brc_hi, c.brc_lo);       // :289-311
    case B200_OP_MUL_MONT: return mred(x, y, q, c.qinv);                                   // :313-334
    case B200_OP_MUL_MONT_LAZY: return mred_lazy(x, y, q, c.qinv);                         // :336-358
    case B200_OP_MUL_MONT_ADD: return csub(z + mred(x, y, q, c.qinv), q);                  // :360-381
    case B200_OP_MUL_MONT_ADD_LAZY: return z + mred(x, y, q, c.qinv);                      // :383-405
    case B200_OP_MUL_MONT_LAZY_ADD_LAZY: return z + mred_lazy(x, y, q, c.qinv);            // :407-429
    case B200_OP_MUL_MONT_SUB: return csub(z + (q - mred(x, y, q, c.qinv)), q);            // :431-453
    case B200_OP_MUL_MONT_SUB_LAZY: return z + (q - mred(x, y, q, c.qinv));                // :455-477
    case B200_OP_MUL_MONT_LAZY_SUB_LAZY: return z + 2 * q - mred_lazy(x, y, q, c.qinv);    // :479-502
    case B200_OP_MUL_MONT_LAZY_NEG: return 2 * q - mred_lazy(x, y, q, c.qinv);             // :504-527
    case B200_OP_MULSCALAR_MONT: return mred(x, sc, q, c.qinv);                            // :664-684
    case B200_OP_MULSCALAR_MONT_LAZY: return mred_lazy(x, sc, q, c.qinv);                  // :686-706
    case B200_OP_MULSCALAR_MONT_ADD: return csub(z + mred(x, sc, q, c.qinv), q);           // :708-728
    case B200_OP_ADD_SCALAR: return csub(x + sc, q);                                       // :575-595
    case B200_OP_SUB_SCALAR: return csub(x + q - sc, q);                                   // :642-662
    }
    return 0;
}

template <int OP> struct traits {
    static constexpr bool binary = (OP <= B200_OP_SUB_LAZY) || (OP >= B200_OP_MUL_BARRETT && OP <= B200_OP_MUL_MONT_LAZY_NEG);
    static constexpr bool accum = OP == B200_OP_MUL_BARRETT_ADD || OP == B200_OP_MUL_BARRETT_ADD_LAZY ||
                                  OP == B200_OP_MUL_MONT_ADD || OP == B200_OP_MUL_MONT_ADD_LAZY ||
                                  OP == B200_OP_MUL_MONT_LAZY_ADD_LAZY || OP == B200_OP_MUL_MONT_SUB ||
                                  OP == B200_OP_MUL_MONT_SUB_LAZY || OP == B200_OP_MUL_MONT_LAZY_SUB_LAZY ||
                                  OP == B200_OP_MULSCALAR_MONT_ADD;
    static constexpr bool scalar = OP >= B200_OP_MULSCALAR_MONT;
};

// grid = (chunks per limb, units); a unit is one (poly, limb) pair; every thread owns 2 words per step.
template <int OP>
__global__ void __launch_bounds__(256)
k_vec_op(const u64 *p1, const u64 *p2, u64 *p3, const LimbConst *__restrict__ consts,
         const u64 *__restrict__ scalars, int N, Span sp, int unit0) {
    const int unit = unit0 + blockIdx.y;
    const int limb = unit % sp.nl;
    const size_t base = ((size_t)(unit / sp.nl) * sp.Lbuf + limb) * (size_t)N;
    const LimbConst c = consts[sp.limb0 + limb];
    const u64 sc = traits<OP>::scalar ? scalars[limb] : 0;
    const ulonglong2 *a = reinterpret_cast<const ulonglong2 *>(p1 + base);
    const ulonglong2 *b = reinterpret_cast<const ulonglong2 *>(traits<OP>::binary ? p2 + base : p1 + base);
    ulonglong2 *o = reinterpret_cast<ulonglong2 *>(p3 + base);
    const int n2 = N >> 1;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += gridDim.x * blockDim.x) {
        ulonglong2 x = a[i], y = make_ulonglong2(0, 0), z = make_ulonglong2(0, 0);
        if (traits<OP>::binary) y = b[i];
        if (traits<OP>::accum) z = o[i];
        z.x = apply<OP>(x.x, y.x, z.x, sc, c);
        z.y = apply<OP>(x.y, y.y, z.y, sc, c);
        o[i] = z;
    }
}

template <int OP>
int run(b200_ctx *ctx, const u64 *p1, const u64 *p2, u64 *p3, Span sp, const u64 *d_scalars) {
    const int N = ctx->N, n2 = N >> 1;
    int chunks = (n2 + 255) / 256;
    if (chunks > 1024) chunks = 1024;
    const int units = sp.batch * sp.nl;
    for (int u0 = 0; u0 < units; u0 += 65535) {
        int nu = units - u0 < 65535 ? units - u0 : 65535;
        LaunchScope ls(ctx, "vec_op", nu);
        k_vec_op<OP><<<dim3(chunks, nu), 256, 0, b200_cur_stream(ctx)>>>(p1, p2, p3, ctx->d_consts, d_scalars, N, sp, u0);
    }
    CUDA_TRY(cudaGetLastError());
    return 0;
}

}  // namespace

int launch_vec_op(b200_ctx *ctx, int op, const u64 *p1, const u64 *p2, u64 *p3, Span sp, const u64 *d_scalars) {
#define CASE(OP) case OP: return run<OP>(ctx, p1, p2, p3, sp, d_scalars);
    switch (op) {
        CASE(B200_OP_ADD) CASE(B200_OP_ADD_LAZY) CASE(B200_OP_SUB) CASE(B200_OP_SUB_LAZY) CASE(B200_OP_NEG)
        CASE(B200_OP_REDUCE) CASE(B200_OP_REDUCE_LAZY) CASE(B200_OP_MFORM) CASE(B200_OP_MFORM_LAZY) CASE(B200_OP_IMFORM)
        CASE(B200_OP_MUL_BARRETT) CASE(B200_OP_MUL_BARRETT_LAZY) CASE(B200_OP_MUL_BARRETT_ADD) CASE(B200_OP_MUL_BARRETT_ADD_LAZY)
        CASE(B200_OP_MUL_MONT) CASE(B200_OP_MUL_MONT_LAZY) CASE(B200_OP_MUL_MONT_ADD) CASE(B200_OP_MUL_MONT_ADD_LAZY)
        CASE(B200_OP_MUL_MONT_LAZY_ADD_LAZY) CASE(B200_OP_MUL_MONT_SUB) CASE(B200_OP_MUL_MONT_SUB_LAZY)
        CASE(B200_OP_MUL_MONT_LAZY_SUB_LAZY) CASE(B200_OP_MUL_MONT_LAZY_NEG)
        CASE(B200_OP_MULSCALAR_MONT) CASE(B200_OP_MULSCALAR_MONT_LAZY) CASE(B200_OP_MULSCALAR_MONT_ADD)
        CASE(B200_OP_ADD_SCALAR) CASE(B200_OP_SUB_SCALAR)
    }
#undef CASE
    b200_set_error("b200_vec_op: unknown op code " + std::to_string(op));
    return B200_EINVAL;
}
