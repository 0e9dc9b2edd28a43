// ewise.cu -- the element-wise operator family (ring/vec_ops.go, ring/operations.go) as ONE templated
// streaming kernel: all batch x (level+1) limbs in a single launch, 128-bit coalesced loads/stores,
// per-limb constants fetched once per thread.  HBM-bound: 16-32 B/element, ~15-40 integer ops.
#include "internal.hpp"

namespace {

template <int OP>
__device__ __forceinline__ u64 apply(u64 x, u64 y, u64 z, u64 sc, u64 sc2, const LimbConst &c) {
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
    case B200_OP_SUB_MULSCALAR_MONT_2Q: return mred(2 * q - y + x, sc, q, c.qinv);         // :752-776
    case B200_OP_ZERO: return 0;                                                           // :847-868
    case B200_OP_MUL_LAZY: return x * y;                                                   // :169-191
    case B200_OP_MUL_LAZY_ADD_LAZY: return z + x * y;                                      // :193-215
    case B200_OP_ADD_LAZY_MULSCALAR_MONT: return mred(x + y, sc, q, c.qinv);               // :529-551
    case B200_OP_ADD_SCALAR_LAZY: return x + sc;                                           // :597-617
    case B200_OP_ADD_SCALAR_LAZY_NEG_2Q_LAZY: return sc + 2 * q - x;                       // :619-640
    case B200_OP_ADD_SCALAR_LAZY_MULSCALAR_MONT: return mred(x + sc, sc2, q, c.qinv);      // :553-573
    case B200_OP_MULSCALAR_MONT_ADD_SCALAR: return csub(mred(x, sc2, q, c.qinv) + sc, q);  // :730-750
    case B200_OP_MASK: return sc >= 64 ? 0 : (x >> sc) & sc2;                              // :870-890 (a Go shift by >= 64 gives 0)
    case B200_OP_ADD_SCALAR_DOUBLE: return csub(x + sc, q);                                // operations.go:167-173
    case B200_OP_SUB_SCALAR_DOUBLE: return csub(x + q - sc, q);                            // :177-183
    case B200_OP_MULSCALAR_MONT_DOUBLE: return mred(x, sc, q, c.qinv);                     // :250-256
    case B200_OP_MULSCALAR_MONT_ADD_DOUBLE: return csub(z + mred(x, sc, q, c.qinv), q);    // :260-266
    }
    return 0;
}

template <int OP> struct traits {
    static constexpr bool binary = (OP <= B200_OP_SUB_LAZY) || (OP >= B200_OP_MUL_BARRETT && OP <= B200_OP_MUL_MONT_LAZY_NEG) ||
                                   OP == B200_OP_SUB_MULSCALAR_MONT_2Q || OP == B200_OP_ADD_LAZY_MULSCALAR_MONT;
    static constexpr bool accum = OP == B200_OP_MUL_BARRETT_ADD || OP == B200_OP_MUL_BARRETT_ADD_LAZY ||
                                  OP == B200_OP_MUL_MONT_ADD || OP == B200_OP_MUL_MONT_ADD_LAZY ||
                                  OP == B200_OP_MUL_MONT_LAZY_ADD_LAZY || OP == B200_OP_MUL_MONT_SUB ||
                                  OP == B200_OP_MUL_MONT_SUB_LAZY || OP == B200_OP_MUL_MONT_LAZY_SUB_LAZY ||
                                  OP == B200_OP_MULSCALAR_MONT_ADD || OP == B200_OP_MUL_LAZY_ADD_LAZY ||
                                  OP == B200_OP_MULSCALAR_MONT_ADD_DOUBLE;
    static constexpr bool scalar = OP >= B200_OP_MULSCALAR_MONT;
    static constexpr bool scalar2 = OP >= B200_OP_ADD_SCALAR_LAZY_MULSCALAR_MONT;   // a second scalar per limb at scalars[nl + limb]
    static constexpr bool dbl = OP >= B200_OP_ADD_SCALAR_DOUBLE;                    // ... used for the second half of the limb
    static constexpr bool nosrc = OP == B200_OP_ZERO;
};

// grid = (chunks per limb, units); a unit is one (poly, limb) pair; every thread owns 2 words per step.
template <int OP, bool BCAST = false>   // BCAST: p2 is one N-word vector shared by every unit (MulByVectorMontgomery)
__global__ void __launch_bounds__(256)
k_vec_op(const u64 *p1, const u64 *p2, u64 *p3, const LimbConst *__restrict__ consts,
         const u64 *__restrict__ scalars, int N, Span sp, int unit0) {
    const int unit = unit0 + blockIdx.y;
    const int limb = unit % sp.nl;
    const size_t base = ((size_t)(unit / sp.nl) * sp.Lbuf + limb) * (size_t)N;
    const LimbConst c = consts[sp.limb0 + limb];
    const u64 sc = traits<OP>::scalar ? scalars[limb] : 0;
    const u64 sc2 = traits<OP>::scalar2 ? scalars[sp.nl + limb] : 0;
    const ulonglong2 *a = reinterpret_cast<const ulonglong2 *>(p1 + base);
    const ulonglong2 *b = reinterpret_cast<const ulonglong2 *>(traits<OP>::binary ? (BCAST ? p2 : p2 + base) : p1 + base);
    ulonglong2 *o = reinterpret_cast<ulonglong2 *>(p3 + base);
    const int n2 = N >> 1;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += gridDim.x * blockDim.x) {
        ulonglong2 x = make_ulonglong2(0, 0), y = make_ulonglong2(0, 0), z = make_ulonglong2(0, 0);
        if (!traits<OP>::nosrc) x = a[i];
        if (traits<OP>::binary) y = b[i];
        if (traits<OP>::accum) z = o[i];
        const bool second = traits<OP>::dbl && 2 * i >= n2;   // word 2i in [N/2, N)
        const u64 s1 = second ? sc2 : sc;
        z.x = apply<OP>(x.x, y.x, z.x, s1, sc2, c);
        z.y = apply<OP>(x.y, y.y, z.y, s1, sc2, c);
        o[i] = z;
    }
}

template <int OP, bool BCAST = false>
int run(b200_ctx *ctx, const u64 *p1, const u64 *p2, u64 *p3, Span sp, const u64 *d_scalars) {
    const int N = ctx->N, n2 = N >> 1;
    int chunks = (n2 + 255) / 256;
    if (chunks > 1024) chunks = 1024;
    const int units = sp.batch * sp.nl;
    for (int u0 = 0; u0 < units; u0 += 65535) {
        int nu = units - u0 < 65535 ? units - u0 : 65535;
        LaunchScope ls(ctx, "vec_op", nu);
        k_vec_op<OP, BCAST><<<dim3(chunks, nu), 256, 0, b200_cur_stream(ctx)>>>(p1, p2, p3, ctx->d_consts, d_scalars, N, sp, u0);
    }
    CUDA_TRY(cudaGetLastError());
    return 0;
}

// Fused degree-1 x degree-1 tensor product in the NTT domain: one pass over 4 inputs and 3 outputs (56 B/word)
// instead of the 2 MForm + 3 MulCoeffsMontgomery + 1 ...ThenAdd passes (136 B/word) of
// schemes/ckks/evaluator.go:821-834 / schemes/matrix_ckks/evaluator.go:163-173.  Results are the canonical
// words those sequences produce.
template <bool MFORM>
__global__ void __launch_bounds__(256)
k_tensor(const u64 *a0, const u64 *a1, const u64 *b0, const u64 *b1, u64 *c0, u64 *c1, u64 *c2,
         const LimbConst *__restrict__ consts, int N, Span sp, int unit0) {
    const int unit = unit0 + blockIdx.y;
    const int limb = unit % sp.nl;
    const size_t base = ((size_t)(unit / sp.nl) * sp.Lbuf + limb) * (size_t)N;
    const LimbConst c = consts[sp.limb0 + limb];
    const int n2 = N >> 1;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += gridDim.x * blockDim.x) {
        ulonglong2 x0 = reinterpret_cast<const ulonglong2 *>(a0 + base)[i], x1 = reinterpret_cast<const ulonglong2 *>(a1 + base)[i];
        const ulonglong2 y0 = reinterpret_cast<const ulonglong2 *>(b0 + base)[i], y1 = reinterpret_cast<const ulonglong2 *>(b1 + base)[i];
        if (MFORM) {
            x0.x = mform(x0.x, c.q, c.brc_hi, c.brc_lo); x0.y = mform(x0.y, c.q, c.brc_hi, c.brc_lo);
            x1.x = mform(x1.x, c.q, c.brc_hi, c.brc_lo); x1.y = mform(x1.y, c.q, c.brc_hi, c.brc_lo);
        }
        ulonglong2 r0, r1, r2;
        r0.x = mred(x0.x, y0.x, c.q, c.qinv); r0.y = mred(x0.y, y0.y, c.q, c.qinv);
        r2.x = mred(x1.x, y1.x, c.q, c.qinv); r2.y = mred(x1.y, y1.y, c.q, c.qinv);
        r1.x = csub(mred(x0.x, y1.x, c.q, c.qinv) + mred(x1.x, y0.x, c.q, c.qinv), c.q);
        r1.y = csub(mred(x0.y, y1.y, c.q, c.qinv) + mred(x1.y, y0.y, c.q, c.qinv), c.q);
        reinterpret_cast<ulonglong2 *>(c0 + base)[i] = r0;
        reinterpret_cast<ulonglong2 *>(c1 + base)[i] = r1;
        reinterpret_cast<ulonglong2 *>(c2 + base)[i] = r2;
    }
}

// Folding layer of the conjugate-invariant transform (ring/ntt.go:764-770 forward, :1149-1157 inverse): with F the
// fourth root of unity RootsForward[1] (inverse: RootsBackward[1])
//   out[j] = in[j] - F*in[N-j] for 0 < j < N (j and N-j are handled by the same thread, so in == out is allowed),
//   out[0] = in[0] (inverse: 2*in[0]).
// Inputs < 2q, outputs canonical.  HBM-bound: 16 B/word, one Shoup multiply per word.
__global__ void __launch_bounds__(256)
k_ci_fold(const u64 *in, u64 *out, const LimbConst *__restrict__ consts, const tw_t *__restrict__ misc, int N, Span sp, int unit0, int inverse) {
    const int unit = unit0 + blockIdx.y;
    const int limb = unit % sp.nl;
    const size_t base = ((size_t)(unit / sp.nl) * sp.Lbuf + limb) * (size_t)N;
    const u64 q = consts[sp.limb0 + limb].q;
    const tw_t F = misc[(sp.limb0 + limb) * MISC_COUNT + (inverse ? MISC_CI_FINV : MISC_CI_F)];
    const u64 *a = in + base;
    u64 *o = out + base;
    const int half = N >> 1;
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j <= half; j += gridDim.x * blockDim.x) {
        if (j == 0) {
            const u64 x = csub(a[0], q);
            o[0] = inverse ? csub(x << 1, q) : x;                     // :770 / :1157 CRed(p2[0] << 1)
        } else {
            const u64 x = csub(a[j], q), y = csub(a[N - j], q);
            const u64 fx = shoup(x, F.x, F.y, q), fy = shoup(y, F.x, F.y, q);
            o[j] = csub(x + q - fy, q);
            if (j != half) o[N - j] = csub(y + q - fx, q);
        }
    }
}

// Ring.UnfoldConjugateInvariantToStandard (ring/conjugate_invariant.go:7-23): the compressed N words of every limb to
// the 2N words of the standard ring, out[j] = in[j], out[2N-1-j] = in[j].
__global__ void __launch_bounds__(256)
k_ci_unfold(const u64 *in, u64 *out, int N, int Lin, int Lout, int nl) {
    const int unit = blockIdx.y, limb = unit % nl, b = unit / nl;
    const u64 *a = in + ((size_t)b * Lin + limb) * (size_t)N;
    u64 *o = out + ((size_t)b * Lout + limb) * (size_t)(2 * N);
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < N; j += gridDim.x * blockDim.x) {
        const u64 v = a[j];
        o[j] = v;
        o[2 * N - 1 - j] = v;
    }
}

// Automorphisms of power-of-two rings (ring/automorphism.go).  NTT domain: a gather through the table of
// AutomorphismNTTIndex (:12-35), computed on the fly when no table is given:
//   index[i] = brev(((gen*(2*brev(i)+1) & (NthRoot-1)) - 1) >> 1)      (brev over log2(NthRoot)-1 bits)
// mode 0: out[j] = in[index[j]]  (AutomorphismNTTWithIndex :50-79);  mode 1: out[j] += in[index[j]]  (...ThenAddLazy :82-110);
// mode 2: out[j] = in[index[j]] + in[j] mod q  (FoldStandardToConjugateInvariant, conjugate_invariant.go:28-44: `in` is a
// polynomial of the standard ring of degree 2N, `out` one of the conjugate-invariant ring of degree N).
struct AutoArgs {
    const u64 *index;      // device table or nullptr
    u64 gen, mask;
    int lognth;
    int n_out, n_in, l_out, l_in, nl, mode;
};
__device__ __forceinline__ unsigned brev_bits(unsigned x, int bits) { return bits ? (__brev(x) >> (32 - bits)) : 0u; }
__global__ void __launch_bounds__(256)
k_automorphism_ntt(const u64 *in, u64 *out, const LimbConst *__restrict__ consts, AutoArgs a, int unit0) {
    const int unit = unit0 + blockIdx.y, limb = unit % a.nl, b = unit / a.nl;
    const u64 *y = in + ((size_t)b * a.l_in + limb) * (size_t)a.n_in;
    u64 *z = out + ((size_t)b * a.l_out + limb) * (size_t)a.n_out;
    const u64 q = consts[limb].q;
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < a.n_out; j += gridDim.x * blockDim.x) {
        u64 x;
        if (a.index) x = a.index[j];
        else {
            const u64 t1 = 2 * (u64)brev_bits((unsigned)j, a.lognth) + 1;
            const u64 t2 = (((a.gen * t1) & a.mask) - 1) >> 1;
            x = brev_bits((unsigned)t2, a.lognth);
        }
        const u64 v = y[x];
        if (a.mode == 0) z[j] = v;
        else if (a.mode == 1) z[j] += v;
        else z[j] = csub(v + y[j], q);
    }
}
// Coefficient domain, Ring.Automorphism (:113-176): out[(i*gen) mod N] = +-in[i]; the conjugate-invariant branch
// (:124-151) walks the 2N exponents of the unfolded polynomial.  A negated zero is written as q, like the Go expression.
__global__ void __launch_bounds__(256)
k_automorphism_coeff(const u64 *in, u64 *out, const LimbConst *__restrict__ consts, int N, int L, int nl, u64 gen, int ci, int unit0) {
    const int unit = unit0 + blockIdx.y, limb = unit % nl, b = unit / nl;
    const size_t base = ((size_t)b * L + limb) * (size_t)N;
    const u64 q = consts[limb].q, n = (u64)N;
    const u64 span = ci ? 2 * n : n, mask = span - 1;
    const int logn = 64 - __clzll(mask);                                   // bits.Len64(mask)
    for (u64 i = blockIdx.x * blockDim.x + threadIdx.x; i < span; i += (u64)gridDim.x * blockDim.x) {
        const u64 raw = i * gen, index = raw & mask;
        u64 tmp = (raw >> logn) & 1, idx = i;
        if (ci) {
            if (index >= n) continue;
            if (idx >= n) { idx = 2 * n - idx; tmp ^= 1; }
        }
        const u64 v = in[base + idx];
        out[base + index] = tmp ? q - v : v;
    }
}

// Ring.Shift (ring/operations.go:278-282, utils.RotateSliceAllocFree): out[j] = in[(j + k) mod N], every limb.
// Ring.MultByMonomial (:306-363): multiplication by X^k with X^N = -1, words exactly as the Go loops leave them
// (a negated zero is q):  shift = (k + 2N) mod 2N, s = shift mod N, neg = shift >= N:
//   j <  s: out[j] = neg ? in[N-s+j] : q - in[N-s+j];      j >= s: out[j] = neg ? q - in[j-s] : in[j-s].
// mode 0 = Shift, 1 = MultByMonomial.  Not in place (the launcher stages aliased operands).
__global__ void __launch_bounds__(256)
k_rotate(const u64 *in, u64 *out, const LimbConst *__restrict__ consts, int N, int L, int nl, int s, int neg, int mode, int unit0) {
    const int unit = unit0 + blockIdx.y, limb = unit % nl, b = unit / nl;
    const size_t base = ((size_t)b * L + limb) * (size_t)N;
    const u64 q = consts[limb].q;
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < N; j += gridDim.x * blockDim.x) {
        if (mode == 0) {
            int src = j + s;
            if (src >= N) src -= N;
            out[base + j] = in[base + src];
        } else {
            const bool wrap = j < s;
            const u64 v = in[base + (wrap ? N - s + j : j - s)];
            out[base + j] = (wrap != (neg != 0)) ? q - v : v;
        }
    }
}

// Rescale helper (ring/scaling.go:137-143): x = coefficient-domain last limb [0, q_l); for every lower limb i
// t_i = (x (+ pHalf mod q_l)) + (q_i - pHalf mod q_i), reduced to [0, q_i).  ws is a [batch][Lbuf][N] buffer whose
// limb `level` holds x; limbs 0..level-1 receive t_i.
__global__ void __launch_bounds__(256)
k_rescale_spread(u64 *ws, const LimbConst *__restrict__ consts, int N, int Lbuf, int level, int round) {
    const int b = blockIdx.y;
    const u64 *x = ws + ((size_t)b * Lbuf + level) * (size_t)N;
    const u64 ql = consts[level].q, phalf = (ql - 1) >> 1;
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < N; j += gridDim.x * blockDim.x) {
        u64 v = x[j];
        if (round) v = csub(v + phalf, ql);
        for (int i = 0; i < level; ++i) {
            const LimbConst c = consts[i];
            const u64 add = round ? c.q - bred_add(phalf, c.q, c.brc_hi) : 0;
            ws[((size_t)b * Lbuf + i) * (size_t)N + j] = csub(bred_add(v, c.q, c.brc_hi) + add, c.q);
        }
    }
}

// Ring.DivFloorByLastModulus / DivRoundByLastModulus (ring/scaling.go:21-30,113-127): the coefficient-domain rescale.
// Unit = (polynomial, limb i < level); x is the word of limb `level`, y the word of limb i, rc = RescaleConstants[level-1][i]:
//   floor: MRed(2q_i - y + x, rc)                                                          (SubThenMulScalarMontgomeryTwoModulus)
//   round: MRed(CRed(x + pHalf, q_level) + ((q_i - BRedAdd(pHalf, q_i)) + 2q_i - y), rc)   (AddScalar, AddScalarLazyThenNegTwoModulusLazy,
//                                                                                          AddLazyThenMulScalarMontgomery)
// in == out is allowed: limb `level` is only read.
template <bool ROUND>
__global__ void __launch_bounds__(256)
k_div_last(const u64 *in, u64 *out, const LimbConst *__restrict__ consts, const u64 *__restrict__ rescale, int N, int Lbuf, int level, int unit0) {
    const int unit = unit0 + blockIdx.y, b = unit / level, i = unit - b * level;
    const LimbConst c = consts[i];
    const u64 ql = consts[level].q, phalf = (ql - 1) >> 1, rc = rescale[i], q = c.q;
    const u64 s = ROUND ? q - bred_add(phalf, q, c.brc_hi) : 0;
    const ulonglong2 *xl = reinterpret_cast<const ulonglong2 *>(in + ((size_t)b * Lbuf + level) * (size_t)N);
    const ulonglong2 *yi = reinterpret_cast<const ulonglong2 *>(in + ((size_t)b * Lbuf + i) * (size_t)N);
    ulonglong2 *o = reinterpret_cast<ulonglong2 *>(out + ((size_t)b * Lbuf + i) * (size_t)N);
    const int n2 = N >> 1;
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n2; j += gridDim.x * blockDim.x) {
        ulonglong2 x = xl[j];
        const ulonglong2 y = yi[j];
        ulonglong2 r;
        if (ROUND) {
            x.x = csub(x.x + phalf, ql); x.y = csub(x.y + phalf, ql);
            r.x = mred(x.x + (s + 2 * q - y.x), rc, q, c.qinv);
            r.y = mred(x.y + (s + 2 * q - y.y), rc, q, c.qinv);
        } else {
            r.x = mred(2 * q - y.x + x.x, rc, q, c.qinv);
            r.y = mred(2 * q - y.y + x.y, rc, q, c.qinv);
        }
        o[j] = r;
    }
}

}  // namespace

int launch_ci_fold(b200_ctx *ctx, const u64 *in, u64 *out, Span sp, bool inverse) {
    const int N = ctx->N;
    int chunks = (N / 2 + 1 + 255) / 256;
    if (chunks > 512) chunks = 512;
    const int units = sp.batch * sp.nl;
    for (int u0 = 0; u0 < units; u0 += 65535) {
        const int nu = units - u0 < 65535 ? units - u0 : 65535;
        LaunchScope ls(ctx, inverse ? "ci_fold_inv" : "ci_fold_fwd", nu);
        k_ci_fold<<<dim3(chunks, nu), 256, 0, b200_cur_stream(ctx)>>>(in, out, ctx->d_consts, ctx->d_misc, N, sp, u0, inverse ? 1 : 0);
    }
    CUDA_TRY(cudaGetLastError());
    return 0;
}

int launch_ci_unfold(b200_ctx *ctx, const u64 *in, u64 *out, int N, int batch, int nl, int Lin, int Lout) {
    int chunks = (N + 255) / 256;
    if (chunks > 512) chunks = 512;
    if ((long long)batch * nl > 65535) { b200_set_error("b200_unfold_ci_to_standard: more than 65535 limb units"); return B200_ESHAPE; }
    LaunchScope ls(ctx, "ci_unfold", (long long)batch * nl);
    k_ci_unfold<<<dim3(chunks, batch * nl), 256, 0, b200_cur_stream(ctx)>>>(in, out, N, Lin, Lout, nl);
    CUDA_TRY(cudaGetLastError());
    return 0;
}

int launch_automorphism_ntt(b200_ctx *ctx, const u64 *in, u64 *out, const u64 *d_index, u64 gen, u64 nthroot, int n_in, int l_in,
                            int batch, int nl, int mode) {
    AutoArgs a{};
    a.index = d_index; a.gen = gen; a.mask = nthroot - 1;
    a.lognth = 0;
    while (((u64)1 << (a.lognth + 1)) < nthroot) a.lognth++;
    a.n_out = ctx->N; a.n_in = n_in; a.l_out = ctx->L; a.l_in = l_in; a.nl = nl; a.mode = mode;
    int chunks = (ctx->N + 255) / 256;
    if (chunks > 512) chunks = 512;
    const int units = batch * nl;
    for (int u0 = 0; u0 < units; u0 += 65535) {
        const int nu = units - u0 < 65535 ? units - u0 : 65535;
        LaunchScope ls(ctx, "automorphism_ntt", nu);
        k_automorphism_ntt<<<dim3(chunks, nu), 256, 0, b200_cur_stream(ctx)>>>(in, out, ctx->d_consts, a, u0);
    }
    CUDA_TRY(cudaGetLastError());
    return 0;
}

int launch_automorphism_coeff(b200_ctx *ctx, const u64 *in, u64 *out, u64 gen, int batch, int nl) {
    const int ci = ctx->ring_type == B200_RING_CONJUGATE_INVARIANT ? 1 : 0;
    int chunks = ((ci ? 2 : 1) * ctx->N + 255) / 256;
    if (chunks > 512) chunks = 512;
    const int units = batch * nl;
    for (int u0 = 0; u0 < units; u0 += 65535) {
        const int nu = units - u0 < 65535 ? units - u0 : 65535;
        LaunchScope ls(ctx, "automorphism", nu);
        k_automorphism_coeff<<<dim3(chunks, nu), 256, 0, b200_cur_stream(ctx)>>>(in, out, ctx->d_consts, ctx->N, ctx->L, nl, gen, ci, u0);
    }
    CUDA_TRY(cudaGetLastError());
    return 0;
}

int launch_rotate(b200_ctx *ctx, const u64 *in, u64 *out, int batch, int nl, int s, int neg, int mode) {
    const int N = ctx->N;
    int chunks = (N + 255) / 256;
    if (chunks > 512) chunks = 512;
    const int units = batch * nl;
    for (int u0 = 0; u0 < units; u0 += 65535) {
        const int nu = units - u0 < 65535 ? units - u0 : 65535;
        LaunchScope ls(ctx, mode ? "mult_by_monomial" : "shift", nu);
        k_rotate<<<dim3(chunks, nu), 256, 0, b200_cur_stream(ctx)>>>(in, out, ctx->d_consts, N, ctx->L, nl, s, neg, mode, u0);
    }
    CUDA_TRY(cudaGetLastError());
    return 0;
}

int launch_rescale_spread(b200_ctx *ctx, u64 *ws, int batch, int level, bool round) {
    const int N = ctx->N;
    int chunks = (N + 255) / 256;
    if (chunks > 256) chunks = 256;
    for (int b0 = 0; b0 < batch; b0 += 65535) {
        const int nb = batch - b0 < 65535 ? batch - b0 : 65535;
        LaunchScope ls(ctx, "rescale_spread", (long long)nb * level);
        k_rescale_spread<<<dim3(chunks, nb), 256, 0, b200_cur_stream(ctx)>>>(ws + (size_t)b0 * ctx->L * N, ctx->d_consts, N, ctx->L, level, round ? 1 : 0);
    }
    CUDA_TRY(cudaGetLastError());
    return 0;
}

int launch_div_last(b200_ctx *ctx, const u64 *in, u64 *out, int batch, int level, bool round) {
    const int N = ctx->N, n2 = N >> 1;
    int chunks = (n2 + 255) / 256;
    if (chunks > 1024) chunks = 1024;
    const long long units = (long long)batch * level;
    const u64 *rescale = ctx->d_rescale + (size_t)(level - 1) * ctx->L;
    for (long long u0 = 0; u0 < units; u0 += 65535) {
        const int nu = (int)(units - u0 < 65535 ? units - u0 : 65535);
        LaunchScope ls(ctx, "div_last", nu);
        if (round) k_div_last<true><<<dim3(chunks, nu), 256, 0, b200_cur_stream(ctx)>>>(in, out, ctx->d_consts, rescale, N, ctx->L, level, (int)u0);
        else k_div_last<false><<<dim3(chunks, nu), 256, 0, b200_cur_stream(ctx)>>>(in, out, ctx->d_consts, rescale, N, ctx->L, level, (int)u0);
    }
    CUDA_TRY(cudaGetLastError());
    return 0;
}

int launch_tensor(b200_ctx *ctx, const u64 *a0, const u64 *a1, const u64 *b0, const u64 *b1, u64 *c0, u64 *c1, u64 *c2,
                  Span sp, bool mform_first) {
    const int N = ctx->N, n2 = N >> 1;
    int chunks = (n2 + 255) / 256;
    if (chunks > 1024) chunks = 1024;
    const int units = sp.batch * sp.nl;
    for (int u0 = 0; u0 < units; u0 += 65535) {
        const int nu = units - u0 < 65535 ? units - u0 : 65535;
        LaunchScope ls(ctx, "tensor", nu);
        if (mform_first) k_tensor<true><<<dim3(chunks, nu), 256, 0, b200_cur_stream(ctx)>>>(a0, a1, b0, b1, c0, c1, c2, ctx->d_consts, N, sp, u0);
        else k_tensor<false><<<dim3(chunks, nu), 256, 0, b200_cur_stream(ctx)>>>(a0, a1, b0, b1, c0, c1, c2, ctx->d_consts, N, sp, u0);
    }
    CUDA_TRY(cudaGetLastError());
    return 0;
}

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
        CASE(B200_OP_ADD_SCALAR) CASE(B200_OP_SUB_SCALAR) CASE(B200_OP_SUB_MULSCALAR_MONT_2Q)
        CASE(B200_OP_ZERO) CASE(B200_OP_MUL_LAZY) CASE(B200_OP_MUL_LAZY_ADD_LAZY) CASE(B200_OP_ADD_LAZY_MULSCALAR_MONT)
        CASE(B200_OP_ADD_SCALAR_LAZY) CASE(B200_OP_ADD_SCALAR_LAZY_NEG_2Q_LAZY) CASE(B200_OP_ADD_SCALAR_LAZY_MULSCALAR_MONT)
        CASE(B200_OP_MULSCALAR_MONT_ADD_SCALAR) CASE(B200_OP_MASK) CASE(B200_OP_ADD_SCALAR_DOUBLE) CASE(B200_OP_SUB_SCALAR_DOUBLE)
        CASE(B200_OP_MULSCALAR_MONT_DOUBLE) CASE(B200_OP_MULSCALAR_MONT_ADD_DOUBLE)
    }
#undef CASE
    b200_set_error("b200_vec_op: unknown op code " + std::to_string(op));
    return B200_EINVAL;
}

// Ring.MulByVectorMontgomery[ThenAddLazy] (ring/operations.go:366-377): p2 of the kernel is one vector for all units
int launch_mul_by_vector(b200_ctx *ctx, const u64 *p1, const u64 *vector, u64 *p3, Span sp, bool then_add_lazy) {
    return then_add_lazy ? run<B200_OP_MUL_MONT_ADD_LAZY, true>(ctx, p1, vector, p3, sp, nullptr)
                         : run<B200_OP_MUL_MONT, true>(ctx, p1, vector, p3, sp, nullptr);
}
