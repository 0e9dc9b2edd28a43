// ntt_big.cu -- the tuned two-pass transforms for large rings (the headline path).
//
// A limb of N words (512 KiB at N = 2^16, 768 KiB at N = 3*2^15) does not fit in one SM, so the
// transform is split over the index i = row*256 + c:
//   * column pass ("cols"): every layer that couples different rows (STANDARD: the first log2(N)-8
//     radix-2 layers; MATRIX: the cyclotomic layer, the radix-3 layer and the first a-9 radix-2
//     layers).  A CTA owns 16 adjacent columns of ALL rows (128-byte row segments => coalesced),
//     runs the layers in two register-resident groups (up to 24 words per thread) with one
//     shared-memory exchange in between.  The few twiddles it needs are the same for every column.
//   * row pass ("rows"): the last 8 radix-2 layers act inside contiguous 256-word rows.  A CTA owns
//     16 (24 for MATRIX) rows, two register groups of 4 layers with one swizzled shared-memory
//     exchange, 256-bit global loads/stores.  For MATRIX rings the row pass also converts between
//     the tree order and the reference slot order k = 6*bitrev(i_low) + 2d + h (slot k <-> w^E[k],
//     ring/ntt_3n.go:46-53) through a staged transposition so that global accesses stay in
//     192-byte runs.
// Every butterfly is Harvey-lazy on Shoup twiddles (one mul.hi + two mul.lo on the integer pipes);
// the twiddle of the sub-transform rooted at block [i0, i0+S) of a layer is table[(N + i0)/S]
// (heap numbering), for both ring types.  Algorithmic HBM traffic is 16 B/word per pass; the
// launcher processes the batch in chunks so the intermediate of a chunk is still in L2 when the
// second pass reads it.
#include <cstdlib>

#include "internal.hpp"

// tuning knobs (compile-time; `build.py --tag=X -DKNOB=v` builds lib/libb200ring_X.so for A/B timing, selected with
// B200RING_LIB).  Occupancy beats ILP on these kernels: the minBlocks values cap registers at 40-56 per thread
// (measured on B200: +45% on the 3N path over letting ptxas use 80-120 registers).
#ifndef B200_CSUB_HI
#define B200_CSUB_HI 0      // 1: compare high words only in the lazy conditional subtraction
#endif
// Conditional subtraction of 4q inside the lazy butterflies, chosen per kernel family (each measured on B200):
//   0 = multiply-add by 2^64-4q on the multiply pipe (compare + select + wide multiply-add)
//   1 = compare / select / subtract on the ALU pipe (6 instructions)
//   2 = subtract, then add 4q back if the result is negative: one sign test + two predicated adds; the subtraction folds
//       into the three-input add that produces the value (inverse butterflies: 5 instead of 8 instructions)
// Mode 2 against the previous choice: inverse rows -3.3 % (3N) / -6.1 % (2^16), inverse columns -0.5 %, forward rows
// -1.5 % / -0.6 %, forward columns 2^16 -7.9 %, forward columns 3N +4.9 % (stays on mode 1).
#ifndef B200_CSUB_FWD_ROWS
#define B200_CSUB_FWD_ROWS 2
#endif
#ifndef B200_CSUB_FWD_COLS_STD
#define B200_CSUB_FWD_COLS_STD 2
#endif
#ifndef B200_CSUB_FWD_COLS_M3
#define B200_CSUB_FWD_COLS_M3 1
#endif
#ifndef B200_CSUB_INV_ROWS
#define B200_CSUB_INV_ROWS 2
#endif
#ifndef B200_CSUB_INV_COLS
#define B200_CSUB_INV_COLS 2
#endif
#ifndef B200_R3INV_APPROX
#define B200_R3INV_APPROX 1   // inverse radix-3 layer of the 3N column kernels on approximate-quotient products (-2.3 % on that kernel)
#endif
#ifndef B200_CYCINV_APPROX
#define B200_CYCINV_APPROX 0   // inverse cyclotomic layer on approximate-quotient products + quotient-estimate reduction: measured 3 % slower
#endif
// products of the cyclotomic / radix-3 / N^-1 layers that must land in [0,2q): 1 = approximate quotient + sign-based
// add-back (shoup2q, ~22 instructions), 0 = exact 64-bit quotient (shoup_lazy, ~30).  Measured on B200: fewer instructions
// but the register-capped column kernels spill (forward 3N columns +40 %, inverse +3-5 %, 2^16 inverse +1 %): off.
#ifndef B200_SHOUP2Q_INV
#define B200_SHOUP2Q_INV 0
#endif
#ifndef B200_SHOUP2Q_FWD
#define B200_SHOUP2Q_FWD 0
#endif
#ifndef B200_CANON_EST
#define B200_CANON_EST 1    // 1: final canonical reduction through a quotient estimate
#endif
// last conditional subtraction of the canonical reduction in the forward row pass (modes of canon_est)
#ifndef B200_CANON_FINAL_M3
#define B200_CANON_FINAL_M3 2
#endif
#ifndef B200_CANON_FINAL_STD
#define B200_CANON_FINAL_STD 2
#endif
// columns per CTA in the hand-tuned column kernels (the generic family uses 8).  Measured on B200: all four tuned
// kernels gain 2-6% from 4-column tiles (192-thread CTAs, shorter barrier waits).
#ifndef B200_CW_M3_FWD
#define B200_CW_M3_FWD 4
#endif
#ifndef B200_CW_M3_INV
#define B200_CW_M3_INV 4
#endif
// the GATHER instantiations also store to peer GPUs: wider tiles make longer contiguous NVLink writes
#ifndef B200_CW_BAL
#define B200_CW_BAL 4     // columns per CTA of the balanced fold column kernels (env B200_CW_BAL: 4, 8 or 16)
#endif
#ifndef B200_CW_GATHER
#define B200_CW_GATHER 16
#endif
#ifndef B200_CW_STD_FWD
#define B200_CW_STD_FWD 4
#endif
#ifndef B200_CW_STD_INV
#define B200_CW_STD_INV 4
#endif
#ifndef B200_VT_STD
#define B200_VT_STD 4         // rows (warps) per power-of-two row CTA; R >= 4 for every covered shape
#endif
#ifndef B200_VT_M3
#define B200_VT_M3 2          // v-tiles (6 rows each) per 3N row CTA
#endif
#ifndef B200_MINB_ROWS_M3
#define B200_MINB_ROWS_M3 4   // CTAs per SM at B200_VT_M3 = 2 (scaled by 2/VT)
#endif
// CTAs per SM at 8 columns per CTA (scaled by 8/CW for narrower tiles), i.e. the register cap: forward 40, inverse 56
#ifndef B200_MINB_COLS_M3_FWD
#define B200_MINB_COLS_M3_FWD 4
#endif
#ifndef B200_MINB_COLS_M3_INV
#define B200_MINB_COLS_M3_INV 3
#endif
#ifndef B200_MINB_ROWS_STD_FWD
#define B200_MINB_ROWS_STD_FWD 8    // 4-warp CTAs per SM: forward 64 registers, inverse 40
#endif
#ifndef B200_MINB_ROWS_STD_INV
#define B200_MINB_ROWS_STD_INV 12
#endif
#ifndef B200_MINB_ROWS_MUL
#define B200_MINB_ROWS_MUL 6
#endif
#ifndef B200_MINB_COLS_STD
#define B200_MINB_COLS_STD 6
#endif

namespace {

struct BigArgs {
    const LimbConst *consts;
    const tw_t *fwd, *inv, *misc;
    int N, R, logU;   // R = N/256 rows; MATRIX: U = R/6 = 2^logU rows per block
    Span sp;
    int unit0, flags;
    int in_dense, out_dense;   // the pointer is a dense [chunk][N] ring buffer indexed by blockIdx.y
    PeerOut peers;             // GATHER kernels only: mirrors of `out` on other GPUs
    u64 div_magic;             // ceil(2^64 / batch): unit / batch = hi64(unit * div_magic) for unit, batch < 2^31; 0 = batch is 1
    // fold arithmetic (ntt_fold.cuh): per-prime constants and tables
    const SolConst *sol;
    const stw_t *fwd_s, *inv_s, *misc_s;
    const tw_t *fwd_r, *inv_r;   // copies of fwd / inv with the entries the row pass reads in row-pass order (b200ring.cu)
};

// word offsets of this CTA's unit in the input and output buffers, and its prime
struct UnitRef { size_t in, out; int limb; };
// Units are numbered LIMB-MAJOR (unit = limb*batch + poly) and sit on blockIdx.x, the tile on blockIdx.y:
// CTAs that are scheduled together work on the same tile of the same prime for consecutive polynomials,
// so they share one small set of twiddles (L1/L2 hits) instead of spreading over all L tables.
// The column pass swaps the roles (COLS: tile on blockIdx.x, unit on blockIdx.y): every column tile of a unit uses the
// same twiddles, and neighbouring tiles share DRAM bursts when they run together (-8 % on the forward 3N kernel).
template <bool COLS = false>
__device__ __forceinline__ UnitRef unit_ref(const BigArgs &a) {
    const int u = COLS ? blockIdx.y : blockIdx.x;
    const int unit = a.unit0 + u;
    // a runtime integer division costs ~30 instructions per thread (reciprocal + fix-ups); the host-prepared multiplier, 4
    const int l = a.div_magic ? (int)__umul64hi((u64)(unsigned)unit, a.div_magic) : unit, b = unit - l * a.sp.batch;
    const size_t layout = ((size_t)b * a.sp.Lbuf + l) * (size_t)a.N;
    const size_t dense = (size_t)u * (size_t)a.N;
    return {a.in_dense ? dense : layout, a.out_dense ? dense : layout, a.sp.limb0 + l};
}

// Per-prime constants of the lazy butterflies
struct QC {
    u64 q, twoq, fourq, nq, nfourq;
    unsigned csh, cmul;
    __device__ __forceinline__ explicit QC(const LimbConst &l)
        : q(l.q), twoq(2 * l.q), fourq(4 * l.q), nq(0 - l.q), nfourq(0 - 4 * l.q), csh(l.canon_shift), cmul(l.canon_mul) {}
};
template <bool ALU>
__device__ __forceinline__ u64 csub4(u64 x, const QC &c) {
#if B200_CSUB_HI
    return csub_fma_hi(x, c.fourq, c.nfourq);
#else
    return ALU ? csub(x, c.fourq) : csub_fma(x, c.fourq, c.nfourq);
#endif
}
// Harvey butterflies on the approximate Shoup product (shoup4: [0,4q)).  q < 2^61 so 8q fits in 64 bits.
//   forward (Cooley-Tukey):   in [0,8q) -> out [0,8q)
//   inverse (Gentleman-Sande): in [0,4q) -> out [0,4q)
template <int MODE>
__device__ __forceinline__ void bf_fwd(u64 &U, u64 &V, const tw_t w, const QC &c) {
    const u64 u = MODE == 2 ? addback_neg(U + c.nfourq, c.fourq) : csub4<MODE == 1>(U, c);
    const u64 t = shoup4(V, w.x, w.y, c.nq);
    U = u + t;
    V = u - t + c.fourq;
}
template <int MODE>
__device__ __forceinline__ void bf_inv(u64 &U, u64 &V, const tw_t w, const QC &c) {
    const u64 d = U - V + c.fourq;
    if (MODE == 2) U = addback_neg(U + V + c.nfourq, c.fourq);
    else U = csub4<MODE == 1>(U + V, c);
    V = shoup4(d, w.x, w.y, c.nq);
}
// [0,8q) -> [0,q) and [0,4q) -> [0,2q)
template <int FINAL = 0>
__device__ __forceinline__ u64 canon8(u64 x, const QC &c) {
#if B200_CANON_EST
    return canon_est<FINAL>(x, c.q, c.nq, c.csh, c.cmul);
#else
    return csub(csub(csub(csub(x, c.fourq), c.twoq), c.q), c.q);
#endif
}
// Conditional subtraction in add-back form for a value the caller already reduced by m (sum_minus_m in [-m, m)).  Used for
// the last step of the power-of-two inverse column pass (-6 % on that kernel); the cyclotomic / radix-3 layers of the 3N
// column kernels keep the compare/select form: converting them measured 8 % (forward) and 2 % (inverse) slower.
__device__ __forceinline__ u64 red_ab(u64 sum_minus_m, u64 m) { return addback_neg(sum_minus_m, m); }
__device__ __forceinline__ u64 to2q(u64 x, const QC &c) { return csub(x, c.twoq); }
template <bool APPROX>
__device__ __forceinline__ u64 mul2q(u64 x, const tw_t w, const QC &c) {
    return APPROX ? shoup2q(x, w.x, w.y, c.nq, c.twoq) : shoup_lazy(x, w.x, w.y, c.q);
}

// 2^LOGM-point sub-transform held in registers; x[k] ordered along the sub-dimension (MSB first);
// P = heap index of the sub-transform's root node.
// Inside a register group the role of every word in the next layer (kept operand U or multiplied operand V) is known at
// compile time.  With B200_FWD_SIGNED_CHAIN a word that will be the U operand of the next layer is produced as value - 4q
// (a free change of the constant in the three-input add), so that layer's conditional subtraction is only the sign test
// and the two predicated adds (3 instructions instead of 5); the group's first layer takes unsigned words, its last layer
// leaves unsigned words in [0,8q), so nothing changes outside the group.
#ifndef B200_FWD_SIGNED_CHAIN
#define B200_FWD_SIGNED_CHAIN 1
#endif
// Where a register group finds the twiddle of level j, block g: heap order (table[(P << j) + g], the column pass and the
// generic kernels) or the row-pass order of the second table copy (b200ring.cu row_pass_order: base[((1 << j) - 1 + g) * STRIDE],
// consecutive lanes of a warp read consecutive entries).
struct TwHeap {
    const tw_t *__restrict__ tw; unsigned P;
    __device__ __forceinline__ tw_t operator()(int j, int g) const { return tw[(P << j) + g]; }
};
template <int STRIDE>
struct TwRow {
    const tw_t *__restrict__ base;
    __device__ __forceinline__ tw_t operator()(int j, int g) const { return base[((1 << j) - 1 + g) * STRIDE]; }
};
template <int LOGM, int MODE, class TW>
__device__ __forceinline__ void fwd_regs_t(u64 (&x)[1 << LOGM], const TW tw, const QC &c) {
#pragma unroll
    for (int j = 0; j < LOGM; ++j) {
        const int half = (1 << LOGM) >> (j + 1);
#pragma unroll
        for (int g = 0; g < (1 << j); ++g) {
            const tw_t w = tw(j, g);
#pragma unroll
            for (int o = 0; o < half; ++o) {
                if (MODE == 2 && B200_FWD_SIGNED_CHAIN) {
                    u64 &U = x[g * 2 * half + o], &V = x[g * 2 * half + o + half];
                    const int i0 = g * 2 * half + o, i1 = i0 + half, nh = half >> 1;
                    const bool last = j + 1 >= LOGM;
                    const bool u_next_is_u = !last && ((i0 / nh) & 1) == 0, v_next_is_u = !last && ((i1 / nh) & 1) == 0;
                    const u64 u = addback_neg(j == 0 ? U + c.nfourq : U, c.fourq);      // layers >= 1 receive value - 4q
                    const u64 t = shoup4(V, w.x, w.y, c.nq);
                    U = u_next_is_u ? u + t + c.nfourq : u + t;
                    V = v_next_is_u ? u - t : u - t + c.fourq;
                } else {
                    bf_fwd<MODE>(x[g * 2 * half + o], x[g * 2 * half + o + half], w, c);
                }
            }
        }
    }
}
template <int LOGM, int MODE = B200_CSUB_FWD_ROWS>
__device__ __forceinline__ void fwd_regs(u64 (&x)[1 << LOGM], unsigned P, const tw_t *__restrict__ tw, const QC &c) {
    fwd_regs_t<LOGM, MODE>(x, TwHeap{tw, P}, c);
}
template <int LOGM, int JMIN, int MODE, class TW>
__device__ __forceinline__ void inv_regs_t(u64 (&x)[1 << LOGM], const TW tw, const QC &c) {
#pragma unroll
    for (int j = LOGM - 1; j >= JMIN; --j) {
        const int half = (1 << LOGM) >> (j + 1);
#pragma unroll
        for (int g = 0; g < (1 << j); ++g) {
            const tw_t w = tw(j, g);
#pragma unroll
            for (int o = 0; o < half; ++o) bf_inv<MODE>(x[g * 2 * half + o], x[g * 2 * half + o + half], w, c);
        }
    }
}
template <int LOGM, int JMIN = 0, int MODE = B200_CSUB_INV_ROWS>
__device__ __forceinline__ void inv_regs(u64 (&x)[1 << LOGM], unsigned P, const tw_t *__restrict__ tw, const QC &c) {
    inv_regs_t<LOGM, JMIN, MODE>(x, TwHeap{tw, P}, c);
}

#ifndef B200_STREAM_DATA
#define B200_STREAM_DATA 0   // 1: polynomial words bypass L1 allocation (ld/st.global.cs) so that L1 keeps the twiddles
#endif
__device__ __forceinline__ u64 ldd(const u64 *p) {
#if B200_STREAM_DATA
    return __ldcs(p);
#else
    return *p;
#endif
}
__device__ __forceinline__ void std_(u64 *p, u64 v) {
#if B200_STREAM_DATA
    __stcs(p, v);
#else
    *p = v;
#endif
}
template <bool GATHER>
__device__ __forceinline__ void store_out(const PeerOut &po, u64 *out, size_t idx, u64 v) {
    std_(out + idx, v);
    if (GATHER) {   // peer-to-peer stores over NVLink; issued from the compute kernel so that they overlap the math
#pragma unroll 1
        for (int i = 0; i < po.n; ++i) po.p[i][idx] = v;
    }
}
__device__ __forceinline__ void ldg256(const u64 *p, u64 &a, u64 &b, u64 &c, u64 &d) {
#if B200_STREAM_DATA
    asm volatile("ld.global.cs.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(p) : "memory");
#else
    asm volatile("ld.global.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(p) : "memory");
#endif
}
__device__ __forceinline__ void stg256(u64 *p, u64 a, u64 b, u64 c, u64 d) {
#if B200_STREAM_DATA
    asm volatile("st.global.cs.v4.u64 [%0], {%1,%2,%3,%4};" ::"l"(p), "l"(a), "l"(b), "l"(c), "l"(d) : "memory");
#else
    asm volatile("st.global.v4.u64 [%0], {%1,%2,%3,%4};" ::"l"(p), "l"(a), "l"(b), "l"(c), "l"(d) : "memory");
#endif
}

__device__ __forceinline__ unsigned rev_bits(unsigned x, int bits) { return bits ? (__brev(x) >> (32 - bits)) : 0u; }

// ====================================================================================================
// row pass: 8 radix-2 layers inside 256-word rows, ONE WARP PER ROW (no CTA barrier on the hot path).
// Three register groups of 3 + 3 + 2 layers on 8 words per lane; the two exchanges go through a
// warp-private shared-memory buffer (index c + c/16: conflict-free for all four access patterns).
//   G1: lane owns c = k*32 + lane            (layers on c bits 7..5; 256-byte coalesced loads)
//   G2: lane owns c = hi3*32 + k*4 + lo2     (bits 4..2; hi3 = lane & 7, lo2 = lane >> 3)
//   G3: lane owns c = lane*8 + k             (bits 1..0 on two 4-groups; 64 contiguous bytes per lane)
// ====================================================================================================
constexpr int WROW = 272;   // 256 words + 16 padding
__device__ __forceinline__ int pad16(int c) { return c + (c >> 4); }

// Staging index of the MATRIX slot permutation: row cr = rev8(c) of the [256][RUN] tile.  A half-warp writes the
// 16 rows cr = 2j (+1) of one column; with the odd pitch RUN+1 alone those fall on 8 bank pairs (2-way conflict),
// the extra word every other 16 rows spreads them over all 16 while row-sequential reads stay conflict-free.
template <int RUN>
__device__ __forceinline__ int stage_row(int cr) { return cr * (RUN + 1) + ((cr >> 4) & 1); }

// MATRIX: a CTA owns the 6 blocks (hd) x VT consecutive v = bitrev(u); rows are ordered hd*VT + vi
template <bool MATRIX, int VT>
__device__ __forceinline__ int row_of(const BigArgs &a, int tile, int w, int &rowcode) {
    if (!MATRIX) { rowcode = 0; return tile * VT + w; }
    const int hd = w / VT, vi = w - hd * VT;
    const int u = (int)rev_bits((unsigned)(tile * VT + vi), a.logU);
    rowcode = vi * 6 + 2 * (hd % 3) + hd / 3;          // position inside a 6*VT run of the reference order
    return (hd << a.logU) + u;
}

// The 8 layers of one 256-word row held by one warp (x[k]: 8 words per lane, buf: the warp's exchange buffer).
// forward: in  c = k*32 + lane  (values < 8q)  ->  out c = lane*8 + k  (values < 8q)
__device__ __forceinline__ void row_fwd_core(u64 (&x)[8], u64 *buf, int lane, const tw_t *__restrict__ twr, const QC &qc) {
    fwd_regs_t<3, B200_CSUB_FWD_ROWS>(x, TwRow<1>{twr}, qc);
#pragma unroll
    for (int k = 0; k < 8; ++k) buf[pad16(k * 32 + lane)] = x[k];
    __syncwarp();
    const int hi3 = lane & 7, lo2 = lane >> 3;
#pragma unroll
    for (int k = 0; k < 8; ++k) x[k] = buf[pad16(hi3 * 32 + k * 4 + lo2)];
    fwd_regs_t<3, B200_CSUB_FWD_ROWS>(x, TwRow<8>{twr + 7 + hi3}, qc);
    __syncwarp();
#pragma unroll
    for (int k = 0; k < 8; ++k) buf[pad16(hi3 * 32 + k * 4 + lo2)] = x[k];
    __syncwarp();
#pragma unroll
    for (int k = 0; k < 8; ++k) x[k] = buf[pad16(lane * 8 + k)];
    u64 (&lo)[4] = *reinterpret_cast<u64 (*)[4]>(&x[0]);
    u64 (&hi)[4] = *reinterpret_cast<u64 (*)[4]>(&x[4]);
    fwd_regs_t<2, B200_CSUB_FWD_ROWS>(lo, TwRow<32>{twr + 63 + lane}, qc);
    fwd_regs_t<2, B200_CSUB_FWD_ROWS>(hi, TwRow<32>{twr + 63 + 96 + lane}, qc);
    __syncwarp();   // the buffer may be reused by the caller
}
// inverse: in  c = lane*8 + k  (values < 4q)  ->  out c = k*32 + lane  (values < 4q)
__device__ __forceinline__ void row_inv_core(u64 (&x)[8], u64 *buf, int lane, const tw_t *__restrict__ twr, const QC &qc) {
    {
        u64 (&lo)[4] = *reinterpret_cast<u64 (*)[4]>(&x[0]);
        u64 (&hi)[4] = *reinterpret_cast<u64 (*)[4]>(&x[4]);
        inv_regs_t<2, 0, B200_CSUB_INV_ROWS>(lo, TwRow<32>{twr + 63 + lane}, qc);
        inv_regs_t<2, 0, B200_CSUB_INV_ROWS>(hi, TwRow<32>{twr + 63 + 96 + lane}, qc);
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) buf[pad16(lane * 8 + k)] = x[k];
    __syncwarp();
    const int hi3 = lane & 7, lo2 = lane >> 3;
#pragma unroll
    for (int k = 0; k < 8; ++k) x[k] = buf[pad16(hi3 * 32 + k * 4 + lo2)];
    inv_regs_t<3, 0, B200_CSUB_INV_ROWS>(x, TwRow<8>{twr + 7 + hi3}, qc);
    __syncwarp();
#pragma unroll
    for (int k = 0; k < 8; ++k) buf[pad16(hi3 * 32 + k * 4 + lo2)] = x[k];
    __syncwarp();
#pragma unroll
    for (int k = 0; k < 8; ++k) x[k] = buf[pad16(k * 32 + lane)];
    inv_regs_t<3, 0, B200_CSUB_INV_ROWS>(x, TwRow<1>{twr}, qc);
}

template <bool MATRIX, int VT>   // warps per CTA: MATRIX 6*VT, STANDARD VT
__global__ void __launch_bounds__((MATRIX ? 6 : 1) * VT * 32, MATRIX ? B200_MINB_ROWS_M3 * 2 / VT : B200_MINB_ROWS_STD_FWD * 4 / VT)
k_rows_fwd(const u64 *in, u64 *out, BigArgs a) {
    constexpr int RUN = 6 * VT;
    extern __shared__ __align__(16) u64 sm[];
    const UnitRef ur = unit_ref(a);
    const int limb = ur.limb;
    const QC qc(a.consts[limb]);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int rowcode;
    const int r = row_of<MATRIX, VT>(a, blockIdx.y, w, rowcode);
    u64 *buf = sm + w * WROW;
    const u64 *src = in + ur.in + (size_t)r * 256;
    const tw_t *__restrict__ twr = a.fwd_r + (size_t)limb * a.N + a.R + 255 * r;
    u64 x[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) x[k] = ldd(src + k * 32 + lane);
    row_fwd_core(x, buf, lane, twr, qc);
#pragma unroll
    for (int k = 0; k < 8; ++k) x[k] = (a.flags & B200_LAZY) ? csub(x[k], qc.fourq) : canon8<(MATRIX ? B200_CANON_FINAL_M3 : B200_CANON_FINAL_STD)>(x[k], qc);   // lazy: [0,4q) <= 6q-2
    if (!MATRIX || (a.flags & B200_TREE_ORDER)) {
        u64 *dst = out + ur.out + (size_t)r * 256 + lane * 8;
        stg256(dst, x[0], x[1], x[2], x[3]);
        stg256(dst + 4, x[4], x[5], x[6], x[7]);
    } else {
        // reference order: slot = rev8(c)*6U + v*6 + (2d+h); stage the tile as [rev8(c)][RUN] (pitch RUN+1)
        __syncthreads();
        const int rl = (int)rev_bits((unsigned)lane, 5);
#pragma unroll
        for (int k = 0; k < 8; ++k) sm[stage_row<RUN>((int)rev_bits((unsigned)k, 3) * 32 + rl) + rowcode] = x[k];
        __syncthreads();
        const size_t run = (size_t)6 << a.logU;                       // 6U words between consecutive rev8(c)
        u64 *dst = out + ur.out + (size_t)blockIdx.y * RUN;
        {   // WARPS*32 = 32*RUN threads: thread -> (cr0, rem) once, then cr advances by 32 per step
            const int cr0 = threadIdx.x / RUN, rem = threadIdx.x - cr0 * RUN;
            // stage_row(cr0 + 32 i) = stage_row(cr0) + 32 i (RUN + 1): constant shared-memory offsets, one running global pointer
            const u64 *sp = sm + stage_row<RUN>(cr0) + rem;
            u64 *gp = dst + (size_t)cr0 * run + rem;
            const size_t gstep = 32 * run;
#pragma unroll
            for (int i = 0; i < 8; ++i, gp += gstep) std_(gp, sp[i * 32 * (RUN + 1)]);
        }
    }
}

template <bool MATRIX, int VT>
__global__ void __launch_bounds__((MATRIX ? 6 : 1) * VT * 32, MATRIX ? B200_MINB_ROWS_M3 * 2 / VT : B200_MINB_ROWS_STD_INV * 4 / VT)
k_rows_inv(const u64 *in, u64 *out, BigArgs a) {
    constexpr int RUN = 6 * VT;
    extern __shared__ __align__(16) u64 sm[];
    const UnitRef ur = unit_ref(a);
    const int limb = ur.limb;
    const QC qc(a.consts[limb]);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int rowcode;
    const int r = row_of<MATRIX, VT>(a, blockIdx.y, w, rowcode);
    u64 *buf = sm + w * WROW;
    const tw_t *__restrict__ twr = a.inv_r + (size_t)limb * a.N + a.R + 255 * r;
    u64 x[8];
    if (!MATRIX || (a.flags & B200_TREE_ORDER)) {
        const u64 *src = in + ur.in + (size_t)r * 256 + lane * 8;
        ldg256(src, x[0], x[1], x[2], x[3]);
        ldg256(src + 4, x[4], x[5], x[6], x[7]);
    } else {
        const size_t run = (size_t)6 << a.logU;
        const u64 *src = in + ur.in + (size_t)blockIdx.y * RUN;
        {
            const int cr0 = threadIdx.x / RUN, rem = threadIdx.x - cr0 * RUN;
            u64 v[8];
            const u64 *gp = src + (size_t)cr0 * run + rem;
            const size_t gstep = 32 * run;
#pragma unroll
            for (int i = 0; i < 8; ++i, gp += gstep) v[i] = ldd(gp);     // all 8 loads in flight
            u64 *sp = sm + stage_row<RUN>(cr0) + rem;
#pragma unroll
            for (int i = 0; i < 8; ++i) sp[i * 32 * (RUN + 1)] = v[i];
        }
        __syncthreads();
        const int rl = (int)rev_bits((unsigned)lane, 5);
#pragma unroll
        for (int k = 0; k < 8; ++k) x[k] = sm[stage_row<RUN>((int)rev_bits((unsigned)k, 3) * 32 + rl) + rowcode];
        __syncthreads();
    }
    row_inv_core(x, buf, lane, twr, qc);
    u64 *dst = out + ur.out + (size_t)r * 256;
#pragma unroll
    for (int k = 0; k < 8; ++k) std_(dst + k * 32 + lane, x[k]);
}

// Fused row stage of the ring product (BASELINE config 4): for one row of a^ = cols_fwd(a) and b^ = cols_fwd(b)
// run the 8 forward layers of both, multiply slot-wise (a*b mod q, the value of MulCoeffsMontgomery(MForm(a), b),
// ring/ring_test.go:106-110) and run the 8 inverse layers -- the NTT-domain polynomials never touch memory and no
// slot permutation is needed because the product is order agnostic.  ta is overwritten with the result.
__global__ void __launch_bounds__(128, B200_MINB_ROWS_MUL)
k_rows_mul(u64 *ta, const u64 *tb, BigArgs a) {
    extern __shared__ __align__(16) u64 sm[];
    const UnitRef ur = unit_ref(a);
    const int limb = ur.limb;
    const LimbConst lc = a.consts[limb];
    const QC qc(lc);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int r = blockIdx.y * 4 + w;
    u64 *buf = sm + w * WROW;
    const tw_t *__restrict__ twr = a.fwd_r + (size_t)limb * a.N + a.R + 255 * r, *__restrict__ twri = a.inv_r + (size_t)limb * a.N + a.R + 255 * r;
    u64 *pa = ta + ur.in + (size_t)r * 256;
    const u64 *pb = tb + ur.in + (size_t)r * 256;
    u64 x[8], y[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) x[k] = pa[k * 32 + lane];
#pragma unroll
    for (int k = 0; k < 8; ++k) y[k] = pb[k * 32 + lane];
    row_fwd_core(x, buf, lane, twr, qc);
    row_fwd_core(y, buf, lane, twr, qc);
#pragma unroll
    for (int k = 0; k < 8; ++k)   // canonical operands -> Barrett product in [0, 2q)
        x[k] = bred_lazy(canon8(x[k], qc), canon8(y[k], qc), lc.q, lc.brc_hi, lc.brc_lo);
    row_inv_core(x, buf, lane, twri, qc);
#pragma unroll
    for (int k = 0; k < 8; ++k) pa[k * 32 + lane] = x[k];
}

// ====================================================================================================
// column pass.  A CTA owns CW adjacent columns of ALL rows (CW*8-byte row segments), CW*32 threads per
// 256 rows.  Shared index of (row, c): pc(row*CW + c), padded every 64 words so that all three groups
// are conflict-free.
// ====================================================================================================
__device__ __forceinline__ int pc(int i) { return i + ((i >> 6) << 3); }

// STANDARD, log2(N) = 16: 256 rows, layers on row bits [7..5] [4..2] [1..0]
template <int CW>
__global__ void __launch_bounds__(CW * 32, B200_MINB_COLS_STD * 8 / CW)
k_cols_fwd_std16(const u64 *in, u64 *out, BigArgs a) {
    extern __shared__ __align__(16) u64 sm[];
    const UnitRef ur = unit_ref<true>(a);
    const int limb = ur.limb;
    const size_t ibase = ur.in + (size_t)blockIdx.x * CW, obase = ur.out + (size_t)blockIdx.x * CW;
    const QC qc(a.consts[limb]);
    const tw_t *__restrict__ tw = a.fwd + (size_t)limb * a.N;
    const int t = threadIdx.x, c = t % CW, g = t / CW;   // g in [0,32)
    u64 x[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) x[k] = ldd(in + ibase + (size_t)(k * 32 + g) * 256 + c);      // rows k*32 + g
    fwd_regs<3, B200_CSUB_FWD_COLS_STD>(x, 1u, tw, qc);
#pragma unroll
    for (int k = 0; k < 8; ++k) sm[pc((k * 32 + g) * CW + c)] = x[k];
    __syncthreads();
    const int lo2 = g & 3, hi3 = g >> 2;                                                 // rows hi3*32 + k*4 + lo2
#pragma unroll
    for (int k = 0; k < 8; ++k) x[k] = sm[pc((hi3 * 32 + k * 4 + lo2) * CW + c)];
    fwd_regs<3, B200_CSUB_FWD_COLS_STD>(x, 8u + hi3, tw, qc);
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 8; ++k) sm[pc((hi3 * 32 + k * 4 + lo2) * CW + c)] = x[k];
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 8; ++k) x[k] = sm[pc((g * 8 + k) * CW + c)];                     // rows g*8 + k
    {
        u64 (&lo)[4] = *reinterpret_cast<u64 (*)[4]>(&x[0]);
        u64 (&hi)[4] = *reinterpret_cast<u64 (*)[4]>(&x[4]);
        fwd_regs<2, B200_CSUB_FWD_COLS_STD>(lo, 64u + g * 2, tw, qc);
        fwd_regs<2, B200_CSUB_FWD_COLS_STD>(hi, 64u + g * 2 + 1, tw, qc);
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) std_(out + obase + (size_t)(g * 8 + k) * 256 + c, x[k]);
}

template <int CW, bool GATHER = false>
__global__ void __launch_bounds__(CW * 32, B200_MINB_COLS_STD * 8 / CW)
k_cols_inv_std16(const u64 *in, u64 *out, BigArgs a) {
    extern __shared__ __align__(16) u64 sm[];
    const UnitRef ur = unit_ref<true>(a);
    const int limb = ur.limb;
    const size_t ibase = ur.in + (size_t)blockIdx.x * CW, obase = ur.out + (size_t)blockIdx.x * CW;
    const QC qc(a.consts[limb]);
    const u64 q = qc.q;
    const tw_t *__restrict__ tw = a.inv + (size_t)limb * a.N;
    const tw_t ninv = a.misc[limb * MISC_COUNT + MISC_NINV], ninv_w1 = a.misc[limb * MISC_COUNT + MISC_NINV_W1];
    const bool lazy = a.flags & B200_LAZY;
    const int t = threadIdx.x, c = t % CW, g = t / CW;
    u64 x[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) x[k] = ldd(in + ibase + (size_t)(g * 8 + k) * 256 + c);
    {
        u64 (&lo)[4] = *reinterpret_cast<u64 (*)[4]>(&x[0]);
        u64 (&hi)[4] = *reinterpret_cast<u64 (*)[4]>(&x[4]);
        inv_regs<2, 0, B200_CSUB_INV_COLS>(lo, 64u + g * 2, tw, qc);
        inv_regs<2, 0, B200_CSUB_INV_COLS>(hi, 64u + g * 2 + 1, tw, qc);
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) sm[pc((g * 8 + k) * CW + c)] = x[k];
    __syncthreads();
    const int lo2 = g & 3, hi3 = g >> 2;
#pragma unroll
    for (int k = 0; k < 8; ++k) x[k] = sm[pc((hi3 * 32 + k * 4 + lo2) * CW + c)];
    inv_regs<3, 0, B200_CSUB_INV_COLS>(x, 8u + hi3, tw, qc);
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 8; ++k) sm[pc((hi3 * 32 + k * 4 + lo2) * CW + c)] = x[k];
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 8; ++k) x[k] = sm[pc((k * 32 + g) * CW + c)];
    inv_regs<3, 1, B200_CSUB_INV_COLS>(x, 1u, tw, qc);
#pragma unroll
    for (int o = 0; o < 4; ++o) {   // last layer carries N^-1
        const u64 U = x[o], V = x[o + 4];
        u64 X = mul2q<(B200_SHOUP2Q_INV != 0)>(U + V, ninv, qc);
        u64 Y = mul2q<(B200_SHOUP2Q_INV != 0)>(U - V + qc.fourq, ninv_w1, qc);   // inputs < 4q
        if (!lazy) { X = red_ab(X - q, q); Y = red_ab(Y - q, q); }
        x[o] = X; x[o + 4] = Y;
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) store_out<GATHER>(a.peers, out, obase + (size_t)(k * 32 + g) * 256 + c, x[k]);
}

// MATRIX, a = 15 (N = 3*2^15): rows (hd, u), hd in [0,6) = 3h + d, u in [0,64).
//   G1: 12 words  (hd x u bit 5): cyclotomic, radix-3, radix-2 layer 0      -- CW*32 threads
//   G2:  8 words  (u bits 4..2): radix-2 layers 1..3                          -- CW*48 threads
//   G3:  8 words  (u bits 2..0, two 4-groups): radix-2 layers 4, 5            -- CW*48 threads
template <int CW>
__global__ void __launch_bounds__(CW * 48, B200_MINB_COLS_M3_FWD * 8 / CW)
k_cols_fwd_m3_15(const u64 *in, u64 *out, BigArgs a) {
    extern __shared__ __align__(16) u64 sm[];
    const UnitRef ur = unit_ref<true>(a);
    const int limb = ur.limb;
    const size_t ibase = ur.in + (size_t)blockIdx.x * CW, obase = ur.out + (size_t)blockIdx.x * CW;
    const QC qc(a.consts[limb]);
    const u64 twoq = qc.twoq, fourq = qc.fourq;
    const tw_t *__restrict__ tw = a.fwd + (size_t)limb * a.N;
    const tw_t om = a.misc[limb * MISC_COUNT + MISC_OMEGA3];
    const int t = threadIdx.x, c = t % CW, g = t / CW;   // g in [0,48)
    if (g < 32) {   // G1: rows hd*64 + b*32 + g
        u64 x[6][2];
#pragma unroll
        for (int hd = 0; hd < 6; ++hd)
#pragma unroll
            for (int b = 0; b < 2; ++b) x[hd][b] = ldd(in + ibase + (size_t)(hd * 64 + b * 32 + g) * 256 + c);
        const tw_t zeta = tw[1];
#pragma unroll
        for (int b = 0; b < 2; ++b) {
#pragma unroll
            for (int d = 0; d < 3; ++d) {   // cyclotomic layer: in [0,2q) -> [0,4q)
                const u64 lo = x[d][b], hi = x[d + 3][b];
                const u64 tt = mul2q<(B200_SHOUP2Q_FWD != 0)>(hi, zeta, qc);
                x[d][b] = lo + tt;
                x[d + 3][b] = csub(lo + hi + twoq - tt, fourq);
            }
#pragma unroll
            for (int h = 0; h < 2; ++h) {   // radix-3 layer: [0,4q) -> [0,4q)
                const tw_t z1 = tw[2 + 2 * h], z2 = tw[3 + 2 * h];
                const u64 b0 = csub(x[3 * h][b], twoq);
                const u64 t1 = mul2q<(B200_SHOUP2Q_FWD != 0)>(x[3 * h + 1][b], z1, qc);
                const u64 t2 = mul2q<(B200_SHOUP2Q_FWD != 0)>(x[3 * h + 2][b], z2, qc);
                const u64 t3 = mul2q<(B200_SHOUP2Q_FWD != 0)>(t1 + twoq - t2, om, qc);
                x[3 * h][b] = csub(b0 + t1 + t2, fourq);
                x[3 * h + 1][b] = csub(b0 + twoq - t2 + t3, fourq);
                x[3 * h + 2][b] = csub(b0 + fourq - t1 - t3, fourq);
            }
        }
#pragma unroll
        for (int hd = 0; hd < 6; ++hd) {
            bf_fwd<B200_CSUB_FWD_COLS_M3>(x[hd][0], x[hd][1], tw[6 + hd], qc);   // radix-2 layer 0 of block hd
#pragma unroll
            for (int b = 0; b < 2; ++b) sm[pc((hd * 64 + b * 32 + g) * CW + c)] = x[hd][b];
        }
    }
    __syncthreads();
    u64 x[8];
    {   // G2: g = top*4 + lo2, top = hd*2 + u5 in [0,12); rows top*32 + k*4 + lo2
        const int lo2 = g & 3, top = g >> 2;
#pragma unroll
        for (int k = 0; k < 8; ++k) x[k] = sm[pc((top * 32 + k * 4 + lo2) * CW + c)];
        fwd_regs<3, B200_CSUB_FWD_COLS_M3>(x, 12u + top, tw, qc);
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 8; ++k) sm[pc((top * 32 + k * 4 + lo2) * CW + c)] = x[k];
    }
    __syncthreads();
    {   // G3: rows g*8 + k; sub-transform roots at depth 4: 96 + g*2 + {0,1}
#pragma unroll
        for (int k = 0; k < 8; ++k) x[k] = sm[pc((g * 8 + k) * CW + c)];
        u64 (&lo)[4] = *reinterpret_cast<u64 (*)[4]>(&x[0]);
        u64 (&hi)[4] = *reinterpret_cast<u64 (*)[4]>(&x[4]);
        fwd_regs<2, B200_CSUB_FWD_COLS_M3>(lo, 96u + g * 2, tw, qc);
        fwd_regs<2, B200_CSUB_FWD_COLS_M3>(hi, 96u + g * 2 + 1, tw, qc);
#pragma unroll
        for (int k = 0; k < 8; ++k) std_(out + obase + (size_t)(g * 8 + k) * 256 + c, x[k]);
    }
}

template <int CW, bool GATHER = false>
__global__ void __launch_bounds__(CW * 48, B200_MINB_COLS_M3_INV * 8 / CW)
k_cols_inv_m3_15(const u64 *in, u64 *out, BigArgs a) {
    extern __shared__ __align__(16) u64 sm[];
    const UnitRef ur = unit_ref<true>(a);
    const int limb = ur.limb;
    const size_t ibase = ur.in + (size_t)blockIdx.x * CW, obase = ur.out + (size_t)blockIdx.x * CW;
    const QC qc(a.consts[limb]);
    const u64 q = qc.q, twoq = qc.twoq, fourq = qc.fourq;
    const tw_t *__restrict__ tw = a.inv + (size_t)limb * a.N;
    const tw_t *mc = a.misc + limb * MISC_COUNT;
    const bool lazy = a.flags & B200_LAZY;
    const int t = threadIdx.x, c = t % CW, g = t / CW;
    {
        u64 x[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) x[k] = ldd(in + ibase + (size_t)(g * 8 + k) * 256 + c);
        u64 (&lo)[4] = *reinterpret_cast<u64 (*)[4]>(&x[0]);
        u64 (&hi)[4] = *reinterpret_cast<u64 (*)[4]>(&x[4]);
        inv_regs<2, 0, B200_CSUB_INV_COLS>(lo, 96u + g * 2, tw, qc);
        inv_regs<2, 0, B200_CSUB_INV_COLS>(hi, 96u + g * 2 + 1, tw, qc);
#pragma unroll
        for (int k = 0; k < 8; ++k) sm[pc((g * 8 + k) * CW + c)] = x[k];
        __syncthreads();
        const int lo2 = g & 3, top = g >> 2;
#pragma unroll
        for (int k = 0; k < 8; ++k) x[k] = sm[pc((top * 32 + k * 4 + lo2) * CW + c)];
        inv_regs<3, 0, B200_CSUB_INV_COLS>(x, 12u + top, tw, qc);
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 8; ++k) sm[pc((top * 32 + k * 4 + lo2) * CW + c)] = x[k];
    }
    __syncthreads();
    if (g < 32) {
        const tw_t om = mc[MISC_OMEGA3], k1 = mc[MISC_K1], k2 = mc[MISC_K2], zeta = mc[MISC_ZETA];
        u64 x[6][2];
#pragma unroll
        for (int hd = 0; hd < 6; ++hd) {
#pragma unroll
            for (int b = 0; b < 2; ++b) x[hd][b] = sm[pc((hd * 64 + b * 32 + g) * CW + c)];
            bf_inv<B200_CSUB_INV_COLS>(x[hd][0], x[hd][1], tw[6 + hd], qc);
        }
#pragma unroll
        for (int b = 0; b < 2; ++b) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {   // inverse radix-3 (factor 3 deferred): [0,2q) -> [0,2q)
                const tw_t z1 = tw[2 + 2 * h], z2 = tw[3 + 2 * h];
                const u64 y0 = to2q(x[3 * h][b], qc), y1 = to2q(x[3 * h + 1][b], qc), y2 = to2q(x[3 * h + 2][b], qc);
#if B200_R3INV_APPROX
                // the three products feed only other products, so the approximate quotient ([0,4q)) is enough:
                // offsets grow by 2q, every operand stays below 8q < 2^64
                const u64 s = shoup4(y1 + twoq - y2, om.x, om.y, qc.nq);
                x[3 * h][b] = csub(csub(y0 + y1 + y2, fourq), twoq);
                x[3 * h + 1][b] = shoup4(y0 + fourq + twoq - y1 - s, z1.x, z1.y, qc.nq);
                x[3 * h + 2][b] = shoup4(y0 + twoq - y2 + s, z2.x, z2.y, qc.nq);
#else
                const u64 s = mul2q<(B200_SHOUP2Q_INV != 0)>(y1 + twoq - y2, om, qc);
                x[3 * h][b] = csub(csub(y0 + y1 + y2, fourq), twoq);
                x[3 * h + 1][b] = mul2q<(B200_SHOUP2Q_INV != 0)>(y0 + fourq - y1 - s, z1, qc);
                x[3 * h + 2][b] = mul2q<(B200_SHOUP2Q_INV != 0)>(y0 + twoq - y2 + s, z2, qc);
#endif
            }
#pragma unroll
            for (int d = 0; d < 3; ++d) {   // inverse cyclotomic layer with the 2/N scaling
                const u64 lo = x[d][b], hi = x[d + 3][b];
                #if B200_CYCINV_APPROX
                u64 hh = shoup4(lo + (B200_R3INV_APPROX ? fourq : twoq) - hi, k1.x, k1.y, qc.nq);                        // [0,4q)
                u64 ll = shoup4(lo, k2.x, k2.y, qc.nq) + fourq - shoup4(hh, zeta.x, zeta.y, qc.nq);                      // (0,8q)
                if (!lazy) { ll = canon_est<2>(ll, q, qc.nq, qc.csh, qc.cmul); hh = canon_est<2>(hh, q, qc.nq, qc.csh, qc.cmul); }
                else { ll = canon_est<3>(ll, q, qc.nq, qc.csh, qc.cmul); hh = csub(hh, twoq); }
#else
                u64 hh = mul2q<(B200_SHOUP2Q_INV != 0)>(lo + (B200_R3INV_APPROX ? fourq : twoq) - hi, k1, qc);
                u64 ll = csub(mul2q<(B200_SHOUP2Q_INV != 0)>(lo, k2, qc) + twoq - mul2q<(B200_SHOUP2Q_INV != 0)>(hh, zeta, qc), twoq);
                if (!lazy) { ll = csub(ll, q); hh = csub(hh, q); }
#endif
                x[d][b] = ll; x[d + 3][b] = hh;
            }
        }
#pragma unroll
        for (int hd = 0; hd < 6; ++hd)
#pragma unroll
            for (int b = 0; b < 2; ++b) store_out<GATHER>(a.peers, out, obase + (size_t)(hd * 64 + b * 32 + g) * 256 + c, x[hd][b]);
    }
}

// ====================================================================================================
// generic column pass: any STANDARD ring with 2^10 <= N <= 2^17 and any MATRIX ring N = 3*2^a, 10 <= a <= 16.
// Same structure as the tuned kernels above (8 columns per CTA, register groups of <= 3 layers, shared-memory
// exchanges), with the group sizes as template parameters.  Row index = hi*2^(b0+S) + k*2^b0 + lo for the
// group that owns row bits [b0, b0+S); the root node of its sub-transform is (ROOTS << (LOG-b0-S)) + hi.
// ====================================================================================================
#ifndef B200_GCW_LOG
#define B200_GCW_LOG 3
#endif
constexpr int GCW_LOG = B200_GCW_LOG, GCW = 1 << GCW_LOG;   // columns per CTA of the generic column kernels

template <int S, int ROOTS, int LOG, bool INV, bool FROM_G, bool TO_G, bool FOLD_NINV, bool GATHER = false>
__device__ __forceinline__ void col_group(const u64 *gin, size_t ibase, u64 *gout, size_t obase, u64 *sm, int b0,
                                          const tw_t *__restrict__ tw, const QC &qc, const tw_t ninv, const tw_t ninv_w1, bool lazy,
                                          const PeerOut *po = nullptr) {
    constexpr int M = 1 << S;
    const int ntiles = (GCW * ROOTS) << (LOG - S);
    for (int t = threadIdx.x; t < ntiles; t += blockDim.x) {
        const int c = t & (GCW - 1), rest = t >> GCW_LOG;
        const int lo = rest & ((1 << b0) - 1), hi = rest >> b0;
        const int row0 = (hi << (b0 + S)) + lo;
        u64 x[M];
#pragma unroll
        for (int k = 0; k < M; ++k) {
            const int row = row0 + (k << b0);
            x[k] = FROM_G ? ldd(gin + ibase + (size_t)row * 256 + c) : sm[pc(row * GCW + c)];
        }
        const unsigned P = (unsigned)((ROOTS << (LOG - b0 - S)) + hi);
        if (!INV) {
            fwd_regs<S, (ROOTS == 1 ? B200_CSUB_FWD_COLS_STD : B200_CSUB_FWD_COLS_M3)>(x, P, tw, qc);
        } else if (!FOLD_NINV) {
            inv_regs<S, 0, B200_CSUB_INV_COLS>(x, P, tw, qc);
        } else {   // top group of the STANDARD inverse: the last layer carries N^-1
            inv_regs<S, 1, B200_CSUB_INV_COLS>(x, P, tw, qc);
#pragma unroll
            for (int o = 0; o < M / 2; ++o) {
                const u64 U = x[o], V = x[o + M / 2];
                u64 X = mul2q<(B200_SHOUP2Q_INV != 0)>(U + V, ninv, qc);
                u64 Y = mul2q<(B200_SHOUP2Q_INV != 0)>(U - V + qc.fourq, ninv_w1, qc);
                if (!lazy) { X = red_ab(X - qc.q, qc.q); Y = red_ab(Y - qc.q, qc.q); }
                x[o] = X; x[o + M / 2] = Y;
            }
        }
#pragma unroll
        for (int k = 0; k < M; ++k) {
            const int row = row0 + (k << b0);
            if (TO_G) { if (GATHER) store_out<true>(*po, gout, obase + (size_t)row * 256 + c, x[k]); else std_(gout + obase + (size_t)row * 256 + c, x[k]); }
            else sm[pc(row * GCW + c)] = x[k];
        }
    }
}

constexpr int std_col_threads(int L) { if (L >= 6) return GCW << (L - 3); return 64; }
constexpr int m3_col_threads(int L) { if (L >= 4) return (GCW * 6) << (L - 3); return 64; }
constexpr int minb_for(int threads, int target) { return target / threads < 1 ? 1 : (target / threads > 16 ? 16 : target / threads); }

template <int LOGR, int S1, int S2, int S3, bool INV, bool GATHER = false>
__global__ void __launch_bounds__(std_col_threads(LOGR), minb_for(std_col_threads(LOGR), 1024))
k_cols_std_g(const u64 *in, u64 *out, BigArgs a) {
    static_assert(S1 + S2 + S3 == LOGR && S1 >= 1 && S2 >= 1, "group sizes");
    extern __shared__ __align__(16) u64 sm[];
    const UnitRef ur = unit_ref<true>(a);
    const int limb = ur.limb;
    const size_t ibase = ur.in + (size_t)blockIdx.x * GCW, obase = ur.out + (size_t)blockIdx.x * GCW;
    const QC qc(a.consts[limb]);
    const tw_t *__restrict__ tw = (INV ? a.inv : a.fwd) + (size_t)limb * a.N;
    const tw_t ninv = a.misc[limb * MISC_COUNT + MISC_NINV], ninv_w1 = a.misc[limb * MISC_COUNT + MISC_NINV_W1];
    const bool lazy = a.flags & B200_LAZY;
    if (!INV) {
        col_group<S1, 1, LOGR, false, true, false, false>(in, ibase, out, obase, sm, S2 + S3, tw, qc, ninv, ninv_w1, lazy);
        __syncthreads();
        if (S3 == 0) {
            col_group<S2, 1, LOGR, false, false, true, false>(in, ibase, out, obase, sm, 0, tw, qc, ninv, ninv_w1, lazy);
        } else {
            col_group<S2, 1, LOGR, false, false, false, false>(in, ibase, out, obase, sm, S3, tw, qc, ninv, ninv_w1, lazy);
            __syncthreads();
            col_group<(S3 ? S3 : 1), 1, LOGR, false, false, true, false>(in, ibase, out, obase, sm, 0, tw, qc, ninv, ninv_w1, lazy);
        }
    } else {
        if (S3 == 0) {
            col_group<S2, 1, LOGR, true, true, false, false>(in, ibase, out, obase, sm, 0, tw, qc, ninv, ninv_w1, lazy);
        } else {
            col_group<(S3 ? S3 : 1), 1, LOGR, true, true, false, false>(in, ibase, out, obase, sm, 0, tw, qc, ninv, ninv_w1, lazy);
            __syncthreads();
            col_group<S2, 1, LOGR, true, false, false, false>(in, ibase, out, obase, sm, S3, tw, qc, ninv, ninv_w1, lazy);
        }
        __syncthreads();
        col_group<S1, 1, LOGR, true, false, true, true, GATHER>(in, ibase, out, obase, sm, S2 + S3, tw, qc, ninv, ninv_w1, lazy, &a.peers);
    }
}

// MATRIX: first group = cyclotomic + radix-3 + S1 radix-2 layers on 6*2^S1 words; then groups on the u bits.
template <int LOGU, int S1, int S2, int S3, bool INV, bool GATHER = false>
__global__ void __launch_bounds__(m3_col_threads(LOGU), minb_for(m3_col_threads(LOGU), 768))
k_cols_m3_g(const u64 *in, u64 *out, BigArgs a) {
    static_assert(S1 + S2 + S3 == LOGU && S2 >= 1, "group sizes");
    extern __shared__ __align__(16) u64 sm[];
    const UnitRef ur = unit_ref<true>(a);
    const int limb = ur.limb;
    const size_t ibase = ur.in + (size_t)blockIdx.x * GCW, obase = ur.out + (size_t)blockIdx.x * GCW;
    const QC qc(a.consts[limb]);
    const u64 q = qc.q, twoq = qc.twoq, fourq = qc.fourq;
    const tw_t *__restrict__ tw = (INV ? a.inv : a.fwd) + (size_t)limb * a.N;
    const tw_t *mc = a.misc + limb * MISC_COUNT;
    const tw_t om = mc[MISC_OMEGA3], none = make_ulonglong2(0, 0);
    const bool lazy = a.flags & B200_LAZY;
    constexpr int MA = 1 << S1, U = 1 << LOGU, ULO = U >> S1;
    if (!INV) {
        for (int t = threadIdx.x; t < GCW * ULO; t += blockDim.x) {
            const int c = t & (GCW - 1), ulo = t >> GCW_LOG;
            u64 x[6][MA];
#pragma unroll
            for (int hd = 0; hd < 6; ++hd)
#pragma unroll
                for (int k = 0; k < MA; ++k) x[hd][k] = ldd(in + ibase + (size_t)(hd * U + k * ULO + ulo) * 256 + c);
            const tw_t zeta = tw[1];
#pragma unroll
            for (int k = 0; k < MA; ++k) {
#pragma unroll
                for (int d = 0; d < 3; ++d) {   // cyclotomic layer: in [0,2q) -> [0,4q)
                    const u64 lo = x[d][k], hi = x[d + 3][k];
                    const u64 tt = mul2q<(B200_SHOUP2Q_FWD != 0)>(hi, zeta, qc);
                    x[d][k] = lo + tt;
                    x[d + 3][k] = csub(lo + hi + twoq - tt, fourq);
                }
#pragma unroll
                for (int h = 0; h < 2; ++h) {   // radix-3 layer: [0,4q) -> [0,4q)
                    const tw_t z1 = tw[2 + 2 * h], z2 = tw[3 + 2 * h];
                    const u64 b0 = csub(x[3 * h][k], twoq);
                    const u64 t1 = mul2q<(B200_SHOUP2Q_FWD != 0)>(x[3 * h + 1][k], z1, qc);
                    const u64 t2 = mul2q<(B200_SHOUP2Q_FWD != 0)>(x[3 * h + 2][k], z2, qc);
                    const u64 t3 = mul2q<(B200_SHOUP2Q_FWD != 0)>(t1 + twoq - t2, om, qc);
                    x[3 * h][k] = csub(b0 + t1 + t2, fourq);
                    x[3 * h + 1][k] = csub(b0 + twoq - t2 + t3, fourq);
                    x[3 * h + 2][k] = csub(b0 + fourq - t1 - t3, fourq);
                }
            }
#pragma unroll
            for (int hd = 0; hd < 6; ++hd) {
                if (S1 > 0) fwd_regs<S1, B200_CSUB_FWD_COLS_M3>(x[hd], (unsigned)(6 + hd), tw, qc);
#pragma unroll
                for (int k = 0; k < MA; ++k) sm[pc((hd * U + k * ULO + ulo) * GCW + c)] = x[hd][k];
            }
        }
        __syncthreads();
        if (S3 == 0) {
            col_group<S2, 6, LOGU, false, false, true, false>(in, ibase, out, obase, sm, 0, tw, qc, none, none, lazy);
        } else {
            col_group<S2, 6, LOGU, false, false, false, false>(in, ibase, out, obase, sm, S3, tw, qc, none, none, lazy);
            __syncthreads();
            col_group<(S3 ? S3 : 1), 6, LOGU, false, false, true, false>(in, ibase, out, obase, sm, 0, tw, qc, none, none, lazy);
        }
    } else {
        if (S3 == 0) {
            col_group<S2, 6, LOGU, true, true, false, false>(in, ibase, out, obase, sm, 0, tw, qc, none, none, lazy);
        } else {
            col_group<(S3 ? S3 : 1), 6, LOGU, true, true, false, false>(in, ibase, out, obase, sm, 0, tw, qc, none, none, lazy);
            __syncthreads();
            col_group<S2, 6, LOGU, true, false, false, false>(in, ibase, out, obase, sm, S3, tw, qc, none, none, lazy);
        }
        __syncthreads();
        const tw_t k1 = mc[MISC_K1], k2 = mc[MISC_K2], zeta = mc[MISC_ZETA];
        for (int t = threadIdx.x; t < GCW * ULO; t += blockDim.x) {
            const int c = t & (GCW - 1), ulo = t >> GCW_LOG;
            u64 x[6][MA];
#pragma unroll
            for (int hd = 0; hd < 6; ++hd) {
#pragma unroll
                for (int k = 0; k < MA; ++k) x[hd][k] = sm[pc((hd * U + k * ULO + ulo) * GCW + c)];
                if (S1 > 0) inv_regs<S1, 0, B200_CSUB_INV_COLS>(x[hd], (unsigned)(6 + hd), tw, qc);
            }
#pragma unroll
            for (int k = 0; k < MA; ++k) {
#pragma unroll
                for (int h = 0; h < 2; ++h) {   // inverse radix-3 (factor 3 deferred)
                    const tw_t z1 = tw[2 + 2 * h], z2 = tw[3 + 2 * h];
                    const u64 y0 = to2q(x[3 * h][k], qc), y1 = to2q(x[3 * h + 1][k], qc), y2 = to2q(x[3 * h + 2][k], qc);
#if B200_R3INV_APPROX
                    const u64 sgm = shoup4(y1 + twoq - y2, om.x, om.y, qc.nq);
                    x[3 * h][k] = csub(csub(y0 + y1 + y2, fourq), twoq);
                    x[3 * h + 1][k] = shoup4(y0 + fourq + twoq - y1 - sgm, z1.x, z1.y, qc.nq);
                    x[3 * h + 2][k] = shoup4(y0 + twoq - y2 + sgm, z2.x, z2.y, qc.nq);
#else
                    const u64 sgm = mul2q<(B200_SHOUP2Q_INV != 0)>(y1 + twoq - y2, om, qc);
                    x[3 * h][k] = csub(csub(y0 + y1 + y2, fourq), twoq);
                    x[3 * h + 1][k] = mul2q<(B200_SHOUP2Q_INV != 0)>(y0 + fourq - y1 - sgm, z1, qc);
                    x[3 * h + 2][k] = mul2q<(B200_SHOUP2Q_INV != 0)>(y0 + twoq - y2 + sgm, z2, qc);
#endif
                }
#pragma unroll
                for (int d = 0; d < 3; ++d) {   // inverse cyclotomic layer with the 2/N scaling
                    const u64 lo = x[d][k], hi = x[d + 3][k];
                    #if B200_CYCINV_APPROX
                    u64 hh = shoup4(lo + (B200_R3INV_APPROX ? fourq : twoq) - hi, k1.x, k1.y, qc.nq);                        // [0,4q)
                    u64 ll = shoup4(lo, k2.x, k2.y, qc.nq) + fourq - shoup4(hh, zeta.x, zeta.y, qc.nq);                      // (0,8q)
                    if (!lazy) { ll = canon_est<2>(ll, q, qc.nq, qc.csh, qc.cmul); hh = canon_est<2>(hh, q, qc.nq, qc.csh, qc.cmul); }
                    else { ll = canon_est<3>(ll, q, qc.nq, qc.csh, qc.cmul); hh = csub(hh, twoq); }
#else
                    u64 hh = mul2q<(B200_SHOUP2Q_INV != 0)>(lo + (B200_R3INV_APPROX ? fourq : twoq) - hi, k1, qc);
                    u64 ll = csub(mul2q<(B200_SHOUP2Q_INV != 0)>(lo, k2, qc) + twoq - mul2q<(B200_SHOUP2Q_INV != 0)>(hh, zeta, qc), twoq);
                    if (!lazy) { ll = csub(ll, q); hh = csub(hh, q); }
#endif
                    x[d][k] = ll; x[d + 3][k] = hh;
                }
            }
#pragma unroll
            for (int hd = 0; hd < 6; ++hd)
#pragma unroll
                for (int k = 0; k < MA; ++k) store_out<GATHER>(a.peers, out, obase + (size_t)(hd * U + k * ULO + ulo) * 256 + c, x[hd][k]);
        }
    }
}

#include "ntt_fold.cuh"

// ---- host side ---------------------------------------------------------------------------------------
int env_int(const char *name, int dflt) {
    const char *v = std::getenv(name);
    return v ? std::atoi(v) : dflt;
}

typedef void (*kern_t)(const u64 *, u64 *, BigArgs);

struct BigPlan {
    bool covered = false;
    int cw_fwd = GCW, cw_inv = GCW, cw_gather = GCW, vt = 2;
    kern_t cols_fwd = nullptr, cols_inv = nullptr, cols_inv_gather = nullptr;
    int col_threads_fwd = 0, col_threads_inv = 0, col_threads_gather = 0;
    size_t chunk_units = 0;   // units per launch group so that a chunk's intermediate stays in L2
    bool fold_rows = true, fold_cols = true;   // B200_FOLD_ROWS / B200_FOLD_COLS = 0 keep a pass on the Shoup kernels (A/B timing)
    // fold-arithmetic column kernels of the shape (null: none, the row pass then folds what the Shoup column pass left)
    kern_t cols_fwd_f = nullptr, cols_inv_f = nullptr, cols_inv_gather_f = nullptr;
    int row_l0 = -1;                           // number of the row pass's first layer after cols_fwd_f
    int cw_bal = 4;                            // columns per CTA of the balanced kernels
    bool bal = false;                          // fold column kernels are the balanced form: CW*16 threads, unpadded tile
};

#define STD_G(L, A, B, C) case L: plan->cols_fwd = (kern_t)k_cols_std_g<L, A, B, C, false>; plan->cols_inv = (kern_t)k_cols_std_g<L, A, B, C, true>; \
                                  plan->cols_inv_gather = (kern_t)k_cols_std_g<L, A, B, C, true, true>; \
                                  plan->cols_fwd_f = (kern_t)k_cols_std_g_f<L, A, B, C, false>; plan->cols_inv_f = (kern_t)k_cols_std_g_f<L, A, B, C, true>; \
                                  plan->cols_inv_gather_f = (kern_t)k_cols_std_g_f<L, A, B, C, true, true>; plan->row_l0 = L; \
                                  plan->col_threads_fwd = plan->col_threads_inv = plan->col_threads_gather = std_col_threads(L); break;
#define M3_G(L, A, B, C) case L: plan->cols_fwd = (kern_t)k_cols_m3_g<L, A, B, C, false>; plan->cols_inv = (kern_t)k_cols_m3_g<L, A, B, C, true>; \
                                 plan->cols_inv_gather = (kern_t)k_cols_m3_g<L, A, B, C, true, true>; \
                                 plan->cols_fwd_f = (kern_t)k_cols_m3_g_f<L, A, B, C, false>; plan->cols_inv_f = (kern_t)k_cols_m3_g_f<L, A, B, C, true>; \
                                 plan->cols_inv_gather_f = (kern_t)k_cols_m3_g_f<L, A, B, C, true, true>; plan->row_l0 = L; \
                                 plan->col_threads_fwd = plan->col_threads_inv = plan->col_threads_gather = m3_col_threads(L); break;
// column-pass kernels for a ring shape: the hand-tuned ones for the two headline shapes, the generic family otherwise
bool pick_col_kernels(BigPlan *plan, bool matrix, int log, bool force_generic) {
    if (!matrix) {
        if (log == 8 && !force_generic) {
            plan->cols_fwd = (kern_t)k_cols_fwd_std16<B200_CW_STD_FWD>; plan->cols_inv = (kern_t)k_cols_inv_std16<B200_CW_STD_INV>;
            plan->cols_inv_gather = (kern_t)k_cols_inv_std16<B200_CW_GATHER, true>;
            plan->cw_fwd = B200_CW_STD_FWD; plan->cw_inv = B200_CW_STD_INV; plan->cw_gather = B200_CW_GATHER;
            plan->col_threads_fwd = B200_CW_STD_FWD * 32; plan->col_threads_inv = B200_CW_STD_INV * 32; plan->col_threads_gather = B200_CW_GATHER * 32;
            plan->cols_fwd_f = (kern_t)k_cols_fwd_std16_f<B200_CW_STD_FWD>; plan->cols_inv_f = (kern_t)k_cols_inv_std16_f<B200_CW_STD_INV>;
            plan->cols_inv_gather_f = (kern_t)k_cols_inv_std16_f<B200_CW_GATHER, true>;
            plan->row_l0 = STD16_ROW_L0;
            return true;
        }
        switch (log) {
            STD_G(2, 1, 1, 0) STD_G(3, 2, 1, 0) STD_G(4, 2, 2, 0) STD_G(5, 3, 2, 0) STD_G(6, 3, 3, 0) STD_G(7, 3, 2, 2) STD_G(8, 3, 3, 2) STD_G(9, 3, 3, 3)
            default: return false;
        }
    } else {
        if (log == 6 && !force_generic) {
            plan->cols_fwd = (kern_t)k_cols_fwd_m3_15<B200_CW_M3_FWD>; plan->cols_inv = (kern_t)k_cols_inv_m3_15<B200_CW_M3_INV>;
            plan->cols_inv_gather = (kern_t)k_cols_inv_m3_15<B200_CW_GATHER, true>;
            plan->cw_fwd = B200_CW_M3_FWD; plan->cw_inv = B200_CW_M3_INV; plan->cw_gather = B200_CW_GATHER;
            plan->col_threads_fwd = B200_CW_M3_FWD * 48; plan->col_threads_inv = B200_CW_M3_INV * 48; plan->col_threads_gather = B200_CW_GATHER * 48;
            plan->cols_fwd_f = (kern_t)k_cols_fwd_m3_15_f<B200_CW_M3_FWD>; plan->cols_inv_f = (kern_t)k_cols_inv_m3_15_f<B200_CW_M3_INV>;
            plan->cols_inv_gather_f = (kern_t)k_cols_inv_m3_15_f<B200_CW_GATHER, true>;
            plan->row_l0 = M3_15_ROW_L0;
            if (env_int("B200_COLS_BAL", 1)) {   // balanced small-CTA column kernels (ntt_fold.cuh); the GATHER form keeps 16-column tiles
                const int cwb = env_int("B200_CW_BAL", B200_CW_BAL);
                if (cwb == 16) { plan->cols_fwd_f = (kern_t)k_cols_fwd_m3_15_b<16>; plan->cols_inv_f = (kern_t)k_cols_inv_m3_15_b<16>; plan->cw_bal = 16; }
                else if (cwb == 8) { plan->cols_fwd_f = (kern_t)k_cols_fwd_m3_15_b<8>; plan->cols_inv_f = (kern_t)k_cols_inv_m3_15_b<8>; plan->cw_bal = 8; }
                else { plan->cols_fwd_f = (kern_t)k_cols_fwd_m3_15_b<4>; plan->cols_inv_f = (kern_t)k_cols_inv_m3_15_b<4>; plan->cw_bal = 4; }
                plan->cols_inv_gather_f = (kern_t)k_cols_inv_m3_15_b<B200_CW_GATHER, true>;
                plan->bal = true;
            }
            return true;
        }
        switch (log) {
            M3_G(1, 0, 1, 0) M3_G(2, 0, 2, 0) M3_G(3, 0, 3, 0) M3_G(4, 1, 3, 0) M3_G(5, 0, 3, 2) M3_G(6, 1, 3, 2) M3_G(7, 1, 3, 3)
            default: return false;
        }
    }
    return true;
}
#undef STD_G
#undef M3_G

u64 unit_div_magic(const Span &sp) {
    // m = ceil(2^64/d): unit*m/2^64 = unit/d + e with 0 <= e < unit/2^64 < 2^-33 < 1/d, so the floor is exact
    if (sp.batch <= 1) return 0;
    return ~0ULL / (u64)sp.batch + 1;
}

template <typename K>
int set_smem(K kern, size_t bytes) {
    if (bytes > 48 * 1024) CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    return 0;
}

}  // namespace

int big_tables_create(b200_ctx *ctx) {
    BigPlan *plan = new BigPlan();
    ctx->big = plan;
    if (env_int("B200_NO_BIG", 0)) return 0;
    const bool matrix = ctx->ring_type == B200_RING_MATRIX;
    const bool generic = env_int("B200_GENERIC_COLS", 0) != 0;
    if (!matrix) plan->covered = ctx->logN >= 10 && ctx->logN <= 17 && pick_col_kernels(plan, false, ctx->logN - 8, generic);
    else plan->covered = ctx->b == 1 && ctx->a >= 10 && ctx->a <= 16 && pick_col_kernels(plan, true, ctx->a - 9, generic);
    if (!ctx->d_fwd_r || !ctx->d_inv_r) plan->covered = false;   // the row kernels read the row-pass copies of the tables
    plan->fold_rows = env_int("B200_FOLD_ROWS", 1) != 0;
    plan->fold_cols = env_int("B200_FOLD_COLS", 1) != 0;
    plan->vt = B200_VT_M3;   // 16-column tiles and 4-v row tiles were measured slower on B200 (fewer resident CTAs)
    int l2 = 0;
    cudaDeviceGetAttribute(&l2, cudaDevAttrL2CacheSize, ctx->device);
    const size_t budget = (size_t)env_int("B200_CHUNK_MB", 512) << 20;
    plan->chunk_units = budget / ((size_t)ctx->N * sizeof(u64));
    if (plan->chunk_units < 1) plan->chunk_units = 1;
    if (plan->chunk_units > 65535) plan->chunk_units = 65535;   // units sit on gridDim.y in the column pass
    return 0;
}
void big_tables_destroy(b200_ctx *ctx) {
    delete (BigPlan *)ctx->big;
    ctx->big = nullptr;
}

// forward row kernel behind a fold column pass that ran radix-2 layers 0..l0-1: the fold schedule (ntt_fold.cuh) depends on
// the layer number modulo FWD_PERIOD only, so five instantiations serve every shape
template <bool MATRIX, int VT>
static kern_t rows_fwd_after_fold_cols(int l0) {
    static_assert(FWD_PERIOD == 5, "one instantiation per residue class of the first row layer");
    switch ((l0 - 1) % 5) {
        case 0: return (kern_t)k_rows_fwd_f<MATRIX, VT, 1>;
        case 1: return (kern_t)k_rows_fwd_f<MATRIX, VT, 2>;
        case 2: return (kern_t)k_rows_fwd_f<MATRIX, VT, 3>;
        case 3: return (kern_t)k_rows_fwd_f<MATRIX, VT, 4>;
        default: return (kern_t)k_rows_fwd_f<MATRIX, VT, 5>;
    }
}
typedef void (*mulkern_t)(u64 *, const u64 *, BigArgs);
static mulkern_t rows_mul_after_fold_cols(int l0) {
    switch ((l0 - 1) % 5) {
        case 0: return k_rows_mul_f<1>;
        case 1: return k_rows_mul_f<2>;
        case 2: return k_rows_mul_f<3>;
        case 3: return k_rows_mul_f<4>;
        default: return k_rows_mul_f<5>;
    }
}

int launch_ntt_big(b200_ctx *ctx, const u64 *in, u64 *out, Span sp, bool inverse, int flags, const PeerOut *peers) {
    BigPlan *plan = (BigPlan *)ctx->big;
    if (!plan || !plan->covered) return 1;
    const bool gather = peers && peers->n > 0;
    if (gather && !inverse) return 1;   // only the inverse transform has a fused all-gather
    const bool matrix = ctx->ring_type == B200_RING_MATRIX;
    const bool permute = matrix && !(flags & B200_TREE_ORDER);
    const bool fold_rows = ctx->fold_ok && plan->fold_rows;
    const bool fold_cols = fold_rows && plan->fold_cols && plan->cols_fwd_f;   // both passes on signed words
    const bool bal = fold_cols && plan->bal;
    const int N = ctx->N, R = N / 256, vt = plan->vt;
    const int cw = inverse ? (gather ? plan->cw_gather : (bal ? plan->cw_bal : plan->cw_inv)) : (bal ? plan->cw_bal : plan->cw_fwd);
    BigArgs a{ctx->d_consts, ctx->d_fwd, ctx->d_inv, ctx->d_misc, N, R, matrix ? ctx->a - 9 : 0, sp, 0, flags, 0, 0};
    a.div_magic = unit_div_magic(sp);
    a.sol = ctx->d_sol; a.fwd_s = ctx->d_fwd_s; a.inv_s = ctx->d_inv_s; a.misc_s = ctx->d_misc_s;
    a.fwd_r = ctx->d_fwd_r; a.inv_r = ctx->d_inv_r;
    kern_t ck = inverse ? (gather ? plan->cols_inv_gather : plan->cols_inv) : plan->cols_fwd;
    if (fold_cols) ck = inverse ? (gather ? plan->cols_inv_gather_f : plan->cols_inv_f) : plan->cols_fwd_f;
    int col_threads = inverse ? (gather ? plan->col_threads_gather : plan->col_threads_inv) : plan->col_threads_fwd;
    if (bal) col_threads = cw * 16;
    kern_t rk;
    int row_warps;
    if (matrix) {
        rk = inverse ? (kern_t)k_rows_inv<true, B200_VT_M3> : (kern_t)k_rows_fwd<true, B200_VT_M3>;
        if (fold_rows) rk = inverse ? (kern_t)k_rows_inv_f<true, B200_VT_M3, false> : (kern_t)k_rows_fwd_f<true, B200_VT_M3, -1>;
        if (fold_cols) rk = inverse ? (kern_t)k_rows_inv_f<true, B200_VT_M3, true> : rows_fwd_after_fold_cols<true, B200_VT_M3>(plan->row_l0);
        row_warps = 6 * vt;
    } else {
        rk = inverse ? (kern_t)k_rows_inv<false, B200_VT_STD> : (kern_t)k_rows_fwd<false, B200_VT_STD>;
        if (fold_rows) rk = inverse ? (kern_t)k_rows_inv_f<false, B200_VT_STD, false> : (kern_t)k_rows_fwd_f<false, B200_VT_STD, -1>;
        if (fold_cols) rk = inverse ? (kern_t)k_rows_inv_f<false, B200_VT_STD, true> : rows_fwd_after_fold_cols<false, B200_VT_STD>(plan->row_l0);
        row_warps = B200_VT_STD;
    }
    const int rows_total = R * cw;
    const size_t col_smem = (bal ? (size_t)rows_total : (size_t)(rows_total + ((rows_total >> 6) << 3) + 8)) * sizeof(u64) + ((size_t)env_int("B200_COL_SMEM_EXTRA_KB", 0) << 10);
    size_t row_smem = (size_t)row_warps * WROW * sizeof(u64);
    if (matrix) row_smem = std::max(row_smem, (size_t)(256 * (6 * vt + 1) + 1) * sizeof(u64));
    row_smem += (size_t)env_int("B200_ROW_SMEM_EXTRA_KB", 0) << 10;   // occupancy experiments only
    if (int rc = set_smem(ck, col_smem)) return rc;
    if (int rc = set_smem(rk, row_smem)) return rc;
    const int units = sp.batch * sp.nl;
    const int chunk = (int)std::min<size_t>(plan->chunk_units, (size_t)units);
    u64 *tmp = out;
    // a slot permutation makes the row pass write rows other CTAs still have to read: go through a buffer
    const bool dense = permute && (!inverse || in == out);
    if (dense) {
        if (int rc = b200_ensure_scratch(ctx, (size_t)chunk * N)) return rc;
        tmp = ctx->d_scratch;
    }
    const int row_tiles = R / row_warps, col_tiles = 256 / cw;
    for (int u0 = 0; u0 < units; u0 += chunk) {
        const int nu = std::min(chunk, units - u0);
        a.unit0 = u0;
        BigArgs first = a, second = a;
        first.out_dense = dense;
        second.in_dense = dense;
        if (gather) second.peers = *peers;
        const dim3 cgrid(col_tiles, nu);   // tiles run together, units on y (chunk_units <= 65535)
        if (!inverse) {
            { LaunchScope ls(ctx, matrix ? "cols_fwd_m3" : "cols_fwd_std", nu);
              ck<<<cgrid, col_threads, col_smem, b200_cur_stream(ctx)>>>(in, tmp, first); }
            { LaunchScope ls(ctx, matrix ? "rows_fwd_m3" : "rows_fwd_std", nu);
              rk<<<dim3(nu, row_tiles), row_warps * 32, row_smem, b200_cur_stream(ctx)>>>(tmp, out, second); }
        } else {
            { LaunchScope ls(ctx, matrix ? "rows_inv_m3" : "rows_inv_std", nu);
              rk<<<dim3(nu, row_tiles), row_warps * 32, row_smem, b200_cur_stream(ctx)>>>(in, tmp, first); }
            { LaunchScope ls(ctx, matrix ? "cols_inv_m3" : "cols_inv_std", nu);
              ck<<<cgrid, col_threads, col_smem, b200_cur_stream(ctx)>>>(tmp, out, second); }
        }
    }
    CUDA_TRY(cudaGetLastError());
    return 0;
}

// c = a * b in the ring: cols_fwd(a), cols_fwd(b) -> fused rows (fwd, fwd, multiply, inv) -> cols_inv, in chunks
// whose two intermediates (dense ring buffers in the context scratch) stay small enough to live in L2.
int launch_polymul_big(b200_ctx *ctx, const u64 *pa, const u64 *pb, u64 *pc, Span sp) {
    BigPlan *plan = (BigPlan *)ctx->big;
    if (!plan || !plan->covered) return 1;
    const bool matrix = ctx->ring_type == B200_RING_MATRIX;
    const int N = ctx->N, R = N / 256;
    const int flags = matrix ? B200_TREE_ORDER : 0;
    BigArgs a{ctx->d_consts, ctx->d_fwd, ctx->d_inv, ctx->d_misc, N, R, matrix ? ctx->a - 9 : 0, sp, 0, flags, 0, 0};
    a.div_magic = unit_div_magic(sp);
    a.sol = ctx->d_sol; a.fwd_s = ctx->d_fwd_s; a.inv_s = ctx->d_inv_s; a.misc_s = ctx->d_misc_s;
    a.fwd_r = ctx->d_fwd_r; a.inv_r = ctx->d_inv_r;
    // fold arithmetic end to end when the ring qualifies: fold column kernels on both sides of k_rows_mul_f
    const bool fold = ctx->fold_ok && plan->fold_rows && plan->fold_cols && plan->cols_fwd_f && plan->cols_inv_f && env_int("B200_FOLD_MUL", 1);
    const kern_t cf = fold ? plan->cols_fwd_f : plan->cols_fwd, ci = fold ? plan->cols_inv_f : plan->cols_inv;
    const bool bal = fold && plan->bal;
    const int cwf = bal ? plan->cw_bal : plan->cw_fwd, cwi = bal ? plan->cw_bal : plan->cw_inv;
    const int thr_f = bal ? cwf * 16 : plan->col_threads_fwd, thr_i = bal ? cwi * 16 : plan->col_threads_inv;
    auto col_smem_of = [R, bal](int cw) { const int rows_total = R * cw; return (bal ? (size_t)rows_total : (size_t)(rows_total + ((rows_total >> 6) << 3) + 8)) * sizeof(u64); };
    const size_t smem_f = col_smem_of(cwf), smem_i = col_smem_of(cwi);
    const size_t row_smem = (size_t)4 * WROW * sizeof(u64);
    if (int rc = set_smem(cf, smem_f)) return rc;
    if (int rc = set_smem(ci, smem_i)) return rc;
    const int units = sp.batch * sp.nl;
    const size_t budget = (size_t)env_int("B200_MUL_CHUNK_MB", 512) << 20;   // per intermediate
    const int chunk = (int)std::max<size_t>(1, std::min<size_t>(std::min<size_t>(budget / ((size_t)N * sizeof(u64)), 65535), (size_t)units));
    if (int rc = b200_ensure_scratch(ctx, (size_t)2 * chunk * N)) return rc;
    u64 *ta = ctx->d_scratch, *tb = ctx->d_scratch + (size_t)chunk * N;
    const int tiles_f = 256 / cwf, tiles_i = 256 / cwi, row_tiles = R / 4;
    cudaStream_t st = b200_cur_stream(ctx);
    for (int u0 = 0; u0 < units; u0 += chunk) {
        const int nu = std::min(chunk, units - u0);
        a.unit0 = u0;
        BigArgs to_dense = a, dense = a, from_dense = a;
        to_dense.out_dense = 1;
        dense.in_dense = dense.out_dense = 1;
        from_dense.in_dense = 1;
        if (fold) from_dense.flags |= B200_FLAG_MONT_SCALE;   // k_rows_mul_f multiplies Montgomery-wise: the inverse column pass scales by N^-1 * 2^64
        const dim3 gf(tiles_f, nu), gi(tiles_i, nu);
        { LaunchScope ls(ctx, matrix ? "cols_fwd_m3" : "cols_fwd_std", nu);
          cf<<<gf, thr_f, smem_f, st>>>(pa, ta, to_dense); }
        { LaunchScope ls(ctx, matrix ? "cols_fwd_m3" : "cols_fwd_std", nu);
          cf<<<gf, thr_f, smem_f, st>>>(pb, tb, to_dense); }
        { LaunchScope ls(ctx, "rows_mul", nu);
          if (!fold) k_rows_mul<<<dim3(nu, row_tiles), 128, row_smem, st>>>(ta, tb, dense);
          else rows_mul_after_fold_cols(plan->row_l0)<<<dim3(nu, row_tiles), 128, row_smem, st>>>(ta, tb, dense); }
        { LaunchScope ls(ctx, matrix ? "cols_inv_m3" : "cols_inv_std", nu);
          ci<<<gi, thr_i, smem_i, st>>>(ta, pc, from_dense); }
    }
    CUDA_TRY(cudaGetLastError());
    return 0;
}
