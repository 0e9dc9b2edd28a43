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

namespace {

struct BigArgs {
    const LimbConst *consts;
    const tw_t *fwd, *inv, *misc;
    int N, R, logU;   // R = N/256 rows; MATRIX: U = R/6 = 2^logU rows per block
    Span sp;
    int unit0, flags;
    int in_dense, out_dense;   // the pointer is a dense [chunk][N] ring buffer indexed by blockIdx.y
};

// word offsets of this CTA's unit in the input and output buffers, and its prime
struct UnitRef { size_t in, out; int limb; };
__device__ __forceinline__ UnitRef unit_ref(const BigArgs &a) {
    const int unit = a.unit0 + blockIdx.y;
    const int l = unit % a.sp.nl;
    const size_t layout = ((size_t)(unit / a.sp.nl) * a.sp.Lbuf + l) * (size_t)a.N;
    const size_t dense = (size_t)blockIdx.y * (size_t)a.N;
    return {a.in_dense ? dense : layout, a.out_dense ? dense : layout, a.sp.limb0 + l};
}

// Harvey butterflies.  forward: in [0,4q) -> out [0,4q);  inverse: in [0,2q) -> out [0,2q)
__device__ __forceinline__ void bf_fwd(u64 &U, u64 &V, const tw_t w, u64 q, u64 twoq) {
    const u64 u = csub(U, twoq);
    const u64 t = shoup_lazy(V, w.x, w.y, q);
    U = u + t;
    V = u - t + twoq;
}
__device__ __forceinline__ void bf_inv(u64 &U, u64 &V, const tw_t w, u64 q, u64 twoq) {
    const u64 s = U + V, d = U - V + twoq;
    U = csub(s, twoq);
    V = shoup_lazy(d, w.x, w.y, q);
}

// 2^LOGM-point sub-transform held in registers; x[k] ordered along the sub-dimension (MSB first);
// P = heap index of the sub-transform's root node.
template <int LOGM>
__device__ __forceinline__ void fwd_regs(u64 (&x)[1 << LOGM], unsigned P, const tw_t *__restrict__ tw, u64 q, u64 twoq) {
#pragma unroll
    for (int j = 0; j < LOGM; ++j) {
        const int half = (1 << LOGM) >> (j + 1);
#pragma unroll
        for (int g = 0; g < (1 << j); ++g) {
            const tw_t w = tw[(P << j) + g];
#pragma unroll
            for (int o = 0; o < half; ++o) bf_fwd(x[g * 2 * half + o], x[g * 2 * half + o + half], w, q, twoq);
        }
    }
}
template <int LOGM, int JMIN = 0>
__device__ __forceinline__ void inv_regs(u64 (&x)[1 << LOGM], unsigned P, const tw_t *__restrict__ tw, u64 q, u64 twoq) {
#pragma unroll
    for (int j = LOGM - 1; j >= JMIN; --j) {
        const int half = (1 << LOGM) >> (j + 1);
#pragma unroll
        for (int g = 0; g < (1 << j); ++g) {
            const tw_t w = tw[(P << j) + g];
#pragma unroll
            for (int o = 0; o < half; ++o) bf_inv(x[g * 2 * half + o], x[g * 2 * half + o + half], w, q, twoq);
        }
    }
}

__device__ __forceinline__ void ldg256(const u64 *p, u64 &a, u64 &b, u64 &c, u64 &d) {
    asm volatile("ld.global.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(p) : "memory");
}
__device__ __forceinline__ void stg256(u64 *p, u64 a, u64 b, u64 c, u64 d) {
    asm volatile("st.global.v4.u64 [%0], {%1,%2,%3,%4};" ::"l"(p), "l"(a), "l"(b), "l"(c), "l"(d) : "memory");
}

__device__ __forceinline__ int swz(int c) { return c + 2 * (c >> 4); }   // 16-word groups padded to 18
constexpr int ROW_PITCH = 288;                                            // 256 words swizzled

__device__ __forceinline__ unsigned rev_bits(unsigned x, int bits) { return bits ? (__brev(x) >> (32 - bits)) : 0u; }

// ====================================================================================================
// row pass: 8 radix-2 layers inside 256-word rows.  ROWS = 16 (STANDARD) or 24 (MATRIX: 6 blocks x 4 v).
// ====================================================================================================
template <int ROWS, bool MATRIX>
__device__ __forceinline__ int row_of(const BigArgs &a, int tile, int row, int &rowcode) {
    if (!MATRIX) { rowcode = 0; return tile * ROWS + row; }
    const int hd = row >> 2, vi = row & 3;
    const int u = (int)rev_bits((unsigned)(tile * 4 + vi), a.logU);
    rowcode = vi * 6 + 2 * (hd % 3) + hd / 3;          // position inside a 24-run of the reference order
    return (hd << a.logU) + u;
}

template <int ROWS, bool MATRIX>
__global__ void __launch_bounds__(ROWS * 16)
k_rows_fwd(const u64 *in, u64 *out, BigArgs a) {
    extern __shared__ __align__(16) u64 sm[];
    const UnitRef ur = unit_ref(a);
    const int limb = ur.limb;
    const u64 q = a.consts[limb].q, twoq = 2 * q;
    const tw_t *__restrict__ tw = a.fwd + (size_t)limb * a.N;
    const int t = threadIdx.x, row = t >> 4, g = t & 15;
    int rowcode;
    const int r = row_of<ROWS, MATRIX>(a, blockIdx.x, row, rowcode);
    const u64 *src = in + ur.in + (size_t)r * 256;
    u64 x[16];
    // group 1: layers on c bits 7..4; thread owns c = k*16 + g
#pragma unroll
    for (int k = 0; k < 16; ++k) x[k] = src[k * 16 + g];
    const unsigned P1 = (unsigned)(a.R + r);
    fwd_regs<4>(x, P1, tw, q, twoq);
#pragma unroll
    for (int k = 0; k < 16; ++k) sm[row * ROW_PITCH + swz(k * 16 + g)] = x[k];
    __syncthreads();
    // group 2: layers on c bits 3..0; thread owns c = g*16 + k (contiguous)
    {
        const ulonglong2 *s2 = reinterpret_cast<const ulonglong2 *>(sm + row * ROW_PITCH + g * 18);
#pragma unroll
        for (int k = 0; k < 8; ++k) { const ulonglong2 v = s2[k]; x[2 * k] = v.x; x[2 * k + 1] = v.y; }
    }
    fwd_regs<4>(x, P1 * 16 + g, tw, q, twoq);
    if (!(a.flags & B200_LAZY)) {
#pragma unroll
        for (int k = 0; k < 16; ++k) x[k] = csub(csub(x[k], twoq), q);
    }
    if (!MATRIX || (a.flags & B200_TREE_ORDER)) {
        u64 *dst = out + ur.out + (size_t)r * 256 + g * 16;
#pragma unroll
        for (int k = 0; k < 4; ++k) stg256(dst + 4 * k, x[4 * k], x[4 * k + 1], x[4 * k + 2], x[4 * k + 3]);
    } else {
        // reference order: slot = rev8(c)*6U + v*6 + (2d+h); stage the tile as [rev8(c)][24] (pitch 25)
        __syncthreads();
        const int rg = (int)rev_bits((unsigned)g, 4);
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            const int cr = (int)rev_bits((unsigned)k, 4) * 16 + rg;   // rev8(g*16 + k)
            sm[cr * 25 + rowcode] = x[k];
        }
        __syncthreads();
        const size_t run = (size_t)6 << a.logU;                       // 6U words between consecutive rev8(c)
        u64 *dst = out + ur.out + (size_t)blockIdx.x * 24;
        for (int o = t; o < 256 * 24; o += ROWS * 16) {
            const int cr = o / 24, rem = o - cr * 24;
            dst[(size_t)cr * run + rem] = sm[cr * 25 + rem];
        }
    }
}

template <int ROWS, bool MATRIX>
__global__ void __launch_bounds__(ROWS * 16)
k_rows_inv(const u64 *in, u64 *out, BigArgs a) {
    extern __shared__ __align__(16) u64 sm[];
    const UnitRef ur = unit_ref(a);
    const int limb = ur.limb;
    const u64 q = a.consts[limb].q, twoq = 2 * q;
    const tw_t *__restrict__ tw = a.inv + (size_t)limb * a.N;
    const int t = threadIdx.x, row = t >> 4, g = t & 15;
    int rowcode;
    const int r = row_of<ROWS, MATRIX>(a, blockIdx.x, row, rowcode);
    u64 x[16];
    if (!MATRIX || (a.flags & B200_TREE_ORDER)) {
        const u64 *src = in + ur.in + (size_t)r * 256 + g * 16;
#pragma unroll
        for (int k = 0; k < 4; ++k) ldg256(src + 4 * k, x[4 * k], x[4 * k + 1], x[4 * k + 2], x[4 * k + 3]);
    } else {
        const size_t run = (size_t)6 << a.logU;
        const u64 *src = in + ur.in + (size_t)blockIdx.x * 24;
        for (int o = t; o < 256 * 24; o += ROWS * 16) {
            const int cr = o / 24, rem = o - cr * 24;
            sm[cr * 25 + rem] = src[(size_t)cr * run + rem];
        }
        __syncthreads();
        const int rg = (int)rev_bits((unsigned)g, 4);
#pragma unroll
        for (int k = 0; k < 16; ++k) x[k] = sm[((int)rev_bits((unsigned)k, 4) * 16 + rg) * 25 + rowcode];
        __syncthreads();
    }
    const unsigned P1 = (unsigned)(a.R + r);
    inv_regs<4>(x, P1 * 16 + g, tw, q, twoq);
    {
        ulonglong2 *s2 = reinterpret_cast<ulonglong2 *>(sm + row * ROW_PITCH + g * 18);
#pragma unroll
        for (int k = 0; k < 8; ++k) s2[k] = make_ulonglong2(x[2 * k], x[2 * k + 1]);
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 16; ++k) x[k] = sm[row * ROW_PITCH + swz(k * 16 + g)];
    inv_regs<4>(x, P1, tw, q, twoq);
    u64 *dst = out + ur.out + (size_t)r * 256;
#pragma unroll
    for (int k = 0; k < 16; ++k) dst[k * 16 + g] = x[k];
}

// ====================================================================================================
// column pass, STANDARD rings: SA + SB = log2(N) - 8 layers; a CTA owns 16 columns of all R rows
// ====================================================================================================
template <int SA, int SB>
__global__ void __launch_bounds__(16 << (SA > SB ? SA : SB))
k_cols_fwd_std(const u64 *in, u64 *out, BigArgs a) {
    extern __shared__ __align__(16) u64 sm[];
    const UnitRef ur = unit_ref(a);
    const int limb = ur.limb;
    const size_t ibase = ur.in + (size_t)blockIdx.x * 16, obase = ur.out + (size_t)blockIdx.x * 16;
    const u64 q = a.consts[limb].q, twoq = 2 * q;
    const tw_t *__restrict__ tw = a.fwd + (size_t)limb * a.N;
    constexpr int MA = 1 << SA, MB = 1 << SB;
    for (int t = threadIdx.x; t < 16 * MB; t += blockDim.x) {   // group A: rows k*MB + r_lo
        const int c = t & 15, rlo = t >> 4;
        u64 x[MA];
#pragma unroll
        for (int k = 0; k < MA; ++k) x[k] = in[ibase + (size_t)(k * MB + rlo) * 256 + c];
        fwd_regs<SA>(x, 1u, tw, q, twoq);
#pragma unroll
        for (int k = 0; k < MA; ++k) sm[(k * MB + rlo) * 16 + c] = x[k];
    }
    __syncthreads();
    for (int t = threadIdx.x; t < 16 * MA; t += blockDim.x) {   // group B: rows r_hi*MB + k
        const int c = t & 15, rhi = t >> 4;
        u64 x[MB];
#pragma unroll
        for (int k = 0; k < MB; ++k) x[k] = sm[(rhi * MB + k) * 16 + c];
        fwd_regs<SB>(x, (unsigned)(MA + rhi), tw, q, twoq);
#pragma unroll
        for (int k = 0; k < MB; ++k) out[obase + (size_t)(rhi * MB + k) * 256 + c] = x[k];
    }
}

template <int SA, int SB>
__global__ void __launch_bounds__(16 << (SA > SB ? SA : SB))
k_cols_inv_std(const u64 *in, u64 *out, BigArgs a) {
    extern __shared__ __align__(16) u64 sm[];
    const UnitRef ur = unit_ref(a);
    const int limb = ur.limb;
    const size_t ibase = ur.in + (size_t)blockIdx.x * 16, obase = ur.out + (size_t)blockIdx.x * 16;
    const u64 q = a.consts[limb].q, twoq = 2 * q;
    const tw_t *__restrict__ tw = a.inv + (size_t)limb * a.N;
    const tw_t ninv = a.misc[limb * MISC_COUNT + MISC_NINV], ninv_w1 = a.misc[limb * MISC_COUNT + MISC_NINV_W1];
    const bool lazy = a.flags & B200_LAZY;
    constexpr int MA = 1 << SA, MB = 1 << SB;
    for (int t = threadIdx.x; t < 16 * MA; t += blockDim.x) {   // group B first (small strides)
        const int c = t & 15, rhi = t >> 4;
        u64 x[MB];
#pragma unroll
        for (int k = 0; k < MB; ++k) x[k] = in[ibase + (size_t)(rhi * MB + k) * 256 + c];
        inv_regs<SB>(x, (unsigned)(MA + rhi), tw, q, twoq);
#pragma unroll
        for (int k = 0; k < MB; ++k) sm[(rhi * MB + k) * 16 + c] = x[k];
    }
    __syncthreads();
    for (int t = threadIdx.x; t < 16 * MB; t += blockDim.x) {   // group A, last layer carries N^-1
        const int c = t & 15, rlo = t >> 4;
        u64 x[MA];
#pragma unroll
        for (int k = 0; k < MA; ++k) x[k] = sm[(k * MB + rlo) * 16 + c];
        inv_regs<SA, 1>(x, 1u, tw, q, twoq);
#pragma unroll
        for (int o = 0; o < MA / 2; ++o) {
            const u64 U = x[o], V = x[o + MA / 2];
            u64 X = shoup_lazy(U + V, ninv.x, ninv.y, q);
            u64 Y = shoup_lazy(U - V + twoq, ninv_w1.x, ninv_w1.y, q);
            if (!lazy) { X = csub(X, q); Y = csub(Y, q); }
            x[o] = X; x[o + MA / 2] = Y;
        }
#pragma unroll
        for (int k = 0; k < MA; ++k) out[obase + (size_t)(k * MB + rlo) * 256 + c] = x[k];
    }
}

// ====================================================================================================
// column pass, MATRIX rings (b = 1): cyclotomic + radix-3 + SA radix-2 layers in registers (6*2^SA words
// per thread), exchange, SB radix-2 layers.  Rows are (hd, u): hd in [0,6) = 3h + d, u in [0, 2^(SA+SB)).
// ====================================================================================================
template <int SA, int SB>
__global__ void __launch_bounds__(SA == 0 ? 16 << SB : 96 << SA)
k_cols_fwd_m3(const u64 *in, u64 *out, BigArgs a) {
    extern __shared__ __align__(16) u64 sm[];
    const UnitRef ur = unit_ref(a);
    const int limb = ur.limb;
    const size_t ibase = ur.in + (size_t)blockIdx.x * 16, obase = ur.out + (size_t)blockIdx.x * 16;
    const u64 q = a.consts[limb].q, twoq = 2 * q, fourq = 4 * q;
    const tw_t *__restrict__ tw = a.fwd + (size_t)limb * a.N;
    const tw_t om = a.misc[limb * MISC_COUNT + MISC_OMEGA3];
    constexpr int MA = 1 << SA, MB = 1 << SB, U = MA * MB;
    for (int t = threadIdx.x; t < 16 * MB; t += blockDim.x) {
        const int c = t & 15, ulo = t >> 4;
        u64 x[6][MA];
#pragma unroll
        for (int hd = 0; hd < 6; ++hd)
#pragma unroll
            for (int k = 0; k < MA; ++k) x[hd][k] = in[ibase + (size_t)(hd * U + k * MB + ulo) * 256 + c];
        const tw_t zeta = tw[1];
#pragma unroll
        for (int k = 0; k < MA; ++k) {
#pragma unroll
            for (int d = 0; d < 3; ++d) {   // cyclotomic layer: in [0,2q) -> [0,4q)
                const u64 lo = x[d][k], hi = x[d + 3][k];
                const u64 tt = shoup_lazy(hi, zeta.x, zeta.y, q);
                x[d][k] = lo + tt;
                x[d + 3][k] = csub(lo + hi + twoq - tt, fourq);
            }
#pragma unroll
            for (int h = 0; h < 2; ++h) {   // radix-3 layer: [0,4q) -> [0,4q)
                const tw_t z1 = tw[2 + 2 * h], z2 = tw[3 + 2 * h];
                const u64 b0 = csub(x[3 * h][k], twoq);
                const u64 t1 = shoup_lazy(x[3 * h + 1][k], z1.x, z1.y, q);
                const u64 t2 = shoup_lazy(x[3 * h + 2][k], z2.x, z2.y, q);
                const u64 t3 = shoup_lazy(t1 + twoq - t2, om.x, om.y, q);
                x[3 * h][k] = csub(b0 + t1 + t2, fourq);
                x[3 * h + 1][k] = csub(b0 + twoq - t2 + t3, fourq);
                x[3 * h + 2][k] = csub(b0 + fourq - t1 - t3, fourq);
            }
        }
#pragma unroll
        for (int hd = 0; hd < 6; ++hd) {
            if (SA > 0) fwd_regs<SA>(x[hd], (unsigned)(6 + hd), tw, q, twoq);
#pragma unroll
            for (int k = 0; k < MA; ++k) sm[(hd * U + k * MB + ulo) * 16 + c] = x[hd][k];
        }
    }
    __syncthreads();
    for (int t = threadIdx.x; t < 16 * 6 * MA; t += blockDim.x) {
        const int c = t & 15, top = t >> 4;   // top = hd*MA + u_top
        u64 x[MB];
#pragma unroll
        for (int k = 0; k < MB; ++k) x[k] = sm[(top * MB + k) * 16 + c];
        fwd_regs<SB>(x, (unsigned)(6 * MA + top), tw, q, twoq);
#pragma unroll
        for (int k = 0; k < MB; ++k) out[obase + (size_t)(top * MB + k) * 256 + c] = x[k];
    }
}

template <int SA, int SB>
__global__ void __launch_bounds__(SA == 0 ? 16 << SB : 96 << SA)
k_cols_inv_m3(const u64 *in, u64 *out, BigArgs a) {
    extern __shared__ __align__(16) u64 sm[];
    const UnitRef ur = unit_ref(a);
    const int limb = ur.limb;
    const size_t ibase = ur.in + (size_t)blockIdx.x * 16, obase = ur.out + (size_t)blockIdx.x * 16;
    const u64 q = a.consts[limb].q, twoq = 2 * q, fourq = 4 * q;
    const tw_t *__restrict__ tw = a.inv + (size_t)limb * a.N;
    const tw_t *mc = a.misc + limb * MISC_COUNT;
    const tw_t om = mc[MISC_OMEGA3], k1 = mc[MISC_K1], k2 = mc[MISC_K2], zeta = mc[MISC_ZETA];
    const bool lazy = a.flags & B200_LAZY;
    constexpr int MA = 1 << SA, MB = 1 << SB, U = MA * MB;
    for (int t = threadIdx.x; t < 16 * 6 * MA; t += blockDim.x) {
        const int c = t & 15, top = t >> 4;
        u64 x[MB];
#pragma unroll
        for (int k = 0; k < MB; ++k) x[k] = in[ibase + (size_t)(top * MB + k) * 256 + c];
        inv_regs<SB>(x, (unsigned)(6 * MA + top), tw, q, twoq);
#pragma unroll
        for (int k = 0; k < MB; ++k) sm[(top * MB + k) * 16 + c] = x[k];
    }
    __syncthreads();
    for (int t = threadIdx.x; t < 16 * MB; t += blockDim.x) {
        const int c = t & 15, ulo = t >> 4;
        u64 x[6][MA];
#pragma unroll
        for (int hd = 0; hd < 6; ++hd) {
#pragma unroll
            for (int k = 0; k < MA; ++k) x[hd][k] = sm[(hd * U + k * MB + ulo) * 16 + c];
            if (SA > 0) inv_regs<SA>(x[hd], (unsigned)(6 + hd), tw, q, twoq);
        }
#pragma unroll
        for (int k = 0; k < MA; ++k) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {   // inverse radix-3 (factor 3 deferred): [0,2q) -> [0,2q)
                const tw_t z1 = tw[2 + 2 * h], z2 = tw[3 + 2 * h];
                const u64 y0 = x[3 * h][k], y1 = x[3 * h + 1][k], y2 = x[3 * h + 2][k];
                const u64 s = shoup_lazy(y1 + twoq - y2, om.x, om.y, q);
                x[3 * h][k] = csub(csub(y0 + y1 + y2, fourq), twoq);
                x[3 * h + 1][k] = shoup_lazy(y0 + fourq - y1 - s, z1.x, z1.y, q);
                x[3 * h + 2][k] = shoup_lazy(y0 + twoq - y2 + s, z2.x, z2.y, q);
            }
#pragma unroll
            for (int d = 0; d < 3; ++d) {   // inverse cyclotomic layer with the 2/N scaling
                const u64 lo = x[d][k], hi = x[d + 3][k];
                u64 hh = shoup_lazy(lo + twoq - hi, k1.x, k1.y, q);
                u64 ll = csub(shoup_lazy(lo, k2.x, k2.y, q) + twoq - shoup_lazy(hh, zeta.x, zeta.y, q), twoq);
                if (!lazy) { ll = csub(ll, q); hh = csub(hh, q); }
                x[d][k] = ll; x[d + 3][k] = hh;
            }
        }
#pragma unroll
        for (int hd = 0; hd < 6; ++hd)
#pragma unroll
            for (int k = 0; k < MA; ++k) out[obase + (size_t)(hd * U + k * MB + ulo) * 256 + c] = x[hd][k];
    }
}

// ---- host side ---------------------------------------------------------------------------------------
struct BigPlan {
    bool covered = false;
    int SA = 0, SB = 0;
    size_t chunk_units = 0;   // units per launch group so that a chunk's intermediate stays in L2
};

int env_int(const char *name, int dflt) {
    const char *v = std::getenv(name);
    return v ? std::atoi(v) : dflt;
}

template <typename K>
int set_smem(K kern, size_t bytes) {
    if (bytes > 48 * 1024) CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    return 0;
}

typedef void (*kern_t)(const u64 *, u64 *, BigArgs);

struct ColKernel { kern_t k; int threads; };

ColKernel col_kernel(bool matrix, bool inverse, int SA, int SB) {
#define STD(A, B) if (!matrix && SA == A && SB == B) return {inverse ? (kern_t)k_cols_inv_std<A, B> : (kern_t)k_cols_fwd_std<A, B>, 16 << (A > B ? A : B)};
#define M3(A, B) if (matrix && SA == A && SB == B) return {inverse ? (kern_t)k_cols_inv_m3<A, B> : (kern_t)k_cols_fwd_m3<A, B>, A == 0 ? 16 << B : 96 << A};
    STD(2, 2) STD(3, 2) STD(3, 3) STD(4, 3) STD(4, 4)
    M3(0, 3) M3(0, 4) M3(1, 4) M3(2, 4)
#undef STD
#undef M3
    return {nullptr, 0};
}

}  // namespace

int big_tables_create(b200_ctx *ctx) {
    BigPlan *plan = new BigPlan();
    ctx->big = plan;
    const int min_std = env_int("B200_BIG_MIN_LOGN", 15);
    const int min_a = env_int("B200_BIG_MIN_A", 13);
    if (ctx->ring_type == B200_RING_STANDARD) {
        const int s = ctx->logN - 8;
        if (ctx->logN >= 12 && ctx->logN <= 16 && ctx->logN >= min_std) {
            plan->covered = true;
            plan->SA = (s + 1) / 2;
            plan->SB = s / 2;
        }
    } else if (ctx->b == 1 && ctx->a >= 12 && ctx->a <= 15 && ctx->a >= min_a) {
        plan->covered = true;
        plan->SB = ctx->a == 12 ? 3 : 4;
        plan->SA = ctx->a - 9 - plan->SB;
    }
    if (plan->covered && !col_kernel(ctx->ring_type == B200_RING_MATRIX, false, plan->SA, plan->SB).k) plan->covered = false;
    // chunking: keep (in + intermediate + out) of one chunk well inside L2
    int l2 = 0;
    cudaDeviceGetAttribute(&l2, cudaDevAttrL2CacheSize, ctx->device);
    const size_t budget = (size_t)env_int("B200_CHUNK_MB", l2 > 0 ? (l2 >> 20) / 4 : 24) << 20;
    plan->chunk_units = budget / ((size_t)ctx->N * sizeof(u64));
    if (plan->chunk_units < 1) plan->chunk_units = 1;
    return 0;
}
void big_tables_destroy(b200_ctx *ctx) {
    delete (BigPlan *)ctx->big;
    ctx->big = nullptr;
}

int launch_ntt_big(b200_ctx *ctx, const u64 *in, u64 *out, Span sp, bool inverse, int flags) {
    BigPlan *plan = (BigPlan *)ctx->big;
    if (!plan || !plan->covered) return 1;
    const bool matrix = ctx->ring_type == B200_RING_MATRIX;
    const bool permute = matrix && !(flags & B200_TREE_ORDER);
    const int N = ctx->N, R = N / 256;
    BigArgs a{ctx->d_consts, ctx->d_fwd, ctx->d_inv, ctx->d_misc, N, R, matrix ? ctx->a - 9 : 0, sp, 0, flags, 0, 0};
    const ColKernel ck = col_kernel(matrix, inverse, plan->SA, plan->SB);
    const size_t col_smem = (size_t)R * 16 * sizeof(u64);
    const int rows = matrix ? 24 : 16;
    const size_t row_smem = (size_t)rows * ROW_PITCH * sizeof(u64);
    kern_t rk = matrix ? (inverse ? (kern_t)k_rows_inv<24, true> : (kern_t)k_rows_fwd<24, true>)
                       : (inverse ? (kern_t)k_rows_inv<16, false> : (kern_t)k_rows_fwd<16, false>);
    if (int rc = set_smem(ck.k, col_smem)) return rc;
    if (int rc = set_smem(rk, row_smem)) return rc;
    // intermediate buffer: in place on `out` unless a slot permutation would make the row pass overwrite
    // rows other CTAs still have to read
    const int units = sp.batch * sp.nl;
    const int chunk = (int)std::min<size_t>(std::min<size_t>(plan->chunk_units, 65535), (size_t)units);
    u64 *tmp = out;
    const bool dense = permute && (!inverse || in == out);
    if (dense) {   // a chunk-sized ring buffer, rewritten by every chunk: it never has to leave L2
        if (int rc = b200_ensure_scratch(ctx, (size_t)chunk * N)) return rc;
        tmp = ctx->d_scratch;
    }
    const int row_tiles = R / rows, col_tiles = 256 / 16;
    for (int u0 = 0; u0 < units; u0 += chunk) {
        const int nu = std::min(chunk, units - u0);
        a.unit0 = u0;
        BigArgs first = a, second = a;
        first.out_dense = dense;
        second.in_dense = dense;
        if (!inverse) {
            { LaunchScope ls(ctx, matrix ? "cols_fwd_m3" : "cols_fwd_std", nu);
              ck.k<<<dim3(col_tiles, nu), ck.threads, col_smem, ctx->stream>>>(in, tmp, first); }
            { LaunchScope ls(ctx, matrix ? "rows_fwd_m3" : "rows_fwd_std", nu);
              rk<<<dim3(row_tiles, nu), rows * 16, row_smem, ctx->stream>>>(tmp, out, second); }
        } else {
            { LaunchScope ls(ctx, matrix ? "rows_inv_m3" : "rows_inv_std", nu);
              rk<<<dim3(row_tiles, nu), rows * 16, row_smem, ctx->stream>>>(in, tmp, first); }
            { LaunchScope ls(ctx, matrix ? "cols_inv_m3" : "cols_inv_std", nu);
              ck.k<<<dim3(col_tiles, nu), ck.threads, col_smem, ctx->stream>>>(tmp, out, second); }
        }
    }
    CUDA_TRY(cudaGetLastError());
    return 0;
}

int launch_polymul_big(b200_ctx *, const u64 *, const u64 *, u64 *, Span) { return 1; }
