// keyswitch.cu -- RNS basis extension, ModDown and the hybrid key-switch gadget product.
//
// Reference: ring/basis_extension.go (GenModUpConstants :95-166, ModUpQtoP/ModUpPtoQ :184-217, ModDownQPtoQNTT
// :241-258, ModUpExact :282-308, Decomposer.DecomposeAndSplit :381-502, reconstructRNS[Centered] :504-594, multSum
// :597-673) and core/rlwe/evaluator_gadget_product.go (GadgetProduct :16-30, ModDown :33-45,
// gadgetProductMultiplePLazy :123-188, DecomposeSingleNTT :455-478).
//
// The ring is ONE context over the concatenated chain Q_0..Q_{LQ-1}, P_0..P_{LP-1}; polynomials keep the usual
// [batch][LQ+LP][N] layout.  The extension is the only cross-limb step of the path: one thread per coefficient
// gathers the source limbs, runs the exact conversion (64-bit Montgomery products, 128-bit accumulation, the
// float64 correction index v computed with the same IEEE operations in the same order as the Go code) and writes
// every requested target limb.  Everything around it is the batched NTT / element-wise kernels of the hot path.
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <string>

#include "hostmath.hpp"
#include "internal.hpp"

namespace hm = hostmath;

namespace {

constexpr int EXT_THREADS = 128;
// Source limbs of one extension: the Go code's y[32] buffer (ring/basis_extension.go:529) is the same cap.  Below it the
// 128-bit accumulator cannot wrap (32 q^2 < 2^127 for q < 2^61) and the generic kernel's ns KiB of shared memory fits the
// 48 KiB a kernel gets without opting in.
constexpr int EXT_MAX_SRC = 32;

// device-side description of one extension: source limbs -> target limbs
struct ExtDesc {
    int ns, nt;
    const int *src_limb, *dst_limb;        // ring limb index (== layout limb index)
    const u64 *half;                       // [ns]  addend before the reconstruction (Q/2 mod q_i), 0 if not centred
    const u64 *qoverqiinvqi;               // [ns]  Montgomery form
    const u64 *qoverqimodp;                // [nt][ns] Montgomery form
    const u64 *vtimesqmodp;                // [nt][ns+1]
    const u64 *sub;                        // [nt]  subtracted after the conversion (Q/2 mod p_j), 0 if not centred
};

// mode 0: exactly the lazy words ModUpQtoP/ModUpPtoQ leave (multSum, then SubScalar = CRed(x + p - s));
// mode 1: the same residue reduced to [0, p) (feeds the transforms inside the gadget product / ModDown)
__global__ void __launch_bounds__(EXT_THREADS)
k_basis_extend(const u64 *src, u64 *dst, const LimbConst *__restrict__ consts, ExtDesc d, int N, int Lbuf, int mode) {
    extern __shared__ u64 ysm[];   // [ns][EXT_THREADS]
    const int b = blockIdx.y;
    const int x = blockIdx.x * EXT_THREADS + threadIdx.x;
    if (x >= N) return;
    const u64 *s = src + (size_t)b * Lbuf * N + x;
    u64 *o = dst + (size_t)b * Lbuf * N + x;
    double vi = 0.0;
    for (int i = 0; i < d.ns; ++i) {                                    // reconstructRNS[Centered]
        const LimbConst c = consts[d.src_limb[i]];
        const u64 y = mred(s[(size_t)d.src_limb[i] * N] + d.half[i], d.qoverqiinvqi[i], c.q, c.qinv);
        ysm[i * EXT_THREADS + threadIdx.x] = y;
        vi = __dadd_rn(vi, __ddiv_rn((double)y, (double)c.q));          // no contraction, same order as the Go loop
    }
    const unsigned v = (unsigned)(u64)vi;
    for (int j = 0; j < d.nt; ++j) {                                    // multSum
        const LimbConst c = consts[d.dst_limb[j]];
        const u64 *row = d.qoverqimodp + (size_t)j * d.ns;
        u64 rlo = 0, rhi = 0;
        for (int i = 0; i < d.ns; ++i) {
            const u64 y = ysm[i * EXT_THREADS + threadIdx.x], k = row[i];
            const u64 mlo = y * k, mhi = __umul64hi(y, k);
            rlo += mlo;
            rhi += mhi + (rlo < mlo);
        }
        u64 r = rhi - __umul64hi(rlo * c.qinv, c.q) + c.q + d.vtimesqmodp[(size_t)j * (d.ns + 1) + v];   // [0, 3p)
        r = csub(r + c.q - d.sub[j], c.q);                                                             // SubScalar
        // canonical form: r < 4p only while the accumulator's high word is below 2p, i.e. ns <= 16 for 61-bit primes
        // (rhi < ns p^2 / 2^64 <= ns p / 8); longer digits go through the Barrett reduction of any 64-bit word
        if (mode == 1) r = d.ns <= 16 ? csub(csub(r, 2 * c.q), c.q) : bred_add(r, c.q, c.brc_hi);
        o[(size_t)d.dst_limb[j] * N] = r;
    }
}

// Specialisation for the digit sizes the key switch uses (NS <= 8 source limbs): all per-prime constants of the CTA
// staged once in shared memory, the source residues in registers, two coefficients per thread so that every constant
// read feeds two 128-bit multiply-accumulates.  Same arithmetic, same order of the float64 operations.
#ifndef B200_EXT_CPT
#define B200_EXT_CPT 2
#endif
constexpr int EXT_CPT = B200_EXT_CPT;   // coefficients per thread
#ifndef B200_EXT_ADDBACK
#define B200_EXT_ADDBACK 0            // 1: final reductions as subtract + sign-based add-back (measured 1.7 % slower here: 1.501 vs 1.476 ms)
#endif
#ifndef B200_EXT_MINB
#define B200_EXT_MINB 5                 // CTAs per SM the register allocation must allow (128-thread CTAs; measured 1..8 on B200: 5 is 4 % faster than 1)
#endif
template <int NS>
__global__ void __launch_bounds__(EXT_THREADS, B200_EXT_MINB)
k_basis_extend_ns(const u64 *src, u64 *dst, const LimbConst *__restrict__ consts, ExtDesc d, int N, int Lbuf, int mode) {
    extern __shared__ __align__(16) u64 csm[];
    // layout: src: q[NS] qinv[NS] half[NS] qinvqi[NS] qd[NS] | per target: q qinv sub k[NS] vq[NS+1]
    u64 *sq = csm, *sqinv = sq + NS, *shalf = sqinv + NS, *sfac = shalf + NS;
    double *sqd = reinterpret_cast<double *>(sfac + NS);
    u64 *tg = reinterpret_cast<u64 *>(sqd + NS);
    constexpr int TSTRIDE = 3 + NS + NS + 1;
    for (int i = threadIdx.x; i < NS; i += EXT_THREADS) {
        const LimbConst c = consts[d.src_limb[i]];
        sq[i] = c.q; sqinv[i] = c.qinv; shalf[i] = d.half[i]; sfac[i] = d.qoverqiinvqi[i]; sqd[i] = (double)c.q;
    }
    for (int t = threadIdx.x; t < d.nt * TSTRIDE; t += EXT_THREADS) {
        const int j = t / TSTRIDE, f = t - j * TSTRIDE;
        u64 v;
        if (f == 0) v = consts[d.dst_limb[j]].q;
        else if (f == 1) v = consts[d.dst_limb[j]].qinv;
        else if (f == 2) v = d.sub[j];
        else if (f < 3 + NS) v = d.qoverqimodp[(size_t)j * NS + (f - 3)];
        else v = d.vtimesqmodp[(size_t)j * (NS + 1) + (f - 3 - NS)];
        tg[t] = v;
    }
    __syncthreads();
    const int b = blockIdx.y;
    const int x0 = blockIdx.x * (EXT_THREADS * EXT_CPT) + threadIdx.x;
    const u64 *s = src + (size_t)b * Lbuf * N;
    u64 *o = dst + (size_t)b * Lbuf * N;
    u64 y[EXT_CPT][NS];
    unsigned v[EXT_CPT];
    bool live[EXT_CPT];
#pragma unroll
    for (int e = 0; e < EXT_CPT; ++e) {
        const int x = x0 + e * EXT_THREADS;
        live[e] = x < N;
        double vi = 0.0;
#pragma unroll
        for (int i = 0; i < NS; ++i) {                                      // reconstructRNS[Centered]
            const u64 w = live[e] ? s[(size_t)d.src_limb[i] * N + x] : 0;
            y[e][i] = mred(w + shalf[i], sfac[i], sq[i], sqinv[i]);
            vi = __dadd_rn(vi, __ddiv_rn((double)y[e][i], sqd[i]));         // no contraction, same order as the Go loop
        }
        v[e] = (unsigned)(u64)vi;
    }
    for (int j = 0; j < d.nt; ++j) {                                        // multSum
        const u64 *t = tg + j * TSTRIDE;
        const u64 q = t[0], qinv = t[1], sub = t[2];
        // 128-bit sum of the NS products y_i * k_i by columns of 32-bit partial products (each computed once):
        //   A  = sum y0*k0 (64 bits + carry count cc),  S1 = sum (y0*k1 + y1*k0),  S2 = sum y1*k1
        // y_i, k_i < 2^61 so their high words are < 2^29: S2 cannot overflow, S1 holds 4 sources (8 terms < 2^61).
        u64 A[EXT_CPT], S1a[EXT_CPT], S1b[EXT_CPT], S2[EXT_CPT];
        unsigned cc[EXT_CPT];
#pragma unroll
        for (int e = 0; e < EXT_CPT; ++e) { A[e] = S1a[e] = S1b[e] = S2[e] = 0; cc[e] = 0; }
#pragma unroll
        for (int i = 0; i < NS; ++i) {
            const u64 k = t[3 + i];
            const unsigned k0 = lo32(k), k1 = hi32(k);
#pragma unroll
            for (int e = 0; e < EXT_CPT; ++e) {
                const unsigned y0 = lo32(y[e][i]), y1 = hi32(y[e][i]);
                const u64 p00 = mulwide_u32(y0, k0);
                asm("add.cc.u64 %0, %0, %2;\n\taddc.u32 %1, %1, 0;" : "+l"(A[e]), "+r"(cc[e]) : "l"(p00));
                if (i < 4) S1a[e] = madwide_u32(y1, k0, madwide_u32(y0, k1, S1a[e]));
                else S1b[e] = madwide_u32(y1, k0, madwide_u32(y0, k1, S1b[e]));
                S2[e] = madwide_u32(y1, k1, S2[e]);
            }
        }
        u64 *oj = o + (size_t)d.dst_limb[j] * N;
#pragma unroll
        for (int e = 0; e < EXT_CPT; ++e) {
            u64 rlo = A[e], rhi = (u64)cc[e] + S2[e] + (S1a[e] >> 32);
            const u64 sa = S1a[e] << 32;
            rlo += sa; rhi += rlo < sa;
            if (NS > 4) {
                const u64 sb = S1b[e] << 32;
                rlo += sb; rhi += (S1b[e] >> 32) + (rlo < sb);
            }
            u64 r = rhi - __umul64hi(rlo * qinv, q) + q + t[3 + NS + v[e]];   // [0, 3p)
#if B200_EXT_ADDBACK
            if (mode == 1) r = addback_neg(addback_neg(r - sub - q, 2 * q) - q, q);   // (r + q - sub) - 2q, then - q: [0,4p) -> [0,p)
            else r = addback_neg(r - sub, q);                                         // SubScalar = CRed(r + q - sub), the lazy word Go leaves
#else
            if (mode == 1) r = csub(csub(r + q - sub, 2 * q), q);               // the same residue, canonical: [0,4p) -> [0,p)
            else r = csub(r + q - sub, q);                                      // SubScalar, the lazy word Go leaves
#endif
            if (live[e]) oj[x0 + e * EXT_THREADS] = r;
        }
    }
}

// single-limb digit of DecomposeAndSplit (:402-428): centred lift of one limb, written canonically
__global__ void __launch_bounds__(256)
k_decompose_single(const u64 *src, u64 *dst, const LimbConst *__restrict__ consts, int src_limb, const int *dst_limb, int nt,
                   int N, int Lbuf) {
    const int b = blockIdx.y;
    const int x = blockIdx.x * 256 + threadIdx.x;
    if (x >= N) return;
    const u64 q = consts[src_limb].q;
    u64 coeff = src[((size_t)b * Lbuf + src_limb) * N + x];
    const bool neg = coeff >= (q >> 1);
    if (neg) coeff = q - coeff;
    for (int j = 0; j < nt; ++j) {
        const LimbConst c = consts[dst_limb[j]];
        const u64 t = bred_add(coeff, c.q, c.brc_hi);
        dst[((size_t)b * Lbuf + dst_limb[j]) * N + x] = neg ? csub(c.q - t, c.q) : t;
    }
}

// Power-of-two digit of gadgetProductSinglePAndBitDecompLazy (evaluator_gadget_product.go:232-235, ring.MaskVec
// vec_ops.go:870-890): cw = (cxInvNTT[src_limb] >> w) & mask, the same small integer for every limb it is transformed
// under (:241-246,:252-257); written reduced, which is the identity whenever 2^pw2 <= q.
__global__ void __launch_bounds__(256)
k_mask_spread(const u64 *src, u64 *dst, const LimbConst *__restrict__ consts, int src_limb, const int *dst_limb, int nt,
              int N, int Lbuf, int w, u64 mask) {
    const int b = blockIdx.y;
    const int x = blockIdx.x * 256 + threadIdx.x;
    if (x >= N) return;
    const u64 v = (src[((size_t)b * Lbuf + src_limb) * N + x] >> w) & mask;
    for (int j = 0; j < nt; ++j) {
        const LimbConst c = consts[dst_limb[j]];
        dst[((size_t)b * Lbuf + dst_limb[j]) * N + x] = bred_add(v, c.q, c.brc_hi);
    }
}

// acc_c[x] (+)= MRed(key[d][c][x], operand[x]) for both components; the operand of the digit's own limbs is the
// NTT-domain input itself (rlwe/evaluator_gadget_product.go:466-472), all others come from the decomposition.
// Up to MAC_DIGITS digits per pass: the accumulators are read (unless `first`) and written once for the whole group
// instead of once per digit (4 of every 7 words a per-digit pass moves).
constexpr int MAC_DIGITS = 4;
struct MacGroup {
    const u64 *op[MAC_DIGITS];             // decomposition of digit g (NTT domain), [batch][L][N]
    const u64 *k0[MAC_DIGITS], *k1[MAC_DIGITS];
    int st[MAC_DIGITS], ed[MAC_DIGITS];    // the digit's own limbs read the input itself
};
__device__ __forceinline__ void mac128(u64 &lo, u64 &hi, u64 x, u64 y) {
    const u64 pl = x * y;
    lo += pl;
    hi += __umul64hi(x, y) + (lo < pl);
}
// a (canonical) + (hi:lo) * 2^-64 mod q, canonical; hi < q
__device__ __forceinline__ u64 fold128(u64 a, u64 lo, u64 hi, u64 q, u64 qinv) {
    const u64 r = hi - __umul64hi(lo * qinv, q) + q;     // MRedLazy of the 128-bit sum: [0, 2q)
    return csub(csub(a + r, 2 * q), q);
}
#ifndef B200_MAC_CHUNKS
#define B200_MAC_CHUNKS 128           // 256 = one pair of words per thread up to N = 2^17; measured 7 % slower (0.715 vs 0.666 ms)
#endif
#ifndef B200_MAC_MINB
#define B200_MAC_MINB 1
#endif
template <int ND>
__global__ void __launch_bounds__(256, B200_MAC_MINB)
k_gadget_mac_group(const u64 *cx, MacGroup m, u64 *acc0, u64 *acc1, const LimbConst *__restrict__ consts, const int *limbs,
                   int nlimbs, int N, int Lbuf, int first) {
    // limb-major units: the polynomials of the batch that share a key limb run back to back, so the 2*ND key limbs
    // they all read (the larger half of the traffic) are served from L2 after the first one
    const int unit = blockIdx.y, batch = gridDim.y / nlimbs;
    const int b = unit % batch, limb = limbs[unit / batch];
    const LimbConst c = consts[limb];
    const size_t base = ((size_t)b * Lbuf + limb) * (size_t)N, kbase = (size_t)limb * N;
    const u64 *op[ND];
#pragma unroll
    for (int g = 0; g < ND; ++g) op[g] = (limb >= m.st[g] && limb < m.ed[g]) ? cx : m.op[g];
    for (int i = 2 * (blockIdx.x * blockDim.x + threadIdx.x); i < N; i += 2 * gridDim.x * blockDim.x) {
        ulonglong2 a0 = make_ulonglong2(0, 0), a1 = make_ulonglong2(0, 0);
        if (!first) {
            a0 = *reinterpret_cast<const ulonglong2 *>(acc0 + base + i);
            a1 = *reinterpret_cast<const ulonglong2 *>(acc1 + base + i);
        }
        // The ND products of one word are summed as 128-bit integers and reduced ONCE: sum < ND*q^2 < q*2^64 for
        // ND <= 4 and q < 2^61, so one Montgomery reduction returns the same residue as the ND reduced products of
        // MulCoeffsMontgomery[Lazy]ThenAdd[Lazy] (evaluator_gadget_product.go:160-176) at half the multiplications.
        u64 lo[4] = {0, 0, 0, 0}, hi[4] = {0, 0, 0, 0};
#pragma unroll
        for (int g = 0; g < ND; ++g) {
            const ulonglong2 v = *reinterpret_cast<const ulonglong2 *>(op[g] + base + i);
            const ulonglong2 x0 = *reinterpret_cast<const ulonglong2 *>(m.k0[g] + kbase + i);
            const ulonglong2 x1 = *reinterpret_cast<const ulonglong2 *>(m.k1[g] + kbase + i);
            mac128(lo[0], hi[0], x0.x, v.x);
            mac128(lo[1], hi[1], x0.y, v.y);
            mac128(lo[2], hi[2], x1.x, v.x);
            mac128(lo[3], hi[3], x1.y, v.y);
        }
        a0.x = fold128(a0.x, lo[0], hi[0], c.q, c.qinv);
        a0.y = fold128(a0.y, lo[1], hi[1], c.q, c.qinv);
        a1.x = fold128(a1.x, lo[2], hi[2], c.q, c.qinv);
        a1.y = fold128(a1.y, lo[3], hi[3], c.q, c.qinv);
        *reinterpret_cast<ulonglong2 *>(acc0 + base + i) = a0;
        *reinterpret_cast<ulonglong2 *>(acc1 + base + i) = a1;
    }
}

struct Extender {
    int ns = 0, nt = 0;
    int *d_src = nullptr, *d_dst = nullptr;
    u64 *d_half = nullptr, *d_qinvqi = nullptr, *d_qmodp = nullptr, *d_vq = nullptr, *d_sub = nullptr;
    ExtDesc desc() const { return ExtDesc{ns, nt, d_src, d_dst, d_half, d_qinvqi, d_qmodp, d_vq, d_sub}; }
    void release() {
        cudaFree(d_src); cudaFree(d_dst); cudaFree(d_half); cudaFree(d_qinvqi); cudaFree(d_qmodp); cudaFree(d_vq); cudaFree(d_sub);
    }
};

// (prod S) >> 1 mod m without big integers: prod S is odd, so ((prod S mod 2m) - 1) / 2 is the residue
u64 half_of_product_mod(const std::vector<u64> &S, u64 m) {
    const u64 m2 = 2 * m;
    u64 r = 1;
    for (u64 s : S) r = hm::mulmod(r, s % m2, m2);
    return ((r - 1) >> 1) % m;
}

}  // namespace

struct b200_keyswitch {
    b200_ctx *ctx = nullptr;
    int LQ = 0, LP = 0;
    std::map<std::string, Extender> cache;
    std::vector<std::vector<u64>> moddown;   // [levelP][i]: MForm((p_0..p_levelP)^-1 mod q_i)  (:24-49)
    u64 *ws = nullptr;                       // cxInv, c2, acc0, acc1: 4 polys of [batch][L][N]
    size_t ws_words = 0;
    u64 *d_moddown = nullptr;                // [LP][LQ]: q_i - moddown[levelP][i], the scalar of the final ModDown pass
    u64 *d_moddown_qp = nullptr;             // [LQ][LP]: p_j - MForm((q_0..q_levelQ)^-1 mod p_j), the same for ModDownQPtoP (modDownConstantsQtoP, :61)
    std::map<long long, int *> limb_lists;   // (q0, q1, p0, p1) -> device list of the limbs the multiply-accumulate touches
};

namespace {

template <typename T>
cudaError_t upload(T **dst, const std::vector<T> &v) {
    cudaError_t e = cudaMalloc(dst, sizeof(T) * std::max<size_t>(v.size(), 1));
    if (e != cudaSuccess) return e;
    return cudaMemcpy(*dst, v.data(), sizeof(T) * v.size(), cudaMemcpyHostToDevice);
}

// key of b200_keyswitch::limb_lists: four limb indices, 16 bits each (chains of any length the ring admits)
static inline long long limb_list_key(int q0, int q1, int p0, int p1) {
    return ((((long long)q0 << 16 | (long long)q1) << 16 | (long long)p0) << 16) | (long long)p1;
}

// GenModUpConstants (:95-166) for source limbs S and target limbs T of the ring, plus the centring scalars
int get_extender(b200_keyswitch *ks, const std::vector<int> &src, const std::vector<int> &dst, bool centred, Extender **out) {
    std::string key = centred ? "c" : "n";
    for (int s : src) key += "s" + std::to_string(s);
    for (int t : dst) key += "t" + std::to_string(t);
    auto it = ks->cache.find(key);
    if (it != ks->cache.end()) { *out = &it->second; return 0; }
    b200_ctx *ctx = ks->ctx;
    const int ns = (int)src.size(), nt = (int)dst.size();
    if (ns > EXT_MAX_SRC) { b200_set_error("basis extension: too many source limbs"); return B200_EINVAL; }
    std::vector<u64> S(ns), T(nt);
    for (int i = 0; i < ns; ++i) S[i] = ctx->sub[src[i]].q;
    for (int j = 0; j < nt; ++j) T[j] = ctx->sub[dst[j]].q;
    std::vector<u64> half(ns, 0), qinvqi(ns), qmodp((size_t)nt * ns), vq((size_t)nt * (ns + 1)), sub(nt, 0);
    for (int i = 0; i < ns; ++i) {
        u64 star = 1;
        for (int j = 0; j < ns; ++j) if (j != i) star = hm::mulmod(star, S[j] % S[i], S[i]);
        qinvqi[i] = hm::mform(hm::invmod_prime(star, S[i]), S[i]);
        if (centred) half[i] = half_of_product_mod(S, S[i]);
    }
    for (int j = 0; j < nt; ++j) {
        const u64 pj = T[j];
        for (int i = 0; i < ns; ++i) {
            u64 star = 1;
            for (int u = 0; u < ns; ++u) if (u != i) star = hm::mulmod(star, S[u] % pj, pj);
            qmodp[(size_t)j * ns + i] = hm::mform(star, pj);
        }
        u64 qmod = 1;
        for (int i = 0; i < ns; ++i) qmod = hm::mulmod(qmod, S[i] % pj, pj);
        const u64 v = pj - qmod;
        vq[(size_t)j * (ns + 1)] = 0;
        for (int i = 1; i <= ns; ++i) { u64 t = vq[(size_t)j * (ns + 1) + i - 1] + v; vq[(size_t)j * (ns + 1) + i] = t >= pj ? t - pj : t; }
        if (centred) sub[j] = half_of_product_mod(S, pj);
    }
    Extender e;
    e.ns = ns; e.nt = nt;
    cudaError_t ce = upload(&e.d_src, src);
    if (ce == cudaSuccess) ce = upload(&e.d_dst, dst);
    if (ce == cudaSuccess) ce = upload(&e.d_half, half);
    if (ce == cudaSuccess) ce = upload(&e.d_qinvqi, qinvqi);
    if (ce == cudaSuccess) ce = upload(&e.d_qmodp, qmodp);
    if (ce == cudaSuccess) ce = upload(&e.d_vq, vq);
    if (ce == cudaSuccess) ce = upload(&e.d_sub, sub);
    if (ce != cudaSuccess) { e.release(); return b200_cuda_fail(ce, "basis extension constants"); }
    *out = &(ks->cache[key] = e);
    return 0;
}

int run_extend(b200_keyswitch *ks, Extender *e, const u64 *src, u64 *dst, int batch, int mode) {
    b200_ctx *ctx = ks->ctx;
    const int N = ctx->N;
    LaunchScope ls(ctx, "basis_extend", (long long)batch * e->nt);
    cudaStream_t st = b200_cur_stream(ctx);
    const int ns = e->ns;
    const size_t csm = sizeof(u64) * ((size_t)5 * ns + (size_t)e->nt * (3 + 2 * ns + 1));
    const dim3 grid2((N + EXT_THREADS * EXT_CPT - 1) / (EXT_THREADS * EXT_CPT), batch);
    static const bool generic_only = std::getenv("B200_EXT_GENERIC") != nullptr;
#define EXT_NS(K) case K: k_basis_extend_ns<K><<<grid2, EXT_THREADS, csm, st>>>(src, dst, ctx->d_consts, e->desc(), N, ctx->L, mode); break;
    if (ns >= 1 && ns <= 8 && csm <= 48 * 1024 && !generic_only) {
        switch (ns) { EXT_NS(1) EXT_NS(2) EXT_NS(3) EXT_NS(4) EXT_NS(5) EXT_NS(6) EXT_NS(7) EXT_NS(8) }
    } else {
        k_basis_extend<<<dim3((N + EXT_THREADS - 1) / EXT_THREADS, batch), EXT_THREADS, sizeof(u64) * ns * EXT_THREADS, st>>>(
            src, dst, ctx->d_consts, e->desc(), N, ctx->L, mode);
    }
#undef EXT_NS
    CUDA_TRY(cudaGetLastError());
    return 0;
}

int ntt_range(b200_ctx *ctx, const u64 *src, u64 *dst, int batch, int limb0, int nl, bool inverse) {
    if (nl <= 0) return 0;
    const size_t off = (size_t)limb0 * ctx->N;
    Span sp{batch, nl, ctx->L, limb0};
    int rc = launch_ntt(ctx, src + off, dst + off, sp, inverse, 0);
    return rc;
}

int ensure_ws(b200_keyswitch *ks, size_t words) {
    if (ks->ws_words >= words) return 0;
    CUDA_TRY(cudaStreamSynchronize(b200_cur_stream(ks->ctx)));
    if (ks->ws) CUDA_TRY(cudaFree(ks->ws));
    ks->ws = nullptr; ks->ws_words = 0;
    CUDA_TRY(cudaMalloc(&ks->ws, words * sizeof(u64)));
    ks->ws_words = words;
    return 0;
}

std::vector<int> range(int a, int b) { std::vector<int> v; for (int i = a; i < b; ++i) v.push_back(i); return v; }

int check_levels(const b200_keyswitch *ks, int levelQ, int levelP, const char *who) {
    if (!ks) { b200_set_error(std::string(who) + ": NULL key-switch context"); return B200_EINVAL; }
    if (levelQ < 0 || levelQ >= ks->LQ || levelP < 0 || levelP >= ks->LP) { b200_set_error(std::string(who) + ": level out of range"); return B200_ESHAPE; }
    return 0;
}
int check_poly(const b200_keyswitch *ks, const b200_poly *p, const char *who) {
    if (!p || p->ctx != ks->ctx) { b200_set_error(std::string(who) + ": polynomial does not belong to the QP ring"); return B200_ESHAPE; }
    return 0;
}

// A rank's share of the chain: Q limbs [q0, q1) and P limbs [p0, p1) (P indices, i.e. ring limbs LQ+p0 ..).
// One GPU owns everything: {0, levelQ+1, 0, levelP+1}.
struct Own { int q0, q1, p0, p1; };

// Final step of ModDownQPtoQNTT (:241-258) on the owned Q limbs.  acc: Q limbs in the NTT domain; its P limbs
// 0..levelP already in the COEFFICIENT domain (ringP.INTTLazy done, all P limbs present).  buf: scratch polynomial.
int mod_down_finish(b200_keyswitch *ks, int levelQ, int levelP, const u64 *acc, u64 *out, u64 *buf, int batch, Own o) {
    b200_ctx *ctx = ks->ctx;
    const int LQ = ks->LQ, nq = o.q1 - o.q0;
    if (nq <= 0) return 0;
    (void)levelQ;
    Extender *e;
    int rc = get_extender(ks, range(LQ, LQ + levelP + 1), range(o.q0, o.q1), true, &e);
    if (rc) return rc;
    if ((rc = run_extend(ks, e, acc, buf, batch, 1))) return rc;                                // ModUpPtoQ
    if ((rc = ntt_range(ctx, buf, buf, batch, o.q0, nq, false))) return rc;                     // ringQ.NTTLazy
    const size_t off = (size_t)o.q0 * ctx->N;
    return launch_vec_op(ctx, B200_OP_SUB_MULSCALAR_MONT_2Q, buf + off, acc + off, out + off, Span{batch, nq, ctx->L, o.q0},
                         ks->d_moddown + (size_t)levelP * LQ + o.q0);
}

// ModDownQPtoQNTT on one GPU: p1 (QP, NTT) -> out limbs 0..levelQ.  buf: a [batch][L][N] scratch polynomial.
int mod_down_ntt(b200_keyswitch *ks, int levelQ, int levelP, const u64 *p1, u64 *out, u64 *buf, int batch) {
    b200_ctx *ctx = ks->ctx;
    const int LQ = ks->LQ;
    // ringP.INTTLazy(p1P, buffP): into the P limbs of buf, then extend from buf; the Q limbs of p1 are read in place
    int rc = ntt_range(ctx, p1, buf, batch, LQ, levelP + 1, true);
    if (rc) return rc;
    Extender *e;
    if ((rc = get_extender(ks, range(LQ, LQ + levelP + 1), range(0, levelQ + 1), true, &e))) return rc;
    if ((rc = run_extend(ks, e, buf, buf, batch, 1))) return rc;
    if ((rc = ntt_range(ctx, buf, buf, batch, 0, levelQ + 1, false))) return rc;
    return launch_vec_op(ctx, B200_OP_SUB_MULSCALAR_MONT_2Q, buf, p1, out, Span{batch, levelQ + 1, ctx->L, 0},
                         ks->d_moddown + (size_t)levelP * LQ);
}

// one multiply-accumulate pass over a group of nd <= MAC_DIGITS digits
int launch_mac(b200_ctx *ctx, const u64 *cx, const MacGroup &mg, int nd, u64 *acc0, u64 *acc1, const int *d_limbs, int nl, int B, bool first) {
    const int N = ctx->N;
    int chunks = (N / 2 + 255) / 256;
    if (chunks > B200_MAC_CHUNKS) chunks = B200_MAC_CHUNKS;
    const dim3 grid(chunks, B * nl);
    LaunchScope ls(ctx, "gadget_mac", (long long)B * nl * nd);
    cudaStream_t st_ = b200_cur_stream(ctx);
    const int f = first ? 1 : 0;
    switch (nd) {
    case 1: k_gadget_mac_group<1><<<grid, 256, 0, st_>>>(cx, mg, acc0, acc1, ctx->d_consts, d_limbs, nl, N, ctx->L, f); break;
    case 2: k_gadget_mac_group<2><<<grid, 256, 0, st_>>>(cx, mg, acc0, acc1, ctx->d_consts, d_limbs, nl, N, ctx->L, f); break;
    case 3: k_gadget_mac_group<3><<<grid, 256, 0, st_>>>(cx, mg, acc0, acc1, ctx->d_consts, d_limbs, nl, N, ctx->L, f); break;
    default: k_gadget_mac_group<4><<<grid, 256, 0, st_>>>(cx, mg, acc0, acc1, ctx->d_consts, d_limbs, nl, N, ctx->L, f); break;
    }
    CUDA_TRY(cudaGetLastError());
    return 0;
}

// DecomposeSingleNTT (evaluator_gadget_product.go:455-478) of digit d restricted to the owned limbs, without the copy of
// the digit's own limbs (:466-468): DecomposeAndSplit (basis_extension.go:381-502) of the coefficient-domain input into every
// owned limb outside the digit and the owned P limbs, then their NTT.
int decompose_digit(b200_keyswitch *ks, int levelQ, int levelP, const u64 *cxinv, u64 *c2d, int B, Own o, int d) {
    b200_ctx *ctx = ks->ctx;
    const int nbPi = levelP + 1, LQ = ks->LQ, N = ctx->N;
    const int st = d * nbPi, ed = std::min(st + nbPi, levelQ + 1);
    std::vector<int> dst;
    for (int j = o.q0; j < o.q1; ++j) if (j < st || j >= ed) dst.push_back(j);
    for (int j = o.p0; j < o.p1; ++j) dst.push_back(LQ + j);
    if (dst.empty()) return 0;
    int rc = 0;
    if (ed - st == 1) {                                                                  // single-limb digit (:402-428)
        Extender *e;   // only used as a cached holder of the target list
        if ((rc = get_extender(ks, range(st, ed), dst, false, &e))) return rc;
        LaunchScope ls(ctx, "decompose_single", (long long)B * (int)dst.size());
        k_decompose_single<<<dim3((N + 255) / 256, B), 256, 0, b200_cur_stream(ctx)>>>(cxinv, c2d, ctx->d_consts, st, e->d_dst, (int)dst.size(), N, ctx->L);
    } else {                                                                             // exact centred conversion (:430-500)
        Extender *e;
        if ((rc = get_extender(ks, range(st, ed), dst, true, &e))) return rc;
        if ((rc = run_extend(ks, e, cxinv, c2d, B, 1))) return rc;
    }
    // NTT of every owned limb outside the digit, and of the owned P limbs (:469-477)
    const int a0 = o.q0, a1 = std::min(o.q1, st), b0 = std::max(o.q0, ed), b1 = o.q1;
    if ((rc = ntt_range(ctx, c2d, c2d, B, a0, a1 - a0, false))) return rc;
    if ((rc = ntt_range(ctx, c2d, c2d, B, b0, b1 - b0, false))) return rc;
    return ntt_range(ctx, c2d, c2d, B, LQ + o.p0, o.p1 - o.p0, false);
}

// Digit loop of gadgetProductMultiplePLazy (:123-188) restricted to the owned limbs: decomposition of every digit
// into the owned targets, their NTT, and the multiply-accumulate with the owned key limbs.  cxinv must hold ALL
// Q limbs 0..levelQ in the coefficient domain (after the all-gather when sharded); cx the owned Q limbs in the NTT domain.
// c2: min(beta, MAC_DIGITS) scratch polynomials of [B][L][N] words.
int decompose_mac(b200_keyswitch *ks, int levelQ, int levelP, const u64 *cx, const u64 *cxinv, const u64 *key, u64 *c2,
                  u64 *acc0, u64 *acc1, int B, Own o) {
    b200_ctx *ctx = ks->ctx;
    const int nbPi = levelP + 1, beta = (levelQ + levelP + 1) / (levelP + 1), LQ = ks->LQ, N = ctx->N;
    std::vector<int> limbs = range(o.q0, o.q1);
    for (int j = o.p0; j < o.p1; ++j) limbs.push_back(LQ + j);
    if (limbs.empty()) return 0;
    int *&d_limbs = ks->limb_lists[limb_list_key(o.q0, o.q1, o.p0, o.p1)];
    if (!d_limbs) {
        cudaError_t ce = upload(&d_limbs, limbs);
        if (ce != cudaSuccess) return b200_cuda_fail(ce, "gadget product limb list");
    }
    // digits are processed in groups of MAC_DIGITS: decomposition + NTT per digit into its own slot of c2, then ONE
    // multiply-accumulate pass over the group
    const size_t words = (size_t)B * ctx->L * N;
    int rc = 0;
    for (int d0 = 0; d0 < beta && rc == 0; d0 += MAC_DIGITS) {
        const int nd = std::min(MAC_DIGITS, beta - d0);
        MacGroup mg{};
        for (int g = 0; g < nd && rc == 0; ++g) {
            const int d = d0 + g;
            const int st = d * nbPi, ed = std::min(st + nbPi, levelQ + 1);
            u64 *c2d = c2 + (size_t)g * words;
            if ((rc = decompose_digit(ks, levelQ, levelP, cxinv, c2d, B, o, d))) break;
            mg.op[g] = c2d;
            mg.k0[g] = key + (size_t)(2 * d) * ctx->L * N;
            mg.k1[g] = key + (size_t)(2 * d + 1) * ctx->L * N;
            mg.st[g] = st; mg.ed[g] = ed;
        }
        if (rc) break;
        rc = launch_mac(ctx, cx, mg, nd, acc0, acc1, d_limbs, (int)limbs.size(), B, d0 == 0);
    }
    return rc;
}

// Digit loop of gadgetProductSinglePAndBitDecompLazy (:190-320) with a power-of-two base (mask != 0): for every Q limb i
// and every window j < sizes[i], the window of the coefficient-domain limb is spread over all limbs of QP, transformed and
// multiply-accumulated with key[i][j][c].  No basis extension.  The key holds 2 * sum(sizes) polynomials ordered [i][j][c].
int bitdecomp_mac(b200_keyswitch *ks, int levelQ, int levelP, int pw2, const int *sizes, const u64 *cxinv, const u64 *key,
                  u64 *c2, u64 *acc0, u64 *acc1, int B) {
    b200_ctx *ctx = ks->ctx;
    const int LQ = ks->LQ, N = ctx->N;
    std::vector<int> limbs = range(0, levelQ + 1);
    for (int j = 0; j <= levelP; ++j) limbs.push_back(LQ + j);
    int *&d_limbs = ks->limb_lists[limb_list_key(0, levelQ + 1, 0, levelP + 1)];
    if (!d_limbs) {
        cudaError_t ce = upload(&d_limbs, limbs);
        if (ce != cudaSuccess) return b200_cuda_fail(ce, "gadget product limb list");
    }
    const u64 mask = pw2 >= 64 ? ~0ULL : ((1ULL << pw2) - 1);
    const size_t words = (size_t)B * ctx->L * N;
    const int nl = (int)limbs.size();
    std::vector<std::pair<int, int>> digits;
    for (int i = 0; i <= levelQ; ++i)
        for (int j = 0; j < sizes[i]; ++j) digits.push_back({i, j});
    int rc = 0;
    for (size_t d0 = 0; d0 < digits.size() && rc == 0; d0 += MAC_DIGITS) {
        const int nd = (int)std::min<size_t>(MAC_DIGITS, digits.size() - d0);
        MacGroup mg{};
        for (int g = 0; g < nd && rc == 0; ++g) {
            const int i = digits[d0 + g].first, j = digits[d0 + g].second;
            u64 *c2d = c2 + (size_t)g * words;
            {
                LaunchScope ls(ctx, "mask_spread", (long long)B * nl);
                k_mask_spread<<<dim3((N + 255) / 256, B), 256, 0, b200_cur_stream(ctx)>>>(cxinv, c2d, ctx->d_consts, i, d_limbs, nl, N, ctx->L, j * pw2, mask);
            }
            if ((rc = ntt_range(ctx, c2d, c2d, B, 0, levelQ + 1, false))) break;
            if ((rc = ntt_range(ctx, c2d, c2d, B, LQ, levelP + 1, false))) break;
            mg.op[g] = c2d;
            mg.k0[g] = key + (size_t)(2 * (d0 + g)) * ctx->L * N;
            mg.k1[g] = key + (size_t)(2 * (d0 + g) + 1) * ctx->L * N;
            mg.st[g] = mg.ed[g] = 0;   // no limb reads the input itself
        }
        if (rc) break;
        rc = launch_mac(ctx, cxinv, mg, nd, acc0, acc1, d_limbs, nl, B, d0 == 0);
    }
    return rc;
}

int check_own(const b200_keyswitch *ks, int levelQ, int levelP, Own o, const char *who) {
    if (o.q0 < 0 || o.q1 > levelQ + 1 || o.q0 > o.q1 || o.p0 < 0 || o.p1 > levelP + 1 || o.p0 > o.p1) {
        b200_set_error(std::string(who) + ": owned limb ranges out of bounds");
        return B200_ESHAPE;
    }
    (void)ks;
    return 0;
}

}  // namespace

extern "C" int b200_keyswitch_create(b200_ctx *ctx, int LQ, b200_keyswitch **out) {
    CtxLock lk(ctx);
    if (!ctx || !out) { b200_set_error("b200_keyswitch_create: NULL argument"); return B200_EINVAL; }
    if (LQ < 1 || LQ >= ctx->L) { b200_set_error("b200_keyswitch_create: the ring must hold LQ >= 1 ciphertext primes followed by >= 1 special primes"); return B200_EINVAL; }
    b200_keyswitch *ks = new b200_keyswitch();
    ks->ctx = ctx; ks->LQ = LQ; ks->LP = ctx->L - LQ;
    ks->moddown.assign(ks->LP, std::vector<u64>(LQ));
    for (int i = 0; i < LQ; ++i) {   // genmodDownConstants (:24-49)
        const u64 qi = ctx->sub[i].q;
        u64 prod = 1;
        for (int j = 0; j < ks->LP; ++j) {
            prod = hm::mulmod(prod, ctx->sub[LQ + j].q % qi, qi);
            ks->moddown[j][i] = hm::mform(hm::invmod_prime(prod, qi), qi);
        }
    }
    cudaSetDevice(ctx->device);
    std::vector<u64> sc((size_t)ks->LP * LQ);
    for (int j = 0; j < ks->LP; ++j)
        for (int i = 0; i < LQ; ++i) sc[(size_t)j * LQ + i] = ctx->sub[i].q - ks->moddown[j][i];
    cudaError_t ce = upload(&ks->d_moddown, sc);
    if (ce != cudaSuccess) { delete ks; return b200_cuda_fail(ce, "b200_keyswitch_create"); }
    std::vector<u64> sp((size_t)LQ * ks->LP);   // genmodDownConstants(ringP, ringQ) (:61): [levelQ][j]
    for (int j = 0; j < ks->LP; ++j) {
        const u64 pj = ctx->sub[LQ + j].q;
        u64 prod = 1;
        for (int i = 0; i < LQ; ++i) {
            prod = hm::mulmod(prod, ctx->sub[i].q % pj, pj);
            sp[(size_t)i * ks->LP + j] = pj - hm::mform(hm::invmod_prime(prod, pj), pj);
        }
    }
    ce = upload(&ks->d_moddown_qp, sp);
    if (ce != cudaSuccess) { cudaFree(ks->d_moddown); delete ks; return b200_cuda_fail(ce, "b200_keyswitch_create"); }
    *out = ks;
    return 0;
}
extern "C" void b200_keyswitch_destroy(b200_keyswitch *ks) {
    if (!ks) return;
    cudaSetDevice(ks->ctx->device);
    for (auto &kv : ks->cache) kv.second.release();
    cudaFree(ks->ws); cudaFree(ks->d_moddown); cudaFree(ks->d_moddown_qp);
    for (auto &kv : ks->limb_lists) cudaFree(kv.second);
    delete ks;
}

// BasisExtender.ModUpQtoP (:184-198, direction 0) / ModUpPtoQ (:203-217, direction 1): lazy words exactly as Go leaves them
extern "C" int b200_mod_up(b200_keyswitch *ks, int direction, int levelQ, int levelP, const b200_poly *in, b200_poly *out) {
    CtxLock lk(ks ? ks->ctx : nullptr);
    if (int rc = check_levels(ks, levelQ, levelP, "b200_mod_up")) return rc;
    if (int rc = check_poly(ks, in, "b200_mod_up")) return rc;
    if (int rc = check_poly(ks, out, "b200_mod_up")) return rc;
    if (in->batch != out->batch) { b200_set_error("b200_mod_up: batch sizes differ"); return B200_ESHAPE; }
    CUDA_TRY(cudaSetDevice(ks->ctx->device));
    Extender *e;
    const std::vector<int> q = range(0, levelQ + 1), p = range(ks->LQ, ks->LQ + levelP + 1);
    if (int rc = get_extender(ks, direction == 0 ? q : p, direction == 0 ? p : q, true, &e)) return rc;
    return run_extend(ks, e, in->d, out->d, in->batch, 0);
}

extern "C" int b200_mod_down_ntt(b200_keyswitch *ks, int levelQ, int levelP, const b200_poly *p1, b200_poly *p2) {
    CtxLock lk(ks ? ks->ctx : nullptr);
    if (int rc = check_levels(ks, levelQ, levelP, "b200_mod_down_ntt")) return rc;
    if (int rc = check_poly(ks, p1, "b200_mod_down_ntt")) return rc;
    if (int rc = check_poly(ks, p2, "b200_mod_down_ntt")) return rc;
    if (p1->batch != p2->batch) { b200_set_error("b200_mod_down_ntt: batch sizes differ"); return B200_ESHAPE; }
    CUDA_TRY(cudaSetDevice(ks->ctx->device));
    const size_t words = (size_t)p1->batch * ks->ctx->L * ks->ctx->N;
    if (int rc = ensure_ws(ks, words)) return rc;
    return mod_down_ntt(ks, levelQ, levelP, p1->d, p2->d, ks->ws, p1->batch);
}

// BasisExtender.ModDownQPtoQ (:223-236, to_p == 0) / ModDownQPtoP (:264-279, to_p != 0): the coefficient-domain forms --
// exact extension of the dropped basis into the kept one, then (p1 - ext) * (dropped product)^-1 on the kept limbs.
extern "C" int b200_mod_down(b200_keyswitch *ks, int to_p, int levelQ, int levelP, const b200_poly *p1, b200_poly *p2) {
    CtxLock lk(ks ? ks->ctx : nullptr);
    if (int rc = check_levels(ks, levelQ, levelP, "b200_mod_down")) return rc;
    if (int rc = check_poly(ks, p1, "b200_mod_down")) return rc;
    if (int rc = check_poly(ks, p2, "b200_mod_down")) return rc;
    if (p1->batch != p2->batch) { b200_set_error("b200_mod_down: batch sizes differ"); return B200_ESHAPE; }
    b200_ctx *ctx = ks->ctx;
    CUDA_TRY(cudaSetDevice(ctx->device));
    const int LQ = ks->LQ, batch = p1->batch;
    if (int rc = ensure_ws(ks, (size_t)batch * ctx->L * ctx->N)) return rc;
    Extender *e;
    const std::vector<int> q = range(0, levelQ + 1), p = range(LQ, LQ + levelP + 1);
    if (int rc = get_extender(ks, to_p ? q : p, to_p ? p : q, true, &e)) return rc;
    if (int rc = run_extend(ks, e, p1->d, ks->ws, batch, 1)) return rc;
    if (!to_p)
        return launch_vec_op(ctx, B200_OP_SUB_MULSCALAR_MONT_2Q, ks->ws, p1->d, p2->d, Span{batch, levelQ + 1, ctx->L, 0},
                             ks->d_moddown + (size_t)levelP * LQ);
    const size_t off = (size_t)LQ * ctx->N;
    return launch_vec_op(ctx, B200_OP_SUB_MULSCALAR_MONT_2Q, ks->ws + off, p1->d + off, p2->d + off, Span{batch, levelP + 1, ctx->L, LQ},
                         ks->d_moddown_qp + (size_t)levelQ * ks->LP);
}

// rlwe.Evaluator.GadgetProduct (NTT-domain operands): evaluator_gadget_product.go:16-30,123-188,455-478.  With one special
// prime (levelP == 0) the reference takes gadgetProductSinglePAndBitDecompLazy (:190-320) with mask == 0, which is the same
// digit loop with one-limb digits (DecomposeAndSplit with nbPi = 1, :213-215) and gives the same canonical result.
extern "C" int b200_gadget_product(b200_keyswitch *ks, int levelQ, int levelP, const b200_poly *cx, const b200_poly *key,
                                   b200_poly *ct0, b200_poly *ct1) {
    CtxLock lk(ks ? ks->ctx : nullptr);
    if (int rc = check_levels(ks, levelQ, levelP, "b200_gadget_product")) return rc;
    for (const b200_poly *p : {cx, key, (const b200_poly *)ct0, (const b200_poly *)ct1})
        if (int rc = check_poly(ks, p, "b200_gadget_product")) return rc;
    b200_ctx *ctx = ks->ctx;
    const int beta = (levelQ + levelP + 1) / (levelP + 1), B = cx->batch, N = ctx->N;
    if (key->batch < 2 * beta) { b200_set_error("b200_gadget_product: the key must hold 2*beta polynomials [digit][component]"); return B200_ESHAPE; }
    if (ct0->batch != B || ct1->batch != B || ct0->d == ct1->d) { b200_set_error("b200_gadget_product: outputs must be two distinct polynomials of the input's batch size"); return B200_ESHAPE; }
    CUDA_TRY(cudaSetDevice(ctx->device));
    const size_t words = (size_t)B * ctx->L * N;
    const int slots = std::min(beta, MAC_DIGITS);
    if (int rc = ensure_ws(ks, (3 + slots) * words)) return rc;
    u64 *cxinv = ks->ws, *acc0 = ks->ws + words, *acc1 = ks->ws + 2 * words, *c2 = ks->ws + 3 * words;
    int rc = ntt_range(ctx, cx->d, cxinv, B, 0, levelQ + 1, true);                               // ringQ.INTT(cxNTT, cxInvNTT)
    if (rc) return rc;
    if ((rc = decompose_mac(ks, levelQ, levelP, cx->d, cxinv, key->d, c2, acc0, acc1, B, Own{0, levelQ + 1, 0, levelP + 1}))) return rc;
    if ((rc = mod_down_ntt(ks, levelQ, levelP, acc0, ct0->d, c2, B))) return rc;                 // eval.ModDown, NTT -> NTT (:40-45)
    return mod_down_ntt(ks, levelQ, levelP, acc1, ct1->d, c2, B);
}

// rlwe.Evaluator.DecomposeNTT (evaluator_gadget_product.go:430-452) for an NTT-domain input: the RNS decomposition of cx,
// every digit extended to all limbs 0..levelQ and P_0..levelP, in the NTT domain and canonical, the digit's own limbs copied
// from cx (:466-468).  `decomp` holds beta*B polynomials ordered [digit][b] (BuffDecompQP[digit]).
extern "C" int b200_decompose_ntt(b200_keyswitch *ks, int levelQ, int levelP, const b200_poly *cx, b200_poly *decomp) {
    CtxLock lk(ks ? ks->ctx : nullptr);
    if (int rc = check_levels(ks, levelQ, levelP, "b200_decompose_ntt")) return rc;
    if (int rc = check_poly(ks, cx, "b200_decompose_ntt")) return rc;
    if (int rc = check_poly(ks, decomp, "b200_decompose_ntt")) return rc;
    b200_ctx *ctx = ks->ctx;
    const int nbPi = levelP + 1, beta = (levelQ + levelP + 1) / (levelP + 1), B = cx->batch, N = ctx->N;
    if (decomp->batch < beta * B || decomp->d == cx->d) { b200_set_error("b200_decompose_ntt: decomp must hold beta*batch polynomials [digit][b], distinct from cx"); return B200_ESHAPE; }
    CUDA_TRY(cudaSetDevice(ctx->device));
    const size_t words = (size_t)B * ctx->L * N;
    if (int rc = ensure_ws(ks, words)) return rc;
    u64 *cxinv = ks->ws;
    int rc = ntt_range(ctx, cx->d, cxinv, B, 0, levelQ + 1, true);                               // ringQ.INTT(polyNTT, polyInvNTT) (:438-441)
    if (rc) return rc;
    const Own o{0, levelQ + 1, 0, levelP + 1};
    for (int d = 0; d < beta; ++d) {
        u64 *c2d = decomp->d + (size_t)d * words;
        if ((rc = decompose_digit(ks, levelQ, levelP, cxinv, c2d, B, o, d))) return rc;
        const int st = d * nbPi, ed = std::min(st + nbPi, levelQ + 1);
        // copy(c2QiQ.Coeffs[x], c2NTT.Coeffs[x]) for the digit's own limbs: (ed-st)*N contiguous words of each polynomial
        CUDA_TRY(cudaMemcpy2DAsync(c2d + (size_t)st * N, (size_t)ctx->L * N * sizeof(u64), cx->d + (size_t)st * N, (size_t)ctx->L * N * sizeof(u64),
                                   (size_t)(ed - st) * N * sizeof(u64), B, cudaMemcpyDeviceToDevice, b200_cur_stream(ctx)));
    }
    return 0;
}
// rlwe.Evaluator.GadgetProductHoisted (:322-339, :341-423): the multiply-accumulate of an already decomposed input with a
// gadget ciphertext, then ModDown.  decomp: the output of b200_decompose_ntt (beta*B polynomials); key as in b200_gadget_product.
extern "C" int b200_gadget_product_hoisted(b200_keyswitch *ks, int levelQ, int levelP, const b200_poly *decomp, const b200_poly *key,
                                           b200_poly *ct0, b200_poly *ct1) {
    CtxLock lk(ks ? ks->ctx : nullptr);
    if (int rc = check_levels(ks, levelQ, levelP, "b200_gadget_product_hoisted")) return rc;
    for (const b200_poly *p : {decomp, key, (const b200_poly *)ct0, (const b200_poly *)ct1})
        if (int rc = check_poly(ks, p, "b200_gadget_product_hoisted")) return rc;
    b200_ctx *ctx = ks->ctx;
    const int beta = (levelQ + levelP + 1) / (levelP + 1), B = ct0->batch, N = ctx->N, LQ = ks->LQ;
    if (key->batch < 2 * beta) { b200_set_error("b200_gadget_product_hoisted: the key must hold 2*beta polynomials [digit][component]"); return B200_ESHAPE; }
    if (decomp->batch < beta * B) { b200_set_error("b200_gadget_product_hoisted: decomp must hold beta*batch polynomials [digit][b]"); return B200_ESHAPE; }
    if (ct1->batch != B || ct0->d == ct1->d) { b200_set_error("b200_gadget_product_hoisted: outputs must be two distinct polynomials of the same batch size"); return B200_ESHAPE; }
    CUDA_TRY(cudaSetDevice(ctx->device));
    const size_t words = (size_t)B * ctx->L * N;
    if (int rc = ensure_ws(ks, 3 * words)) return rc;
    u64 *acc0 = ks->ws, *acc1 = ks->ws + words, *buf = ks->ws + 2 * words;
    std::vector<int> limbs = range(0, levelQ + 1);
    for (int j = 0; j <= levelP; ++j) limbs.push_back(LQ + j);
    int *&d_limbs = ks->limb_lists[limb_list_key(0, levelQ + 1, 0, levelP + 1)];
    if (!d_limbs) {
        cudaError_t ce = upload(&d_limbs, limbs);
        if (ce != cudaSuccess) return b200_cuda_fail(ce, "gadget product limb list");
    }
    int rc = 0;
    for (int d0 = 0; d0 < beta && rc == 0; d0 += MAC_DIGITS) {
        const int nd = std::min(MAC_DIGITS, beta - d0);
        MacGroup mg{};
        for (int g = 0; g < nd; ++g) {
            const int d = d0 + g;
            mg.op[g] = decomp->d + (size_t)d * words;
            mg.k0[g] = key->d + (size_t)(2 * d) * ctx->L * N;
            mg.k1[g] = key->d + (size_t)(2 * d + 1) * ctx->L * N;
            mg.st[g] = mg.ed[g] = 0;   // the own limbs are already in the decomposition
        }
        rc = launch_mac(ctx, decomp->d, mg, nd, acc0, acc1, d_limbs, (int)limbs.size(), B, d0 == 0);
    }
    if (rc) return rc;
    if ((rc = mod_down_ntt(ks, levelQ, levelP, acc0, ct0->d, buf, B))) return rc;
    return mod_down_ntt(ks, levelQ, levelP, acc1, ct1->d, buf, B);
}

// rlwe.Evaluator.GadgetProduct for a gadget ciphertext with a power-of-two base (BaseTwoDecomposition = pw2 > 0, one
// special prime): gadgetProductSinglePAndBitDecompLazy (evaluator_gadget_product.go:190-320) followed by ModDown (:33-45).
// sizes[i] = BaseTwoDecompositionVectorSize()[i] = len(gadgetCt.Value[i]) for i <= levelQ; the key holds
// 2 * sum(sizes) polynomials ordered [i][j][component].  pw2 == 0 is b200_gadget_product.
extern "C" int b200_gadget_product_bit_decomp(b200_keyswitch *ks, int levelQ, int levelP, int pw2, const int *sizes, const b200_poly *cx,
                                              const b200_poly *key, b200_poly *ct0, b200_poly *ct1) {
    if (pw2 == 0) return b200_gadget_product(ks, levelQ, levelP, cx, key, ct0, ct1);
    CtxLock lk(ks ? ks->ctx : nullptr);
    if (int rc = check_levels(ks, levelQ, levelP, "b200_gadget_product_bit_decomp")) return rc;
    for (const b200_poly *p : {cx, key, (const b200_poly *)ct0, (const b200_poly *)ct1})
        if (int rc = check_poly(ks, p, "b200_gadget_product_bit_decomp")) return rc;
    if (levelP != 0) { b200_set_error("b200_gadget_product_bit_decomp: a power-of-two base needs exactly one special prime (levelP == 0, rlwe/params.go:621)"); return B200_EINVAL; }
    if (pw2 < 0 || pw2 > 60 || !sizes) { b200_set_error("b200_gadget_product_bit_decomp: invalid base"); return B200_EINVAL; }
    long long total = 0;
    for (int i = 0; i <= levelQ; ++i) {
        if (sizes[i] < 1 || (long long)(sizes[i] - 1) * pw2 >= 64) { b200_set_error("b200_gadget_product_bit_decomp: invalid BaseTwoDecompositionVectorSize"); return B200_ESHAPE; }
        total += sizes[i];
    }
    b200_ctx *ctx = ks->ctx;
    const int B = cx->batch, N = ctx->N;
    if (key->batch < 2 * total) { b200_set_error("b200_gadget_product_bit_decomp: the key must hold 2*sum(sizes) polynomials [limb][window][component]"); return B200_ESHAPE; }
    if (ct0->batch != B || ct1->batch != B || ct0->d == ct1->d) { b200_set_error("b200_gadget_product_bit_decomp: outputs must be two distinct polynomials of the input's batch size"); return B200_ESHAPE; }
    CUDA_TRY(cudaSetDevice(ctx->device));
    const size_t words = (size_t)B * ctx->L * N;
    const int slots = (int)std::min<long long>(total, MAC_DIGITS);
    if (int rc = ensure_ws(ks, (3 + slots) * words)) return rc;
    u64 *cxinv = ks->ws, *acc0 = ks->ws + words, *acc1 = ks->ws + 2 * words, *c2 = ks->ws + 3 * words;
    int rc = ntt_range(ctx, cx->d, cxinv, B, 0, levelQ + 1, true);                               // ringQ.INTT(cx, cxInvNTT) (:201-203)
    if (rc) return rc;
    if ((rc = bitdecomp_mac(ks, levelQ, levelP, pw2, sizes, cxinv, key->d, c2, acc0, acc1, B))) return rc;
    if ((rc = mod_down_ntt(ks, levelQ, levelP, acc0, ct0->d, c2, B))) return rc;
    return mod_down_ntt(ks, levelQ, levelP, acc1, ct1->d, c2, B);
}

// ---- limb-sharded execution (SURVEY.md 8e / BASELINE config 5): the same gadget product in three device steps,
// separated by the two exchanges of the path (an all-gather of the coefficient-domain input limbs, and one of the
// P limbs of the accumulators), which the caller performs with NCCL on the polynomials' device memory.
extern "C" int b200_ks_decompose_mac(b200_keyswitch *ks, int levelQ, int levelP, const b200_poly *cx, const b200_poly *cxinv,
                                     const b200_poly *key, b200_poly *acc0, b200_poly *acc1, int q0, int q1, int p0, int p1) {
    CtxLock lk(ks ? ks->ctx : nullptr);
    if (int rc = check_levels(ks, levelQ, levelP, "b200_ks_decompose_mac")) return rc;
    for (const b200_poly *p : {cx, cxinv, key, (const b200_poly *)acc0, (const b200_poly *)acc1})
        if (int rc = check_poly(ks, p, "b200_ks_decompose_mac")) return rc;
    const Own o{q0, q1, p0, p1};
    if (int rc = check_own(ks, levelQ, levelP, o, "b200_ks_decompose_mac")) return rc;
    b200_ctx *ctx = ks->ctx;
    const int beta = (levelQ + levelP + 1) / (levelP + 1), B = cx->batch;
    if (key->batch < 2 * beta || cxinv->batch != B || acc0->batch != B || acc1->batch != B || acc0->d == acc1->d) {
        b200_set_error("b200_ks_decompose_mac: batch sizes do not match");
        return B200_ESHAPE;
    }
    CUDA_TRY(cudaSetDevice(ctx->device));
    const size_t words = (size_t)B * ctx->L * ctx->N;
    if (int rc = ensure_ws(ks, (size_t)std::min(beta, MAC_DIGITS) * words)) return rc;
    return decompose_mac(ks, levelQ, levelP, cx->d, cxinv->d, key->d, ks->ws, acc0->d, acc1->d, B, o);
}

extern "C" int b200_ks_mod_down_limbs(b200_keyswitch *ks, int levelQ, int levelP, const b200_poly *acc, b200_poly *out, int q0, int q1) {
    CtxLock lk(ks ? ks->ctx : nullptr);
    if (int rc = check_levels(ks, levelQ, levelP, "b200_ks_mod_down_limbs")) return rc;
    if (int rc = check_poly(ks, acc, "b200_ks_mod_down_limbs")) return rc;
    if (int rc = check_poly(ks, out, "b200_ks_mod_down_limbs")) return rc;
    const Own o{q0, q1, 0, 0};
    if (int rc = check_own(ks, levelQ, levelP, o, "b200_ks_mod_down_limbs")) return rc;
    if (acc->batch != out->batch) { b200_set_error("b200_ks_mod_down_limbs: batch sizes differ"); return B200_ESHAPE; }
    b200_ctx *ctx = ks->ctx;
    CUDA_TRY(cudaSetDevice(ctx->device));
    const size_t words = (size_t)acc->batch * ctx->L * ctx->N;
    if (int rc = ensure_ws(ks, words)) return rc;
    return mod_down_finish(ks, levelQ, levelP, acc->d, out->d, ks->ws, acc->batch, o);
}

// Ciphertext x ciphertext multiply with relinearisation in the NTT domain (BASELINE config 5):
// tensor product (schemes/ckks/evaluator.go:821-834, MForm first; schemes/matrix_ckks/evaluator.go:163-173 is the
// same without MForm) followed by rlwe.Evaluator.Relinearize = GadgetProduct(c2, rlk) added to (c0, c1)
// (schemes/ckks/evaluator.go:836-853, core/rlwe/evaluator_evaluationkey.go:125-153).  The reference never runs
// this composition on a 3N ring: it is defined here as exactly these reference steps in sequence.
extern "C" int b200_mul_relin(b200_keyswitch *ks, int levelQ, int levelP, const b200_poly *a0, const b200_poly *a1,
                              const b200_poly *b0, const b200_poly *b1, const b200_poly *rlk, b200_poly *out0, b200_poly *out1,
                              int mform_first) {
    CtxLock lk(ks ? ks->ctx : nullptr);
    if (int rc = check_levels(ks, levelQ, levelP, "b200_mul_relin")) return rc;
    for (const b200_poly *p : {a0, a1, b0, b1, rlk, (const b200_poly *)out0, (const b200_poly *)out1})
        if (int rc = check_poly(ks, p, "b200_mul_relin")) return rc;
    b200_ctx *ctx = ks->ctx;
    const int B = a0->batch, N = ctx->N, beta = (levelQ + levelP + 1) / (levelP + 1);
    if (a1->batch != B || b0->batch != B || b1->batch != B || out0->batch != B || out1->batch != B || rlk->batch < 2 * beta || out0->d == out1->d) {
        b200_set_error("b200_mul_relin: batch sizes do not match");
        return B200_ESHAPE;
    }
    CUDA_TRY(cudaSetDevice(ctx->device));
    const size_t words = (size_t)B * ctx->L * N;
    if (int rc = ensure_ws(ks, (size_t)(7 + std::min(beta, MAC_DIGITS)) * words)) return rc;
    u64 *cxinv = ks->ws, *acc0 = ks->ws + words, *acc1 = ks->ws + 2 * words;
    u64 *t0 = ks->ws + 3 * words, *t1 = ks->ws + 4 * words, *t2 = ks->ws + 5 * words, *g = ks->ws + 6 * words, *c2buf = ks->ws + 7 * words;
    const Span spq{B, levelQ + 1, ctx->L, 0};
    int rc = launch_tensor(ctx, a0->d, a1->d, b0->d, b1->d, t0, t1, t2, spq, mform_first != 0);
    if (rc) return rc;
    if ((rc = ntt_range(ctx, t2, cxinv, B, 0, levelQ + 1, true))) return rc;
    if ((rc = decompose_mac(ks, levelQ, levelP, t2, cxinv, rlk->d, c2buf, acc0, acc1, B, Own{0, levelQ + 1, 0, levelP + 1}))) return rc;
    if ((rc = mod_down_ntt(ks, levelQ, levelP, acc0, g, c2buf, B))) return rc;
    if ((rc = launch_vec_op(ctx, B200_OP_ADD, t0, g, out0->d, spq, nullptr))) return rc;          // ringQ.Add(c0, tmpCt.Value[0], opOut.Value[0])
    if ((rc = mod_down_ntt(ks, levelQ, levelP, acc1, g, c2buf, B))) return rc;
    return launch_vec_op(ctx, B200_OP_ADD, t1, g, out1->d, spq, nullptr);
}

// rlwe.Evaluator.Automorphism (core/rlwe/evaluator_automorphism.go:13-60) on an NTT-domain ciphertext of degree 1 over a
// power-of-two ring: (t0, t1) = GadgetProduct(ct1, galois key); t0 += ct0; out_c = AutomorphismNTT(t_c, galEl).
// With decomp != NULL the key switch is the hoisted one (AutomorphismHoisted, :62-108) on the caller's DecomposeNTT(ct1).
extern "C" int b200_automorphism_ct(b200_keyswitch *ks, int levelQ, int levelP, const b200_poly *ct0, const b200_poly *ct1,
                                    const b200_poly *decomp, uint64_t galEl, const b200_poly *key, b200_poly *out0, b200_poly *out1) {
    CtxLock lk(ks ? ks->ctx : nullptr);
    if (int rc = check_levels(ks, levelQ, levelP, "b200_automorphism_ct")) return rc;
    for (const b200_poly *p : {ct0, ct1, key, (const b200_poly *)out0, (const b200_poly *)out1})
        if (int rc = check_poly(ks, p, "b200_automorphism_ct")) return rc;
    if (decomp) if (int rc = check_poly(ks, decomp, "b200_automorphism_ct")) return rc;
    b200_ctx *ctx = ks->ctx;
    if (ctx->ring_type != B200_RING_STANDARD) { b200_set_error("N must be a power of two"); return B200_ESHAPE; }   // AutomorphismNTTIndex (ring/automorphism.go:14-16)
    const int B = ct0->batch, N = ctx->N;
    if (ct1->batch != B || out0->batch != B || out1->batch != B || out0->d == out1->d) { b200_set_error("b200_automorphism_ct: batch sizes do not match"); return B200_ESHAPE; }
    CUDA_TRY(cudaSetDevice(ctx->device));
    const size_t words = (size_t)B * ctx->L * N;
    // two result polynomials of our own next to the key switch's workspace (which it may re-allocate: grow it first)
    const int beta = (levelQ + levelP + 1) / (levelP + 1);
    if (int rc = ensure_ws(ks, (size_t)(3 + std::min(beta, MAC_DIGITS)) * words)) return rc;
    if (int rc = b200_ensure_scratch2(ctx, 2 * words)) return rc;
    u64 *t0 = ctx->d_ws, *t1 = ctx->d_ws + words;
    b200_poly p0{ctx, t0, B, false}, p1{ctx, t1, B, false};
    int rc = decomp ? b200_gadget_product_hoisted(ks, levelQ, levelP, decomp, key, &p0, &p1)
                    : b200_gadget_product(ks, levelQ, levelP, ct1, key, &p0, &p1);
    if (rc) return rc;
    const Span spq{B, levelQ + 1, ctx->L, 0};
    if ((rc = launch_vec_op(ctx, B200_OP_ADD, t0, ct0->d, t0, spq, nullptr))) return rc;          // ringQ.Add(ctTmp.Value[0], ctIn.Value[0], ctTmp.Value[0]) (:43)
    const u64 nth = 2 * (u64)N;
    if ((rc = launch_automorphism_ntt(ctx, t0, out0->d, nullptr, galEl, nth, N, ctx->L, B, levelQ + 1, 0))) return rc;
    return launch_automorphism_ntt(ctx, t1, out1->d, nullptr, galEl, nth, N, ctx->L, B, levelQ + 1, 0);
}
