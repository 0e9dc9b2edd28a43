// b200ring.cu -- the C ABI of include/b200ring.h: context creation (per-prime constants and twiddle
// tables, bit-identical to the reference's ring/subring.go:129-214), device polynomial management and
// dispatch to the sm_100a kernels.  No CPU fallback: every compute entry point needs a CUDA device.
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <set>

#include "hostmath.hpp"
#include "internal.hpp"

namespace hm = hostmath;

// ---- error plumbing -------------------------------------------------------------------------------
static thread_local std::string g_err;
void b200_set_error(const std::string &msg) { g_err = msg; }
int b200_cuda_fail(cudaError_t e, const char *what) {
    g_err = std::string("CUDA error: ") + cudaGetErrorString(e) + " in " + what;
    if (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver) return B200_ENODEVICE;
    if (e == cudaErrorMemoryAllocation) return B200_ENOMEM;
    return B200_ECUDA;
}
static int fail(int code, const std::string &msg) {
    g_err = msg;
    return code;
}

extern "C" const char *b200_last_error(void) { return g_err.c_str(); }
extern "C" const char *b200_version(void) { return "b200ring 0.1 (sm_100a)"; }
extern "C" int b200_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

static thread_local cudaStream_t tl_stream = nullptr;
static thread_local bool tl_stream_on = false;
cudaStream_t b200_cur_stream(b200_ctx *ctx) { return tl_stream_on ? tl_stream : ctx->stream; }
void b200_set_stream_override(cudaStream_t s, bool active) { tl_stream = s; tl_stream_on = active; }

int b200_ensure_scratch(b200_ctx *ctx, size_t words) {
    if (ctx->scratch_words >= words) return 0;
    CUDA_TRY(cudaStreamSynchronize(b200_cur_stream(ctx)));
    if (ctx->d_scratch) CUDA_TRY(cudaFree(ctx->d_scratch));
    ctx->d_scratch = nullptr;
    ctx->scratch_words = 0;
    CUDA_TRY(cudaMalloc(&ctx->d_scratch, words * sizeof(u64)));
    ctx->scratch_words = words;
    return 0;
}

// second workspace (ctx->d_ws) for composite operators whose steps use d_scratch themselves
int b200_ensure_scratch2(b200_ctx *ctx, size_t words) {
    if (ctx->ws_words >= words) return 0;
    CUDA_TRY(cudaStreamSynchronize(b200_cur_stream(ctx)));
    if (ctx->d_ws) CUDA_TRY(cudaFree(ctx->d_ws));
    ctx->d_ws = nullptr; ctx->ws_words = 0;
    CUDA_TRY(cudaMalloc(&ctx->d_ws, words * sizeof(u64)));
    ctx->ws_words = words;
    return 0;
}

// ---- ring-degree validation (subring.go:14-30, ring.go:266,319-323) -------------------------------------
static bool valid_3n_degree(int N, int *a_out, int *b_out) {
    if (N <= 0) return false;
    int a = 0, b = 0, t = N;
    while (t % 2 == 0) t /= 2, ++a;
    while (t % 3 == 0) t /= 3, ++b;
    if (a_out) *a_out = a;
    if (b_out) *b_out = b;
    return t == 1;
}

static tw_t mk_tw(u64 w, u64 q) { return make_ulonglong2(w, hm::shoup(w, q)); }

// STANDARD tables: Go-format (Montgomery, bit-reversed; subring.go:183-205) + plain/Shoup device format
static void build_standard(b200_ctx *ctx, int i, std::vector<tw_t> &fwd, std::vector<tw_t> &inv, tw_t *misc) {
    SubRingHost &s = ctx->sub[i];
    const int N = ctx->N, logN = ctx->logN;
    const u64 q = s.q, nth = 2 * (u64)N;
    s.ninv = hm::mform(hm::invmod_prime(nth >> 1, q), q);                       // subring.go:183
    const u64 psi = hm::powmod(s.g, (q - 1) / nth, q);                           // :187
    const u64 psi_inv = hm::powmod(s.g, q - ((q - 1) / nth) - 1, q);             // :188
    s.root = psi;
    const u64 psi_m = hm::mform(psi, q), psi_inv_m = hm::mform(psi_inv, q);
    s.roots_fwd.assign(N, 0);
    s.roots_bwd.assign(N, 0);
    s.roots_fwd[0] = s.roots_bwd[0] = hm::mform(1, q);
    u64 pf = 1, pb = 1;   // plain-domain running powers
    fwd[0] = inv[0] = mk_tw(1, q);
    for (u64 j = 1; j < (u64)N; ++j) {                                           // :198-205
        const u64 prev = hm::bitrev(j - 1, logN), cur = hm::bitrev(j, logN);
        s.roots_fwd[cur] = hm::mred(s.roots_fwd[prev], psi_m, q, s.qinv);
        s.roots_bwd[cur] = hm::mred(s.roots_bwd[prev], psi_inv_m, q, s.qinv);
        pf = hm::mulmod(pf, psi, q);
        pb = hm::mulmod(pb, psi_inv, q);
        fwd[cur] = mk_tw(pf, q);
        inv[cur] = mk_tw(pb, q);
    }
    const u64 ninv = hm::invmod_prime((u64)N % q, q);
    misc[MISC_NINV] = mk_tw(ninv, q);
    misc[MISC_NINV_W1] = mk_tw(hm::mulmod(ninv, inv[1].x, q), q);
    const u64 r64 = hm::mform(1, q);   // 2^64 mod q
    misc[MISC_NINVM] = mk_tw(hm::mulmod(ninv, r64, q), q);
    misc[MISC_NINV_W1M] = mk_tw(hm::mulmod(hm::mulmod(ninv, inv[1].x, q), r64, q), q);
}

// CONJUGATE_INVARIANT tables (NewRingConjugateInvariant, ring.go:280-282: NthRoot = 4N).  The Go-format tables hold 2N
// entries (subring.go:183-205).  The transform (ntt.go:754-786) is a folding layer with RootsForward[1] followed by the
// layers m = 2, 4, .., N that read RootsForward[m + i], i < m/2: the left half of every level of the 2N-point tree.
// Relabelled level m' = m/2 that is exactly the heap order of an N-point transform, table[m' + i] = Roots4N[2m' + i],
// so the STANDARD kernels run it unchanged; the inverse (ntt.go:1104-1158) mirrors it with NInv = (2N)^-1.
static void build_ci(b200_ctx *ctx, int i, std::vector<tw_t> &fwd, std::vector<tw_t> &inv, tw_t *misc) {
    SubRingHost &s = ctx->sub[i];
    const int N = ctx->N, logH = ctx->logN + 1;
    const u64 q = s.q, nth = 4 * (u64)N, half = nth >> 1;
    s.ninv = hm::mform(hm::invmod_prime(half % q, q), q);                        // subring.go:183
    const u64 psi = hm::powmod(s.g, (q - 1) / nth, q);                           // :187
    const u64 psi_inv = hm::powmod(s.g, q - ((q - 1) / nth) - 1, q);             // :188
    s.root = psi;
    const u64 psi_m = hm::mform(psi, q), psi_inv_m = hm::mform(psi_inv, q);
    s.roots_fwd.assign(half, 0);
    s.roots_bwd.assign(half, 0);
    s.roots_fwd[0] = s.roots_bwd[0] = hm::mform(1, q);
    std::vector<u64> pf(half), pb(half);   // plain-domain copies of the two tables
    pf[0] = pb[0] = 1;
    u64 f = 1, b = 1;
    for (u64 j = 1; j < half; ++j) {                                             // :198-205
        const u64 prev = hm::bitrev(j - 1, logH), cur = hm::bitrev(j, logH);
        s.roots_fwd[cur] = hm::mred(s.roots_fwd[prev], psi_m, q, s.qinv);
        s.roots_bwd[cur] = hm::mred(s.roots_bwd[prev], psi_inv_m, q, s.qinv);
        f = hm::mulmod(f, psi, q);
        b = hm::mulmod(b, psi_inv, q);
        pf[cur] = f; pb[cur] = b;
    }
    fwd[0] = inv[0] = mk_tw(1, q);
    for (u64 m = 1; m < (u64)N; m <<= 1)
        for (u64 k = 0; k < m; ++k) {
            fwd[m + k] = mk_tw(pf[2 * m + k], q);
            inv[m + k] = mk_tw(pb[2 * m + k], q);
        }
    const u64 ninv = hm::invmod_prime(half % q, q);
    misc[MISC_NINV] = mk_tw(ninv, q);
    misc[MISC_NINV_W1] = mk_tw(hm::mulmod(ninv, inv[1].x, q), q);
    misc[MISC_CI_F] = mk_tw(pf[1], q);
    misc[MISC_CI_FINV] = mk_tw(pb[1], q);
}

// MATRIX tables: zetas in the order of the exponent tree: index 1 = omega^(N/2),
// radix-3 layer with nb blocks at [nb + 2 blk, +1], radix-2 layer with nb blocks at [nb + blk].
static void build_matrix(b200_ctx *ctx, int i, std::vector<tw_t> &fwd, std::vector<tw_t> &inv, tw_t *misc,
                         std::vector<int> *leaf_out) {
    SubRingHost &s = ctx->sub[i];
    const int N = ctx->N;
    const u64 q = s.q, omega = s.root, M = 3 * (u64)N;
    std::vector<u64> pw(M);   // omega^e, e in [0, 3N)
    pw[0] = 1;
    for (u64 e = 1; e < M; ++e) pw[e] = hm::mulmod(pw[e - 1], omega, q);
    auto set = [&](size_t idx, u64 e) {
        e %= M;
        fwd[idx] = mk_tw(pw[e], q);
        inv[idx] = mk_tw(pw[(M - e) % M], q);
    };
    std::vector<u64> cur{M / 6, 5 * M / 6}, nxt;
    set(0, 0);
    set(1, cur[0]);
    for (int l = 0; l < ctx->b; ++l) {              // radix-3 levels: e -> e/3 + {0, N, 2N}
        const size_t nb = cur.size();
        nxt.assign(3 * nb, 0);
        for (size_t k = 0; k < nb; ++k) {
            const u64 e = cur[k] / 3;
            nxt[3 * k] = e; nxt[3 * k + 1] = e + M / 3; nxt[3 * k + 2] = e + 2 * (M / 3);
            set(nb + 2 * k, e);
            set(nb + 2 * k + 1, 2 * e);
        }
        cur.swap(nxt);
    }
    for (int l = 0; l < ctx->a - 1; ++l) {          // radix-2 levels: e -> e/2 + {0, 3N/2}
        const size_t nb = cur.size();
        nxt.assign(2 * nb, 0);
        for (size_t k = 0; k < nb; ++k) {
            const u64 e = cur[k] / 2;
            nxt[2 * k] = e; nxt[2 * k + 1] = e + M / 2;
            set(nb + k, e);
        }
        cur.swap(nxt);
    }
    if (leaf_out) {
        leaf_out->resize(N);
        for (int k = 0; k < N; ++k) (*leaf_out)[k] = (int)(cur[k] % M);
    }
    const u64 zeta = pw[N / 2], zeta5 = pw[(5 * (u64)N / 2) % M];
    const u64 c = hm::invmod_prime(hm::submod(zeta, zeta5, q), q);
    const u64 two_over_n = hm::mulmod(2, hm::invmod_prime((u64)N % q, q), q);
    misc[MISC_OMEGA3] = mk_tw(pw[N], q);
    misc[MISC_K1] = mk_tw(hm::mulmod(c, two_over_n, q), q);
    misc[MISC_K2] = mk_tw(two_over_n, q);
    misc[MISC_ZETA] = mk_tw(zeta, q);
    misc[MISC_K4] = mk_tw(hm::invmod_prime((u64)N % q, q), q);
    const u64 r64 = hm::mform(1, q);   // 2^64 mod q
    misc[MISC_K1M] = mk_tw(hm::mulmod(hm::mulmod(c, two_over_n, q), r64, q), q);
    misc[MISC_K4M] = mk_tw(hm::mulmod(hm::invmod_prime((u64)N % q, q), r64, q), q);
}

// Row-pass order of a twiddle table (the row kernels of ntt_big.cu / ntt_fold.cuh): the 255 twiddles of the subtree under row
// node R + r (R = N/256) move to the block [R + 255 r, R + 255 (r + 1)) in the order the lanes load them -- 7 entries of the
// first register group (warp-uniform), 7 runs of 8 entries (second group, indexed by lane & 7), 6 runs of 32 entries (third
// group, indexed by lane) -- so that one load instruction of a warp reads consecutive entries instead of a strided set
// (L1 data-pipe wavefronts were the busiest resource of the row kernels).  Heap indices below R, all the column pass uses,
// stay where they are.
static int row_pass_index(int N, int h) {
    const int R = N / 256;
    if (h < R) return h;
    int d = 0;
    while ((h >> d) >= 2 * R) ++d;
    const int r = (h >> d) - R, k = h & ((1 << d) - 1);
    int off;
    if (d < 3) off = (1 << d) - 1 + k;
    else if (d < 6) { const int j = d - 3, hi3 = k >> j, g = k & ((1 << j) - 1); off = 7 + ((1 << j) - 1 + g) * 8 + hi3; }
    else { const int j = d - 6, sub = k >> j, g = k & ((1 << j) - 1), lane = sub >> 1, s2 = sub & 1; off = 63 + (s2 * 3 + (1 << j) - 1 + g) * 32 + lane; }
    return R + 255 * r + off;
}
extern "C" int b200_row_pass_index(int N, int heap_index) {
    if (N < 512 || N % 256 != 0 || heap_index < 0 || heap_index >= N) return -1;
    return row_pass_index(N, heap_index);
}
template <typename T>
static void row_pass_order(std::vector<T> &t, int L, int N) {
    const int R = N / 256;
    std::vector<T> o(t);
    for (int i = 0; i < L; ++i) {
        const T *src = &t[(size_t)i * N];
        T *dst = &o[(size_t)i * N];
        for (int h = R; h < N; ++h) dst[row_pass_index(N, h)] = src[h];
    }
    t.swap(o);
}

// ---- fold arithmetic tables (solinas.cuh) ---------------------------------------------------------------
// q = 2^k + c with the smallest |c|
static void fold_shape(u64 q, int *k, long long *c) {
    int b = 63;
    while (b > 0 && !((q >> b) & 1)) --b;
    const u64 lo = 1ULL << b;
    if (b < 63 && q - lo > (lo << 1) - q) { *k = b + 1; *c = -(long long)((lo << 1) - q); }
    else { *k = b; *c = (long long)(q - lo); }
}
static bool fold_split(long long v, int s, int *lo, int *hi) {
    const long long h = (v + (1LL << (s - 1))) >> s;
    if (h > 0x7fffffffLL || h < -0x80000000LL) return false;
    *hi = (int)h;
    *lo = (int)(v - h * (1LL << s));
    return true;
}
// {w, w*2^32 mod q} as balanced residues split at bit k-32; false when a half does not fit 32 bits (a residue within
// 2^(k-33) of +-q/2 for q > 2^k: about 2^-33 of all residues)
static bool fold_entry(u64 w, u64 q, int k, stw_t *out) {
    const u64 v = hm::mulmod(w, (1ULL << 32) % q, q);
    const long long wb = w > q / 2 ? (long long)w - (long long)q : (long long)w;
    const long long vb = v > q / 2 ? (long long)v - (long long)q : (long long)v;
    return fold_split(wb, k - 32, &out->w_lo, &out->w_hi) && fold_split(vb, k - 32, &out->v_lo, &out->v_hi);
}
// The windows of the fold arithmetic are stated in units of 2^k (tools/fold_ranges.py): the term c*B.hi of a product is at
// most |c| * 2^30 / 2^k of a unit, so |c| <= 2^(k-33) keeps it at the 1/8 the interval model assumes for every k (2^27 at
// k = 60; a bound of 2^27 for all k would let a 55-bit prime with a 27-bit c push products four units off).
static bool fold_prime_ok(u64 q) {
    int k; long long c;
    fold_shape(q, &k, &c);
    return k >= 50 && k <= 60 && (c < 0 ? -c : c) <= (1LL << (k - 33));
}
// builds the device tables; leaves ctx->fold_ok false (and no tables) when the ring does not qualify
static cudaError_t fold_setup(b200_ctx *ctx, const std::vector<tw_t> &fwd, const std::vector<tw_t> &inv, const std::vector<tw_t> &misc) {
    const char *off = std::getenv("B200_NO_FOLD");
    if (off && std::atoi(off)) return cudaSuccess;
    const int L = ctx->L;
    std::vector<SolConst> sol(L);
    for (int i = 0; i < L; ++i) {
        const u64 q = ctx->sub[i].q;
        if (!fold_prime_ok(q)) return cudaSuccess;
        int k; long long c;
        fold_shape(q, &k, &c);
        sol[i].q = q;
        sol[i].bias = ((1ULL << 63) / q) * q;
        sol[i].negc = (int)-c;
        sol[i].two_s = 1u << (k - 32);
        sol[i].sh = (unsigned)(k - 32);
        sol[i].mask = (1u << (k - 32)) - 1;
    }
    auto conv = [&](const std::vector<tw_t> &src, size_t per_limb, std::vector<stw_t> &dst) {
        dst.resize(src.size());
        for (int i = 0; i < L; ++i) {
            const u64 q = ctx->sub[i].q;
            const int k = (int)sol[i].sh + 32;
            for (size_t j = 0; j < per_limb; ++j)
                if (!fold_entry(src[(size_t)i * per_limb + j].x, q, k, &dst[(size_t)i * per_limb + j])) return false;
        }
        return true;
    };
    std::vector<stw_t> f, v, m;
    if (!conv(fwd, (size_t)ctx->N, f) || !conv(inv, (size_t)ctx->N, v) || !conv(misc, (size_t)MISC_COUNT, m)) return cudaSuccess;
    if (ctx->N >= 512 && ctx->N % 256 == 0) {
        row_pass_order(f, L, ctx->N);
        row_pass_order(v, L, ctx->N);
    }
    auto up = [&](void **dst, const void *src, size_t bytes) -> cudaError_t {
        cudaError_t e = cudaMalloc(dst, bytes);
        if (e != cudaSuccess) return e;
        return cudaMemcpy(*dst, src, bytes, cudaMemcpyHostToDevice);
    };
    cudaError_t e = up((void **)&ctx->d_sol, sol.data(), sizeof(SolConst) * L);
    if (e == cudaSuccess) e = up((void **)&ctx->d_fwd_s, f.data(), sizeof(stw_t) * f.size());
    if (e == cudaSuccess) e = up((void **)&ctx->d_inv_s, v.data(), sizeof(stw_t) * v.size());
    if (e == cudaSuccess) e = up((void **)&ctx->d_misc_s, m.data(), sizeof(stw_t) * m.size());
    if (e == cudaSuccess) ctx->fold_ok = true;
    return e;
}

extern "C" int b200_ctx_create(int ring_type, int N, int L, const uint64_t *moduli, const uint64_t *roots,
                               int device, b200_ctx **out) {
    if (!out) return fail(B200_EINVAL, "b200_ctx_create: out is NULL");
    *out = nullptr;
    int a = 0, b = 0;
    const bool pow2 = N >= 8 && (N & (N - 1)) == 0;
    const bool is3n = valid_3n_degree(N, &a, &b);
    if (ring_type != B200_RING_STANDARD && ring_type != B200_RING_MATRIX && ring_type != B200_RING_CONJUGATE_INVARIANT)
        return fail(B200_EINVAL, "invalid ring type");
    if (!pow2 && !is3n)   // ring.go:319-323
        return fail(B200_EINVAL, "invalid ring degree: must be a power of 2 greater than 8 or satisfy 3N = 2^a * 3^{b+1} condition");
    if (ring_type == B200_RING_STANDARD && !pow2)
        return fail(B200_EINVAL, "invalid ring degree: the Standard ring needs N to be a power of 2 greater than 8");
    if (ring_type == B200_RING_CONJUGATE_INVARIANT && !pow2)   // the reference's 6N variant runs radix-2 butterflies on a non-power-of-two length
        return fail(B200_EINVAL, "invalid ring degree: the ConjugateInvariant ring needs N to be a power of 2 greater than 8");
    if (ring_type == B200_RING_MATRIX && !(N % 3 == 0 && is3n && a >= 1))   // ring.go:301-303, 319-323: every N = 2^a 3^b with a, b >= 1, N = 6 included
        return fail(B200_EINVAL, "matrix ring type requires N to satisfy 3N = 2^a * 3^{b+1} condition, got N=" + std::to_string(N));
    if (L <= 0 || !moduli) return fail(B200_EINVAL, "invalid ModuliChain (must be a non-empty []uint64)");   // ring.go:325-327
    {
        std::set<u64> seen(moduli, moduli + L);
        if ((int)seen.size() != L) return fail(B200_EINVAL, "invalid ModuliChain (moduli are not distinct)");   // ring.go:329-331
    }
    const u64 nthroot = ring_type == B200_RING_MATRIX ? 3 * (u64)N : ring_type == B200_RING_CONJUGATE_INVARIANT ? 4 * (u64)N : 2 * (u64)N;
    for (int i = 0; i < L; ++i) {
        const u64 q = moduli[i];
        if (!hm::is_prime(q)) return fail(B200_EINVAL, "invalid modulus: " + std::to_string(q) + " is not prime)");      // subring.go:140-142
        if (q % nthroot != 1) return fail(B200_EINVAL, "invalid modulus: " + std::to_string(q) + " != 1 mod NthRoot)");   // subring.go:144-146
        if (q >> 61) return fail(B200_EINVAL, "invalid modulus: " + std::to_string(q) + " exceeds 61 bits (lazy butterflies need 8q < 2^64, ring/ntt.go:169)");
    }
    int ndev = b200_device_count();
    if (ndev <= 0) return fail(B200_ENODEVICE, "no CUDA device visible: b200ring has no CPU fallback");
    if (device < 0 || device >= ndev) return fail(B200_EINVAL, "invalid device index " + std::to_string(device));
    CUDA_TRY(cudaSetDevice(device));

    b200_ctx *ctx = new b200_ctx();
    ctx->ring_type = ring_type; ctx->N = N; ctx->L = L; ctx->device = device;
    ctx->core_type = ring_type == B200_RING_MATRIX ? B200_RING_MATRIX : B200_RING_STANDARD;
    ctx->a = a; ctx->b = b;
    while ((1 << ctx->logN) < N) ctx->logN++;
    cudaDeviceGetAttribute(&ctx->max_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device);
    cudaDeviceGetAttribute(&ctx->sm_count, cudaDevAttrMultiProcessorCount, device);
    ctx->sub.resize(L);

    std::vector<LimbConst> consts(L);
    std::vector<tw_t> fwd((size_t)L * N), inv((size_t)L * N), misc((size_t)L * MISC_COUNT, make_ulonglong2(0, 0));
    std::vector<int> leaf;
    for (int i = 0; i < L; ++i) {
        SubRingHost &s = ctx->sub[i];
        s.q = moduli[i];
        s.qinv = hm::mred_constant(s.q);
        hm::bred_constant(s.q, s.brc);
        int bl = 0;
        while (bl < 64 && ((s.q - 1) >> bl)) ++bl;
        s.mask = bl >= 64 ? ~0ULL : ((1ULL << bl) - 1);   // subring.go:94
        {
            // quotient estimate constants (canon_est, modarith.cuh): xs = x >> sh has <= 32 bits for x < 16q; qs = q >> sh has
            // 28 bits.  mul = floor(2^32 / (qs + 1)) lies in [16, 32], so its truncation alone is worth up to 1/17 of the
            // quotient: k = (xs * mul) >> 32 never exceeds floor(x / q) and falls short of it by at most one ONLY because
            // x < 16q (deficit < 16/17 + 2^-27 * 16 < 1, plus the floor).  Do not reuse these constants for a wider input range;
            // tests/test_canon_est.py restates the bound on random and extreme q.
            int bl = 0;
            while (bl < 64 && (s.q >> bl)) ++bl;
            const unsigned sh = bl > 28 ? (unsigned)(bl - 28) : 0u;
            const u64 qs = s.q >> sh;
            const unsigned mul = (unsigned)((1ULL << 32) / (qs + 1));
            consts[i] = LimbConst{s.q, s.qinv, s.brc[0], s.brc[1], sh, mul, 0};
        }
        std::vector<tw_t> f(N), v(N);
        const u64 given = roots ? roots[i] : 0;
        if (ring_type != B200_RING_MATRIX) {
            if (given) {
                std::vector<u64> fac = hm::unique_factors(s.q - 1);
                if (!hm::is_primitive_root(given % s.q, s.q, fac)) { delete ctx; return fail(B200_EINVAL, "roots[" + std::to_string(i) + "] is not a primitive root"); }
                s.g = given;
            } else {
                s.g = hm::primitive_root(s.q, hm::unique_factors(s.q - 1));   // subring.go:218-251
            }
            if (ring_type == B200_RING_STANDARD) build_standard(ctx, i, f, v, &misc[(size_t)i * MISC_COUNT]);
            else build_ci(ctx, i, f, v, &misc[(size_t)i * MISC_COUNT]);
        } else {
            if (given) {
                const u64 w = given % s.q;
                if (hm::powmod(w, nthroot, s.q) != 1 || hm::powmod(w, nthroot / 2, s.q) == 1 || hm::powmod(w, nthroot / 3, s.q) == 1) {
                    delete ctx;
                    return fail(B200_EINVAL, "roots[" + std::to_string(i) + "] is not a primitive 3N-th root of unity");
                }
                s.root = w;
            } else {
                s.g = hm::primitive_root(s.q, hm::unique_factors(s.q - 1));
                s.root = hm::powmod(s.g, (s.q - 1) / nthroot, s.q);           // subring.go:255-290
            }
            build_matrix(ctx, i, f, v, &misc[(size_t)i * MISC_COUNT], i == 0 ? &leaf : nullptr);
        }
        std::copy(f.begin(), f.end(), fwd.begin() + (size_t)i * N);
        std::copy(v.begin(), v.end(), inv.begin() + (size_t)i * N);
    }

    // RescaleConstants (ring/ring.go:361-379): rc[j-1][i] = MForm(q_i - q_j^-1 mod q_i) for i < j
    ctx->rescale.assign((size_t)(L > 1 ? L - 1 : 0) * L, 0);
    for (int j = 1; j < L; ++j)
        for (int i = 0; i < j; ++i) {
            const u64 qi = ctx->sub[i].q, qj = ctx->sub[j].q;
            ctx->rescale[(size_t)(j - 1) * L + i] = hm::mform(qi - hm::invmod_prime(qj % qi, qi), qi);
        }

    auto up = [&](void **dst, const void *src, size_t bytes) -> cudaError_t {
        cudaError_t e = cudaMalloc(dst, bytes);
        if (e != cudaSuccess) return e;
        return cudaMemcpy(*dst, src, bytes, cudaMemcpyHostToDevice);
    };
    cudaError_t e = up((void **)&ctx->d_consts, consts.data(), sizeof(LimbConst) * L);
    if (e == cudaSuccess) e = up((void **)&ctx->d_fwd, fwd.data(), sizeof(tw_t) * fwd.size());
    if (e == cudaSuccess) e = up((void **)&ctx->d_inv, inv.data(), sizeof(tw_t) * inv.size());
    if (e == cudaSuccess) e = up((void **)&ctx->d_misc, misc.data(), sizeof(tw_t) * misc.size());
    if (e == cudaSuccess && !ctx->rescale.empty()) e = up((void **)&ctx->d_rescale, ctx->rescale.data(), sizeof(u64) * ctx->rescale.size());
    if (e == cudaSuccess && ring_type == B200_RING_MATRIX) {
        // reference slot order: slot k <-> omega^E[k], E ascending totatives of 3N (ntt_3n.go:46-53,235-243)
        std::vector<int> pos(3 * (size_t)N, -1), go2fast(N);
        int k = 0;
        for (int ex = 1; ex < 3 * N; ++ex)
            if (ex % 2 != 0 && ex % 3 != 0) pos[ex] = k++;
        for (int idx = 0; idx < N; ++idx) go2fast[pos[leaf[idx]]] = idx;
        e = up((void **)&ctx->d_go2fast, go2fast.data(), sizeof(int) * N);
    }
    if (e == cudaSuccess) e = fold_setup(ctx, fwd, inv, misc);
    if (e == cudaSuccess && N >= 512 && N % 256 == 0) {   // shapes the two-pass kernels may cover: row-pass copies of the Shoup tables
        row_pass_order(fwd, L, N);
        row_pass_order(inv, L, N);
        e = up((void **)&ctx->d_fwd_r, fwd.data(), sizeof(tw_t) * fwd.size());
        if (e == cudaSuccess) e = up((void **)&ctx->d_inv_r, inv.data(), sizeof(tw_t) * inv.size());
    }
    if (e != cudaSuccess) { b200_ctx_destroy(ctx); return b200_cuda_fail(e, "context table upload"); }
    int rc = big_tables_create(ctx);
    if (rc < 0) { b200_ctx_destroy(ctx); return rc; }
    *out = ctx;
    return B200_OK;
}

extern "C" void b200_ctx_destroy(b200_ctx *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    big_tables_destroy(ctx);
    cudaFree(ctx->d_consts); cudaFree(ctx->d_fwd); cudaFree(ctx->d_inv); cudaFree(ctx->d_misc); cudaFree(ctx->d_fwd_r); cudaFree(ctx->d_inv_r);
    cudaFree(ctx->d_sol); cudaFree(ctx->d_fwd_s); cudaFree(ctx->d_inv_s); cudaFree(ctx->d_misc_s);
    cudaFree(ctx->d_go2fast); cudaFree(ctx->d_scratch); cudaFree(ctx->d_ws); cudaFree(ctx->d_rescale); cudaFree(ctx->d_index);
    if (ctx->pipe.ready) {
        for (int i = 0; i < b200_ctx::HostPipe::NB; ++i) {
            cudaFree(ctx->pipe.buf[i]);
            cudaEventDestroy(ctx->pipe.up[i]); cudaEventDestroy(ctx->pipe.done[i]); cudaEventDestroy(ctx->pipe.freed[i]);
        }
        cudaEventDestroy(ctx->pipe.entry);
        cudaStreamDestroy(ctx->pipe.h2d); cudaStreamDestroy(ctx->pipe.comp); cudaStreamDestroy(ctx->pipe.d2h);
    }
    delete ctx;
}
extern "C" int b200_ctx_set_stream(b200_ctx *ctx, void *s) {
    CtxLock lk(ctx);
    if (!ctx) return fail(B200_EINVAL, "ctx is NULL");
    ctx->stream = (cudaStream_t)s;
    return 0;
}
extern "C" int b200_sync(b200_ctx *ctx) {
    CtxLock lk(ctx);
    if (!ctx) return fail(B200_EINVAL, "ctx is NULL");
    CUDA_TRY(cudaSetDevice(ctx->device));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return 0;
}
extern "C" int b200_ctx_n(const b200_ctx *ctx) { return ctx ? ctx->N : B200_EINVAL; }
extern "C" int b200_ctx_moduli_count(const b200_ctx *ctx) { return ctx ? ctx->L : B200_EINVAL; }
extern "C" int b200_ctx_ring_type(const b200_ctx *ctx) { return ctx ? ctx->ring_type : B200_EINVAL; }
extern "C" uint64_t b200_ctx_launch_count(const b200_ctx *ctx) { return ctx ? ctx->launches.load() : 0; }

// ---- per-kernel device timing ------------------------------------------------------------------------
extern "C" int b200_profile_begin(b200_ctx *ctx) {
    CtxLock lk(ctx);
    if (!ctx) return fail(B200_EINVAL, "ctx is NULL");
    CUDA_TRY(cudaSetDevice(ctx->device));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    for (auto &r : ctx->prof) { cudaEventDestroy(r.e0); cudaEventDestroy(r.e1); }
    ctx->prof.clear();
    ctx->profiling = true;
    return 0;
}
extern "C" int b200_profile_end(b200_ctx *ctx, char *out, size_t cap) {
    CtxLock lk(ctx);
    if (!ctx || !out || cap == 0) return fail(B200_EINVAL, "b200_profile_end: bad argument");
    CUDA_TRY(cudaSetDevice(ctx->device));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    ctx->profiling = false;
    struct Acc { double ms = 0; long long launches = 0, units = 0; };
    std::vector<std::pair<std::string, Acc>> acc;
    for (auto &r : ctx->prof) {
        float ms = 0;
        cudaEventElapsedTime(&ms, r.e0, r.e1);
        size_t i = 0;
        for (; i < acc.size(); ++i) if (acc[i].first == r.name) break;
        if (i == acc.size()) acc.push_back({r.name, Acc()});
        acc[i].second.ms += ms; acc[i].second.launches++; acc[i].second.units += r.units;
        cudaEventDestroy(r.e0); cudaEventDestroy(r.e1);
    }
    ctx->prof.clear();
    std::string s;
    for (auto &kv : acc) {
        char line[160];
        snprintf(line, sizeof line, "%s %.6f %lld %lld\n", kv.first.c_str(), kv.second.ms, kv.second.launches, kv.second.units);
        s += line;
    }
    if (s.size() + 1 > cap) return fail(B200_ESHAPE, "b200_profile_end: buffer too small");
    std::memcpy(out, s.c_str(), s.size() + 1);
    return 0;
}

extern "C" int b200_ctx_subring(const b200_ctx *ctx, int limb, uint64_t out[8]) {
    if (!ctx || !out || limb < 0 || limb >= ctx->L) return fail(B200_EINVAL, "b200_ctx_subring: bad argument");
    const SubRingHost &s = ctx->sub[limb];
    out[0] = s.q; out[1] = s.qinv; out[2] = s.brc[0]; out[3] = s.brc[1]; out[4] = s.mask;
    out[5] = s.g; out[6] = s.ninv; out[7] = s.root;
    return 0;
}
extern "C" int b200_ctx_roots(const b200_ctx *ctx, int limb, int inverse, uint64_t *out, size_t n) {
    if (!ctx || !out || limb < 0 || limb >= ctx->L) return fail(B200_EINVAL, "b200_ctx_roots: bad argument");
    if (ctx->ring_type == B200_RING_MATRIX) return fail(B200_EINVAL, "b200_ctx_roots: Matrix rings carry no RootsForward/RootsBackward");
    const std::vector<u64> &t = inverse ? ctx->sub[limb].roots_bwd : ctx->sub[limb].roots_fwd;
    if (n < t.size()) return fail(B200_ESHAPE, "b200_ctx_roots: len(out) < NthRoot/2");
    std::memcpy(out, t.data(), sizeof(u64) * t.size());
    return 0;
}

// ---- device polynomials ---------------------------------------------------------------------------
static size_t poly_words(const b200_poly *p) { return (size_t)p->batch * p->ctx->L * p->ctx->N; }

extern "C" int b200_poly_alloc(b200_ctx *ctx, int batch, b200_poly **out) {
    if (!ctx || !out || batch <= 0) return fail(B200_EINVAL, "b200_poly_alloc: bad argument");
    CUDA_TRY(cudaSetDevice(ctx->device));
    b200_poly *p = new b200_poly();
    p->ctx = ctx; p->batch = batch; p->owned = true;
    const size_t bytes = poly_words(p) * sizeof(u64);
    cudaError_t e = cudaMalloc(&p->d, bytes);
    if (e == cudaSuccess) e = cudaMemsetAsync(p->d, 0, bytes, ctx->stream);
    if (e != cudaSuccess) { cudaFree(p->d); delete p; return b200_cuda_fail(e, "b200_poly_alloc"); }
    *out = p;
    return 0;
}
extern "C" int b200_poly_wrap(b200_ctx *ctx, void *device_ptr, int batch, b200_poly **out) {
    if (!ctx || !out || !device_ptr || batch <= 0) return fail(B200_EINVAL, "b200_poly_wrap: bad argument");
    if ((uintptr_t)device_ptr % 16) return fail(B200_EINVAL, "b200_poly_wrap: device pointer must be 16-byte aligned");
    b200_poly *p = new b200_poly();
    p->ctx = ctx; p->batch = batch; p->owned = false; p->d = (u64 *)device_ptr;
    *out = p;
    return 0;
}
extern "C" void b200_poly_free(b200_poly *p) {
    if (!p) return;
    if (p->owned) { cudaSetDevice(p->ctx->device); cudaStreamSynchronize(p->ctx->stream); cudaFree(p->d); }
    delete p;
}
extern "C" void *b200_poly_device_ptr(b200_poly *p) { return p ? p->d : nullptr; }
extern "C" int b200_poly_batch(const b200_poly *p) { return p ? p->batch : B200_EINVAL; }

static int limb_check(const b200_poly *p, int b, int limb, const void *host, size_t n, const char *who) {
    if (!p || !host) return fail(B200_EINVAL, std::string(who) + ": NULL argument");
    if (b < 0 || b >= p->batch || limb < 0 || limb >= p->ctx->L) return fail(B200_ESHAPE, std::string(who) + ": poly/limb index out of range");
    if (n < (size_t)p->ctx->N) return fail(B200_ESHAPE, std::string(who) + ": len(slice)=" + std::to_string(n) + " < N=" + std::to_string(p->ctx->N));
    return 0;
}
extern "C" int b200_poly_upload_limb(b200_poly *p, int b, int limb, const uint64_t *host, size_t n) {
    CtxLock lk(p ? p->ctx : nullptr);
    if (int rc = limb_check(p, b, limb, host, n, "b200_poly_upload_limb")) return rc;
    b200_ctx *c = p->ctx;
    CUDA_TRY(cudaSetDevice(c->device));
    CUDA_TRY(cudaMemcpyAsync(p->d + ((size_t)b * c->L + limb) * c->N, host, sizeof(u64) * c->N, cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));   // the Go slice may be reused as soon as we return
    return 0;
}
extern "C" int b200_poly_download_limb(b200_poly *p, int b, int limb, uint64_t *host, size_t n) {
    CtxLock lk(p ? p->ctx : nullptr);
    if (int rc = limb_check(p, b, limb, host, n, "b200_poly_download_limb")) return rc;
    b200_ctx *c = p->ctx;
    CUDA_TRY(cudaSetDevice(c->device));
    CUDA_TRY(cudaMemcpyAsync(host, p->d + ((size_t)b * c->L + limb) * c->N, sizeof(u64) * c->N, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return 0;
}
extern "C" int b200_poly_upload(b200_poly *p, const uint64_t *host, size_t n_words) {
    CtxLock lk(p ? p->ctx : nullptr);
    if (!p || !host) return fail(B200_EINVAL, "b200_poly_upload: NULL argument");
    if (n_words != poly_words(p)) return fail(B200_ESHAPE, "b200_poly_upload: expected batch*L*N words");
    CUDA_TRY(cudaSetDevice(p->ctx->device));
    CUDA_TRY(cudaMemcpyAsync(p->d, host, n_words * sizeof(u64), cudaMemcpyHostToDevice, p->ctx->stream));
    return 0;   // asynchronous when `host` is pinned; ordered before later work on the stream
}
extern "C" int b200_poly_download(b200_poly *p, uint64_t *host, size_t n_words) {
    CtxLock lk(p ? p->ctx : nullptr);
    if (!p || !host) return fail(B200_EINVAL, "b200_poly_download: NULL argument");
    if (n_words != poly_words(p)) return fail(B200_ESHAPE, "b200_poly_download: expected batch*L*N words");
    CUDA_TRY(cudaSetDevice(p->ctx->device));
    CUDA_TRY(cudaMemcpyAsync(host, p->d, n_words * sizeof(u64), cudaMemcpyDeviceToHost, p->ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(p->ctx->stream));
    return 0;
}

// Poly.Copy / CopyLvl (ring/poly.go:73-91): limbs 0..level of every polynomial of src into dst, device to device.
// level < 0: all limbs.  Identical handles are skipped like the aliased limbs of the Go code (utils.Alias1D).
extern "C" int b200_poly_copy(b200_poly *dst, const b200_poly *src, int level) {
    CtxLock lk(dst ? dst->ctx : nullptr);
    if (!dst || !src) return fail(B200_EINVAL, "b200_poly_copy: NULL argument");
    if (dst->ctx != src->ctx) return fail(B200_ESHAPE, "b200_poly_copy: polynomials belong to different rings");
    if (dst->batch != src->batch) return fail(B200_ESHAPE, "b200_poly_copy: batch sizes differ");
    b200_ctx *ctx = dst->ctx;
    if (level < 0) level = ctx->L - 1;
    if (level >= ctx->L) return fail(B200_ESHAPE, "b200_poly_copy: level cannot be larger than max level");
    if (dst->d == src->d) return 0;
    CUDA_TRY(cudaSetDevice(ctx->device));
    const size_t pitch = (size_t)ctx->L * ctx->N * sizeof(u64);
    CUDA_TRY(cudaMemcpy2DAsync(dst->d, pitch, src->d, pitch, (size_t)(level + 1) * ctx->N * sizeof(u64), dst->batch,
                               cudaMemcpyDeviceToDevice, b200_cur_stream(ctx)));
    return 0;
}

// ---- wire format ------------------------------------------------------------------
// ring.Poly serialises as structs.Matrix[uint64] (ring/poly.go:97-160, utils/structs/matrix.go:57-128,
// utils/structs/vector.go:43-100): u64 row count, then per limb a u64 length followed by the words, all
// little-endian.  The device layout is the same limb-major stream without the 8-byte headers, so each limb is
// one contiguous copy straight between the wire buffer and HBM.
static inline void put_le64(unsigned char *p, u64 v) { for (int i = 0; i < 8; ++i) p[i] = (unsigned char)(v >> (8 * i)); }
static inline u64 get_le64(const unsigned char *p) { u64 v = 0; for (int i = 0; i < 8; ++i) v |= (u64)p[i] << (8 * i); return v; }

extern "C" int b200_poly_binary_size(const b200_ctx *ctx, int level, size_t *out) {
    if (!ctx || !out) return fail(B200_EINVAL, "b200_poly_binary_size: NULL argument");
    if (level < 0 || level >= ctx->L) return fail(B200_ESHAPE, "b200_poly_binary_size: level out of range");
    *out = 8 + (size_t)(level + 1) * (8 + (size_t)ctx->N * 8);
    return 0;
}
extern "C" int b200_poly_marshal(const b200_poly *p, int b, int level, void *buf, size_t cap, size_t *written) {
    CtxLock lk(p ? p->ctx : nullptr);
    if (!p || !buf) return fail(B200_EINVAL, "b200_poly_marshal: NULL argument");
    const b200_ctx *ctx = p->ctx;
    if (b < 0 || b >= p->batch) return fail(B200_ESHAPE, "b200_poly_marshal: polynomial index out of range");
    if (level < 0 || level >= ctx->L) return fail(B200_ESHAPE, "b200_poly_marshal: level out of range");
    const size_t N = (size_t)ctx->N, row = 8 + N * 8, need = 8 + (size_t)(level + 1) * row;
    if (cap < need) return fail(B200_ESHAPE, "b200_poly_marshal: buffer smaller than BinarySize()");
    CUDA_TRY(cudaSetDevice(ctx->device));
    unsigned char *w = static_cast<unsigned char *>(buf);
    put_le64(w, (u64)(level + 1));
    for (int l = 0; l <= level; ++l) {
        unsigned char *r = w + 8 + (size_t)l * row;
        put_le64(r, (u64)N);
        CUDA_TRY(cudaMemcpyAsync(r + 8, p->d + ((size_t)b * ctx->L + l) * N, N * 8, cudaMemcpyDeviceToHost, ctx->stream));
    }
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    if (written) *written = need;
    return 0;
}
extern "C" int b200_poly_unmarshal(b200_poly *p, int b, const void *buf, size_t len, int *level_out, size_t *consumed) {
    CtxLock lk(p ? p->ctx : nullptr);
    if (!p || !buf) return fail(B200_EINVAL, "b200_poly_unmarshal: NULL argument");
    const b200_ctx *ctx = p->ctx;
    if (b < 0 || b >= p->batch) return fail(B200_ESHAPE, "b200_poly_unmarshal: polynomial index out of range");
    const unsigned char *r = static_cast<const unsigned char *>(buf);
    if (len < 8) return fail(B200_ESHAPE, "b200_poly_unmarshal: truncated input (no row count)");
    const u64 rows = get_le64(r);
    if (rows > (u64)ctx->L) return fail(B200_ESHAPE, "b200_poly_unmarshal: more limbs than the ring has moduli");
    const size_t N = (size_t)ctx->N, row = 8 + N * 8;
    // validate every header before touching the device so a malformed stream leaves the polynomial unchanged
    size_t off = 8;
    for (u64 l = 0; l < rows; ++l, off += row) {
        if (len < off + 8) return fail(B200_ESHAPE, "b200_poly_unmarshal: truncated input (limb header)");
        if (get_le64(r + off) != (u64)N) return fail(B200_ESHAPE, "b200_poly_unmarshal: limb length differs from the ring degree");
        if (len < off + row) return fail(B200_ESHAPE, "b200_poly_unmarshal: truncated input (limb data)");
    }
    CUDA_TRY(cudaSetDevice(ctx->device));
    for (u64 l = 0; l < rows; ++l)
        CUDA_TRY(cudaMemcpyAsync(p->d + ((size_t)b * ctx->L + l) * N, r + 8 + (size_t)l * row + 8, N * 8, cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));   // the caller may reuse `buf` on return
    if (level_out) *level_out = (int)rows - 1;
    if (consumed) *consumed = off;
    return 0;
}
// SubRing.Factors (ring/subring.go:45): unique prime factors of Modulus-1, ascending -- what
// subRingParametersLiteral stores next to PrimitiveRoot (ring/subring.go:333-360).  Returns the count.
extern "C" int b200_ctx_factors(const b200_ctx *ctx, int limb, uint64_t *out, int cap) {
    if (!ctx) return fail(B200_EINVAL, "b200_ctx_factors: ctx is NULL");
    if (limb < 0 || limb >= ctx->L) return fail(B200_ESHAPE, "b200_ctx_factors: limb out of range");
    const std::vector<u64> f = hm::unique_factors(ctx->sub[limb].q - 1);
    if (out) {
        if (cap < (int)f.size()) return fail(B200_ESHAPE, "b200_ctx_factors: output too small");
        for (size_t i = 0; i < f.size(); ++i) out[i] = f[i];
    }
    return (int)f.size();
}

// ---- dispatch ---------------------------------------------------------------------------------------
static int check_polys(b200_ctx *ctx, int level, const char *who, std::initializer_list<const b200_poly *> ps, int *batch) {
    if (!ctx) return fail(B200_EINVAL, std::string(who) + ": ctx is NULL");
    if (level < 0) return fail(B200_ESHAPE, std::string(who) + ": level cannot be negative");                       // ring.go:197-199
    if (level >= ctx->L) return fail(B200_ESHAPE, std::string(who) + ": level cannot be larger than max level");    // ring.go:202-204
    int bt = -1;
    for (const b200_poly *p : ps) {
        if (!p) return fail(B200_EINVAL, std::string(who) + ": NULL polynomial");
        if (p->ctx != ctx) return fail(B200_ESHAPE, std::string(who) + ": polynomial belongs to another ring");
        if (bt < 0) bt = p->batch;
        if (p->batch != bt) return fail(B200_ESHAPE, std::string(who) + ": batch sizes differ");
    }
    *batch = bt;
    return 0;
}

// The transform of the ring: tuned two-pass kernels where the shape is covered, generic kernels otherwise.  Conjugate-
// invariant rings add their folding layer (launch_ci_fold): before the butterflies of the forward transform, after those
// of the inverse (ring/ntt.go:754-786,1104-1158).
int launch_ntt(b200_ctx *ctx, const u64 *in, u64 *out, Span sp, bool inverse, int flags) {
    const bool ci = ctx->ring_type == B200_RING_CONJUGATE_INVARIANT;
    if (ci && !inverse) {
        if (int rc = launch_ci_fold(ctx, in, out, sp, false)) return rc;
        in = out;
    }
    int rc = launch_ntt_big(ctx, in, out, sp, inverse, flags);
    if (rc == 1) rc = launch_ntt_generic(ctx, in, out, sp, inverse, flags);
    if (rc == 0 && ci && inverse) rc = launch_ci_fold(ctx, out, out, sp, true);
    return rc;
}

static int transform(b200_ctx *ctx, const b200_poly *in, b200_poly *out, int level, int flags, bool inverse, const char *who) {
    int batch;
    if (int rc = check_polys(ctx, level, who, {in, out}, &batch)) return rc;
    CUDA_TRY(cudaSetDevice(ctx->device));
    Span sp{batch, level + 1, ctx->L, 0};
    int rc = launch_ntt(ctx, in->d, out->d, sp, inverse, flags);
    return rc;
}
// Ring.NTT / INTT restricted to limbs [limb0, limb0+nl): what a rank of a limb-sharded ring runs on its share
extern "C" int b200_ntt_limbs(b200_ctx *ctx, const b200_poly *in, b200_poly *out, int limb0, int nl, int inverse, int flags) {
    CtxLock lk(ctx);
    int batch;
    if (int rc = check_polys(ctx, 0, "b200_ntt_limbs", {in, out}, &batch)) return rc;
    if (limb0 < 0 || nl < 0 || limb0 + nl > ctx->L) return fail(B200_ESHAPE, "b200_ntt_limbs: limb range out of bounds");
    if (nl == 0) return 0;
    CUDA_TRY(cudaSetDevice(ctx->device));
    const size_t off = (size_t)limb0 * ctx->N;
    Span sp{batch, nl, ctx->L, limb0};
    int rc = launch_ntt(ctx, in->d + off, out->d + off, sp, inverse != 0, flags);
    return rc;
}

// ---- fused all-gather over NVLink peer memory (key-switch exchanges) --------------------------
// One process per GPU: a rank exports its (library-allocated) polynomial with CUDA IPC, the peers map it, and the
// inverse transform of the owned limbs writes its result into every rank's copy from inside the last kernel.
extern "C" int b200_poly_ipc_export(const b200_poly *p, unsigned char handle[64]) {
    if (!p || !handle) return fail(B200_EINVAL, "b200_poly_ipc_export: NULL argument");
    if (!p->owned) return fail(B200_EINVAL, "b200_poly_ipc_export: only polynomials allocated by b200_poly_alloc can be exported");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    CUDA_TRY(cudaSetDevice(p->ctx->device));
    cudaIpcMemHandle_t h;
    CUDA_TRY(cudaIpcGetMemHandle(&h, p->d));
    std::memcpy(handle, &h, 64);
    return 0;
}
extern "C" int b200_peer_open(b200_ctx *ctx, const unsigned char handle[64], void **device_ptr) {
    if (!ctx || !handle || !device_ptr) return fail(B200_EINVAL, "b200_peer_open: NULL argument");
    CUDA_TRY(cudaSetDevice(ctx->device));
    cudaIpcMemHandle_t h;
    std::memcpy(&h, handle, 64);
    CUDA_TRY(cudaIpcOpenMemHandle(device_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return 0;
}
extern "C" int b200_peer_close(b200_ctx *ctx, void *device_ptr) {
    if (!ctx || !device_ptr) return fail(B200_EINVAL, "b200_peer_close: NULL argument");
    CUDA_TRY(cudaSetDevice(ctx->device));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    CUDA_TRY(cudaIpcCloseMemHandle(device_ptr));
    return 0;
}

// shape-agnostic fallback of the fused gather: copy the freshly written limbs to the peers
__global__ void __launch_bounds__(256)
k_scatter_limbs(const u64 *src, PeerOut po, Span sp, int N) {
    const int unit = blockIdx.y, b = unit / sp.nl, l = unit - b * sp.nl;
    const size_t base = ((size_t)b * sp.Lbuf + l) * (size_t)N;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < N; i += gridDim.x * blockDim.x) {
        const u64 v = src[base + i];
        for (int k = 0; k < po.n; ++k) po.p[k][base + i] = v;
    }
}

// Ring.INTT on limbs [limb0, limb0+nl) of `in` into `out` AND into the same limbs of npeers polynomials that live on
// other GPUs (device pointers from b200_peer_open, same [batch][L][N] layout): the all-gather of the key switch
// (coefficient-domain limbs before the basis extension, basis_extension.go:381-502) done by the producing kernel.
// The caller orders the peers' reads after this call with any cross-GPU barrier on the stream.
extern "C" int b200_intt_limbs_gather(b200_ctx *ctx, const b200_poly *in, b200_poly *out, int limb0, int nl, int flags,
                                      int npeers, void *const *peer_polys) {
    CtxLock lk(ctx);
    int batch;
    if (int rc = check_polys(ctx, 0, "b200_intt_limbs_gather", {in, out}, &batch)) return rc;
    if (limb0 < 0 || nl < 0 || limb0 + nl > ctx->L) return fail(B200_ESHAPE, "b200_intt_limbs_gather: limb range out of bounds");
    if (npeers < 0 || npeers > 7 || (npeers && !peer_polys)) return fail(B200_EINVAL, "b200_intt_limbs_gather: at most 7 peers");
    if (nl == 0) return 0;
    CUDA_TRY(cudaSetDevice(ctx->device));
    const size_t off = (size_t)limb0 * ctx->N;
    PeerOut po;
    po.n = npeers;
    for (int i = 0; i < npeers; ++i) {
        if (!peer_polys[i]) return fail(B200_EINVAL, "b200_intt_limbs_gather: NULL peer pointer");
        po.p[i] = static_cast<u64 *>(peer_polys[i]) + off;
    }
    Span sp{batch, nl, ctx->L, limb0};
    // the fused peer stores live in the last butterfly pass, which is not the last step of a conjugate-invariant transform
    int rc = ctx->ring_type == B200_RING_CONJUGATE_INVARIANT ? 1 : launch_ntt_big(ctx, in->d + off, out->d + off, sp, true, flags, &po);
    if (rc != 1) return rc;
    rc = launch_ntt(ctx, in->d + off, out->d + off, sp, true, flags);
    if (rc || npeers == 0) return rc;
    int chunks = (ctx->N + 255) / 256;
    if (chunks > 64) chunks = 64;
    {
        LaunchScope ls(ctx, "scatter_limbs", (long long)batch * nl);
        k_scatter_limbs<<<dim3(chunks, batch * nl), 256, 0, b200_cur_stream(ctx)>>>(out->d + off, po, sp, ctx->N);
    }
    CUDA_TRY(cudaGetLastError());
    return 0;
}

extern "C" int b200_ntt(b200_ctx *ctx, const b200_poly *in, b200_poly *out, int level, int flags) {
    CtxLock lk(ctx);
    return transform(ctx, in, out, level, flags, false, "b200_ntt");
}
extern "C" int b200_intt(b200_ctx *ctx, const b200_poly *in, b200_poly *out, int level, int flags) {
    CtxLock lk(ctx);
    return transform(ctx, in, out, level, flags, true, "b200_intt");
}

extern "C" int b200_vec_op(b200_ctx *ctx, int op, const b200_poly *p1, const b200_poly *p2, b200_poly *p3, int level,
                           const uint64_t *scalars) {
    CtxLock lk(ctx);
    int batch;
    const bool binary = op <= B200_OP_SUB_LAZY || (op >= B200_OP_MUL_BARRETT && op <= B200_OP_MUL_MONT_LAZY_NEG) ||
                        op == B200_OP_SUB_MULSCALAR_MONT_2Q || op == B200_OP_ADD_LAZY_MULSCALAR_MONT;
    if (!binary) {
        if (int rc = check_polys(ctx, level, "b200_vec_op", {op == B200_OP_ZERO && !p1 ? p3 : p1, p3}, &batch)) return rc;
    } else {
        if (int rc = check_polys(ctx, level, "b200_vec_op", {p1, p2, p3}, &batch)) return rc;
    }
    if (op >= B200_OP_ADD_SCALAR_DOUBLE && op <= B200_OP_MULSCALAR_MONT_ADD_DOUBLE && (ctx->N & 3))
        return fail(B200_ESHAPE, "b200_vec_op: the DoubleRNSScalar operators need N % 4 == 0");
    CUDA_TRY(cudaSetDevice(ctx->device));
    const u64 *d_sc = nullptr;
    if (op >= B200_OP_MULSCALAR_MONT) {
        if (!scalars) return fail(B200_EINVAL, "b200_vec_op: scalar op needs one scalar per limb");
        const size_t ns = (size_t)(level + 1) * (op >= B200_OP_ADD_SCALAR_LAZY_MULSCALAR_MONT ? 2 : 1);
        // small per-call constant block; lives in the scratch-independent misc area of the stream order
        u64 *tmp = nullptr;
        CUDA_TRY(cudaMallocAsync(&tmp, sizeof(u64) * ns, b200_cur_stream(ctx)));
        CUDA_TRY(cudaMemcpyAsync(tmp, scalars, sizeof(u64) * ns, cudaMemcpyHostToDevice, b200_cur_stream(ctx)));
        CUDA_TRY(cudaStreamSynchronize(b200_cur_stream(ctx)));   // `scalars` is a pageable Go/host slice
        d_sc = tmp;
    }
    int rc = launch_vec_op(ctx, op, p1 ? p1->d : p3->d, p2 ? p2->d : nullptr, p3->d, Span{batch, level + 1, ctx->L, 0}, d_sc);
    if (d_sc) cudaFreeAsync((void *)d_sc, b200_cur_stream(ctx));
    return rc;
}

extern "C" int b200_mul_by_vector_montgomery(b200_ctx *ctx, const b200_poly *p1, const b200_poly *vector, b200_poly *p2,
                                             int level, int then_add_lazy) {
    CtxLock lk(ctx);
    int batch;
    if (int rc = check_polys(ctx, level, "b200_mul_by_vector_montgomery", {p1, p2}, &batch)) return rc;
    if (!vector || vector->ctx != ctx) return fail(B200_ESHAPE, "b200_mul_by_vector_montgomery: vector must be a polynomial of this ring");
    CUDA_TRY(cudaSetDevice(ctx->device));
    return launch_mul_by_vector(ctx, p1->d, vector->d, p2->d, Span{batch, level + 1, ctx->L, 0}, then_add_lazy != 0);
}

// Ring.DivRoundByLastModulusNTT / DivFloorByLastModulusNTT (ring/scaling.go:6-19,130-156) for the whole batch:
// INTT of the last limb -> spread to the lower limbs -> NTT of those -> SubThenMulScalarMontgomeryTwoModulus.
static int div_ntt_locked(b200_ctx *ctx, const b200_poly *p0, b200_poly *p1, int level, int round);
extern "C" int b200_div_by_last_modulus_ntt(b200_ctx *ctx, const b200_poly *p0, b200_poly *p1, int level, int round) {
    CtxLock lk(ctx);
    return div_ntt_locked(ctx, p0, p1, level, round);
}
static int div_ntt_locked(b200_ctx *ctx, const b200_poly *p0, b200_poly *p1, int level, int round) {
    int batch;
    if (int rc = check_polys(ctx, level, "b200_div_by_last_modulus_ntt", {p0, p1}, &batch)) return rc;
    if (level < 1) return fail(B200_ESHAPE, "b200_div_by_last_modulus_ntt: needs level >= 1");
    CUDA_TRY(cudaSetDevice(ctx->device));
    const size_t words = (size_t)batch * ctx->L * ctx->N;
    if (int rc = b200_ensure_scratch2(ctx, words)) return rc;
    u64 *ws = ctx->d_ws;
    const size_t off = (size_t)level * ctx->N;
    auto ntt = [&](const u64 *src, u64 *dst, Span sp, bool inv) {
        int r = launch_ntt(ctx, src, dst, sp, inv, 0);
        return r;
    };
    int rc = ntt(p0->d + off, ws + off, Span{batch, 1, ctx->L, level}, true);          // scaling.go:134
    if (rc) return rc;
    if ((rc = launch_rescale_spread(ctx, ws, batch, level, round != 0))) return rc;      // :139-142
    if ((rc = ntt(ws, ws, Span{batch, level, ctx->L, 0}, false))) return rc;             // :143
    return launch_vec_op(ctx, B200_OP_SUB_MULSCALAR_MONT_2Q, ws, p0->d, p1->d, Span{batch, level, ctx->L, 0},
                         ctx->d_rescale + (size_t)(level - 1) * ctx->L);                         // :144
}

// Ring.DivFloorByLastModulus / DivRoundByLastModulus (ring/scaling.go:21-30,113-127), coefficient domain
extern "C" int b200_div_by_last_modulus(b200_ctx *ctx, const b200_poly *p0, b200_poly *p1, int level, int round) {
    CtxLock lk(ctx);
    int batch;
    if (int rc = check_polys(ctx, level, "b200_div_by_last_modulus", {p0, p1}, &batch)) return rc;
    if (level < 1) return fail(B200_ESHAPE, "b200_div_by_last_modulus: needs level >= 1");
    CUDA_TRY(cudaSetDevice(ctx->device));
    return launch_div_last(ctx, p0->d, p1->d, batch, level, round != 0);
}

// Ring.Div{Floor,Round}ByLastModulusMany[NTT] (ring/scaling.go:32-91,130-192): nb_rescales divisions by the last moduli
// of levels level, level-1, ...; p1 limbs 0..level-nb_rescales hold the result.
extern "C" int b200_div_by_last_modulus_many(b200_ctx *ctx, const b200_poly *p0, b200_poly *p1, int level, int nb_rescales,
                                             int round, int ntt) {
    CtxLock lk(ctx);
    int batch;
    if (int rc = check_polys(ctx, level, "b200_div_by_last_modulus_many", {p0, p1}, &batch)) return rc;
    if (nb_rescales < 0 || nb_rescales > level) return fail(B200_ESHAPE, "b200_div_by_last_modulus_many: nb_rescales must be in [0, level]");
    CUDA_TRY(cudaSetDevice(ctx->device));
    const size_t N = ctx->N, words = (size_t)batch * ctx->L * N;
    if (nb_rescales == 0) {   // :33-36: p1.Copy(p0)
        if (p0->d != p1->d)
            CUDA_TRY(cudaMemcpy2DAsync(p1->d, ctx->L * N * sizeof(u64), p0->d, ctx->L * N * sizeof(u64), (size_t)(level + 1) * N * sizeof(u64), batch,
                                       cudaMemcpyDeviceToDevice, b200_cur_stream(ctx)));
        return 0;
    }
    if (ntt && nb_rescales == 1) return div_ntt_locked(ctx, p0, p1, level, round);   // :150-152 (and the same residues for the floor form)
    if (!ntt && nb_rescales == 1) return launch_div_last(ctx, p0->d, p1->d, batch, level, round != 0);
    if (int rc = b200_ensure_scratch2(ctx, words)) return rc;
    u64 *ws = ctx->d_ws;
    int rc = 0;
    const u64 *src = p0->d;
    if (ntt) {
        if ((rc = launch_ntt(ctx, p0->d, ws, Span{batch, level + 1, ctx->L, 0}, true, 0))) return rc;   // rCpy.INTT(p0, buff)
        src = ws;
    }
    for (int i = 0; i < nb_rescales; ++i) {
        u64 *dst = (!ntt && i == nb_rescales - 1) ? p1->d : ws;
        if ((rc = launch_div_last(ctx, src, dst, batch, level - i, round != 0))) return rc;
        src = ws;
    }
    if (ntt) rc = launch_ntt(ctx, ws, p1->d, Span{batch, level + 1 - nb_rescales, ctx->L, 0}, false, 0);   // rCpy.NTT(buff, p1)
    return rc;
}

extern "C" int b200_tensor(b200_ctx *ctx, const b200_poly *a0, const b200_poly *a1, const b200_poly *b0, const b200_poly *b1,
                           b200_poly *c0, b200_poly *c1, b200_poly *c2, int level, int mform_first) {
    CtxLock lk(ctx);
    int batch;
    if (int rc = check_polys(ctx, level, "b200_tensor", {a0, a1, b0, b1, c0, c1, c2}, &batch)) return rc;
    if (c0 == c1 || c0 == c2 || c1 == c2 || c0->d == c1->d || c0->d == c2->d || c1->d == c2->d)
        return fail(B200_ESHAPE, "b200_tensor: the three outputs must be distinct polynomials");
    CUDA_TRY(cudaSetDevice(ctx->device));
    return launch_tensor(ctx, a0->d, a1->d, b0->d, b1->d, c0->d, c1->d, c2->d, Span{batch, level + 1, ctx->L, 0}, mform_first != 0);
}

// Ring.Shift (ring/operations.go:278-282): p2 = p1 rotated left by k positions, ALL limbs of the polynomial (the Go loop
// ranges over p1.Coeffs, not over the ring's level).  Ring.MultByMonomial (:306-363): p2 = X^k * p1 on limbs 0..level.
// Both accept p1 == p2 like the reference (the operand is staged).
static int rotate_common(b200_ctx *ctx, const b200_poly *p1, b200_poly *p2, int nl, int s, int neg, int mode, const char *who) {
    int batch;
    if (int rc = check_polys(ctx, nl - 1, who, {p1, p2}, &batch)) return rc;
    CUDA_TRY(cudaSetDevice(ctx->device));
    const u64 *src = p1->d;
    if (p1->d == p2->d) {
        const size_t words = (size_t)batch * ctx->L * ctx->N;
        if (int rc = b200_ensure_scratch(ctx, words)) return rc;
        CUDA_TRY(cudaMemcpyAsync(ctx->d_scratch, p1->d, words * sizeof(u64), cudaMemcpyDeviceToDevice, b200_cur_stream(ctx)));
        src = ctx->d_scratch;
    }
    return launch_rotate(ctx, src, p2->d, batch, nl, s, neg, mode);
}
extern "C" int b200_shift(b200_ctx *ctx, const b200_poly *p1, int k, b200_poly *p2) {
    CtxLock lk(ctx);
    if (!ctx) return fail(B200_EINVAL, "b200_shift: ctx is NULL");
    int s = k % ctx->N;
    if (s < 0) s += ctx->N;
    return rotate_common(ctx, p1, p2, ctx->L, s, 0, 0, "b200_shift");
}
extern "C" int b200_mult_by_monomial(b200_ctx *ctx, const b200_poly *p1, int k, b200_poly *p2, int level) {
    CtxLock lk(ctx);
    if (!ctx) return fail(B200_EINVAL, "b200_mult_by_monomial: ctx is NULL");
    if (level < 0 || level >= ctx->L) return fail(B200_ESHAPE, "b200_mult_by_monomial: level out of range");
    const long long twoN = 2LL * ctx->N;
    const long long shift = (((long long)k + twoN) % twoN + twoN) % twoN;      // (k + 2N) % 2N of the reference for k > -2N
    if (shift == 0) {                                                          // :312-320: plain copy of limbs 0..level
        if (p1 && p2 && p1->d == p2->d) return 0;
        return rotate_common(ctx, p1, p2, level + 1, 0, 0, 0, "b200_mult_by_monomial");
    }
    return rotate_common(ctx, p1, p2, level + 1, (int)(shift % ctx->N), shift >= ctx->N ? 1 : 0, 1, "b200_mult_by_monomial");
}

// ---- automorphisms X -> X^gen (ring/automorphism.go) ------------------------------------------------------------
static u64 ring_nthroot(const b200_ctx *ctx) {
    return (ctx->ring_type == B200_RING_MATRIX ? 3 : ctx->ring_type == B200_RING_CONJUGATE_INVARIANT ? 4 : 2) * (u64)ctx->N;
}
// AutomorphismNTTIndex (:12-35), a pure host function
extern "C" int b200_automorphism_ntt_index(int N, uint64_t NthRoot, uint64_t GalEl, uint64_t *index) {
    if (!index) return fail(B200_EINVAL, "b200_automorphism_ntt_index: index is NULL");
    if (N <= 0 || (N & (N - 1))) return fail(B200_EINVAL, "N must be a power of two");                 // :14-16
    if (NthRoot == 0 || (NthRoot & (NthRoot - 1))) return fail(B200_EINVAL, "NthRoot must be w power of two");   // :18-20
    int lognth = 0;
    while (((u64)1 << (lognth + 1)) < NthRoot) lognth++;
    const u64 mask = NthRoot - 1;
    for (int i = 0; i < N; ++i) {
        const u64 t1 = 2 * hm::bitrev((u64)i, lognth) + 1;
        index[i] = hm::bitrev((((GalEl * t1) & mask) - 1) >> 1, lognth);
    }
    return 0;
}
static int stage_index(b200_ctx *ctx, const uint64_t *index, size_t n, size_t need, u64 bound, const char *who, const u64 **d_out) {
    if (!index || n < need) return fail(B200_ESHAPE, std::string(who) + ": len(index) < N");
    for (size_t i = 0; i < need; ++i)
        if (index[i] >= bound) return fail(B200_ESHAPE, std::string(who) + ": index out of range");   // the Go loop panics
    if (ctx->index_words < need) {
        CUDA_TRY(cudaStreamSynchronize(b200_cur_stream(ctx)));
        if (ctx->d_index) CUDA_TRY(cudaFree(ctx->d_index));
        ctx->d_index = nullptr; ctx->index_words = 0;
        CUDA_TRY(cudaMalloc(&ctx->d_index, need * sizeof(u64)));
        ctx->index_words = need;
    }
    // stream-ordered after the previous call's kernel; a pageable source is staged before the call returns
    CUDA_TRY(cudaMemcpyAsync(ctx->d_index, index, need * sizeof(u64), cudaMemcpyHostToDevice, b200_cur_stream(ctx)));
    *d_out = ctx->d_index;
    return 0;
}
// Ring.AutomorphismNTT (:38-46) with add_lazy == 0, the table computed on the device; add_lazy != 0 accumulates like
// AutomorphismNTTWithIndexThenAddLazy (:82-110).  Not in place (:37).
extern "C" int b200_automorphism_ntt(b200_ctx *ctx, const b200_poly *in, uint64_t gen, b200_poly *out, int level, int add_lazy) {
    CtxLock lk(ctx);
    int batch;
    if (int rc = check_polys(ctx, level, "b200_automorphism_ntt", {in, out}, &batch)) return rc;
    if (ctx->ring_type == B200_RING_MATRIX) return fail(B200_ESHAPE, "N must be a power of two");          // panic(err), :41-43
    if (in->d == out->d) return fail(B200_ESHAPE, "b200_automorphism_ntt: the result cannot be in-place");
    const u64 nth = ring_nthroot(ctx);
    if (ctx->ring_type == B200_RING_CONJUGATE_INVARIANT && (gen & 3) != 1)   // slots >= N of the 2N-point transform: the Go loop indexes out of range
        return fail(B200_ESHAPE, "b200_automorphism_ntt: index out of range (conjugate-invariant rings need gen = 1 mod 4)");
    CUDA_TRY(cudaSetDevice(ctx->device));
    return launch_automorphism_ntt(ctx, in->d, out->d, nullptr, gen, nth, ctx->N, ctx->L, batch, level + 1, add_lazy ? 1 : 0);
}
// Ring.AutomorphismNTTWithIndex (:50-79) / ...ThenAddLazy (:82-110) with the caller's table
extern "C" int b200_automorphism_ntt_with_index(b200_ctx *ctx, const b200_poly *in, const uint64_t *index, size_t n, b200_poly *out,
                                                int level, int add_lazy) {
    CtxLock lk(ctx);
    int batch;
    if (int rc = check_polys(ctx, level, "b200_automorphism_ntt_with_index", {in, out}, &batch)) return rc;
    if (in->d == out->d) return fail(B200_ESHAPE, "b200_automorphism_ntt_with_index: the result cannot be in-place");
    CUDA_TRY(cudaSetDevice(ctx->device));
    const u64 *d_index;
    if (int rc = stage_index(ctx, index, n, (size_t)ctx->N, (u64)ctx->N, "b200_automorphism_ntt_with_index", &d_index)) return rc;
    return launch_automorphism_ntt(ctx, in->d, out->d, d_index, 0, ring_nthroot(ctx), ctx->N, ctx->L, batch, level + 1, add_lazy ? 1 : 0);
}
// Ring.Automorphism (:113-176), coefficient domain; not in place
extern "C" int b200_automorphism(b200_ctx *ctx, const b200_poly *in, uint64_t gen, b200_poly *out, int level) {
    CtxLock lk(ctx);
    int batch;
    if (int rc = check_polys(ctx, level, "b200_automorphism", {in, out}, &batch)) return rc;
    if (ctx->ring_type == B200_RING_MATRIX) return fail(B200_ESHAPE, "b200_automorphism: N must be a power of two");
    if (in->d == out->d) return fail(B200_ESHAPE, "b200_automorphism: the result cannot be in-place");
    if (!(gen & 1)) return fail(B200_ESHAPE, "b200_automorphism: gen must be odd");   // an even gen is not a permutation: racing writes
    CUDA_TRY(cudaSetDevice(ctx->device));
    return launch_automorphism_coeff(ctx, in->d, out->d, gen, batch, level + 1);
}
// Ring.FoldStandardToConjugateInvariant (ring/conjugate_invariant.go:28-44): ctx is the conjugate-invariant ring (degree N),
// st a polynomial of a ring of degree 2N over the same moduli, index the table permuteNTTIndexInv (N entries < 2N):
// ci[j] = st[index[j]] + st[j] mod q.
extern "C" int b200_fold_standard_to_ci(b200_ctx *ctx, const b200_poly *st, const uint64_t *index, size_t n, b200_poly *ci, int level) {
    CtxLock lk(ctx);
    if (!ctx || !ci || !st || ci->ctx != ctx || !st->ctx) return fail(B200_EINVAL, "b200_fold_standard_to_ci: NULL argument or foreign polynomial");
    if (st->ctx->N != 2 * ctx->N)   // :31-33 (panic)
        return fail(B200_ESHAPE, "cannot FoldStandardToConjugateInvariant: Ring degree of polyStandard must be 2N and ring degree of polyConjugateInvariant must be N");
    if (level < 0 || level >= ctx->L || level >= st->ctx->L) return fail(B200_ESHAPE, "b200_fold_standard_to_ci: level out of range");
    if (ci->batch != st->batch) return fail(B200_ESHAPE, "b200_fold_standard_to_ci: batch sizes differ");
    if (ci->ctx->device != st->ctx->device) return fail(B200_ESHAPE, "b200_fold_standard_to_ci: polynomials live on different devices");
    for (int i = 0; i <= level; ++i)
        if (ctx->sub[i].q != st->ctx->sub[i].q) return fail(B200_ESHAPE, "b200_fold_standard_to_ci: the rings do not share their moduli");
    CUDA_TRY(cudaSetDevice(ctx->device));
    const u64 *d_index;
    if (int rc = stage_index(ctx, index, n, (size_t)ctx->N, 2 * (u64)ctx->N, "b200_fold_standard_to_ci", &d_index)) return rc;
    return launch_automorphism_ntt(ctx, st->d, ci->d, d_index, 0, 4 * (u64)ctx->N, st->ctx->N, st->ctx->L, ci->batch, level + 1, 2);
}

// Ring.UnfoldConjugateInvariantToStandard (ring/conjugate_invariant.go:7-23): limbs 0..level of a polynomial of a
// conjugate-invariant ring of degree N into a polynomial of a ring of degree 2N over the same moduli.
extern "C" int b200_unfold_ci_to_standard(b200_ctx *ctx, const b200_poly *ci, b200_poly *st, int level) {
    CtxLock lk(ctx);
    if (!ctx || !ci || !st || ci->ctx != ctx || !st->ctx) return fail(B200_EINVAL, "b200_unfold_ci_to_standard: NULL argument or foreign polynomial");
    if (2 * ctx->N != st->ctx->N)   // conjugate_invariant.go:10-12 (panic)
        return fail(B200_ESHAPE, "cannot UnfoldConjugateInvariantToStandard: Ring degree of polyConjugateInvariant must be twice the ring degree of polyStandard");
    if (level < 0 || level >= ctx->L || level >= st->ctx->L) return fail(B200_ESHAPE, "b200_unfold_ci_to_standard: level out of range");
    if (ci->batch != st->batch) return fail(B200_ESHAPE, "b200_unfold_ci_to_standard: batch sizes differ");
    if (ci->ctx->device != st->ctx->device) return fail(B200_ESHAPE, "b200_unfold_ci_to_standard: polynomials live on different devices");
    CUDA_TRY(cudaSetDevice(ctx->device));
    return launch_ci_unfold(ctx, ci->d, st->d, ctx->N, ci->batch, level + 1, ctx->L, st->ctx->L);
}

extern "C" int b200_polymul(b200_ctx *ctx, const b200_poly *a, const b200_poly *b, b200_poly *c, int level) {
    CtxLock lk(ctx);
    int batch;
    if (int rc = check_polys(ctx, level, "b200_polymul", {a, b, c}, &batch)) return rc;
    CUDA_TRY(cudaSetDevice(ctx->device));
    Span sp{batch, level + 1, ctx->L, 0};
    int rc = ctx->ring_type == B200_RING_CONJUGATE_INVARIANT ? 1 : launch_polymul_big(ctx, a->d, b->d, c->d, sp);   // the fused rows skip the folding layer
    if (rc != 1) return rc;
    // generic composition: NTT(a), NTT(b) in tree order -> Barrett product (== MRed(MForm(a^), b^), canonical,
    // ring/ring_test.go:106-110) -> INTT.  Two scratch polynomials.
    const size_t words = (size_t)batch * ctx->L * ctx->N;
    if (int r = b200_ensure_scratch(ctx, 2 * words)) return r;
    u64 *ta = ctx->d_scratch, *tb = ctx->d_scratch + words;
    const int tf = ctx->ring_type == B200_RING_MATRIX ? B200_TREE_ORDER : 0;
    auto ntt = [&](const u64 *src, u64 *dst, bool inv) {
        int r = launch_ntt(ctx, src, dst, sp, inv, tf);
        return r;
    };
    if ((rc = ntt(a->d, ta, false))) return rc;
    if ((rc = ntt(b->d, tb, false))) return rc;
    if ((rc = launch_vec_op(ctx, B200_OP_MUL_BARRETT, ta, tb, ta, sp, nullptr))) return rc;
    return ntt(ta, c->d, true);
}

// ---- pipelined host-buffer transforms ---------------------------------------------------------------------
// The call a caller WITHOUT device-resident polynomials makes: the batch streams through the GPU in chunks,
// H2D copy -> transforms -> D2H copy on three streams so that both PCIe directions and the SMs overlap.
extern "C" int b200_transform_host(b200_ctx *ctx, const uint64_t *host_in, uint64_t *host_out, int batch, int level,
                                   const int *ops, int n_ops, int flags) {
    CtxLock lk(ctx);
    if (!ctx || !host_in || !host_out || !ops) return fail(B200_EINVAL, "b200_transform_host: NULL argument");
    if (batch <= 0 || n_ops <= 0) return fail(B200_EINVAL, "b200_transform_host: empty batch or op list");
    if (level < 0 || level >= ctx->L) return fail(B200_ESHAPE, "b200_transform_host: level out of range");
    CUDA_TRY(cudaSetDevice(ctx->device));
    typedef b200_ctx::HostPipe Pipe;
    Pipe &p = ctx->pipe;
    const size_t poly_words = (size_t)ctx->L * ctx->N;
    // chunk = as many polynomials as fit ~128 MiB (at least 1): big enough for PCIe efficiency, small enough to overlap
    int cpolys = (int)std::max<size_t>(1, ((size_t)128 << 20) / (poly_words * sizeof(u64)));
    cpolys = std::min(cpolys, batch);
    if (!p.ready) {
        CUDA_TRY(cudaStreamCreateWithFlags(&p.h2d, cudaStreamNonBlocking));
        CUDA_TRY(cudaStreamCreateWithFlags(&p.comp, cudaStreamNonBlocking));
        CUDA_TRY(cudaStreamCreateWithFlags(&p.d2h, cudaStreamNonBlocking));
        for (int i = 0; i < Pipe::NB; ++i) {
            CUDA_TRY(cudaEventCreateWithFlags(&p.up[i], cudaEventDisableTiming));
            CUDA_TRY(cudaEventCreateWithFlags(&p.done[i], cudaEventDisableTiming));
            CUDA_TRY(cudaEventCreateWithFlags(&p.freed[i], cudaEventDisableTiming));
        }
        CUDA_TRY(cudaEventCreateWithFlags(&p.entry, cudaEventDisableTiming));
        p.ready = true;
    }
    if (p.buf_words < (size_t)cpolys * poly_words) {
        CUDA_TRY(cudaDeviceSynchronize());
        for (int i = 0; i < Pipe::NB; ++i) {
            if (p.buf[i]) CUDA_TRY(cudaFree(p.buf[i]));
            p.buf[i] = nullptr;
            CUDA_TRY(cudaMalloc(&p.buf[i], (size_t)cpolys * poly_words * sizeof(u64)));
        }
        p.buf_words = (size_t)cpolys * poly_words;
    }
    int rc = 0;
    // The transforms below run on the pipeline's own stream but share the context scratch (slot permutation, layered
    // large-N path) with whatever was queued earlier on the context's stream: wait for that work first.  The pipeline is
    // drained before this call returns, so later work on the context's stream cannot overlap it either.
    CUDA_TRY(cudaEventRecord(p.entry, ctx->stream));
    CUDA_TRY(cudaStreamWaitEvent(p.comp, p.entry, 0));
    b200_set_stream_override(p.comp, true);
    // scratch for the slot permutation of MATRIX rings is sized once, before any work is in flight
    if (ctx->ring_type == B200_RING_MATRIX && !(flags & B200_TREE_ORDER))
        rc = b200_ensure_scratch(ctx, (size_t)cpolys * poly_words);
    int nchunks = (batch + cpolys - 1) / cpolys;
    for (int c = 0; c < nchunks && rc == 0; ++c) {
        const int b0 = c * cpolys, nb = std::min(cpolys, batch - b0), s = c % Pipe::NB;
        const size_t words = (size_t)nb * poly_words;
        cudaError_t e = cudaSuccess;
        if (c >= Pipe::NB) e = cudaStreamWaitEvent(p.h2d, p.freed[s], 0);          // staging buffer drained
        if (e == cudaSuccess) e = cudaMemcpyAsync(p.buf[s], host_in + (size_t)b0 * poly_words, words * sizeof(u64), cudaMemcpyHostToDevice, p.h2d);
        if (e == cudaSuccess) e = cudaEventRecord(p.up[s], p.h2d);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(p.comp, p.up[s], 0);
        if (e != cudaSuccess) { rc = b200_cuda_fail(e, "b200_transform_host upload"); break; }
        Span sp{nb, level + 1, ctx->L, 0};
        for (int o = 0; o < n_ops && rc == 0; ++o) {
            const bool inv = ops[o] != 0;
            rc = launch_ntt(ctx, p.buf[s], p.buf[s], sp, inv, flags);
        }
        if (rc) break;
        e = cudaEventRecord(p.done[s], p.comp);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(p.d2h, p.done[s], 0);
        if (e == cudaSuccess) e = cudaMemcpyAsync(host_out + (size_t)b0 * poly_words, p.buf[s], words * sizeof(u64), cudaMemcpyDeviceToHost, p.d2h);
        if (e == cudaSuccess) e = cudaEventRecord(p.freed[s], p.d2h);
        if (e != cudaSuccess) { rc = b200_cuda_fail(e, "b200_transform_host download"); break; }
    }
    b200_set_stream_override(nullptr, false);
    cudaError_t e = cudaStreamSynchronize(p.d2h);
    if (e == cudaSuccess) e = cudaStreamSynchronize(p.comp);
    if (e == cudaSuccess) e = cudaStreamSynchronize(p.h2d);
    if (rc == 0 && e != cudaSuccess) rc = b200_cuda_fail(e, "b200_transform_host sync");
    return rc;
}

extern "C" int b200_subring_ntt_host(b200_ctx *ctx, int limb, const uint64_t *p1, uint64_t *p2, size_t n, int inverse, int flags) {
    CtxLock lk(ctx);
    if (!ctx || !p1 || !p2) return fail(B200_EINVAL, "b200_subring_ntt_host: NULL argument");
    if (limb < 0 || limb >= ctx->L) return fail(B200_ESHAPE, "b200_subring_ntt_host: limb out of range");
    if (n < (size_t)ctx->N)   // ring/ntt.go:211-214, ntt_3n.go:85-87
        return fail(B200_ESHAPE, "cannot NTT: ensure that len(p1), len(p2)=" + std::to_string(n) + " >= N=" + std::to_string(ctx->N));
    CUDA_TRY(cudaSetDevice(ctx->device));
    u64 *d = nullptr;
    CUDA_TRY(cudaMallocAsync(&d, sizeof(u64) * ctx->N, ctx->stream));
    CUDA_TRY(cudaMemcpyAsync(d, p1, sizeof(u64) * ctx->N, cudaMemcpyHostToDevice, ctx->stream));
    Span sp{1, 1, 1, limb};
    int rc = launch_ntt(ctx, d, d, sp, inverse != 0, flags);
    if (rc == 0) {
        cudaError_t e = cudaMemcpyAsync(p2, d, sizeof(u64) * ctx->N, cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) rc = b200_cuda_fail(e, "b200_subring_ntt_host download");
    }
    cudaFreeAsync(d, ctx->stream);
    return rc;
}
