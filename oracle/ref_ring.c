/*
 * oracle/ref_ring.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement (plain C, unsigned __int128 == Go math/bits.Mul64) of the
 * hot path of swanhong/matrix-fhe-lattigo, package ring/.  It is the checker
 * for the CUDA library (tests/, __graft_entry__.smoke(), bench.py cpu_baseline
 * and --impl reference).  Nothing under matrix-fhe-lattigo_b200/ may link,
 * import or call it.
 *
 * Parity pin: the power-of-two path (tables + NTT + INTT) reproduces the six
 * known-answer vectors of ring/ntt_test.go:10-89 bit-for-bit (checked by
 * tests/test_oracle_kat.py against tests/golden/ntt_kat.json).  The Go
 * toolchain is absent from the build container, so the reference itself is
 * not executed; the 3N path has NO golden vectors upstream and
 * is pinned by definition (Horner evaluation at w^E[k], ring/ntt_3n.go:99-108)
 * and against vectors produced by the reference's own Python prototype
 * (references/integer_dft.py, tests/golden/ntt3n_proto.json).
 *
 * Each function cites the reference file:line it restates.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <unistd.h>

typedef uint64_t u64;
typedef unsigned __int128 u128;

#define EXPORT __attribute__((visibility("default")))

static inline u64 mulhi(u64 a, u64 b) { return (u64)(((u128)a * b) >> 64); }

/* tiny pthread parallel-for (dynamic chunks of 1) so the oracle has no OpenMP dependency */
typedef void (*pf_body)(int idx, void *arg);
typedef struct { pf_body body; void *arg; int n; volatile int next; } pf_state;
static void *pf_worker(void *p) {
    pf_state *st = p;
    for (;;) { int i = __sync_fetch_and_add(&st->next, 1); if (i >= st->n) break; st->body(i, st->arg); }
    return NULL;
}
static int host_threads(void) { long n = sysconf(_SC_NPROCESSORS_ONLN); return n < 1 ? 1 : (int)n; }
static void parallel_for(int n, int threads, pf_body body, void *arg) {
    if (threads <= 0) threads = host_threads();
    if (threads > n) threads = n;
    pf_state st = {body, arg, n, 0};
    if (threads <= 1) { pf_worker(&st); return; }
    pthread_t *th = malloc(sizeof(pthread_t) * threads);
    int started = 0;
    for (int t = 0; t < threads - 1; t++) if (pthread_create(&th[started], NULL, pf_worker, &st) == 0) started++;
    pf_worker(&st);
    for (int t = 0; t < started; t++) pthread_join(th[t], NULL);
    free(th);
}

/* ---------------------------------------------------------------------------
 * ring/modular_reduction.go
 * ------------------------------------------------------------------------- */

/* GenMRedConstant, modular_reduction.go:68-75 : q^(2^63-1) = q^-1 mod 2^64 */
EXPORT u64 ref_gen_mred_constant(u64 q) {
    u64 m = 1;
    for (int i = 0; i < 63; i++) { m *= q; q *= q; }
    return m;
}

/* GenBRedConstant, modular_reduction.go:99-107 : floor(2^128/q) as {hi, lo} */
EXPORT void ref_gen_bred_constant(u64 q, u64 out[2]) {
    /* 2^128 / q = ((2^128-1) / q) unless q | 2^128 (q odd > 1 -> never) */
    u128 all = ~(u128)0;
    u128 r = all / q;
    if ((all % q) == (u128)(q - 1)) r += 1; /* exact division guard (q=1 only) */
    out[0] = (u64)(r >> 64);
    out[1] = (u64)r;
}

/* MForm, modular_reduction.go:11-35 */
EXPORT u64 ref_mform(u64 a, u64 q, const u64 brc[2]) {
    u64 mhi = mulhi(a, brc[1]);
    u64 r = -(a * brc[0] + mhi) * q;
    if (r >= q) r -= q;
    return r;
}
/* MFormLazy, modular_reduction.go:40-45 */
EXPORT u64 ref_mform_lazy(u64 a, u64 q, const u64 brc[2]) {
    u64 mhi = mulhi(a, brc[1]);
    return -(a * brc[0] + mhi) * q;
}
/* IMForm, modular_reduction.go:49-56 */
EXPORT u64 ref_imform(u64 a, u64 q, u64 qinv) {
    u64 r = mulhi(a * qinv, q);
    r = q - r;
    if (r >= q) r -= q;
    return r;
}
/* IMFormLazy, modular_reduction.go:61-65 */
EXPORT u64 ref_imform_lazy(u64 a, u64 q, u64 qinv) {
    return q - mulhi(a * qinv, q);
}
/* MRed, modular_reduction.go:78-86 */
EXPORT u64 ref_mred(u64 x, u64 y, u64 q, u64 qinv) {
    u128 m = (u128)x * y;
    u64 hhi = mulhi((u64)m * qinv, q);
    u64 r = (u64)(m >> 64) - hhi + q;
    if (r >= q) r -= q;
    return r;
}
/* MRedLazy, modular_reduction.go:90-95 */
EXPORT u64 ref_mred_lazy(u64 x, u64 y, u64 q, u64 qinv) {
    u128 m = (u128)x * y;
    u64 H = mulhi((u64)m * qinv, q);
    return (u64)(m >> 64) - H + q;
}
/* BRedAdd, modular_reduction.go:110-117 */
EXPORT u64 ref_bred_add(u64 a, u64 q, const u64 brc[2]) {
    u64 r = a - mulhi(a, brc[0]) * q;
    if (r >= q) r -= q;
    return r;
}
/* BRedAddLazy, modular_reduction.go:121-124 */
EXPORT u64 ref_bred_add_lazy(u64 a, u64 q, const u64 brc[2]) {
    return a - mulhi(a, brc[0]) * q;
}
/* BRedLazy, modular_reduction.go:166-197 (same carry chain as BRed) */
EXPORT u64 ref_bred_lazy(u64 x, u64 y, u64 q, const u64 brc[2]) {
    u128 m = (u128)x * y;
    u64 mhi = (u64)(m >> 64), mlo = (u64)m;
    u64 r = mhi * brc[0];
    u128 h = (u128)mlo * brc[0];
    u64 hhi = (u64)(h >> 64), hlo = (u64)h;
    r += hhi;
    u64 lhi = mulhi(mlo, brc[1]);
    u64 s0 = hlo + lhi;
    u64 carry = s0 < hlo;
    r += carry;
    h = (u128)mhi * brc[1];
    hhi = (u64)(h >> 64); hlo = (u64)h;
    r += hhi;
    u64 s1 = hlo + s0;
    carry = s1 < hlo;
    r += carry;
    return mlo - r * q;
}
/* BRed, modular_reduction.go:127-162 */
EXPORT u64 ref_bred(u64 x, u64 y, u64 q, const u64 brc[2]) {
    u64 r = ref_bred_lazy(x, y, q, brc);
    if (r >= q) r -= q;
    return r;
}
/* CRed, modular_reduction.go:200-205 */
static inline u64 cred(u64 a, u64 q) { return a >= q ? a - q : a; }
EXPORT u64 ref_cred(u64 a, u64 q) { return cred(a, q); }

/* ModExp, ring/utils.go:29-39 (BRed square-and-multiply) */
EXPORT u64 ref_modexp(u64 x, u64 e, u64 p) {
    u64 brc[2];
    ref_gen_bred_constant(p, brc);
    u64 result = 1;
    for (u64 i = e; i > 0; i >>= 1) {
        if (i & 1) result = ref_bred(result, x, p, brc);
        x = ref_bred(x, x, p, brc);
    }
    return result;
}

static inline u64 mulmod(u64 a, u64 b, u64 q) { return (u64)(((u128)a * b) % q); }

/* ---------------------------------------------------------------------------
 * primes / roots: ring/primes.go:11-13 (IsPrime), subring.go:218-251
 * (PrimitiveRoot), subring.go:255-290 (Find3NPrimitiveRoot),
 * primes_3n.go:11-43 (Find3NRNSPrimes), utils/factorization (unique factors)
 * ------------------------------------------------------------------------- */
static u64 powmod(u64 b, u64 e, u64 m) {
    u64 r = 1; b %= m;
    while (e) { if (e & 1) r = mulmod(r, b, m); b = mulmod(b, b, m); e >>= 1; }
    return r;
}
/* deterministic Miller-Rabin for 64-bit (equivalent verdict to Baillie-PSW < 2^64) */
EXPORT int ref_is_prime(u64 n) {
    if (n < 2) return 0;
    static const u64 small[] = {2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37};
    for (int i = 0; i < 12; i++) { if (n % small[i] == 0) return n == small[i]; }
    u64 d = n - 1; int s = 0;
    while (!(d & 1)) { d >>= 1; s++; }
    for (int i = 0; i < 12; i++) {
        u64 x = powmod(small[i], d, n);
        if (x == 1 || x == n - 1) continue;
        int comp = 1;
        for (int r = 1; r < s; r++) { x = mulmod(x, x, n); if (x == n - 1) { comp = 0; break; } }
        if (comp) return 0;
    }
    return 1;
}
static u64 gcd64(u64 a, u64 b) { while (b) { u64 t = a % b; a = b; b = t; } return a; }
static u64 pollard_rho(u64 n) {
    if (!(n & 1)) return 2;
    for (u64 c = 1;; c++) {
        u64 x = 2, y = 2, d = 1;
        while (d == 1) {
            x = (mulmod(x, x, n) + c) % n;
            y = (mulmod(y, y, n) + c) % n;
            y = (mulmod(y, y, n) + c) % n;
            d = gcd64(x > y ? x - y : y - x, n);
        }
        if (d != n) return d;
    }
}
static void factor_rec(u64 n, u64 *f, int *nf) {
    if (n == 1) return;
    if (ref_is_prime(n)) {
        for (int i = 0; i < *nf; i++) if (f[i] == n) return;
        f[(*nf)++] = n; return;
    }
    u64 d = pollard_rho(n);
    factor_rec(d, f, nf); factor_rec(n / d, f, nf);
}
/* unique prime factors of m, ascending; returns count (<= 16) */
EXPORT int ref_unique_factors(u64 m, u64 *out) {
    int nf = 0;
    for (u64 p = 2; p < 4096 && p * p <= m; p++) {
        if (m % p == 0) { out[nf++] = p; while (m % p == 0) m /= p; }
    }
    factor_rec(m, out, &nf);
    for (int i = 1; i < nf; i++) { u64 v = out[i]; int j = i - 1; while (j >= 0 && out[j] > v) { out[j + 1] = out[j]; j--; } out[j + 1] = v; }
    return nf;
}
/* PrimitiveRoot, subring.go:218-251 : smallest g >= 3 (g=2 is never tried) */
EXPORT u64 ref_primitive_root(u64 q) {
    u64 f[16]; int nf = ref_unique_factors(q - 1, f);
    for (u64 g = 3;; g++) {
        int ok = 1;
        for (int i = 0; i < nf; i++) if (ref_modexp(g, (q - 1) / f[i], q) == 1) { ok = 0; break; }
        if (ok) return g;
    }
}
/* Find3NPrimitiveRoot, subring.go:255-290 : deterministic omega = g^((p-1)/3N) */
EXPORT u64 ref_find_3n_primitive_root(u64 p, u64 nthroot) {
    if ((p - 1) % nthroot) return 0;
    return ref_modexp(ref_primitive_root(p), (p - 1) / nthroot, p);
}
/* Find3NRNSPrimes, primes_3n.go:11-43 */
EXPORT int ref_find_3n_rns_primes(u64 N, int bitsize, int count, int budget, u64 *out) {
    u64 threeN = 3 * N, base = (u64)1 << bitsize;
    u64 cand = ((base - 1) / threeN + 1) * threeN + 1;
    int n = 0;
    for (int steps = 0; steps < budget && n < count; steps++) {
        if (cand % threeN == 1 && ref_is_prime(cand)) out[n++] = cand;
        if (cand > ~(u64)0 - threeN) break;
        cand += threeN;
    }
    return n;
}

/* ---------------------------------------------------------------------------
 * power-of-two ring: tables (subring.go:129-214) and transforms (ntt.go)
 * ------------------------------------------------------------------------- */
typedef struct {
    int N, logN;
    u64 q, qinv, brc[2], g, ninv;
    u64 *rf, *rb; /* RootsForward / RootsBackward, Montgomery form, bit-reversed */
} ref_pow2;

static u64 bitrev(u64 x, int bits) { /* utils/utils.go:34-36 */
    u64 r = 0;
    for (int i = 0; i < bits; i++) { r = (r << 1) | ((x >> i) & 1); }
    return r;
}

/* NewSubRingWithCustomNTT(subring.go:74-111) + generateNTTConstants(:129-214), NthRoot = 2N */
EXPORT ref_pow2 *ref_pow2_new(int N, u64 q) {
    if (N < 2 || (N & (N - 1))) return NULL;
    u64 nth = 2 * (u64)N;
    if (!ref_is_prime(q) || q % nth != 1) return NULL;
    ref_pow2 *s = calloc(1, sizeof *s);
    s->N = N; s->q = q;
    while ((1 << s->logN) < N) s->logN++;
    ref_gen_bred_constant(q, s->brc);
    s->qinv = ref_gen_mred_constant(q);
    s->g = ref_primitive_root(q);
    s->ninv = ref_mform(ref_modexp(nth >> 1, q - 2, q), q, s->brc);                  /* :183 */
    u64 psi = ref_mform(ref_modexp(s->g, (q - 1) / nth, q), q, s->brc);              /* :187 */
    u64 psiinv = ref_mform(ref_modexp(s->g, q - ((q - 1) / nth) - 1, q), q, s->brc); /* :188 */
    s->rf = malloc(sizeof(u64) * N); s->rb = malloc(sizeof(u64) * N);
    s->rf[0] = s->rb[0] = ref_mform(1, q, s->brc);
    for (u64 j = 1; j < (u64)N; j++) {                                               /* :198-205 */
        u64 prev = bitrev(j - 1, s->logN), cur = bitrev(j, s->logN);
        s->rf[cur] = ref_mred(s->rf[prev], psi, q, s->qinv);
        s->rb[cur] = ref_mred(s->rb[prev], psiinv, q, s->qinv);
    }
    return s;
}
EXPORT void ref_pow2_free(ref_pow2 *s) { if (s) { free(s->rf); free(s->rb); free(s); } }
EXPORT void ref_pow2_constants(const ref_pow2 *s, u64 out[6]) {
    out[0] = s->qinv; out[1] = s->brc[0]; out[2] = s->brc[1]; out[3] = s->g; out[4] = s->ninv; out[5] = (u64)s->logN;
}
EXPORT void ref_pow2_tables(const ref_pow2 *s, u64 *rf, u64 *rb) {
    memcpy(rf, s->rf, sizeof(u64) * s->N); memcpy(rb, s->rb, sizeof(u64) * s->N);
}

/* butterfly, ntt.go:155-161 */
static inline void bfly(u64 *X, u64 *Y, u64 psi, u64 twoq, u64 fourq, u64 q, u64 qinv) {
    u64 U = *X, V = *Y;
    if (U >= fourq) U -= fourq;
    V = ref_mred_lazy(V, psi, q, qinv);
    *X = U + V; *Y = U + twoq - V;
}
/* invbutterfly, ntt.go:164-171 */
static inline void ibfly(u64 *X, u64 *Y, u64 psi, u64 twoq, u64 fourq, u64 q, u64 qinv) {
    u64 U = *X, V = *Y;
    u64 x = U + V;
    if (x >= twoq) x -= twoq;
    *X = x; *Y = ref_mred_lazy(U + fourq - V, psi, q, qinv);
}
/* nttLazy, ntt.go:223-257 (the readable spec; the unrolled version differs only in which
 * stages apply the 4q correction, so canonical outputs are identical) */
static void ntt_lazy(const ref_pow2 *s, const u64 *p1, u64 *p2) {
    int N = s->N; u64 q = s->q, twoq = 2 * q, fourq = 4 * q;
    if (p1 != p2) memcpy(p2, p1, sizeof(u64) * N);
    int t = N;
    for (int m = 1; m < N; m <<= 1) {
        t >>= 1;
        for (int i = 0; i < m; i++) {
            int j1 = (i * t) << 1; u64 F = s->rf[m + i];
            for (int j = j1; j < j1 + t; j++) bfly(&p2[j], &p2[j + t], F, twoq, fourq, q, s->qinv);
        }
    }
}
/* inttLazy, ntt.go:568-606 */
static void intt_lazy(const ref_pow2 *s, const u64 *p1, u64 *p2) {
    int N = s->N; u64 q = s->q, twoq = 2 * q, fourq = 4 * q;
    if (p1 != p2) memcpy(p2, p1, sizeof(u64) * N);
    int t = 1;
    for (int m = N; m > 1; m >>= 1) {
        int h = m >> 1;
        for (int i = 0, j1 = 0; i < h; i++, j1 += 2 * t) {
            u64 F = s->rb[h + i];
            for (int j = j1; j < j1 + t; j++) ibfly(&p2[j], &p2[j + t], F, twoq, fourq, q, s->qinv);
        }
        t <<= 1;
    }
}
/* NTTStandard / NTTStandardLazy, ntt.go:174-182 */
EXPORT void ref_pow2_ntt(const ref_pow2 *s, const u64 *p1, u64 *p2, int lazy) {
    ntt_lazy(s, p1, p2);
    if (!lazy) for (int i = 0; i < s->N; i++) p2[i] = ref_bred_add(p2[i], s->q, s->brc); /* reducevec */
}
/* INTTStandard / INTTStandardLazy, ntt.go:185-206.  (The Go "Lazy" variant calls the
 * non-lazy mulscalarmontgomeryvec for N>=16, ntt.go:204 -- restated as such.) */
EXPORT void ref_pow2_intt(const ref_pow2 *s, const u64 *p1, u64 *p2, int lazy) {
    intt_lazy(s, p1, p2);
    for (int i = 0; i < s->N; i++)
        p2[i] = (lazy && s->N < 16) ? ref_mred_lazy(p2[i], s->ninv, s->q, s->qinv)
                                    : ref_mred(p2[i], s->ninv, s->q, s->qinv);
}

/* ---------------------------------------------------------------------------
 * conjugate-invariant ring Z_q[X + X^-1]/(X^2N + 1): ring/ntt.go:80-124 (transformer), :716-786 (forward),
 * :1088-1158 (inverse); constructor NewRingConjugateInvariant (ring/ring.go:280-282) = NthRoot 4N, so the tables
 * of generateNTTConstants (subring.go:183-205) hold 2N entries, bit-reversed over log2(2N) bits, and
 * NInv = (2N)^-1 in Montgomery form.
 * ------------------------------------------------------------------------- */
typedef struct {
    int N, logN;      /* ring degree N (the compressed representation), log2 N */
    u64 q, qinv, brc[2], g, ninv;
    u64 *rf, *rb;     /* 2N entries each */
} ref_ci;

EXPORT ref_ci *ref_ci_new(int N, u64 q) {
    if (N < 2 || (N & (N - 1))) return NULL;
    u64 nth = 4 * (u64)N;
    if (!ref_is_prime(q) || q % nth != 1) return NULL;
    ref_ci *s = calloc(1, sizeof *s);
    s->N = N; s->q = q;
    while ((1 << s->logN) < N) s->logN++;
    ref_gen_bred_constant(q, s->brc);
    s->qinv = ref_gen_mred_constant(q);
    s->g = ref_primitive_root(q);
    s->ninv = ref_mform(ref_modexp(nth >> 1, q - 2, q), q, s->brc);                  /* subring.go:183 */
    u64 psi = ref_mform(ref_modexp(s->g, (q - 1) / nth, q), q, s->brc);              /* :187 */
    u64 psiinv = ref_mform(ref_modexp(s->g, q - ((q - 1) / nth) - 1, q), q, s->brc); /* :188 */
    u64 half = nth >> 1;
    s->rf = malloc(sizeof(u64) * half); s->rb = malloc(sizeof(u64) * half);
    s->rf[0] = s->rb[0] = ref_mform(1, q, s->brc);
    for (u64 j = 1; j < half; j++) {                                                 /* :198-205 */
        u64 prev = bitrev(j - 1, s->logN + 1), cur = bitrev(j, s->logN + 1);
        s->rf[cur] = ref_mred(s->rf[prev], psi, q, s->qinv);
        s->rb[cur] = ref_mred(s->rb[prev], psiinv, q, s->qinv);
    }
    return s;
}
EXPORT void ref_ci_free(ref_ci *s) { if (s) { free(s->rf); free(s->rb); free(s); } }
EXPORT void ref_ci_constants(const ref_ci *s, u64 out[6]) {
    out[0] = s->qinv; out[1] = s->brc[0]; out[2] = s->brc[1]; out[3] = s->g; out[4] = s->ninv; out[5] = (u64)s->logN;
}
EXPORT void ref_ci_tables(const ref_ci *s, u64 *rf, u64 *rb) {
    memcpy(rf, s->rf, sizeof(u64) * 2 * s->N); memcpy(rb, s->rb, sizeof(u64) * 2 * s->N);
}
/* nttConjugateInvariantLazy, ntt.go:754-786 (p1 and p2 may alias: the mirrored pair is read before it is written) */
static void ci_ntt_lazy(const ref_ci *s, const u64 *p1, u64 *p2) {
    int N = s->N; u64 q = s->q, twoq = 2 * q, fourq = 4 * q;
    u64 F = s->rf[1];
    for (int jx = 1, jy = N - 1; jx < (N >> 1); jx++, jy--) {
        u64 a = p1[jx], b = p1[jy];
        p2[jx] = a + twoq - ref_mred_lazy(b, F, q, s->qinv);
        p2[jy] = b + twoq - ref_mred_lazy(a, F, q, s->qinv);
    }
    p2[N >> 1] = p1[N >> 1] + twoq - ref_mred_lazy(p1[N >> 1], F, q, s->qinv);
    p2[0] = p1[0];
    int t = N;
    for (int m = 2; m < 2 * N; m <<= 1) {
        t >>= 1;
        int h = m >> 1;
        for (int i = 0, j1 = 0; i < h; i++, j1 += 2 * t) {
            u64 Fm = s->rf[m + i];
            for (int j = j1; j < j1 + t; j++) bfly(&p2[j], &p2[j + t], Fm, twoq, fourq, q, s->qinv);
        }
    }
}
/* inttConjugateInvariantLazy, ntt.go:1104-1158 */
static void ci_intt_lazy(const ref_ci *s, const u64 *p1, u64 *p2) {
    int N = s->N; u64 q = s->q, twoq = 2 * q, fourq = 4 * q;
    if (p1 != p2) memcpy(p2, p1, sizeof(u64) * N);
    int t = 1, h = N >> 1;
    for (int i = 0, j1 = 0; i < h; i++, j1 += 2 * t) {
        u64 F = s->rb[N + i];
        for (int j = j1; j < j1 + t; j++) ibfly(&p2[j], &p2[j + t], F, twoq, fourq, q, s->qinv);
    }
    t <<= 1;
    for (int m = N >> 1; m > 1; m >>= 1) {
        h = m >> 1;
        for (int i = 0, j1 = 0; i < h; i++, j1 += 2 * t) {
            u64 F = s->rb[m + i];
            for (int j = j1; j < j1 + t; j++) ibfly(&p2[j], &p2[j + t], F, twoq, fourq, q, s->qinv);
        }
        t <<= 1;
    }
    u64 F = s->rb[1];
    for (int jx = 1, jy = N - 1; jx < (N >> 1); jx++, jy--) {
        u64 a = p2[jx], b = p2[jy];
        p2[jx] = a + twoq - ref_mred_lazy(b, F, q, s->qinv);
        p2[jy] = b + twoq - ref_mred_lazy(a, F, q, s->qinv);
    }
    p2[N >> 1] = p2[N >> 1] + twoq - ref_mred_lazy(p2[N >> 1], F, q, s->qinv);
    p2[0] = cred(p2[0] << 1, q);
}
/* NTTConjugateInvariant / ...Lazy, ntt.go:716-726 */
EXPORT void ref_ci_ntt(const ref_ci *s, const u64 *p1, u64 *p2, int lazy) {
    ci_ntt_lazy(s, p1, p2);
    if (!lazy) for (int i = 0; i < s->N; i++) p2[i] = ref_bred_add(p2[i], s->q, s->brc);
}
/* INTTConjugateInvariant / ...Lazy, ntt.go:728-738 */
EXPORT void ref_ci_intt(const ref_ci *s, const u64 *p1, u64 *p2, int lazy) {
    ci_intt_lazy(s, p1, p2);
    for (int i = 0; i < s->N; i++)
        p2[i] = lazy ? ref_mred_lazy(p2[i], s->ninv, s->q, s->qinv) : ref_mred(p2[i], s->ninv, s->q, s->qinv);
}

/* ---------------------------------------------------------------------------
 * automorphisms X -> X^gen of power-of-two rings: ring/automorphism.go
 * ------------------------------------------------------------------------- */
/* AutomorphismNTTIndex, automorphism.go:12-35.  Returns 0, or -1 if N or NthRoot is not a power of two. */
EXPORT int ref_automorphism_ntt_index(int N, u64 nthroot, u64 galel, u64 *index) {
    if (N <= 0 || (N & (N - 1)) || nthroot == 0 || (nthroot & (nthroot - 1))) return -1;
    int lognth = 0;
    while (((u64)1 << (lognth + 1)) < nthroot) lognth++;          /* bits.Len64(NthRoot-1) - 1 */
    u64 mask = nthroot - 1;
    for (int i = 0; i < N; i++) {
        u64 tmp1 = 2 * bitrev((u64)i, lognth) + 1;
        u64 tmp2 = (((galel * tmp1) & mask) - 1) >> 1;
        index[i] = bitrev(tmp2, lognth);
    }
    return 0;
}
/* Ring.Automorphism on one limb, automorphism.go:113-176 (ci != 0: the conjugate-invariant branch :124-151).  A negated
 * zero coefficient is written as q, exactly like the Go expression. */
EXPORT void ref_automorphism(int N, u64 q, int ci, u64 gen, const u64 *in, u64 *out) {
    u64 n = (u64)N;
    if (ci) {
        u64 mask = 2 * n - 1;
        int logn = 0; while (((u64)1 << logn) <= mask) logn++;    /* bits.Len64(mask) */
        for (u64 i = 0; i < 2 * n; i++) {
            u64 raw = i * gen, index = raw & mask, tmp = (raw >> logn) & 1;
            if (index < n) {
                u64 idx = i;
                if (idx >= n) { idx = 2 * n - idx; tmp ^= 1; }
                out[index] = (in[idx] * (tmp ^ 1)) | ((q - in[idx]) * tmp);
            }
        }
    } else {
        u64 mask = n - 1;
        int logn = 0; while (((u64)1 << logn) <= mask) logn++;
        for (u64 i = 0; i < n; i++) {
            u64 raw = i * gen, index = raw & mask, tmp = (raw >> logn) & 1;
            out[index] = (in[i] * (tmp ^ 1)) | ((q - in[i]) * tmp);
        }
    }
}

/* ---------------------------------------------------------------------------
 * element-wise family: ring/vec_ops.go (op codes shared with include/b200ring.h)
 * ------------------------------------------------------------------------- */
enum {
    OP_ADD = 0, OP_ADD_LAZY, OP_SUB, OP_SUB_LAZY, OP_NEG, OP_REDUCE, OP_REDUCE_LAZY,
    OP_MFORM, OP_MFORM_LAZY, OP_IMFORM, OP_ZERO,
    OP_MUL_BARRETT = 16, OP_MUL_BARRETT_LAZY, OP_MUL_BARRETT_ADD, OP_MUL_BARRETT_ADD_LAZY, OP_MUL_LAZY, OP_MUL_LAZY_ADD_LAZY,
    OP_MUL_MONT = 32, OP_MUL_MONT_LAZY, OP_MUL_MONT_ADD, OP_MUL_MONT_ADD_LAZY, OP_MUL_MONT_LAZY_ADD_LAZY,
    OP_MUL_MONT_SUB, OP_MUL_MONT_SUB_LAZY, OP_MUL_MONT_LAZY_SUB_LAZY, OP_MUL_MONT_LAZY_NEG,
    OP_MULSCALAR_MONT = 48, OP_MULSCALAR_MONT_LAZY, OP_MULSCALAR_MONT_ADD, OP_ADD_SCALAR, OP_SUB_SCALAR,
    OP_SUB_MULSCALAR_MONT_2Q, OP_ADD_LAZY_MULSCALAR_MONT, OP_ADD_SCALAR_LAZY, OP_ADD_SCALAR_LAZY_NEG_2Q_LAZY,
    OP_ADD_SCALAR_LAZY_MULSCALAR_MONT, OP_MULSCALAR_MONT_ADD_SCALAR, OP_MASK,
    OP_ADD_SCALAR_DOUBLE, OP_SUB_SCALAR_DOUBLE, OP_MULSCALAR_MONT_DOUBLE, OP_MULSCALAR_MONT_ADD_DOUBLE
};
/* p3 = op(p1, p2 [, p3]) on one limb; scalar (scalar0) and scalar1 used by the *_SCALAR ops: two-scalar operators take
 * (scalar0, scalarMont1) in the Go argument order, MaskVec (w, mask), the Double operators of ring/operations.go:165-266
 * scalar0 for words [0, N/2) and scalar1 for [N/2, N).  Returns 0, or -1 on unknown op. */
EXPORT int ref_vec_op2(int op, int N, u64 q, const u64 *p1, const u64 *p2, u64 *p3, u64 scalar, u64 scalar1);
EXPORT int ref_vec_op(int op, int N, u64 q, const u64 *p1, const u64 *p2, u64 *p3, u64 scalar) {
    return ref_vec_op2(op, N, q, p1, p2, p3, scalar, 0);
}
EXPORT int ref_vec_op2(int op, int N, u64 q, const u64 *p1, const u64 *p2, u64 *p3, u64 scalar, u64 scalar1) {
    if (op >= OP_ADD_SCALAR_DOUBLE && op <= OP_MULSCALAR_MONT_ADD_DOUBLE) {   /* two half-limb calls of the plain operator */
        static const int plain[4] = {OP_ADD_SCALAR, OP_SUB_SCALAR, OP_MULSCALAR_MONT, OP_MULSCALAR_MONT_ADD};
        const int h = N >> 1, o = plain[op - OP_ADD_SCALAR_DOUBLE];
        if (ref_vec_op2(o, h, q, p1, NULL, p3, scalar, 0)) return -1;
        return ref_vec_op2(o, N - h, q, p1 + h, NULL, p3 + h, scalar1, 0);
    }
    u64 brc[2]; ref_gen_bred_constant(q, brc);
    u64 qinv = ref_gen_mred_constant(q), twoq = 2 * q;
    for (int i = 0; i < N; i++) {
        u64 x = p1 ? p1[i] : 0, y = p2 ? p2[i] : 0, z = p3[i];
        switch (op) {
        case OP_ADD: z = cred(x + y, q); break;                                  /* vec_ops.go:7-29 */
        case OP_ADD_LAZY: z = x + y; break;                                      /* :31-53 */
        case OP_SUB: z = cred((x + q) - y, q); break;                            /* :55-77 */
        case OP_SUB_LAZY: z = x + q - y; break;                                  /* :79-101 */
        case OP_NEG: z = q - x; break;                                           /* :103-123 */
        case OP_REDUCE: z = ref_bred_add(x, q, brc); break;                      /* :125-145 */
        case OP_REDUCE_LAZY: z = ref_bred_add_lazy(x, q, brc); break;            /* :147-167 */
        case OP_MFORM: z = ref_mform(x, q, brc); break;                          /* :778-798 */
        case OP_MFORM_LAZY: z = ref_mform_lazy(x, q, brc); break;                /* :800-820 */
        case OP_IMFORM: z = ref_imform(x, q, qinv); break;                       /* :822-845 */
        case OP_MUL_BARRETT: z = ref_bred(x, y, q, brc); break;                  /* :217-239 */
        case OP_MUL_BARRETT_LAZY: z = ref_bred_lazy(x, y, q, brc); break;        /* :241-263 */
        case OP_MUL_BARRETT_ADD: z = cred(z + ref_bred(x, y, q, brc), q); break; /* :265-287 */
        case OP_MUL_BARRETT_ADD_LAZY: z += ref_bred(x, y, q, brc); break;        /* :289-311 */
        case OP_MUL_MONT: z = ref_mred(x, y, q, qinv); break;                    /* :313-334 */
        case OP_MUL_MONT_LAZY: z = ref_mred_lazy(x, y, q, qinv); break;          /* :336-358 */
        case OP_MUL_MONT_ADD: z = cred(z + ref_mred(x, y, q, qinv), q); break;   /* :360-381 */
        case OP_MUL_MONT_ADD_LAZY: z += ref_mred(x, y, q, qinv); break;          /* :383-405 */
        case OP_MUL_MONT_LAZY_ADD_LAZY: z += ref_mred_lazy(x, y, q, qinv); break;/* :407-429 */
        case OP_MUL_MONT_SUB: z = cred(z + (q - ref_mred(x, y, q, qinv)), q); break;      /* :431-453 */
        case OP_MUL_MONT_SUB_LAZY: z += (q - ref_mred(x, y, q, qinv)); break;             /* :455-477 */
        case OP_MUL_MONT_LAZY_SUB_LAZY: z += twoq - ref_mred_lazy(x, y, q, qinv); break;  /* :479-502 */
        case OP_MUL_MONT_LAZY_NEG: z = twoq - ref_mred_lazy(x, y, q, qinv); break;        /* :504-527 */
        case OP_MULSCALAR_MONT: z = ref_mred(x, scalar, q, qinv); break;                  /* :664-684 */
        case OP_MULSCALAR_MONT_LAZY: z = ref_mred_lazy(x, scalar, q, qinv); break;        /* :686-706 */
        case OP_MULSCALAR_MONT_ADD: z = cred(z + ref_mred(x, scalar, q, qinv), q); break; /* :708-728 */
        case OP_ADD_SCALAR: z = cred(x + scalar, q); break;                               /* :575-595 */
        case OP_SUB_SCALAR: z = cred(x + q - scalar, q); break;                           /* :642-662 */
        case OP_SUB_MULSCALAR_MONT_2Q: z = ref_mred(twoq - y + x, scalar, q, qinv); break; /* :752-776 */
        case OP_ZERO: z = 0; break;                                                       /* :847-868 */
        case OP_MUL_LAZY: z = x * y; break;                                               /* :169-191 */
        case OP_MUL_LAZY_ADD_LAZY: z += x * y; break;                                     /* :193-215 */
        case OP_ADD_LAZY_MULSCALAR_MONT: z = ref_mred(x + y, scalar, q, qinv); break;     /* :529-551 */
        case OP_ADD_SCALAR_LAZY: z = x + scalar; break;                                   /* :597-617 */
        case OP_ADD_SCALAR_LAZY_NEG_2Q_LAZY: z = scalar + twoq - x; break;                /* :619-640 */
        case OP_ADD_SCALAR_LAZY_MULSCALAR_MONT: z = ref_mred(x + scalar, scalar1, q, qinv); break;         /* :553-573 */
        case OP_MULSCALAR_MONT_ADD_SCALAR: z = cred(ref_mred(x, scalar1, q, qinv) + scalar, q); break;     /* :730-750 */
        case OP_MASK: z = scalar >= 64 ? 0 : (x >> scalar) & scalar1; break;              /* :870-890; Go: x >> w is 0 for w >= 64 */
        default: return -1;
        }
        p3[i] = z;
    }
    return 0;
}

/* ---------------------------------------------------------------------------
 * 3N ring Z_q[X]/(X^N - X^(N/2) + 1), N = 2^a 3^b
 * ------------------------------------------------------------------------- */
static int gcd_int(int a, int b) { while (b) { int t = a % b; a = b; b = t; } return a; }

/* buildPrimitive3NExponents, ntt_3n.go:235-243 : ascending totatives of 3N */
EXPORT int ref_3n_exponents(int threeN, int *out) {
    int n = 0;
    for (int e = 1; e < threeN; e++) if (gcd_int(e, threeN) == 1) out[n++] = e;
    return n;
}

/* T0 -- definitional forward: NumberTheoreticTransformer3N.Forward, ntt_3n.go:82-109.
 * out[k] = sum_j p[j] (omega^E[k])^j by Horner.  O(N^2).  In-place safe (temp copy, :90-96). */
/* single output slot k of the definitional forward (spot checks at large N) */
EXPORT u64 ref_3n_forward_horner_at(int N, u64 q, u64 omega, const u64 *p1, int exponent) {
    u64 xk = ref_modexp(omega, (u64)exponent, q), acc = 0;
    for (int j = N - 1; j >= 0; j--) { acc = mulmod(acc, xk, q); acc = cred(acc + p1[j], q); }
    return acc;
}
typedef struct { int N; u64 q, omega; const u64 *p1; u64 *tmp; const int *E; } horner_job;
static void horner_body(int k, void *arg) {
    horner_job *j = arg;
    j->tmp[k] = ref_3n_forward_horner_at(j->N, j->q, j->omega, j->p1, j->E[k]);
}
EXPORT void ref_3n_forward_horner(int N, u64 q, u64 omega, const u64 *p1, u64 *p2) {
    int *E = malloc(sizeof(int) * 3 * N);
    ref_3n_exponents(3 * N, E);
    u64 *tmp = malloc(sizeof(u64) * N);
    horner_job job = {N, q, omega, p1, tmp, E};
    parallel_for(N, 0, horner_body, &job);   /* threads only speed the checker up; result is exact */
    memcpy(p2, tmp, sizeof(u64) * N);
    free(tmp); free(E);
}

/* T0 -- definitional inverse: NumberTheoreticTransformer3N.Backward, ntt_3n.go:118-151 with
 * solveVandermondeGaussian :170-222.  O(N^3); only for N <= a few hundred (as upstream). */
EXPORT int ref_3n_backward_gauss(int N, u64 q, u64 omega, const u64 *p1, u64 *p2) {
    int *E = malloc(sizeof(int) * 3 * N);
    ref_3n_exponents(3 * N, E);
    u64 *V = malloc(sizeof(u64) * (size_t)N * N), *y = malloc(sizeof(u64) * N), *a = calloc(N, sizeof(u64));
    for (int i = 0; i < N; i++) {
        u64 x = ref_modexp(omega, (u64)E[i], q);
        V[(size_t)i * N] = 1;
        for (int j = 1; j < N; j++) V[(size_t)i * N + j] = mulmod(V[(size_t)i * N + j - 1], x, q);
        y[i] = p1[i];
    }
    int ok = 1;
    for (int col = 0; col < N && ok; col++) {
        int piv = col;
        while (piv < N && V[(size_t)piv * N + col] == 0) piv++;
        if (piv == N) { ok = 0; break; }
        if (piv != col) {
            for (int j = 0; j < N; j++) { u64 t = V[(size_t)col * N + j]; V[(size_t)col * N + j] = V[(size_t)piv * N + j]; V[(size_t)piv * N + j] = t; }
            u64 t = y[col]; y[col] = y[piv]; y[piv] = t;
        }
        u64 inv = ref_modexp(V[(size_t)col * N + col] % q, q - 2, q);
        for (int j = col; j < N; j++) V[(size_t)col * N + j] = mulmod(V[(size_t)col * N + j], inv, q);
        y[col] = mulmod(y[col], inv, q);
        for (int row = col + 1; row < N; row++) {
            u64 f = V[(size_t)row * N + col];
            if (!f) continue;
            for (int j = col; j < N; j++) V[(size_t)row * N + j] = cred(V[(size_t)row * N + j] + q - mulmod(f, V[(size_t)col * N + j], q), q);
            y[row] = cred(y[row] + q - mulmod(f, y[col], q), q);
        }
    }
    if (ok) for (int i = N - 1; i >= 0; i--) {
        u64 sum = 0;
        for (int j = i + 1; j < N; j++) sum = cred(sum + mulmod(V[(size_t)i * N + j], a[j], q), q);
        a[i] = cred(y[i] + q - sum, q);
    }
    for (int i = 0; i < N; i++) p2[i] = a[i] % q;
    free(V); free(y); free(a); free(E);
    return ok ? 0 : -1;
}

/* naiveCyclotomicMultiply3N, ntt_3n_test.go:312-364 : schoolbook product reduced by
 * X^N = X^(N/2) - 1 (high to low). */
EXPORT void ref_3n_schoolbook(int N, u64 q, const u64 *a, const u64 *b, u64 *c) {
    u64 *t = calloc(2 * (size_t)N, sizeof(u64));
    for (int i = 0; i < N; i++) {
        if (!a[i]) continue;
        for (int j = 0; j < N; j++) t[i + j] = cred(t[i + j] + mulmod(a[i], b[j], q), q);
    }
    for (int k = 2 * N - 2; k >= N; k--) {
        u64 v = t[k]; if (!v) continue;
        t[k] = 0;
        t[k - N / 2] = cred(t[k - N / 2] + v, q);        /* + v X^(k-N/2) */
        t[k - N] = cred(t[k - N] + q - v, q);            /* - v X^(k-N)   */
    }
    memcpy(c, t, sizeof(u64) * N);
    free(t);
}
/* negacyclic schoolbook product mod X^N + 1 (ground truth for the power-of-two multiply pipeline) */
EXPORT void ref_negacyclic_schoolbook(int N, u64 q, const u64 *a, const u64 *b, u64 *c) {
    u64 *t = calloc(2 * (size_t)N, sizeof(u64));
    for (int i = 0; i < N; i++) {
        if (!a[i]) continue;
        for (int j = 0; j < N; j++) t[i + j] = cred(t[i + j] + mulmod(a[i], b[j], q), q);
    }
    for (int k = 0; k < N; k++) c[k] = cred(t[k] + q - t[k + N], q);
    free(t);
}

/* T2 -- fast 3N transform.  Structure restated from the reference's Python prototype
 * references/integer_dft.py:150-183 (exponent tree + zetas), :266-348 (forward),
 * :350-432 (inverse); boundary semantics (slot k <-> omega^E[k]) from ntt_3n.go:46-53,99-108.
 * Twiddle multiplications use Shoup's method (w, floor(w 2^64/q)); all stored values canonical. */
typedef struct {
    int N, a, b;
    u64 q, omega, Om, Om_s;       /* Om = omega^N (cube root of unity) */
    u64 *z, *zs;                  /* zetas[0..N) and Shoup companions   */
    u64 cinv, cinv_s;             /* (zeta - zeta^5)^-1                 */
    u64 ninv, ninv_s, ninv2, ninv2_s;
    int *leaf;                    /* leaf exponent t_i (mod 3N) of fast slot i */
    int *go2fast;                 /* Go slot k -> fast slot i            */
} ref_3n;

static inline u64 shoup_of(u64 w, u64 q) { return (u64)((((u128)w) << 64) / q); }
static inline u64 mul_shoup(u64 x, u64 w, u64 ws, u64 q) {
    u64 r = x * w - mulhi(x, ws) * q;
    return r >= q ? r - q : r;
}

EXPORT ref_3n *ref_3n_new(int N, u64 q, u64 omega) {
    int a = 0, b = 0, t = N;
    while (t % 2 == 0) { t /= 2; a++; }
    while (t % 3 == 0) { t /= 3; b++; }
    if (t != 1 || a < 1 || b < 1) return NULL;
    u64 threeN = 3 * (u64)N;
    if ((q - 1) % threeN) return NULL;
    /* omega must have exact order 3N = 2^a 3^(b+1) */
    if (powmod(omega, threeN, q) != 1 || powmod(omega, threeN / 2, q) == 1 || powmod(omega, threeN / 3, q) == 1) return NULL;
    ref_3n *s = calloc(1, sizeof *s);
    s->N = N; s->a = a; s->b = b; s->q = q; s->omega = omega;
    s->Om = powmod(omega, N, q); s->Om_s = shoup_of(s->Om, q);
    int level = a + b;
    /* exponent tree, integer_dft.py:156-183 */
    u64 *cur = malloc(sizeof(u64) * N), *nxt = malloc(sizeof(u64) * N);
    s->z = malloc(sizeof(u64) * N); s->zs = malloc(sizeof(u64) * N);
    int nz = 0, width;
    s->z[nz++] = 1;
    cur[0] = threeN / 6; cur[1] = 5 * threeN / 6; width = 2;
    s->z[nz++] = powmod(omega, cur[0], q);
    for (int l = 1; l <= b; l++) {
        for (int i = 0; i < width; i++) {
            u64 e = cur[i] / 3;
            nxt[3 * i] = e; nxt[3 * i + 1] = e + threeN / 3; nxt[3 * i + 2] = e + 2 * threeN / 3;
            s->z[nz++] = powmod(omega, e, q);
            s->z[nz++] = powmod(omega, 2 * e, q);
        }
        width *= 3; u64 *sw = cur; cur = nxt; nxt = sw;
    }
    for (int l = b + 1; l < level; l++) {
        for (int i = 0; i < width; i++) {
            u64 e = cur[i] / 2;
            nxt[2 * i] = e; nxt[2 * i + 1] = e + threeN / 2;
            s->z[nz++] = powmod(omega, e, q);
        }
        width *= 2; u64 *sw = cur; cur = nxt; nxt = sw;
    }
    if (nz != N || width != N) { free(cur); free(nxt); free(s->z); free(s->zs); free(s); return NULL; }
    for (int i = 0; i < N; i++) s->zs[i] = shoup_of(s->z[i], q);
    s->leaf = malloc(sizeof(int) * N);
    for (int i = 0; i < N; i++) s->leaf[i] = (int)(cur[i] % threeN);
    free(cur); free(nxt);
    /* Go order: slot k evaluates at omega^E[k], E ascending totatives (ntt_3n.go:46-53) */
    int *E = malloc(sizeof(int) * 3 * N), *pos = malloc(sizeof(int) * 3 * N);
    int n = ref_3n_exponents(3 * N, E);
    for (int i = 0; i < 3 * N; i++) pos[i] = -1;
    for (int k = 0; k < n; k++) pos[E[k]] = k;
    s->go2fast = malloc(sizeof(int) * N);
    for (int i = 0; i < N; i++) s->go2fast[pos[s->leaf[i]]] = i;
    free(E); free(pos);
    u64 zeta = s->z[1], zeta5 = powmod(zeta, 5, q);
    s->cinv = powmod(cred(zeta + q - zeta5, q), q - 2, q); s->cinv_s = shoup_of(s->cinv, q);
    s->ninv = powmod((u64)N % q, q - 2, q); s->ninv_s = shoup_of(s->ninv, q);
    s->ninv2 = cred(2 * s->ninv, q); s->ninv2_s = shoup_of(s->ninv2, q);
    return s;
}
EXPORT void ref_3n_free(ref_3n *s) {
    if (s) { free(s->z); free(s->zs); free(s->leaf); free(s->go2fast); free(s); }
}
EXPORT void ref_3n_leaf_exponents(const ref_3n *s, int *out) { memcpy(out, s->leaf, sizeof(int) * s->N); }
EXPORT void ref_3n_go2fast(const ref_3n *s, int *out) { memcpy(out, s->go2fast, sizeof(int) * s->N); }
EXPORT void ref_3n_zetas(const ref_3n *s, u64 *out) { memcpy(out, s->z, sizeof(u64) * s->N); }

/* forward, tree order, in place; input canonical */
static void fwd_tree(const ref_3n *s, u64 *b) {
    int N = s->N; u64 q = s->q; int zi = 1;
    {   /* cyclotomic layer, integer_dft.py:288-296 */
        u64 z = s->z[zi], zs = s->zs[zi]; zi++;
        for (int i = 0; i < N / 2; i++) {
            u64 t = mul_shoup(b[i + N / 2], z, zs, q);
            u64 lo = b[i], hi = b[i + N / 2];
            b[i + N / 2] = cred(cred(lo + hi, q) + q - t, q);
            b[i] = cred(lo + t, q);
        }
    }
    int step = N / 6;
    while (step >= (1 << (s->a - 1))) {   /* radix-3 layers, :301-323 */
        for (int st = 0; st < N; st += 3 * step) {
            u64 z1 = s->z[zi], z1s = s->zs[zi], z2 = s->z[zi + 1], z2s = s->zs[zi + 1]; zi += 2;
            for (int i = st; i < st + step; i++) {
                u64 t1 = mul_shoup(b[i + step], z1, z1s, q);
                u64 t2 = mul_shoup(b[i + 2 * step], z2, z2s, q);
                u64 t3 = mul_shoup(cred(t1 + q - t2, q), s->Om, s->Om_s, q);
                u64 b0 = b[i];
                b[i + 2 * step] = cred(cred(b0 + q - t1, q) + q - t3, q);
                b[i + step] = cred(cred(b0 + q - t2, q) + t3, q);
                b[i] = cred(cred(b0 + t1, q) + t2, q);
            }
        }
        step /= 3;
    }
    step = s->a >= 2 ? 1 << (s->a - 2) : 0;
    while (step >= 1) {                    /* radix-2 layers, :325-343 */
        for (int st = 0; st < N; st += 2 * step) {
            u64 z = s->z[zi], zs = s->zs[zi]; zi++;
            for (int i = st; i < st + step; i++) {
                u64 t = mul_shoup(b[i + step], z, zs, q);
                u64 u = b[i];
                b[i + step] = cred(u + q - t, q);
                b[i] = cred(u + t, q);
            }
        }
        step >>= 1;
    }
}
/* inverse, tree order, in place */
static void inv_tree(const ref_3n *s, u64 *b) {
    int N = s->N; u64 q = s->q; int zi = N - 1;
    int step = 1;
    if (s->a >= 2) while (step <= (1 << (s->a - 2))) {   /* inverse radix-2, :371-388 */
        for (int st = 0; st < N; st += 2 * step) {
            u64 z = s->z[zi], zs = s->zs[zi]; zi--;
            for (int i = st; i < st + step; i++) {
                u64 t = b[i + step], u = b[i];
                b[i + step] = mul_shoup(cred(t + q - u, q), z, zs, q);
                b[i] = cred(t + u, q);
            }
        }
        step <<= 1;
    }
    while (step <= N / 6) {                                /* inverse radix-3, :391-414 */
        for (int st = 0; st < N; st += 3 * step) {
            u64 z2 = s->z[zi], z2s = s->zs[zi]; zi--;
            u64 z1 = s->z[zi], z1s = s->zs[zi]; zi--;
            for (int i = st; i < st + step; i++) {
                u64 b0 = b[i], b1 = b[i + step], b2 = b[i + 2 * step];
                u64 t1 = mul_shoup(cred(b1 + q - b0, q), s->Om, s->Om_s, q);
                u64 t2 = mul_shoup(cred(cred(b2 + q - b0, q) + t1, q), z1, z1s, q);
                u64 t3 = mul_shoup(cred(cred(b2 + q - b1, q) + q - t1, q), z2, z2s, q);
                b[i] = cred(cred(b0 + b1, q) + b2, q);
                b[i + step] = t2; b[i + 2 * step] = t3;
            }
        }
        step *= 3;
    }
    for (int i = 0; i < N / 2; i++) {                      /* final layer, :417-429 */
        u64 lo = b[i], hi = b[i + N / 2];
        u64 t1 = cred(lo + hi, q);
        u64 t2 = mul_shoup(cred(lo + q - hi, q), s->cinv, s->cinv_s, q);
        b[i] = mul_shoup(cred(t1 + q - t2, q), s->ninv, s->ninv_s, q);
        b[i + N / 2] = mul_shoup(t2, s->ninv2, s->ninv2_s, q);
    }
}
/* order: 0 = Go order (slot k <-> omega^E[k], the API contract), 1 = tree order (prototype's) */
EXPORT void ref_3n_forward(const ref_3n *s, const u64 *p1, u64 *p2, int order) {
    int N = s->N;
    u64 *t = malloc(sizeof(u64) * N);
    for (int i = 0; i < N; i++) t[i] = p1[i] % s->q;
    fwd_tree(s, t);
    if (order == 1) memcpy(p2, t, sizeof(u64) * N);
    else for (int k = 0; k < N; k++) p2[k] = t[s->go2fast[k]];
    free(t);
}
EXPORT void ref_3n_backward(const ref_3n *s, const u64 *p1, u64 *p2, int order) {
    int N = s->N;
    u64 *t = malloc(sizeof(u64) * N);
    if (order == 1) for (int i = 0; i < N; i++) t[i] = p1[i] % s->q;
    else for (int k = 0; k < N; k++) t[s->go2fast[k]] = p1[k] % s->q;
    inv_tree(s, t);
    memcpy(p2, t, sizeof(u64) * N);
    free(t);
}

/* ---------------------------------------------------------------------------
 * RNS rescale in the NTT domain: Ring.DivRoundByLastModulusNTT (ring/scaling.go:130-156) and
 * Ring.DivFloorByLastModulusNTT (ring/scaling.go:6-19); RescaleConstants from ring/ring.go:361-379.
 * kind 0 = power-of-two rings (ref_pow2 handles), 1 = 3N rings (ref_3n handles).  p0 = [nl][N] (NTT domain,
 * canonical), p1 = [nl-1][N].  The Go "INTTLazy" is canonical for N >= 16 (ring/ntt.go:197-206) and for
 * the 3N transformer (ntt_3n.go:154-156), which is what makes the result unique.
 * ------------------------------------------------------------------------- */
EXPORT void ref_div_by_last_modulus_ntt(int kind, void *const *rings, int nl, int N, const u64 *p0, u64 *p1, int round) {
    const int level = nl - 1;
    u64 *buffl = malloc(sizeof(u64) * N), *buffi = malloc(sizeof(u64) * N);
    u64 ql;
    if (kind == 1) { const ref_3n *r = rings[level]; ql = r->q; ref_3n_backward(r, p0 + (size_t)level * N, buffl, 0); }
    else { const ref_pow2 *r = rings[level]; ql = r->q; ref_pow2_intt(r, p0 + (size_t)level * N, buffl, 1); }
    const u64 phalf = (ql - 1) >> 1;
    if (round) for (int j = 0; j < N; j++) buffl[j] = cred(buffl[j] + phalf, ql);                 /* AddScalar, scaling.go:139 */
    for (int i = 0; i < level; i++) {
        u64 qi, qinv, brc[2];
        qi = kind == 1 ? ((const ref_3n *)rings[i])->q : ((const ref_pow2 *)rings[i])->q;
        ref_gen_bred_constant(qi, brc);
        qinv = ref_gen_mred_constant(qi);
        const u64 rc = ref_mform(qi - ref_modexp(ql, qi - 2, qi), qi, brc);                         /* ring.go:372-374 */
        const u64 add = round ? qi - ref_bred_add(phalf, qi, brc) : 0;                              /* scaling.go:142 */
        for (int j = 0; j < N; j++) buffi[j] = buffl[j] + add;                                      /* AddScalarLazy */
        if (kind == 1) ref_3n_forward(rings[i], buffi, buffi, 0); else ref_pow2_ntt(rings[i], buffi, buffi, 1);   /* NTTLazy */
        for (int j = 0; j < N; j++)                                                                  /* vec_ops.go:752-776 */
            p1[(size_t)i * N + j] = ref_mred(2 * qi - p0[(size_t)i * N + j] + buffi[j], rc, qi, qinv);
    }
    free(buffl); free(buffi);
}

/* ---------------------------------------------------------------------------
 * RNS basis extension: ring/basis_extension.go
 * ------------------------------------------------------------------------- */
typedef struct {
    int ns, nt;          /* source / target modulus counts */
    u64 *S, *T;
    u64 *qoverqiinvqi;   /* [ns]        ((Q/q_i)^-1 mod q_i), Montgomery form          :117-129 */
    u64 *qoverqimodp;    /* [nt][ns]    (Q/q_i mod p_j), Montgomery form                :131-142 */
    u64 *vtimesqmodp;    /* [nt][ns+1]  (-v*Q mod p_j), v = 0..ns                        :145-161 */
} ref_muc;

/* GenModUpConstants, basis_extension.go:95-166 */
EXPORT ref_muc *ref_gen_modup_constants(const u64 *Q, int nq, const u64 *P, int np) {
    ref_muc *m = calloc(1, sizeof *m);
    m->ns = nq; m->nt = np;
    m->S = malloc(sizeof(u64) * nq); m->T = malloc(sizeof(u64) * np);
    memcpy(m->S, Q, sizeof(u64) * nq); memcpy(m->T, P, sizeof(u64) * np);
    m->qoverqiinvqi = malloc(sizeof(u64) * nq);
    m->qoverqimodp = malloc(sizeof(u64) * (size_t)np * nq);
    m->vtimesqmodp = malloc(sizeof(u64) * (size_t)np * (nq + 1));
    for (int i = 0; i < nq; i++) {
        u64 qi = Q[i], brc[2]; ref_gen_bred_constant(qi, brc);
        u64 qinv = ref_gen_mred_constant(qi);
        u64 star = ref_mform(1, qi, brc);
        for (int j = 0; j < nq; j++) if (j != i) star = ref_mred(star, ref_mform(Q[j], qi, brc), qi, qinv);
        /* ModexpMontgomery(star, qi-2): Montgomery form of the inverse */
        u64 acc = ref_mform(1, qi, brc), base = star;
        for (u64 e = qi - 2; e; e >>= 1) { if (e & 1) acc = ref_mred(acc, base, qi, qinv); base = ref_mred(base, base, qi, qinv); }
        m->qoverqiinvqi[i] = acc;
        for (int j = 0; j < np; j++) {
            u64 pj = P[j], brp[2]; ref_gen_bred_constant(pj, brp);
            u64 pinv = ref_gen_mred_constant(pj), st = 1;
            for (int u = 0; u < nq; u++) if (u != i) st = ref_mred(st, ref_mform(Q[u], pj, brp), pj, pinv);
            m->qoverqimodp[(size_t)j * nq + i] = ref_mform(st, pj, brp);
        }
    }
    for (int j = 0; j < np; j++) {
        u64 pj = P[j], brp[2]; ref_gen_bred_constant(pj, brp);
        u64 pinv = ref_gen_mred_constant(pj), qmod = 1;
        for (int i = 0; i < nq; i++) qmod = ref_mred(qmod, ref_mform(Q[i], pj, brp), pj, pinv);
        u64 v = pj - qmod, *row = m->vtimesqmodp + (size_t)j * (nq + 1);
        row[0] = 0;
        for (int i = 1; i < nq + 1; i++) row[i] = cred(row[i - 1] + v, pj);
    }
    return m;
}
EXPORT void ref_muc_free(ref_muc *m) {
    if (m) { free(m->S); free(m->T); free(m->qoverqiinvqi); free(m->qoverqimodp); free(m->vtimesqmodp); free(m); }
}

/* One coefficient of reconstructRNS[Centered] (:504-594) + multSum (:597-673).
 * half == NULL: ModUpExact (:282-308).  half != NULL: the centered variant used by DecomposeAndSplit (:430-470).
 * src[i] = word of source limb i; dst[j] receives the lazy result of multSum for target j in `targets`. */
static void ext_one(const ref_muc *m, const u64 *src, const u64 *half, const int *targets, int ntargets, u64 *dst) {
    u64 y[64];
    double vi = 0;
    for (int i = 0; i < m->ns; i++) {
        u64 qi = m->S[i];
        y[i] = ref_mred(src[i] + (half ? half[i] : 0), m->qoverqiinvqi[i], qi, ref_gen_mred_constant(qi));
        vi += (double)y[i] / (double)qi;
    }
    const u64 v = (u64)vi;
    for (int t = 0; t < ntargets; t++) {
        const int j = targets[t];
        const u64 pj = m->T[j], pinv = ref_gen_mred_constant(pj);
        u128 acc = 0;
        for (int i = 0; i < m->ns; i++) acc += (u128)y[i] * m->qoverqimodp[(size_t)j * m->ns + i];
        const u64 hhi = mulhi((u64)acc * pinv, pj);
        dst[t] = (u64)(acc >> 64) - hhi + pj + m->vtimesqmodp[(size_t)j * (m->ns + 1) + v];
    }
}
/* p1 = [ns][N] source limbs, p2 = [ntargets][N]; targets = indices into the constants' target list */
EXPORT void ref_basis_extend(const ref_muc *m, const u64 *p1, const u64 *half, const int *targets, int ntargets, int N, u64 *p2) {
    u64 src[64], dst[128];
    for (int x = 0; x < N; x++) {
        for (int i = 0; i < m->ns; i++) src[i] = p1[(size_t)i * N + x];
        ext_one(m, src, half, targets, ntargets, dst);
        for (int t = 0; t < ntargets; t++) p2[(size_t)t * N + x] = dst[t];
    }
}
/* single-limb digit of DecomposeAndSplit (:402-428): centred lift of one limb to the targets */
EXPORT void ref_decompose_single(u64 q, const u64 *p0, const u64 *T, int nt, int N, u64 *p1) {
    for (int j = 0; j < N; j++) {
        u64 coeff = p0[j], pos = 1, neg = 0;
        if (coeff >= (q >> 1)) { coeff = q - coeff; pos = 0; neg = 1; }
        for (int i = 0; i < nt; i++) {
            u64 brc[2]; ref_gen_bred_constant(T[i], brc);
            const u64 tmp = ref_bred_add(coeff, T[i], brc);
            p1[(size_t)i * N + j] = tmp * pos + (T[i] - tmp) * neg;
        }
    }
}

/* ---------------------------------------------------------------------------
 * Synthetic inputs: splitmix64 counter generator, rejection-sampled with the
 * prime's bit mask like ring/sampler_uniform.go:69-95.  Shared by C / Python / CUDA harness.
 * ------------------------------------------------------------------------- */
static inline u64 splitmix64(u64 x) {
    x += 0x9E3779B97F4A7C15ULL;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL;
    return x ^ (x >> 31);
}
EXPORT void ref_fill_uniform(u64 seed, u64 poly, u64 limb, u64 q, int N, u64 *out) {
    u64 mask = 1; while (mask < q - 1) mask = (mask << 1) | 1;
    for (int i = 0; i < N; i++) {
        u64 ctr = seed ^ ((poly << 40) | (limb << 24) | (u64)i), v, tries = 0;
        do { v = splitmix64(ctr + (tries++ << 56)) & mask; } while (v >= q);
        out[i] = v;
    }
}

/* ---------------------------------------------------------------------------
 * Batched CPU drivers for the timed baseline (bench.py cpu_baseline / --impl reference).
 * Ring.NTT is a serial loop over limbs (ring/ntt.go:127-131); the threaded variant spreads
 * independent (poly, limb) units over host cores.  data = [units][N], unit u uses ring u % L.
 * ------------------------------------------------------------------------- */
EXPORT int ref_num_threads(void) { return host_threads(); }
typedef struct { void *const *rings; int L, N, inverse, kind; u64 *data; } batch_job;
static void batch_body(int u, void *arg) {
    batch_job *j = arg;
    u64 *p = j->data + (size_t)u * j->N;
    if (j->kind == 1) {
        const ref_3n *r = j->rings[u % j->L];
        if (j->inverse) ref_3n_backward(r, p, p, 0); else ref_3n_forward(r, p, p, 0);
    } else {
        const ref_pow2 *r = j->rings[u % j->L];
        if (j->inverse) ref_pow2_intt(r, p, p, 0); else ref_pow2_ntt(r, p, p, 0);
    }
}
EXPORT void ref_3n_batch(ref_3n *const *rings, int L, u64 *data, int units, int inverse, int threads) {
    batch_job j = {(void *const *)rings, L, rings[0]->N, inverse, 1, data};
    parallel_for(units, threads, batch_body, &j);
}
EXPORT void ref_pow2_batch(ref_pow2 *const *rings, int L, u64 *data, int units, int inverse, int threads) {
    batch_job j = {(void *const *)rings, L, rings[0]->N, inverse, 0, data};
    parallel_for(units, threads, batch_body, &j);
}
