/*
 * b200ring.h -- C ABI of the B200-native RNS NTT / NTT-domain multiply library.
 *
 * This is the drop-in boundary for the polynomial-arithmetic hot path of the Lattigo fork
 * swanhong/matrix-fhe-lattigo (package ring/).  Every entry point names the reference
 * interface it replaces (file:line, relative to the reference root).  A thin cgo file in
 * package `ring` binds these symbols; tests and bench.py bind them
 * with ctypes.  Plain pointers and sizes only; no C++ or torch types.
 *
 * Conventions
 *   - every function returns 0 on success or a negative B200_E* code; the message of the
 *     last failure on the calling thread is b200_last_error().  The reference's constructors
 *     return `error`, its hot functions panic (ring/ntt.go:211-214): the binding converts.
 *   - a context is immutable after creation and may be used from many host threads
 *     (ring/ring.go:192-194); work is enqueued on the context's CUDA stream and is
 *     asynchronous until b200_sync() or a download.
 *   - device polynomials are one contiguous [batch][L][N] uint64 allocation, limb-major like
 *     ring.Poly.Coeffs (ring/poly.go:13-24); `level` selects limbs 0..level like
 *     Ring.AtLevel (ring/ring.go:194-213, ring/ntt.go:128).
 *   - in-place calls (in == out) are allowed everywhere (ring/ntt_3n.go:90-96).
 *   - there is no CPU fallback: without a CUDA device every compute call fails with
 *     B200_ENODEVICE.
 */
#ifndef B200RING_H
#define B200RING_H

#include <stddef.h>
#include <stdint.h>

#if defined(__GNUC__)
#define B200_API __attribute__((visibility("default")))
#else
#define B200_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

typedef struct b200_ctx b200_ctx;   /* ring.Ring (ring/ring.go:76-89) + its SubRings on one GPU */
typedef struct b200_poly b200_poly; /* ring.Poly (ring/poly.go:13-24), device resident, batched  */

/* ring.Type (ring/ring.go:30-34) */
#define B200_RING_STANDARD 0            /* Z_q[X]/(X^N + 1), N = 2^k                                  */
#define B200_RING_CONJUGATE_INVARIANT 1 /* Z_q[X + X^-1]/(X^2N + 1), N = 2^k coefficients (ring/ntt.go:80-124) */
#define B200_RING_MATRIX 2              /* Z_q[X]/(X^N - X^(N/2) + 1), N = 2^a 3^b                    */

/* error codes */
#define B200_OK 0
#define B200_EINVAL (-1)    /* bad argument / invalid ring parameters (constructor `error`) */
#define B200_ENODEVICE (-2) /* no CUDA device or extension built without one                */
#define B200_ECUDA (-3)     /* CUDA runtime failure; see b200_last_error()                  */
#define B200_ESHAPE (-4)    /* polynomial / level / length mismatch (the reference panics)  */
#define B200_ENOMEM (-5)

/* flags of b200_ntt / b200_intt */
#define B200_LAZY 1       /* NTTLazy / INTTLazy: output only guaranteed in [0,6q-2] / [0,2q-1]
                             (ring/ntt.go:63-78); a canonical result is a valid lazy result */
#define B200_TREE_ORDER 2 /* MATRIX rings only: keep NTT-domain slots in the transform's native
                             tree order instead of the reference order (slot k <-> omega^E[k],
                             E ascending totatives of 3N, ring/ntt_3n.go:46-53,235-243).
                             MulCoeffs* is order agnostic, so fused pipelines may use it.    */

/* Element-wise operators: op codes of b200_vec_op.  One code per function of ring/vec_ops.go /
 * method of ring/operations.go; results are bit-identical to the Go functions, lazy ones included. */
#define B200_OP_ADD 0                  /* Ring.Add              operations.go:11  vec_ops.go:7    */
#define B200_OP_ADD_LAZY 1             /* Ring.AddLazy          operations.go:18  vec_ops.go:31   */
#define B200_OP_SUB 2                  /* Ring.Sub              operations.go:25  vec_ops.go:55   */
#define B200_OP_SUB_LAZY 3             /* Ring.SubLazy          operations.go:32  vec_ops.go:79   */
#define B200_OP_NEG 4                  /* Ring.Neg              operations.go:39  vec_ops.go:103  */
#define B200_OP_REDUCE 5               /* Ring.Reduce           operations.go:46  vec_ops.go:125  */
#define B200_OP_REDUCE_LAZY 6          /* Ring.ReduceLazy       operations.go:53  vec_ops.go:147  */
#define B200_OP_MFORM 7                /* Ring.MForm            operations.go:285 vec_ops.go:778  */
#define B200_OP_MFORM_LAZY 8           /* Ring.MFormLazy        operations.go:292 vec_ops.go:800  */
#define B200_OP_IMFORM 9               /* Ring.IMForm           operations.go:299 vec_ops.go:822  */
#define B200_OP_ZERO 10                /* ring.ZeroVec                                    vec_ops.go:847  (p1 ignored) */
#define B200_OP_MUL_BARRETT 16         /* Ring.MulCoeffsBarrett             operations.go:60  */
#define B200_OP_MUL_BARRETT_LAZY 17    /* Ring.MulCoeffsBarrettLazy         operations.go:67  */
#define B200_OP_MUL_BARRETT_ADD 18     /* Ring.MulCoeffsBarrettThenAdd      operations.go:74  */
#define B200_OP_MUL_BARRETT_ADD_LAZY 19/* Ring.MulCoeffsBarrettThenAddLazy  operations.go:81  */
#define B200_OP_MUL_LAZY 20            /* SubRing.MulCoeffsLazy            subring_ops.go:55  vec_ops.go:169: low word of p1*p2 */
#define B200_OP_MUL_LAZY_ADD_LAZY 21   /* SubRing.MulCoeffsLazyThenAddLazy subring_ops.go:62  vec_ops.go:193 */
#define B200_OP_MUL_MONT 32            /* Ring.MulCoeffsMontgomery                 operations.go:88  */
#define B200_OP_MUL_MONT_LAZY 33       /* Ring.MulCoeffsMontgomeryLazy             operations.go:95  */
#define B200_OP_MUL_MONT_ADD 34        /* Ring.MulCoeffsMontgomeryThenAdd          operations.go:109 */
#define B200_OP_MUL_MONT_ADD_LAZY 35   /* Ring.MulCoeffsMontgomeryThenAddLazy      operations.go:116 */
#define B200_OP_MUL_MONT_LAZY_ADD_LAZY 36 /* Ring.MulCoeffsMontgomeryLazyThenAddLazy operations.go:123 */
#define B200_OP_MUL_MONT_SUB 37        /* Ring.MulCoeffsMontgomeryThenSub          operations.go:130 */
#define B200_OP_MUL_MONT_SUB_LAZY 38   /* Ring.MulCoeffsMontgomeryThenSubLazy      operations.go:137 */
#define B200_OP_MUL_MONT_LAZY_SUB_LAZY 39 /* Ring.MulCoeffsMontgomeryLazyThenSubLazy operations.go:144 */
#define B200_OP_MUL_MONT_LAZY_NEG 40   /* Ring.MulCoeffsMontgomeryLazyThenNeg      operations.go:102 */
#define B200_OP_MULSCALAR_MONT 48      /* SubRing.MulScalarMontgomery        subring_ops.go:202 */
#define B200_OP_MULSCALAR_MONT_LAZY 49 /* SubRing.MulScalarMontgomeryLazy    subring_ops.go:209 */
#define B200_OP_MULSCALAR_MONT_ADD 50  /* SubRing.MulScalarMontgomeryThenAdd subring_ops.go:216 */
#define B200_OP_ADD_SCALAR 51          /* SubRing.AddScalar                  subring_ops.go:174 */
#define B200_OP_SUB_SCALAR 52          /* SubRing.SubScalar                  subring_ops.go:195 */
#define B200_OP_SUB_MULSCALAR_MONT_2Q 53 /* SubRing.SubThenMulScalarMontgomeryTwoModulus subring_ops.go:230:
                                            p3 = MRed(2q - p2 + p1, scalar_i)                          */
#define B200_OP_ADD_LAZY_MULSCALAR_MONT 54 /* SubRing.AddLazyThenMulScalarMontgomery subring_ops.go:160: MRed(p1 + p2, scalar_i) */
#define B200_OP_ADD_SCALAR_LAZY 55     /* SubRing.AddScalarLazy              subring_ops.go:181: p1 + scalar_i          */
#define B200_OP_ADD_SCALAR_LAZY_NEG_2Q_LAZY 56 /* SubRing.AddScalarLazyThenNegTwoModulusLazy subring_ops.go:188: scalar_i + 2q - p1 */
/* two scalars per limb: scalars[0..level] = scalar0_i, scalars[level+1..2*level+1] = scalar1_i */
#define B200_OP_ADD_SCALAR_LAZY_MULSCALAR_MONT 57 /* SubRing.AddScalarLazyThenMulScalarMontgomery subring_ops.go:167: MRed(p1 + scalar0_i, scalarMont1_i) */
#define B200_OP_MULSCALAR_MONT_ADD_SCALAR 58      /* SubRing.MulScalarMontgomeryThenAddScalar subring_ops.go:223: CRed(MRed(p1, scalarMont1_i) + scalar0_i) */
#define B200_OP_MASK 59                /* ring.MaskVec vec_ops.go:870: (p1 >> scalar0_i) & scalar1_i                     */
/* one scalar per limb and half (N % 4 == 0): words [0, N/2) use scalars[i], words [N/2, N) use scalars[level+1+i] */
#define B200_OP_ADD_SCALAR_DOUBLE 60         /* Ring.AddDoubleRNSScalar        operations.go:167 */
#define B200_OP_SUB_SCALAR_DOUBLE 61         /* Ring.SubDoubleRNSScalar        operations.go:177 */
#define B200_OP_MULSCALAR_MONT_DOUBLE 62     /* Ring.MulDoubleRNSScalar        operations.go:250 (scalars already in Montgomery form) */
#define B200_OP_MULSCALAR_MONT_ADD_DOUBLE 63 /* Ring.MulDoubleRNSScalarThenAdd operations.go:260 */

/* ---- library ------------------------------------------------------------------------------- */
B200_API const char *b200_last_error(void);   /* message of the last failure on this thread ("" if none) */
B200_API const char *b200_version(void);
B200_API int b200_device_count(void);         /* 0 when no CUDA device is visible */
/* Where the twiddle of heap index h (h = (R + row) * 2^d + k, R = N/256, the subtree under one 256-word row) sits in the
 * row-pass copies of the tables the two-pass kernels read (no reference counterpart: the Go tables are the heap order,
 * ring/subring.go:187-205; b200_ctx_roots still returns that order).  Host only, no device needed: returns the index in
 * [R, N), h itself for h < R, -1 for an invalid N or h. */
B200_API int b200_row_pass_index(int N, int heap_index);

/* ---- context: ring.NewRing / NewRingFromType / NewRingWithCustomNTT (ring/ring.go:264-356) ----
 * Validates like the reference (ring degree ring.go:319-323 / subring.go:14-30, non-empty and
 * distinct moduli ring.go:325-331, prime and = 1 mod NthRoot subring.go:140-146), generates the
 * per-prime constants and twiddle tables (subring.go:74-111,129-214; ntt_3n.go:35-79) and uploads
 * them to `device`.
 *   roots == NULL or roots[i] == 0 : deterministic choice of the reference --
 *       STANDARD: g = smallest primitive root >= 3 (subring.go:236-248), psi = g^((q-1)/2N);
 *       CONJUGATE_INVARIANT (NewRingConjugateInvariant, ring.go:280-282): the same g, psi = g^((q-1)/4N), q = 1 mod 4N;
 *       MATRIX  : omega = g^((q-1)/3N) (subring.go:255-290).
 *   roots[i] != 0 : STANDARD: the primitive root g of q_i (SubRing.PrimitiveRoot, skips factoring q-1);
 *                   MATRIX  : the primitive 3N-th root omega the Go transformer drew at random
 *                             (NumberTheoreticTransformer3N.psi3N, ntt_3n.go:39), so NTT-domain
 *                             values agree with that Go ring.
 */
B200_API int b200_ctx_create(int ring_type, int N, int L, const uint64_t *moduli, const uint64_t *roots,
                    int device, b200_ctx **out);
B200_API void b200_ctx_destroy(b200_ctx *ctx);
B200_API int b200_ctx_set_stream(b200_ctx *ctx, void *cuda_stream); /* cudaStream_t; NULL = default stream */
B200_API int b200_sync(b200_ctx *ctx);                              /* wait for all enqueued work */
B200_API int b200_ctx_n(const b200_ctx *ctx);                       /* Ring.N()                 */
B200_API int b200_ctx_moduli_count(const b200_ctx *ctx);            /* Ring.ModuliChainLength() */
B200_API int b200_ctx_ring_type(const b200_ctx *ctx);
/* SubRing fields of limb i (ring/subring.go:35-55, ring/ntt.go:38-44):
 *   out[0]=Modulus out[1]=MRedConstant out[2..3]=BRedConstant out[4]=Mask
 *   out[5]=PrimitiveRoot g out[6]=NInv (Montgomery form; not MATRIX) out[7]=root used (psi or omega) */
B200_API int b200_ctx_subring(const b200_ctx *ctx, int limb, uint64_t out[8]);
/* STANDARD / CONJUGATE_INVARIANT rings: NTTTable.RootsForward (inverse=0) / RootsBackward (inverse=1), NthRoot/2 =
 * N (2N for conjugate-invariant rings) words, Montgomery form, bit-reversed order -- bit-identical to subring.go:187-205. */
B200_API int b200_ctx_roots(const b200_ctx *ctx, int limb, int inverse, uint64_t *out, size_t n);

/* ---- device polynomials: ring.Poly / Ring.NewPoly (ring/poly.go:18-24) ---------------------- */
B200_API int b200_poly_alloc(b200_ctx *ctx, int batch, b200_poly **out); /* zero-filled [batch][L][N]   */
/* wrap caller-owned device memory of batch*L*N words (e.g. a torch tensor); never freed here */
B200_API int b200_poly_wrap(b200_ctx *ctx, void *device_ptr, int batch, b200_poly **out);
B200_API void b200_poly_free(b200_poly *p);
B200_API void *b200_poly_device_ptr(b200_poly *p);
B200_API int b200_poly_batch(const b200_poly *p);
/* one limb of one polynomial <-> a Go []uint64 (Poly.Coeffs[limb]); n must be >= N (ntt.go:212) */
B200_API int b200_poly_upload_limb(b200_poly *p, int b, int limb, const uint64_t *host, size_t n);
B200_API int b200_poly_download_limb(b200_poly *p, int b, int limb, uint64_t *host, size_t n);
/* the whole [batch][L][N] block from/to one contiguous host buffer (pinned for full PCIe rate) */
B200_API int b200_poly_upload(b200_poly *p, const uint64_t *host, size_t n_words);
B200_API int b200_poly_download(b200_poly *p, uint64_t *host, size_t n_words);
/* Poly.Copy / CopyLvl (ring/poly.go:73-91): limbs 0..level (level < 0: all) of every polynomial of src into dst, on the
 * device, ordered on the ring's stream.  dst == src is a no-op. */
B200_API int b200_poly_copy(b200_poly *dst, const b200_poly *src, int level);

/* ---- transforms ---------------------------------------------------------------------------
 * Ring.NTT / NTTLazy (ring/ntt.go:127-138) and Ring.INTT / INTTLazy (ring/ntt.go:141-152) on all
 * batch x (level+1) limbs in one enqueue.  STANDARD: NumberTheoreticTransformerStandard
 * (ntt.go:59-78; out[i] = f(psi^(2 bitrev(i)+1)), canonical).  MATRIX:
 * NumberTheoreticTransformer3N.Forward/Backward (ntt_3n.go:82-156) computed in O(N log N). */
B200_API int b200_ntt(b200_ctx *ctx, const b200_poly *in, b200_poly *out, int level, int flags);
B200_API int b200_intt(b200_ctx *ctx, const b200_poly *in, b200_poly *out, int level, int flags);

/* ---- wire format -------------------------------------------------------------
 * Poly.BinarySize / WriteTo / MarshalBinary / ReadFrom / UnmarshalBinary (ring/poly.go:97-160): the polynomial is a
 * structs.Matrix[uint64] (utils/structs/matrix.go:57-128, vector.go:43-100) -- u64 limb count, then per limb a u64
 * length and the words, little-endian.  `b` selects the polynomial of a batched handle; limbs 0..level are written.
 * Unmarshal accepts at most L limbs of exactly N words (anything else is B200_ESHAPE and leaves the polynomial
 * untouched), reports the level found (rows-1) and the bytes consumed, and leaves limbs above it unchanged. */
B200_API int b200_poly_binary_size(const b200_ctx *ctx, int level, size_t *out);
B200_API int b200_poly_marshal(const b200_poly *p, int b, int level, void *buf, size_t cap, size_t *written);
B200_API int b200_poly_unmarshal(b200_poly *p, int b, const void *buf, size_t len, int *level_out, size_t *consumed);
/* SubRing.Factors (ring/subring.go:45, :333-360): unique prime factors of Modulus-1, ascending; returns the count
 * (out may be NULL to query it) or a negative error. */
B200_API int b200_ctx_factors(const b200_ctx *ctx, int limb, uint64_t *out, int cap);

/* ---- fused all-gather over NVLink peer memory (one process per GPU) ------------------------------
 * The key switch is the one step of the path with a real exchange (ring/basis_extension.go:381-502 needs every limb
 * of a digit per coefficient).  b200_poly_ipc_export gives a 64-byte CUDA IPC handle of a polynomial allocated by
 * b200_poly_alloc; a peer process maps it with b200_peer_open (and unmaps with b200_peer_close).
 * b200_intt_limbs_gather = Ring.INTT on limbs [limb0, limb0+nl) of `in` into `out` and into the same limbs of the
 * npeers (<= 7) mapped polynomials, the stores to the peers issued by the transform's last kernel.  The caller
 * orders the peers' reads after it (any cross-GPU barrier on the stream, e.g. a one-word all-reduce). */
B200_API int b200_poly_ipc_export(const b200_poly *p, unsigned char handle[64]);
B200_API int b200_peer_open(b200_ctx *ctx, const unsigned char handle[64], void **device_ptr);
B200_API int b200_peer_close(b200_ctx *ctx, void *device_ptr);
B200_API int b200_intt_limbs_gather(b200_ctx *ctx, const b200_poly *in, b200_poly *out, int limb0, int nl, int flags,
                           int npeers, void *const *peer_polys);

/* ---- element-wise family: p3 = op(p1, p2 [, p3]) (ring/operations.go, ring/vec_ops.go) -------
 * p2 may be NULL for unary ops.  For *_SCALAR / MULSCALAR ops `scalars` holds one word per limb
 * (level+1 of them; Montgomery form where the Go method says so), two per limb for op codes 57..63; otherwise NULL. */
B200_API int b200_vec_op(b200_ctx *ctx, int op, const b200_poly *p1, const b200_poly *p2, b200_poly *p3,
                int level, const uint64_t *scalars);
/* Ring.MulByVectorMontgomery / MulByVectorMontgomeryThenAddLazy (ring/operations.go:366-377): every limb of every
 * polynomial of p1 times the same N words (limb 0 of polynomial 0 of `vector`), p2 = MRed(p1, vector) or p2 += ... */
B200_API int b200_mul_by_vector_montgomery(b200_ctx *ctx, const b200_poly *p1, const b200_poly *vector, b200_poly *p2,
                                           int level, int then_add_lazy);

/* ---- key switching: RNS basis extension, ModDown, gadget product ----------------------
 * The ring of `ctx` is the concatenated chain Q_0..Q_{LQ-1}, P_0..P_{LP-1} (what ringqp.Ring fans out over,
 * ring/ringqp/ring.go:12-17); polynomials keep the [batch][LQ+LP][N] layout, Q limbs first.
 * b200_keyswitch_create precomputes what ring.NewBasisExtender / ring.NewDecomposer hold
 * (ring/basis_extension.go:24-73,95-166,319-377); extension constants are generated on first use and cached. */
typedef struct b200_keyswitch b200_keyswitch;
B200_API int b200_keyswitch_create(b200_ctx *ctx, int LQ, b200_keyswitch **out);
B200_API void b200_keyswitch_destroy(b200_keyswitch *ks);
/* BasisExtender.ModUpQtoP (direction 0, basis_extension.go:184-198: reads Q limbs 0..levelQ of `in`, writes P limbs
 * 0..levelP of `out`) / ModUpPtoQ (direction 1, :203-217).  Coefficient domain.  The words written are exactly the
 * lazy representatives the Go code leaves (ModUpExact :282-308 then SubScalarBigint), float64 correction included. */
B200_API int b200_mod_up(b200_keyswitch *ks, int direction, int levelQ, int levelP, const b200_poly *in, b200_poly *out);
/* BasisExtender.ModDownQPtoQNTT (:241-258): p1 in basis QP, NTT domain -> p2 limbs 0..levelQ = round(p1 / P), canonical */
B200_API int b200_mod_down_ntt(b200_keyswitch *ks, int levelQ, int levelP, const b200_poly *p1, b200_poly *p2);
/* BasisExtender.ModDownQPtoQ (to_p == 0, :223-236: p2 limbs 0..levelQ = round(p1 / P)) and ModDownQPtoP (to_p != 0, :264-279:
 * p2 limbs LQ..LQ+levelP = (p1 - [p1]_Q) / Q with [.]_Q the centred residue) on coefficient-domain polynomials, canonical. */
B200_API int b200_mod_down(b200_keyswitch *ks, int to_p, int levelQ, int levelP, const b200_poly *p1, b200_poly *p2);
/* rlwe.Evaluator.GadgetProduct for NTT-domain operands (core/rlwe/evaluator_gadget_product.go:16-30, 123-188, 455-478;
 * with one special prime, levelP == 0, the reference's :190-320 with mask == 0, which gives the same result): ct = (<decomp(cx), key[.][0]>, <decomp(cx), key[.][1]>) / P mod Q.  cx: batch B, Q limbs used.
 * key: 2*beta polynomials ordered [digit][component], NTT domain and Montgomery form over QP
 * (GadgetCiphertext.Value[digit][0][component]), beta = (levelQ+levelP+1)/(levelP+1).  ct0, ct1: batch B, canonical. */
B200_API int b200_gadget_product(b200_keyswitch *ks, int levelQ, int levelP, const b200_poly *cx, const b200_poly *key,
                        b200_poly *ct0, b200_poly *ct1);
/* The same for a gadget ciphertext with a power-of-two base (BaseTwoDecomposition = pw2 > 0, which the reference allows with
 * one special prime only, core/rlwe/params.go:615-632): gadgetProductSinglePAndBitDecompLazy with mask != 0
 * (evaluator_gadget_product.go:190-320) + ModDown.  sizes[i] = len(gadgetCt.Value[i]) for i <= levelQ; `key` holds
 * 2*sum(sizes) polynomials ordered [limb i][window j][component].  pw2 == 0 forwards to b200_gadget_product. */
B200_API int b200_gadget_product_bit_decomp(b200_keyswitch *ks, int levelQ, int levelP, int pw2, const int *sizes, const b200_poly *cx,
                                   const b200_poly *key, b200_poly *ct0, b200_poly *ct1);
/* Hoisted key switch (one decomposition, many keys: rotations of a linear transform).
 * b200_decompose_ntt: rlwe.Evaluator.DecomposeNTT (core/rlwe/evaluator_gadget_product.go:430-478) for an NTT-domain cx:
 *   decomp receives beta*B polynomials ordered [digit][b], every digit over all limbs 0..levelQ and P_0..levelP, NTT domain,
 *   canonical (BuffDecompQP[digit]).
 * b200_gadget_product_hoisted: GadgetProductHoisted (:322-423): multiply-accumulate of `decomp` with `key` + ModDown. */
B200_API int b200_decompose_ntt(b200_keyswitch *ks, int levelQ, int levelP, const b200_poly *cx, b200_poly *decomp);
B200_API int b200_gadget_product_hoisted(b200_keyswitch *ks, int levelQ, int levelP, const b200_poly *decomp, const b200_poly *key,
                                b200_poly *ct0, b200_poly *ct1);
/* rlwe.Evaluator.Automorphism (core/rlwe/evaluator_automorphism.go:13-60) on an NTT-domain ciphertext (ct0, ct1) of a
 * power-of-two ring: (t0, t1) = GadgetProduct(ct1, key); t0 += ct0; out_c = AutomorphismNTT(t_c, galEl).  With decomp != NULL
 * (the output of b200_decompose_ntt on ct1) it is AutomorphismHoisted (:62-108). */
B200_API int b200_automorphism_ct(b200_keyswitch *ks, int levelQ, int levelP, const b200_poly *ct0, const b200_poly *ct1,
                         const b200_poly *decomp, uint64_t galEl, const b200_poly *key, b200_poly *out0, b200_poly *out1);

/* Ciphertext x ciphertext multiply with relinearisation, NTT domain in and out (BASELINE config 5):
 * (c0, c1, c2) = tensor(a, b) as b200_tensor, then (out0, out1) = (c0, c1) + GadgetProduct(c2, rlk)
 * (schemes/ckks/evaluator.go:821-853, core/rlwe/evaluator_evaluationkey.go:125-153).  The reference never composes
 * this on a 3N ring (matrix_ckks.Evaluator.Mul stops at the tensor product): defined here as those steps in sequence. */
B200_API int b200_mul_relin(b200_keyswitch *ks, int levelQ, int levelP, const b200_poly *a0, const b200_poly *a1,
                   const b200_poly *b0, const b200_poly *b1, const b200_poly *rlk, b200_poly *out0, b200_poly *out1,
                   int mform_first);

/* Limb-sharded execution of the same gadget product (BASELINE config 5): every rank holds the QP ring
 * but owns the Q limbs [q0, q1) and the P limbs [p0, p1); the two exchanges of the path are NCCL all-gathers the
 * caller runs on the polynomials' device memory between these steps (see sharding.gadget_product_sharded):
 *   1. b200_ntt_limbs(inverse) on the owned Q limbs of cx          -> all-gather the Q limbs of cxinv
 *   2. b200_ks_decompose_mac: every digit into the owned targets, NTT, multiply-accumulate with the owned key limbs
 *   3. b200_ntt_limbs(inverse) on the owned P limbs of acc0/acc1   -> all-gather the P limbs
 *   4. b200_ks_mod_down_limbs: extension P -> owned Q limbs, NTT, final pass -> owned Q limbs of the result */
B200_API int b200_ntt_limbs(b200_ctx *ctx, const b200_poly *in, b200_poly *out, int limb0, int nl, int inverse, int flags);
B200_API int b200_ks_decompose_mac(b200_keyswitch *ks, int levelQ, int levelP, const b200_poly *cx, const b200_poly *cxinv,
                          const b200_poly *key, b200_poly *acc0, b200_poly *acc1, int q0, int q1, int p0, int p1);
B200_API int b200_ks_mod_down_limbs(b200_keyswitch *ks, int levelQ, int levelP, const b200_poly *acc, b200_poly *out, int q0, int q1);

/* ---- RNS rescale in the NTT domain ------------------------------------------------
 * Ring.DivRoundByLastModulusNTT (round != 0, ring/scaling.go:130-156) / DivFloorByLastModulusNTT (round == 0,
 * ring/scaling.go:6-19) at `level` >= 1 for the whole batch: p1 limbs 0..level-1 = (p0 - [p0]_{q_level}) / q_level.
 * p0 in the NTT domain, canonical; p1 may equal p0 (limb `level` of p1 is left untouched). */
B200_API int b200_div_by_last_modulus_ntt(b200_ctx *ctx, const b200_poly *p0, b200_poly *p1, int level, int round);
/* Ring.DivFloorByLastModulus / DivRoundByLastModulus (ring/scaling.go:21-30,113-127): the same division on
 * coefficient-domain polynomials.  p0 is only read (the Go DivRound leaves intermediate words in p0); p1 may equal p0. */
B200_API int b200_div_by_last_modulus(b200_ctx *ctx, const b200_poly *p0, b200_poly *p1, int level, int round);
/* Ring.Div{Floor,Round}ByLastModulusMany (ntt == 0, ring/scaling.go:57-91,160-192) / ...ManyNTT (ntt != 0, :32-55,130-158):
 * nb_rescales (0..level) successive divisions by q_level, q_level-1, ...; the result is in limbs 0..level-nb_rescales of
 * p1 (nb_rescales == 0 copies limbs 0..level).  The `buff` polynomial of the Go methods is internal. */
B200_API int b200_div_by_last_modulus_many(b200_ctx *ctx, const b200_poly *p0, b200_poly *p1, int level, int nb_rescales,
                                           int round, int ntt);

/* ---- fused ciphertext tensor product in the NTT domain ----------------------------
 * (a0, a1) x (b0, b1) -> c0 = a0*b0, c1 = a0*b1 + a1*b0, c2 = a1*b1, slot-wise, in ONE pass over 4 inputs and
 * 3 outputs.  mform_first != 0: ckks.Evaluator.mulRelin (schemes/ckks/evaluator.go:821-834): MForm(a0), MForm(a1),
 * then MulCoeffsMontgomery / ...ThenAdd, i.e. c = a*b mod q.  mform_first == 0: matrix_ckks.Evaluator.Mul
 * (schemes/matrix_ckks/evaluator.go:163-173): plain MulCoeffsMontgomery, c = a*b*2^-64 mod q.  Bit-identical
 * to those sequences (also for squaring, a == b).  Outputs may alias inputs but must be distinct from each other. */
B200_API int b200_tensor(b200_ctx *ctx, const b200_poly *a0, const b200_poly *a1, const b200_poly *b0, const b200_poly *b1,
                b200_poly *c0, b200_poly *c1, b200_poly *c2, int level, int mform_first);

/* Ring.Shift (ring/operations.go:278-282): p2[j] = p1[(j + k) mod N] on every limb of the polynomial.
 * Ring.MultByMonomial (:306-363): p2 = X^k * p1 with X^N = -1 on limbs 0..level; a negated zero is written as q, exactly
 * like the Go loops.  p1 == p2 is allowed. */
B200_API int b200_shift(b200_ctx *ctx, const b200_poly *p1, int k, b200_poly *p2);
B200_API int b200_mult_by_monomial(b200_ctx *ctx, const b200_poly *p1, int k, b200_poly *p2, int level);

/* ---- automorphisms X -> X^gen of power-of-two rings (ring/automorphism.go) -----------------------------------------
 * b200_automorphism_ntt_index: AutomorphismNTTIndex (:12-35), host only, N words.
 * b200_automorphism_ntt: Ring.AutomorphismNTT (:38-46), the table computed on the device; with add_lazy != 0 the result is
 *   added to `out` like AutomorphismNTTWithIndexThenAddLazy (:82-110).  b200_automorphism_ntt_with_index takes the
 *   caller's table (AutomorphismNTTWithIndex :50-79).  b200_automorphism: Ring.Automorphism (:113-176), coefficient
 *   domain, Standard and ConjugateInvariant branches.  None of them works in place (the reference says so, :37,:49,:112);
 *   Matrix rings and out-of-range tables return B200_ESHAPE where the Go code panics.
 * b200_fold_standard_to_ci: Ring.FoldStandardToConjugateInvariant (ring/conjugate_invariant.go:28-44). */
B200_API int b200_automorphism_ntt_index(int N, uint64_t NthRoot, uint64_t GalEl, uint64_t *index);
B200_API int b200_automorphism_ntt(b200_ctx *ctx, const b200_poly *in, uint64_t gen, b200_poly *out, int level, int add_lazy);
B200_API int b200_automorphism_ntt_with_index(b200_ctx *ctx, const b200_poly *in, const uint64_t *index, size_t n, b200_poly *out,
                                     int level, int add_lazy);
B200_API int b200_automorphism(b200_ctx *ctx, const b200_poly *in, uint64_t gen, b200_poly *out, int level);
B200_API int b200_fold_standard_to_ci(b200_ctx *ctx, const b200_poly *st, const uint64_t *index, size_t n, b200_poly *ci, int level);

/* Ring.UnfoldConjugateInvariantToStandard (ring/conjugate_invariant.go:7-23): limbs 0..level of `ci` (a polynomial of the
 * conjugate-invariant ring `ctx`, N words per limb) into `st`, a polynomial of a ring of degree 2N over the same moduli:
 * st[j] = ci[j], st[2N-1-j] = ci[j]. */
B200_API int b200_unfold_ci_to_standard(b200_ctx *ctx, const b200_poly *ci, b200_poly *st, int level);

/* ---- fused ring product (BASELINE config 4; ring/ring_test.go:106-110 sequence) ---------------
 * c = INTT( MulCoeffsMontgomery( MForm(NTT(a)), NTT(b) ) ) = a*b in the ring, coefficient domain,
 * canonical.  a, b are left untouched unless they alias c. */
B200_API int b200_polymul(b200_ctx *ctx, const b200_poly *a, const b200_poly *b, b200_poly *c, int level);

/* ---- host-buffer transforms (a caller without device-resident polynomials) ----------------------
 * Applies ops[0..n_ops) in order (0 = Ring.NTT, 1 = Ring.INTT, ring/ntt.go:127-152) to `batch` polynomials
 * held in HOST memory as [batch][L][N] words and writes the result to host_out (may equal host_in).  The
 * batch is streamed through the GPU in chunks with the H2D copy, the kernels and the D2H copy of
 * successive chunks overlapped on three streams; pinned host memory gives the full PCIe rate.
 * Synchronous: the result is in host_out on return. */
B200_API int b200_transform_host(b200_ctx *ctx, const uint64_t *host_in, uint64_t *host_out, int batch, int level,
                        const int *ops, int n_ops, int flags);

/* ---- plugin granularity: ring.NumberTheoreticTransformer (ring/ntt.go:15-22) -----------------
 * One limb, HOST slices in and out (upload -> transform -> download, synchronous).  This is what a
 * NewRingWithCustomNTT factory (ring/ring.go:314) binds; it is correct but not the fast path. */
B200_API int b200_subring_ntt_host(b200_ctx *ctx, int limb, const uint64_t *p1, uint64_t *p2, size_t n,
                          int inverse, int flags);

/* ---- modulus and root search of the ring constructors (host only: no device is touched) ---------------
 * What a caller runs before b200_ctx_create: core/rlwe/params.go:924,966,998 and core/rlwe/params_3n.go:804-805 pick
 * the moduli chain with these, ring/ntt_3n.go:39 picks omega.  Same integers as the Go functions. */
B200_API int b200_is_prime(uint64_t x);                                /* ring.IsPrime, ring/primes.go:11-13: 1 / 0 */
/* ring.Find3NRNSPrimes (ring/primes_3n.go:11-43): the first `count` primes = 1 mod 3N above 2^bitSize, ascending,
 * giving up after searchBudget candidates (B200_EINVAL, "could not find enough 3N-friendly primes (found i/count)"). */
B200_API int b200_find_3n_rns_primes(int N, int bitSize, int count, int searchBudget, uint64_t *out);
/* ring.Find3NFriendlyPrime (ring/primes.go:16-67): first prime k*3N+1 at or above 2^bitSize (1000 candidates), else below. */
B200_API int b200_find_3n_friendly_prime(int N, int bitSize, uint64_t *out);
B200_API int b200_order_mod(uint64_t g, uint64_t p, uint64_t *out);     /* ring.OrderMod, ring/primes_3n.go:78-91 */
B200_API int b200_has_primitive_mth_root(uint64_t q, uint64_t m);       /* ring.HasPrimitiveMthRoot, :121-123: 1 / 0 */
B200_API int b200_primitive_root(uint64_t q, uint64_t *g_out);          /* ring.PrimitiveRoot, ring/subring.go:218-251 */
/* ring.FindPrimitiveRootOfUnity (ring/primes_3n.go:127-149): an element of exact order m.  The reference draws up to
 * 256 random bases from crypto/rand, so any element of order m is a valid answer; here the bases come from
 * splitmix64(seed).  seed == 0 returns the deterministic root g^((q-1)/m), g = b200_primitive_root(q)
 * (ring.Find3NPrimitiveRoot, ring/subring.go:255-290) -- the omega b200_ctx_create uses when roots == NULL. */
B200_API int b200_find_primitive_root_of_unity(uint64_t q, uint64_t m, uint64_t seed, uint64_t *out);
/* ring.Derive3NRoots (ring/primes_3n.go:165-181): out = {omega, omega^(3^(b+1)) of order 2^a, omega^(2^a) of order
 * 3^(b+1)}; B200_EINVAL unless omega has exact order 3N = 3 * 2^a * 3^b modulo the prime q. */
B200_API int b200_derive_3n_roots(uint64_t q, int a, int b, uint64_t omega, uint64_t out[3]);
/* ring.NTTFriendlyPrimesGenerator (ring/primes.go:69-277): primes 2^BitSize +/- k*NthRoot + 1.  The struct is the Go
 * struct, field for field.  direction +1 = NextUpstreamPrimes(k), -1 = NextDownstreamPrimes(k), 0 =
 * NextAlternatingPrimes(k); B200_EINVAL with the reference's message once a side overlaps the neighbouring bit size
 * (where the Go NextUpstreamPrime would spin forever on a second call, primes.go:147-149, the error is returned again). */
typedef struct b200_prime_gen {
    double Size;
    uint64_t NextPrime, PrevPrime, NthRoot;
    int CheckNextPrime, CheckPrevPrime;
} b200_prime_gen;
B200_API int b200_prime_gen_init(b200_prime_gen *g, uint64_t BitSize, uint64_t NthRoot);
B200_API int b200_prime_gen_next(b200_prime_gen *g, int direction, int k, uint64_t *out);

/* Per-kernel device timing for the roofline report: between begin and end every kernel launch of this
 * context is bracketed by CUDA events on its stream.  end() writes one line per kernel name:
 * "<kernel> <total_ms> <launches> <limb-units processed>\n". */
B200_API int b200_profile_begin(b200_ctx *ctx);
B200_API int b200_profile_end(b200_ctx *ctx, char *out, size_t cap);

/* kernel launches enqueued by this context so far (bench.py's gpu_launches evidence) */
B200_API uint64_t b200_ctx_launch_count(const b200_ctx *ctx);

#ifdef __cplusplus
}
#endif
#endif /* B200RING_H */
