// b200ring.hpp -- C++ host-side mirror of the reference's `ring` package surface for the hot path
// (ring.Ring, ring.SubRing, ring.Poly; ring/ring.go, ring/poly.go, ring/ntt.go, ring/operations.go),
// header-only, over the C ABI of b200ring.h.  The reference is Go; no Go toolchain exists in the build
// image, so the compiled host side is C++ with the same names, argument meaning and error behaviour:
// constructors throw std::invalid_argument where Go returns `error`, hot functions throw ring::Panic
// where Go panics.  INTEGRATION.md shows the equivalent cgo file for package `ring`.
#pragma once
#include <cstdint>
#include <array>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "b200ring.h"

namespace ring {

enum Type { Standard = 0, ConjugateInvariant = 1, Matrix = 2 };   // ring/ring.go:30-34

struct Panic : std::runtime_error {   // the reference's panic(...) on the hot path
    int code;
    Panic(int c, const std::string &m) : std::runtime_error(m), code(c) {}
};

inline void check(int rc) {
    if (rc == B200_OK) return;
    if (rc == B200_EINVAL) throw std::invalid_argument(b200_last_error());
    throw Panic(rc, b200_last_error());
}

// isValidRingDegreeFor3N, ring/subring.go:14-30
inline bool IsValidRingDegreeFor3N(int N) {
    if (N <= 0) return false;
    long t = 3L * N;
    while (t % 3 == 0) t /= 3;
    return t > 0 && (t & (t - 1)) == 0;
}

// ---- ring/primes.go, ring/primes_3n.go: modulus and root search (host only); Go `error` -> std::invalid_argument
inline bool IsPrime(uint64_t x) { return b200_is_prime(x) != 0; }                                             // primes.go:11-13
inline std::vector<uint64_t> Find3NRNSPrimes(int N, int bitSize, int count, int searchBudget) {                // primes_3n.go:11-43
    std::vector<uint64_t> out(count > 0 ? count : 1);
    check(b200_find_3n_rns_primes(N, bitSize, count, searchBudget, out.data()));
    out.resize(count);
    return out;
}
inline uint64_t Find3NFriendlyPrime(int N, int bitSize) { uint64_t q = 0; check(b200_find_3n_friendly_prime(N, bitSize, &q)); return q; }   // primes.go:16-67
inline uint64_t OrderMod(uint64_t g, uint64_t p) { uint64_t o = 0; check(b200_order_mod(g, p, &o)); return o; }                    // primes_3n.go:78-91
inline bool HasPrimitiveMthRoot(uint64_t q, uint64_t m) { return b200_has_primitive_mth_root(q, m) != 0; }                         // primes_3n.go:121-123
inline uint64_t PrimitiveRoot(uint64_t q) { uint64_t g = 0; check(b200_primitive_root(q, &g)); return g; }                         // subring.go:218-251
inline uint64_t FindPrimitiveRootOfUnity(uint64_t q, uint64_t m, uint64_t seed = 0) {                                              // primes_3n.go:127-149
    uint64_t w = 0;
    check(b200_find_primitive_root_of_unity(q, m, seed, &w));
    return w;
}
inline uint64_t Find3NPrimitiveRoot(uint64_t p, uint64_t NthRoot) { return FindPrimitiveRootOfUnity(p, NthRoot, 0); }              // subring.go:255-290
struct Roots3N { uint64_t omega3N, omega2Pow, omega3Pow; };
inline Roots3N Derive3NRoots(uint64_t q, int a, int b, uint64_t omega) {                                                           // primes_3n.go:165-181
    uint64_t o[3];
    check(b200_derive_3n_roots(q, a, b, omega, o));
    return Roots3N{o[0], o[1], o[2]};
}
class NTTFriendlyPrimesGenerator {                                                                                                 // primes.go:69-277
  public:
    NTTFriendlyPrimesGenerator(uint64_t BitSize, uint64_t NthRoot) { check(b200_prime_gen_init(&g_, BitSize, NthRoot)); }
    std::vector<uint64_t> NextUpstreamPrimes(int k) { return next(1, k); }
    std::vector<uint64_t> NextDownstreamPrimes(int k) { return next(-1, k); }
    std::vector<uint64_t> NextAlternatingPrimes(int k) { return next(0, k); }
    uint64_t NextUpstreamPrime() { return next(1, 1)[0]; }
    uint64_t NextDownstreamPrime() { return next(-1, 1)[0]; }
    uint64_t NextAlternatingPrime() { return next(0, 1)[0]; }

  private:
    std::vector<uint64_t> next(int direction, int k) {
        std::vector<uint64_t> out(k > 0 ? k : 1);
        check(b200_prime_gen_next(&g_, direction, k, out.data()));
        out.resize(k > 0 ? k : 0);
        return out;
    }
    b200_prime_gen g_;
};
inline NTTFriendlyPrimesGenerator NewNTTFriendlyPrimesGenerator(uint64_t BitSize, uint64_t NthRoot) { return NTTFriendlyPrimesGenerator(BitSize, NthRoot); }

class Ring;

// ring.Poly (ring/poly.go:13-24): device resident, `batch` polynomials of [L][N] words
class Poly {
  public:
    Poly(const Ring &r, int batch = 1);
    Poly(const Poly &) = delete;
    Poly &operator=(const Poly &) = delete;
    Poly(Poly &&o) noexcept : h_(o.h_), words_(o.words_), n_(o.n_), batch_(o.batch_) { o.h_ = nullptr; }
    ~Poly() { if (h_) b200_poly_free(h_); }
    int Batch() const { return batch_; }
    // Poly.Coeffs[limb] of polynomial b  <->  a host slice of >= N words
    void SetLimb(int b, int limb, const std::vector<uint64_t> &v) { check(b200_poly_upload_limb(h_, b, limb, v.data(), v.size())); }
    std::vector<uint64_t> Limb(int b, int limb, size_t n) const {
        std::vector<uint64_t> v(n);
        check(b200_poly_download_limb(h_, b, limb, v.data(), v.size()));
        return v;
    }
    void Set(const std::vector<uint64_t> &all) { check(b200_poly_upload(h_, all.data(), all.size())); }
    std::vector<uint64_t> Get() const {
        std::vector<uint64_t> v(words_);
        check(b200_poly_download(h_, v.data(), v.size()));
        return v;
    }
    // Poly.MarshalBinary / UnmarshalBinary (ring/poly.go:97-160) of polynomial b: structs.Matrix[uint64] bytes
    std::vector<unsigned char> MarshalBinary(int b, int level) const {
        std::vector<unsigned char> buf(8 + (size_t)(level + 1) * (8 + 8 * n_));
        size_t n = 0;
        check(b200_poly_marshal(h_, b, level, buf.data(), buf.size(), &n));
        buf.resize(n);
        return buf;
    }
    int UnmarshalBinary(int b, const std::vector<unsigned char> &data, size_t *consumed = nullptr) {   // returns the level read
        int level = -1;
        check(b200_poly_unmarshal(h_, b, data.data(), data.size(), &level, consumed));
        return level;
    }
    std::array<unsigned char, 64> IpcHandle() const {   // CUDA IPC handle for Ring::OpenPeer in another process
        std::array<unsigned char, 64> h{};
        check(b200_poly_ipc_export(h_, h.data()));
        return h;
    }
    void *DevicePtr() const { return b200_poly_device_ptr(h_); }
    b200_poly *handle() const { return h_; }

  private:
    b200_poly *h_ = nullptr;
    size_t words_ = 0, n_ = 0;
    int batch_ = 0;
};

// ring.SubRing (ring/subring.go:35-55) view of one prime
struct SubRing {
    int N = 0;
    uint64_t Modulus = 0, MRedConstant = 0, BRedConstant[2] = {0, 0}, Mask = 0, PrimitiveRoot = 0, NInv = 0, Root = 0;
};

class Ring {
  public:
    // ring.NewRing (ring/ring.go:264-272)
    static Ring NewRing(int N, const std::vector<uint64_t> &Moduli, int device = 0, const std::vector<uint64_t> *roots = nullptr) {
        const bool is3n = N % 3 == 0 && IsValidRingDegreeFor3N(N);
        return NewRingFromType(N, Moduli, is3n ? Matrix : Standard, device, roots);
    }
    // ring.NewRingConjugateInvariant (ring/ring.go:280-282): NthRoot = 4N
    static Ring NewRingConjugateInvariant(int N, const std::vector<uint64_t> &Moduli, int device = 0) {
        return NewRingFromType(N, Moduli, ConjugateInvariant, device);
    }
    // ring.NewRingFromType (ring/ring.go:291-308)
    static Ring NewRingFromType(int N, const std::vector<uint64_t> &Moduli, Type t, int device = 0,
                                const std::vector<uint64_t> *roots = nullptr) {
        b200_ctx *c = nullptr;
        check(b200_ctx_create((int)t, N, (int)Moduli.size(), Moduli.data(), roots ? roots->data() : nullptr, device, &c));
        return Ring(std::shared_ptr<b200_ctx>(c, b200_ctx_destroy), (int)Moduli.size() - 1);
    }
    int N() const { return b200_ctx_n(ctx_.get()); }
    int Level() const { return level_; }
    int MaxLevel() const { return b200_ctx_moduli_count(ctx_.get()) - 1; }
    int ModuliChainLength() const { return b200_ctx_moduli_count(ctx_.get()); }
    Type GetType() const { return (Type)b200_ctx_ring_type(ctx_.get()); }
    Ring AtLevel(int level) const {   // ring/ring.go:194-213
        if (level < 0) throw Panic(B200_ESHAPE, "level cannot be negative");
        if (level > MaxLevel()) throw Panic(B200_ESHAPE, "level cannot be larger than max level");
        return Ring(ctx_, level);
    }
    SubRing SubRingAt(int i) const {
        uint64_t o[8];
        check(b200_ctx_subring(ctx_.get(), i, o));
        SubRing s;
        s.N = N(); s.Modulus = o[0]; s.MRedConstant = o[1]; s.BRedConstant[0] = o[2]; s.BRedConstant[1] = o[3];
        s.Mask = o[4]; s.PrimitiveRoot = o[5]; s.NInv = o[6]; s.Root = o[7];
        return s;
    }
    Poly NewPoly(int batch = 1) const { return Poly(*this, batch); }
    void Sync() const { check(b200_sync(ctx_.get())); }

    // ring/ntt.go:127-152
    void NTT(const Poly &p1, Poly &p2) const { check(b200_ntt(ctx_.get(), p1.handle(), p2.handle(), level_, 0)); }
    void NTTLazy(const Poly &p1, Poly &p2) const { check(b200_ntt(ctx_.get(), p1.handle(), p2.handle(), level_, B200_LAZY)); }
    void INTT(const Poly &p1, Poly &p2) const { check(b200_intt(ctx_.get(), p1.handle(), p2.handle(), level_, 0)); }
    void INTTLazy(const Poly &p1, Poly &p2) const { check(b200_intt(ctx_.get(), p1.handle(), p2.handle(), level_, B200_LAZY)); }
    // ring/operations.go (a representative subset by name; every op code is reachable through VecOp)
    // `scalars`: one word per limb (two per limb, scalar0 block then scalar1 block, for op codes 57..63)
    void VecOp(int op, const Poly &p1, const Poly *p2, Poly &p3, const std::vector<uint64_t> *scalars = nullptr) const {
        if (scalars && scalars->size() < (size_t)(level_ + 1) * (op >= B200_OP_ADD_SCALAR_LAZY_MULSCALAR_MONT ? 2 : 1))
            throw Panic(B200_ESHAPE, "need one scalar per limb");
        check(b200_vec_op(ctx_.get(), op, p1.handle(), p2 ? p2->handle() : nullptr, p3.handle(), level_, scalars ? scalars->data() : nullptr));
    }
    void Zero(Poly &p) const { VecOp(B200_OP_ZERO, p, nullptr, p); }                                                          // vec_ops.go:847
    void MulByVectorMontgomery(const Poly &p1, const Poly &vector, Poly &p2) const {                                           // operations.go:366
        check(b200_mul_by_vector_montgomery(ctx_.get(), p1.handle(), vector.handle(), p2.handle(), level_, 0));
    }
    void MulByVectorMontgomeryThenAddLazy(const Poly &p1, const Poly &vector, Poly &p2) const {                                // operations.go:373
        check(b200_mul_by_vector_montgomery(ctx_.get(), p1.handle(), vector.handle(), p2.handle(), level_, 1));
    }
    void Add(const Poly &a, const Poly &b, Poly &c) const { VecOp(B200_OP_ADD, a, &b, c); }
    void Sub(const Poly &a, const Poly &b, Poly &c) const { VecOp(B200_OP_SUB, a, &b, c); }
    void Reduce(const Poly &a, Poly &c) const { VecOp(B200_OP_REDUCE, a, nullptr, c); }
    void MForm(const Poly &a, Poly &c) const { VecOp(B200_OP_MFORM, a, nullptr, c); }
    void IMForm(const Poly &a, Poly &c) const { VecOp(B200_OP_IMFORM, a, nullptr, c); }
    void MulCoeffsMontgomery(const Poly &a, const Poly &b, Poly &c) const { VecOp(B200_OP_MUL_MONT, a, &b, c); }
    void MulCoeffsMontgomeryThenAdd(const Poly &a, const Poly &b, Poly &c) const { VecOp(B200_OP_MUL_MONT_ADD, a, &b, c); }
    void MulCoeffsBarrett(const Poly &a, const Poly &b, Poly &c) const { VecOp(B200_OP_MUL_BARRETT, a, &b, c); }
    // schemes/ckks/evaluator.go:821-834 (mform) / schemes/matrix_ckks/evaluator.go:163-173 (!mform): fused tensor product
    void TensorProduct(const Poly &a0, const Poly &a1, const Poly &b0, const Poly &b1, Poly &c0, Poly &c1, Poly &c2, bool mform = true) const {
        check(b200_tensor(ctx_.get(), a0.handle(), a1.handle(), b0.handle(), b1.handle(), c0.handle(), c1.handle(), c2.handle(), level_, mform));
    }
    // ring/scaling.go:130-156 and :6-19 (the Go `buff` argument is internal)
    void DivRoundByLastModulusNTT(const Poly &p0, Poly &p1) const { check(b200_div_by_last_modulus_ntt(ctx_.get(), p0.handle(), p1.handle(), level_, 1)); }
    void DivFloorByLastModulusNTT(const Poly &p0, Poly &p1) const { check(b200_div_by_last_modulus_ntt(ctx_.get(), p0.handle(), p1.handle(), level_, 0)); }
    // ring/scaling.go:21-30,113-127 (coefficient domain) and :32-91,130-192 (nbRescales divisions; `buff` is internal)
    void DivFloorByLastModulus(const Poly &p0, Poly &p1) const { check(b200_div_by_last_modulus(ctx_.get(), p0.handle(), p1.handle(), level_, 0)); }
    void DivRoundByLastModulus(const Poly &p0, Poly &p1) const { check(b200_div_by_last_modulus(ctx_.get(), p0.handle(), p1.handle(), level_, 1)); }
    void DivFloorByLastModulusMany(int nb, const Poly &p0, Poly &p1) const { check(b200_div_by_last_modulus_many(ctx_.get(), p0.handle(), p1.handle(), level_, nb, 0, 0)); }
    void DivFloorByLastModulusManyNTT(int nb, const Poly &p0, Poly &p1) const { check(b200_div_by_last_modulus_many(ctx_.get(), p0.handle(), p1.handle(), level_, nb, 0, 1)); }
    void DivRoundByLastModulusMany(int nb, const Poly &p0, Poly &p1) const { check(b200_div_by_last_modulus_many(ctx_.get(), p0.handle(), p1.handle(), level_, nb, 1, 0)); }
    void DivRoundByLastModulusManyNTT(int nb, const Poly &p0, Poly &p1) const { check(b200_div_by_last_modulus_many(ctx_.get(), p0.handle(), p1.handle(), level_, nb, 1, 1)); }
    // ckks.Evaluator.Rescale on the two polynomials of a ciphertext (schemes/ckks/evaluator.go:500-541); the caller lowers its level by nb
    void Rescale(int nb, const Poly &c0, const Poly &c1, Poly &out0, Poly &out1) const { DivRoundByLastModulusManyNTT(nb, c0, out0); DivRoundByLastModulusManyNTT(nb, c1, out1); }
    // NTT / INTT sequence on host-resident polynomials, copies overlapped with compute
    void TransformHost(const std::vector<uint64_t> &in, std::vector<uint64_t> &out, const std::vector<int> &ops) const {
        const size_t per = (size_t)ModuliChainLength() * N();
        check(b200_transform_host(ctx_.get(), in.data(), out.data(), (int)(in.size() / per), level_, ops.data(), (int)ops.size(), 0));
    }
    // a rank's share of a limb-sharded ring: transforms on limbs [limb0, limb0+nl) only
    void NTTLimbs(const Poly &p1, Poly &p2, int limb0, int nl, bool inverse = false) const {
        check(b200_ntt_limbs(ctx_.get(), p1.handle(), p2.handle(), limb0, nl, inverse, 0));
    }
    // inverse transform whose last kernel also stores into the same limbs of polynomials mapped from other GPUs
    // (Poly::IpcHandle in the owning process, OpenPeer here): the key switch's all-gather fused into its producer
    void *OpenPeer(const std::array<unsigned char, 64> &handle) const {
        void *ptr = nullptr;
        check(b200_peer_open(ctx_.get(), handle.data(), &ptr));
        return ptr;
    }
    void ClosePeer(void *ptr) const { check(b200_peer_close(ctx_.get(), ptr)); }
    void INTTLimbsGather(const Poly &p1, Poly &p2, int limb0, int nl, const std::vector<void *> &peers) const {
        check(b200_intt_limbs_gather(ctx_.get(), p1.handle(), p2.handle(), limb0, nl, 0, (int)peers.size(), peers.data()));
    }
    // fused NTT, NTT, MForm, MulCoeffsMontgomery, INTT (BASELINE config 4)
    void MulPoly(const Poly &a, const Poly &b, Poly &c) const { check(b200_polymul(ctx_.get(), a.handle(), b.handle(), c.handle(), level_)); }
    // ring/automorphism.go, ring/operations.go:278-363, ring/conjugate_invariant.go (none of the automorphisms works in place)
    static std::vector<uint64_t> AutomorphismNTTIndex(int N, uint64_t NthRoot, uint64_t GalEl) {
        std::vector<uint64_t> index(N > 0 ? N : 0);
        check(b200_automorphism_ntt_index(N, NthRoot, GalEl, index.data()));
        return index;
    }
    void AutomorphismNTT(const Poly &in, uint64_t gen, Poly &out) const { check(b200_automorphism_ntt(ctx_.get(), in.handle(), gen, out.handle(), level_, 0)); }
    void AutomorphismNTTWithIndex(const Poly &in, const std::vector<uint64_t> &index, Poly &out) const {
        check(b200_automorphism_ntt_with_index(ctx_.get(), in.handle(), index.data(), index.size(), out.handle(), level_, 0));
    }
    void AutomorphismNTTWithIndexThenAddLazy(const Poly &in, const std::vector<uint64_t> &index, Poly &out) const {
        check(b200_automorphism_ntt_with_index(ctx_.get(), in.handle(), index.data(), index.size(), out.handle(), level_, 1));
    }
    void Automorphism(const Poly &in, uint64_t gen, Poly &out) const { check(b200_automorphism(ctx_.get(), in.handle(), gen, out.handle(), level_)); }
    void Shift(const Poly &p1, int k, Poly &p2) const { check(b200_shift(ctx_.get(), p1.handle(), k, p2.handle())); }
    void MultByMonomial(const Poly &p1, int k, Poly &p2) const { check(b200_mult_by_monomial(ctx_.get(), p1.handle(), k, p2.handle(), level_)); }
    // the receiver is the conjugate-invariant ring; polyStandard lives in a ring of twice the degree over the same moduli
    void UnfoldConjugateInvariantToStandard(const Poly &polyCI, Poly &polyStandard) const {
        check(b200_unfold_ci_to_standard(ctx_.get(), polyCI.handle(), polyStandard.handle(), level_));
    }
    void FoldStandardToConjugateInvariant(const Poly &polyStandard, const std::vector<uint64_t> &permuteNTTIndexInv, Poly &polyCI) const {
        check(b200_fold_standard_to_ci(ctx_.get(), polyStandard.handle(), permuteNTTIndexInv.data(), permuteNTTIndexInv.size(), polyCI.handle(), level_));
    }
    // NumberTheoreticTransformer (ring/ntt.go:15-22) for one prime, host slices
    void SubRingNTT(int limb, const std::vector<uint64_t> &p1, std::vector<uint64_t> &p2, bool inverse = false, bool lazy = false) const {
        check(b200_subring_ntt_host(ctx_.get(), limb, p1.data(), p2.data(), p1.size() < p2.size() ? p1.size() : p2.size(), inverse, lazy ? B200_LAZY : 0));
    }
    b200_ctx *handle() const { return ctx_.get(); }

  private:
    Ring(std::shared_ptr<b200_ctx> c, int level) : ctx_(std::move(c)), level_(level) {}
    std::shared_ptr<b200_ctx> ctx_;
    int level_;
};

// ring.BasisExtender + ring.Decomposer + rlwe.Evaluator.GadgetProduct over one QP ring (Q limbs first)
class KeySwitcher {
  public:
    KeySwitcher(const Ring &ringQP, int LQ) : ring_(ringQP) {
        b200_keyswitch *k = nullptr;
        check(b200_keyswitch_create(ringQP.handle(), LQ, &k));
        ks_.reset(k, b200_keyswitch_destroy);
    }
    void ModUpQtoP(int levelQ, int levelP, const Poly &polQ, Poly &polP) const { check(b200_mod_up(ks_.get(), 0, levelQ, levelP, polQ.handle(), polP.handle())); }
    void ModUpPtoQ(int levelP, int levelQ, const Poly &polP, Poly &polQ) const { check(b200_mod_up(ks_.get(), 1, levelQ, levelP, polP.handle(), polQ.handle())); }
    void ModDownQPtoQNTT(int levelQ, int levelP, const Poly &p1, Poly &p2) const { check(b200_mod_down_ntt(ks_.get(), levelQ, levelP, p1.handle(), p2.handle())); }
    void ModDownQPtoQ(int levelQ, int levelP, const Poly &p1, Poly &p2) const { check(b200_mod_down(ks_.get(), 0, levelQ, levelP, p1.handle(), p2.handle())); }   // :223-236
    void ModDownQPtoP(int levelQ, int levelP, const Poly &p1, Poly &p2) const { check(b200_mod_down(ks_.get(), 1, levelQ, levelP, p1.handle(), p2.handle())); }   // :264-279
    void GadgetProduct(int levelQ, int levelP, const Poly &cx, const Poly &key, Poly &ct0, Poly &ct1) const {
        check(b200_gadget_product(ks_.get(), levelQ, levelP, cx.handle(), key.handle(), ct0.handle(), ct1.handle()));
    }
    // one special prime with a power-of-two base (evaluator_gadget_product.go:190-320); sizes = BaseTwoDecompositionVectorSize
    void GadgetProductBitDecomp(int levelQ, int levelP, int pw2, const std::vector<int> &sizes, const Poly &cx, const Poly &key, Poly &ct0, Poly &ct1) const {
        check(b200_gadget_product_bit_decomp(ks_.get(), levelQ, levelP, pw2, sizes.data(), cx.handle(), key.handle(), ct0.handle(), ct1.handle()));
    }
    // hoisted form (evaluator_gadget_product.go:322-478): decomp holds beta*batch polynomials [digit][b]
    void DecomposeNTT(int levelQ, int levelP, const Poly &cx, Poly &decomp) const { check(b200_decompose_ntt(ks_.get(), levelQ, levelP, cx.handle(), decomp.handle())); }
    void GadgetProductHoisted(int levelQ, int levelP, const Poly &decomp, const Poly &key, Poly &ct0, Poly &ct1) const {
        check(b200_gadget_product_hoisted(ks_.get(), levelQ, levelP, decomp.handle(), key.handle(), ct0.handle(), ct1.handle()));
    }
    // rlwe.Evaluator.Automorphism / AutomorphismHoisted (evaluator_automorphism.go:13-108), NTT domain
    void Automorphism(int levelQ, int levelP, const Poly &ct0, const Poly &ct1, uint64_t galEl, const Poly &key, Poly &out0, Poly &out1,
                      const Poly *decomp = nullptr) const {
        check(b200_automorphism_ct(ks_.get(), levelQ, levelP, ct0.handle(), ct1.handle(), decomp ? decomp->handle() : nullptr, galEl, key.handle(),
                                   out0.handle(), out1.handle()));
    }
    void MulRelin(int levelQ, int levelP, const Poly &a0, const Poly &a1, const Poly &b0, const Poly &b1, const Poly &rlk, Poly &o0, Poly &o1,
                  bool mform = true) const {
        check(b200_mul_relin(ks_.get(), levelQ, levelP, a0.handle(), a1.handle(), b0.handle(), b1.handle(), rlk.handle(), o0.handle(), o1.handle(), mform));
    }
    // rlwe.Evaluator.Relinearize (core/rlwe/evaluator_evaluationkey.go:125-153): GadgetProduct(c2, rlk) into two temporaries,
    // then out_i = c_i + tmp_i on the limbs 0..levelQ
    void Relinearize(int levelQ, int levelP, const Poly &c0, const Poly &c1, const Poly &c2, const Poly &rlk, Poly &out0, Poly &out1) const {
        Poly t0 = ring_.NewPoly(c2.Batch()), t1 = ring_.NewPoly(c2.Batch());
        GadgetProduct(levelQ, levelP, c2, rlk, t0, t1);
        const Ring rq = ring_.AtLevel(levelQ);
        rq.Add(c0, t0, out0);
        rq.Add(c1, t1, out1);
    }
    // rlwe.Evaluator.ApplyEvaluationKey for equal ring degrees (:37-117): GadgetProduct(c1, evk), out0 = c0 + tmp0, out1 = tmp1
    void ApplyEvaluationKey(int levelQ, int levelP, const Poly &c0, const Poly &c1, const Poly &evk, Poly &out0, Poly &out1) const {
        Poly t0 = ring_.NewPoly(c1.Batch());
        GadgetProduct(levelQ, levelP, c1, evk, t0, out1);
        ring_.AtLevel(levelQ).Add(c0, t0, out0);
    }

  private:
    Ring ring_;
    std::shared_ptr<b200_keyswitch> ks_;
};

inline Poly::Poly(const Ring &r, int batch) : batch_(batch) {
    check(b200_poly_alloc(r.handle(), batch, &h_));
    words_ = (size_t)batch * r.ModuliChainLength() * r.N();
    n_ = (size_t)r.N();
}

}  // namespace ring
