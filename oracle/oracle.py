"""ctypes front-end of the CPU oracle (oracle/ref_ring.c) -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  The product package must never do so (it fails loudly without its CUDA .so).

Also holds `py_*` pure-Python-int definitional functions (tier T0 of SURVEY.md 8c) used to
cross-check the C restatement itself on small sizes.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libref_ring.so")

u64 = C.c_uint64
P64 = C.POINTER(C.c_uint64)
PI = C.POINTER(C.c_int)

# op codes: identical to include/b200ring.h (B200_OP_*) and oracle/ref_ring.c
OPS = {
    "Add": 0, "AddLazy": 1, "Sub": 2, "SubLazy": 3, "Neg": 4, "Reduce": 5, "ReduceLazy": 6,
    "MForm": 7, "MFormLazy": 8, "IMForm": 9,
    "MulCoeffsBarrett": 16, "MulCoeffsBarrettLazy": 17, "MulCoeffsBarrettThenAdd": 18,
    "MulCoeffsBarrettThenAddLazy": 19,
    "MulCoeffsMontgomery": 32, "MulCoeffsMontgomeryLazy": 33, "MulCoeffsMontgomeryThenAdd": 34,
    "MulCoeffsMontgomeryThenAddLazy": 35, "MulCoeffsMontgomeryLazyThenAddLazy": 36,
    "MulCoeffsMontgomeryThenSub": 37, "MulCoeffsMontgomeryThenSubLazy": 38,
    "MulCoeffsMontgomeryLazyThenSubLazy": 39, "MulCoeffsMontgomeryLazyThenNeg": 40,
    "MulScalarMontgomery": 48, "MulScalarMontgomeryLazy": 49, "MulScalarMontgomeryThenAdd": 50,
    "AddScalar": 51, "SubScalar": 52, "SubThenMulScalarMontgomeryTwoModulus": 53,
    "ZeroVec": 10, "MulCoeffsLazy": 20, "MulCoeffsLazyThenAddLazy": 21,
    "AddLazyThenMulScalarMontgomery": 54, "AddScalarLazy": 55, "AddScalarLazyThenNegTwoModulusLazy": 56,
    "AddScalarLazyThenMulScalarMontgomery": 57, "MulScalarMontgomeryThenAddScalar": 58, "MaskVec": 59,
    "AddDoubleScalar": 60, "SubDoubleScalar": 61, "MulDoubleScalarMontgomery": 62, "MulDoubleScalarMontgomeryThenAdd": 63,
}


def build(force=False):
    """Compile oracle/ref_ring.c -> oracle/_build/libref_ring.so (gcc only, a few seconds)."""
    src = os.path.join(_HERE, "ref_ring.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        sig = {
            "ref_gen_mred_constant": (u64, [u64]),
            "ref_gen_bred_constant": (None, [u64, P64]),
            "ref_mform": (u64, [u64, u64, P64]), "ref_mform_lazy": (u64, [u64, u64, P64]),
            "ref_imform": (u64, [u64, u64, u64]), "ref_imform_lazy": (u64, [u64, u64, u64]),
            "ref_mred": (u64, [u64, u64, u64, u64]), "ref_mred_lazy": (u64, [u64, u64, u64, u64]),
            "ref_bred": (u64, [u64, u64, u64, P64]), "ref_bred_lazy": (u64, [u64, u64, u64, P64]),
            "ref_bred_add": (u64, [u64, u64, P64]), "ref_bred_add_lazy": (u64, [u64, u64, P64]),
            "ref_cred": (u64, [u64, u64]), "ref_modexp": (u64, [u64, u64, u64]),
            "ref_is_prime": (C.c_int, [u64]), "ref_unique_factors": (C.c_int, [u64, P64]),
            "ref_primitive_root": (u64, [u64]), "ref_find_3n_primitive_root": (u64, [u64, u64]),
            "ref_find_3n_rns_primes": (C.c_int, [u64, C.c_int, C.c_int, C.c_int, P64]),
            "ref_pow2_new": (C.c_void_p, [C.c_int, u64]), "ref_pow2_free": (None, [C.c_void_p]),
            "ref_pow2_constants": (None, [C.c_void_p, P64]),
            "ref_pow2_tables": (None, [C.c_void_p, P64, P64]),
            "ref_pow2_ntt": (None, [C.c_void_p, P64, P64, C.c_int]),
            "ref_pow2_intt": (None, [C.c_void_p, P64, P64, C.c_int]),
            "ref_ci_new": (C.c_void_p, [C.c_int, u64]), "ref_ci_free": (None, [C.c_void_p]),
            "ref_ci_constants": (None, [C.c_void_p, P64]),
            "ref_ci_tables": (None, [C.c_void_p, P64, P64]),
            "ref_ci_ntt": (None, [C.c_void_p, P64, P64, C.c_int]),
            "ref_ci_intt": (None, [C.c_void_p, P64, P64, C.c_int]),
            "ref_automorphism_ntt_index": (C.c_int, [C.c_int, u64, u64, P64]),
            "ref_automorphism": (None, [C.c_int, u64, C.c_int, u64, P64, P64]),
            "ref_vec_op": (C.c_int, [C.c_int, C.c_int, u64, P64, P64, P64, u64]),
            "ref_vec_op2": (C.c_int, [C.c_int, C.c_int, u64, P64, P64, P64, u64, u64]),
            "ref_3n_exponents": (C.c_int, [C.c_int, PI]),
            "ref_3n_forward_horner": (None, [C.c_int, u64, u64, P64, P64]),
            "ref_3n_forward_horner_at": (u64, [C.c_int, u64, u64, P64, C.c_int]),
            "ref_3n_backward_gauss": (C.c_int, [C.c_int, u64, u64, P64, P64]),
            "ref_3n_schoolbook": (None, [C.c_int, u64, P64, P64, P64]),
            "ref_negacyclic_schoolbook": (None, [C.c_int, u64, P64, P64, P64]),
            "ref_3n_new": (C.c_void_p, [C.c_int, u64, u64]), "ref_3n_free": (None, [C.c_void_p]),
            "ref_3n_leaf_exponents": (None, [C.c_void_p, PI]), "ref_3n_go2fast": (None, [C.c_void_p, PI]),
            "ref_3n_zetas": (None, [C.c_void_p, P64]),
            "ref_3n_forward": (None, [C.c_void_p, P64, P64, C.c_int]),
            "ref_3n_backward": (None, [C.c_void_p, P64, P64, C.c_int]),
            "ref_div_by_last_modulus_ntt": (None, [C.c_int, C.POINTER(C.c_void_p), C.c_int, C.c_int, P64, P64, C.c_int]),
            "ref_gen_modup_constants": (C.c_void_p, [P64, C.c_int, P64, C.c_int]), "ref_muc_free": (None, [C.c_void_p]),
            "ref_basis_extend": (None, [C.c_void_p, P64, P64, PI, C.c_int, C.c_int, P64]),
            "ref_decompose_single": (None, [u64, P64, P64, C.c_int, C.c_int, P64]),
            "ref_fill_uniform": (None, [u64, u64, u64, u64, C.c_int, P64]),
            "ref_num_threads": (C.c_int, []),
            "ref_3n_batch": (None, [C.POINTER(C.c_void_p), C.c_int, P64, C.c_int, C.c_int, C.c_int]),
            "ref_pow2_batch": (None, [C.POINTER(C.c_void_p), C.c_int, P64, C.c_int, C.c_int, C.c_int]),
        }
        for name, (res, args) in sig.items():
            f = getattr(L, name)
            f.restype, f.argtypes = res, args
        _lib = L
    return _lib


def _p(a):
    assert a.dtype == np.uint64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(P64)


def _arr(x):
    return np.ascontiguousarray(np.asarray(x, dtype=np.uint64))


def bred_constant(q):
    out = np.zeros(2, dtype=np.uint64)
    lib().ref_gen_bred_constant(q, _p(out))
    return out


def fill_uniform(seed, poly, limb, q, N):
    out = np.empty(N, dtype=np.uint64)
    lib().ref_fill_uniform(seed, poly, limb, q, N, _p(out))
    return out


def find_3n_rns_primes(N, bitsize, count, budget=1 << 20):
    out = np.zeros(count, dtype=np.uint64)
    n = lib().ref_find_3n_rns_primes(N, bitsize, count, budget, _p(out))
    if n < count:
        raise ValueError("could not find enough 3N-friendly primes (found %d/%d)" % (n, count))
    return [int(x) for x in out]


def vec_op(name, q, p1, p2=None, p3=None, scalar=0, scalar1=0):
    """One limb of ring/vec_ops.go; returns the new p3.  scalar1: second scalar of the two-scalar operators."""
    p1 = _arr(p1)
    N = len(p1)
    p2a = _arr(p2) if p2 is not None else None
    out = _arr(p3).copy() if p3 is not None else np.zeros(N, dtype=np.uint64)
    rc = lib().ref_vec_op2(OPS[name], N, q, _p(p1), _p(p2a) if p2a is not None else None, _p(out), scalar, scalar1)
    if rc:
        raise ValueError(name)
    return out


class Pow2SubRing:
    """ring.SubRing with the Standard transformer (subring.go:74-214 + ntt.go)."""

    def __init__(self, N, q):
        self.N, self.q = N, q
        self.h = lib().ref_pow2_new(N, q)
        if not self.h:
            raise ValueError("invalid power-of-two subring N=%d q=%d" % (N, q))
        c = np.zeros(6, dtype=np.uint64)
        lib().ref_pow2_constants(self.h, _p(c))
        self.MRedConstant, self.BRedConstant = int(c[0]), (int(c[1]), int(c[2]))
        self.PrimitiveRoot, self.NInv = int(c[3]), int(c[4])

    def tables(self):
        rf = np.empty(self.N, dtype=np.uint64)
        rb = np.empty(self.N, dtype=np.uint64)
        lib().ref_pow2_tables(self.h, _p(rf), _p(rb))
        return rf, rb

    def NTT(self, p, lazy=False):
        p = _arr(p)
        out = np.empty_like(p)
        lib().ref_pow2_ntt(self.h, _p(p), _p(out), int(lazy))
        return out

    def INTT(self, p, lazy=False):
        p = _arr(p)
        out = np.empty_like(p)
        lib().ref_pow2_intt(self.h, _p(p), _p(out), int(lazy))
        return out

    def __del__(self):
        if getattr(self, "h", None) and _lib is not None:
            _lib.ref_pow2_free(self.h)
            self.h = None


def automorphism_ntt_index(N, nthroot, galel):
    """ring.AutomorphismNTTIndex (ring/automorphism.go:12-35)."""
    out = np.empty(N, dtype=np.uint64)
    if lib().ref_automorphism_ntt_index(N, nthroot, galel % (1 << 64), _p(out)):
        raise ValueError("N and NthRoot must be powers of two")
    return out


def automorphism(N, q, gen, p, conjugate_invariant=False):
    """Ring.Automorphism on one limb, coefficient domain (ring/automorphism.go:113-176).  Not in place."""
    p = _arr(p)
    out = np.zeros_like(p)
    lib().ref_automorphism(N, q, int(conjugate_invariant), gen % (1 << 64), _p(p), _p(out))
    return out


class ConjugateInvariantSubRing:
    """ring.SubRing with the ConjugateInvariant transformer (NthRoot = 4N; ring/ntt.go:80-124,716-786,1088-1158)."""

    def __init__(self, N, q):
        self.N, self.q = N, q
        self.h = lib().ref_ci_new(N, q)
        if not self.h:
            raise ValueError("invalid conjugate-invariant subring N=%d q=%d" % (N, q))
        c = np.zeros(6, dtype=np.uint64)
        lib().ref_ci_constants(self.h, _p(c))
        self.MRedConstant, self.BRedConstant = int(c[0]), (int(c[1]), int(c[2]))
        self.PrimitiveRoot, self.NInv = int(c[3]), int(c[4])

    def tables(self):
        rf = np.empty(2 * self.N, dtype=np.uint64)
        rb = np.empty(2 * self.N, dtype=np.uint64)
        lib().ref_ci_tables(self.h, _p(rf), _p(rb))
        return rf, rb

    def NTT(self, p, lazy=False):
        p = _arr(p)
        out = np.empty_like(p)
        lib().ref_ci_ntt(self.h, _p(p), _p(out), int(lazy))
        return out

    def INTT(self, p, lazy=False):
        p = _arr(p)
        out = np.empty_like(p)
        lib().ref_ci_intt(self.h, _p(p), _p(out), int(lazy))
        return out

    def __del__(self):
        if getattr(self, "h", None) and _lib is not None:
            _lib.ref_ci_free(self.h)
            self.h = None


class ThreeNSubRing:
    """3N ring Z_q[X]/(X^N - X^(N/2) + 1): fast transform (T2) with the Go slot order."""

    def __init__(self, N, q, omega=None):
        if omega is None:
            omega = lib().ref_find_3n_primitive_root(q, 3 * N)  # subring.go:255-290 (deterministic)
        self.N, self.q, self.omega = N, q, int(omega)
        self.h = lib().ref_3n_new(N, q, self.omega)
        if not self.h:
            raise ValueError("invalid 3N subring N=%d q=%d omega=%d" % (N, q, self.omega))

    def leaf_exponents(self):
        out = np.empty(self.N, dtype=np.int32)
        lib().ref_3n_leaf_exponents(self.h, out.ctypes.data_as(PI))
        return out

    def go2fast(self):
        out = np.empty(self.N, dtype=np.int32)
        lib().ref_3n_go2fast(self.h, out.ctypes.data_as(PI))
        return out

    def zetas(self):
        out = np.empty(self.N, dtype=np.uint64)
        lib().ref_3n_zetas(self.h, _p(out))
        return out

    def NTT(self, p, order=0):
        p = _arr(p)
        out = np.empty_like(p)
        lib().ref_3n_forward(self.h, _p(p), _p(out), order)
        return out

    def INTT(self, p, order=0):
        p = _arr(p)
        out = np.empty_like(p)
        lib().ref_3n_backward(self.h, _p(p), _p(out), order)
        return out

    def NTT_horner(self, p):
        p = _arr(p)
        out = np.empty_like(p)
        lib().ref_3n_forward_horner(self.N, self.q, self.omega, _p(p), _p(out))
        return out

    def horner_at(self, p, exponent):
        return int(lib().ref_3n_forward_horner_at(self.N, self.q, self.omega, _p(_arr(p)), int(exponent)))

    def INTT_gauss(self, p):
        p = _arr(p)
        out = np.empty_like(p)
        if lib().ref_3n_backward_gauss(self.N, self.q, self.omega, _p(p), _p(out)):
            raise ValueError("singular Vandermonde")
        return out

    def __del__(self):
        if getattr(self, "h", None) and _lib is not None:
            _lib.ref_3n_free(self.h)
            self.h = None


def schoolbook_3n(N, q, a, b):
    a, b = _arr(a), _arr(b)
    c = np.empty(N, dtype=np.uint64)
    lib().ref_3n_schoolbook(N, q, _p(a), _p(b), _p(c))
    return c


def schoolbook_negacyclic(N, q, a, b):
    a, b = _arr(a), _arr(b)
    c = np.empty(N, dtype=np.uint64)
    lib().ref_negacyclic_schoolbook(N, q, _p(a), _p(b), _p(c))
    return c


def div_by_last_modulus_ntt(rings, p0, round=True):
    """Ring.DivRoundByLastModulusNTT / DivFloorByLastModulusNTT (ring/scaling.go:6-19,130-156) on one polynomial
    p0[nl][N] in the NTT domain; returns p1[nl-1][N]."""
    p0 = _arr(p0)
    nl, N = p0.shape
    hs = (C.c_void_p * nl)(*[r.h for r in rings[:nl]])
    out = np.empty((nl - 1, N), dtype=np.uint64)
    lib().ref_div_by_last_modulus_ntt(1 if isinstance(rings[0], ThreeNSubRing) else 0, hs, nl, N, _p(p0), _p(out), int(round))
    return out


def rescale_constants(moduli, level):
    """Ring.RescaleConstants[level-1] (ring/ring.go:361-379): MForm(q_i - q_level^-1 mod q_i) for i < level."""
    ql = int(moduli[level])
    return [(int(q) - pow(ql, int(q) - 2, int(q))) * (1 << 64) % int(q) for q in moduli[:level]]


def div_by_last_modulus(moduli, p0, round=True):
    """Ring.DivRoundByLastModulus / DivFloorByLastModulus (ring/scaling.go:21-30,113-127) on one coefficient-domain
    polynomial p0[nl][N], statement by statement with the vector functions of ring/vec_ops.go; returns p1[nl-1][N]
    (the Go DivRound also overwrites p0 with its intermediates: the copies made here play that role)."""
    p0 = _arr(p0).copy()
    level = p0.shape[0] - 1
    rc = rescale_constants(moduli, level)
    out = np.empty((level, p0.shape[1]), dtype=np.uint64)
    if not round:
        for i in range(level):   # :24-26
            out[i] = vec_op("SubThenMulScalarMontgomeryTwoModulus", int(moduli[i]), p0[level], p0[i], scalar=rc[i])
        return out
    ql = int(moduli[level])
    phalf = (ql - 1) >> 1
    p0[level] = vec_op("AddScalar", ql, p0[level], scalar=phalf)                                   # :120
    for i in range(level):
        q = int(moduli[i])
        p0[i] = vec_op("AddScalarLazyThenNegTwoModulusLazy", q, p0[i], scalar=q - phalf % q)       # :123 (BRedAdd(pHalf) = pHalf mod q)
        out[i] = vec_op("AddLazyThenMulScalarMontgomery", q, p0[level], p0[i], scalar=rc[i])       # :124
    return out


def div_by_last_modulus_many(rings, moduli, p0, nb, round=True, ntt=False):
    """Ring.Div{Floor,Round}ByLastModulusMany[NTT] (ring/scaling.go:32-91,130-192) on one polynomial p0[nl][N];
    returns the nl - nb limbs of the result.  `rings`: the oracle SubRings (used by the NTT forms only)."""
    p0 = _arr(p0)
    nl = p0.shape[0]
    if nb == 0:
        return p0.copy()
    if ntt and round and nb == 1:
        return div_by_last_modulus_ntt(rings, p0, round=True)                                      # :150-152
    buff = np.stack([rings[i].INTT(p0[i]) for i in range(nl)]) if ntt else p0.copy()               # :40 / :140
    for k in range(nb):
        buff = div_by_last_modulus(moduli[: nl - k], buff, round=round)
    return np.stack([rings[i].NTT(buff[i]) for i in range(nl - nb)]) if ntt else buff


def totatives(threeN):
    out = np.empty(threeN, dtype=np.int32)
    n = lib().ref_3n_exponents(threeN, out.ctypes.data_as(PI))
    return out[:n].copy()


def batch_transform(rings, data, inverse=False, threads=0):
    """In-place transform of data[units][N]; unit u uses rings[u % len(rings)] (timed CPU baseline)."""
    assert data.dtype == np.uint64 and data.flags["C_CONTIGUOUS"] and data.ndim == 2
    hs = (C.c_void_p * len(rings))(*[r.h for r in rings])
    fn = lib().ref_3n_batch if isinstance(rings[0], ThreeNSubRing) else lib().ref_pow2_batch
    fn(hs, len(rings), _p(data), data.shape[0], int(inverse), threads)


def num_threads():
    return lib().ref_num_threads()


# ---------------------------------------------------------------------------------------------
# T0: pure-Python-int definitional functions (small sizes only)
# ---------------------------------------------------------------------------------------------
def py_bitrev(x, bits):
    return int(format(x, "0%db" % bits)[::-1], 2) if bits else 0


def py_pow2_ntt(N, q, g, coeffs):
    """out[i] = f(psi^(2 brev(i)+1)), psi = g^((q-1)/2N): the definition ring/ntt.go computes."""
    psi = pow(g, (q - 1) // (2 * N), q)
    logn = N.bit_length() - 1
    out = []
    for i in range(N):
        x = pow(psi, 2 * py_bitrev(i, logn) + 1, q)
        acc = 0
        for c in reversed(coeffs):
            acc = (acc * x + int(c)) % q
        out.append(acc)
    return out


def py_3n_ntt(N, q, omega, coeffs):
    """ring/ntt_3n.go:82-109 with Python ints."""
    from math import gcd
    E = [e for e in range(1, 3 * N) if gcd(e, 3 * N) == 1]
    out = []
    for e in E:
        x = pow(omega, e, q)
        acc = 0
        for c in reversed(coeffs):
            acc = (acc * x + int(c)) % q
        out.append(acc)
    return out


# ---------------------------------------------------------------------------------------------
# RNS basis extension, decomposition and gadget product: ring/basis_extension.go,
# core/rlwe/evaluator_gadget_product.go.  Constants use Python big ints; the per-coefficient loops are C.
# ---------------------------------------------------------------------------------------------
class ModUpConstants:
    """ring.GenModUpConstants(Q, P) (basis_extension.go:95-166)."""

    def __init__(self, Q, P):
        self.Q, self.P = [int(q) for q in Q], [int(p) for p in P]
        qa, pa = _arr(self.Q), _arr(self.P)
        self.h = lib().ref_gen_modup_constants(_p(qa), len(self.Q), _p(pa), len(self.P))

    def extend(self, p1, targets=None, half=None):
        """multSum of reconstructRNS[Centered] for the given target indices: lazy words, [len(targets)][N]."""
        p1 = _arr(p1)
        ns, N = p1.shape
        assert ns == len(self.Q)
        targets = list(range(len(self.P))) if targets is None else list(targets)
        t = np.ascontiguousarray(np.asarray(targets, dtype=np.int32))
        out = np.empty((len(targets), N), dtype=np.uint64)
        h = _arr(half) if half is not None else None
        lib().ref_basis_extend(self.h, _p(p1), _p(h) if h is not None else None, t.ctypes.data_as(PI), len(targets), N, _p(out))
        return out

    def __del__(self):
        if getattr(self, "h", None) and _lib is not None:
            _lib.ref_muc_free(self.h)
            self.h = None


def _prod(xs):
    r = 1
    for x in xs:
        r *= int(x)
    return r


class BasisExtender:
    """ring.BasisExtender (basis_extension.go:12-73,184-279) over moduli chains Q and P."""

    def __init__(self, Q, P):
        self.Q, self.P = [int(q) for q in Q], [int(p) for p in P]
        self.QtoP = [ModUpConstants(self.Q[: i + 1], self.P) for i in range(len(self.Q))]
        self.PtoQ = [ModUpConstants(self.P[: i + 1], self.Q) for i in range(len(self.P))]
        rinv = lambda q: pow(2**64, 1, q)
        # genmodDownConstants (:24-49): Montgomery form of (p_0 ... p_j)^-1 mod q_i
        self.modDownPtoQ = [[(pow(_prod(self.P[: j + 1]), -1, q) * rinv(q)) % q for q in self.Q] for j in range(len(self.P))]
        self.modDownQtoP = [[(pow(_prod(self.Q[: i + 1]), -1, p) * rinv(p)) % p for p in self.P] for i in range(len(self.Q))]   # :61

    def ModUpQtoP(self, levelQ, levelP, polQ):
        """:184-198  polQ [levelQ+1][N] -> polP [levelP+1][N] (lazy words exactly as the Go code leaves them)."""
        QHalf = _prod(self.Q[: levelQ + 1]) >> 1
        buff = np.stack([vec_op("AddScalar", q, polQ[i], scalar=QHalf % q) for i, q in enumerate(self.Q[: levelQ + 1])])
        ext = self.QtoP[levelQ].extend(buff, range(levelP + 1))
        return np.stack([vec_op("SubScalar", p, ext[j], scalar=QHalf % p) for j, p in enumerate(self.P[: levelP + 1])])

    def ModUpPtoQ(self, levelP, levelQ, polP):
        """:203-217"""
        PHalf = _prod(self.P[: levelP + 1]) >> 1
        buff = np.stack([vec_op("AddScalar", p, polP[j], scalar=PHalf % p) for j, p in enumerate(self.P[: levelP + 1])])
        ext = self.PtoQ[levelP].extend(buff, range(levelQ + 1))
        return np.stack([vec_op("SubScalar", q, ext[i], scalar=PHalf % q) for i, q in enumerate(self.Q[: levelQ + 1])])

    def ModDownQPtoQ(self, levelQ, levelP, p1Q, p1P):
        """:223-236  (coefficient domain) -> p2Q [levelQ+1][N], canonical."""
        buffQ = self.ModUpPtoQ(levelP, levelQ, p1P)
        return np.stack([vec_op("SubThenMulScalarMontgomeryTwoModulus", q, buffQ[i], p1Q[i], scalar=q - self.modDownPtoQ[levelP][i])
                         for i, q in enumerate(self.Q[: levelQ + 1])])

    def ModDownQPtoP(self, levelQ, levelP, p1Q, p1P):
        """:264-279  (coefficient domain) -> p2P [levelP+1][N], canonical."""
        buffP = self.ModUpQtoP(levelQ, levelP, p1Q)
        return np.stack([vec_op("SubThenMulScalarMontgomeryTwoModulus", p, buffP[j], p1P[j], scalar=p - self.modDownQtoP[levelQ][j])
                         for j, p in enumerate(self.P[: levelP + 1])])

    def ModDownQPtoQNTT(self, levelQ, levelP, p1Q, p1P, ringsQ, ringsP):
        """:241-258  (inputs NTT domain) -> p2Q [levelQ+1][N], canonical."""
        buffP = np.stack([ringsP[j].INTT(p1P[j]) for j in range(levelP + 1)])
        buffQ = self.ModUpPtoQ(levelP, levelQ, buffP)
        out = []
        for i, q in enumerate(self.Q[: levelQ + 1]):
            b = ringsQ[i].NTT(buffQ[i] % np.uint64(q))
            out.append(vec_op("SubThenMulScalarMontgomeryTwoModulus", q, b, p1Q[i], scalar=q - self.modDownPtoQ[levelP][i]))
        return np.stack(out)


def decompose_and_split(Q, P, levelQ, levelP, nbPi, digit, p0Q):
    """ring.Decomposer.DecomposeAndSplit (basis_extension.go:381-502).  Returns (p1Q [levelQ+1][N], p1P [levelP+1][N]);
    the limbs of the digit itself are returned as zeros (the caller overwrites them, rlwe/evaluator_gadget_product.go:466-472)."""
    Q, P = [int(q) for q in Q], [int(p) for p in P]
    p0Q = _arr(p0Q)
    N = p0Q.shape[1]
    st = digit * nbPi
    decompLvl = nbPi - 2 if levelQ > nbPi * (digit + 1) - 1 else (levelQ % nbPi) - 1
    outQ = np.zeros((levelQ + 1, N), dtype=np.uint64)
    outP = np.zeros((levelP + 1, N), dtype=np.uint64)
    if decompLvl < 0:
        T = _arr(Q[: levelQ + 1] + P[: levelP + 1])
        tmp = np.empty((len(T), N), dtype=np.uint64)
        lib().ref_decompose_single(Q[st], _p(np.ascontiguousarray(p0Q[st])), _p(T), len(T), N, _p(tmp))
        outQ[:], outP[:] = tmp[: levelQ + 1], tmp[levelQ + 1:]
        return outQ, outP
    ed = min(st + nbPi, levelQ + 1)
    # NewDecomposer (:319-377): constants over Qi = Q[st : st+decompLvl+2], targets = Q (whole chain) + P[:nbPi]
    muc = ModUpConstants(Q[st: st + decompLvl + 2], Q + P[:nbPi])
    QBig = _prod(Q[st:ed])
    QHalf = QBig >> 1
    half = [QHalf % q for q in Q[st:ed]]
    tq = [j for j in range(levelQ + 1) if not (st <= j < ed)]
    tp = [len(Q) + j for j in range(levelP + 1)]
    ext = muc.extend(np.ascontiguousarray(p0Q[st:ed]), tq + tp, half=half)
    for k, j in enumerate(tq):
        outQ[j] = vec_op("SubScalar", Q[j], ext[k], scalar=QHalf % Q[j])
    for k, j in enumerate(range(levelP + 1)):
        outP[j] = vec_op("SubScalar", P[j], ext[len(tq) + k], scalar=QHalf % P[j])
    return outQ, outP


def gadget_product(ringsQ, ringsP, levelQ, levelP, cx_ntt, key):
    """rlwe.Evaluator.GadgetProduct for NTT-domain operands
    (core/rlwe/evaluator_gadget_product.go:16-30, 33-45, 123-188, 455-478).  With levelP == 0 this is also the result of
    gadgetProductSinglePAndBitDecompLazy with mask == 0 (:190-320): one-limb digits through DecomposeAndSplit(nbPi = 1).
    cx_ntt: [levelQ+1][N]; key[i][c]: (Q part [LQ][N], P part [LP][N]) for digit i, component c, NTT + Montgomery form.
    Returns (ct0, ct1), each [levelQ+1][N], canonical."""
    Q = [r.q for r in ringsQ]
    P = [r.q for r in ringsP]
    nbPi = levelP + 1
    beta = (levelQ + levelP + 1) // (levelP + 1)
    cx_inv = np.stack([ringsQ[i].INTT(cx_ntt[i]) for i in range(levelQ + 1)])
    N = cx_inv.shape[1]
    accQ = [np.zeros((levelQ + 1, N), dtype=np.uint64) for _ in range(2)]
    accP = [np.zeros((levelP + 1, N), dtype=np.uint64) for _ in range(2)]
    for d in range(beta):
        c2Q, c2P = decompose_and_split(Q, P, levelQ, levelP, nbPi, d, cx_inv)
        st, ed = d * nbPi, d * nbPi + nbPi
        for x in range(levelQ + 1):
            c2Q[x] = cx_ntt[x] if st <= x < ed else ringsQ[x].NTT(c2Q[x] % np.uint64(Q[x]))
        for j in range(levelP + 1):
            c2P[j] = ringsP[j].NTT(c2P[j] % np.uint64(P[j]))
        for c in range(2):
            kq, kp = key[d][c]
            for x in range(levelQ + 1):   # MulCoeffsMontgomeryLazy[ThenAddLazy] + Reduce: canonical sum
                accQ[c][x] = vec_op("Add", Q[x], accQ[c][x], vec_op("MulCoeffsMontgomery", Q[x], kq[x], c2Q[x]))
            for j in range(levelP + 1):
                accP[c][j] = vec_op("Add", P[j], accP[c][j], vec_op("MulCoeffsMontgomery", P[j], kp[j], c2P[j]))
    be = BasisExtender(Q, P)
    return tuple(be.ModDownQPtoQNTT(levelQ, levelP, accQ[c], accP[c], ringsQ, ringsP) for c in range(2))


def decompose_ntt(ringsQ, ringsP, levelQ, levelP, cx_ntt):
    """rlwe.Evaluator.DecomposeNTT for an NTT-domain input (evaluator_gadget_product.go:430-478): per digit the pair
    (c2Q [levelQ+1][N], c2P [levelP+1][N]), NTT domain, canonical, the digit's own limbs copied from cx_ntt."""
    Q = [r.q for r in ringsQ]
    P = [r.q for r in ringsP]
    nbPi = levelP + 1
    beta = (levelQ + levelP + 1) // (levelP + 1)
    cx_inv = np.stack([ringsQ[i].INTT(cx_ntt[i]) for i in range(levelQ + 1)])
    out = []
    for d in range(beta):
        c2Q, c2P = decompose_and_split(Q, P, levelQ, levelP, nbPi, d, cx_inv)
        st, ed = d * nbPi, d * nbPi + nbPi
        for x in range(levelQ + 1):
            c2Q[x] = cx_ntt[x] if st <= x < ed else ringsQ[x].NTT(c2Q[x] % np.uint64(Q[x]))
        for j in range(levelP + 1):
            c2P[j] = ringsP[j].NTT(c2P[j] % np.uint64(P[j]))
        out.append((c2Q, c2P))
    return out


def gadget_product_bit_decomp(ringsQ, ringsP, levelQ, pw2, sizes, cx_ntt, key):
    """rlwe.Evaluator.GadgetProduct for one special prime and a power-of-two base pw2 > 0:
    gadgetProductSinglePAndBitDecompLazy with mask != 0 (core/rlwe/evaluator_gadget_product.go:190-320) + ModDown (:33-45).
    key[i][j][c]: (Q part [LQ][N], P part [1][N]) for limb i, window j < sizes[i], component c.  Canonical outputs."""
    Q = [r.q for r in ringsQ]
    P = [r.q for r in ringsP]
    levelP = 0
    mask = np.uint64((1 << pw2) - 1)
    cx_inv = np.stack([ringsQ[i].INTT(cx_ntt[i]) for i in range(levelQ + 1)])
    N = cx_inv.shape[1]
    accQ = [np.zeros((levelQ + 1, N), dtype=np.uint64) for _ in range(2)]
    accP = [np.zeros((levelP + 1, N), dtype=np.uint64) for _ in range(2)]
    for i in range(levelQ + 1):
        for j in range(sizes[i]):
            cw = (cx_inv[i] >> np.uint64(j * pw2)) & mask                     # ring.MaskVec (vec_ops.go:870-890)
            for c in range(2):
                kq, kp = key[i][j][c]
                for x in range(levelQ + 1):
                    cwn = ringsQ[x].NTT(cw % np.uint64(Q[x]))
                    accQ[c][x] = vec_op("Add", Q[x], accQ[c][x], vec_op("MulCoeffsMontgomery", Q[x], kq[x], cwn))
                for u in range(levelP + 1):
                    cwn = ringsP[u].NTT(cw % np.uint64(P[u]))
                    accP[c][u] = vec_op("Add", P[u], accP[c][u], vec_op("MulCoeffsMontgomery", P[u], kp[u], cwn))
    be = BasisExtender(Q, P)
    return tuple(be.ModDownQPtoQNTT(levelQ, levelP, accQ[c], accP[c], ringsQ, ringsP) for c in range(2))


# ---- wire format -------------------------------------------------------------------
def poly_marshal(coeffs):
    """ring.Poly.MarshalBinary (ring/poly.go:97-160) = structs.Matrix[uint64].WriteTo (utils/structs/matrix.go:76-104)
    over structs.Vector[uint64].WriteTo (utils/structs/vector.go:82-102): u64 row count, then per row a u64 length and
    the words, little-endian (utils/buffer/writer.go:325).  coeffs: list of 1-D uint64 rows (rows may be ragged)."""
    out = [len(coeffs).to_bytes(8, "little")]
    for row in coeffs:
        row = np.asarray(row, dtype="<u8")
        out.append(len(row).to_bytes(8, "little"))
        out.append(row.tobytes())
    return b"".join(out)


def poly_unmarshal(data):
    """ring.Poly.UnmarshalBinary (matrix.go:119-153, vector.go ReadFrom): returns (rows, bytes consumed)."""
    rows = int.from_bytes(data[:8], "little")
    off, out = 8, []
    for _ in range(rows):
        n = int.from_bytes(data[off:off + 8], "little")
        off += 8
        if len(data) < off + 8 * n:
            raise ValueError("truncated input")
        out.append(np.frombuffer(data, dtype="<u8", count=n, offset=off).astype(np.uint64))
        off += 8 * n
    return out, off


def unique_factors(m):
    """utils/factorization.GetFactors (factorization.go:19-86): unique prime factors, ascending."""
    buf = np.zeros(64, dtype=np.uint64)
    n = lib().ref_unique_factors(m, _p(buf))
    return [int(x) for x in buf[:n]]


def ring_literal_json(N, moduli, nthroot_over_n, roots=None):
    """Ring.MarshalJSON (ring/ring.go:565-589) of subRingParametersLiteral (ring/subring.go:333-360) in Go's
    encoding/json form (declaration-order keys, no whitespace).  Type is 0 for Standard and for 3N sub-rings
    (subring.go:114-123); LogN = bits.Len64(N-1); PrimitiveRoot = smallest primitive root g >= 3 (subring.go:218-251),
    or omega = g^((q-1)/3N) for a 3N ring (subring.go:165, 255-290) unless given."""
    import json
    lit = []
    for i, q in enumerate(moduli):
        if roots is not None:
            root = int(roots[i])
        elif nthroot_over_n == 3:
            root = int(lib().ref_find_3n_primitive_root(q, 3 * N))
        else:
            root = int(lib().ref_primitive_root(q))
        lit.append({"Type": 0, "LogN": (N - 1).bit_length(), "NthRoot": nthroot_over_n, "Modulus": int(q),
                    "Factors": unique_factors(q - 1), "PrimitiveRoot": root})
    return json.dumps(lit, separators=(",", ":")).encode()
