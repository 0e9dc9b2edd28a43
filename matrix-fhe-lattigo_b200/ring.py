"""Host-side mirror of the reference's `ring` package surface for the hot path, over the C ABI.

Same names, argument meaning and error behaviour as the Go code so that tests read like the
reference's own (ring/ring.go, ring/operations.go, ring/ntt.go, ring/poly.go, ring/subring_ops.go):
constructors raise ValueError where Go returns an `error`; hot functions raise B200Error(ESHAPE)
where Go panics.  A Poly here is a *batch* of device-resident polynomials ([batch][L][N] uint64);
batch=1 is exactly a ring.Poly.
"""
import ctypes as C

import numpy as np

from . import capi
from . import scalar as _scalar
from .capi import B200Error, check, lib


class Type(int):
    """ring.Type (ring/ring.go:27-34)."""

    def __str__(self):
        return {0: "Standard", 1: "ConjugateInvariant", 2: "Matrix"}.get(int(self), "Invalid")


Standard, ConjugateInvariant, Matrix = Type(0), Type(1), Type(2)

_OPS = {
    "Add": 0, "AddLazy": 1, "Sub": 2, "SubLazy": 3, "Neg": 4, "Reduce": 5, "ReduceLazy": 6,
    "MForm": 7, "MFormLazy": 8, "IMForm": 9,
    "MulCoeffsBarrett": 16, "MulCoeffsBarrettLazy": 17, "MulCoeffsBarrettThenAdd": 18,
    "MulCoeffsBarrettThenAddLazy": 19,
    "MulCoeffsMontgomery": 32, "MulCoeffsMontgomeryLazy": 33, "MulCoeffsMontgomeryThenAdd": 34,
    "MulCoeffsMontgomeryThenAddLazy": 35, "MulCoeffsMontgomeryLazyThenAddLazy": 36,
    "MulCoeffsMontgomeryThenSub": 37, "MulCoeffsMontgomeryThenSubLazy": 38,
    "MulCoeffsMontgomeryLazyThenSubLazy": 39, "MulCoeffsMontgomeryLazyThenNeg": 40,
    "MulCoeffsLazy": 20, "MulCoeffsLazyThenAddLazy": 21,   # SubRing level only in the reference (subring_ops.go:55-66)
}
_UNARY = {"Neg", "Reduce", "ReduceLazy", "MForm", "MFormLazy", "IMForm"}
_BINARY_SCALAR_OPS = {"SubThenMulScalarMontgomeryTwoModulus": 53, "AddLazyThenMulScalarMontgomery": 54}
# two scalars per limb, in the Go argument order (scalar0, scalarMont1): subring_ops.go:167, :223; MaskVec(p1, w, mask, p2)
_TWO_SCALAR_OPS = {"AddScalarLazyThenMulScalarMontgomery": 57, "MulScalarMontgomeryThenAddScalar": 58}
# Ring.*DoubleRNSScalar* (ring/operations.go:165-266): scalar0 for words [0, N/2), scalar1 for [N/2, N)
_DOUBLE_OPS = {"AddDoubleRNSScalar": (60, False), "SubDoubleRNSScalar": (61, False), "MulDoubleRNSScalar": (62, True),
               "MulDoubleRNSScalarThenAdd": (63, True)}
_SCALAR_OPS = {"MulScalarMontgomery": 48, "MulScalarMontgomeryLazy": 49, "MulScalarMontgomeryThenAdd": 50,
               "AddScalar": 51, "SubScalar": 52, "AddScalarLazy": 55, "AddScalarLazyThenNegTwoModulusLazy": 56}


def _is_valid_ring_degree_for_3n(N):   # ring/subring.go:14-30
    if N <= 0:
        return False
    t = 3 * N
    while t % 3 == 0:
        t //= 3
    return t > 0 and (t & (t - 1)) == 0


def _u64_array(x):
    return np.ascontiguousarray(np.asarray(x, dtype=np.uint64))


class _Context:
    """Owns one b200_ctx (shared by a Ring and its AtLevel views)."""

    def __init__(self, ring_type, N, moduli, roots, device):
        self.h = C.c_void_p()
        q = (C.c_uint64 * max(len(moduli), 1))(*moduli)
        r = (C.c_uint64 * len(moduli))(*roots) if roots is not None else None
        rc = lib().b200_ctx_create(int(ring_type), N, len(moduli), q, r, device, C.byref(self.h))
        if rc == -1:
            raise ValueError(lib().b200_last_error().decode())   # the Go constructor's `error`
        check(rc)

    def __del__(self):
        if getattr(self, "h", None) and capi._lib is not None:
            capi._lib.b200_ctx_destroy(self.h)
            self.h = None


class SubRing:
    """ring.SubRing (ring/subring.go:35-55): per-prime constants + the plugin-granularity transforms."""

    def __init__(self, ring, index):
        self._ring, self._i = ring, index
        out = (C.c_uint64 * 8)()
        check(lib().b200_ctx_subring(ring._ctx.h, index, out))
        self.N = ring.N()
        self.Modulus, self.MRedConstant = int(out[0]), int(out[1])
        self.BRedConstant = (int(out[2]), int(out[3]))
        self.Mask, self.PrimitiveRoot, self.NInv, self.Root = int(out[4]), int(out[5]), int(out[6]), int(out[7])

    def _roots(self, inverse):
        n = self._ring.NthRoot() // 2        # subring.go:190-191: NthRoot/2 entries
        out = np.empty(n, dtype=np.uint64)
        check(lib().b200_ctx_roots(self._ring._ctx.h, self._i, inverse, out.ctypes.data_as(capi.P64), n))
        return out

    @property
    def RootsForward(self):
        return self._roots(0)

    @property
    def RootsBackward(self):
        return self._roots(1)

    def _host(self, p1, p2, inverse, flags):
        a = _u64_array(p1)
        out = p2 if (isinstance(p2, np.ndarray) and p2.dtype == np.uint64 and p2.flags["C_CONTIGUOUS"]) else np.empty(max(len(a), 0), dtype=np.uint64)
        n = min(len(a), len(out))
        check(lib().b200_subring_ntt_host(self._ring._ctx.h, self._i, a.ctypes.data_as(capi.P64),
                                          out.ctypes.data_as(capi.P64), n, inverse, flags))
        return out

    # NumberTheoreticTransformer (ring/ntt.go:15-22) on host slices
    def NTT(self, p1, p2=None):
        return self._host(p1, p2, 0, 0)

    def NTTLazy(self, p1, p2=None):
        return self._host(p1, p2, 0, capi.B200_LAZY)

    def INTT(self, p1, p2=None):
        return self._host(p1, p2, 1, 0)

    def INTTLazy(self, p1, p2=None):
        return self._host(p1, p2, 1, capi.B200_LAZY)


class Poly:
    """ring.Poly (ring/poly.go:13-24), device resident and batched: [batch][L][N] uint64."""

    def __init__(self, ring, batch=1, _wrap=None, _keep=None):
        self.ring, self.batch = ring, batch
        self.h = C.c_void_p()
        self._keep = _keep   # keeps a wrapped torch tensor alive
        if _wrap is None:
            check(lib().b200_poly_alloc(ring._ctx.h, batch, C.byref(self.h)))
        else:
            check(lib().b200_poly_wrap(ring._ctx.h, C.c_void_p(_wrap), batch, C.byref(self.h)))

    @classmethod
    def from_torch(cls, ring, tensor):
        """Wrap a CUDA int64/uint64 torch tensor of shape [batch, L, N] without copying."""
        L, N = ring.ModuliChainLength(), ring.N()
        assert tensor.is_cuda and tensor.is_contiguous() and tensor.element_size() == 8
        assert tuple(tensor.shape[-2:]) == (L, N), (tuple(tensor.shape), L, N)
        batch = tensor.numel() // (L * N)
        return cls(ring, batch, _wrap=tensor.data_ptr(), _keep=tensor)

    def shape(self):
        return (self.batch, self.ring.ModuliChainLength(), self.ring.N())

    def N(self):
        return self.ring.N()

    def Level(self):
        return self.ring.MaxLevel()

    def set(self, host):
        """Upload a [batch][L][N] (or [L][N] when batch == 1) host array."""
        a = _u64_array(host).reshape(self.shape())
        check(lib().b200_poly_upload(self.h, a.ctypes.data_as(C.c_void_p), a.size))
        check(lib().b200_sync(self.ring._ctx.h))
        return self

    def get(self):
        out = np.empty(self.shape(), dtype=np.uint64)
        check(lib().b200_poly_download(self.h, out.ctypes.data_as(C.c_void_p), out.size))
        return out

    @property
    def Coeffs(self):
        """Host copy of the limbs: [L][N] for batch 1 (like Poly.Coeffs), else [batch][L][N]."""
        a = self.get()
        return a[0] if self.batch == 1 else a

    def set_limb(self, b, limb, host):
        a = _u64_array(host)
        check(lib().b200_poly_upload_limb(self.h, b, limb, a.ctypes.data_as(capi.P64), len(a)))

    def get_limb(self, b, limb, n=None):
        out = np.empty(self.ring.N() if n is None else n, dtype=np.uint64)
        check(lib().b200_poly_download_limb(self.h, b, limb, out.ctypes.data_as(capi.P64), len(out)))
        return out

    def DevicePtr(self):
        return lib().b200_poly_device_ptr(self.h)

    def IpcHandle(self):
        """64-byte CUDA IPC handle of this (library-allocated) polynomial for Ring.OpenPeer in another process."""
        buf = C.create_string_buffer(64)
        check(lib().b200_poly_ipc_export(self.h, buf))
        return buf.raw

    # -- wire format (ring/poly.go:97-160): structs.Matrix[uint64], little-endian --
    def BinarySize(self, level=None):
        out = C.c_size_t()
        check(lib().b200_poly_binary_size(self.ring._ctx.h, self.ring.level if level is None else level, C.byref(out)))
        return out.value

    def MarshalBinary(self, b=0, level=None):
        """Poly.MarshalBinary for polynomial `b` of the batch, limbs 0..level (default: the ring's level)."""
        level = self.ring.level if level is None else level
        buf = np.empty(self.BinarySize(level), dtype=np.uint8)
        n = C.c_size_t()
        check(lib().b200_poly_marshal(self.h, b, level, buf.ctypes.data_as(C.c_void_p), buf.size, C.byref(n)))
        return buf[: n.value].tobytes()

    def UnmarshalBinary(self, data, b=0):
        """Poly.UnmarshalBinary into polynomial `b`; returns (level found, bytes consumed)."""
        a = np.frombuffer(bytes(data), dtype=np.uint8)
        lvl, used = C.c_int(), C.c_size_t()
        check(lib().b200_poly_unmarshal(self.h, b, a.ctypes.data_as(C.c_void_p), a.size, C.byref(lvl), C.byref(used)))
        return lvl.value, used.value

    def WriteTo(self, w, b=0, level=None):
        """Poly.WriteTo: writes exactly BinarySize() bytes on the binary file object `w`."""
        data = self.MarshalBinary(b, level)
        w.write(data)
        return len(data)

    def ReadFrom(self, r, b=0):
        """Poly.ReadFrom: consumes one serialised polynomial from the binary file object `r`."""
        head = r.read(8)
        if len(head) != 8:
            raise B200Error(capi.B200_ESHAPE, "Poly.ReadFrom: truncated input (no row count)")
        rows = int.from_bytes(head, "little")
        if rows > self.ring.ModuliChainLength():
            raise B200Error(capi.B200_ESHAPE, "Poly.ReadFrom: more limbs than the ring has moduli")
        body = r.read(rows * (8 + 8 * self.ring.N()))
        return self.UnmarshalBinary(head + body, b)[1]

    def Copy(self, other):
        """Poly.Copy (ring/poly.go:73-77); a device-to-device copy when `other` is a Poly, an upload otherwise."""
        if isinstance(other, Poly):
            check(lib().b200_poly_copy(self.h, other.h, -1))
        else:
            self.set(other)
        return self

    def CopyLvl(self, level, other):
        """Poly.CopyLvl (ring/poly.go:81-91): limbs 0..level only."""
        check(lib().b200_poly_copy(self.h, other.h, int(level)))
        return self

    def CopyNew(self):
        """Poly.CopyNew (ring/poly.go:63-70)."""
        return Poly(self.ring, self.batch).Copy(self)

    def Zero(self):
        """Poly.Zero (ring/poly.go:56-60): every limb of every polynomial of the batch."""
        r = self.ring
        (r if r.level == r.MaxLevel() else r.AtLevel(r.MaxLevel())).Zero(self)
        return self

    def Equal(self, other):
        return bool((self.get() == other.get()).all())

    def __del__(self):
        if getattr(self, "h", None) and capi._lib is not None:
            capi._lib.b200_poly_free(self.h)
            self.h = None


class Ring:
    """ring.Ring (ring/ring.go:76-89) on one GPU."""

    def __init__(self, ctx, ring_type, N, moduli, level, device):
        self._ctx, self._type, self._N, self._moduli, self.level, self.device = ctx, ring_type, N, list(moduli), level, device
        self.SubRings = [SubRing(self, i) for i in range(len(moduli))]

    # -- accessors (ring/ring.go) --
    def N(self):
        return self._N

    def NthRoot(self):
        return {int(Matrix): 3, int(ConjugateInvariant): 4}.get(int(self._type), 2) * self._N

    def ConjugateInvariantRing(self):
        """Ring.ConjugateInvariantRing (ring/ring.go:95-110): the conjugate-invariant ring of half the degree over the
        same moduli (the receiver itself if it already is one)."""
        if self._type == ConjugateInvariant:
            return self
        if self._type != Standard:
            raise ValueError("cannot ConjugateInvariantRing: the ring is not a power-of-two ring")
        return NewRingFromType(self._N >> 1, self._moduli[: self.level + 1], ConjugateInvariant, self.device)

    def StandardRing(self):
        """Ring.StandardRing (ring/ring.go:112-127): the standard ring of twice the degree over the same moduli."""
        if self._type == Standard:
            return self
        if self._type != ConjugateInvariant:
            raise ValueError("cannot StandardRing: the ring is not a conjugate-invariant ring")
        return NewRingFromType(self._N << 1, self._moduli[: self.level + 1], Standard, self.device)

    # -- host-side accessors and big-integer glue of ring/ring.go (no device arithmetic: CRT on downloaded words) --
    def LogN(self):   # ring/ring.go:159-161
        return (self._N - 1).bit_length()

    def Modulus(self):   # :232-234: the product of the moduli at this level
        m = 1
        for q in self._moduli[: self.level + 1]:
            m *= q
        return m

    def LogModuli(self):   # :170-175
        import math
        return sum(math.log2(q) for q in self._moduli[: self.level + 1])

    def MRedConstants(self):   # :238-246
        return [s.MRedConstant for s in self.SubRings[: self.level + 1]]

    def BRedConstants(self):   # :249-257
        return [s.BRedConstant for s in self.SubRings[: self.level + 1]]

    # ring/scalar.go: RNSScalar helpers at this ring's level (host arithmetic, scalar.py)
    def _chain(self):
        return self._moduli[: self.level + 1]

    def NewRNSScalar(self):
        return [0] * (self.level + 1)

    def NewRNSScalarFromUInt64(self, v):
        return _scalar.NewRNSScalarFromUInt64(self._chain(), v)

    def NewRNSScalarFromBigint(self, v):
        return _scalar.NewRNSScalarFromBigint(self._chain(), v)

    def MFormRNSScalar(self, s1):
        return _scalar.MFormRNSScalar(self._chain(), s1)

    def NegRNSScalar(self, s1):
        return _scalar.NegRNSScalar(self._chain(), s1)

    def SubRNSScalar(self, s1, s2):
        return _scalar.SubRNSScalar(self._chain(), s1, s2)

    def MulRNSScalar(self, s1, s2):
        return _scalar.MulRNSScalar(self._chain(), s1, s2)

    def Inverse(self, a):
        return _scalar.Inverse(self._chain(), a)

    def NewMonomialXi(self, i):
        """Ring.NewMonomialXi (ring/ring.go:408-430): the polynomial X^i (X^N = -1), coefficient domain."""
        N = self._N
        i &= (N << 1) - 1
        if i >= N:
            i -= N << 1
        host = np.zeros((1, len(self._moduli), N), dtype=np.uint64)
        for k, q in enumerate(self._moduli[: self.level + 1]):
            if i < 0:
                host[0, k, N + i] = q - 1
            else:
                host[0, k, i] = 1
        return self.NewPoly().set(host)

    def SetCoefficientsBigint(self, coeffs, p1, b=0):
        """Ring.SetCoefficientsBigint (ring/ring.go:433-447): limb i of polynomial b of p1 <- coeffs mod q_i."""
        for i, q in enumerate(self._moduli[: self.level + 1]):
            limb = p1.get_limb(b, i)
            limb[: len(coeffs)] = np.array([int(c) % q for c in coeffs], dtype=np.uint64)
            p1.set_limb(b, i, limb)

    def _crt(self, p1, gap, b, centred):
        Q = self.Modulus()
        rec = []
        for q in self._moduli[: self.level + 1]:
            m = Q // q
            rec.append(m * pow(m, -1, q))
        limbs = [p1.get_limb(b, k) for k in range(self.level + 1)]
        out = []
        for j in range(0, self._N, gap):
            x = sum(int(limbs[k][j]) * rec[k] for k in range(self.level + 1)) % Q
            out.append(x - Q if centred and x >= (Q >> 1) else x)
        return out

    def PolyToBigint(self, p1, gap=1, b=0):
        """Ring.PolyToBigint (ring/ring.go:467-496): CRT reconstruction of the coefficients X^(i*gap), in [0, Q)."""
        return self._crt(p1, gap, b, False)

    def PolyToBigintCentered(self, p1, gap=1, b=0):
        """Ring.PolyToBigintCentered (ring/ring.go:503-543): the same, in [-Q/2, Q/2) (values >= floor(Q/2) wrap)."""
        return self._crt(p1, gap, b, True)

    def PolyToString(self, p1, b=0):   # ring/ring.go:450-462
        return [str(x) for x in self.PolyToBigint(p1, 1, b)]

    def Shift(self, p1, k, p2):
        """Ring.Shift (ring/operations.go:278-282): rotate every limb left by k."""
        check(lib().b200_shift(self._ctx.h, p1.h, int(k), p2.h))

    def MultByMonomial(self, p1, k, p2):
        """Ring.MultByMonomial (ring/operations.go:306-363): p2 = X^k * p1, X^N = -1."""
        check(lib().b200_mult_by_monomial(self._ctx.h, p1.h, int(k), p2.h, self.level))

    # -- host-side scalar preparation + the scalar operators (ring/operations.go:151-267) --
    def _mform_scalars(self, scalar_of_limb):
        out = []
        for s in self.SubRings[: self.level + 1]:
            out.append(int(scalar_of_limb(s)) % s.Modulus * (1 << 64) % s.Modulus)      # MForm(scalar, q, BRedConstant)
        return out

    def MulScalar(self, p1, scalar, p2):                                                # :201-205
        self.MulScalarMontgomery(p1, self._mform_scalars(lambda s: scalar), p2)

    def MulScalarThenAdd(self, p1, scalar, p2):                                         # :208-212
        self.MulScalarMontgomeryThenAdd(p1, self._mform_scalars(lambda s: scalar), p2)

    def MulScalarThenSub(self, p1, scalar, p2):                                         # :223-228
        self.MulScalarMontgomeryThenAdd(p1, self._mform_scalars(lambda s: s.Modulus - int(scalar) % s.Modulus), p2)

    def MulRNSScalarMontgomery(self, p1, scalar, p2):                                   # :216-220
        self.MulScalarMontgomery(p1, [int(v) for v in scalar[: self.level + 1]], p2)

    def MulScalarBigint(self, p1, scalar, p2):                                          # :231-237
        self.MulScalar(p1, scalar, p2)

    def MulScalarBigintThenAdd(self, p1, scalar, p2):                                   # :240-247
        self.MulScalarThenAdd(p1, scalar, p2)

    def AddScalarBigint(self, p1, scalar, p2):                                          # :158-163
        self.AddScalar(p1, [int(scalar) % s.Modulus for s in self.SubRings[: self.level + 1]], p2)

    def SubScalarBigint(self, p1, scalar, p2):                                          # :193-198
        self.SubScalar(p1, [int(scalar) % s.Modulus for s in self.SubRings[: self.level + 1]], p2)

    def EvalPolyScalar(self, p1, scalar, p2):                                           # :269-275
        p2.Copy(p1[-1])
        for i in range(len(p1) - 1, 0, -1):
            self.MulScalar(p2, scalar, p2)
            self.Add(p2, p1[i - 1], p2)

    # -- automorphisms X -> X^gen (ring/automorphism.go); none of them works in place --
    def AutomorphismNTT(self, polIn, gen, polOut):
        check(lib().b200_automorphism_ntt(self._ctx.h, polIn.h, int(gen) % (1 << 64), polOut.h, self.level, 0))

    def _auto_index(self, polIn, index, polOut, add):
        idx = _u64_array(index)
        check(lib().b200_automorphism_ntt_with_index(self._ctx.h, polIn.h, idx.ctypes.data_as(capi.P64), len(idx), polOut.h,
                                                     self.level, add))

    def AutomorphismNTTWithIndex(self, polIn, index, polOut):
        self._auto_index(polIn, index, polOut, 0)

    def AutomorphismNTTWithIndexThenAddLazy(self, polIn, index, polOut):
        self._auto_index(polIn, index, polOut, 1)

    def Automorphism(self, polIn, gen, polOut):
        check(lib().b200_automorphism(self._ctx.h, polIn.h, int(gen) % (1 << 64), polOut.h, self.level))

    def FoldStandardToConjugateInvariant(self, polyStandard, permuteNTTIndexInv, polyConjugateInvariant):
        """ring/conjugate_invariant.go:28-44 (the receiver is the conjugate-invariant ring)."""
        idx = _u64_array(permuteNTTIndexInv)
        check(lib().b200_fold_standard_to_ci(self._ctx.h, polyStandard.h, idx.ctypes.data_as(capi.P64), len(idx),
                                             polyConjugateInvariant.h, self.level))

    def UnfoldConjugateInvariantToStandard(self, polyConjugateInvariant, polyStandard):
        """ring/conjugate_invariant.go:7-23 (the receiver is the conjugate-invariant ring)."""
        check(lib().b200_unfold_ci_to_standard(self._ctx.h, polyConjugateInvariant.h, polyStandard.h, self.level))

    def Type(self):
        return self._type

    # -- parameters literal (ring/ring.go:560-608, ring/subring.go:330-360) --
    def parametersLiteral(self):
        """ringParametersLiteral: one subRingParametersLiteral per prime, field for field as the Go struct.
        Reference quirks kept on purpose: a 3N SubRing reports Type Standard (subring.go:114-123), LogN =
        bits.Len64(N-1) (so N = 3*2^15 gives 17) and PrimitiveRoot = omega (subring.go:165)."""
        out = []
        for i, s in enumerate(self.SubRings):
            n = lib().b200_ctx_factors(self._ctx.h, i, None, 0)
            if n < 0:
                check(n)
            f = (C.c_uint64 * max(n, 1))()
            rc = lib().b200_ctx_factors(self._ctx.h, i, f, n)
            if rc < 0:
                check(rc)
            out.append({"Type": int(self._type) if self._type == ConjugateInvariant else 0,
                        "LogN": (self._N - 1).bit_length(), "NthRoot": self.NthRoot() // self._N,
                        "Modulus": s.Modulus, "Factors": [int(f[k]) for k in range(n)],
                        "PrimitiveRoot": s.Root if self._type == Matrix else s.PrimitiveRoot})
        return out

    def MarshalJSON(self):
        """Ring.MarshalJSON (ring/ring.go:587-589): the bytes encoding/json produces for the literal."""
        import json
        return json.dumps(self.parametersLiteral(), separators=(",", ":")).encode()

    def MarshalBinary(self):   # ring/ring.go:577-579
        return self.MarshalJSON()

    def Is3N(self):   # ring/ring.go:353 quirk: true for powers of two as well
        return _is_valid_ring_degree_for_3n(self._N)

    def Level(self):
        return self.level

    def MaxLevel(self):
        return len(self._moduli) - 1

    def ModuliChain(self):
        return list(self._moduli)

    def ModuliChainLength(self):
        return len(self._moduli)

    def AtLevel(self, level):   # ring/ring.go:194-213 (panics -> B200Error)
        if level < 0:
            raise B200Error(-4, "level cannot be negative")
        if level > self.MaxLevel():
            raise B200Error(-4, "level cannot be larger than max level")
        r = Ring.__new__(Ring)
        r.__dict__.update(self.__dict__)
        r.level = level
        return r

    def NewPoly(self, batch=1):
        return Poly(self, batch)

    def SetStream(self, cuda_stream):
        """Enqueue on this cudaStream_t (e.g. torch.cuda.current_stream().cuda_stream)."""
        check(lib().b200_ctx_set_stream(self._ctx.h, C.c_void_p(cuda_stream)))

    def Sync(self):
        check(lib().b200_sync(self._ctx.h))

    def LaunchCount(self):
        return int(lib().b200_ctx_launch_count(self._ctx.h))

    def Profile(self, fn):
        """Run fn() with every kernel launch bracketed by CUDA events; returns
        {kernel: {"ms": total, "launches": n, "units": limb-units}} (bench.py roofline)."""
        check(lib().b200_profile_begin(self._ctx.h))
        try:
            fn()
        finally:
            buf = C.create_string_buffer(1 << 16)
            check(lib().b200_profile_end(self._ctx.h, buf, len(buf)))
        out = {}
        for line in buf.value.decode().splitlines():
            name, ms, launches, units = line.split()
            out[name] = {"ms": float(ms), "launches": int(launches), "units": int(units)}
        return out

    # -- transforms (ring/ntt.go:127-152) --
    def NTT(self, p1, p2, tree_order=False):
        check(lib().b200_ntt(self._ctx.h, p1.h, p2.h, self.level, capi.B200_TREE_ORDER if tree_order else 0))

    def NTTLazy(self, p1, p2, tree_order=False):
        check(lib().b200_ntt(self._ctx.h, p1.h, p2.h, self.level, capi.B200_LAZY | (capi.B200_TREE_ORDER if tree_order else 0)))

    def INTT(self, p1, p2, tree_order=False):
        check(lib().b200_intt(self._ctx.h, p1.h, p2.h, self.level, capi.B200_TREE_ORDER if tree_order else 0))

    def INTTLazy(self, p1, p2, tree_order=False):
        check(lib().b200_intt(self._ctx.h, p1.h, p2.h, self.level, capi.B200_LAZY | (capi.B200_TREE_ORDER if tree_order else 0)))

    # -- host-buffer transforms: H2D, kernels and D2H of successive chunks overlapped --
    def TransformHost(self, host_in, host_out, ops, tree_order=False):
        """Apply ops (sequence of "NTT"/"INTT") to [batch][L][N] HOST polynomials; host_in/host_out are numpy
        uint64 arrays or (pinned) torch int64 tensors.  Returns host_out."""
        code = (C.c_int * len(ops))(*[{"NTT": 0, "INTT": 1}[o] for o in ops])

        def ptr_words(a):
            if hasattr(a, "data_ptr"):
                return a.data_ptr(), a.numel()
            return a.ctypes.data, a.size
        pi, ni = ptr_words(host_in)
        po, no = ptr_words(host_out)
        per = self.ModuliChainLength() * self._N
        if ni != no or ni % per:
            raise B200Error(-4, "host buffers must hold [batch][L][N] words")
        check(lib().b200_transform_host(self._ctx.h, C.c_void_p(pi), C.c_void_p(po), ni // per, self.level, code, len(ops),
                                        capi.B200_TREE_ORDER if tree_order else 0))
        return host_out

    def NTTLimbs(self, p1, p2, limb0, nl, inverse=False):
        """Ring.NTT / INTT on limbs [limb0, limb0+nl) only (a rank's share of a limb-sharded ring)."""
        check(lib().b200_ntt_limbs(self._ctx.h, p1.h, p2.h, limb0, nl, int(inverse), 0))

    # -- fused all-gather over NVLink peer memory (one process per GPU) --
    def OpenPeer(self, handle):
        """Map a polynomial exported by another process (Poly.IpcHandle) into this rank; returns its device pointer."""
        ptr = C.c_void_p()
        check(lib().b200_peer_open(self._ctx.h, bytes(handle), C.byref(ptr)))
        return ptr.value

    def ClosePeer(self, ptr):
        check(lib().b200_peer_close(self._ctx.h, C.c_void_p(ptr)))

    def INTTLimbsGather(self, p1, p2, limb0, nl, peers):
        """Ring.INTT on limbs [limb0, limb0+nl) of p1 into p2 and into the same limbs of the mapped peer polynomials
        (device pointers from OpenPeer), written from inside the transform's last kernel."""
        arr = (C.c_void_p * max(len(peers), 1))(*peers)
        check(lib().b200_intt_limbs_gather(self._ctx.h, p1.h, p2.h, limb0, nl, 0, len(peers), arr))

    # -- fused ring product (BASELINE config 4) --
    def MulPoly(self, p1, p2, p3):
        """p3 = p1 * p2 in the ring (NTT, NTT, MForm, MulCoeffsMontgomery, INTT fused on device)."""
        check(lib().b200_polymul(self._ctx.h, p1.h, p2.h, p3.h, self.level))

    def DivRoundByLastModulusNTT(self, p0, p1):
        """ring/scaling.go:130-156 at this ring's level (the `buff` argument of the Go method is internal here)."""
        check(lib().b200_div_by_last_modulus_ntt(self._ctx.h, p0.h, p1.h, self.level, 1))

    def DivFloorByLastModulusNTT(self, p0, p1):
        """ring/scaling.go:6-19."""
        check(lib().b200_div_by_last_modulus_ntt(self._ctx.h, p0.h, p1.h, self.level, 0))

    def DivFloorByLastModulus(self, p0, p1):
        """ring/scaling.go:21-30 (coefficient domain)."""
        check(lib().b200_div_by_last_modulus(self._ctx.h, p0.h, p1.h, self.level, 0))

    def DivRoundByLastModulus(self, p0, p1):
        """ring/scaling.go:113-127 (coefficient domain; p0 is left unchanged unless p1 is p0)."""
        check(lib().b200_div_by_last_modulus(self._ctx.h, p0.h, p1.h, self.level, 1))

    def DivFloorByLastModulusMany(self, nbRescales, p0, p1):      # ring/scaling.go:57-91
        check(lib().b200_div_by_last_modulus_many(self._ctx.h, p0.h, p1.h, self.level, int(nbRescales), 0, 0))

    def DivFloorByLastModulusManyNTT(self, nbRescales, p0, p1):   # :32-55
        check(lib().b200_div_by_last_modulus_many(self._ctx.h, p0.h, p1.h, self.level, int(nbRescales), 0, 1))

    def DivRoundByLastModulusMany(self, nbRescales, p0, p1):      # :160-192
        check(lib().b200_div_by_last_modulus_many(self._ctx.h, p0.h, p1.h, self.level, int(nbRescales), 1, 0))

    def DivRoundByLastModulusManyNTT(self, nbRescales, p0, p1):   # :130-158
        check(lib().b200_div_by_last_modulus_many(self._ctx.h, p0.h, p1.h, self.level, int(nbRescales), 1, 1))

    def Rescale(self, nbRescales, c0, c1, out0, out1):
        """ckks.Evaluator.Rescale on the two polynomials of a ciphertext (schemes/ckks/evaluator.go:500-541, and the
        matrix_ckks form :208-243): DivRoundByLastModulusManyNTT of each component; the caller lowers its level by
        nbRescales (use AtLevel(level - nbRescales) afterwards)."""
        self.DivRoundByLastModulusManyNTT(nbRescales, c0, out0)
        self.DivRoundByLastModulusManyNTT(nbRescales, c1, out1)

    def TensorProduct(self, a0, a1, b0, b1, c0, c1, c2, mform=True):
        """(a0,a1) x (b0,b1) -> (c0,c1,c2) in one pass: the MForm/MulCoeffsMontgomery[ThenAdd] sequence of
        ckks.Evaluator.mulRelin (mform=True) or matrix_ckks.Evaluator.Mul (mform=False)."""
        check(lib().b200_tensor(self._ctx.h, a0.h, a1.h, b0.h, b1.h, c0.h, c1.h, c2.h, self.level, int(mform)))

    def _vec(self, op, p1, p2, p3, scalars=None):
        sc = None
        if scalars is not None:
            if len(scalars) == 2 and not np.isscalar(scalars[0]):   # two scalars per limb: ([scalar0_i], [scalar1_i])
                if min(len(scalars[0]), len(scalars[1])) < self.level + 1:
                    raise B200Error(-4, "need one scalar per limb")
                vals = [int(s) for s in scalars[0][: self.level + 1]] + [int(s) for s in scalars[1][: self.level + 1]]
            else:
                vals = [int(s) for s in scalars]
                if len(vals) < self.level + 1:
                    raise B200Error(-4, "need one scalar per limb")
                vals = vals[: self.level + 1]
            sc = (C.c_uint64 * len(vals))(*vals)
        check(lib().b200_vec_op(self._ctx.h, op, p1.h, p2.h if p2 is not None else None, p3.h, self.level, sc))

    def Zero(self, p):
        """ring.ZeroVec on limbs 0..level of every polynomial of the batch (ring/vec_ops.go:847-868, Poly.Zero ring/poly.go)."""
        self._vec(10, p, None, p)

    def MaskVec(self, p1, w, mask, p2):
        """ring.MaskVec (ring/vec_ops.go:870-890) on every limb: p2 = (p1 >> w) & mask."""
        n = self.level + 1
        self._vec(59, p1, None, p2, ([int(w)] * n, [int(mask)] * n))

    def MulByVectorMontgomery(self, p1, vector, p2):
        """Ring.MulByVectorMontgomery (ring/operations.go:366-370); `vector`: a Poly whose first limb holds the N words."""
        check(lib().b200_mul_by_vector_montgomery(self._ctx.h, p1.h, vector.h, p2.h, self.level, 0))

    def MulByVectorMontgomeryThenAddLazy(self, p1, vector, p2):
        """Ring.MulByVectorMontgomeryThenAddLazy (ring/operations.go:373-377)."""
        check(lib().b200_mul_by_vector_montgomery(self._ctx.h, p1.h, vector.h, p2.h, self.level, 1))

    def Equal(self, p1, p2):   # ring/ring.go:546-558: reduces both sides, then compares limbs 0..level
        a, b = self.NewPoly(p1.batch), self.NewPoly(p2.batch)
        self.Reduce(p1, a)
        self.Reduce(p2, b)
        return bool((a.get()[:, : self.level + 1] == b.get()[:, : self.level + 1]).all())


class KeySwitcher:
    """ring.BasisExtender + ring.Decomposer + rlwe.Evaluator.GadgetProduct over one QP ring.
    `ringQP` is a Ring over Q_0..Q_{LQ-1}, P_0..P_{LP-1}; all polynomials use its [batch][LQ+LP][N] layout."""

    def __init__(self, ringQP, LQ):
        self.ring, self.LQ, self.LP = ringQP, LQ, ringQP.ModuliChainLength() - LQ
        self.h = C.c_void_p()
        rc = lib().b200_keyswitch_create(ringQP._ctx.h, LQ, C.byref(self.h))
        if rc == -1:
            raise ValueError(lib().b200_last_error().decode())
        check(rc)

    def ModUpQtoP(self, levelQ, levelP, polQ, polP):   # ring/basis_extension.go:184-198
        check(lib().b200_mod_up(self.h, 0, levelQ, levelP, polQ.h, polP.h))

    def ModUpPtoQ(self, levelP, levelQ, polP, polQ):   # :203-217
        check(lib().b200_mod_up(self.h, 1, levelQ, levelP, polP.h, polQ.h))

    def ModDownQPtoQNTT(self, levelQ, levelP, p1, p2):   # :241-258
        check(lib().b200_mod_down_ntt(self.h, levelQ, levelP, p1.h, p2.h))

    def ModDownQPtoQ(self, levelQ, levelP, p1, p2):   # :223-236 (coefficient domain)
        check(lib().b200_mod_down(self.h, 0, levelQ, levelP, p1.h, p2.h))

    def ModDownQPtoP(self, levelQ, levelP, p1, p2):   # :264-279 (coefficient domain; result in the P limbs of p2)
        check(lib().b200_mod_down(self.h, 1, levelQ, levelP, p1.h, p2.h))

    def GadgetProduct(self, levelQ, levelP, cx, key, ct0, ct1):   # core/rlwe/evaluator_gadget_product.go:16-30
        check(lib().b200_gadget_product(self.h, levelQ, levelP, cx.h, key.h, ct0.h, ct1.h))

    def DecomposeNTT(self, levelQ, levelP, cx, decomp):
        """rlwe.Evaluator.DecomposeNTT (evaluator_gadget_product.go:430-478): decomp = beta*batch polynomials [digit][b]."""
        check(lib().b200_decompose_ntt(self.h, levelQ, levelP, cx.h, decomp.h))

    def GadgetProductHoisted(self, levelQ, levelP, decomp, key, ct0, ct1):   # :322-423
        check(lib().b200_gadget_product_hoisted(self.h, levelQ, levelP, decomp.h, key.h, ct0.h, ct1.h))

    def Automorphism(self, levelQ, levelP, ct0, ct1, galEl, key, out0, out1, decomp=None):
        """rlwe.Evaluator.Automorphism / AutomorphismHoisted (core/rlwe/evaluator_automorphism.go:13-108), NTT domain."""
        check(lib().b200_automorphism_ct(self.h, levelQ, levelP, ct0.h, ct1.h, decomp.h if decomp is not None else None,
                                         int(galEl) % (1 << 64), key.h, out0.h, out1.h))

    def GadgetProductBitDecomp(self, levelQ, levelP, pw2, sizes, cx, key, ct0, ct1):
        """GadgetProduct with a power-of-two base (evaluator_gadget_product.go:190-320); sizes = BaseTwoDecompositionVectorSize."""
        sz = (C.c_int * len(sizes))(*[int(v) for v in sizes])
        check(lib().b200_gadget_product_bit_decomp(self.h, levelQ, levelP, pw2, sz, cx.h, key.h, ct0.h, ct1.h))

    def MulRelin(self, levelQ, levelP, a0, a1, b0, b1, rlk, out0, out1, mform=True):
        """ct x ct multiply + relinearisation in the NTT domain: tensor product, GadgetProduct(c2, rlk), add
        (schemes/ckks/evaluator.go:821-853)."""
        check(lib().b200_mul_relin(self.h, levelQ, levelP, a0.h, a1.h, b0.h, b1.h, rlk.h, out0.h, out1.h, int(mform)))

    def _tmp(self, batch):
        t = getattr(self, "_tmp_polys", None)
        if t is None or t[0].batch != batch:
            t = self._tmp_polys = (self.ring.NewPoly(batch), self.ring.NewPoly(batch))
        return t

    def Relinearize(self, levelQ, levelP, c0, c1, c2, rlk, out0, out1):
        """rlwe.Evaluator.Relinearize (core/rlwe/evaluator_evaluationkey.go:125-153): GadgetProduct(c2, rlk) into the
        evaluator's buffers, then out_i = c_i + tmp_i on the limbs 0..levelQ."""
        t0, t1 = self._tmp(c2.batch)
        self.GadgetProduct(levelQ, levelP, c2, rlk, t0, t1)
        rq = self.ring.AtLevel(levelQ)
        rq.Add(c0, t0, out0)
        rq.Add(c1, t1, out1)

    def ApplyEvaluationKey(self, levelQ, levelP, c0, c1, evk, out0, out1):
        """rlwe.Evaluator.ApplyEvaluationKey for equal ring degrees (evaluator_evaluationkey.go:37-108 -> applyEvaluationKey
        :110-117): GadgetProduct(c1, evk), out0 = c0 + tmp0, out1 = tmp1 (CopyLvl)."""
        t0, t1 = self._tmp(c1.batch)
        self.GadgetProduct(levelQ, levelP, c1, evk, t0, t1)
        self.ring.AtLevel(levelQ).Add(c0, t0, out0)
        out1.CopyLvl(levelQ, t1)

    # limb-sharded steps (see sharding.gadget_product_sharded)
    def DecomposeMac(self, levelQ, levelP, cx, cxinv, key, acc0, acc1, q0, q1, p0, p1):
        check(lib().b200_ks_decompose_mac(self.h, levelQ, levelP, cx.h, cxinv.h, key.h, acc0.h, acc1.h, q0, q1, p0, p1))

    def ModDownLimbs(self, levelQ, levelP, acc, out, q0, q1):
        check(lib().b200_ks_mod_down_limbs(self.h, levelQ, levelP, acc.h, out.h, q0, q1))

    def __del__(self):
        if getattr(self, "h", None) and capi._lib is not None:
            capi._lib.b200_keyswitch_destroy(self.h)
            self.h = None


def _mk_binary(name, code):
    def f(self, p1, p2, p3):
        self._vec(code, p1, p2, p3)
    f.__name__ = name
    f.__doc__ = "Ring.%s(p1, p2, p3) -- ring/operations.go" % name
    return f


def _mk_unary(name, code):
    def f(self, p1, p2):
        self._vec(code, p1, None, p2)
    f.__name__ = name
    f.__doc__ = "Ring.%s(p1, p2) -- ring/operations.go" % name
    return f


def _mk_scalar(name, code):
    def f(self, p1, scalars, p2):
        self._vec(code, p1, None, p2, scalars)
    f.__name__ = name
    f.__doc__ = "per-limb SubRing.%s(p1, scalar_i, p2) -- ring/subring_ops.go" % name
    return f


for _n, _c in _OPS.items():
    setattr(Ring, _n, _mk_unary(_n, _c) if _n in _UNARY else _mk_binary(_n, _c))
for _n, _c in _SCALAR_OPS.items():
    setattr(Ring, _n, _mk_scalar(_n, _c))


def _mk_binary_scalar(name, code):
    def f(self, p1, p2, scalars, p3):
        self._vec(code, p1, p2, p3, scalars)
    f.__name__ = name
    f.__doc__ = "per-limb SubRing.%s(p1, p2, scalar_i, p3) -- ring/subring_ops.go" % name
    return f


for _n, _c in _BINARY_SCALAR_OPS.items():
    setattr(Ring, _n, _mk_binary_scalar(_n, _c))


def _mk_two_scalar(name, code):
    def f(self, p1, scalars0, scalars1, p2):
        self._vec(code, p1, None, p2, (list(scalars0), list(scalars1)))
    f.__name__ = name
    f.__doc__ = "per-limb SubRing.%s(p1, scalar0_i, scalarMont1_i, p2) -- ring/subring_ops.go" % name
    return f


def _mk_double(name, code, mform):
    def f(self, p1, scalar0, scalar1, p2):
        s0, s1 = list(scalar0), list(scalar1)
        if mform:   # MForm(scalar_i, q_i, BRedConstant) on the host, like operations.go:253-254
            qs = [sr.Modulus for sr in self.SubRings[: self.level + 1]]
            s0 = [int(v) % q * (1 << 64) % q for v, q in zip(s0, qs)]
            s1 = [int(v) % q * (1 << 64) % q for v, q in zip(s1, qs)]
        self._vec(code, p1, None, p2, (s0, s1))
    f.__name__ = name
    f.__doc__ = "Ring.%s(p1, scalar0, scalar1 RNSScalar, p2) -- ring/operations.go:165-266" % name
    return f


for _n, _c in _TWO_SCALAR_OPS.items():
    setattr(Ring, _n, _mk_two_scalar(_n, _c))
for _n, (_c, _m) in _DOUBLE_OPS.items():
    setattr(Ring, _n, _mk_double(_n, _c, _m))


def NewRingFromType(N, Moduli, ringType, device=0, roots=None):
    """ring.NewRingFromType (ring/ring.go:291-308).  `roots`: see b200_ctx_create."""
    ringType = Type(ringType)
    if ringType not in (Standard, ConjugateInvariant, Matrix):
        raise ValueError("invalid ring type")
    moduli = [int(q) for q in Moduli]
    ctx = _Context(ringType, N, moduli, roots, device)
    return Ring(ctx, ringType, N, moduli, len(moduli) - 1, device)


def AutomorphismNTTIndex(N, NthRoot, GalEl):
    """ring.AutomorphismNTTIndex (ring/automorphism.go:12-35); raises ValueError where Go returns an error."""
    out = np.empty(max(N, 0), dtype=np.uint64)
    rc = lib().b200_automorphism_ntt_index(N, int(NthRoot) % (1 << 64), int(GalEl) % (1 << 64), out.ctypes.data_as(capi.P64))
    if rc == -1:
        raise ValueError(lib().b200_last_error().decode())
    check(rc)
    return out


def NewRingConjugateInvariant(N, Moduli, device=0, roots=None):
    """ring.NewRingConjugateInvariant (ring/ring.go:280-282): Z_q[X + X^-1]/(X^2N + 1), moduli = 1 mod 4N."""
    return NewRingFromType(N, Moduli, ConjugateInvariant, device, roots)


def NewRing(N, Moduli, device=0, roots=None):
    """ring.NewRing (ring/ring.go:264-272): 3N transformer iff N % 3 == 0 and 3N = 2^a 3^(b+1)."""
    is3n = N % 3 == 0 and _is_valid_ring_degree_for_3n(N)
    return NewRingFromType(N, Moduli, Matrix if is3n else Standard, device, roots)


def UnmarshalJSON(data, device=0, N=None):
    """Ring.UnmarshalJSON / UnmarshalBinary (ring/ring.go:582-657, ring/subring.go:362-417).

    Power-of-two rings are rebuilt from the literal exactly as the reference does (same primitive root, hence the
    same tables).  A literal written by a 3N ring (NthRoot == 3) does not determine N -- the reference stores
    only bits.Len64(N-1) and itself rebuilds a wrong power-of-two ring from it -- so it is refused unless the caller
    supplies the ring degree with `N=`."""
    import json
    lit = json.loads(bytes(data).decode())
    if not isinstance(lit, list) or not lit:
        raise ValueError("invalid ring literal")
    for k, p in enumerate(lit):
        if p["Type"] not in (int(Standard), int(ConjugateInvariant)):
            raise ValueError("invalid ring type")
        if k and (p["LogN"] != lit[0]["LogN"] or p["NthRoot"] != lit[0]["NthRoot"]):
            raise ValueError("invalid SubRings: all SubRings must have the same ring degree and NthRoot")
    nth = lit[0]["NthRoot"]
    moduli = [int(p["Modulus"]) for p in lit]
    roots = [int(p["PrimitiveRoot"]) for p in lit]
    if nth == 3:
        if N is None:
            raise ValueError("3N ring literal does not determine the ring degree (LogN = %d); pass N=" % lit[0]["LogN"])
        if (N - 1).bit_length() != lit[0]["LogN"] or N % 3 != 0 or not _is_valid_ring_degree_for_3n(N):
            raise ValueError("N does not match the literal")
        return NewRingFromType(N, moduli, Matrix, device, roots)
    if nth < 2:
        raise ValueError("invalid ring type: NthRoot must be at least 2N but is %dN" % nth)
    if lit[0]["Type"] == int(ConjugateInvariant):
        if nth != 4:
            raise ValueError("invalid NthRoot: a ConjugateInvariant ring has NthRoot 4N")
        return NewRingFromType(1 << lit[0]["LogN"], moduli, ConjugateInvariant, device, roots)
    if nth != 2:
        raise ValueError("invalid NthRoot: only 2N (Standard), 4N (ConjugateInvariant) and 3N (Matrix) rings are supported")
    return NewRingFromType(1 << lit[0]["LogN"], moduli, Standard, device, roots)


def NewPoly(ring, batch=1):
    return Poly(ring, batch)
