"""Host-side mirror of ring/primes.go and ring/primes_3n.go over the C ABI (b200_find_3n_rns_primes, ...): the modulus
and root-of-unity search a caller runs before a ring constructor.  Errors: ValueError where Go returns an `error`."""
import ctypes as C

import numpy as np

from . import capi
from .capi import check, lib

_M64 = (1 << 64) - 1


def _err(rc):
    if rc == capi.B200_EINVAL:
        raise ValueError(lib().b200_last_error().decode())
    check(rc)


def IsPrime(x):
    """ring.IsPrime (ring/primes.go:11-13)."""
    return bool(lib().b200_is_prime(int(x) & _M64))


def Find3NRNSPrimes(N, bitSize, count, searchBudget):
    """ring.Find3NRNSPrimes (ring/primes_3n.go:11-43)."""
    out = np.zeros(max(int(count), 1), dtype=np.uint64)
    _err(lib().b200_find_3n_rns_primes(int(N), int(bitSize), int(count), int(searchBudget), out.ctypes.data_as(capi.P64)))
    return [int(q) for q in out[:count]]


def Find3NFriendlyPrime(N, bitSize):
    """ring.Find3NFriendlyPrime (ring/primes.go:16-67)."""
    out = C.c_uint64(0)
    _err(lib().b200_find_3n_friendly_prime(int(N), int(bitSize), C.byref(out)))
    return out.value


def OrderMod(g, p):
    """ring.OrderMod (ring/primes_3n.go:78-91)."""
    out = C.c_uint64(0)
    _err(lib().b200_order_mod(int(g) & _M64, int(p) & _M64, C.byref(out)))
    return out.value


def HasPrimitiveMthRoot(q, m):
    """ring.HasPrimitiveMthRoot (ring/primes_3n.go:121-123)."""
    return bool(lib().b200_has_primitive_mth_root(int(q) & _M64, int(m) & _M64))


def PrimitiveRoot(q):
    """ring.PrimitiveRoot (ring/subring.go:218-251): the smallest primitive root >= 3."""
    out = C.c_uint64(0)
    _err(lib().b200_primitive_root(int(q) & _M64, C.byref(out)))
    return out.value


def FindPrimitiveRootOfUnity(q, m, seed=0):
    """ring.FindPrimitiveRootOfUnity (ring/primes_3n.go:127-149).  seed == 0: the deterministic root of
    ring.Find3NPrimitiveRoot (ring/subring.go:255-290); otherwise bases drawn from splitmix64(seed)."""
    out = C.c_uint64(0)
    _err(lib().b200_find_primitive_root_of_unity(int(q) & _M64, int(m) & _M64, int(seed) & _M64, C.byref(out)))
    return out.value


def Find3NPrimitiveRoot(p, NthRoot):
    """ring.Find3NPrimitiveRoot (ring/subring.go:255-290): g^((p-1)/NthRoot) for the smallest primitive root g >= 3."""
    return FindPrimitiveRootOfUnity(p, NthRoot, 0)


def Derive3NRoots(q, a, b, omega):
    """ring.Derive3NRoots (ring/primes_3n.go:165-181): (omega, omega1 of order 2^a, omega2 of order 3^(b+1))."""
    out = (C.c_uint64 * 3)()
    _err(lib().b200_derive_3n_roots(int(q) & _M64, int(a), int(b), int(omega) & _M64, out))
    return int(out[0]), int(out[1]), int(out[2])


def Uint64Pow(base, exp):
    """ring.Uint64Pow (ring/primes_3n.go:183-193): wraps modulo 2^64 like the Go code."""
    return pow(int(base), int(exp), 1 << 64)


class _Gen(C.Structure):   # b200_prime_gen
    _fields_ = [("Size", C.c_double), ("NextPrime", C.c_uint64), ("PrevPrime", C.c_uint64), ("NthRoot", C.c_uint64),
                ("CheckNextPrime", C.c_int), ("CheckPrevPrime", C.c_int)]


class NTTFriendlyPrimesGenerator:
    """ring.NTTFriendlyPrimesGenerator (ring/primes.go:69-277): primes 2^BitSize +/- k*NthRoot + 1."""

    def __init__(self, BitSize, NthRoot):
        self._g = _Gen()
        _err(lib().b200_prime_gen_init(C.byref(self._g), int(BitSize), int(NthRoot)))

    def _next(self, direction, k):
        out = np.zeros(max(int(k), 1), dtype=np.uint64)
        _err(lib().b200_prime_gen_next(C.byref(self._g), direction, int(k), out.ctypes.data_as(capi.P64)))
        return [int(q) for q in out[:k]]

    def NextUpstreamPrimes(self, k):
        return self._next(1, k)

    def NextDownstreamPrimes(self, k):
        return self._next(-1, k)

    def NextAlternatingPrimes(self, k):
        return self._next(0, k)

    def NextUpstreamPrime(self):
        return self._next(1, 1)[0]

    def NextDownstreamPrime(self):
        return self._next(-1, 1)[0]

    def NextAlternatingPrime(self):
        return self._next(0, 1)[0]


def NewNTTFriendlyPrimesGenerator(BitSize, NthRoot):
    return NTTFriendlyPrimesGenerator(BitSize, NthRoot)
