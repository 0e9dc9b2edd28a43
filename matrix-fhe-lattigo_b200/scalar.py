"""ring.RNSScalar (ring/scalar.go) and the word-level primitives it uses (ring/modular_reduction.go, ring/utils.go:29-69): host
arithmetic on one word per modulus, restated bit for bit with Python integers.  These feed the `scalars` argument of
b200_vec_op; they never touch the device, so they take the moduli chain instead of a ring handle (Ring.* wrappers in ring.py)."""
_M = (1 << 64) - 1


def MRedConstant(q):   # GenMRedConstant, modular_reduction.go:68-75: q^-1 mod 2^64
    return pow(q, -1, 1 << 64)


def BRedConstant(q):   # GenBRedConstant, :99-107: floor(2^128 / q) as (hi, lo)
    u = (1 << 128) // q
    return (u >> 64, u & _M)


def CRed(a, q):   # :200-205
    return a - q if a >= q else a


def MRedLazy(x, y, q, qinv):   # :90-95, result in [0, 2q-1]
    m = x * y
    h = (((m & _M) * qinv & _M) * q) >> 64
    return ((m >> 64) - h + q) & _M


def MRed(x, y, q, qinv):   # :78-86
    return CRed(MRedLazy(x, y, q, qinv), q)


def MForm(a, q, brc=None):   # :11-35: a * 2^64 mod q for any 64-bit a
    hi, lo = brc if brc is not None else BRedConstant(q)
    r = (-(a * hi + ((a * lo) >> 64)) * q) & _M
    return CRed(r, q)


def IMForm(a, q, qinv):   # :49-56
    return CRed(q - ((((a * qinv) & _M) * q) >> 64), q)


def BRedAdd(a, q, brc=None):   # :110-117
    hi = (brc if brc is not None else BRedConstant(q))[0]
    return CRed((a - ((a * hi) >> 64) * q) & _M, q)


def ModExp(x, e, p):   # ring/utils.go:29-39
    return pow(x % p, e, p)


def ModexpMontgomery(x, e, q, qinv, brc=None):   # ring/utils.go:56-69: Montgomery form in and out
    result = MForm(1, q, brc)
    while e > 0:
        if e & 1:
            result = MRed(result, x, q, qinv)
        x = MRed(x, x, q, qinv)
        e >>= 1
    return result


# ---- ring/scalar.go: one word per modulus of the chain (levels 0..level = the first level+1 moduli)
def NewRNSScalar(moduli):
    return [0] * len(moduli)


def NewRNSScalarFromUInt64(moduli, v):   # :16-22
    return [(int(v) & _M) % q for q in moduli]


def NewRNSScalarFromBigint(moduli, v):   # :25-33 (big.Int.Mod: Euclidean, non-negative)
    return [int(v) % q for q in moduli]


def MFormRNSScalar(moduli, s1):   # :36-40
    return [MForm(a, q) for a, q in zip(s1, moduli)]


def NegRNSScalar(moduli, s1):   # :43-47 (q - s: the Go code leaves q for a zero residue)
    return [q - a for a, q in zip(s1, moduli)]


def SubRNSScalar(moduli, s1, s2):   # :50-58
    return [a + q - b if b > a else a - b for a, b, q in zip(s1, s2, moduli)]


def MulRNSScalar(moduli, s1, s2):   # :62-66: MRedLazy, result in [0, 2q-1]
    return [MRedLazy(a, b, q, MRedConstant(q)) for a, b, q in zip(s1, s2, moduli)]


def Inverse(moduli, a):   # :69-74: a^(q-2) in Montgomery form, a in Montgomery form
    return [ModexpMontgomery(x, q - 2, q, MRedConstant(q)) for x, q in zip(a, moduli)]
