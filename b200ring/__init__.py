"""Import alias for the product package directory `matrix-fhe-lattigo_b200/` (a hyphen is not a valid
Python identifier).  `import b200ring` resolves sub-modules from that directory."""
import os as _os

_PKG = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "matrix-fhe-lattigo_b200")
__path__.insert(0, _PKG)

from .capi import lib, B200Error, load_library  # noqa: E402,F401
from .ring import (  # noqa: E402,F401
    Type, Standard, ConjugateInvariant, Matrix, Ring, SubRing, Poly, KeySwitcher, NewRing, NewRingFromType, NewRingConjugateInvariant, NewPoly, AutomorphismNTTIndex,
    UnmarshalJSON,
)
from .primes import (  # noqa: E402,F401
    IsPrime, Find3NRNSPrimes, Find3NFriendlyPrime, OrderMod, HasPrimitiveMthRoot, PrimitiveRoot, FindPrimitiveRootOfUnity,
    Find3NPrimitiveRoot, Derive3NRoots, Uint64Pow, NTTFriendlyPrimesGenerator, NewNTTFriendlyPrimesGenerator,
)
from . import scalar  # noqa: E402,F401
from . import sharding  # noqa: E402,F401
