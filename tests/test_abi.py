"""CPU-only checks of the drop-in boundary: the C-ABI library loads, exports every symbol the header
declares, and the constructors validate exactly like ring.NewRing (ring/ring.go:314-356,
ring/subring.go:129-146; cases of ring/ring_test.go:128-167).  No compute calls."""
import os
import re

import pytest

from conftest import ROOT, QI60
import b200ring
from b200ring import capi


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "b200ring.h")).read()
    declared = set(re.findall(r"^B200_API [^;(]*?\b(b200_\w+)\(", hdr, re.M))
    assert len(declared) >= 25
    assert declared == set(capi.SYMBOLS), declared ^ set(capi.SYMBOLS)
    L = capi.load_library()           # binds every symbol or raises AttributeError
    for name in declared:
        assert getattr(L, name) is not None
    assert b"sm_100a" in L.b200_version()


def test_header_cites_reference_for_every_entry_point():
    hdr = open(os.path.join(ROOT, "include", "b200ring.h")).read()
    for needle in ("ring/ring.go:264-356", "ring/ntt.go:127-138", "ring/ntt.go:15-22", "ring/poly.go", "ring/operations.go"):
        assert needle in hdr


@pytest.mark.parametrize("N,moduli,msg", [
    (15, [QI60[0]], "invalid ring degree"),
    (4, [QI60[0]], "invalid ring degree"),
    (16, [], "invalid ModuliChain (must be a non-empty []uint64)"),
    (16, [QI60[0], QI60[0]], "moduli are not distinct"),
    (16, [QI60[0] + 2], "is not prime"),
    (1 << 21, [QI60[0]], "!= 1 mod NthRoot"),
    (24, [QI60[0]], "!= 1 mod NthRoot"),            # 3N ring needs q = 1 mod 72
])
def test_constructor_errors_like_reference(N, moduli, msg):
    with pytest.raises(ValueError) as e:
        b200ring.NewRing(N, moduli)
    assert msg in str(e.value)


def test_matrix_type_rejects_power_of_two_degree():
    with pytest.raises(ValueError) as e:
        b200ring.NewRingFromType(16, [QI60[0]], b200ring.Matrix)   # ring.go:301-303
    assert "matrix ring type requires" in str(e.value)


def test_no_cpu_fallback_without_device():
    if capi.lib().b200_device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(b200ring.B200Error) as e:
        b200ring.NewRing(16, [576460752303439873])
    assert e.value.code == capi.ENODEVICE and "no CPU fallback" in str(e.value)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "matrix-fhe-lattigo_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text.replace("no oracle", ""), os.path.join(dirpath, f)


def test_tools_never_import_the_oracle():
    """tools/ are measurement helpers of the product path: like the package they must not execute anything under oracle/
    (only tests/, smoke() and bench.py's CPU arms may)."""
    for f in os.listdir(os.path.join(ROOT, "tools")):
        if f.endswith(".py"):
            text = open(os.path.join(ROOT, "tools", f)).read()
            assert "from oracle" not in text and "import oracle" not in text, f


def test_null_arguments_are_errors_not_crashes():
    """Every entry point checks its handles before touching the device (the Go side would panic on a nil Poly)."""
    import ctypes as C
    L = capi.lib()
    out = C.c_uint64(0)
    assert L.b200_poly_copy(None, None, 0) == capi.B200_EINVAL
    assert L.b200_vec_op(None, 0, None, None, None, 0, None) == capi.B200_EINVAL
    assert L.b200_ntt(None, None, None, 0, 0) == capi.B200_EINVAL
    assert L.b200_div_by_last_modulus(None, None, None, 1, 0) == capi.B200_EINVAL
    assert L.b200_div_by_last_modulus_many(None, None, None, 1, 1, 0, 0) == capi.B200_EINVAL
    assert L.b200_mul_by_vector_montgomery(None, None, None, None, 0, 0) == capi.B200_EINVAL
    assert L.b200_mod_down(None, 0, 0, 0, None, None) == capi.B200_EINVAL
    assert L.b200_find_3n_rns_primes(12, 60, 1, 10, None) == capi.B200_EINVAL
    assert L.b200_order_mod(2, 7, None) == capi.B200_EINVAL
    assert L.b200_prime_gen_next(None, 0, 1, C.byref(out)) == capi.B200_EINVAL
    assert L.b200_last_error() != b""


@pytest.mark.parametrize("N", [1 << 10, 1 << 16, 3 << 10, 3 << 15])
def test_row_pass_order_is_a_permutation_of_the_row_subtrees(N):
    """The row-pass copies of the twiddle tables permute the heap indices [R, N): host-only entry point,
    no device.  Every row subtree {(R + r) * 2^d + k} lands in its own block [R + 255 r, R + 255 (r + 1)), first the 7
    entries of depths 0..2 in heap order, then depths 3..5 as runs of 8 indexed by the top three bits of k, then depths
    6, 7 as runs of 32 indexed by lane = (k >> (d - 6)) >> 1."""
    L = capi.load_library()
    R = N // 256
    assert L.b200_row_pass_index(N, 0) == 0 and L.b200_row_pass_index(N, R - 1) == R - 1      # column-pass indices stay
    assert L.b200_row_pass_index(N, N) == -1 and L.b200_row_pass_index(100, 5) == -1
    seen = set()
    for r in (0, 1, R // 2, R - 1):
        block = []
        for d in range(8):
            for k in range(1 << d):
                i = L.b200_row_pass_index(N, ((R + r) << d) + k)
                assert R + 255 * r <= i < R + 255 * (r + 1)
                block.append(i)
                if d < 3:
                    assert i == R + 255 * r + (1 << d) - 1 + k
                elif d < 6:
                    j = d - 3
                    assert i == R + 255 * r + 7 + ((1 << j) - 1 + (k & ((1 << j) - 1))) * 8 + (k >> j)
                else:
                    j = d - 6
                    sub = k >> j
                    assert i == R + 255 * r + 63 + ((sub & 1) * 3 + (1 << j) - 1 + (k & ((1 << j) - 1))) * 32 + (sub >> 1)
        assert sorted(block) == list(range(R + 255 * r, R + 255 * (r + 1)))
        seen.update(block)
    if N <= 3 << 10:
        assert sorted(L.b200_row_pass_index(N, h) for h in range(R, N)) == list(range(R, N))
