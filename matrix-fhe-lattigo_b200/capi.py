"""ctypes binding of include/b200ring.h.  There is no fallback: if the CUDA library is missing the
import of any compute symbol raises, and without a GPU every compute call returns B200_ENODEVICE."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("B200RING_LIB") or os.path.join(_HERE, "lib", "libb200ring.so")   # env: A/B tuning builds

u64 = C.c_uint64
P64 = C.POINTER(C.c_uint64)
VP = C.c_void_p

# every symbol include/b200ring.h declares: name -> (restype, argtypes)
SYMBOLS = {
    "b200_last_error": (C.c_char_p, []),
    "b200_version": (C.c_char_p, []),
    "b200_device_count": (C.c_int, []),
    "b200_row_pass_index": (C.c_int, [C.c_int, C.c_int]),
    "b200_ctx_create": (C.c_int, [C.c_int, C.c_int, C.c_int, P64, P64, C.c_int, C.POINTER(VP)]),
    "b200_ctx_destroy": (None, [VP]),
    "b200_ctx_set_stream": (C.c_int, [VP, VP]),
    "b200_sync": (C.c_int, [VP]),
    "b200_ctx_n": (C.c_int, [VP]),
    "b200_ctx_moduli_count": (C.c_int, [VP]),
    "b200_ctx_ring_type": (C.c_int, [VP]),
    "b200_ctx_subring": (C.c_int, [VP, C.c_int, P64]),
    "b200_ctx_roots": (C.c_int, [VP, C.c_int, C.c_int, P64, C.c_size_t]),
    "b200_ctx_factors": (C.c_int, [VP, C.c_int, P64, C.c_int]),
    "b200_poly_ipc_export": (C.c_int, [VP, C.c_char_p]),
    "b200_peer_open": (C.c_int, [VP, C.c_char_p, C.POINTER(VP)]),
    "b200_peer_close": (C.c_int, [VP, VP]),
    "b200_intt_limbs_gather": (C.c_int, [VP, VP, VP, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(VP)]),
    "b200_poly_binary_size": (C.c_int, [VP, C.c_int, C.POINTER(C.c_size_t)]),
    "b200_poly_marshal": (C.c_int, [VP, C.c_int, C.c_int, VP, C.c_size_t, C.POINTER(C.c_size_t)]),
    "b200_poly_unmarshal": (C.c_int, [VP, C.c_int, VP, C.c_size_t, C.POINTER(C.c_int), C.POINTER(C.c_size_t)]),
    "b200_poly_alloc": (C.c_int, [VP, C.c_int, C.POINTER(VP)]),
    "b200_poly_wrap": (C.c_int, [VP, VP, C.c_int, C.POINTER(VP)]),
    "b200_poly_free": (None, [VP]),
    "b200_poly_device_ptr": (VP, [VP]),
    "b200_poly_batch": (C.c_int, [VP]),
    "b200_poly_upload_limb": (C.c_int, [VP, C.c_int, C.c_int, P64, C.c_size_t]),
    "b200_poly_download_limb": (C.c_int, [VP, C.c_int, C.c_int, P64, C.c_size_t]),
    "b200_poly_upload": (C.c_int, [VP, VP, C.c_size_t]),
    "b200_poly_download": (C.c_int, [VP, VP, C.c_size_t]),
    "b200_poly_copy": (C.c_int, [VP, VP, C.c_int]),
    "b200_ntt": (C.c_int, [VP, VP, VP, C.c_int, C.c_int]),
    "b200_intt": (C.c_int, [VP, VP, VP, C.c_int, C.c_int]),
    "b200_vec_op": (C.c_int, [VP, C.c_int, VP, VP, VP, C.c_int, P64]),
    "b200_mul_by_vector_montgomery": (C.c_int, [VP, VP, VP, VP, C.c_int, C.c_int]),
    "b200_keyswitch_create": (C.c_int, [VP, C.c_int, C.POINTER(VP)]),
    "b200_keyswitch_destroy": (None, [VP]),
    "b200_mod_up": (C.c_int, [VP, C.c_int, C.c_int, C.c_int, VP, VP]),
    "b200_mod_down_ntt": (C.c_int, [VP, C.c_int, C.c_int, VP, VP]),
    "b200_mod_down": (C.c_int, [VP, C.c_int, C.c_int, C.c_int, VP, VP]),
    "b200_gadget_product": (C.c_int, [VP, C.c_int, C.c_int, VP, VP, VP, VP]),
    "b200_decompose_ntt": (C.c_int, [VP, C.c_int, C.c_int, VP, VP]),
    "b200_gadget_product_hoisted": (C.c_int, [VP, C.c_int, C.c_int, VP, VP, VP, VP]),
    "b200_automorphism_ct": (C.c_int, [VP, C.c_int, C.c_int, VP, VP, VP, C.c_uint64, VP, VP, VP]),
    "b200_gadget_product_bit_decomp": (C.c_int, [VP, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int), VP, VP, VP, VP]),
    "b200_mul_relin": (C.c_int, [VP, C.c_int, C.c_int, VP, VP, VP, VP, VP, VP, VP, C.c_int]),
    "b200_ntt_limbs": (C.c_int, [VP, VP, VP, C.c_int, C.c_int, C.c_int, C.c_int]),
    "b200_ks_decompose_mac": (C.c_int, [VP, C.c_int, C.c_int, VP, VP, VP, VP, VP, C.c_int, C.c_int, C.c_int, C.c_int]),
    "b200_ks_mod_down_limbs": (C.c_int, [VP, C.c_int, C.c_int, VP, VP, C.c_int, C.c_int]),
    "b200_div_by_last_modulus_ntt": (C.c_int, [VP, VP, VP, C.c_int, C.c_int]),
    "b200_div_by_last_modulus": (C.c_int, [VP, VP, VP, C.c_int, C.c_int]),
    "b200_div_by_last_modulus_many": (C.c_int, [VP, VP, VP, C.c_int, C.c_int, C.c_int, C.c_int]),
    "b200_tensor": (C.c_int, [VP, VP, VP, VP, VP, VP, VP, VP, C.c_int, C.c_int]),
    "b200_polymul": (C.c_int, [VP, VP, VP, VP, C.c_int]),
    "b200_shift": (C.c_int, [VP, VP, C.c_int, VP]),
    "b200_mult_by_monomial": (C.c_int, [VP, VP, C.c_int, VP, C.c_int]),
    "b200_unfold_ci_to_standard": (C.c_int, [VP, VP, VP, C.c_int]),
    "b200_automorphism_ntt_index": (C.c_int, [C.c_int, C.c_uint64, C.c_uint64, P64]),
    "b200_automorphism_ntt": (C.c_int, [VP, VP, C.c_uint64, VP, C.c_int, C.c_int]),
    "b200_automorphism_ntt_with_index": (C.c_int, [VP, VP, P64, C.c_size_t, VP, C.c_int, C.c_int]),
    "b200_automorphism": (C.c_int, [VP, VP, C.c_uint64, VP, C.c_int]),
    "b200_fold_standard_to_ci": (C.c_int, [VP, VP, P64, C.c_size_t, VP, C.c_int]),
    "b200_transform_host": (C.c_int, [VP, VP, VP, C.c_int, C.c_int, C.POINTER(C.c_int), C.c_int, C.c_int]),
    "b200_subring_ntt_host": (C.c_int, [VP, C.c_int, P64, P64, C.c_size_t, C.c_int, C.c_int]),
    "b200_is_prime": (C.c_int, [C.c_uint64]),
    "b200_find_3n_rns_primes": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, P64]),
    "b200_find_3n_friendly_prime": (C.c_int, [C.c_int, C.c_int, P64]),
    "b200_order_mod": (C.c_int, [C.c_uint64, C.c_uint64, P64]),
    "b200_has_primitive_mth_root": (C.c_int, [C.c_uint64, C.c_uint64]),
    "b200_primitive_root": (C.c_int, [C.c_uint64, P64]),
    "b200_find_primitive_root_of_unity": (C.c_int, [C.c_uint64, C.c_uint64, C.c_uint64, P64]),
    "b200_derive_3n_roots": (C.c_int, [C.c_uint64, C.c_int, C.c_int, C.c_uint64, P64]),
    "b200_prime_gen_init": (C.c_int, [VP, C.c_uint64, C.c_uint64]),
    "b200_prime_gen_next": (C.c_int, [VP, C.c_int, C.c_int, P64]),
    "b200_ctx_launch_count": (C.c_uint64, [VP]),
    "b200_profile_begin": (C.c_int, [VP]),
    "b200_profile_end": (C.c_int, [VP, C.c_char_p, C.c_size_t]),
}

B200_LAZY, B200_TREE_ORDER = 1, 2
ENODEVICE = -2
B200_EINVAL, B200_ENODEVICE, B200_ECUDA, B200_ESHAPE, B200_ENOMEM = -1, -2, -3, -4, -5


class B200Error(RuntimeError):
    """A failed C-ABI call: `.code` is the B200_E* value, the message is b200_last_error()."""

    def __init__(self, code, msg):
        super().__init__(msg)
        self.code = code


_lib = None


def load_library(path=None):
    """dlopen the CUDA library and bind every declared symbol; raises if it was not built."""
    global _lib
    if _lib is None:
        path = path or LIB_PATH
        if not os.path.exists(path):
            raise ImportError(
                "b200ring: CUDA extension %s is missing -- build it with `python matrix-fhe-lattigo_b200/build.py` "
                "(there is no CPU fallback)" % path)
        L = C.CDLL(path)
        for name, (res, args) in SYMBOLS.items():
            f = getattr(L, name)   # AttributeError if the header and the library disagree
            f.restype, f.argtypes = res, args
        _lib = L
    return _lib


def lib():
    return load_library()


def check(rc):
    if rc != 0:
        raise B200Error(rc, lib().b200_last_error().decode())
    return rc
