"""ctypes loader for libabc_b200.so (the C ABI of include/abc_b200.h).

The library is built in-tree by __graft_entry__.build() (nvcc, sm_100a).  If it is missing this module
raises ImportError: there is no Python/CPU fallback for the ciphertext path.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("ABC_B200_LIB") or os.path.join(_HERE, "lib", "libabc_b200.so")  # override: A/B builds of the same ABI

# every symbol include/abc_b200.h declares: name -> (restype, argtypes)
vp, sz, u64, u32, i32, f32p = C.c_void_p, C.c_size_t, C.c_uint64, C.c_uint32, C.c_int, C.POINTER(C.c_float)
vpp = C.POINTER(C.c_void_p)


class AbcParams(C.Structure):
    _fields_ = [("poly_degree", u32), ("n_primes", u32), ("primes", C.POINTER(u64)), ("plain_modulus", u64),
                ("device", C.c_int32), ("batch", u32), ("seed", u64)]


SYMBOLS = {
    "abc_ctx_create": (i32, [C.POINTER(AbcParams), vpp]),
    "abc_ctx_destroy": (None, [vp]),
    "abc_last_error": (C.c_char_p, [vp]),
    "abc_sync": (i32, [vp]),
    "abc_host_alloc": (i32, [vp, sz, vpp]),
    "abc_host_free": (None, [vp]),
    "abc_host_register": (i32, [vp, vp, sz]),
    "abc_host_unregister": (None, [vp]),
    "abc_faulted": (i32, [vp]),
    "abc_clear_fault": (i32, [vp]),
    "abc_poly_degree": (u32, [vp]),
    "abc_n_primes": (u32, [vp]),
    "abc_n_limbs": (u32, [vp]),
    "abc_batch": (u32, [vp]),
    "abc_plain_modulus": (u64, [vp]),
    "abc_get_primes": (i32, [vp, vp]),
    "abc_get_aux_primes": (i32, [vp, vp, C.POINTER(u32)]),
    "abc_keygen": (i32, [vp]),
    "abc_keygen_select": (i32, [vp, vp, sz]),
    "abc_galois_elt_from_step": (u32, [vp, i32]),
    "abc_key_words": (sz, [vp, i32]),
    "abc_key_export": (i32, [vp, i32, u32, vp, sz]),
    "abc_key_import": (i32, [vp, i32, u32, vp, sz]),
    "abc_has_galois_key": (i32, [vp, u32]),
    "abc_ct_alloc": (i32, [vp, vpp]),
    "abc_ct_free": (None, [vp]),
    "abc_ct_clone": (i32, [vp, vp, vpp]),
    "abc_ct_words": (sz, [vp]),
    "abc_ct_export": (i32, [vp, vp, vp, sz]),
    "abc_ct_import": (i32, [vp, vp, vp, sz]),
    "abc_encode_encrypt": (i32, [vp, vp, sz, i32, vpp]),
    "abc_decrypt_decode": (i32, [vp, vp, vp]),
    "abc_decrypt_decode_async": (i32, [vp, vp, vp]),
    "abc_decrypt_wait": (i32, [vp]),
    "abc_set_encrypt_nonce": (i32, [vp, u64]),
    "abc_set_rng_key": (i32, [vp, vp]),
    "abc_noise_budget": (i32, [vp, vp, vp]),
    "abc_is_transparent": (i32, [vp, vp, vp]),
    "abc_add": (i32, [vp, vp, vp, vp]),
    "abc_sub": (i32, [vp, vp, vp, vp]),
    "abc_negate": (i32, [vp, vp, vp]),
    "abc_mul_relin": (i32, [vp, vp, vp, vp]),
    "abc_rotate_rows": (i32, [vp, vp, vp, i32]),
    "abc_rotate_rows_add": (i32, [vp, vp, vp, i32, vp]),
    "abc_ct_shared": (i32, [vp]),
    "abc_ct_deferred": (i32, [vp]),
    "abc_add_plain": (i32, [vp, vp, vp, vp, sz, i32]),
    "abc_sub_plain": (i32, [vp, vp, vp, vp, sz, i32]),
    "abc_mul_plain": (i32, [vp, vp, vp, vp, sz, i32]),
    "abc_pt_encode": (i32, [vp, vp, sz, i32, vpp]),
    "abc_pt_free": (None, [vp]),
    "abc_add_plain_pt": (i32, [vp, vp, vp, vp]),
    "abc_sub_plain_pt": (i32, [vp, vp, vp, vp]),
    "abc_mul_plain_pt": (i32, [vp, vp, vp, vp]),
    "abc_encrypt_pt": (i32, [vp, vp, vpp]),
    "abc_probe_ntt": (i32, [vp, i32, u32, vp, sz]),
    "abc_probe_multiply": (i32, [vp, vp, vp, vp, sz]),
    "abc_bench_ntt": (i32, [vp, i32, u32, sz, i32, f32p]),
    "abc_comm_unique_id": (i32, [vp, vp]),
    "abc_comm_init": (i32, [vp, i32, i32, vp]),
    "abc_comm_stats": (i32, [vp, C.POINTER(u64), C.POINTER(u64)]),
    "abc_comm_rank": (i32, [vp]),
    "abc_comm_world": (i32, [vp]),
    "abc_owned_limbs": (i32, [vp, C.POINTER(u32), C.POINTER(u32)]),
    "abc_ct_allgather": (i32, [vp, vp]),
    "abc_timer_start": (i32, [vp]),
    "abc_timer_stop": (i32, [vp, f32p]),
    "abc_flush_l2": (i32, [vp, sz]),
    "abc_launch_count": (u64, [vp]),
    "abc_key_switch_count": (u64, [vp]),
    "abc_profile_enable": (i32, [vp, i32]),
    "abc_profile_json": (C.c_char_p, [vp]),
    "abc_measure_int_peak": (i32, [vp, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "abc_measure_butterfly_peak": (i32, [vp, i32, C.POINTER(C.c_double)]),
    "abc_ntt_arith_class": (i32, [vp]),
    "abc_seal_parms_id": (i32, [vp, i32, vp]),
    "abc_seal_params_save": (i32, [vp, i32, vp, sz, C.POINTER(sz)]),
    "abc_seal_params_parse": (i32, [vp, sz, C.POINTER(AbcParams), vp, sz]),
    "abc_seal_ct_save": (i32, [vp, vp, u32, i32, vp, sz, C.POINTER(sz)]),
    "abc_seal_ct_load": (i32, [vp, vp, u32, vp, sz]),
    "abc_seal_key_save": (i32, [vp, i32, i32, vp, sz, C.POINTER(sz)]),
    "abc_seal_key_load": (i32, [vp, i32, vp, sz]),
    "abc_ct_export_instance": (i32, [vp, vp, u32, vp, sz]),
    "abc_ct_import_instance": (i32, [vp, vp, u32, vp, sz]),
    "abc_galois_elts": (i32, [vp, vp, sz, C.POINTER(sz)]),
}

_lib = None


def load():
    """Load the shared library and bind every declared symbol (no device call is made)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "abc_b200: %s is missing. Build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). There is no CPU fallback." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        f = getattr(lib, name)  # AttributeError if the .so lacks a declared entry point
        f.restype, f.argtypes = res, args
    _lib = lib
    return lib
