"""ctypes binding of libbellman_b200.so (the C ABI in include/bellman_b200.h).

There is no fallback: if the shared library is missing or no CUDA device is usable the
import / context creation fails loudly.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# BMPC_LIB_PATH: another build of the same library (A/B runs of kernel variants)
LIB_PATH = os.environ.get("BMPC_LIB_PATH") or os.path.join(_HERE, "libbellman_b200.so")

(OK, ERR_UNEXPECTED_IDENTITY, ERR_UNEXPECTED_EOF, ERR_DEGREE_TOO_LARGE, ERR_LENGTH_MISMATCH, ERR_CUDA, ERR_INVALID,
 ERR_INVALID_DATA) = range(8)
G1, G2 = 1, 2
FORM_UNCOMPRESSED_BE, FORM_MONT_XY = 0, 1
FFT, IFFT, COSET_FFT, ICOSET_FFT = 0, 1, 2, 3

# every symbol include/bellman_b200.h declares (checked by tests/test_abi.py)
EXPORTS = [
    "bmpc_ctx_create", "bmpc_ctx_destroy", "bmpc_last_error", "bmpc_ctx_set_tuning",
    "bmpc_ctx_reload_env", "bmpc_ctx_launch_count", "bmpc_ctx_profile", "bmpc_ctx_profile_read",
    "bmpc_bases_register", "bmpc_bases_register_dev", "bmpc_bases_precompute", "bmpc_bases_len", "bmpc_bases_group",
    "bmpc_bases_read", "bmpc_bases_dev_ptr", "bmpc_bases_free",
    "bmpc_multiexp", "bmpc_multiexp_dev", "bmpc_multiexp_partial_dev", "bmpc_multiexp_shard_dev",
    "bmpc_msm_flags_status", "bmpc_sum_partials",
    "bmpc_multiexp_shard_enqueue_dev", "bmpc_shard_record_bytes", "bmpc_fold_shard_records",
    "bmpc_partial_bytes", "bmpc_msm_geometry", "bmpc_msm_accumulate_info",
    "bmpc_domain_from_coeffs", "bmpc_domain_from_coeffs_dev", "bmpc_domain_len", "bmpc_domain_exp",
    "bmpc_domain_into_coeffs", "bmpc_domain_dev_ptr", "bmpc_domain_free", "bmpc_domain_transform",
    "bmpc_domain_distribute_powers", "bmpc_domain_z", "bmpc_domain_divide_by_z_on_coset",
    "bmpc_domain_mul_assign", "bmpc_domain_sub_assign", "bmpc_ntt_dev", "bmpc_ntt",
    "bmpc_ntt_batch_dev", "bmpc_fr_swap01_dev", "bmpc_ntt_fourstep_twiddle_dev", "bmpc_fr_scale_pow_dev",
    "bmpc_h_coefficients", "bmpc_h_coefficients_dev", "bmpc_h_coset_evals_dev", "bmpc_h_from_coset_evals_dev",
    "bmpc_fr_to_canonical_dev",
    "bmpc_create_proof", "bmpc_create_proof_partials", "bmpc_create_proof_finish", "bmpc_batch_scalar_mul", "bmpc_fixed_base_mul",
    "bmpc_list_mul_matrix",
    "bmpc_params_read", "bmpc_params_write", "bmpc_params_free",
    "bmpc_r1cs_eval", "bmpc_generate_parameters",
    "bmpc_multiexp_async", "bmpc_waiter_wait",
    "bmpc_multi_create", "bmpc_multi_destroy", "bmpc_multi_size", "bmpc_multi_ctx", "bmpc_multi_last_error",
    "bmpc_multi_bases_register", "bmpc_multi_bases_precompute", "bmpc_multi_bases_len", "bmpc_multi_bases_part",
    "bmpc_multi_bases_free", "bmpc_multi_multiexp", "bmpc_multi_create_proof",
]


class Params(C.Structure):
    """bmpc_params"""
    _fields_ = [("h", C.c_void_p), ("l", C.c_void_p), ("a", C.c_void_p), ("b_g1", C.c_void_p),
                ("b_g2", C.c_void_p),
                ("alpha_g1", C.c_uint8 * 96), ("beta_g1", C.c_uint8 * 96), ("beta_g2", C.c_uint8 * 192),
                ("delta_g1", C.c_uint8 * 96), ("delta_g2", C.c_uint8 * 192)]


class Assignment(C.Structure):
    """bmpc_assignment"""
    _fields_ = [("a", C.c_void_p), ("b", C.c_void_p), ("c", C.c_void_p), ("num_constraints", C.c_size_t),
                ("input_assignment", C.c_void_p), ("num_inputs", C.c_size_t),
                ("aux_assignment", C.c_void_p), ("num_aux", C.c_size_t),
                ("a_aux_density", C.c_void_p), ("b_input_density", C.c_void_p),
                ("b_aux_density", C.c_void_p)]


class MultiParams(C.Structure):
    """bmpc_multi_params"""
    _fields_ = Params._fields_


PROOF_PARTIAL_BYTES = 1920      # BMPC_PROOF_PARTIAL_BYTES: 6 G1 XYZZ (192 B) + 2 G2 XYZZ (384 B)


class ProofShard(C.Structure):
    """bmpc_proof_shard"""
    _fields_ = [("base_offset", C.c_size_t * 8), ("h_lo", C.c_size_t), ("h_hi", C.c_size_t),
                ("n_total", C.c_size_t * 8)]


class Csr(C.Structure):
    """bmpc_csr"""
    _fields_ = [("row_ptr", C.c_void_p), ("col", C.c_void_p), ("coeff", C.c_void_p),
                ("num_rows", C.c_size_t), ("nnz", C.c_size_t)]


class ParametersFile(C.Structure):
    """bmpc_parameters"""
    _fields_ = [("p", Params), ("gamma_g2", C.c_uint8 * 192), ("ic", C.c_void_p)]


_lib = None


def load():
    """Load the shared library (built in-tree by `make` / __graft_entry__.build())."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `make -j8` (or __graft_entry__.build()); "
            "bellman_mpc_b200 has no CPU fallback")
    lib = C.CDLL(LIB_PATH)
    vp, sz, u32, i32 = C.c_void_p, C.c_size_t, C.c_uint32, C.c_int
    sig = {
        "bmpc_ctx_create": (i32, [i32, C.POINTER(vp)]),
        "bmpc_ctx_destroy": (None, [vp]),
        "bmpc_last_error": (C.c_char_p, [vp]),
        "bmpc_ctx_set_tuning": (i32, [vp, i32, i32]),
        "bmpc_ctx_reload_env": (i32, [vp]),
        "bmpc_ctx_launch_count": (C.c_uint64, [vp]),
        "bmpc_ctx_profile": (i32, [vp, i32]),
        "bmpc_ctx_profile_read": (i32, [vp, i32, C.POINTER(C.c_double), C.POINTER(C.c_uint64)]),
        "bmpc_bases_register": (i32, [vp, i32, vp, sz, sz, i32, C.POINTER(vp)]),
        "bmpc_bases_register_dev": (i32, [vp, i32, vp, sz, C.POINTER(vp), vp]),
        "bmpc_bases_precompute": (i32, [vp, vp, i32]),
        "bmpc_bases_len": (sz, [vp]),
        "bmpc_bases_group": (i32, [vp]),
        "bmpc_bases_read": (i32, [vp, vp, sz, sz, vp]),
        "bmpc_bases_dev_ptr": (vp, [vp]),
        "bmpc_bases_free": (None, [vp, vp]),
        "bmpc_multiexp": (i32, [vp, vp, sz, vp, sz, vp, sz, vp]),
        "bmpc_multiexp_dev": (i32, [vp, vp, sz, vp, sz, vp, sz, vp, vp]),
        "bmpc_multiexp_partial_dev": (i32, [vp, vp, sz, vp, sz, vp, sz, vp, vp]),
        "bmpc_multiexp_shard_dev": (i32, [vp, vp, sz, vp, sz, vp, sz, sz, vp, C.POINTER(u32), vp]),
        "bmpc_msm_flags_status": (i32, [u32]),
        "bmpc_multiexp_shard_enqueue_dev": (i32, [vp, vp, sz, vp, sz, vp, sz, sz, vp, vp]),
        "bmpc_shard_record_bytes": (sz, [i32]),
        "bmpc_fold_shard_records": (i32, [vp, i32, vp, sz, sz, vp, C.POINTER(u32), vp]),
        "bmpc_sum_partials": (i32, [vp, i32, vp, sz, vp, vp]),
        "bmpc_partial_bytes": (sz, [i32]),
        "bmpc_msm_geometry": (i32, [vp, vp, sz, C.POINTER(u32), C.POINTER(u32), C.POINTER(u32)]),
        "bmpc_msm_accumulate_info": (i32, [vp, vp, sz, C.POINTER(u32 * 8)]),
        "bmpc_domain_from_coeffs": (i32, [vp, vp, sz, C.POINTER(vp)]),
        "bmpc_domain_from_coeffs_dev": (i32, [vp, vp, sz, C.POINTER(vp), vp]),
        "bmpc_domain_len": (sz, [vp]),
        "bmpc_domain_exp": (u32, [vp]),
        "bmpc_domain_into_coeffs": (i32, [vp, vp, vp]),
        "bmpc_domain_dev_ptr": (vp, [vp]),
        "bmpc_domain_free": (None, [vp, vp]),
        "bmpc_domain_transform": (i32, [vp, vp, i32, vp]),
        "bmpc_domain_distribute_powers": (i32, [vp, vp, vp, vp]),
        "bmpc_domain_z": (i32, [vp, vp, vp, vp]),
        "bmpc_domain_divide_by_z_on_coset": (i32, [vp, vp, vp]),
        "bmpc_domain_mul_assign": (i32, [vp, vp, vp, vp]),
        "bmpc_domain_sub_assign": (i32, [vp, vp, vp, vp]),
        "bmpc_ntt_dev": (i32, [vp, vp, u32, i32, vp]),
        "bmpc_ntt": (i32, [vp, vp, u32, i32]),
        "bmpc_ntt_batch_dev": (i32, [vp, vp, u32, u32, i32, vp]),
        "bmpc_fr_swap01_dev": (i32, [vp, vp, vp, u32, u32, u32, vp]),
        "bmpc_ntt_fourstep_twiddle_dev": (i32, [vp, vp, u32, u32, u32, u32, i32, vp]),
        "bmpc_fr_scale_pow_dev": (i32, [vp, vp, sz, u32, u32, i32, vp]),
        "bmpc_h_coefficients": (i32, [vp, vp, vp, vp, sz, vp, C.POINTER(sz)]),
        "bmpc_h_coefficients_dev": (i32, [vp, vp, vp, vp, u32, vp]),
        "bmpc_fr_to_canonical_dev": (i32, [vp, vp, sz, vp]),
        "bmpc_h_coset_evals_dev": (i32, [vp, vp, u32, vp, sz, vp]),
        "bmpc_h_from_coset_evals_dev": (i32, [vp, vp, vp, vp, u32, vp]),
        "bmpc_create_proof": (i32, [vp, C.POINTER(Params), C.POINTER(Assignment), vp, vp, vp]),
        "bmpc_create_proof_partials": (i32, [vp, C.POINTER(Params), C.POINTER(Assignment), C.POINTER(ProofShard), vp,
                                             C.POINTER(u32 * 8)]),
        "bmpc_create_proof_finish": (i32, [vp, C.POINTER(Params), vp, sz, vp, vp, vp]),
        "bmpc_batch_scalar_mul": (i32, [vp, vp, vp, i32, C.POINTER(vp)]),
        "bmpc_fixed_base_mul": (i32, [vp, i32, vp, vp, sz, i32, C.POINTER(vp)]),
        "bmpc_list_mul_matrix": (i32, [vp, vp, vp, vp, vp, sz, C.POINTER(vp)]),
        "bmpc_params_read": (i32, [vp, vp, sz, i32, C.POINTER(ParametersFile)]),
        "bmpc_params_write": (i32, [vp, C.POINTER(ParametersFile), vp, sz, C.POINTER(sz)]),
        "bmpc_params_free": (None, [vp, C.POINTER(ParametersFile)]),
        "bmpc_r1cs_eval": (i32, [vp, C.POINTER(Csr), C.POINTER(Csr), C.POINTER(Csr), sz, sz, vp, vp, vp, vp, vp, vp, vp, vp]),
        "bmpc_generate_parameters": (i32, [vp, C.POINTER(Csr), C.POINTER(Csr), C.POINTER(Csr), sz, sz, sz, vp, vp,
                                           vp, vp, vp, vp, vp, C.POINTER(ParametersFile)]),
        "bmpc_multiexp_async": (i32, [vp, vp, sz, vp, sz, vp, sz, C.POINTER(vp)]),
        "bmpc_waiter_wait": (i32, [vp, vp]),
        "bmpc_multi_create": (i32, [C.POINTER(i32), i32, C.POINTER(vp)]),
        "bmpc_multi_destroy": (None, [vp]),
        "bmpc_multi_size": (i32, [vp]),
        "bmpc_multi_ctx": (vp, [vp, i32]),
        "bmpc_multi_last_error": (C.c_char_p, [vp]),
        "bmpc_multi_bases_register": (i32, [vp, i32, vp, sz, sz, i32, C.POINTER(vp)]),
        "bmpc_multi_bases_precompute": (i32, [vp, vp, i32]),
        "bmpc_multi_bases_len": (sz, [vp]),
        "bmpc_multi_bases_part": (vp, [vp, i32, C.POINTER(sz)]),
        "bmpc_multi_bases_free": (None, [vp, vp]),
        "bmpc_multi_multiexp": (i32, [vp, vp, sz, vp, sz, vp, sz, vp]),
        "bmpc_multi_create_proof": (i32, [vp, C.POINTER(MultiParams), C.POINTER(Assignment), vp, vp, vp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib
