"""ctypes binding of libapsu_b200.so (include/apsu_b200.h).

The library is the product: if it is missing or no CUDA device is present, everything here raises —
there is no CPU fallback and nothing in this package imports the oracle.
"""
from __future__ import annotations

import ctypes as C
import pathlib

import numpy as np

_HERE = pathlib.Path(__file__).resolve().parent
LIB_PATH = _HERE / "libapsu_b200.so"

MAX_COEFF_MODULUS = 8
MAX_QUERY_POWERS = 1024

OK, ERR_INVALID_ARGUMENT, ERR_LOGIC, ERR_RUNTIME, ERR_CUDA = 0, -1, -2, -3, -4


class CParams(C.Structure):
    _fields_ = [
        ("poly_modulus_degree", C.c_uint32),
        ("coeff_modulus_count", C.c_uint32),
        ("plain_modulus", C.c_uint64),
        ("coeff_modulus", C.c_uint64 * MAX_COEFF_MODULUS),
        ("hash_func_count", C.c_uint32),
        ("table_size", C.c_uint32),
        ("max_items_per_bin", C.c_uint32),
        ("felts_per_item", C.c_uint32),
        ("ps_low_degree", C.c_uint32),
        ("query_power_count", C.c_uint32),
        ("query_powers", C.c_uint32 * MAX_QUERY_POWERS),
        ("item_bit_count_per_felt", C.c_uint32),
        ("item_bit_count", C.c_uint32),
        ("items_per_bundle", C.c_uint32),
        ("bins_per_bundle", C.c_uint32),
        ("bundle_idx_count", C.c_uint32),
    ]


class CTimings(C.Structure):
    _fields_ = [
        ("compute_powers_ms", C.c_float),
        ("eval_ms", C.c_float),
        ("db_stream_ms", C.c_float),
        ("db_stream_bytes", C.c_uint64),
        ("db_stream_launches", C.c_uint32),
        ("kernel_launches", C.c_uint32),
    ]


class CudaUnavailable(RuntimeError):
    """raised for APSU_B200_ERR_CUDA (no device / CUDA failure): the path has no CPU fallback"""


_EXC = {ERR_INVALID_ARGUMENT: ValueError, ERR_LOGIC: AssertionError, ERR_RUNTIME: RuntimeError, ERR_CUDA: CudaUnavailable}

_lib = None

u64p = np.ctypeslib.ndpointer(dtype=np.uint64, flags="C_CONTIGUOUS")
u32p = np.ctypeslib.ndpointer(dtype=np.uint32, flags="C_CONTIGUOUS")
vp = C.c_void_p

# name -> (restype, argtypes); every symbol include/apsu_b200.h declares
SIGNATURES = {
    "apsu_b200_last_error": (C.c_char_p, []),
    "apsu_b200_version": (C.c_char_p, []),
    "apsu_b200_params_load_json": (C.c_int, [C.c_char_p, C.POINTER(CParams)]),
    "apsu_b200_params_validate": (C.c_int, [C.POINTER(CParams)]),
    "apsu_b200_coeff_modulus_create": (C.c_int, [C.c_uint32, C.POINTER(C.c_int), C.c_uint32, u64p]),
    "apsu_b200_plain_modulus_batching": (C.c_int, [C.c_uint32, C.c_int, C.POINTER(C.c_uint64)]),
    "apsu_b200_powers_dag": (C.c_int, [C.POINTER(CParams), C.c_uint32, u32p, u32p, u32p, u32p, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]),
    "apsu_b200_ctx_create": (C.c_int, [C.POINTER(CParams), C.c_int, C.POINTER(vp)]),
    "apsu_b200_ctx_destroy": (None, [vp]),
    "apsu_b200_host_alloc": (C.c_int, [C.c_size_t, C.POINTER(vp)]),
    "apsu_b200_host_free": (None, [vp]),
    "apsu_b200_mgpu_unique_id": (C.c_int, [vp]),
    "apsu_b200_mgpu_create": (C.c_int, [vp, vp, C.c_uint32, C.c_uint32, C.POINTER(vp)]),
    "apsu_b200_mgpu_destroy": (None, [vp]),
    "apsu_b200_mgpu_commit": (C.c_int, [vp, vp, C.c_int]),
    "apsu_b200_mgpu_info": (C.c_int, [vp, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "apsu_b200_mgpu_run_query": (C.c_int, [vp, u32p, C.c_uint32, vp, vp, vp, C.c_uint32, vp, vp, vp]),
    "apsu_b200_mgpu_run_query_shared": (C.c_int, [vp, u32p, C.c_uint32, vp, vp, vp, C.c_uint32, vp, vp, vp]),
    "apsu_b200_mgpu_run_query_local": (C.c_int, [vp, u32p, C.c_uint32, vp, vp, vp, C.c_uint32, vp, vp, vp]),
    "apsu_b200_mgpu_local_count": (C.c_int, [vp, C.POINTER(C.c_uint32)]),
    "apsu_b200_mgpu_compute_powers": (C.c_int, [vp]),
    "apsu_b200_ctx_set_stream": (C.c_int, [vp, vp]),
    "apsu_b200_ctx_get_stream": (C.c_int, [vp, C.POINTER(vp)]),
    "apsu_b200_ctx_synchronize": (C.c_int, [vp]),
    "apsu_b200_ctx_level": (C.c_int, [vp, C.c_int, C.POINTER(C.c_uint32)]),
    "apsu_b200_db_add_binbundle": (C.c_int, [vp, C.c_uint32, C.POINTER(vp), C.c_uint32, C.POINTER(C.c_uint32)]),
    "apsu_b200_db_add_binbundle_synthetic": (C.c_int, [vp, C.c_uint32, C.c_uint32, C.c_uint64, C.POINTER(C.c_uint32)]),
    "apsu_b200_db_add_binbundle_from_bins": (C.c_int, [vp, C.c_uint32, u32p, u64p, C.POINTER(C.c_uint32)]),
    "apsu_b200_db_set_data": (C.c_int, [vp, vp, vp, C.c_uint64, vp]),
    "apsu_b200_db_set_data_device": (C.c_int, [vp, vp, vp, C.c_uint64, vp]),
    "apsu_b200_db_bin_bundle_count": (C.c_int, [vp, C.c_uint32, C.POINTER(C.c_uint32)]),
    "apsu_b200_db_total_bin_bundle_count": (C.c_int, [vp, C.POINTER(C.c_uint32)]),
    "apsu_b200_db_binbundle_ncoeffs": (C.c_int, [vp, C.c_uint32, C.c_uint32, C.POINTER(C.c_uint32)]),
    "apsu_b200_db_binbundle_coeff": (C.c_int, [vp, C.c_uint32, C.c_uint32, C.c_uint32, vp, C.POINTER(C.c_uint32)]),
    "apsu_b200_db_stream_bytes": (C.c_int, [vp, C.POINTER(C.c_uint64)]),
    "apsu_b200_db_clear": (C.c_int, [vp]),
    "apsu_b200_set_relin_keys": (C.c_int, [vp, vp]),
    "apsu_b200_query_begin": (C.c_int, [vp, u32p, C.c_uint32, u64p]),
    "apsu_b200_compute_powers": (C.c_int, [vp]),
    "apsu_b200_get_power": (C.c_int, [vp, C.c_uint32, C.c_uint32, vp, C.POINTER(C.c_uint32), C.POINTER(C.c_int)]),
    "apsu_b200_set_masks": (C.c_int, [vp, u64p, C.c_uint32]),
    "apsu_b200_encode_masks": (C.c_int, [vp, u64p, C.c_uint32, u64p]),
    "apsu_b200_generate_masks": (C.c_int, [vp, vp, vp, C.c_uint32, vp, vp]),
    "apsu_b200_query_begin_seeded": (C.c_int, [vp, u32p, C.c_uint32, vp, vp]),
    "apsu_b200_set_relin_keys_seeded": (C.c_int, [vp, vp, vp]),
    "apsu_b200_op_prng_stream": (C.c_int, [vp, vp, C.c_uint64, vp, C.c_uint64]),
    "apsu_b200_op_expand_seeds": (C.c_int, [vp, C.c_uint32, vp, C.c_uint32, vp]),
    "apsu_b200_decrypt_results": (C.c_int, [vp, vp, vp, C.c_uint32, vp, vp, vp]),
    "apsu_b200_set_powers_partition": (C.c_int, [vp, C.c_uint32, C.c_uint32]),
    "apsu_b200_powers_stage_count": (C.c_int, [vp, C.POINTER(C.c_uint32)]),
    "apsu_b200_compute_powers_stage": (C.c_int, [vp, C.c_uint32]),
    "apsu_b200_powers_exchange_regions": (C.c_int, [vp, C.c_uint32, C.POINTER(C.c_void_p), C.POINTER(C.c_uint64), C.c_uint32, C.POINTER(C.c_uint32)]),
    "apsu_b200_eval_all": (C.c_int, [vp]),
    "apsu_b200_fetch_results": (C.c_int, [vp, vp, vp, vp]),
    "apsu_b200_eval_all_stream": (C.c_int, [vp, vp, vp, vp]),
    "apsu_b200_ctx_set_eval_chunk": (C.c_int, [vp, C.c_uint32]),
    "apsu_b200_run_query": (C.c_int, [vp, u32p, C.c_uint32, vp, vp, vp, C.c_uint32, vp, vp, vp]),
    "apsu_b200_run_query_seeded": (C.c_int, [vp, u32p, C.c_uint32, vp, vp, vp, vp, vp, vp, C.c_uint32, vp, vp, vp, vp]),
    "apsu_b200_query_begin_device": (C.c_int, [vp, u32p, C.c_uint32, vp]),
    "apsu_b200_set_relin_keys_device": (C.c_int, [vp, vp]),
    "apsu_b200_set_masks_device": (C.c_int, [vp, vp, C.c_uint32]),
    "apsu_b200_results_device": (C.c_int, [vp, C.POINTER(vp), C.POINTER(C.c_uint64)]),
    "apsu_b200_copy_results_device": (C.c_int, [vp, vp]),
    "apsu_b200_ctx_modulus_index": (C.c_int, [vp, C.c_int, C.c_uint32, C.POINTER(C.c_uint32)]),
    "apsu_b200_op_ntt": (C.c_int, [vp, u64p, C.c_uint32, u32p, C.c_uint32, C.c_int]),
    "apsu_b200_op_multiply": (C.c_int, [vp, C.c_uint32, u64p, u64p, u64p, C.c_uint32]),
    "apsu_b200_op_relinearize": (C.c_int, [vp, C.c_uint32, u64p, u64p, C.c_uint32]),
    "apsu_b200_op_mod_switch_next": (C.c_int, [vp, C.c_uint32, u64p, u64p, C.c_uint32]),
    "apsu_b200_last_timings": (C.c_int, [vp, C.POINTER(CTimings)]),
    "apsu_b200_bench_ntt": (C.c_int, [vp, C.c_uint32, C.c_uint32, C.c_int, C.POINTER(C.c_float)]),
    "apsu_b200_set_profiling": (C.c_int, [vp, C.c_int]),
}


def lib():
    """Load libapsu_b200.so; raises if it has not been built (python __graft_entry__.py build)."""
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise ImportError(f"{LIB_PATH} is missing: build it with `make -C apsu_b200/csrc` (no CPU fallback exists)")
        L = C.CDLL(str(LIB_PATH))
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(rc: int):
    if rc != OK:
        msg = lib().apsu_b200_last_error().decode()
        raise _EXC.get(rc, RuntimeError)(msg)


def ptr(a):
    if a is None:
        return None
    return a.ctypes.data_as(C.c_void_p)
