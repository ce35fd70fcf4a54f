"""ctypes binding of libgwsem_b200.so (the C ABI in include/gwsem_b200.h).

The library is the product: if it is missing or no B200 is visible the package raises --
there is no CPU path to fall back to.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libgwsem_b200.so")

NA_BITS = 0x7FF00000000007A2


class ModelSpec(C.Structure):
    _fields_ = [("n_vars", C.c_int32), ("n_manifest", C.c_int32), ("n_params", C.c_int32),
                ("n_cells", C.c_int32), ("cells", C.POINTER(C.c_double)),
                ("par_start", C.POINTER(C.c_double)), ("par_lbound", C.POINTER(C.c_double)),
                ("fit", C.c_int32), ("min_variance", C.c_double), ("n_rows", C.c_int32),
                ("manifest_kind", C.POINTER(C.c_int32)),
                ("manifest_data", C.POINTER(C.POINTER(C.c_double))),
                ("n_categories", C.POINTER(C.c_int32)), ("n_exo", C.c_int32),
                ("exo_data", C.POINTER(C.POINTER(C.c_double))), ("exo_latent", C.POINTER(C.c_int32)),
                ("exo_free", C.POINTER(C.c_uint8)), ("n_thresholds", C.c_int32),
                ("thresholds", C.POINTER(C.c_double)), ("exo_kind", C.POINTER(C.c_int32))]


class Result(C.Structure):
    _fields_ = [("est", C.c_void_p), ("vcov", C.c_void_p), ("fit", C.c_void_p), ("maf", C.c_void_p),
                ("n_used", C.c_void_p), ("status", C.c_void_p), ("catch_code", C.c_void_p),
                ("catch_var", C.c_void_p), ("iterations", C.c_void_p)]


class Job(C.Structure):
    _fields_ = [("snp", C.POINTER(C.c_int32)), ("n_snp", C.c_int32), ("start_from", C.c_int32),
                ("out_path", C.c_char_p), ("par_names", C.POINTER(C.c_char_p)),
                ("context_path", C.c_char_p), ("vcov_pairs", C.POINTER(C.c_int32)),
                ("n_vcov_pairs", C.c_int32), ("batch_snps", C.c_int32), ("model_name", C.c_char_p),
                ("manifest_names", C.POINTER(C.c_char_p))]


# every symbol include/gwsem_b200.h declares: (restype, argtypes)
SYMBOLS = {
    "gwsem_abi_version": (C.c_int, []),
    "gwsem_last_error": (C.c_char_p, []),
    "gwsem_device_count": (C.c_int, []),
    "gwsem_engine_create": (C.c_int, [C.c_int, C.POINTER(C.c_void_p)]),
    "gwsem_engine_destroy": (None, [C.c_void_p]),
    "gwsem_engine_set_stream": (C.c_int, [C.c_void_p, C.c_void_p]),
    "gwsem_engine_synchronize": (C.c_int, [C.c_void_p]),
    "gwsem_engine_launch_count": (C.c_int64, [C.c_void_p]),
    "gwsem_engine_series_columns": (C.c_int64, [C.c_void_p]),
    "gwsem_engine_profile": (C.c_int, [C.c_void_p, C.c_int]),
    "gwsem_engine_profile_read": (C.c_int, [C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_double)]),
    "gwsem_engine_fp64_peak": (C.c_int, [C.c_void_p, C.POINTER(C.c_double)]),
    "gwsem_source_open": (C.c_int, [C.c_char_p, C.c_char_p, C.c_int, C.POINTER(C.c_void_p)]),
    "gwsem_source_close": (None, [C.c_void_p]),
    "gwsem_source_num_variants": (C.c_int, [C.c_void_p]),
    "gwsem_source_num_samples": (C.c_int, [C.c_void_p]),
    "gwsem_source_load_counter": (C.c_int, [C.c_void_p]),
    "gwsem_source_checkpoint_columns": (C.c_char_p, [C.c_void_p]),
    "gwsem_source_context": (C.c_int, [C.c_void_p, C.c_int, C.c_char_p, C.c_size_t]),
    "gwsem_source_load_rows": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p,
                                         C.c_void_p, C.c_void_p]),
    "gwsem_decode_2bit_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int,
                                           C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "gwsem_decode_2bit": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int,
                                    C.c_void_p, C.c_void_p, C.c_void_p]),
    "gwsem_model_create": (C.c_int, [C.c_void_p, C.POINTER(ModelSpec), C.POINTER(C.c_void_p)]),
    "gwsem_model_destroy": (None, [C.c_void_p]),
    "gwsem_model_set_row_filter": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int]),
    "gwsem_fit_2bit_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.POINTER(Result)]),
    "gwsem_fit_2bit": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.POINTER(Result)]),
    "gwsem_fit_rows_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.POINTER(Result)]),
    "gwsem_gwas_run": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(Job), C.POINTER(C.c_int64)]),
}

_lib = None


def lib():
    """The loaded library.  Raises (loudly) when the CUDA extension has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(gwsem_b200 has no CPU fallback)")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


class GwsemError(RuntimeError):
    pass


def check(rc):
    if rc != 0:
        raise GwsemError(lib().gwsem_last_error().decode(errors="replace"))
