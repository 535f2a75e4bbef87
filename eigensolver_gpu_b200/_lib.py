"""ctypes loader for libeigb200.so -- the C-ABI drop-in library (include/eigb200.h).

There is no CPU fallback: if the library is missing or does not load, importing the compute API raises."""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("EIGB200_LIB_PATH") or os.path.join(HERE, "lib", "libeigb200.so")      # (override: debugging builds)

_lib = None

_i, _d, _p, _c = C.c_int, C.c_double, C.c_void_p, C.c_char
_ip = C.POINTER(C.c_int)

# name -> (restype, argtypes); mirrors include/eigb200.h one to one
SIGNATURES = {
    "eigb200_init": (_i, []),
    "eigb200_finalize": (_i, []),
    "eigb200_last_error": (C.c_char_p, []),
    "eigb200_set_stream": (_i, [_p]),
    "eigb200_set_a_ready_event": (_i, [_p]),
    "eigb200_version": (_i, []),
    "eigb200_scratch_bytes": (C.c_int64, [_i, _i]),
    "eigb200_mg_alloc": (_i, [C.c_longlong, C.POINTER(C.c_void_p), C.c_char_p]),
    "eigb200_mg_open": (_i, [C.c_char_p, C.POINTER(C.c_void_p)]),
    "eigb200_mg_config": (_i, [_i, _i, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.c_longlong, _p]),
    "eigb200_mg_flag_bytes": (_i, [_i, _i]),
    "eigb200_mg_unique_id": (_i, [C.c_char_p]),
    "eigb200_mg_init": (_i, [_i, _i, C.c_char_p]),
    "eigb200_mg_finalize": (_i, []),
    "eigb200_mg_allgather_columns": (_i, [_p, _i, _i, _i]),
    "eigb200_mg_column_range": (_i, [_i, _i, _i, _ip, _ip]),
    "eigb200_prof_enable": (_i, [_i]),
    "eigb200_prof_reset": (_i, []),
    "eigb200_prof_collect": (_i, [_p, _p, _p]),
    "eigb200_trace_read": (C.c_longlong, [_p, C.c_longlong]),
    "eigb200_probe_peaks": (_i, [_p]),
    "eigb200_set_option": (_i, [C.c_char_p, _i]),
    "eigb200_get_option": (_i, [C.c_char_p]),
    "eigb200_dsygvdx": (_i, [_i, _p, _i, _p, _i, _p, _i, _i, _i, _p, _p, _i, _p, _i, _p, _i, _p, _i, _p, _ip, _i]),
    "eigb200_zhegvdx": (_i, [_i, _p, _i, _p, _i, _p, _i, _i, _i, _p, _p, _i, _p, _i, _p, _i, _p, _i, _p, _i, _p, _i,
                             _p, _ip, _i]),
    "eigb200_dsygvdx_mg": (_i, [_i, _p, _i, _p, _i, _p, _i, _i, _i, _p, _p, _i, _p, _i, _p, _i, _p, _i, _p, _ip, _i]),
    "eigb200_zhegvdx_mg": (_i, [_i, _p, _i, _p, _i, _p, _i, _i, _i, _p, _p, _i, _p, _i, _p, _i, _p, _i, _p, _i, _p, _i,
                                _p, _ip, _i]),
    "eigb200_dsyevd": (_i, [_i, _i, _i, _p, _i, _p, _i, _p, _p, _i, _p, _i, _p, _i, _p, _i, _p, _ip]),
    "eigb200_zheevd": (_i, [_i, _i, _i, _p, _i, _p, _i, _p, _p, _i, _p, _i, _p, _i, _p, _i, _p, _i, _p, _i, _p, _ip]),
    "eigb200_dpotrf": (_i, [_i, _p, _i, _ip]),
    "eigb200_zpotrf": (_i, [_i, _p, _i, _ip]),
    "eigb200_dsygst": (_i, [_i, _p, _i, _p, _i]),
    "eigb200_zhegst": (_i, [_i, _p, _i, _p, _i]),
    "eigb200_dsytrd": (_i, [_i, _p, _i, _p, _p, _p]),
    "eigb200_zhetrd": (_i, [_i, _p, _i, _p, _p, _p]),
    "eigb200_dsymv": (_i, [_i, _p, _i, _p, _p]),
    "eigb200_zhemv": (_i, [_i, _p, _i, _p, _p]),
    "eigb200_dsyr2k": (_i, [_i, _i, _d, _p, _i, _p, _i, _d, _p, _i]),
    "eigb200_zher2k": (_i, [_i, _i, _d, _p, _i, _p, _i, _d, _p, _i]),
    "eigb200_dgemm": (_i, [_c, _c, _i, _i, _i, _d, _p, _i, _p, _i, _d, _p, _i]),
    "eigb200_zgemm": (_i, [_c, _c, _i, _i, _i, _d, _p, _i, _p, _i, _d, _p, _i]),
    "eigb200_dstedc": (_i, [_i, _p, _p, _p, _i]),
    "eigb200_dstedc_range": (_i, [_i, _p, _p, _p, _i, _i, _i]),
    "eigb200_dormtr": (_i, [_i, _i, _p, _i, _p, _p, _i]),
    "eigb200_zunmtr": (_i, [_i, _i, _p, _i, _p, _p, _i]),
    "eigb200_dtrsm": (_i, [_c, _c, _i, _i, _p, _i, _p, _i]),
    "eigb200_ztrsm": (_i, [_c, _c, _i, _i, _p, _i, _p, _i]),
}


def load(strict=False):
    """Loads the library and sets the prototypes. strict=False tolerates symbols not built yet (dev only)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"eigb200: {LIB_PATH} is missing -- build it with `python -m eigensolver_gpu_b200.build` "
            "(there is no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        try:
            fn = getattr(lib, name)
        except AttributeError:
            if strict:
                raise RuntimeError(f"eigb200: symbol {name} missing from {LIB_PATH}")
            continue
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def sync_stream():
    """Points the library at torch's CURRENT stream; every Python entry point calls this first, so that a call made
    inside `with torch.cuda.stream(s)` and the next one made outside of it are each issued where the caller is."""
    import torch
    lib = load()
    lib.eigb200_set_stream(C.c_void_p(torch.cuda.current_stream().cuda_stream))
    return lib


def last_error():
    return load().eigb200_last_error().decode()


class Eigb200Error(RuntimeError):
    pass


def check(rc, what):
    if rc != 0:
        raise Eigb200Error(f"{what} failed (rc={rc}): {last_error()}")
