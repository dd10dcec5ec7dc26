"""ctypes loader for libocrpp.so (the C-ABI declared in include/ocrpp.h).

There is no CPU fallback: if the shared library cannot be loaded (and cannot be built because
nvcc is missing) importing the operators fails loudly."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("OCRPP_LIB_PATH") or os.path.join(_HERE, "csrc", "libocrpp.so")   # the override is a development aid

OK = 0
F32, F16 = 0, 1
IMG_RUN_OVERFLOW, IMG_CANDIDATES_TRUNCATED, IMG_VALUE_OUT_OF_RANGE = 1, 2, 4
TUNE_DB_PATH, TUNE_DB_SPLIT, TUNE_DB_PRIO, TUNE_DB_SCAN, TUNE_DB_SCAN4_STAGES, TUNE_DB_SCAN4_CTAS = 0, 1, 2, 3, 4, 5
DB_SEMANTICS_CPP, DB_SEMANTICS_PYTHON, DB_SCORE_POLY, DB_SCORE_BOX = 0, 1, 0, 1

_lib = None

_i32p = C.POINTER(C.c_int32)

# name -> (restype, argtypes); must list every symbol include/ocrpp.h declares
# (tests/test_abi.py cross-checks this table against the header and the built library)
SIGNATURES = {
    "ocrpp_abi_version": (C.c_int, []),
    "ocrpp_last_error": (C.c_char_p, []),
    "ocrpp_launch_count": (C.c_int64, []),
    "ocrpp_reset_launch_count": (None, []),
    "ocrpp_set_tuning": (C.c_int, [C.c_int, C.c_int]),
    "ocrpp_profile_enable": (None, [C.c_int]),
    "ocrpp_profile_reset": (None, []),
    "ocrpp_profile_read": (C.c_int, [C.POINTER(C.c_float), C.c_int, C.POINTER(C.c_int)]),
    "ocrpp_profile_phase_name": (C.c_char_p, [C.c_int]),
    "ocrpp_ctc_greedy": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int64, C.c_int64,
                                   C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "ocrpp_db_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int, C.c_int]),
    "ocrpp_db_postprocess": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int64, C.c_int64,
                                       C.c_void_p, C.c_float, C.c_float, C.c_float, C.c_int, C.c_int, C.c_int, C.c_int,
                                       C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                       C.c_void_p, C.c_size_t, C.c_void_p]),
    "ocrpp_db_postprocess_ex": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int64, C.c_int64,
                                          C.c_void_p, C.c_float, C.c_double, C.c_double, C.c_int, C.c_int, C.c_int, C.c_int,
                                          C.c_int, C.c_int,
                                          C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                          C.c_void_p, C.c_size_t, C.c_void_p]),
    "ocrpp_pse_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int64]),
    "ocrpp_pse_postprocess": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                        C.c_int64, C.c_int64, C.c_int64, C.c_int, C.c_int, C.c_void_p,
                                        C.c_float, C.c_float, C.c_float, C.c_float, C.c_int, C.c_int, C.c_int64,
                                        C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                        C.c_void_p, C.c_size_t, C.c_void_p]),
    "ocrpp_pan_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int64]),
    "ocrpp_pan_postprocess": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                        C.c_int64, C.c_int64, C.c_int64, C.c_int, C.c_int, C.c_void_p,
                                        C.c_float, C.c_float, C.c_float, C.c_float, C.c_int, C.c_int, C.c_int64,
                                        C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                        C.c_void_p, C.c_size_t, C.c_void_p]),
    "ocrpp_crop_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int]),
    "ocrpp_crop_boxes": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int64, C.c_int64,
                                   C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_size_t,
                                   C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "ocrpp_rec_preprocess": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                       C.c_void_p, C.c_void_p]),
}
IMG_MODE = {"GRAY": 0, "RGB": 1, "BGR": 2}


class OcrppError(RuntimeError):
    pass


def lib():
    """Loads (building first if the binary is absent and nvcc exists) and returns the CDLL."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        from .csrc.build import build
        try:
            build()
        except Exception as e:  # no nvcc, compile error, ...
            raise OcrppError(
                "libocrpp.so is not built and could not be built (%s). Run "
                "`python -m pytorchocr_b200.csrc.build`; there is no CPU fallback." % (e,))
    L = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        if not hasattr(L, name):
            continue  # entry point not built yet; check() reports it when it is called
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    if L.ocrpp_abi_version() != 2:
        raise OcrppError("libocrpp.so ABI version mismatch")
    _lib = L
    return L


def check(status):
    if status != OK:
        msg = lib().ocrpp_last_error()
        raise OcrppError("libocrpp call failed (status %d): %s" % (status, (msg or b"").decode("utf-8", "replace")))


def require_cuda():
    import torch
    if not torch.cuda.is_available():
        raise OcrppError("pytorchocr_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    return torch


def profile_read():
    """-> (calls, [(phase_name, total_ms), ...]) since the last ocrpp_profile_reset()."""
    L = lib()
    ms = (C.c_float * 16)()
    calls = C.c_int(0)
    n = L.ocrpp_profile_read(ms, 16, C.byref(calls))
    return calls.value, [(L.ocrpp_profile_phase_name(i).decode(), float(ms[i])) for i in range(n)]
