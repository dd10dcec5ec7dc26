"""TEST INFRASTRUCTURE ONLY (oracle). ctypes bindings to oracle/_build/liboracle.so."""
import ctypes as C

import numpy as np

from .build_oracle import build

_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(build())
        i32p, u8p, f32p = C.POINTER(C.c_int32), C.POINTER(C.c_uint8), C.POINTER(C.c_float)
        L.oracle_ccl4.argtypes = [u8p, C.c_int, C.c_int, i32p]
        L.oracle_ccl4.restype = C.c_int
        L.oracle_pse.argtypes = [u8p, C.c_int, C.c_int, C.c_int, C.c_float, i32p]
        L.oracle_pse.restype = None
        L.oracle_pa_expand.argtypes = [u8p, f32p, i32p, C.c_int, C.c_int, i32p, f32p, i32p]
        L.oracle_pa_expand.restype = None
        L.oracle_ctc_greedy.argtypes = [f32p, C.c_int, C.c_int, C.c_int, C.c_int64, C.c_int64,
                                        i32p, f32p, i32p, i32p]
        L.oracle_ctc_greedy.restype = None
        _lib = L
    return _lib


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def ccl4(img):
    """uint8 [H,W] -> (label_num, int32 labels) == cv2.connectedComponents(img, connectivity=4)."""
    img = np.ascontiguousarray(img, dtype=np.uint8)
    H, W = img.shape
    out = np.empty((H, W), np.int32)
    n = lib().oracle_ccl4(_p(img, C.c_uint8), H, W, _p(out, C.c_int32))
    return n, out


def pse(kernels, min_area):
    kernels = np.ascontiguousarray(kernels, dtype=np.uint8)
    K, H, W = kernels.shape
    out = np.empty((H, W), np.int32)
    lib().oracle_pse(_p(kernels, C.c_uint8), K, H, W, float(min_area), _p(out, C.c_int32))
    return out


def pa_expand(text, emb, label, flag, mean_emb):
    text = np.ascontiguousarray(text, dtype=np.uint8)
    emb = np.ascontiguousarray(emb, dtype=np.float32)
    label = np.ascontiguousarray(label, dtype=np.int32)
    flag = np.ascontiguousarray(flag, dtype=np.int32)
    mean_emb = np.ascontiguousarray(mean_emb, dtype=np.float32)
    H, W = text.shape
    out = np.empty((H, W), np.int32)
    lib().oracle_pa_expand(_p(text, C.c_uint8), _p(emb, C.c_float), _p(label, C.c_int32), H, W,
                           _p(flag, C.c_int32), _p(mean_emb, C.c_float), _p(out, C.c_int32))
    return out


def ctc_greedy(probs_tbc):
    """float32 [T,B,C] (any strides with unit class stride) -> idx[B,T], prob[B,T], len[B], raw[B,T]."""
    a = probs_tbc
    assert a.dtype == np.float32 and a.ndim == 3 and a.strides[2] == 4
    T, B, Cc = a.shape
    idx = np.zeros((B, T), np.int32)
    prob = np.zeros((B, T), np.float32)
    ln = np.zeros((B,), np.int32)
    raw = np.zeros((B, T), np.int32)
    lib().oracle_ctc_greedy(_p(a, C.c_float), T, B, Cc, a.strides[0] // 4, a.strides[1] // 4,
                            _p(idx, C.c_int32), _p(prob, C.c_float), _p(ln, C.c_int32),
                            _p(raw, C.c_int32))
    return idx, prob, ln, raw
