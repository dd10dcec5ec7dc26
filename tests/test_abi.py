"""CPU checks of the drop-in boundary: the C-ABI library builds for sm_100a, loads, and exports
every symbol include/ocrpp.h declares; the ctypes table covers the header; the product package
does not import the oracle."""
import ctypes
import os
import re
import subprocess

import pytest

from conftest import ROOT


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "ocrpp.h")).read()
    return sorted(set(re.findall(r"OCRPP_API\s+[\w\s\*]+?\b(ocrpp_\w+)\s*\(", src)))


def test_header_declares_entry_points():
    syms = _header_symbols()
    for s in ("ocrpp_ctc_greedy", "ocrpp_db_postprocess", "ocrpp_last_error", "ocrpp_abi_version"):
        assert s in syms


def test_library_builds_and_exports_every_declared_symbol():
    from pytorchocr_b200.csrc.build import build
    path = build()
    L = ctypes.CDLL(path)
    missing = [s for s in _header_symbols() if not hasattr(L, s)]
    assert not missing, missing
    L.ocrpp_abi_version.restype = ctypes.c_int
    assert L.ocrpp_abi_version() == 2
    # sm_100a code is really in the binary
    out = subprocess.check_output(["/usr/local/cuda/bin/cuobjdump", "-lelf", path]).decode()
    assert "sm_100a" in out


def test_ctypes_table_matches_header():
    from pytorchocr_b200 import _lib
    assert sorted(_lib.SIGNATURES) == _header_symbols()


def test_invalid_arguments_fail_without_gpu():
    """Argument validation happens before any CUDA call, so it is checkable on CPU."""
    from pytorchocr_b200 import _lib
    L = _lib.lib()
    st = L.ocrpp_ctc_greedy(None, 7, 1, 1, 1, 1, 1, None, None, None, None, None, None)
    assert st == 1
    assert b"dtype" in L.ocrpp_last_error()


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "pytorchocr_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, re.M), os.path.join(dp, f)
                # no include / dlopen / path construction that reaches into oracle/ (comments may cite it)
                assert not re.search(r'#include\s+"[^"]*oracle', txt), os.path.join(dp, f)
                assert not re.search(r'(CDLL|join|open)\([^)]*oracle', txt), os.path.join(dp, f)


def test_flag_off_fails_loudly():
    from pytorchocr_b200 import _lib
    from pytorchocr_b200.postprocess import build_post_process
    with pytest.raises(_lib.OcrppError):
        build_post_process({"name": "CTCLabelDecode"})
    assert build_post_process({"name": "None"}) is None
