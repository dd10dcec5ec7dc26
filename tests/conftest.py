import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

REF_ROOT = "/root/reference"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "needs_reference: needs /root/reference (authoring container only)")


def _cuda_ok():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    have_gpu = _cuda_ok()
    have_ref = os.path.isdir(os.path.join(REF_ROOT, "pytocr"))
    for it in items:
        if "gpu" in it.keywords and not have_gpu:
            it.add_marker(pytest.mark.skip(reason="no CUDA device"))
        if "needs_reference" in it.keywords and not have_ref:
            it.add_marker(pytest.mark.skip(reason="/root/reference not present"))


@pytest.fixture(scope="session")
def ref_modules():
    """The reference's own compiled Cython modules (oracle/_ref, built by oracle/build_ref.py)."""
    p = os.path.join(ROOT, "oracle", "_ref")
    if not any(f.startswith("pse.") and f.endswith(".so") for f in os.listdir(p) if os.path.isdir(p)):
        pytest.skip("oracle/_ref not built")
    if p not in sys.path:
        sys.path.insert(0, p)
    import pa
    import pse
    return pse, pa
