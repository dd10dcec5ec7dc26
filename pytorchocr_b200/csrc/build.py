"""Builds pytorchocr_b200/csrc/libocrpp.so for sm_100a with nvcc (cross-compiles without a GPU).

    python -m pytorchocr_b200.csrc.build [--force] [--verbose]

The .so is built IN-TREE (git-ignored, shipped to the GPU box by gpurun)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
OUT = os.path.join(HERE, "libocrpp.so")
SOURCES = ["api.cu", "ctc.cu", "db.cu", "expand.cu", "crop.cu", "prep.cu"]
EXTRA = os.environ.get("OCRPP_NVCC_EXTRA", "").split()
NVCC_FLAGS = EXTRA + ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "--shared", "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden",
              "--expt-relaxed-constexpr", "-Xptxas", "-warn-spills"]


def sources():
    return [os.path.join(HERE, s) for s in SOURCES if os.path.exists(os.path.join(HERE, s))]


def _deps():
    d = sources()
    d += [os.path.join(HERE, f) for f in os.listdir(HERE) if f.endswith((".cuh", ".h"))]
    d.append(os.path.join(ROOT, "include", "ocrpp.h"))
    return d


def build(force=False, verbose=False):
    if not force and os.path.exists(OUT):
        t = os.path.getmtime(OUT)
        if all(os.path.getmtime(p) <= t for p in _deps()):
            return OUT
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + sources() + ["-o", OUT]
    print("[build]", " ".join(cmd), flush=True)
    subprocess.check_call(cmd)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
