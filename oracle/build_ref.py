"""TEST INFRASTRUCTURE ONLY (oracle). Builds the parts of the *reference itself* that
compile in this image, from the sources where they lie under /root/reference, with
outputs only into oracle/_ref/ (git-ignored, shipped to the GPU box by gpurun).

What is built (no reference source is copied into the repo):
  * libclipper_ref.so  - R/pytocr/postprocess/db_postprocess_fast/src/clipper.cpp
                         (vendored Clipper 6.4.2) + oracle/clipper_shim.cpp
  * pse.<abi>.so       - R/pytocr/postprocess/pse_postprocess_fast/pse.pyx, unmodified
                         (same flags as its setup.py:1-19: language c++, -O3)
  * pa.<abi>.so        - R/pytocr/postprocess/pan_postprocess_fast/pa.pyx, unmodified

What is NOT buildable here: db_postprocess.cpp needs OpenCV C++ headers/libs
(db_postprocess.h:3-5, Makefile:5) which this image lacks; the DB oracle is therefore a
restatement over cv2-python (oracle/db_oracle.py) that calls the real Clipper built here.

Run:  python oracle/build_ref.py            (no-op when /root/reference is absent)
"""
import os
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
REF = os.environ.get("OCR_REFERENCE_ROOT", "/root/reference")
PP = os.path.join(REF, "pytocr", "postprocess")


def _run(cmd):
    print("[oracle/build_ref]", " ".join(cmd), flush=True)
    subprocess.check_call(cmd)


def _newer(target, *sources):
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(s) <= t for s in sources)


def build_clipper():
    src = os.path.join(PP, "db_postprocess_fast", "src", "clipper.cpp")
    inc = os.path.join(PP, "db_postprocess_fast", "include")
    shim = os.path.join(HERE, "clipper_shim.cpp")
    out = os.path.join(OUT, "libclipper_ref.so")
    if _newer(out, src, shim):
        return out
    _run(["g++", "-O3", "-std=c++11", "-fPIC", "-shared", "-w", "-I", inc, shim, src, "-o", out])
    return out


def build_pyx(name, subdir):
    import numpy

    pyx = os.path.join(PP, subdir, name + ".pyx")
    cpp = os.path.join(OUT, name + ".cpp")
    so = os.path.join(OUT, name + sysconfig.get_config_var("EXT_SUFFIX"))
    if _newer(so, pyx):
        return so
    _run([sys.executable, "-m", "cython", "--cplus", pyx, "-o", cpp])
    _run(["g++", "-O3", "-fPIC", "-shared", "-w",
          "-I", sysconfig.get_paths()["include"], "-I", numpy.get_include(),
          "-DNPY_NO_DEPRECATED_API=0", cpp, "-o", so])
    os.remove(cpp)  # generated from the reference's .pyx: keep only the binary
    return so


def build_all(verbose=True):
    """Returns the list of built files; empty when the reference tree is not present."""
    if not os.path.isdir(PP):
        if verbose:
            print("[oracle/build_ref] %s not present: using prebuilt oracle/_ref if any" % REF)
        return []
    os.makedirs(OUT, exist_ok=True)
    return [build_clipper(),
            build_pyx("pse", "pse_postprocess_fast"),
            build_pyx("pa", "pan_postprocess_fast")]


if __name__ == "__main__":
    for f in build_all():
        print("built", f)
