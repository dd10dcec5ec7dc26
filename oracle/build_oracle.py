"""TEST INFRASTRUCTURE ONLY (oracle). Compiles oracle/c/ocr_oracle.c -> oracle/_build/liboracle.so."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "c", "ocr_oracle.c")
OUT = os.path.join(HERE, "_build", "liboracle.so")


def build(force=False):
    if (not force and os.path.exists(OUT) and os.path.exists(SRC)
            and os.path.getmtime(OUT) >= os.path.getmtime(SRC)):
        return OUT
    if not os.path.exists(SRC):
        if os.path.exists(OUT):
            return OUT
        raise FileNotFoundError(SRC)
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    subprocess.check_call(["gcc", "-O3", "-std=c99", "-fPIC", "-shared", "-fno-fast-math",
                           SRC, "-lm", "-o", OUT])
    return OUT


if __name__ == "__main__":
    print(build(force=True))
