"""TEST INFRASTRUCTURE ONLY.  Build recipe for the C oracle (oracle/ah_oracle.c -> oracle/libah_oracle.so).

The Python reference has no compiled sources, so there is no ``oracle/_ref`` to build for this
project; the C restatement is validated against the live Python reference instead.
"""
from __future__ import annotations

import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "ah_oracle.c")
LIB = os.path.join(HERE, "libah_oracle.so")

# -ffp-contract=off : no compiler-introduced FMA; the only fma() calls are the explicit numpy-dot ones
# -fno-builtin-pow  : keep `** 2` as libm pow(), as CPython / numpy scalars evaluate it
# -march=x86-64-v3  : hardware FMA for fma() without tying the .so to the build machine's exact CPU
CFLAGS = ["-O2", "-std=c11", "-fPIC", "-shared", "-ffp-contract=off", "-fno-fast-math", "-fno-builtin-pow",
          "-march=x86-64-v3", "-fopenmp", "-Wall", "-Wextra", "-Wno-unused-function"]


def build(force: bool = False) -> str:
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= os.path.getmtime(SRC):
        return LIB
    cmd = ["gcc", *CFLAGS, "-o", LIB, SRC, "-lm"]
    subprocess.run(cmd, check=True, cwd=HERE)
    return LIB


if __name__ == "__main__":
    print(build(force=True))
