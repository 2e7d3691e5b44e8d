"""Build recipe of the CUDA extension (csrc/ah_b200.cu -> libah_b200.so, in-tree, sm_100a only).

    python -m air_hockey_training_environment_b200.build

The library is a plain C-ABI shared object (include/ah_b200.h), loaded with ctypes by ``_capi.py``; it
links the CUDA runtime statically and has no dependency on torch.
"""
from __future__ import annotations

import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "csrc", "ah_b200.cu")
DEPS = [SRC, *(os.path.join(HERE, "csrc", h) for h in ("ah_physics.cuh", "ah_rng.cuh", "ah_step_packed.cuh", "ah_policy.cuh")),
        os.path.join(os.path.dirname(HERE), "include", "ah_b200.h")]
LIB = os.path.join(HERE, "libah_b200.so")

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
# -fmad=false: the FP64 validation build must not contract a*b+c (the only FMAs are the explicit ones that
#              reproduce numpy's dot); the FP32 build writes its fmaf() calls out explicitly.
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo", "-fmad=false",
         "-shared", "-Xcompiler", "-fPIC", "-Xptxas", "-v", "-cudart", "static"]


def stale() -> bool:
    return not os.path.exists(LIB) or any(os.path.getmtime(d) > os.path.getmtime(LIB) for d in DEPS if os.path.exists(d))


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not stale():
        return LIB
    cmd = [NVCC, *FLAGS, "-o", LIB, SRC]
    res = subprocess.run(cmd, cwd=HERE, capture_output=True, text=True)
    if verbose or res.returncode != 0:
        print(res.stdout)
        print(res.stderr)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed: " + " ".join(cmd))
    with open(os.path.join(HERE, "csrc", "ptxas_info.txt"), "w") as f:
        f.write(res.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force=True, verbose=True))
