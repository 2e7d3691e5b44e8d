"""CPU-side checks of what the compiler made of the hot kernels (no GPU needed: nvcc cross-compiles, cuobjdump reads the
in-tree library): the facts DESIGN.md and profiles/ quote about ``ah::step_packed_kernel<float>`` -- the kernel bench.py
times -- must hold for the library that ships."""
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "air_hockey_training_environment_b200")
LIB = os.path.join(PKG, "libah_b200.so")
HEADLINE = "_ZN2ah18step_packed_kernelIfLi0E"          # step_packed_kernel<float, IO_PACKED>

pytestmark = pytest.mark.skipif(shutil.which("cuobjdump") is None or not os.path.exists(LIB),
                                reason="needs cuobjdump and the built library")


@pytest.fixture(scope="module")
def sass():
    txt = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    out = {}
    for f in re.split(r"\n\s+Function : ", txt)[1:]:
        name = f.split("\n", 1)[0].strip()
        out[name] = [m.group(1).strip() for m in re.finditer(r"/\*[0-9a-f]{4,5}\*/\s+(.*?);", f)]
    return out


def kernel(sass, prefix):
    names = [k for k in sass if k.startswith(prefix)]
    assert len(names) == 1, names
    return sass[names[0]]


def ops(ins):
    return [(t.split()[1] if t.startswith("@") else t.split()[0]) for t in ins]


def test_headline_kernel_fits_nine_blocks_per_sm_without_spills():
    info = open(os.path.join(PKG, "csrc", "ptxas_info.txt")).read()
    m = re.search(re.escape(HEADLINE) + r".*?\n.*?(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads\n.*?Used (\d+) registers",
                  info, re.S)
    assert m, "ptxas info of the headline kernel not found (python -m air_hockey_training_environment_b200.build)"
    stack, st, ld, regs = map(int, m.groups())
    assert (stack, st, ld) == (0, 0, 0)
    assert regs <= 56                                   # 9 blocks of 128 threads per SM: 65,536 / (9 * 128) = 56


def test_headline_kernel_memory_and_arithmetic_instructions(sass):
    o = ops(kernel(sass, HEADLINE))
    # state rows, meta and the two halves of obs_next move as 128-bit accesses; the action / event words as 32-bit ones
    assert sum(x.startswith("LDG.E.128") for x in o) >= 4 and sum(x.startswith("STG.E.128") for x in o) >= 6
    assert not any(x.startswith(("LDL", "STL")) for x in o)                  # no local memory
    # Blackwell packed FP32 (add.f32x2 / fma.rn.f32x2 / mul.f32x2) carries the (x, y) pairs
    assert sum(x.split(".")[0] in ("FADD2", "FFMA2", "FMUL2") for x in o) >= 60
    # FP32 build: no double-precision arithmetic, and the statistics go through IDP.4A / REDUX
    assert not any(x.split(".")[0] in ("DADD", "DMUL", "DFMA") for x in o)
    assert any(x.startswith("IDP") for x in o) and any(x.startswith("REDUX") for x in o)
