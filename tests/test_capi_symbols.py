"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol that
include/ah_b200.h declares, and refuses to compute without a CUDA device (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest
import torch

from air_hockey_training_environment_b200 import _capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "ah_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ah_[a-z0-9_]+)\s*\(", src)))


def test_library_is_built_in_tree():
    assert os.path.exists(_capi.LIB_PATH), "run `python -m air_hockey_training_environment_b200.build`"
    assert os.path.dirname(_capi.LIB_PATH).startswith(ROOT)


def test_every_declared_symbol_is_exported_and_bound():
    lib = C.CDLL(_capi.LIB_PATH)
    declared = header_symbols()
    assert len(declared) >= 15
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/ah_b200.h but not exported"
    assert sorted(_capi.SYMBOLS) == declared, "the ctypes binding list must match the header"
    bound = _capi.lib()
    for name in declared:
        assert getattr(bound, name).argtypes is not None, f"{name} has no ctypes signature"


def test_abi_version_and_strerror():
    lib = _capi.lib()
    assert lib.ah_abi_version() == 3
    assert _capi.strerror(_capi.AH_OK) == "ok"
    assert "agent" in _capi.strerror(_capi.AH_EAGENT)
    assert "CUDA" in _capi.strerror(_capi.AH_ECUDA)


def test_struct_layouts_match_header():
    # sizes the C side assumes (LP64): ah_config 32 B, ah_ring 56 B, ah_step_io 160 B
    assert C.sizeof(_capi.AhConfig) == 32
    assert C.sizeof(_capi.AhRing) == 80
    assert C.sizeof(_capi.AhStepIO) == 168


def test_argument_validation_needs_no_gpu():
    lib = _capi.lib()
    h = C.c_void_p()
    assert lib.ah_create(C.byref(h), 0, _capi.AH_F32, 0, 0, 0, None) == _capi.AH_EINVAL
    assert lib.ah_create(C.byref(h), 16, 7, 0, 0, 0, None) == _capi.AH_EINVAL
    assert lib.ah_create(None, 16, _capi.AH_F32, 0, 0, 0, None) == _capi.AH_EINVAL
    assert lib.ah_destroy(None) == _capi.AH_EINVAL
    assert lib.ah_step(None, None, 1, None) == _capi.AH_EINVAL


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback():
    lib = _capi.lib()
    h = C.c_void_p()
    assert lib.ah_create(C.byref(h), 16, _capi.AH_F32, 0, 0, 0, None) == _capi.AH_ECUDA
    from air_hockey_training_environment_b200 import BatchedAirHockey
    with pytest.raises(ValueError):
        BatchedAirHockey(4, device="cpu")
    with pytest.raises(Exception):
        BatchedAirHockey(4, device="cuda")


def test_product_package_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "air_hockey_training_environment_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text.replace("no oracle", ""), f"{f} mentions the oracle"


def test_header_is_plain_c_and_struct_sizes_agree(tmp_path):
    """include/ah_b200.h must compile as C99 (no C++/torch types on the boundary) and its struct sizes must equal
    the ctypes mirrors'."""
    import subprocess
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include "ah_b200.h"\n'
                   'int main(void){ printf("%zu %zu %zu %d %d\\n", sizeof(ah_config), sizeof(ah_ring), sizeof(ah_step_io), '
                   'AH_NSTATS, AH_ACTOR_NWEIGHTS); return 0; }\n')
    exe = tmp_path / "sz"
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"),
                    str(src), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()
    assert [int(v) for v in out] == [C.sizeof(_capi.AhConfig), C.sizeof(_capi.AhRing), C.sizeof(_capi.AhStepIO),
                                     _capi.AH_NSTATS, _capi.AH_ACTOR_NWEIGHTS]
