"""CPU: the C-ABI library builds, loads and exports every declared symbol."""
import ctypes
import os
import re

from optyx_b200 import build as b2build
from optyx_b200 import runtime as rt

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_builds_and_exports_header_symbols():
    path = b2build.build()
    lib = ctypes.CDLL(path)
    header = open(os.path.join(ROOT, "include", "optyx_b200.h")).read()
    declared = set(re.findall(r"\b(b2_[a-z0-9_]+)\s*\(", header))
    assert declared, "no symbols parsed from the header"
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in optyx_b200.h but not exported"
    assert declared == set(rt.EXPORTS)
    assert lib.b2_abi_version() == 1


def test_struct_sizes_match_header():
    assert ctypes.sizeof(rt.LeafDesc) == 72
    assert ctypes.sizeof(rt.Modes) == 8 + 64 * 4 + 3 * 64 * 8
    assert ctypes.sizeof(rt.GemmModes) == 16 + 64 * 4 + 3 * 64 * 8
    assert rt.SP_STEP_DTYPE.itemsize == rt.SP_COPY_DTYPE.itemsize == rt.SP_TASK_DTYPE.itemsize == 32
    assert ctypes.sizeof(rt.SmallProgramArgs) == 6 * 8 + 3 * 8 + 8 + 4 * 4


def test_no_cpu_fallback_without_device():
    import pytest
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    lib = rt.load()
    # a compute entry point must fail loudly, not fall back
    desc = rt.Modes()
    desc.n_out = 0
    desc.n_red = 0
    buf = (ctypes.c_double * 2)()
    st = lib.b2_contract_direct(ctypes.byref(desc), buf, buf, buf, 1.0, 0.0, 0, 0, None)
    assert st != 0
    assert b"CUDA" in lib.b2_last_error() or b"device" in lib.b2_last_error()
    from optyx_b200.executor import require_cuda
    with pytest.raises(rt.B2Error):
        require_cuda()
