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
    assert ctypes.sizeof(rt.SmallProgramArgs) == 4 * 8 + 3 * 8 + 8 + 4 * 4
    assert ctypes.sizeof(rt.PlanStep) == 8 * 4 + 4 * 8 + 2 * 16 * 4 + 2 * 16 * 8 + ctypes.sizeof(rt.GemmModes)


def test_ctypes_layouts_equal_what_a_c_compiler_sees(tmp_path):
    """sizeof / offsetof of every struct of include/optyx_b200.h as gcc lays them out
    == the ctypes mirrors in optyx_b200/runtime.py (a drifted field is a silent
    memory corruption on the device)."""
    import subprocess
    src = tmp_path / "sizes.c"
    src.write_text(
        '#include <stdio.h>\n#include <stddef.h>\n#include "optyx_b200.h"\n'
        'int main(void) { printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu\\n", sizeof(b2_leaf_desc), '
        'sizeof(b2_modes), sizeof(b2_gemm_modes), sizeof(b2_small_program), sizeof(b2_plan_step), '
        'offsetof(b2_plan_step, desc), offsetof(b2_plan_step, a_slice_stride), '
        'offsetof(b2_small_program, sync_dev), offsetof(b2_small_program, smem_bytes)); return 0; }\n')
    exe = tmp_path / "sizes"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    got = [int(x) for x in subprocess.run([str(exe)], capture_output=True, text=True,
                                          check=True).stdout.split()]
    want = [ctypes.sizeof(rt.LeafDesc), ctypes.sizeof(rt.Modes), ctypes.sizeof(rt.GemmModes),
            ctypes.sizeof(rt.SmallProgramArgs), ctypes.sizeof(rt.PlanStep),
            rt.PlanStep.desc.offset, rt.PlanStep.a_slice_stride.offset,
            rt.SmallProgramArgs.sync_dev.offset, rt.SmallProgramArgs.smem_bytes.offset]
    assert got == want


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
