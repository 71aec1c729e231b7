"""ctypes binding of ``liboptyx_b200.so`` (the C ABI in include/optyx_b200.h).

There is no CPU fallback: if the library is missing, or a compute entry point
is called without a CUDA device, this module raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import build as _build

B2_C128, B2_C64 = 0, 1
MAX_LEAF_RANK = 8
MAX_MODES = 64
MAX_RED_DIRECT = 16

EXPORTS = [
    "b2_abi_version", "b2_last_error", "b2_last_kernel", "b2_device_sms", "b2_build_leaves",
    "b2_contract_direct", "b2_contract_gemm", "b2_run_small_program", "b2_axpy", "b2_fill_zero",
    "b2_probs_amp", "b2_probs_dm", "b2_compact_nonzero", "b2_permanent_amplitudes", "b2_permanent_histogram",
    "b2_fp64_peak_kernel", "b2_run_plan", "b2_workspace_bytes", "b2_build_leaves_indexed",
    "b2_plan_greedy", "b2_plan_cost", "b2_plan_reconfigure", "b2_conj", "b2_bond_compress", "b2_bond_compress_scratch_bytes",
]


class LeafDesc(C.Structure):
    _fields_ = [
        ("kind", C.c_int32), ("ndim", C.c_int32),
        ("dims", C.c_int32 * MAX_LEAF_RANK),
        ("ipar", C.c_int32), ("shared", C.c_int32),
        ("offset", C.c_int64), ("nelem", C.c_int64), ("poff", C.c_int64),
    ]


LEAF_DESC_DTYPE = np.dtype([
    ("kind", "<i4"), ("ndim", "<i4"), ("dims", "<i4", (MAX_LEAF_RANK,)),
    ("ipar", "<i4"), ("shared", "<i4"),
    ("offset", "<i8"), ("nelem", "<i8"), ("poff", "<i8"),
])
assert LEAF_DESC_DTYPE.itemsize == C.sizeof(LeafDesc) == 72


class Modes(C.Structure):
    _fields_ = [
        ("n_out", C.c_int32), ("n_red", C.c_int32),
        ("dim", C.c_int32 * MAX_MODES),
        ("sa", C.c_int64 * MAX_MODES),
        ("sb", C.c_int64 * MAX_MODES),
        ("sc", C.c_int64 * MAX_MODES),
    ]


class GemmModes(C.Structure):
    _fields_ = [
        ("n_batch", C.c_int32), ("n_m", C.c_int32),
        ("n_n", C.c_int32), ("n_k", C.c_int32),
        ("dim", C.c_int32 * MAX_MODES),
        ("sa", C.c_int64 * MAX_MODES),
        ("sb", C.c_int64 * MAX_MODES),
        ("sc", C.c_int64 * MAX_MODES),
    ]


class SmallProgramArgs(C.Structure):
    """b2_small_program (include/optyx_b200.h)."""
    _fields_ = [
        ("blob_dev", C.c_void_p), ("blob_off_dev", C.c_void_p),
        ("out_tab_dev", C.c_void_p), ("red_tab_dev", C.c_void_p),
        ("base_dev", C.c_void_p * 3), ("sync_dev", C.c_void_p),
        ("n_tasks", C.c_int32), ("batch", C.c_int32), ("smem_bytes", C.c_int32),
        ("reserved", C.c_int32),
    ]


MAX_SLICED = 16


class _StepDesc(C.Union):
    _fields_ = [("direct", Modes), ("gemm", GemmModes)]


class PlanStep(C.Structure):
    """b2_plan_step (include/optyx_b200.h): one row of the step table b2_run_plan walks."""
    _fields_ = [
        ("kind", C.c_int32), ("per_slice", C.c_int32), ("final", C.c_int32),
        ("a_space", C.c_int32), ("b_space", C.c_int32), ("c_space", C.c_int32),
        ("n_a_slice", C.c_int32), ("n_b_slice", C.c_int32),
        ("a_off", C.c_int64), ("b_off", C.c_int64), ("c_off", C.c_int64), ("c_elems", C.c_int64),
        ("a_slice_pos", C.c_int32 * MAX_SLICED), ("b_slice_pos", C.c_int32 * MAX_SLICED),
        ("a_slice_stride", C.c_int64 * MAX_SLICED), ("b_slice_stride", C.c_int64 * MAX_SLICED),
        ("desc", _StepDesc),
    ]


STEP_DIRECT, STEP_GEMM = 0, 1
RUN_ZERO, RUN_INVARIANT, RUN_SLICES = 1, 2, 4


class B2Error(RuntimeError):
    pass


_lib = None


def lib_path() -> str:
    return _build.LIB


def load():
    """Load (building first if needed) the CUDA library.  Raises if it cannot
    be built or loaded -- the product path never degrades to a CPU path."""
    global _lib
    if _lib is not None:
        return _lib
    # build() returns at once when the library's stamp matches the sources; a stale
    # library (sources edited since the last build: struct layouts may have drifted)
    # is rebuilt, never loaded
    try:
        path = _build.build()
    except (RuntimeError, OSError) as exc:
        raise B2Error(f"liboptyx_b200.so is missing or stale and cannot be rebuilt: {exc}") from exc
    lib = C.CDLL(path)
    vp, i32, i64, dbl = C.c_void_p, C.c_int32, C.c_int64, C.c_double
    lib.b2_abi_version.restype = C.c_int
    lib.b2_last_error.restype = C.c_char_p
    lib.b2_last_kernel.restype = C.c_char_p
    lib.b2_device_sms.restype = C.c_int
    lib.b2_build_leaves.argtypes = [vp, i32, i64, vp, i64, vp, i64, i32, i32, vp]
    lib.b2_build_leaves_indexed.argtypes = [vp, i32, vp, i64, vp, i64, vp, vp, i64, i32, i32, vp]
    lib.b2_contract_direct.argtypes = [C.POINTER(Modes), vp, vp, vp, dbl, dbl, i32, i32, vp]
    lib.b2_contract_gemm.argtypes = [C.POINTER(GemmModes), vp, vp, vp, dbl, dbl, i32, i32, vp]
    lib.b2_run_small_program.argtypes = [C.POINTER(SmallProgramArgs), i32, vp]
    lib.b2_run_plan.argtypes = [C.POINTER(PlanStep), i32, C.POINTER(i32), i32,
                                C.POINTER(SmallProgramArgs), vp, vp, vp, i64, i32, i64, i64,
                                dbl, dbl, i32, vp, C.POINTER(i32)]
    lib.b2_workspace_bytes.argtypes = [C.POINTER(PlanStep), i32, i32]
    lib.b2_conj.argtypes = [i64, vp, vp, i32, vp]
    lib.b2_bond_compress.argtypes = [vp, vp, i32, i32, dbl, vp, vp, vp, vp, vp, vp]
    lib.b2_bond_compress_scratch_bytes.argtypes = [i32]
    lib.b2_plan_greedy.argtypes = [i32, vp, vp, i32, vp, vp, C.c_uint64, dbl, dbl, vp]
    lib.b2_plan_cost.argtypes = [i32, vp, vp, i32, vp, vp, vp, vp, vp, vp]
    lib.b2_plan_reconfigure.argtypes = [i32, vp, vp, i32, vp, vp, vp, i32, vp, vp]
    lib.b2_axpy.argtypes = [i64, dbl, dbl, vp, vp, i32, vp]
    lib.b2_fill_zero.argtypes = [i64, vp, i32, vp]
    lib.b2_probs_amp.argtypes = [i64, vp, vp, vp, i32, vp]
    lib.b2_probs_dm.argtypes = [i32, C.POINTER(i32), C.POINTER(i32), vp, vp, vp, vp, i32, vp]
    lib.b2_compact_nonzero.argtypes = [i64, vp, i32, i64, vp, vp, vp, vp]
    lib.b2_permanent_amplitudes.argtypes = [vp, i32, vp, i32, vp, vp, i64, vp, vp]
    lib.b2_permanent_histogram.argtypes = [vp, i32, i32, vp, vp, dbl, i64, i32, vp, vp]
    lib.b2_fp64_peak_kernel.argtypes = [i32, i32, vp, C.POINTER(dbl), vp]
    for name in EXPORTS:
        fn = getattr(lib, name)
        if name not in ("b2_abi_version", "b2_last_error", "b2_last_kernel", "b2_device_sms",
                        "b2_workspace_bytes", "b2_bond_compress_scratch_bytes"):
            fn.restype = C.c_int
    lib.b2_workspace_bytes.restype = C.c_int64
    lib.b2_bond_compress_scratch_bytes.restype = C.c_int64
    if lib.b2_abi_version() != 1:
        raise B2Error("liboptyx_b200.so ABI version mismatch")
    _lib = lib
    return lib


def check(status: int, what: str = "") -> None:
    if status != 0:
        msg = load().b2_last_error().decode(errors="replace")
        raise B2Error(f"{what}: {msg}" if what else msg)


def dtype_code(dtype) -> int:
    dt = np.dtype(dtype)
    if dt == np.complex128:
        return B2_C128
    if dt == np.complex64:
        return B2_C64
    raise ValueError(f"dtype must be complex128 or complex64, got {dt}")
