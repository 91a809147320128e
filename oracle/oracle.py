"""ctypes wrapper around oracle/libnsx_oracle.so (the CPU restatement in nsx_oracle.c).

TEST INFRASTRUCTURE ONLY - see the header of nsx_oracle.c.  It reuses the product's ctypes
struct declarations (the C structs are shared via include/nsx_b200.h) but nothing in the
product imports this module.
"""

from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

from network_flow_solver_b200._capi import (
    CallFrame,
    EngineOptions,
    NsxOptions,
    NsxProblem,
    NsxResult,
    NsxWarmStart,
    RawSolution,
)
from network_flow_solver_b200.canonical import CanonicalProblem

_HERE = Path(__file__).resolve().parent
_LIB = _HERE / "libnsx_oracle.so"
_lib = None


def build(force: bool = False) -> Path:
    src = _HERE / "nsx_oracle.c"
    header = _HERE.parent / "include" / "nsx_b200.h"
    if force or not _LIB.exists() or _LIB.stat().st_mtime < max(src.stat().st_mtime, header.stat().st_mtime):
        subprocess.run(["make", "-C", str(_HERE), "-B", "libnsx_oracle.so"], check=True,
                       capture_output=True)
    return _LIB


def _load():
    global _lib
    if _lib is None:
        build()
        lib = C.CDLL(str(_LIB))
        lib.nsx_oracle_solve.argtypes = [
            C.POINTER(NsxProblem), C.POINTER(NsxOptions), C.POINTER(NsxResult), C.c_int,
        ]
        lib.nsx_oracle_solve.restype = C.c_int
        lib.nsx_oracle_solve_warm.argtypes = [
            C.POINTER(NsxProblem), C.POINTER(NsxOptions), C.POINTER(NsxWarmStart), C.POINTER(NsxResult), C.c_int,
        ]
        lib.nsx_oracle_solve_warm.restype = C.c_int
        lib.nsx_oracle_max_threads.restype = C.c_int
        _lib = lib
    return _lib


def max_threads() -> int:
    return int(_load().nsx_oracle_max_threads())


def solve_canonical(cp: CanonicalProblem, opts: EngineOptions, threads: int = 1, warm=None) -> RawSolution:
    """Run the restatement on one canonical problem; `threads` parallelises the pricing sweep only
    (order-preserving reduction, result independent of the thread count); `warm` = warm_start.WarmStart."""
    lib = _load()
    frame = CallFrame(cp, opts)
    if warm is not None:
        w = NsxWarmStart.of(warm)
        rc = lib.nsx_oracle_solve_warm(C.byref(frame.problem), C.byref(frame.options), C.byref(w),
                                       C.byref(frame.result), int(threads))
        if rc != 0:
            raise RuntimeError(f"nsx_oracle_solve_warm returned {rc}")
        return frame.harvest()
    rc = lib.nsx_oracle_solve(C.byref(frame.problem), C.byref(frame.options),
                              C.byref(frame.result), int(threads))
    if rc != 0:
        raise RuntimeError(f"nsx_oracle_solve returned {rc}")
    return frame.harvest()
