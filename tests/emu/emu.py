"""ctypes wrapper for tests/emu/libnsx_emu.so - the serial host emulation of the device pivot code.
Test infrastructure only (see nsx_emu.cpp)."""

from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

from network_flow_solver_b200._capi import CallFrame, EngineOptions, NsxOptions, NsxProblem, NsxResult, NsxWarmStart, RawSolution

_HERE = Path(__file__).resolve().parent
_LIB = _HERE / "libnsx_emu.so"
_CORE = _HERE.parents[1] / "network_flow_solver_b200" / "csrc" / "nsx_core.cuh"
_WARM = _HERE.parents[1] / "network_flow_solver_b200" / "csrc" / "nsx_warm.h"
_lib = None


_FLAGS = ["-fPIC", "-std=c++17", "-ffp-contract=off", "-fno-fast-math", "-x", "c++", "-shared"]


def build_mt(out: Path, sanitize: bool = False) -> Path:
    """The multi-threaded variant (NSX_HOST_MT): the pivot CTA as real threads (NSX_EMU_THREADS, default 8) with
    NSX_SYNC as a barrier; `sanitize` adds ThreadSanitizer.  Select it with NSX_EMU_LIB=<path>."""
    extra = ["-fsanitize=thread", "-O1", "-g"] if sanitize else ["-O2"]
    subprocess.run(["/usr/bin/g++", *extra, *_FLAGS, "-DNSX_HOST_MT", "-pthread", "-o", str(out), str(_HERE / "nsx_emu.cpp")],
                   check=True, capture_output=True)
    return out


def build(force: bool = False) -> Path:
    if os.environ.get("NSX_EMU_LIB"):
        return Path(os.environ["NSX_EMU_LIB"])
    src = _HERE / "nsx_emu.cpp"
    header = _HERE.parents[1] / "include" / "nsx_b200.h"
    newest = max(src.stat().st_mtime, _CORE.stat().st_mtime, _WARM.stat().st_mtime, header.stat().st_mtime)
    if force or not _LIB.exists() or _LIB.stat().st_mtime < newest:
        subprocess.run(
            ["/usr/bin/g++", "-O2", "-fPIC", "-std=c++17", "-ffp-contract=off", "-fno-fast-math",
             "-x", "c++", "-shared", "-o", str(_LIB), str(src)],
            check=True, capture_output=True,
        )
    return _LIB


def solve_canonical(cp, opts: EngineOptions, warm=None) -> RawSolution:
    global _lib
    if _lib is None:
        _lib = C.CDLL(str(build()))
        _lib.nsx_emu_solve.argtypes = [C.POINTER(NsxProblem), C.POINTER(NsxOptions), C.POINTER(NsxResult)]
        _lib.nsx_emu_solve.restype = C.c_int
        _lib.nsx_emu_solve_warm.argtypes = [C.POINTER(NsxProblem), C.POINTER(NsxOptions), C.POINTER(NsxWarmStart), C.POINTER(NsxResult)]
        _lib.nsx_emu_solve_warm.restype = C.c_int
    frame = CallFrame(cp, opts)
    if warm is not None:
        w = NsxWarmStart.of(warm)
        rc = _lib.nsx_emu_solve_warm(C.byref(frame.problem), C.byref(frame.options), C.byref(w), C.byref(frame.result))
        if rc != 0:
            raise RuntimeError(f"nsx_emu_solve_warm returned {rc}")
        return frame.harvest()
    rc = _lib.nsx_emu_solve(C.byref(frame.problem), C.byref(frame.options), C.byref(frame.result))
    if rc != 0:
        raise RuntimeError(f"nsx_emu_solve returned {rc}")
    return frame.harvest()
