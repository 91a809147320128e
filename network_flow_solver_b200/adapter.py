"""Benchmark-harness plugin: the reference compares solvers through ``SolverAdapter`` classes
(benchmarks/solvers/base.py:25-79: classmethods ``solve(problem, timeout_s) -> SolverResult``, ``is_available()``,
``get_version()``; registered in benchmarks/solvers/__init__.py:27-61).  ``B200Adapter`` offers that interface for the
GPU engine.  The harness' own ``SolverResult`` / ``SolverAdapter`` classes live in the reference's ``benchmarks`` tree;
when they are importable the adapter subclasses / returns them, otherwise it falls back to structurally identical local
definitions (same field names), so the class works inside and outside the reference checkout.

Like the reference's own adapter (benchmarks/solvers/network_solver_adapter.py:19-75) errors are reported through
``status='error'`` instead of being raised, and the default ``SolverOptions()`` are used.
"""

from __future__ import annotations

import time
from dataclasses import dataclass

from . import __version__, _capi
from .data import NetworkProblem, SolverOptions
from .exceptions import DeviceEngineError, NetworkSolverError
from .solver import solve_min_cost_flow

try:  # inside the reference checkout
    from benchmarks.solvers.base import SolverAdapter as _Base, SolverResult  # type: ignore
except Exception:  # standalone

    @dataclass
    class SolverResult:  # benchmarks/solvers/base.py:9-22
        solver_name: str
        problem_name: str
        status: str
        objective: float | None
        solve_time_ms: float
        iterations: int | None
        error_message: str | None = None
        metadata: dict | None = None

    class _Base:
        name = "base"
        display_name = "Base Solver"
        description = "Abstract base solver"


class B200Adapter(_Base):
    name = "network_solver_b200"
    display_name = "Network Solver (B200 engine)"
    description = "Device-resident network simplex on NVIDIA B200 behind the network_solver API"

    @classmethod
    def solve(cls, problem: NetworkProblem, timeout_s: float = 60.0, options: SolverOptions | None = None) -> SolverResult:
        start = time.perf_counter()
        try:
            result = solve_min_cost_flow(problem, options if options is not None else SolverOptions())
        except NetworkSolverError as exc:
            return SolverResult(cls.name, "", "error", None, (time.perf_counter() - start) * 1e3, None, error_message=str(exc))
        elapsed = (time.perf_counter() - start) * 1e3
        return SolverResult(cls.name, "", result.status, result.objective if result.status == "optimal" else None,
                            elapsed, result.iterations, metadata={"flows": len(result.flows)})

    @classmethod
    def is_available(cls) -> bool:
        try:
            return _capi.load_library().nsx_device_count() > 0
        except DeviceEngineError:
            return False

    @classmethod
    def get_version(cls) -> str | None:
        return __version__
