"""Error convention of the drop-in boundary.

Same class names and constructor arguments as the reference's hierarchy
(reference: src/network_solver/exceptions.py:6-167) so that callers' ``except``
clauses keep working when they switch packages.
"""

from __future__ import annotations


class NetworkSolverError(Exception):
    """Root of every error raised by this package."""


class InvalidProblemError(NetworkSolverError):
    """Problem or option values are malformed (unbalanced supplies, bad arcs, ...)."""


class InfeasibleProblemError(NetworkSolverError):
    def __init__(self, message: str, iterations: int = 0):
        super().__init__(message)
        self.iterations = iterations


class UnboundedProblemError(NetworkSolverError):
    """Ratio test found no blocking arc (reference: simplex.py:1231-1246)."""

    def __init__(self, message, entering_arc=None, reduced_cost=None):
        super().__init__(message)
        self.entering_arc = entering_arc
        self.reduced_cost = reduced_cost


class NumericalInstabilityError(NetworkSolverError):
    def __init__(self, message: str, condition_number=None):
        super().__init__(message)
        self.condition_number = condition_number


class IterationLimitError(NetworkSolverError):
    def __init__(self, message, iterations=0, objective=None, status="unknown"):
        super().__init__(message)
        self.iterations = iterations
        self.objective = objective
        self.status = status


class SolverConfigurationError(NetworkSolverError):
    """Unsupported or inconsistent solver configuration."""


class DeviceEngineError(NetworkSolverError):
    """The CUDA engine (libnsx_b200.so) is missing or reported a device error.

    There is deliberately no CPU fallback behind ``solve_min_cost_flow``.
    """
