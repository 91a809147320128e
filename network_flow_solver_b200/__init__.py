"""network_flow_solver_b200 - B200-native network-simplex engine behind the reference's API.

Drop-in for the hot path of jeffreyhorn/network_flow_solver:
``solve_min_cost_flow(NetworkProblem, SolverOptions) -> FlowResult``.
"""

from .data import (
    Arc,
    Basis,
    FlowResult,
    NetworkProblem,
    Node,
    ProgressCallback,
    ProgressInfo,
    SolverOptions,
    build_problem,
)
from .exceptions import (
    DeviceEngineError,
    InfeasibleProblemError,
    InvalidProblemError,
    IterationLimitError,
    NetworkSolverError,
    NumericalInstabilityError,
    SolverConfigurationError,
    UnboundedProblemError,
)
from .dimacs import load_dimacs_canonical, parse_dimacs_file, parse_dimacs_string
from .io import load_problem, save_result
from .preprocessing import PreprocessingResult, preprocess_and_solve, preprocess_problem, translate_result
from .solver import solve_min_cost_flow

__version__ = "0.1.0"

__all__ = [
    "Arc",
    "Basis",
    "DeviceEngineError",
    "FlowResult",
    "InfeasibleProblemError",
    "InvalidProblemError",
    "IterationLimitError",
    "NetworkProblem",
    "NetworkSolverError",
    "Node",
    "NumericalInstabilityError",
    "PreprocessingResult",
    "ProgressCallback",
    "ProgressInfo",
    "SolverConfigurationError",
    "SolverOptions",
    "UnboundedProblemError",
    "build_problem",
    "load_dimacs_canonical",
    "load_problem",
    "parse_dimacs_file",
    "parse_dimacs_string",
    "preprocess_and_solve",
    "preprocess_problem",
    "save_result",
    "solve_min_cost_flow",
    "translate_result",
    "__version__",
]
