"""Boundary types of the drop-in: problem, options and result containers.

These mirror the reference's public dataclasses field for field
(reference: src/network_solver/data.py:12-567) because they ARE the plug-in
interface of the hot path: ``solve_min_cost_flow(NetworkProblem, SolverOptions)
-> FlowResult``.  Nothing here runs on the device; the canonicaliser
(`canonical.py`) turns these objects into the structure-of-arrays the C ABI takes.
"""

from __future__ import annotations

import math
from collections.abc import Callable, Iterable, Sequence
from dataclasses import dataclass, field

from .exceptions import InvalidProblemError

_STRATEGIES = ("devex", "dantzig", "candidate_list", "adaptive")


@dataclass(frozen=True)
class Node:
    """A node with supply (>0), demand (<0) or neither (reference: data.py:12-37)."""

    id: str
    supply: float = 0.0


@dataclass(frozen=True)
class Arc:
    """A directed arc; ``capacity=None`` means uncapacitated (reference: data.py:40-88)."""

    tail: str
    head: str
    capacity: float | None
    cost: float
    lower: float = 0.0

    def __post_init__(self) -> None:
        if self.tail == self.head:
            raise InvalidProblemError(
                f"Self-loop detected on node '{self.tail}'. Self-loops are not supported."
            )
        if self.capacity is not None and self.capacity < self.lower:
            raise InvalidProblemError(
                f"Arc {self.tail} -> {self.head} has capacity ({self.capacity}) less than "
                f"lower bound ({self.lower})."
            )


@dataclass
class NetworkProblem:
    """A min-cost-flow instance (reference: data.py:91-223)."""

    directed: bool
    nodes: dict[str, Node]
    arcs: list[Arc]
    tolerance: float = 1e-3

    def validate(self) -> None:
        balance = sum(n.supply for n in self.nodes.values())
        if abs(balance) > self.tolerance:
            raise InvalidProblemError(
                f"Problem is unbalanced: total supply {balance:.6f} exceeds tolerance "
                f"{self.tolerance}."
            )
        known = self.nodes
        for a in self.arcs:
            if a.tail not in known:
                raise InvalidProblemError(f"Arc tail '{a.tail}' not found in node set.")
            if a.head not in known:
                raise InvalidProblemError(f"Arc head '{a.head}' not found in node set.")

    def undirected_expansion(self) -> Sequence[Arc]:
        """Undirected edge {u,v} cap C -> one arc (u,v) with lower = -C (reference: data.py:162-223)."""
        if self.directed:
            return tuple(self.arcs)
        out: list[Arc] = []
        for a in self.arcs:
            if a.capacity is None:
                raise InvalidProblemError(
                    f"Undirected edge {a.tail} -- {a.head} has infinite capacity; undirected "
                    f"graphs require finite capacity on all edges."
                )
            cap = float(a.capacity)
            custom = abs(a.lower) > 1e-12 and not math.isclose(
                a.lower, -cap, rel_tol=0.0, abs_tol=1e-12
            )
            if custom:
                raise InvalidProblemError(
                    f"Undirected edge {a.tail} -- {a.head} has custom lower bound ({a.lower}); "
                    f"undirected edges do not support custom lower bounds."
                )
            out.append(Arc(tail=a.tail, head=a.head, capacity=cap, cost=a.cost, lower=-cap))
        return tuple(out)


@dataclass
class Basis:
    """Spanning-tree basis handed back for warm starts (reference: data.py:226-266)."""

    tree_arcs: set[tuple[str, str]] = field(default_factory=set)
    arc_flows: dict[tuple[str, str], float] = field(default_factory=dict)


@dataclass
class FlowResult:
    """Solver output (reference: data.py:269-322)."""

    objective: float
    flows: dict[tuple[str, str], float] = field(default_factory=dict)
    status: str = "optimal"
    iterations: int = 0
    duals: dict[str, float] = field(default_factory=dict)
    basis: Basis | None = None


@dataclass(frozen=True)
class ProgressInfo:
    """Payload of the progress callback (reference: data.py:325-343)."""

    iteration: int
    max_iterations: int
    phase: int
    phase_iterations: int
    objective_estimate: float
    elapsed_time: float


ProgressCallback = Callable[[ProgressInfo], None]


@dataclass
class SolverOptions:
    """Solver configuration; same fields and defaults as the reference (data.py:459-529).

    Fields that only steer the reference's dense basis factorisation
    (projection_cache_size, use_dense_inverse, use_jit, adaptive_ft_*) are accepted
    and validated but have no effect here: on a spanning-tree basis the factorised
    quantity they serve equals a tree-path length.
    """

    max_iterations: int | None = None
    tolerance: float = 1e-6
    pricing_strategy: str = "adaptive"
    explicit_pricing_strategy: bool = False
    block_size: int | str | None = None
    ft_update_limit: int = 64
    projection_cache_size: int = 100
    auto_scale: bool = True
    adaptive_refactorization: bool = True
    condition_check_interval: int = 50
    condition_number_threshold: float = 1e12
    adaptive_ft_min: int = 20
    adaptive_ft_max: int = 200
    use_dense_inverse: bool | None = None
    use_vectorized_pricing: bool = True
    use_jit: bool = True

    def __post_init__(self) -> None:
        if self.tolerance <= 0:
            raise InvalidProblemError(f"Tolerance must be positive, got {self.tolerance}.")
        if self.pricing_strategy not in _STRATEGIES:
            raise InvalidProblemError(
                f"Invalid pricing strategy '{self.pricing_strategy}'. "
                f"Must be 'devex', 'dantzig', 'candidate_list', or 'adaptive'."
            )
        bs = self.block_size
        if bs is not None:
            if isinstance(bs, str):
                if bs != "auto":
                    raise InvalidProblemError(
                        f"Invalid block_size '{bs}'. Must be a positive integer, 'auto', or None."
                    )
            elif bs <= 0:
                raise InvalidProblemError(f"Block size must be positive, got {bs}.")
        if self.ft_update_limit <= 0:
            raise InvalidProblemError(
                f"FT update limit must be positive, got {self.ft_update_limit}."
            )
        if self.condition_number_threshold <= 1:
            raise InvalidProblemError(
                f"Condition number threshold must be > 1, got {self.condition_number_threshold}."
            )
        if self.adaptive_ft_min <= 0 or self.adaptive_ft_min > self.adaptive_ft_max:
            raise InvalidProblemError(
                f"Adaptive FT min must be positive and <= max, got min={self.adaptive_ft_min}, "
                f"max={self.adaptive_ft_max}."
            )
        if self.use_dense_inverse is None:
            # The reference resolves this to "not has_sparse_lu()"; no factorisation exists here.
            object.__setattr__(self, "use_dense_inverse", False)


def build_problem(
    nodes: Iterable[dict], arcs: Iterable[dict], directed: bool, tolerance: float
) -> NetworkProblem:
    """Assemble and validate a problem from plain dicts (reference: data.py:532-567)."""
    node_map: dict[str, Node] = {}
    for spec in nodes:
        nid = str(spec["id"])
        if nid in node_map:
            raise InvalidProblemError(f"Duplicate node id '{nid}'.")
        node_map[nid] = Node(id=nid, supply=float(spec.get("supply", 0.0)))
    arc_list: list[Arc] = []
    for spec in arcs:
        cap = spec.get("capacity")
        arc_list.append(
            Arc(
                tail=str(spec["tail"]),
                head=str(spec["head"]),
                capacity=None if cap is None else float(cap),
                cost=float(spec.get("cost", 0.0)),
                lower=float(spec.get("lower", 0.0)),
            )
        )
    problem = NetworkProblem(
        directed=directed, nodes=node_map, arcs=arc_list, tolerance=float(tolerance)
    )
    problem.validate()
    return problem
