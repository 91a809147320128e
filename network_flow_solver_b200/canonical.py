"""Canonicaliser: NetworkProblem (or raw arrays) -> the engine's structure-of-arrays.

The device engine works in the reference's *internal index space* so that arc
indices, tie-breaks and therefore the entering-arc sequence coincide:

* node order: artificial root at 0, then ``sorted(node ids)`` as Python strings
  (reference: simplex.py:149-152);
* arc order: stable sort by ``(tail id, head id)`` strings, lower bounds shifted
  out (reference: simplex.py:392-432);
* cost perturbation ``c + 1e-10 * 1.00001**i`` with the growth factor multiplied
  up sequentially (reference: simplex.py:36-37,1431-1440);
* artificial-arc penalty ``max|c| * (n_nodes + 1)`` (reference: simplex.py:161-163);
* structure detection that makes the reference override the configured pricing
  rule (reference: specializations.py:60-288, simplex.py:1058-1064).

Everything here is host-side NumPy; nothing is priced or pivoted on the CPU.
"""

from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np

from .data import NetworkProblem
from .exceptions import InvalidProblemError

PERTURB_EPS_BASE = 1e-10  # reference: simplex.py:36
PERTURB_GROWTH = 1.00001  # reference: simplex.py:37

NET_GENERAL = "general"
NET_TRANSPORTATION = "transportation"
NET_ASSIGNMENT = "assignment"
NET_BIPARTITE_MATCHING = "bipartite_matching"
NET_MAX_FLOW = "max_flow"
NET_SHORTEST_PATH = "shortest_path"


@dataclass
class CanonicalProblem:
    """Arrays in the engine's index space.  ``n_nodes`` includes the root (index 0)."""

    n_nodes: int
    tail: np.ndarray  # int32[M]
    head: np.ndarray  # int32[M]
    orig_cost: np.ndarray  # float64[M]  (objective)
    pert_cost: np.ndarray  # float64[M]  (perturbed Phase-2 cost)
    upper: np.ndarray  # float64[M]  (+inf = uncapacitated), after the lower-bound shift
    shift: np.ndarray  # float64[M]  (= original lower bound)
    supply: np.ndarray  # float64[n_nodes], adjusted for the shift; supply[0] = 0
    penalty: float
    network_type: str = NET_GENERAL
    node_ids: list[str] | None = None  # index -> id (None for array-native problems)
    arc_keys: list[tuple[str, str]] | None = None
    n_supply_nodes: int = 0
    n_demand_nodes: int = 0
    meta: dict = field(default_factory=dict)

    @property
    def n_arcs(self) -> int:
        return int(self.tail.shape[0])


def perturbed_costs(cost: np.ndarray, eps_base: float = PERTURB_EPS_BASE) -> np.ndarray:
    """``cost[i] + eps_base * f_i`` with ``f_0 = 1, f_{i+1} = f_i * 1.00001`` (sequential products).

    np.cumprod accumulates left to right, which is the reference's loop order
    (simplex.py:1431-1438), so the factors are bit-identical.
    """
    m = cost.shape[0]
    if m == 0:
        return cost.astype(np.float64).copy()
    with np.errstate(over="ignore"):
        growth = np.empty(m, dtype=np.float64)
        growth[0] = 1.0
        if m > 1:
            growth[1:] = np.cumprod(np.full(m - 1, PERTURB_GROWTH, dtype=np.float64))
        return cost.astype(np.float64) + eps_base * growth


def _penalty(orig_cost: np.ndarray, n_nodes: int) -> float:
    max_cost = float(np.max(np.abs(orig_cost))) if orig_cost.size else 1.0
    return max_cost * (n_nodes + 1)


# ----------------------------------------------------------------------------------------------
# Structure detection (host mirror of specializations.py; only the *type* matters to the engine)
# ----------------------------------------------------------------------------------------------
def _is_bipartite(problem: NetworkProblem) -> bool:
    if not problem.nodes:
        return False
    adj: dict[str, set[str]] = {nid: set() for nid in problem.nodes}
    for a in problem.arcs:
        adj[a.tail].add(a.head)
        adj[a.head].add(a.tail)
    colour: dict[str, int] = {}
    for start in problem.nodes:
        if start in colour:
            continue
        colour[start] = 0
        frontier = [start]
        while frontier:
            nxt = []
            for u in frontier:
                want = 1 - colour[u]
                for v in adj[u]:
                    c = colour.get(v)
                    if c is None:
                        colour[v] = want
                        nxt.append(v)
                    elif c != want:
                        return False
            frontier = nxt
    return True


def detect_network_type(problem: NetworkProblem) -> str:
    """Classify the instance the way the reference does (specializations.py:187-288).

    Uses ``problem.tolerance`` (not the solver tolerance), like the reference.
    """
    tol = problem.tolerance
    sources = {nid for nid, n in problem.nodes.items() if n.supply > tol}
    sinks = {nid for nid, n in problem.nodes.items() if n.supply < -tol}
    n_trans = len(problem.nodes) - len(sources) - len(sinks)
    total_supply = sum(problem.nodes[s].supply for s in sources)
    total_demand = sum(abs(problem.nodes[s].supply) for s in sinks)
    balanced = abs(total_supply - total_demand) <= tol
    has_lower = any(a.lower > tol for a in problem.arcs)
    bipartite = _is_bipartite(problem)

    if n_trans == 0 and sources and sinks and bipartite and not has_lower:
        if all(a.tail in sources and a.head in sinks for a in problem.arcs):
            if balanced and len(sources) == len(sinks):
                unit_s = all(abs(problem.nodes[s].supply - 1.0) <= tol for s in sources)
                unit_d = all(abs(problem.nodes[s].supply + 1.0) <= tol for s in sinks)
                if unit_s and unit_d:
                    return NET_ASSIGNMENT
            return NET_TRANSPORTATION
    if len(sources) == 1 and len(sinks) == 1:
        s = next(iter(sources))
        t = next(iter(sinks))
        if (
            abs(problem.nodes[s].supply - 1.0) <= tol
            and abs(problem.nodes[t].supply + 1.0) <= tol
        ):
            return NET_SHORTEST_PATH
    if bipartite and not has_lower:
        if all(
            abs(abs(n.supply) - 1.0) <= tol or abs(n.supply) <= tol
            for n in problem.nodes.values()
        ):
            return NET_BIPARTITE_MATCHING
    if len(sources) == 1 and len(sinks) == 1 and not has_lower:
        if all(abs(a.cost) <= tol for a in problem.arcs) or all(
            abs(a.cost - 1.0) <= tol for a in problem.arcs
        ):
            return NET_MAX_FLOW
    return NET_GENERAL


def is_likely_goto(problem: NetworkProblem, n_arcs: int, tol: float) -> bool:
    """Grid-on-torus heuristic that flips the reference to Dantzig (simplex.py:329-363)."""
    n = len(problem.nodes)
    if n == 0:
        return False
    n_sup = sum(1 for v in problem.nodes.values() if v.supply > tol)
    n_dem = sum(1 for v in problem.nodes.values() if v.supply < -tol)
    trans_pct = (n - n_sup - n_dem) / n
    return (
        n_sup + n_dem <= 4
        and trans_pct > 0.98
        and (2 * n_arcs) / n >= 8
        and 6 <= (n_arcs / n) <= 12
    )


# ----------------------------------------------------------------------------------------------
# Object path
# ----------------------------------------------------------------------------------------------
def canonicalize(
    problem: NetworkProblem, tolerance: float, eps_base: float = PERTURB_EPS_BASE
) -> CanonicalProblem:
    """NetworkProblem -> CanonicalProblem in the reference's index space."""
    node_ids = ["__network_simplex_root__"] + sorted(problem.nodes.keys())
    index = {nid: i for i, nid in enumerate(node_ids)}
    n_nodes = len(node_ids)

    supply = [0.0] * n_nodes
    for i in range(1, n_nodes):
        supply[i] = problem.nodes[node_ids[i]].supply
    total = sum(supply)
    if abs(total) > tolerance:
        raise InvalidProblemError(
            f"Supplies do not balance after lower-bound adjustment: total supply "
            f"{total:.6f} exceeds tolerance {tolerance}. The sum of all node "
            f"supplies must equal zero for a valid flow problem."
        )

    arcs = list(problem.undirected_expansion())
    arcs.sort(key=lambda a: (a.tail, a.head))
    m = len(arcs)
    tail = np.empty(m, dtype=np.int32)
    head = np.empty(m, dtype=np.int32)
    cost = np.empty(m, dtype=np.float64)
    upper = np.empty(m, dtype=np.float64)
    shift = np.empty(m, dtype=np.float64)
    keys: list[tuple[str, str]] = []
    for i, a in enumerate(arcs):
        t = index[a.tail]
        h = index[a.head]
        lower = a.lower
        if a.capacity is None:
            up = math.inf
        else:
            up = float(a.capacity) - lower
            if up < -tolerance:
                raise InvalidProblemError(
                    f"Arc capacity ({a.capacity}) is less than lower bound ({lower}) "
                    f"for arc {a.tail} -> {a.head}. Capacity must be >= lower bound."
                )
            up = max(0.0, up)
        if lower:
            supply[t] -= lower
            supply[h] += lower
        tail[i] = t
        head[i] = h
        cost[i] = a.cost
        upper[i] = up
        shift[i] = lower
        keys.append((a.tail, a.head))

    tol_p = problem.tolerance
    return CanonicalProblem(
        n_nodes=n_nodes,
        tail=tail,
        head=head,
        orig_cost=cost,
        pert_cost=perturbed_costs(cost, eps_base),
        upper=upper,
        shift=shift,
        supply=np.asarray(supply, dtype=np.float64),
        penalty=_penalty(cost, n_nodes),
        network_type=detect_network_type(problem),
        node_ids=node_ids,
        arc_keys=keys,
        n_supply_nodes=sum(1 for v in problem.nodes.values() if v.supply > tol_p),
        n_demand_nodes=sum(1 for v in problem.nodes.values() if v.supply < -tol_p),
    )


# ----------------------------------------------------------------------------------------------
# Array-native path (instances too large to hold as Python objects)
# ----------------------------------------------------------------------------------------------
def canonicalize_arrays(
    n_problem_nodes: int,
    tail: np.ndarray,
    head: np.ndarray,
    cost: np.ndarray,
    capacity: np.ndarray,
    supply: np.ndarray,
    *,
    network_type: str = NET_GENERAL,
    eps_base: float = PERTURB_EPS_BASE,
    tolerance: float = 1e-6,
    presorted: bool = False,
) -> CanonicalProblem:
    """Build a CanonicalProblem from 0-based node indices.

    Node ``k`` (0-based) stands for an id whose string sort rank is ``k`` (e.g. zero-padded
    decimal ids), so it becomes engine node ``k + 1``.  Arcs are stably sorted by
    ``(tail, head)`` like the reference does with id strings.  ``capacity`` uses ``inf`` for
    uncapacitated arcs; lower bounds are zero on this path.
    """
    tail = np.asarray(tail)
    head = np.asarray(head)
    if not presorted and tail.shape[0] > 1:
        key = tail.astype(np.int64) * (n_problem_nodes + 1) + head.astype(np.int64)
        presorted = bool(np.all(key[1:] >= key[:-1]))
    if not presorted:
        order = np.lexsort((head, tail))  # stable, primary key = tail
        tail, head = tail[order], head[order]
        cost = np.asarray(cost)[order]
        capacity = np.asarray(capacity)[order]
    cost = np.ascontiguousarray(cost, dtype=np.float64)
    sup = np.zeros(n_problem_nodes + 1, dtype=np.float64)
    sup[1:] = supply
    total = float(np.sum(sup))
    if abs(total) > tolerance:
        raise InvalidProblemError(
            f"Supplies do not balance after lower-bound adjustment: total supply "
            f"{total:.6f} exceeds tolerance {tolerance}."
        )
    n_nodes = n_problem_nodes + 1
    return CanonicalProblem(
        n_nodes=n_nodes,
        tail=np.ascontiguousarray(tail, dtype=np.int32) + np.int32(1),
        head=np.ascontiguousarray(head, dtype=np.int32) + np.int32(1),
        orig_cost=cost,
        pert_cost=perturbed_costs(cost, eps_base),
        upper=np.ascontiguousarray(capacity, dtype=np.float64),
        shift=np.zeros(cost.shape[0], dtype=np.float64),
        supply=sup,
        penalty=_penalty(cost, n_nodes),
        network_type=network_type,
        n_supply_nodes=int(np.sum(sup > tolerance)),
        n_demand_nodes=int(np.sum(sup < -tolerance)),
    )


def initial_block_size(arc_count: int) -> int:
    """Static Devex block-size heuristic (reference: simplex_adaptive.py:70-96)."""
    if arc_count < 1000:
        return max(1, arc_count // 4)
    if arc_count < 10000:
        return max(1, arc_count // 8)
    return max(1, arc_count // 16)
