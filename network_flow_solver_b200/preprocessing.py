"""Problem reductions applied before the device solve and undone afterwards.

Host-side caller of the hot path, same public surface and the same observable behaviour as the reference's
``preprocessing.py`` (reference: src/network_solver/preprocessing.py:27-57 result record, :59-195 driver,
:198-263 parallel-arc merge, :266-310 component count, :313-456 series contraction, :459-522 pendant removal,
:525-692 result translation, :695-759 ``preprocess_and_solve``).  The reduced instance is what goes to the GPU
engine, so the ORDER of the surviving nodes and arcs matters (it fixes the canonical arc indices and therefore the
pivot sequence): every pass below reproduces the reference's output order, including its key-based
(``(tail, head)``) bookkeeping of which original arc ended up where.  The restatement is pinned to recorded
reference outputs in ``tests/test_next_preprocessing.py``.

Passes, in the reference's order:
  1. parallel arcs with equal (tail, head, cost, lower) collapse into the first one, capacities added;
  2. connected components are counted (diagnostic only);
  3. zero-supply nodes with one arc in and one arc out are contracted (repeated until nothing changes);
  4. zero-supply nodes with a single incident arc are dropped together with that arc (one pass).
"""

from __future__ import annotations

import logging
import time
from dataclasses import dataclass, field
from typing import Any

from .data import Arc, FlowResult, NetworkProblem, Node

_log = logging.getLogger(__name__)

ArcKey = tuple[str, str]


@dataclass
class PreprocessingResult:
    """Reduced problem plus statistics and the original -> reduced maps (preprocessing.py:27-57).

    ``arc_mapping[i]`` is the ``(tail, head)`` key of the reduced arc that original arc ``i`` ended up in, or
    ``None`` when it was dropped; ``node_mapping[id]`` is the node's own id, or ``None`` when it was removed.
    """

    problem: NetworkProblem
    removed_arcs: int = 0
    removed_nodes: int = 0
    merged_arcs: int = 0
    redundant_arcs: int = 0
    disconnected_components: int = 0
    preprocessing_time_ms: float = 0.0
    optimizations: dict[str, int] = field(default_factory=dict)
    arc_mapping: dict[int, ArcKey | None] = field(default_factory=dict)
    node_mapping: dict[str, str | None] = field(default_factory=dict)


def _retarget(arc_mapping: dict[int, ArcKey | None] | None, old_keys: tuple[ArcKey, ...], new_key: ArcKey | None) -> None:
    """Every original arc currently filed under one of `old_keys` moves to `new_key` (the reference tracks arcs by
    their (tail, head) key, so parallel arcs that share a key move together: preprocessing.py:397-406, :504-511)."""
    if arc_mapping is None:
        return
    for i, key in arc_mapping.items():
        if key in old_keys:
            arc_mapping[i] = new_key


def _remove_redundant_arcs(problem: NetworkProblem, arc_mapping: dict[int, ArcKey | None] | None = None) -> int:
    """Pass 1 (preprocessing.py:198-263): one arc per distinct (tail, head, cost, lower), in order of first
    appearance; capacity = sum in index order, or unbounded as soon as one member is unbounded."""
    members: dict[tuple[str, str, float, float], list[int]] = {}
    for i, a in enumerate(problem.arcs):
        members.setdefault((a.tail, a.head, a.cost, a.lower), []).append(i)
    out: list[Arc] = []
    dropped = 0
    for (tail, head, cost, lower), idx in members.items():
        if len(idx) == 1:
            out.append(problem.arcs[idx[0]])
            continue
        caps = [problem.arcs[i].capacity for i in idx]
        capacity = None if any(c is None for c in caps) else sum(caps)
        out.append(Arc(tail=tail, head=head, capacity=capacity, cost=cost, lower=lower))
        if arc_mapping is not None:
            for i in idx:
                arc_mapping[i] = (tail, head)
        dropped += len(idx) - 1
    problem.arcs = out
    return dropped


def _detect_disconnected_components(problem: NetworkProblem) -> int:
    """Pass 2 (preprocessing.py:266-310): number of weakly connected components that contain a problem node."""
    rep: dict[str, str] = {}

    def find(x: str) -> str:
        root = x
        while rep.setdefault(root, root) != root:
            root = rep[root]
        while rep[x] != root:
            rep[x], x = root, rep[x]
        return root

    for a in problem.arcs:
        ra, rb = find(a.tail), find(a.head)
        if ra != rb:
            rep[ra] = rb
    return len({find(v) for v in problem.nodes})


def _simplify_series_arcs(
    problem: NetworkProblem,
    arc_mapping: dict[int, ArcKey | None] | None = None,
    node_mapping: dict[str, str | None] | None = None,
) -> tuple[int, int]:
    """Pass 3 (preprocessing.py:313-456): u -> v -> w with supply(v) = 0 and no other arc at v becomes u -> w with
    cost = sum, capacity = min, lower = max.  Rounds repeat until a round contracts nothing; inside a round the
    nodes are taken in dictionary order and a node is skipped when a neighbour was already contracted in the same
    round.  Survivors keep their order, the new arcs are appended in contraction order."""
    contracted_total = 0
    tol = problem.tolerance
    while True:
        arcs = problem.arcs
        arcs_in: dict[str, list[int]] = {}
        arcs_out: dict[str, list[int]] = {}
        for i, a in enumerate(arcs):
            arcs_out.setdefault(a.tail, []).append(i)
            arcs_in.setdefault(a.head, []).append(i)
        gone: set[str] = set()
        gone_order: list[str] = []
        used: set[int] = set()
        fresh: list[Arc] = []
        for v, node in problem.nodes.items():
            if abs(node.supply) > tol:
                continue
            ins, outs = arcs_in.get(v, ()), arcs_out.get(v, ())
            if len(ins) != 1 or len(outs) != 1:
                continue
            a, b = arcs[ins[0]], arcs[outs[0]]
            if a.tail == b.head:  # would become a self-loop
                continue
            # candidates were fixed at the start of the round; the neighbour test uses the nodes contracted so far
            if a.tail in gone or b.head in gone:
                continue
            if a.capacity is None:
                capacity = b.capacity
            elif b.capacity is None:
                capacity = a.capacity
            else:
                capacity = min(a.capacity, b.capacity)
            fresh.append(Arc(tail=a.tail, head=b.head, capacity=capacity, cost=a.cost + b.cost, lower=max(a.lower, b.lower)))
            _retarget(arc_mapping, ((a.tail, a.head), (b.tail, b.head)), (a.tail, b.head))
            if node_mapping is not None:
                node_mapping[v] = None
            gone.add(v)
            gone_order.append(v)
            used.add(ins[0])
            used.add(outs[0])
        if not gone_order:
            break
        for v in gone_order:
            del problem.nodes[v]
        problem.arcs = [
            a for i, a in enumerate(arcs) if i not in used and a.tail not in gone and a.head not in gone
        ] + fresh
        contracted_total += len(gone_order)
    return contracted_total, contracted_total


def _remove_zero_supply_nodes(
    problem: NetworkProblem,
    arc_mapping: dict[int, ArcKey | None] | None = None,
    node_mapping: dict[str, str | None] | None = None,
) -> int:
    """Pass 4 (preprocessing.py:459-522): zero-supply nodes touched by exactly one arc end (a self-loop counts
    twice) disappear with that arc.  Single pass - nodes that become pendant as a result stay."""
    touching: dict[str, list[int]] = {}
    for i, a in enumerate(problem.arcs):
        touching.setdefault(a.tail, []).append(i)
        touching.setdefault(a.head, []).append(i)
    tol = problem.tolerance
    victims = [
        v for v, node in problem.nodes.items() if abs(node.supply) <= tol and len(touching.get(v, ())) == 1
    ]
    if not victims:
        return 0
    doomed = {touching[v][0] for v in victims}
    if node_mapping is not None:
        for v in victims:
            node_mapping[v] = None
    for i in doomed:
        a = problem.arcs[i]
        _retarget(arc_mapping, ((a.tail, a.head),), None)
    for v in victims:
        del problem.nodes[v]
    problem.arcs = [a for i, a in enumerate(problem.arcs) if i not in doomed]
    return len(victims)


def preprocess_problem(
    problem: NetworkProblem,
    remove_redundant: bool = True,
    detect_disconnected: bool = True,
    simplify_series: bool = True,
    remove_zero_supply: bool = True,
) -> PreprocessingResult:
    """Reduce `problem` without changing its optimum (preprocessing.py:59-195).  The input is not modified."""
    t0 = time.time()
    reduced = NetworkProblem(
        directed=problem.directed,
        nodes={k: Node(id=k, supply=n.supply) for k, n in problem.nodes.items()},
        arcs=[Arc(tail=a.tail, head=a.head, capacity=a.capacity, cost=a.cost, lower=a.lower) for a in problem.arcs],
        tolerance=problem.tolerance,
    )
    res = PreprocessingResult(problem=reduced)
    res.node_mapping = {k: k for k in problem.nodes}
    res.arc_mapping = {i: (a.tail, a.head) for i, a in enumerate(problem.arcs)}
    if remove_redundant:
        n = _remove_redundant_arcs(reduced, res.arc_mapping)
        res.redundant_arcs = n
        res.removed_arcs += n
        res.optimizations["redundant_arcs_removed"] = n
        if n > 0:
            _log.info(f"Removed {n} redundant parallel arcs")
    if detect_disconnected:
        n = _detect_disconnected_components(reduced)
        res.disconnected_components = n
        res.optimizations["disconnected_components"] = n
        if n > 1:
            _log.warning(f"Detected {n} disconnected components - problem may be infeasible")
    if simplify_series:
        nodes_gone, merged = _simplify_series_arcs(reduced, res.arc_mapping, res.node_mapping)
        res.removed_nodes += nodes_gone
        res.merged_arcs = merged
        res.removed_arcs += merged
        res.optimizations["series_arcs_merged"] = merged
        res.optimizations["series_nodes_removed"] = nodes_gone
        if nodes_gone > 0:
            _log.info(f"Simplified {merged} series arcs, removed {nodes_gone} nodes")
    if remove_zero_supply:
        n = _remove_zero_supply_nodes(reduced, res.arc_mapping, res.node_mapping)
        res.removed_nodes += n
        res.removed_arcs += n
        res.optimizations["zero_supply_nodes_removed"] = n
        if n > 0:
            _log.info(f"Removed {n} zero-supply transshipment nodes")
    res.preprocessing_time_ms = (time.time() - t0) * 1000
    if res.removed_arcs > 0 or res.removed_nodes > 0:
        _log.info(
            f"Preprocessing complete: removed {res.removed_arcs} arcs, "
            f"{res.removed_nodes} nodes in {res.preprocessing_time_ms:.2f}ms"
        )
    return res


def _share_of_parallel_flow(flow: float, caps: list[float], position: int) -> float | None:
    """Part of a merged parallel arc's flow that original member `position` carries (preprocessing.py:585-633):
    unbounded members split it evenly (bounded ones get nothing, signalled by None), otherwise in proportion to
    capacity; all-zero capacities split evenly."""
    inf = float("inf")
    unbounded = [i for i, c in enumerate(caps) if c == inf]
    if len(unbounded) == len(caps):
        return flow / len(caps)
    if unbounded:
        return flow / len(unbounded) if position in unbounded else None
    total = sum(caps)
    if total > 0:
        return flow * (caps[position] / total)
    return flow / len(caps)


def translate_result(
    flow_result: FlowResult, preproc_result: PreprocessingResult, original_problem: NetworkProblem
) -> FlowResult:
    """Solution of the reduced problem -> flows / duals keyed like the original problem (preprocessing.py:525-692).

    Dropped arcs carry 0; every arc of a contracted chain carries the chain's flow; members of a merged parallel
    group share its flow (see `_share_of_parallel_flow`); flows of original arcs with the same key add up.  Duals
    of removed nodes are the mean of the values implied by their already-known neighbours (arcs in, then arcs
    out, original arc order), 0 without any.  The basis is not translated (None)."""
    arcs = original_problem.arcs
    mapping = preproc_result.arc_mapping
    sharing: dict[ArcKey | None, list[int]] = {}
    for i, key in mapping.items():
        sharing.setdefault(key, []).append(i)
    parallel_group: dict[ArcKey, bool] = {}
    flows: dict[ArcKey, float] = {}
    for i, arc in enumerate(arcs):
        own = (arc.tail, arc.head)
        key = mapping.get(i)
        if key is None:
            flows.setdefault(own, 0.0)
            continue
        f = flow_result.flows.get(key, 0.0)
        group = sharing[key]
        if key not in parallel_group:
            parallel_group[key] = all((arcs[j].tail, arcs[j].head) == key for j in group)
        part: float | None = f
        if len(group) > 1 and parallel_group[key]:
            caps = [arcs[j].capacity if arcs[j].capacity is not None else float("inf") for j in group]
            part = _share_of_parallel_flow(f, caps, group.index(i))
        if part is None:
            flows.setdefault(own, 0.0)
        else:
            flows[own] = flows.get(own, 0.0) + part
    duals: dict[str, float] = {}
    for v in original_problem.nodes:
        kept = preproc_result.node_mapping.get(v)
        if kept is not None:
            duals[v] = flow_result.duals.get(kept, 0.0)
    for v in original_problem.nodes:
        if preproc_result.node_mapping.get(v) is not None:
            continue
        implied = [duals[a.tail] + a.cost for a in arcs if a.head == v and a.tail in duals]
        implied += [duals[a.head] - a.cost for a in arcs if a.tail == v and a.head in duals]
        duals[v] = sum(implied) / len(implied) if implied else 0.0
    return FlowResult(
        objective=flow_result.objective,
        flows=flows,
        status=flow_result.status,
        iterations=flow_result.iterations,
        duals=duals,
        basis=None,
    )


def preprocess_and_solve(problem: NetworkProblem, **solve_kwargs: Any) -> tuple[PreprocessingResult, FlowResult]:
    """Reduce, solve the reduced instance on the GPU engine, translate back (preprocessing.py:695-759).  A
    `warm_start_basis` is dropped when the reductions changed the structure."""
    from .solver import solve_min_cost_flow

    pre = preprocess_problem(problem)
    changed = pre.removed_arcs > 0 or pre.removed_nodes > 0 or pre.merged_arcs > 0
    if changed and "warm_start_basis" in solve_kwargs:
        _log.warning(
            "Dropping warm_start_basis: preprocessing made structural changes "
            f"(removed {pre.removed_arcs} arcs, {pre.removed_nodes} nodes, merged {pre.merged_arcs} arc series). "
            "Basis from original problem is incompatible with preprocessed problem."
        )
        solve_kwargs = {k: v for k, v in solve_kwargs.items() if k != "warm_start_basis"}
    result = solve_min_cost_flow(pre.problem, **solve_kwargs)
    if changed:
        result = translate_result(result, pre, problem)
    return pre, result
