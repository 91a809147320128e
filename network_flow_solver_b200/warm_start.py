"""Warm start: the reference's ``Basis`` -> the initial spanning tree and flows of the device engine.

Host-side pre-step of the hot path.  Restates what the reference does between
``solve(warm_start_basis=...)`` and its first pricing call:

* ``NetworkSimplex._apply_warm_start_basis`` (simplex.py:740-911): validation, basis arcs into the tree, one
  artificial arc per remaining component (chosen by the sign of the component's net supply), arc-count check;
* ``NetworkSimplex._recompute_tree_flows`` (simplex.py:913-1021): tree-arc flows from flow conservation, bottom up,
  a node's children taken in ascending arc index (the order of the reference's adjacency lists);
* the phase decision of ``solve`` (simplex.py:1494-1530): no artificial arc in the tree => Phase 1 is skipped.

The result goes over the C ABI as ``nsx_warm_start`` (include/nsx_b200.h); the engine lays the tree out as its
preorder array and runs the resident pivot loop from there.  ``None`` means what a ``False`` return means in the
reference: the basis is unusable, cold start (the reference then rebuilds the all-artificial tree, simplex.py:1525-1529,
which is the engine's normal initial state).

Reference behaviour that IS reproduced, because it decides the pivot sequence: after a successful warm start the
reference never refreshes its NumPy residual mirrors (``_sync_vectorized_arrays`` is only called from
``_initialize_tree``, simplex.py:728; a pivot refreshes the arcs of its own cycle only, simplex.py:1276-1283), so its
ratio test sees the COLD-START flow of every initial tree arc until that arc has been on a pivot cycle.  The engine
marks those arcs ``NSX_ARC_STALE`` and evaluates their residuals the same way (csrc/nsx_core.cuh); with that, 64 recorded
reference warm solves (Dantzig, Devex, candidate list, adaptive; accepted and rejected bases) are reproduced pivot for
pivot (tests/test_next_warm_start.py).  Not reproduced: one recorded Devex run in which the reference crashes
("Failed to locate cycle path") because its vectorised search prices a basis arc as non-tree; the engine keeps true tree
flags there and returns the correct status.
"""

from __future__ import annotations

import logging
from dataclasses import dataclass

import numpy as np

from .canonical import CanonicalProblem
from .data import Basis

_log = logging.getLogger(__name__)


@dataclass
class WarmStart:
    """Initial engine state: arcs ``0..M-1`` real, ``M + v - 1`` the artificial arc of node ``v``."""

    in_tree: np.ndarray  # uint8[M + n - 1]
    flow: np.ndarray  # float64[M + n - 1]
    start_phase: int  # 1, or 2 when no artificial arc is needed
    basis_arcs: int = 0
    artificial_in_tree: int = 0


def _components(n: int, tail: np.ndarray, head: np.ndarray) -> np.ndarray:
    """Component label per node of the undirected graph (tail[i], head[i]); union-find with path halving."""
    rep = list(range(n))
    for a, b in zip(tail.tolist(), head.tolist()):
        while rep[a] != a:
            rep[a] = rep[rep[a]]
            a = rep[a]
        while rep[b] != b:
            rep[b] = rep[rep[b]]
            b = rep[b]
        if a != b:
            rep[a] = b
    out = np.empty(n, dtype=np.int64)
    for v in range(n):
        r = v
        while rep[r] != r:
            r = rep[r]
        rep[v] = r
        out[v] = r
    return out


def apply_basis(cp: CanonicalProblem, basis: Basis, tolerance: float) -> WarmStart | None:
    """Mirror of _apply_warm_start_basis + _recompute_tree_flows on the canonical arrays."""
    if len(basis.tree_arcs) == 0:
        _log.warning("Warm-start basis is empty. Falling back to cold start.")
        return None
    if cp.arc_keys is None:
        raise ValueError("a Basis is keyed by (tail id, head id) and this problem has no arc keys: use apply_tree_arcs")
    index_of = {key: i for i, key in enumerate(cp.arc_keys)}  # parallel arcs: the last one wins (simplex.py:766-769)
    chosen: list[int] = []
    for key in basis.tree_arcs:
        i = index_of.get(key)
        if i is None:
            _log.warning(f"Warm-start basis contains arc {key} not in current problem. Falling back to cold start.")
            return None
        chosen.append(i)
    flows = {i: basis.arc_flows[cp.arc_keys[i]] for i in chosen if cp.arc_keys[i] in basis.arc_flows}
    return apply_tree_arcs(cp, chosen, flows, tolerance)


def apply_tree_arcs(cp: CanonicalProblem, chosen, flows, tolerance: float) -> WarmStart | None:
    """Array-native form of `apply_basis`: `chosen` = indices of the real arcs of the previous tree (canonical arc order),
    `flows` = {arc index: previous flow} for the ones whose flow is known (it is only validated, simplex.py:785-803; the
    tree flows are recomputed from conservation).  For instances built without NetworkProblem objects
    (INTEGRATION.md section 5): `np.flatnonzero(raw.state[:M] & ARC_IN_TREE)` of the previous RawSolution."""
    chosen = [int(i) for i in chosen]
    n, m = cp.n_nodes, cp.n_arcs
    ma = m + n - 1
    tol = tolerance
    upper = cp.upper
    flow = np.zeros(ma, dtype=np.float64)
    in_tree = np.zeros(ma, dtype=np.uint8)
    for i in chosen:
        if i in flows:
            f = flows[i]
            key = cp.arc_keys[i] if cp.arc_keys is not None else i
            if f < 0.0 - tol:  # internal lower bounds are 0 after the shift (simplex.py:416-428)
                _log.warning(f"Warm-start basis has flow {f:.2f} below lower bound 0.00 on arc {key}. Falling back to cold start.")
                return None
            if f > upper[i] + tol:
                _log.warning(
                    f"Warm-start basis has flow {f:.2f} exceeding capacity {upper[i]:.2f} on arc {key}. Falling back to cold start."
                )
                return None
            flow[i] = f
        in_tree[i] = 1
    real = np.asarray(sorted(chosen), dtype=np.int64)
    tail, head = cp.tail, cp.head
    supply = np.asarray(cp.supply, dtype=np.float64)

    # one artificial arc per component that does not contain the root (simplex.py:828-869)
    label = _components(n, tail[real], head[real])
    comp_supply = np.zeros(n, dtype=np.float64)
    np.add.at(comp_supply, label[1:], supply[1:])  # unbuffered: accumulates in node order like the reference's sum()
    want = np.where(comp_supply > tol, 1, np.where(comp_supply < -tol, -1, 0))[label]  # per node: what its component needs
    kind = np.where(supply > tol, 1, -1)  # artificial arc of v: v -> root when supply > tol, else root -> v (simplex.py:645-698)
    nodes = np.arange(n)
    usable = (nodes > 0) & (label != label[0]) & ((want == 0) | (want == kind))
    cand = nodes[usable]
    _, first = np.unique(label[cand], return_index=True)  # lowest node index per component
    picked = np.sort(cand[first])
    in_tree[m + picked - 1] = 1
    total = len(real) + len(picked)
    if total != n - 1:
        _log.warning(f"Warm-start tree has {total} arcs, expected {n - 1}. Falling back to cold start.")
        return None

    # tree structure from the root (simplex.py:929-951) ...
    arcs = np.concatenate([real, m + picked - 1])
    a_tail = np.concatenate([tail[real], np.where(kind[picked] > 0, picked, 0)]).tolist()
    a_head = np.concatenate([head[real], np.where(kind[picked] > 0, 0, picked)]).tolist()
    a_upper = np.concatenate([upper[real], np.where(np.abs(supply[picked]) <= tol, np.inf, np.abs(supply[picked]))]).tolist()
    a_index = arcs.tolist()  # ascending: real arcs sorted, artificial arcs in node order
    incident: list[list[int]] = [[] for _ in range(n)]
    for k in range(len(a_index)):
        incident[a_tail[k]].append(k)
        incident[a_head[k]].append(k)
    parent = [-1] * n
    parent_arc = [-1] * n
    parent[0] = 0
    order = [0]
    for u in order:
        for k in incident[u]:
            v = a_head[k] if a_tail[k] == u else a_tail[k]
            if parent[v] < 0:
                parent[v] = u
                parent_arc[v] = k
                order.append(v)
    if len(order) != n:  # only possible with a cyclic "basis"; the reference would carry on with a broken tree
        _log.warning("Warm-start basis arcs contain a cycle. Falling back to cold start.")
        return None
    # ... and the flows that conservation forces on them, leaves first (simplex.py:956-1010)
    sup = supply.tolist()
    f_tree = [0.0] * len(a_index)
    for v in reversed(order[1:]):
        balance = sup[v]
        for k in incident[v]:  # children in ascending arc index
            if k == parent_arc[v]:
                continue
            balance = balance + f_tree[k] if a_tail[k] != v else balance - f_tree[k]
        k = parent_arc[v]
        need = balance if a_tail[k] == v else -balance
        if need < -tol or need > a_upper[k] + tol:
            _log.warning("Warm-start basis incompatible with current capacities. Falling back to cold start.")
            return None
        f_tree[k] = max(0.0, min(a_upper[k], need))
    flow[arcs] = np.asarray(f_tree, dtype=np.float64)
    artificial = int(len(picked))
    _log.info(f"Successfully applied warm-start basis with {len(chosen)} basis arcs")
    return WarmStart(
        in_tree=in_tree,
        flow=flow,
        start_phase=2 if artificial == 0 else 1,
        basis_arcs=len(chosen),
        artificial_in_tree=artificial,
    )
