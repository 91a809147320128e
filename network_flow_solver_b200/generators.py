"""Synthetic instance generators for the BASELINE.json configurations (array-native).

Every generator returns an ``ArcArrays`` record with 0-based node indices; node ``k`` is meant to
carry the zero-padded decimal id ``f"{k:0{w}d}"`` so that the reference's lexicographic id sort
(simplex.py:149) coincides with numeric order.  ``to_network_problem`` materialises the same
instance as reference-style objects for sizes where that is affordable, so the object path and
the array path can be checked against each other.

All values are integers stored as float64 (the reference has no integer arithmetic,
SURVEY.md section 0).  Families follow SURVEY.md section 8(d).
"""

from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from .canonical import (
    NET_ASSIGNMENT,
    NET_GENERAL,
    NET_MAX_FLOW,
    NET_SHORTEST_PATH,
    NET_TRANSPORTATION,
    PERTURB_EPS_BASE,
    CanonicalProblem,
    canonicalize_arrays,
)
from .data import Arc, NetworkProblem, Node


@dataclass
class ArcArrays:
    n_nodes: int  # problem nodes (no root)
    tail: np.ndarray
    head: np.ndarray
    cost: np.ndarray
    capacity: np.ndarray  # inf = uncapacitated
    supply: np.ndarray
    network_type: str = NET_GENERAL
    family: str = ""
    seed: int = 0

    @property
    def n_arcs(self) -> int:
        return int(self.tail.shape[0])

    def canonical(
        self, eps_base: float = PERTURB_EPS_BASE, tolerance: float = 1e-6
    ) -> CanonicalProblem:
        cp = canonicalize_arrays(
            self.n_nodes,
            self.tail,
            self.head,
            self.cost,
            self.capacity,
            self.supply,
            network_type=self.network_type,
            eps_base=eps_base,
            tolerance=tolerance,
        )
        cp.meta.update(family=self.family, seed=self.seed)
        return cp


def node_id(k: int, n_nodes: int) -> str:
    width = max(1, len(str(n_nodes - 1)))
    return f"{k:0{width}d}"


def to_network_problem(a: ArcArrays, tolerance: float = 1e-3) -> NetworkProblem:
    """Object form of the same instance (ids zero-padded so sorted order = numeric order)."""
    ids = [node_id(k, a.n_nodes) for k in range(a.n_nodes)]
    nodes = {ids[k]: Node(id=ids[k], supply=float(a.supply[k])) for k in range(a.n_nodes)}
    arcs = [
        Arc(
            tail=ids[int(t)],
            head=ids[int(h)],
            capacity=None if np.isinf(c) else float(c),
            cost=float(w),
        )
        for t, h, c, w in zip(a.tail.tolist(), a.head.tolist(), a.capacity.tolist(), a.cost.tolist())
    ]
    return NetworkProblem(directed=True, nodes=nodes, arcs=arcs, tolerance=tolerance)


def _dedupe(tail: np.ndarray, head: np.ndarray, n: int) -> tuple[np.ndarray, np.ndarray]:
    key = tail.astype(np.int64) * n + head.astype(np.int64)
    _, first = np.unique(key, return_index=True)
    first.sort()
    return tail[first], head[first]


def netgen_like(
    n_nodes: int,
    n_arcs: int,
    *,
    n_sources: int = 8,
    n_sinks: int = 8,
    supply_each: int = 1000,
    cost_max: int = 10000,
    cap_max: int = 1000,
    seed: int = 0,
) -> ArcArrays:
    """NETGEN-style sparse instance: a ring skeleton (capacity = total supply, so every source can
    reach every sink) plus uniformly random arcs; integer costs U[1,cost_max], capacities
    U[1,cap_max] (config 2: n=2^16, m=2^20, 256 sources/sinks, seed 1601)."""
    rng = np.random.default_rng(seed)
    ring_t = np.arange(n_nodes, dtype=np.int64)
    ring_h = (ring_t + 1) % n_nodes
    extra = max(0, n_arcs - n_nodes)
    t = rng.integers(0, n_nodes, size=int(extra * 1.05) + 16)
    h = rng.integers(0, n_nodes, size=t.shape[0])
    keep = t != h
    t, h = t[keep], h[keep]
    tail = np.concatenate([ring_t, t])
    head = np.concatenate([ring_h, h])
    tail, head = _dedupe(tail, head, n_nodes)
    tail, head = tail[:n_arcs], head[:n_arcs]
    m = tail.shape[0]
    cost = rng.integers(1, cost_max + 1, size=m).astype(np.float64)
    cap = rng.integers(1, cap_max + 1, size=m).astype(np.float64)
    total = float(n_sources * supply_each)
    cap[:n_nodes] = total  # ring arcs survive _dedupe at the front (first occurrences)
    picks = rng.permutation(n_nodes)[: n_sources + n_sinks]
    supply = np.zeros(n_nodes, dtype=np.float64)
    supply[picks[:n_sources]] = supply_each
    per_sink = total / n_sinks
    supply[picks[n_sources:]] = -per_sink
    return ArcArrays(n_nodes, tail, head, cost, cap, supply, NET_GENERAL, "netgen_like", seed)


def transportation(
    n_sources: int,
    n_sinks: int,
    *,
    cost_max: int = 1000,
    supply_each: int | None = None,
    capacity: float = np.inf,
    seed: int = 0,
) -> ArcArrays:
    """Dense transportation problem: arc i*n_sinks + j joins source i to sink j (config 3:
    4096 x 4096, costs U[1,1000], seed 4096).  Sources come first in id order."""
    rng = np.random.default_rng(seed)
    n = n_sources + n_sinks
    tail = np.repeat(np.arange(n_sources, dtype=np.int64), n_sinks)
    head = np.tile(np.arange(n_sources, n, dtype=np.int64), n_sources)
    cost = rng.integers(1, cost_max + 1, size=tail.shape[0]).astype(np.float64)
    cap = np.full(tail.shape[0], capacity, dtype=np.float64)
    each = n_sinks if supply_each is None else supply_each
    supply = np.zeros(n, dtype=np.float64)
    supply[:n_sources] = each
    total = each * n_sources
    base, rem = divmod(total, n_sinks)
    dem = np.full(n_sinks, base, dtype=np.float64)
    dem[:rem] += 1
    supply[n_sources:] = -dem
    return ArcArrays(n, tail, head, cost, cap, supply, NET_TRANSPORTATION, "transportation", seed)


def goto_like(
    side: int, *, cost_max: int = 10000, cap_max: int = 1000, seed: int = 0
) -> ArcArrays:
    """GOTO-style grid-on-torus: side*side nodes, 4 torus neighbours + 4 random jump arcs per node,
    one source and one sink (config 4: side=64 -> 4096 nodes / 32768 arcs)."""
    rng = np.random.default_rng(seed)
    n = side * side
    idx = np.arange(n, dtype=np.int64)
    r, c = idx // side, idx % side
    nbrs = [
        ((r + 1) % side) * side + c,
        ((r - 1) % side) * side + c,
        r * side + (c + 1) % side,
        r * side + (c - 1) % side,
    ]
    tails = [idx] * 4
    heads = list(nbrs)
    for _ in range(4):
        j = rng.integers(0, n, size=n)
        j = np.where(j == idx, (j + 1) % n, j)
        tails.append(idx)
        heads.append(j)
    tail = np.concatenate(tails)
    head = np.concatenate(heads)
    tail, head = _dedupe(tail, head, n)
    m = tail.shape[0]
    cost = rng.integers(1, cost_max + 1, size=m).astype(np.float64)
    cap = rng.integers(1, cap_max + 1, size=m).astype(np.float64)
    src, snk = (int(x) for x in rng.permutation(n)[:2])
    out_cap = float(cap[tail == src].sum())
    in_cap = float(cap[head == snk].sum())
    amount = float(max(1, int(0.25 * min(out_cap, in_cap))))
    supply = np.zeros(n, dtype=np.float64)
    supply[src] = amount
    supply[snk] = -amount
    return ArcArrays(n, tail, head, cost, cap, supply, NET_GENERAL, "goto_like", seed)


def gridgen_like(
    width: int = 16,
    n_arcs: int = 2056,
    *,
    n_sources: int = 16,
    n_sinks: int = 16,
    supply_each: int = 1000,
    cost_max: int = 10000,
    cap_max: int = 1000,
    seed: int = 808,
) -> ArcArrays:
    """GRIDGEN-8-style stand-in for gridgen_8_08a (absent from the reference tree): width^2 grid
    nodes plus one super node, grid skeleton (right/down, wrapping) with capacity = total supply,
    random extra arcs up to n_arcs (config 1: 257 nodes / 2056 arcs, seed 808)."""
    rng = np.random.default_rng(seed)
    g = width * width
    n = g + 1
    idx = np.arange(g, dtype=np.int64)
    r, c = idx // width, idx % width
    right = r * width + (c + 1) % width
    down = ((r + 1) % width) * width + c
    sup_t = np.array([g, 0], dtype=np.int64)  # super node joined to the grid both ways
    sup_h = np.array([0, g], dtype=np.int64)
    skel_t = np.concatenate([idx, idx, sup_t])
    skel_h = np.concatenate([right, down, sup_h])
    extra = max(0, n_arcs - skel_t.shape[0])
    t = rng.integers(0, n, size=int(extra * 1.2) + 16)
    h = rng.integers(0, n, size=t.shape[0])
    keep = t != h
    tail = np.concatenate([skel_t, t[keep]])
    head = np.concatenate([skel_h, h[keep]])
    tail, head = _dedupe(tail, head, n)
    tail, head = tail[:n_arcs], head[:n_arcs]
    m = tail.shape[0]
    cost = rng.integers(1, cost_max + 1, size=m).astype(np.float64)
    cap = rng.integers(1, cap_max + 1, size=m).astype(np.float64)
    total = float(n_sources * supply_each)
    cap[: skel_t.shape[0]] = total
    picks = rng.permutation(n)[: n_sources + n_sinks]
    supply = np.zeros(n, dtype=np.float64)
    supply[picks[:n_sources]] = supply_each
    supply[picks[n_sources:]] = -total / n_sinks
    return ArcArrays(n, tail, head, cost, cap, supply, NET_GENERAL, "gridgen_like", seed)


def assignment(n: int, *, cost_max: int = 200, seed: int = 0) -> ArcArrays:
    """Dense n x n assignment instance: unit supplies / demands, unit capacities, integer costs U[1, cost_max].  The
    reference recognises the structure (specializations.py:214-245) and runs its assignment pivot rule first."""
    rng = np.random.default_rng(seed)
    tail = np.repeat(np.arange(n, dtype=np.int64), n)
    head = n + np.tile(np.arange(n, dtype=np.int64), n)
    supply = np.concatenate([np.ones(n), -np.ones(n)])
    cost = rng.integers(1, cost_max + 1, size=n * n).astype(np.float64)
    return ArcArrays(2 * n, tail, head, cost, np.ones(n * n), supply, NET_ASSIGNMENT, "assignment", seed)


def shortest_path(n_nodes: int, n_arcs: int, *, cost_max: int = 500, cut_fraction: int = 8, seed: int = 0) -> ArcArrays:
    """One unit from a source to a sink over an uncapacitated NETGEN-style graph (ring skeleton + random arcs).  The last
    1/cut_fraction of the nodes keep their arcs into the rest but lose every arc coming from it, so they are not
    reachable from the source - the part of the graph the reference's shortest-path rule never prices forward
    (specialized_pivots.py:396-399).  Source and sink are placed in the reachable part."""
    a = netgen_like(n_nodes, n_arcs, n_sources=1, n_sinks=1, supply_each=1, cost_max=cost_max, seed=seed)
    first_cut = n_nodes - n_nodes // cut_fraction
    keep = ~((a.head >= first_cut) & (a.tail < first_cut))
    supply = np.zeros(n_nodes, dtype=np.float64)
    rng = np.random.default_rng(seed + 7)
    src, dst = rng.choice(first_cut, size=2, replace=False)
    supply[src], supply[dst] = 1.0, -1.0
    m = int(keep.sum())
    return ArcArrays(n_nodes, a.tail[keep], a.head[keep], a.cost[keep], np.full(m, np.inf), supply, NET_SHORTEST_PATH,
                     "shortest_path", seed)


def max_flow(n_nodes: int, n_arcs: int, *, flow: int = 40, seed: int = 0) -> ArcArrays:
    """`flow` units from one source to one sink, every arc at cost 1 (the uniform-cost form the reference classifies as
    max flow, specializations.py:268-286), capacities U[1, flow] on the random arcs and `flow` on the ring skeleton."""
    a = netgen_like(n_nodes, n_arcs, n_sources=1, n_sinks=1, supply_each=flow, cost_max=1, cap_max=flow, seed=seed)
    a.network_type = NET_MAX_FLOW
    a.family = "max_flow"
    return a
