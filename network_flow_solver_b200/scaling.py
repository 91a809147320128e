"""Automatic problem scaling - the host pre/post step the reference wraps around its solver
(scaling.py:37-267, applied at simplex.py:103-114 and undone at simplex.py:1753-1756).

Restated, not copied: when ``SolverOptions.auto_scale`` is on (the default) and the value ranges of the instance
differ by more than 1e6 (within costs, capacities or supplies, or across them), costs are multiplied by the inverse
geometric mean of the non-zero |costs|, capacities and lower bounds by the inverse geometric mean of the finite positive
capacities, supplies by the inverse geometric mean of the non-zero |supplies|.  The pivot loop then runs on the scaled
instance - every number the device sees changes - and flows / objective are divided back afterwards (duals and the
basis stay in scaled units, as in the reference).  The geometric means use the same NumPy expression as the reference
so that the factors are bit-identical.
"""

from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np

from .data import Arc, NetworkProblem, Node


@dataclass
class ScalingFactors:
    cost_scale: float = 1.0
    capacity_scale: float = 1.0
    supply_scale: float = 1.0
    enabled: bool = False


def _groups(problem: NetworkProblem) -> tuple[list[float], list[float], list[float]]:
    costs = [abs(a.cost) for a in problem.arcs if a.cost != 0]
    caps = [
        a.capacity
        for a in problem.arcs
        if a.capacity is not None and math.isfinite(a.capacity) and a.capacity > 0
    ]
    supplies = [abs(n.supply) for n in problem.nodes.values() if n.supply != 0]
    return costs, caps, supplies


def _spread_exceeds(values: list[float], threshold: float) -> bool:
    if len(values) < 2:
        return False
    lo, hi = min(values), max(values)
    return lo > 0 and hi / lo > threshold


def should_scale_problem(problem: NetworkProblem, threshold: float = 1e6) -> bool:
    """scaling.py:37-95."""
    costs, caps, supplies = _groups(problem)
    if any(_spread_exceeds(group, threshold) for group in (costs, caps, supplies)):
        return True
    return _spread_exceeds(costs + caps + supplies, threshold)


def _inverse_geometric_mean(values: list[float]) -> float:
    if not values:
        return 1.0
    geo_mean = float(np.exp(np.mean(np.log(values))))
    return 1.0 / geo_mean if geo_mean > 0 else 1.0


def compute_scaling_factors(problem: NetworkProblem) -> ScalingFactors:
    """scaling.py:98-159."""
    costs, caps, supplies = _groups(problem)
    f = ScalingFactors(_inverse_geometric_mean(costs), _inverse_geometric_mean(caps), _inverse_geometric_mean(supplies))
    eps = 1e-10
    f.enabled = abs(f.cost_scale - 1.0) > eps or abs(f.capacity_scale - 1.0) > eps or abs(f.supply_scale - 1.0) > eps
    return f


def scale_problem(problem: NetworkProblem, f: ScalingFactors) -> NetworkProblem:
    """scaling.py:162-222: a scaled copy (infinite capacities stay infinite)."""
    if not f.enabled:
        return problem
    nodes = {nid: Node(id=n.id, supply=n.supply * f.supply_scale) for nid, n in problem.nodes.items()}
    arcs = []
    for a in problem.arcs:
        cap = a.capacity
        if cap is not None and math.isfinite(cap):
            cap = cap * f.capacity_scale
        arcs.append(Arc(tail=a.tail, head=a.head, capacity=cap, cost=a.cost * f.cost_scale, lower=a.lower * f.capacity_scale))
    return NetworkProblem(directed=problem.directed, nodes=nodes, arcs=arcs, tolerance=problem.tolerance)


def unscale_solution(flows: dict, objective, f: ScalingFactors):
    """scaling.py:225-267: flows / supply_scale, objective / (cost_scale * supply_scale)."""
    if not f.enabled:
        return flows, objective
    return {k: v / f.supply_scale for k, v in flows.items()}, objective / (f.cost_scale * f.supply_scale)
