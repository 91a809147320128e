"""Host-side trigger test for the reference's automatic scaling (scaling.py:37-95).

Scaling itself is a pre/post step around the solver, outside the accelerated path; the drop-in
only needs to know whether the reference *would* have rescaled the instance, because that changes
every number the pivot loop sees.
"""

from __future__ import annotations

import math

from .data import NetworkProblem


def _spread_exceeds(values: list[float], threshold: float) -> bool:
    if len(values) < 2:
        return False
    lo, hi = min(values), max(values)
    return lo > 0 and hi / lo > threshold


def should_scale_problem(problem: NetworkProblem, threshold: float = 1e6) -> bool:
    costs = [abs(a.cost) for a in problem.arcs if a.cost != 0]
    caps = [
        a.capacity
        for a in problem.arcs
        if a.capacity is not None and math.isfinite(a.capacity) and a.capacity > 0
    ]
    supplies = [abs(n.supply) for n in problem.nodes.values() if n.supply != 0]
    if any(_spread_exceeds(group, threshold) for group in (costs, caps, supplies)):
        return True
    return _spread_exceeds(costs + caps + supplies, threshold)
