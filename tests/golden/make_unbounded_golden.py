#!/usr/bin/env python
"""Golden vectors for the unbounded exit of the ratio test (theta = +inf, simplex.py:1231-1246), recorded from the
UNMODIFIED reference:   NUMBA_CACHE_DIR=/tmp/numba_cache python tests/golden/make_unbounded_golden.py
Instances with a negative-cost cycle of uncapacitated arcs; stored: the pivots up to the failing one, the arc the
reference names in UnboundedProblemError and the reduced cost it reports."""

from __future__ import annotations

import gzip
import io
import json
import random
import sys
from contextlib import redirect_stdout
from pathlib import Path

REPO = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(REPO / "tests" / "golden"))
import make_golden as mg  # noqa: E402

from network_solver import SolverOptions as RefOptions  # noqa: E402
from network_solver.exceptions import UnboundedProblemError as RefUnbounded  # noqa: E402
from network_solver.simplex import NetworkSimplex  # noqa: E402

from network_flow_solver_b200 import generators as gen  # noqa: E402


def run(problem, opts):
    solver = NetworkSimplex(problem, RefOptions(**opts))
    trace, pivot = [], solver._pivot

    def rec(arc_idx, direction):
        trace.append(int(arc_idx) * 2 + (1 if direction < 0 else 0))
        return pivot(arc_idx, direction)

    solver._pivot = rec
    with redirect_stdout(io.StringIO()):
        try:
            result = solver.solve()
        except RefUnbounded as exc:
            return {"options": opts, "status": "unbounded", "unbounded_arc": list(exc.entering_arc), "reduced_cost": float(exc.reduced_cost),
                    "message": str(exc), "trace": trace}
    return {"options": opts, "status": result.status, "iterations": result.iterations, "objective": result.objective, "trace": trace,
            "flows": [[k[0], k[1], v] for k, v in result.flows.items()], "duals": [[k, v] for k, v in result.duals.items()]}


def planted_cycle(n, m, seed, length=4, gain=-30.0):
    rng = random.Random(seed)
    p = gen.to_network_problem(gen.netgen_like(n, m, n_sources=4, n_sinks=4, cost_max=50, cap_max=40, supply_each=20, seed=seed))
    nodes = [{"id": x.id, "supply": x.supply} for x in p.nodes.values()]
    arcs = [{"tail": a.tail, "head": a.head, "capacity": a.capacity, "cost": a.cost} for a in p.arcs]
    ring = rng.sample([x["id"] for x in nodes], length)
    have = {(a["tail"], a["head"]) for a in arcs}
    for a, b in zip(ring, ring[1:] + ring[:1]):
        if (a, b) in have:
            arcs = [x for x in arcs if (x["tail"], x["head"]) != (a, b)]
        arcs.append({"tail": a, "head": b, "capacity": None, "cost": gain / length})
    rng.shuffle(arcs)
    return mg.ref_build(nodes, arcs, directed=True, tolerance=1e-6)


def main():
    DZ, DX = mg.DZ, mg.DX
    CL = {"pricing_strategy": "candidate_list", "explicit_pricing_strategy": True, "auto_scale": False}
    AD = {"auto_scale": False}
    cases = []

    def case(name, problem, variants=(DZ, DX, CL, AD)):
        runs = [run(problem, dict(v)) for v in variants]
        cases.append({"name": name, "problem": mg.problem_to_spec(problem), "runs": runs})
        print(name, [(r["status"], len(r["trace"]), r.get("unbounded_arc"), r.get("reduced_cost")) for r in runs], flush=True)

    case("triangle", mg.ref_build(
        [{"id": "a", "supply": 2.0}, {"id": "b", "supply": 0.0}, {"id": "c", "supply": -2.0}],
        [{"tail": "a", "head": "b", "capacity": None, "cost": -3.0}, {"tail": "b", "head": "c", "capacity": None, "cost": -3.0},
         {"tail": "c", "head": "a", "capacity": None, "cost": 1.0}], directed=True, tolerance=1e-6))
    case("triangle_capped_elsewhere", mg.ref_build(
        [{"id": "a", "supply": 2.0}, {"id": "b", "supply": 0.0}, {"id": "c", "supply": -2.0}, {"id": "d", "supply": 0.0}],
        [{"tail": "a", "head": "d", "capacity": 5.0, "cost": 1.0}, {"tail": "d", "head": "c", "capacity": 5.0, "cost": 1.0},
         {"tail": "a", "head": "b", "capacity": None, "cost": -3.0}, {"tail": "b", "head": "c", "capacity": None, "cost": -3.0},
         {"tail": "c", "head": "a", "capacity": None, "cost": 1.0}], directed=True, tolerance=1e-6))
    case("planted_cycle_48", planted_cycle(48, 300, 71))
    case("planted_cycle_128", planted_cycle(128, 900, 72, length=6, gain=-12.0))
    path = REPO / "tests" / "golden" / "next" / "unbounded.json.gz"
    with gzip.GzipFile(path, "wb", mtime=0) as fh:
        fh.write(json.dumps({"cases": cases}, separators=(",", ":")).encode())
    print(f"wrote {path} ({path.stat().st_size / 1024:.1f} KiB)")


if __name__ == "__main__":
    main()
