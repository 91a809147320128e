#!/usr/bin/env python
"""Golden vectors for the canonicalisation row SURVEY.md 8/a9 beyond the plain directed case: arcs with LOWER BOUNDS
(shifted out, simplex.py:392-432, added back in the result, simplex.py:1703-1721) and UNDIRECTED problems (edge = one arc
with lower = -capacity, data.py:162-223), recorded from the UNMODIFIED reference:
    NUMBA_CACHE_DIR=/tmp/numba_cache python tests/golden/make_bounds_golden.py"""

from __future__ import annotations

import gzip
import json
import random
import sys
from pathlib import Path

REPO = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(REPO / "tests" / "golden"))
import make_golden as mg  # noqa: E402

from network_flow_solver_b200 import generators as gen  # noqa: E402


def with_lower_bounds(n, m, seed, share=0.15):
    """netgen-like instance; a share of the arcs must carry at least `lower` units.  Supplies are rebalanced so that the
    forced flows are feasible: lower units enter at the tail and leave at the head through the node balances."""
    rng = random.Random(seed)
    p = gen.to_network_problem(gen.netgen_like(n, m, n_sources=4, n_sinks=4, cost_max=100, cap_max=60, supply_each=40, seed=seed))
    nodes = {x.id: x.supply for x in p.nodes.values()}
    arcs = []
    for a in p.arcs:
        lower = 0.0
        if rng.random() < share:
            lower = float(rng.randint(1, max(1, int(min(a.capacity, 6)))))
            nodes[a.tail] += lower   # the forced units are supplied at the tail ...
            nodes[a.head] -= lower   # ... and consumed at the head, so the instance stays balanced and feasible
        arcs.append({"tail": a.tail, "head": a.head, "capacity": a.capacity, "cost": a.cost, "lower": lower})
    return mg.ref_build([{"id": k, "supply": v} for k, v in nodes.items()], arcs, directed=True, tolerance=1e-6)


def undirected_grid(side, seed):
    rng = random.Random(seed)
    ids = [[f"g{r:02d}{c:02d}" for c in range(side)] for r in range(side)]
    nodes = {v: 0.0 for row in ids for v in row}
    edges = []
    for r in range(side):
        for c in range(side):
            if c + 1 < side:
                edges.append((ids[r][c], ids[r][c + 1]))
            if r + 1 < side:
                edges.append((ids[r][c], ids[r + 1][c]))
    rng.shuffle(edges)
    arcs = [{"tail": (a if rng.random() < 0.5 else b), "head": (b if rng.random() < 0.5 else a), "capacity": float(rng.randint(5, 30)),
             "cost": float(rng.randint(1, 40))} for a, b in edges]
    arcs = [x for x in arcs if x["tail"] != x["head"]]
    flat = [v for row in ids for v in row]
    for v in rng.sample(flat, 4):
        nodes[v] += 6.0
    for v in rng.sample(flat, 4):
        nodes[v] -= 6.0
    return mg.ref_build([{"id": k, "supply": v} for k, v in nodes.items()], arcs, directed=False, tolerance=1e-6)


def main():
    DZ, DX = mg.DZ, mg.DX
    CL = {"pricing_strategy": "candidate_list", "explicit_pricing_strategy": True, "auto_scale": False}
    AD = {"auto_scale": False}
    cases = []

    def case(name, problem, variants=(DZ, DX, CL, AD)):
        runs = [mg.run_reference(problem, dict(v)) for v in variants]
        cases.append({"name": name, "problem": mg.problem_to_spec(problem), "runs": runs})
        print(name, [(r["status"], r.get("iterations"), r.get("objective")) for r in runs], flush=True)

    case("lower_bounds_48", with_lower_bounds(48, 300, 51))
    case("lower_bounds_128", with_lower_bounds(128, 900, 52, share=0.3))
    case("lower_bound_equals_capacity", mg.ref_build(
        [{"id": "a", "supply": 5.0}, {"id": "b", "supply": 0.0}, {"id": "c", "supply": -5.0}],
        [{"tail": "a", "head": "b", "capacity": 5.0, "cost": 1.0, "lower": 5.0}, {"tail": "b", "head": "c", "capacity": 9.0, "cost": 2.0, "lower": 2.0},
         {"tail": "a", "head": "c", "capacity": 4.0, "cost": 1.0}], directed=True, tolerance=1e-6))
    und = lambda nodes, edges: mg.ref_build([{"id": k, "supply": float(v)} for k, v in nodes.items()],
                                            [{"tail": a, "head": b, "capacity": float(c), "cost": float(w)} for a, b, c, w in edges],
                                            directed=False, tolerance=1e-6)
    # the scenarios of the reference's tests/unit/test_undirected_graphs.py and tests/integration/test_undirected_performance.py
    case("undirected_chain_3", und({"A": 10, "B": 0, "C": -10}, [("A", "B", 15, 2), ("B", "C", 15, 3)]))               # objective 50
    case("undirected_against_orientation", und({"A": -10, "B": 10}, [("A", "B", 20, 4)]))                             # flow = -10 on (A, B)
    case("undirected_two_paths", und({"A": 10, "B": 0, "C": 0, "D": -10},
                                     [("A", "B", 8, 1), ("B", "D", 8, 2), ("A", "C", 8, 1.5), ("C", "D", 8, 1.5)]))
    case("undirected_parallel_edges", und({"S": 15, "T": -15}, [("S", "T", 10, 2), ("S", "T", 10, 5)]))
    case("undirected_star_mixed", und({"A": 4, "B": -9, "C": 5}, [("A", "B", 15, 1), ("A", "C", 15, 1)]))           # C -> A -> B: negative flow on (A, C)
    case("undirected_chain_150", und({**{f"n{i:03d}": 0 for i in range(150)}, "n000": 900, "n149": -900},
                                     [(f"n{i:03d}", f"n{i + 1:03d}", 1000, 2 + i % 5) for i in range(149)]), variants=(DZ, AD))
    case("undirected_grid_5", undirected_grid(5, 61))
    case("undirected_grid_9", undirected_grid(9, 62))
    path = REPO / "tests" / "golden" / "next" / "bounds_and_undirected.json.gz"
    with gzip.GzipFile(path, "wb", mtime=0) as fh:
        fh.write(json.dumps({"cases": cases}, separators=(",", ":")).encode())
    print(f"wrote {path} ({path.stat().st_size / 1024:.1f} KiB)")


if __name__ == "__main__":
    main()
