#!/usr/bin/env python
"""Golden vectors for the structure-specific pivot rules (specialized_pivots.py:150-450),
recorded from the UNMODIFIED reference:
    NUMBA_CACHE_DIR=/tmp/numba_cache python tests/golden/make_special_golden.py
Assignment, max-flow and shortest-path instances (the structure test of specializations.py picks the rule), each under
Dantzig / Devex / candidate-list / default options: entering-arc trace, internal state and public result."""

from __future__ import annotations

import gzip
import json
import random
import sys
from pathlib import Path

REPO = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(REPO / "tests" / "golden"))
import make_golden as mg  # noqa: E402


def assignment(n, seed, density=1.0, cost_max=50):
    rng = random.Random(seed)
    nodes = [{"id": f"w{i:02d}", "supply": 1.0} for i in range(n)] + [{"id": f"j{i:02d}", "supply": -1.0} for i in range(n)]
    arcs = []
    for i in range(n):
        for j in range(n):
            if i == j or rng.random() < density:  # the diagonal keeps it feasible
                arcs.append({"tail": f"w{i:02d}", "head": f"j{j:02d}", "capacity": 1.0, "cost": float(rng.randint(1, cost_max))})
    return mg.ref_build(nodes, arcs, directed=True, tolerance=1e-6)


def max_flow(n, m, seed, flow, unit_cost):
    rng = random.Random(seed)
    ids = [f"v{i:03d}" for i in range(n)]
    nodes = [{"id": v, "supply": 0.0} for v in ids]
    nodes[0]["supply"], nodes[-1]["supply"] = float(flow), -float(flow)
    arcs, seen = [], set()
    for i in range(n - 1):  # a backbone path with room for the whole flow, then random extra arcs
        arcs.append({"tail": ids[i], "head": ids[i + 1], "capacity": float(flow), "cost": unit_cost})
        seen.add((i, i + 1))
    while len(arcs) < m:
        a, b = rng.randrange(n), rng.randrange(n)
        if a == b or (a, b) in seen:
            continue
        seen.add((a, b))
        arcs.append({"tail": ids[a], "head": ids[b], "capacity": float(rng.randint(1, flow)), "cost": unit_cost})
    rng.shuffle(arcs)
    return mg.ref_build(nodes, arcs, directed=True, tolerance=1e-6)


def shortest_path(n, m, seed, unreachable=4, cost_max=20, sink_in_core=True):
    rng = random.Random(seed)
    ids = [f"v{i:03d}" for i in range(n)]
    nodes = [{"id": v, "supply": 0.0} for v in ids]
    core = list(range(n - unreachable))  # the last nodes only have arcs INTO the core: not reachable from the source
    src, dst = 1, (core[-1] if sink_in_core else n - 2)
    nodes[src]["supply"], nodes[dst]["supply"] = 1.0, -1.0
    arcs, seen = [], set()
    order = [src] + [v for v in core if v != src]
    rng.shuffle(order[1:])
    for a, b in zip(order, order[1:]):  # a Hamiltonian path from the source keeps the core reachable
        arcs.append({"tail": ids[a], "head": ids[b], "capacity": None, "cost": float(rng.randint(1, cost_max))})
        seen.add((a, b))
    while len(arcs) < m:
        a, b = rng.randrange(n), rng.choice(core)
        if a == b or (a, b) in seen:
            continue
        seen.add((a, b))
        arcs.append({"tail": ids[a], "head": ids[b], "capacity": None if rng.random() < 0.7 else float(rng.randint(1, 3)),
                     "cost": float(rng.randint(1, cost_max))})
    rng.shuffle(arcs)
    return mg.ref_build(nodes, arcs, directed=True, tolerance=1e-6)


def main():
    DZ, DX = mg.DZ, mg.DX
    CL = {"pricing_strategy": "candidate_list", "explicit_pricing_strategy": True, "auto_scale": False}
    AD = {"auto_scale": False}
    cases = []

    def case(name, problem, variants=(DZ, DX, CL, AD), **kw):
        runs = [mg.run_reference(problem, dict(v), **kw) for v in variants]
        cases.append({"name": name, "problem": mg.problem_to_spec(problem), "runs": runs})
        print(name, runs[0].get("network_type"), [(r["status"], r.get("iterations"), r.get("objective"), r.get("row_scan")) for r in runs], flush=True)

    case("assignment_6", assignment(6, 1))
    case("assignment_12_ties", assignment(12, 2, cost_max=4))
    case("assignment_16", assignment(16, 3))
    case("assignment_24_sparse", assignment(24, 4, density=0.3))
    case("assignment_16_limit", assignment(16, 3), variants=(DZ, DX), max_iterations=20)
    case("max_flow_24_zero_cost", max_flow(24, 90, 5, 9, 0.0))
    case("max_flow_24_unit_cost", max_flow(24, 90, 6, 9, 1.0))
    case("max_flow_48_unit_cost", max_flow(48, 260, 7, 15, 1.0))
    case("shortest_path_20", shortest_path(20, 60, 8))
    case("shortest_path_40", shortest_path(40, 160, 9, unreachable=6))
    case("shortest_path_64_ties", shortest_path(64, 300, 10, unreachable=8, cost_max=3))
    case("shortest_path_96", shortest_path(96, 500, 11, unreachable=10, cost_max=40))
    case("shortest_path_20_sink_cut_off", shortest_path(20, 60, 8, sink_in_core=False))
    # more than 1024 arcs: the device scan walks these in several chunks
    case("assignment_40", assignment(40, 12), variants=(DZ, DX, AD))
    case("max_flow_200_unit_cost", max_flow(200, 1400, 13, 25, 1.0), variants=(DZ, AD))
    case("shortest_path_300", shortest_path(300, 1500, 14, unreachable=20, cost_max=60), variants=(DZ, DX, AD))
    path = REPO / "tests" / "golden" / "next" / "special_pivots.json.gz"
    with gzip.GzipFile(path, "wb", mtime=0) as fh:
        fh.write(json.dumps({"cases": cases}, separators=(",", ":")).encode())
    print(f"wrote {path} ({path.stat().st_size / 1024:.1f} KiB)")


if __name__ == "__main__":
    main()
