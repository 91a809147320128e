#!/usr/bin/env python
"""Golden vectors for automatic scaling (scaling.py), recorded from the UNMODIFIED reference:
    NUMBA_CACHE_DIR=/tmp/numba_cache python tests/golden/make_scaling_golden.py
Instances whose value ranges differ by > 1e6, solved with auto_scale=True (the default): the reference rescales costs,
capacities and supplies, solves, and divides flows / objective back.  Stored: problem, options, the scaling factors the
reference computed, and the public result (trace in the scaled index space)."""

from __future__ import annotations

import gzip
import json
import sys
from pathlib import Path

REPO = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(REPO / "tests" / "golden"))
import make_golden as mg  # noqa: E402

import numpy as np  # noqa: E402
from network_solver.scaling import compute_scaling_factors, should_scale_problem  # noqa: E402

from network_flow_solver_b200 import generators as gen  # noqa: E402


def badly_scaled(n, m, seed, cost_mul, cap_mul):
    a = gen.netgen_like(n, m, n_sources=4, n_sinks=4, cost_max=1000, cap_max=500, seed=seed)
    p = gen.to_network_problem(a)
    nodes = [{"id": x.id, "supply": x.supply * cap_mul} for x in p.nodes.values()]
    arcs = [{"tail": x.tail, "head": x.head, "capacity": None if x.capacity is None else x.capacity * cap_mul,
             "cost": x.cost * cost_mul, "lower": x.lower} for x in p.arcs]
    return mg.ref_build(nodes, arcs, directed=True, tolerance=p.tolerance)


def main() -> None:
    cases = [("small_costs_big_caps", badly_scaled(64, 512, 21, 1e-4, 1e4)),
             ("big_costs", badly_scaled(128, 1024, 22, 1e5, 1.0)),
             ("mild", badly_scaled(64, 512, 23, 1.0, 1.0))]  # does not trigger scaling
    out = []
    for name, problem in cases:
        triggered = bool(should_scale_problem(problem))
        f = compute_scaling_factors(problem)
        runs = []
        for opts in ({}, {"pricing_strategy": "devex", "explicit_pricing_strategy": True},
                     {"pricing_strategy": "dantzig", "explicit_pricing_strategy": True}):
            runs.append(mg.run_reference(problem, dict(opts)))
        out.append({"name": name, "problem": mg.problem_to_spec(problem), "triggered": triggered,
                    "factors": [f.cost_scale, f.capacity_scale, f.supply_scale, bool(f.enabled)], "runs": runs})
        print(name, triggered, [f.cost_scale, f.capacity_scale, f.supply_scale],
              [(r["status"], r["iterations"], r["objective"]) for r in runs], flush=True)
    path = REPO / "tests" / "golden" / "next" / "scaling.json.gz"
    with gzip.GzipFile(path, "wb", mtime=0) as fh:
        fh.write(json.dumps({"cases": out}, separators=(",", ":")).encode())
    print(f"wrote {path} ({path.stat().st_size / 1024:.1f} KiB)")


if __name__ == "__main__":
    main()
