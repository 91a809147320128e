#!/usr/bin/env python
"""Golden vectors for the loop-based Devex rule (SolverOptions.use_vectorized_pricing=False, simplex_pricing.py:205-269),
recorded from the UNMODIFIED reference:
    NUMBA_CACHE_DIR=/tmp/numba_cache python tests/golden/make_devex_loop_golden.py"""

from __future__ import annotations

import gzip
import json
import sys
from pathlib import Path

REPO = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(REPO / "tests" / "golden"))
import make_golden as mg  # noqa: E402
import make_special_golden as sp  # noqa: E402

from network_flow_solver_b200 import generators as gen  # noqa: E402


def main():
    LX = {"pricing_strategy": "devex", "explicit_pricing_strategy": True, "auto_scale": False, "use_vectorized_pricing": False}
    cases = []

    def case(name, problem, variants, **kw):
        runs = [mg.run_reference(problem, dict(v), **kw) for v in variants]
        cases.append({"name": name, "problem": mg.problem_to_spec(problem), "runs": runs})
        print(name, [(r["status"], r.get("iterations"), r.get("objective"), r.get("row_scan")) for r in runs], flush=True)

    def fam(arrays, tol=1e-3):
        p = gen.to_network_problem(arrays)
        return mg.ref_build([{"id": n.id, "supply": n.supply} for n in p.nodes.values()],
                            [{"tail": a.tail, "head": a.head, "capacity": a.capacity, "cost": a.cost} for a in p.arcs],
                            directed=True, tolerance=tol)

    case("netgen_64", fam(gen.netgen_like(64, 512, n_sources=4, n_sinks=4, seed=11)), [LX, {**LX, "block_size": 64}, {**LX, "block_size": 500}])
    case("netgen_256", fam(gen.netgen_like(256, 2048, n_sources=8, n_sinks=8, seed=12)), [LX, {**LX, "block_size": 32}, {**LX, "block_size": 128}])
    case("netgen_128_ties", fam(gen.netgen_like(128, 1024, n_sources=4, n_sinks=4, cost_max=5, cap_max=3, supply_each=6, seed=14)), [LX, {**LX, "block_size": 100}])
    case("gridgen_257", fam(gen.gridgen_like()), [LX])
    case("goto_16", fam(gen.goto_like(16, seed=32)), [LX])
    case("transport_32", fam(gen.transportation(32, 32, cost_max=100, seed=22)), [LX, {**LX, "block_size": 200}])  # row scan first
    case("assignment_16", sp.assignment(16, 3), [LX])                                                               # assignment rule first
    case("netgen_64_limit", fam(gen.netgen_like(64, 512, n_sources=4, n_sinks=4, seed=11)), [LX], max_iterations=60)
    path = REPO / "tests" / "golden" / "next" / "devex_loop.json.gz"
    with gzip.GzipFile(path, "wb", mtime=0) as fh:
        fh.write(json.dumps({"cases": cases}, separators=(",", ":")).encode())
    print(f"wrote {path} ({path.stat().st_size / 1024:.1f} KiB)")


if __name__ == "__main__":
    main()
