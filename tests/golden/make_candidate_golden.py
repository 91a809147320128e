#!/usr/bin/env python
"""Golden vectors for the reference's candidate-list / adaptive pricing, recorded from
the UNMODIFIED reference:  NUMBA_CACHE_DIR=/tmp/numba_cache python tests/golden/make_candidate_golden.py
They pin the restatement in oracle/nsx_oracle.c (cl_select / cl_scan / cl_refresh); the CUDA engine does not
implement this pricing rule yet."""

from __future__ import annotations

import gzip
import json
import sys
from pathlib import Path

REPO = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(REPO / "tests" / "golden"))
import make_golden as mg  # noqa: E402  (sets up sys.path for the reference)

from network_flow_solver_b200 import generators as gen  # noqa: E402

CL = {"pricing_strategy": "candidate_list", "explicit_pricing_strategy": True, "auto_scale": False}
AD = {"auto_scale": False}  # the reference's defaults: pricing_strategy="adaptive", not explicit
AD_EXPLICIT = {"pricing_strategy": "adaptive", "explicit_pricing_strategy": True, "auto_scale": False}


def ref_problem(arrays):
    p = gen.to_network_problem(arrays)
    nodes = [{"id": n.id, "supply": n.supply} for n in p.nodes.values()]
    arcs = [{"tail": a.tail, "head": a.head, "capacity": a.capacity, "cost": a.cost, "lower": a.lower} for a in p.arcs]
    return mg.ref_build(nodes, arcs, directed=True, tolerance=p.tolerance)


def main() -> None:
    cases = [
        ("netgen_64", gen.netgen_like(64, 512, n_sources=4, n_sinks=4, seed=11), [CL, AD]),
        ("netgen_256", gen.netgen_like(256, 2048, n_sources=8, n_sinks=8, seed=12), [CL, AD]),
        ("netgen_512_ties", gen.netgen_like(512, 4096, n_sources=8, n_sinks=8, cost_max=20, seed=13), [CL, AD_EXPLICIT]),
        ("gridgen_257", gen.gridgen_like(seed=808), [CL, AD]),
        ("goto_16", gen.goto_like(16, seed=3), [CL, AD]),           # AD: auto-detected GOTO structure -> Dantzig
        ("transport_24", gen.transportation(24, 24, cost_max=100, seed=5), [CL, AD]),  # row scan first
    ]
    out = []
    for name, arrays, option_sets in cases:
        problem = ref_problem(arrays)
        runs = [mg.run_reference(problem, dict(o)) for o in option_sets]
        runs.append(mg.run_reference(problem, dict(CL), max_iterations=max(5, runs[0]["iterations"] // 2)))
        out.append({"name": name, "problem": mg.problem_to_spec(problem), "runs": runs})
        print(name, [(r["status"], r["iterations"], r["strategy"], r["objective"]) for r in runs], flush=True)
    path = REPO / "tests" / "golden" / "next" / "candidate_list.json.gz"
    path.parent.mkdir(exist_ok=True)
    with gzip.GzipFile(path, "wb", mtime=0) as fh:
        fh.write(json.dumps({"cases": out}, separators=(",", ":")).encode())
    print(f"wrote {path} ({path.stat().st_size / 1024:.1f} KiB)")


if __name__ == "__main__":
    main()
