#!/usr/bin/env python
"""Generate golden vectors by running the UNMODIFIED reference in the build container.

    NUMBA_CACHE_DIR=/tmp/numba_cache python tests/golden/make_golden.py

The reference (pure Python, /root/reference/src) cannot travel to the GPU box, so its outputs
are committed here as small fixtures: for every case the entering-arc sequence (recorded by
wrapping NetworkSimplex._pivot), the internal per-arc flows, tree flags and node potentials, and
the public FlowResult (objective, flows, duals, status, iterations).  The problem itself is
stored too (node ids, arcs), so tests rebuild it without the generator.

Cases: (a) the reference's own example problems and their expected objectives
(tests/integration/test_solver_end_to_end.py:57-121), (b) the iteration pins of
tests/unit/test_simplex.py:57-102, (c) synthetic families of SURVEY.md section 8(d) at sizes
the reference finishes in seconds, under dantzig / devex(auto block) / devex(fixed block) and
the transportation override.
"""

from __future__ import annotations

import gzip
import io
import json
import os
import sys
from contextlib import redirect_stdout
from pathlib import Path

REPO = Path(__file__).resolve().parents[2]
REF = Path(os.environ.get("NSX_REFERENCE", "/root/reference"))
sys.path.insert(0, str(REPO))
sys.path.insert(0, str(REF / "src"))
os.environ.setdefault("NUMBA_CACHE_DIR", "/tmp/numba_cache")

import numpy as np  # noqa: E402
from network_solver import SolverOptions as RefOptions  # noqa: E402
from network_solver import build_problem as ref_build  # noqa: E402
from network_solver import load_problem as ref_load  # noqa: E402
from network_solver.exceptions import UnboundedProblemError as RefUnbounded  # noqa: E402
from network_solver.simplex import NetworkSimplex  # noqa: E402

from network_flow_solver_b200 import generators as gen  # noqa: E402

OUT = Path(__file__).resolve().parent


def problem_to_spec(problem) -> dict:
    return {
        "directed": problem.directed,
        "tolerance": problem.tolerance,
        "nodes": [[n.id, n.supply] for n in problem.nodes.values()],
        "arcs": [[a.tail, a.head, a.capacity, a.cost, a.lower] for a in problem.arcs],
    }


def run_reference(problem, opt_kwargs: dict, max_iterations=None) -> dict:
    options = RefOptions(**opt_kwargs)
    solver = NetworkSimplex(problem, options)
    trace: list[int] = []
    original = solver._pivot

    def recording_pivot(arc_idx, direction):
        trace.append(int(arc_idx) * 2 + (1 if direction < 0 else 0))
        return original(arc_idx, direction)

    solver._pivot = recording_pivot
    rec: dict = {"options": opt_kwargs, "max_iterations": max_iterations}
    with redirect_stdout(io.StringIO()):
        try:
            result = solver.solve(max_iterations=max_iterations)
        except RefUnbounded as exc:
            rec.update(status="unbounded", unbounded_arc=list(exc.entering_arc), trace=trace)
            return rec
    m = solver.actual_arc_count
    rec.update(
        status=result.status,
        iterations=result.iterations,
        objective=result.objective,
        flows=[[k[0], k[1], v] for k, v in result.flows.items()],
        duals=[[k, v] for k, v in result.duals.items()],
        trace=trace,
        strategy=type(solver.pricing_strategy).__name__,
        row_scan=solver.specialized_pivot_strategy is not None,
        network_type=solver.network_structure.network_type.value,
        internal_flow=[float(a.flow) for a in solver.arcs],
        internal_in_tree=[int(a.in_tree) for a in solver.arcs],
        internal_potential=[float(p) for p in solver.basis.potential],
        internal_np_typed=[int(isinstance(a.flow, np.floating)) for a in solver.arcs[:m]],
        degenerate_pivots=int(solver.degenerate_pivots),
        final_block_size=int(solver.adaptive_tuner.block_size),
        node_ids=solver.node_ids[1:],
    )
    return rec


def dump(name: str, problem, runs: list[dict], note: str = "") -> None:
    doc = {"name": name, "note": note, "problem": problem_to_spec(problem), "runs": runs}
    raw = json.dumps(doc, separators=(",", ":")).encode()
    path = OUT / f"{name}.json.gz"
    with gzip.GzipFile(path, "wb", mtime=0) as fh:
        fh.write(raw)
    print(f"{name}: {len(runs)} runs, {path.stat().st_size / 1024:.1f} KiB", flush=True)


DZ = {"pricing_strategy": "dantzig", "explicit_pricing_strategy": True, "auto_scale": False}
DX = {"pricing_strategy": "devex", "explicit_pricing_strategy": True, "auto_scale": False}


def main() -> None:
    only = set(sys.argv[1:])

    def want(name):
        return not only or name in only

    # (a) reference example problems (golden objectives 8.0 / 85.0 / 100.0 / 20.0)
    for stem in ("dimacs_small", "textbook_transport", "large_transport", "sample"):
        name = f"ref_{stem}"
        if not want(name):
            continue
        problem = ref_load(REF / "examples" / f"{stem}_problem.json")
        runs = [run_reference(problem, DZ), run_reference(problem, DX)]
        dump(name, problem, runs, note=f"examples/{stem}_problem.json")

    # (b) iteration pins of tests/unit/test_simplex.py:57-102 (devex: 4 -> limit, 5 -> optimal)
    if want("ref_iteration_pins"):
        nodes = [
            {"id": "s", "supply": 10.0},
            {"id": "a", "supply": 0.0},
            {"id": "b", "supply": 0.0},
            {"id": "c", "supply": 0.0},
            {"id": "t", "supply": -10.0},
        ]
        arcs = [
            {"tail": "s", "head": "a", "capacity": 10.0, "cost": 5.0},
            {"tail": "s", "head": "b", "capacity": 10.0, "cost": 4.0},
            {"tail": "a", "head": "c", "capacity": 10.0, "cost": 1.0},
            {"tail": "b", "head": "c", "capacity": 10.0, "cost": 2.0},
            {"tail": "c", "head": "t", "capacity": 10.0, "cost": 1.0},
        ]
        problem = ref_build(nodes, arcs, directed=True, tolerance=1e-6)
        runs = []
        for k in (1, 2, 3, 4, 5, 6, 100):
            runs.append(run_reference(problem, DX, max_iterations=k))
            runs.append(run_reference(problem, DZ, max_iterations=k))
        dump("ref_iteration_pins", problem, runs)

    # (c) synthetic families
    def family(name, arrays, variants):
        if not want(name):
            return
        problem = gen.to_network_problem(arrays)
        # rebuild through the reference's own builder types
        rp = ref_build(
            [{"id": n.id, "supply": n.supply} for n in problem.nodes.values()],
            [
                {"tail": a.tail, "head": a.head, "capacity": a.capacity, "cost": a.cost}
                for a in problem.arcs
            ],
            directed=True,
            tolerance=1e-3,
        )
        runs = [run_reference(rp, v) for v in variants]
        dump(name, rp, runs, note=f"{arrays.family} seed={arrays.seed}")

    family("netgen_64", gen.netgen_like(64, 512, n_sources=4, n_sinks=4, seed=11),
           [DZ, DX, {**DX, "block_size": 64}, {**DX, "block_size": 500}])
    family("netgen_256", gen.netgen_like(256, 2048, n_sources=8, n_sinks=8, seed=12),
           [DZ, DX, {**DX, "block_size": 32}, {**DX, "block_size": 128}])
    family("netgen_512", gen.netgen_like(512, 4096, n_sources=8, n_sinks=8, seed=13), [DZ, DX])
    family("gridgen_257", gen.gridgen_like(), [DX, DZ])
    family("transport_16", gen.transportation(16, 16, cost_max=100, seed=21), [DZ, DX])
    family("transport_32", gen.transportation(32, 32, cost_max=100, seed=22), [DZ, DX])
    family("transport_48", gen.transportation(48, 48, cost_max=100, seed=23), [DX])
    family("transport_24x40", gen.transportation(24, 40, cost_max=1000, supply_each=50, seed=24),
           [DZ])
    family("goto_8", gen.goto_like(8, seed=31), [DZ, {**DZ, "explicit_pricing_strategy": False}])
    family("goto_16", gen.goto_like(16, seed=32), [DZ, DX])
    # a small-cost-range sparse instance: many integer ties, exercises the perturbation ordering
    family("netgen_128_ties",
           gen.netgen_like(128, 1024, n_sources=4, n_sinks=4, cost_max=5, cap_max=3,
                           supply_each=6, seed=14),
           [DZ, DX])


if __name__ == "__main__":
    main()
