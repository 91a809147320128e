#!/usr/bin/env python
"""Golden vectors for warm starts (simplex.py:740-1039,1494-1530), recorded from the
UNMODIFIED reference:
    NUMBA_CACHE_DIR=/tmp/numba_cache python tests/golden/make_warm_start_golden.py
Each case: problem A is solved cold; its FlowResult.basis warm-starts problem B (A itself or an edited copy).  Stored:
the basis, whether the reference accepted it, the tree flags / flows it built before the first pricing call, whether it
skipped Phase 1, the entering-arc trace of the warm solve and its public result."""

from __future__ import annotations

import gzip
import io
import json
import sys
from contextlib import redirect_stdout
from pathlib import Path

REPO = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(REPO / "tests" / "golden"))
import make_golden as mg  # noqa: E402

import numpy as np  # noqa: E402
from network_solver import SolverOptions as RefOptions  # noqa: E402
from network_solver.data import Basis as RefBasis  # noqa: E402
from network_solver.simplex import NetworkSimplex  # noqa: E402

from network_flow_solver_b200 import generators as gen  # noqa: E402


def spec_to_ref(spec):
    return mg.ref_build([{"id": a, "supply": b} for a, b in spec["nodes"]],
                        [{"tail": t, "head": h, "capacity": cap, "cost": c, "lower": lo} for t, h, cap, c, lo in spec["arcs"]],
                        directed=spec["directed"], tolerance=spec["tolerance"])


def run_warm(problem, opt_kwargs, basis):
    solver = NetworkSimplex(problem, RefOptions(**opt_kwargs))
    trace, seen = [], {}
    pivot, apply = solver._pivot, solver._apply_warm_start_basis

    def rec_pivot(arc_idx, direction):
        trace.append(int(arc_idx) * 2 + (1 if direction < 0 else 0))
        return pivot(arc_idx, direction)

    def rec_apply(b):
        ok = apply(b)
        seen["applied"] = bool(ok)
        if ok:
            seen["in_tree"] = [int(a.in_tree) for a in solver.arcs]
            seen["flow"] = [float(a.flow) for a in solver.arcs]
            seen["artificial_in_tree"] = sum(1 for a in solver.arcs if a.in_tree and a.artificial)
        return ok

    solver._pivot, solver._apply_warm_start_basis = rec_pivot, rec_apply
    try:
        with redirect_stdout(io.StringIO()):
            result = solver.solve(warm_start_basis=basis)
    except RuntimeError as exc:  # the reference's own failure mode (stale pricing mirrors after a warm start)
        return {"options": opt_kwargs, **seen, "trace": trace, "status": "reference_error", "error": str(exc)}
    m = solver.actual_arc_count
    return {"options": opt_kwargs, **seen, "trace": trace, "status": result.status, "iterations": result.iterations,
            "objective": result.objective, "flows": [[k[0], k[1], v] for k, v in result.flows.items()],
            "duals": [[k, v] for k, v in result.duals.items()],
            "internal_flow": [float(a.flow) for a in solver.arcs], "internal_in_tree": [int(a.in_tree) for a in solver.arcs],
            "internal_potential": [float(p) for p in solver.basis.potential],
            "internal_np_typed": [int(isinstance(a.flow, np.floating)) for a in solver.arcs[:m]],
            "degenerate_pivots": int(solver.degenerate_pivots), "final_block_size": int(solver.adaptive_tuner.block_size),
            "basis_out": None if result.basis is None else
            {"tree_arcs": sorted(list(k) for k in result.basis.tree_arcs), "arc_flows": [[k[0], k[1], v] for k, v in sorted(result.basis.arc_flows.items())]}}


def edit(spec, **kw):
    s = json.loads(json.dumps(spec))
    for i, mul in kw.get("capacity_mul", {}).items():
        if s["arcs"][i][2] is not None:
            s["arcs"][i][2] = s["arcs"][i][2] * mul
    for i, c in kw.get("cost_set", {}).items():
        s["arcs"][i][3] = c
    for i, d in kw.get("supply_add", {}).items():
        s["nodes"][i][1] += d
    for i in sorted(kw.get("drop_arcs", []), reverse=True):
        del s["arcs"][i]
    for a in kw.get("add_arcs", []):
        s["arcs"].append(a)
    return s


def main():
    cases = []

    def case(name, spec_a, spec_b, variants, basis_edit=None):
        pa, pb = spec_to_ref(spec_a), spec_to_ref(spec_b)
        runs = []
        for opts in variants:
            with redirect_stdout(io.StringIO()):
                first = NetworkSimplex(pa, RefOptions(**opts)).solve()
            basis = first.basis
            if basis_edit is not None:
                basis = basis_edit(basis)
            rec = run_warm(pb, dict(opts), basis)
            rec["basis_in"] = {"tree_arcs": sorted(list(k) for k in basis.tree_arcs),
                               "arc_flows": [[k[0], k[1], v] for k, v in sorted(basis.arc_flows.items())]}
            rec["cold_iterations"] = first.iterations
            runs.append(rec)
            print(name, opts.get("pricing_strategy", "default"), "applied" if rec["applied"] else "REJECTED",
                  rec.get("artificial_in_tree"), rec["status"], rec.get("iterations"), "cold", first.iterations, rec.get("objective"), flush=True)
        cases.append({"name": name, "problem_a": spec_a, "problem": spec_b, "runs": runs})

    DZ, DX, CL = mg.DZ, mg.DX, {"pricing_strategy": "candidate_list", "explicit_pricing_strategy": True, "auto_scale": False}
    AD = {"auto_scale": False}
    allv = [DZ, DX, CL, AD]
    small = mg.problem_to_spec(mg.ref_build(
        [{"id": "s", "supply": 10.0}, {"id": "a", "supply": 0.0}, {"id": "b", "supply": 0.0}, {"id": "c", "supply": 0.0}, {"id": "t", "supply": -10.0}],
        [{"tail": "s", "head": "a", "capacity": 10.0, "cost": 5.0}, {"tail": "s", "head": "b", "capacity": 10.0, "cost": 4.0},
         {"tail": "a", "head": "c", "capacity": 10.0, "cost": 1.0}, {"tail": "b", "head": "c", "capacity": 10.0, "cost": 2.0},
         {"tail": "c", "head": "t", "capacity": 10.0, "cost": 1.0}], directed=True, tolerance=1e-6))
    case("small_identical", small, small, allv)
    case("small_supply_up", small, edit(small, supply_add={0: 3.0, 4: -3.0}, capacity_mul={0: 2, 1: 2, 2: 2, 3: 2, 4: 2}), allv)
    case("small_cost_flip", small, edit(small, cost_set={1: 9.0}), allv)
    case("small_arc_missing", small, edit(small, drop_arcs=[3]), allv)
    case("small_capacity_cut", small, edit(small, capacity_mul={4: 0.5}), allv)
    case("small_empty_basis", small, small, [DZ], basis_edit=lambda b: RefBasis(tree_arcs=set(), arc_flows={}))
    case("small_partial_basis", small, small, allv,
         basis_edit=lambda b: RefBasis(tree_arcs={("s", "b")}, arc_flows={("s", "b"): 10.0}))

    # found by scripts/fuzz_warm_start_vs_reference.py: the reference's stale residual mirrors let the first pivot push 7
    # units through an arc with 5 units of room, the flow update clamps it back and the conservation check after Phase 1
    # (simplex.py:1575-1598) turns the run into "infeasible" although the edited instance is feasible
    over = mg.problem_to_spec(mg.ref_build(
        [{"id": "v0", "supply": 5.0}, {"id": "v1", "supply": 2.0}, {"id": "v2", "supply": -3.0}, {"id": "v3", "supply": -4.0}],
        [{"tail": a, "head": b, "capacity": c, "cost": w} for a, b, c, w in
         [("v0", "v1", 7.0, 1.0), ("v2", "v1", 7.0, 0.0), ("v0", "v3", 6.0, 3.0), ("v3", "v1", 6.0, 4.0), ("v0", "v2", 5.0, 1.0),
          ("v1", "v2", 7.0, 2.0), ("v2", "v3", 7.0, 7.0)]], directed=True, tolerance=1e-6))
    case("overpush_breaks_conservation", over, edit(over, capacity_mul={3: 4.0 / 6.0}), [DZ, CL, AD])

    def fam(arrays, tol=1e-3):
        p = gen.to_network_problem(arrays)
        return mg.problem_to_spec(mg.ref_build([{"id": n.id, "supply": n.supply} for n in p.nodes.values()],
                                               [{"tail": a.tail, "head": a.head, "capacity": a.capacity, "cost": a.cost} for a in p.arcs],
                                               directed=True, tolerance=tol))

    n64 = fam(gen.netgen_like(64, 512, n_sources=4, n_sinks=4, seed=11))
    case("netgen_64_identical", n64, n64, allv)
    m64 = len(n64["arcs"])
    case("netgen_64_costs_changed", n64, edit(n64, cost_set={i: float(3 + (i * 7) % 40) for i in range(0, m64, 9)}), allv)
    case("netgen_64_capacity_up", n64, edit(n64, capacity_mul={i: 2.0 for i in range(0, m64, 5)}), allv)
    u64 = json.loads(json.dumps(n64))  # uncapacitated copy: no arc rests at its upper bound, so the basis is accepted
    for a in u64["arcs"]:
        a[2] = None
    case("uncap_64_identical", u64, u64, allv)
    case("uncap_64_costs_changed", u64, edit(u64, cost_set={i: float(3 + (i * 7) % 40) for i in range(0, m64, 9)}), allv)
    case("uncap_64_arcs_dropped", u64, edit(u64, drop_arcs=list(range(5, m64, 23))), allv)
    src = [i for i, (_, s) in enumerate(n64["nodes"]) if s > 0][:2]
    dst = [i for i, (_, s) in enumerate(n64["nodes"]) if s < 0][:2]
    case("netgen_64_supply_shift", n64, edit(n64, supply_add={src[0]: 2.0, src[1]: -2.0, dst[0]: -1.0, dst[1]: 1.0}), allv)
    n256 = fam(gen.netgen_like(256, 2048, n_sources=8, n_sinks=8, seed=12))
    for a in n256["arcs"]:
        a[2] = None
    m256 = len(n256["arcs"])
    case("uncap_256_costs_changed", n256, edit(n256, cost_set={i: float(1 + (i * 13) % 60) for i in range(0, m256, 17)}), [DZ, DX, CL, AD])
    s256 = [i for i, (_, s) in enumerate(n256["nodes"]) if s > 0][:2]
    d256 = [i for i, (_, s) in enumerate(n256["nodes"]) if s < 0][:2]
    case("uncap_256_supply_shift", n256, edit(n256, supply_add={s256[0]: 5.0, s256[1]: -5.0, d256[0]: -2.0, d256[1]: 2.0}), [DZ, CL, AD])
    t16 = fam(gen.transportation(16, 16, cost_max=100, seed=21))
    case("transport_16_identical", t16, t16, [DZ, DX])
    case("transport_16_costs_changed", t16, edit(t16, cost_set={i: float(1 + (i * 31) % 97) for i in range(0, 256, 7)}), [DZ, DX])
    g8 = fam(gen.goto_like(8, seed=31))
    case("goto_8_costs_changed", g8, edit(g8, cost_set={i: float(1 + (i * 11) % 50) for i in range(0, len(g8["arcs"]), 6)}), [DZ, AD])

    path = REPO / "tests" / "golden" / "next" / "warm_start.json.gz"
    with gzip.GzipFile(path, "wb", mtime=0) as fh:
        fh.write(json.dumps({"cases": cases}, separators=(",", ":")).encode())
    print(f"wrote {path} ({path.stat().st_size / 1024:.1f} KiB)")


if __name__ == "__main__":
    main()
