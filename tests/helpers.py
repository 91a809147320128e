"""Shared helpers for the parity tests: golden-fixture loading and result comparison."""

from __future__ import annotations

import gzip
import json
from pathlib import Path

import numpy as np

from network_flow_solver_b200 import SolverOptions, build_problem
from network_flow_solver_b200 import _capi
from network_flow_solver_b200.solver import finish, prepare

GOLDEN = Path(__file__).resolve().parent / "golden"


def golden_names() -> list[str]:
    return sorted(p.name[: -len(".json.gz")] for p in GOLDEN.glob("*.json.gz"))


def load_golden(name: str) -> dict:
    with gzip.open(GOLDEN / f"{name}.json.gz", "rb") as fh:
        return json.loads(fh.read().decode())


def rebuild_problem(spec: dict):
    nodes = [{"id": nid, "supply": s} for nid, s in spec["nodes"]]
    arcs = [
        {"tail": t, "head": h, "capacity": cap, "cost": c, "lower": lo}
        for t, h, cap, c, lo in spec["arcs"]
    ]
    return build_problem(nodes, arcs, directed=spec["directed"], tolerance=spec["tolerance"])


def golden_cases():
    """(fixture name, run index) for every recorded reference run."""
    out = []
    for name in golden_names():
        doc = load_golden(name)
        for i in range(len(doc["runs"])):
            out.append((name, i))
    return out


def prepare_run(doc: dict, run: dict, trace_capacity: int = 1 << 20):
    problem = rebuild_problem(doc["problem"])
    options = SolverOptions(**run["options"])
    cp, plan, options = prepare(
        problem, options, run.get("max_iterations"), trace_capacity=trace_capacity
    )
    return problem, cp, plan, options


def assert_matches_reference(run: dict, cp, raw: _capi.RawSolution, options) -> None:
    """Bit-exact comparison of an engine/oracle solution with a recorded reference run."""
    if run["status"] == "unbounded":
        assert raw.status == _capi.STATUS_UNBOUNDED
        assert list(cp.arc_keys[raw.unbounded_arc]) == run["unbounded_arc"]
        return
    trace = raw.trace.tolist()
    ref_trace = run["trace"]
    if trace != ref_trace:
        k = next((i for i, (a, b) in enumerate(zip(trace, ref_trace)) if a != b), min(len(trace), len(ref_trace)))
        raise AssertionError(
            f"entering-arc sequence diverges at pivot {k}: got {trace[k:k+3]} want {ref_trace[k:k+3]} "
            f"(lengths {len(trace)} vs {len(ref_trace)})"
        )
    assert raw.iterations == run["iterations"]
    result = finish(cp, raw, options)
    assert result.status == run["status"]
    assert result.objective == run["objective"]
    assert {(a, b): v for a, b, v in run["flows"]} == result.flows
    assert {k: v for k, v in run["duals"]} == result.duals
    if run["status"] in ("optimal", "iteration_limit") and run["flows"] is not None and "internal_flow" in run:
        if result.flows or run["objective"] != 0.0 or run["status"] == "optimal":
            np.testing.assert_array_equal(raw.flow, np.asarray(run["internal_flow"]))
            np.testing.assert_array_equal(
                (raw.state & _capi.ARC_IN_TREE).astype(np.int64), np.asarray(run["internal_in_tree"])
            )
            np.testing.assert_array_equal(raw.potential, np.asarray(run["internal_potential"]))
            m = cp.n_arcs
            np.testing.assert_array_equal(
                ((raw.state[:m] & _capi.ARC_TOUCHED) != 0).astype(np.int64),
                np.asarray(run["internal_np_typed"]),
            )
            assert raw.degenerate_pivots == run["degenerate_pivots"]
            assert raw.final_block_size == run["final_block_size"]
