"""JSON problem / result files (reference schema, io.py:33-69) and the benchmark-harness adapter surface."""

import json

import pytest

from helpers import load_golden, rebuild_problem
from network_flow_solver_b200 import FlowResult, InvalidProblemError, load_problem, save_result
from network_flow_solver_b200.adapter import B200Adapter, SolverResult


def test_load_problem_reads_the_reference_schema(tmp_path):
    spec = load_golden("ref_textbook_transport")["problem"]
    doc = {"directed": spec["directed"], "tolerance": spec["tolerance"],
           "nodes": [{"id": i, "supply": s} for i, s in spec["nodes"]],
           "edges": [{"tail": t, "head": h, "capacity": cap, "cost": c, "lower": lo} for t, h, cap, c, lo in spec["arcs"]]}
    path = tmp_path / "p.json"
    path.write_text(json.dumps(doc))
    got, want = load_problem(path), rebuild_problem(spec)
    assert (got.directed, got.tolerance) == (want.directed, want.tolerance)
    assert [(n.id, n.supply) for n in got.nodes.values()] == [(n.id, n.supply) for n in want.nodes.values()]
    assert [(a.tail, a.head, a.capacity, a.cost, a.lower) for a in got.arcs] == \
           [(a.tail, a.head, a.capacity, a.cost, a.lower) for a in want.arcs]
    doc["arcs"] = doc.pop("edges")  # the alternative key
    path.write_text(json.dumps(doc))
    assert len(load_problem(path).arcs) == len(want.arcs)


@pytest.mark.parametrize("doc", [{"nodes": []}, {"nodes": [], "edges": [{"tail": "a"}]}, {"edges": []}])
def test_load_problem_rejects_malformed_documents(tmp_path, doc):
    path = tmp_path / "bad.json"
    path.write_text(json.dumps(doc))
    with pytest.raises(InvalidProblemError):
        load_problem(path)


def test_save_result_layout(tmp_path):
    result = FlowResult(objective=12.5, flows={("b", "c"): 2.0, ("a", "b"): 3.0}, status="optimal", iterations=4,
                        duals={"b": 1.0, "a": 0.0})
    save_result(tmp_path / "r.json", result)
    doc = json.loads((tmp_path / "r.json").read_text())
    assert doc == {"status": "optimal", "objective": 12.5, "iterations": 4,
                   "flows": [{"tail": "a", "head": "b", "flow": 3.0}, {"tail": "b", "head": "c", "flow": 2.0}],
                   "duals": {"a": 0.0, "b": 1.0}}


def test_adapter_surface():
    assert B200Adapter.name and B200Adapter.display_name and isinstance(B200Adapter.get_version(), str)
    assert isinstance(B200Adapter.is_available(), bool)
    fields = {"solver_name", "problem_name", "status", "objective", "solve_time_ms", "iterations", "error_message", "metadata"}
    assert fields <= set(SolverResult.__dataclass_fields__)


@pytest.mark.gpu
def test_adapter_solves_on_the_gpu():
    spec = load_golden("ref_textbook_transport")
    r = B200Adapter.solve(rebuild_problem(spec["problem"]))
    assert r.status == "optimal" and r.objective == 85.0 and r.iterations > 0 and r.solve_time_ms > 0

