"""Scope row SURVEY.md section 8f-4 (second half): the preprocessing wrappers around the solve (preprocessing.py).
Vectors recorded from the unmodified reference (tests/golden/make_preprocessing_golden.py): the reduced problem
(node and arc ORDER included - it fixes the canonical arc indices the engine pivots on), the statistics, both maps,
the reference's solve of the reduced problem and the translated result.  CPU: reductions + translation against the
recording, pivot loop by the oracle.  GPU: preprocess_and_solve() end to end through the CUDA engine."""

import gzip
import json
from pathlib import Path

import pytest

from helpers import rebuild_problem
from network_flow_solver_b200 import FlowResult, SolverOptions, _capi
from network_flow_solver_b200.preprocessing import (
    PreprocessingResult,
    preprocess_and_solve,
    preprocess_problem,
    translate_result,
)
from network_flow_solver_b200.solver import finish, prepare
from oracle import oracle

DOC = json.loads(gzip.open(Path(__file__).resolve().parent / "golden" / "next" / "preprocessing.json.gz", "rb").read().decode())
CASES = {c["name"]: c for c in DOC["cases"]}
RUNS = [(c["name"], i) for c in DOC["cases"] for i in range(len(c["runs"]))]


def spec_of(problem):
    return {
        "directed": problem.directed,
        "tolerance": problem.tolerance,
        "nodes": [[n.id, n.supply] for n in problem.nodes.values()],
        "arcs": [[a.tail, a.head, a.capacity, a.cost, a.lower] for a in problem.arcs],
    }


@pytest.mark.parametrize("name", sorted(CASES))
def test_reductions_match_the_reference(name):
    case = CASES[name]
    problem = rebuild_problem(case["problem"])
    before = spec_of(problem)
    pre = preprocess_problem(problem)
    assert isinstance(pre, PreprocessingResult)
    assert spec_of(problem) == before  # the input is not modified
    assert spec_of(pre.problem) == case["reduced"]  # same nodes, same arcs, same order
    stats = case["stats"]
    assert (pre.removed_arcs, pre.removed_nodes, pre.merged_arcs, pre.redundant_arcs, pre.disconnected_components) == (
        stats["removed_arcs"], stats["removed_nodes"], stats["merged_arcs"], stats["redundant_arcs"], stats["disconnected_components"])
    assert pre.optimizations == stats["optimizations"]
    assert [[i, None if k is None else list(k)] for i, k in pre.arc_mapping.items()] == case["arc_mapping"]
    assert [[k, v] for k, v in pre.node_mapping.items()] == case["node_mapping"]


def test_individual_passes_can_be_switched_off():
    problem = rebuild_problem(CASES["decorated_48"]["problem"])
    off = preprocess_problem(problem, remove_redundant=False, detect_disconnected=False, simplify_series=False, remove_zero_supply=False)
    assert spec_of(off.problem) == spec_of(problem) and off.removed_arcs == 0 and off.optimizations == {}
    only_parallel = preprocess_problem(problem, simplify_series=False, remove_zero_supply=False)
    assert only_parallel.removed_arcs == CASES["decorated_48"]["stats"]["redundant_arcs"]
    assert len(only_parallel.problem.nodes) == len(problem.nodes)


def final_of(case, run, inner: FlowResult):
    problem = rebuild_problem(case["problem"])
    pre = preprocess_problem(problem)
    return translate_result(inner, pre, problem) if case["changed"] else inner


def assert_final(result: FlowResult, want: dict):
    assert (result.status, result.iterations, result.objective) == (want["status"], want["iterations"], want["objective"])
    assert list(result.flows.items()) == [((a, b), v) for a, b, v in want["flows"]]  # values AND key order
    assert list(result.duals.items()) == [(k, v) for k, v in want["duals"]]


@pytest.mark.parametrize("name,i", RUNS)
def test_translation_of_the_recorded_inner_result(name, i):
    case, run = CASES[name], CASES[name]["runs"][i]
    inner = run["inner"]
    fr = FlowResult(objective=inner["objective"], flows={(a, b): v for a, b, v in inner["flows"]}, status=inner["status"],
                    iterations=inner["iterations"], duals=dict(inner["duals"]))
    assert_final(final_of(case, run, fr), run["final"])


def solve_reduced(case, run, solve):
    pre = preprocess_problem(rebuild_problem(case["problem"]))
    cp, plan, options = prepare(pre.problem, SolverOptions(**run["inner"]["options"]), trace_capacity=1 << 16)
    raw = solve(cp, plan.engine)
    assert raw.trace.tolist() == run["inner"]["trace"]
    return finish(cp, raw, options, plan.scaling)


@pytest.mark.parametrize("name,i", RUNS)
def test_oracle_on_the_reduced_instance_then_translation(name, i):
    case, run = CASES[name], CASES[name]["runs"][i]
    inner = solve_reduced(case, run, lambda cp, eng: oracle.solve_canonical(cp, eng))
    assert_final(final_of(case, run, inner), run["final"])


@pytest.mark.gpu
@pytest.mark.parametrize("name,i", RUNS)
def test_engine_on_the_reduced_instance_then_translation(name, i):
    case, run = CASES[name], CASES[name]["runs"][i]
    inner = solve_reduced(case, run, lambda cp, eng: _capi.solve_canonical(cp, eng))
    assert_final(final_of(case, run, inner), run["final"])


@pytest.mark.parametrize("backend", ["oracle", pytest.param("engine", marks=pytest.mark.gpu)])
@pytest.mark.parametrize("name", ["chain_of_three", "decorated_96", "nothing_to_do"])
def test_preprocess_and_solve_end_to_end(name, backend, capsys, monkeypatch):
    if backend == "oracle":  # test-only stand-in for the C-ABI call: runs the host half of the public call on CPU
        from network_flow_solver_b200 import solver as solver_module

        monkeypatch.setattr(solver_module._capi, "solve_canonical",
                            lambda cp, opts, out=None, warm=None: oracle.solve_canonical(cp, opts, warm=warm))
    case = CASES[name]
    for run in case["runs"]:
        pre, result = preprocess_and_solve(rebuild_problem(case["problem"]), options=SolverOptions(**run["inner"]["options"]))
        assert pre.removed_arcs == case["stats"]["removed_arcs"]
        assert_final(result, run["final"])
        assert (result.basis is None) == case["changed"]
