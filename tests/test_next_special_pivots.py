"""Scope row SURVEY.md section 8f-4 (third part): the structure-specific entering rules the reference switches on by
itself when it recognises an assignment, max-flow or shortest-path instance (specializations.py:187-288,
specialized_pivots.py:150-450, wiring simplex.py:259-261,1058-1075).  Vectors recorded from the unmodified reference
(tests/golden/make_special_golden.py), including its own failure modes on these classes (zero-cost max-flow instances end
"infeasible" / at the iteration limit because the artificial penalty max|c|*(N+2) is 0).  CPU: oracle and the emulated
device core; GPU: the CUDA engine."""

import gzip
import json
from pathlib import Path

import numpy as np
import pytest

from emu import emu
from helpers import assert_matches_reference, rebuild_problem
from network_flow_solver_b200 import SolverConfigurationError, SolverOptions, _capi, build_problem, solve_min_cost_flow
from network_flow_solver_b200.solver import finish, prepare, reachable_from
from oracle import oracle

DOC = json.loads(gzip.open(Path(__file__).resolve().parent / "golden" / "next" / "special_pivots.json.gz", "rb").read().decode())
CASES = {c["name"]: c for c in DOC["cases"]}
RUNS = [(c["name"], i) for c in DOC["cases"] for i in range(len(c["runs"]))]
RULE = {"assignment": _capi.SPECIAL_ASSIGNMENT, "max_flow": _capi.SPECIAL_MAX_FLOW, "shortest_path": _capi.SPECIAL_SHORTEST_PATH}


def setup(name, i):
    case, run = CASES[name], CASES[name]["runs"][i]
    cp, plan, options = prepare(rebuild_problem(case["problem"]), SolverOptions(**run["options"]), run.get("max_iterations"),
                                trace_capacity=1 << 16)
    return run, cp, plan, options


@pytest.mark.parametrize("name", sorted(CASES))
def test_structure_detection_and_rule_choice(name):
    run, cp, plan, options = setup(name, 0)
    kind = name.rsplit("_", 10)[0]
    kind = "assignment" if name.startswith("assignment") else "max_flow" if name.startswith("max_flow") else "shortest_path"
    assert cp.network_type == kind == run.get("network_type", kind)
    assert plan.engine.row_scan_first == RULE[kind]
    assert (plan.engine.node_mask is not None) == (kind == "shortest_path")
    if kind == "shortest_path":  # some nodes are cut off from the source on purpose: the mask is not all-ones
        assert 0 < int(plan.engine.node_mask[1:].sum()) < cp.n_nodes - 1


def test_generators_build_what_the_structure_test_recognises():
    """Array-native instances carry their network type as a declaration (no NetworkProblem objects to analyse): it must be
    what the structure test (specializations.py:187-288) says about the same instance."""
    from network_flow_solver_b200 import generators as gen
    from network_flow_solver_b200.canonical import detect_network_type

    for arrays in (gen.assignment(12, seed=1), gen.shortest_path(96, 400, seed=2), gen.max_flow(64, 300, seed=3),
                   gen.transportation(8, 12, seed=4), gen.netgen_like(64, 256, n_sources=4, n_sinks=4, seed=5)):
        assert detect_network_type(gen.to_network_problem(arrays, tolerance=1e-6)) == arrays.network_type, arrays.family


def test_reachability_mask():
    tail, head = np.array([1, 2, 2, 4, 5]), np.array([2, 3, 1, 2, 5])
    assert reachable_from(6, tail, head, 1).tolist() == [0, 1, 1, 1, 0, 0]
    assert reachable_from(6, tail, head, 4).tolist() == [0, 1, 1, 1, 1, 0]
    assert reachable_from(6, tail, head, 3).tolist() == [0, 0, 0, 1, 0, 0]


def _matching_typed_instances():
    """Instances the reference classifies as bipartite_matching (specializations.py:247-266: bipartite, no lower bounds,
    supplies in {+1, -1, 0}) with the objective the unmodified reference returned for them
    (status optimal; 3.0 under dantzig / devex / adaptive, 60.0 with default options; recorded 2026-10)."""
    nodes = [{"id": "a", "supply": 1.0}, {"id": "b", "supply": 1.0}, {"id": "x", "supply": -1.0}, {"id": "y", "supply": 0.0},
             {"id": "z", "supply": -1.0}]
    arcs = [{"tail": "a", "head": "x", "capacity": 1.0, "cost": 1.0}, {"tail": "b", "head": "y", "capacity": 1.0, "cost": 1.0},
            {"tail": "y", "head": "z", "capacity": 1.0, "cost": 1.0}, {"tail": "x", "head": "b", "capacity": 1.0, "cost": 2.0}]
    yield build_problem(nodes, arcs, directed=True, tolerance=1e-6), 3.0
    # three unit sources and three unit sinks on a 6 x 6 grid (right / down arcs, capacity 3)
    costs = [5, 6, 9, 1, 8, 4, 1, 3, 2, 6, 8, 4, 7, 9, 2, 4, 1, 4, 7, 5, 3, 7, 3, 2, 3, 8, 3, 3, 1, 1, 4, 4, 3, 3, 5, 6, 4, 9, 4,
             3, 4, 7, 5, 1, 6, 7, 3, 3, 5, 2, 6, 5, 1, 6, 2, 5, 6, 5, 8, 6]
    W, k, garcs = 6, 0, []
    for r in range(W):
        for c in range(W):
            if c + 1 < W:
                garcs.append({"tail": f"n{r}{c}", "head": f"n{r}{c + 1}", "capacity": 3.0, "cost": float(costs[k])}); k += 1
            if r + 1 < W:
                garcs.append({"tail": f"n{r}{c}", "head": f"n{r + 1}{c}", "capacity": 3.0, "cost": float(costs[k])}); k += 1
    supply = {"n00": 1.0, "n02": 1.0, "n20": 1.0, "n55": -1.0, "n53": -1.0, "n35": -1.0}
    gnodes = [{"id": f"n{r}{c}", "supply": supply.get(f"n{r}{c}", 0.0)} for r in range(W) for c in range(W)]
    yield build_problem(gnodes, garcs, directed=True, tolerance=1e-6), 60.0


def test_matching_typed_instances_fall_through_to_the_configured_rule(caplog):
    """The reference's matching rule is hash-seed dependent (no sequence to reproduce): the instance is NOT refused, it is
    solved with the configured pricing rule and must reach the reference's optimal objective."""
    from network_flow_solver_b200.canonical import NET_BIPARTITE_MATCHING

    for problem, objective in _matching_typed_instances():
        for strategy in ("dantzig", "devex", "adaptive"):
            with caplog.at_level("WARNING"):
                cp, plan, options = prepare(problem, SolverOptions(pricing_strategy=strategy))
            assert cp.network_type == NET_BIPARTITE_MATCHING and plan.engine.row_scan_first == _capi.SPECIAL_NONE
            assert "hash-seed dependent" in caplog.text
            for solve in (oracle.solve_canonical, emu.solve_canonical):
                result = finish(cp, solve(cp, plan.engine), options)
                assert result.status == "optimal" and result.objective == objective


@pytest.mark.gpu
def test_matching_typed_instances_on_the_gpu():
    for problem, objective in _matching_typed_instances():
        result = solve_min_cost_flow(problem, SolverOptions())
        assert result.status == "optimal" and result.objective == objective


@pytest.mark.parametrize("name,i", RUNS)
def test_oracle_matches_reference(name, i):
    run, cp, plan, options = setup(name, i)
    assert_matches_reference(run, cp, oracle.solve_canonical(cp, plan.engine), options)


@pytest.mark.parametrize("name,i", RUNS)
def test_emulated_device_core_matches_reference(name, i):
    run, cp, plan, options = setup(name, i)
    assert_matches_reference(run, cp, emu.solve_canonical(cp, plan.engine), options)


@pytest.mark.gpu
@pytest.mark.timeout(300, method="thread")
@pytest.mark.parametrize("name,i", RUNS)
def test_engine_matches_reference(name, i):
    run, cp, plan, options = setup(name, i)
    assert_matches_reference(run, cp, _capi.solve_canonical(cp, plan.engine), options)


@pytest.mark.gpu
@pytest.mark.timeout(300, method="thread")
def test_public_api_picks_the_rule_by_itself(capsys):
    for name in ("assignment_16", "max_flow_48_unit_cost", "shortest_path_96"):
        case = CASES[name]
        run = case["runs"][3]  # default options apart from auto_scale
        result = solve_min_cost_flow(rebuild_problem(case["problem"]), SolverOptions(**run["options"]))
        assert (result.status, result.iterations, result.objective) == (run["status"], run["iterations"], run["objective"])
        assert result.flows == {(a, b): v for a, b, v in run["flows"]}
