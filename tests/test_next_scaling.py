"""Scope row SURVEY.md section 8f-4 (first half): automatic problem scaling, the host pre/post step the reference wraps
around its solver when value ranges differ by more than 1e6 (scaling.py, applied with the DEFAULT options).  Vectors
recorded from the unmodified reference (tests/golden/make_scaling_golden.py): the scaling factors, and - with the
pivot loop run by the oracle (CPU) or the CUDA engine (GPU) on the scaled instance - the reference's entering-arc
sequence, unscaled flows, objective and (scaled) duals."""

import gzip
import json
from pathlib import Path

import pytest

from helpers import rebuild_problem
from network_flow_solver_b200 import SolverOptions, _capi
from network_flow_solver_b200.scaling import compute_scaling_factors, should_scale_problem
from network_flow_solver_b200.solver import finish, prepare
from oracle import oracle

DOC = json.loads(gzip.open(Path(__file__).resolve().parent / "golden" / "next" / "scaling.json.gz", "rb").read().decode())
CASES = {c["name"]: c for c in DOC["cases"]}
RUNS = [(c["name"], i) for c in DOC["cases"] for i in range(len(c["runs"]))]


@pytest.mark.parametrize("name", sorted(CASES))
def test_scaling_trigger_and_factors_match_the_reference(name):
    case = CASES[name]
    problem = rebuild_problem(case["problem"])
    assert should_scale_problem(problem) == case["triggered"]
    f = compute_scaling_factors(problem)
    assert [f.cost_scale, f.capacity_scale, f.supply_scale, f.enabled] == case["factors"]  # bit-identical


def check(name, i, solve):
    case, run = CASES[name], CASES[name]["runs"][i]
    problem = rebuild_problem(case["problem"])
    cp, plan, options = prepare(problem, SolverOptions(**run["options"]), trace_capacity=1 << 16)
    assert (plan.scaling is not None) == case["triggered"]
    raw = solve(cp, plan.engine)
    assert raw.trace.tolist() == run["trace"]
    result = finish(cp, raw, options, plan.scaling)
    assert (result.status, result.iterations, result.objective) == (run["status"], run["iterations"], run["objective"])
    assert result.flows == {(a, b): v for a, b, v in run["flows"]}
    assert result.duals == dict((k, v) for k, v in run["duals"])


@pytest.mark.parametrize("name,i", RUNS)
def test_oracle_on_rescaled_instance_matches_reference(name, i):
    check(name, i, lambda cp, eng: oracle.solve_canonical(cp, eng))


@pytest.mark.gpu
@pytest.mark.parametrize("name,i", RUNS)
def test_engine_on_rescaled_instance_matches_reference(name, i):
    check(name, i, lambda cp, eng: _capi.solve_canonical(cp, eng))


@pytest.mark.gpu
def test_public_api_with_pure_defaults_on_a_badly_scaled_instance():
    from network_flow_solver_b200 import solve_min_cost_flow

    case = CASES["small_costs_big_caps"]
    run = case["runs"][0]
    result = solve_min_cost_flow(rebuild_problem(case["problem"]))  # SolverOptions() defaults: adaptive + auto_scale
    assert (result.status, result.iterations, result.objective) == (run["status"], run["iterations"], run["objective"])
