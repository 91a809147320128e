"""Canonicalisation row SURVEY.md 8/a9 beyond plain directed instances: arcs with lower bounds (shifted out before the
solve, simplex.py:392-432; added back in flows and objective, simplex.py:1703-1721) and undirected problems (an edge is
one arc with lower = -capacity, data.py:162-223, so a flow against the stored orientation comes back NEGATIVE and counts
negatively in the objective - the reference's convention, pinned here).  Vectors recorded from the unmodified reference
(tests/golden/make_bounds_golden.py), among them the scenarios of its tests/unit/test_undirected_graphs.py and the
150-node chain of tests/integration/test_undirected_performance.py (objective 534600)."""

import gzip
import json
from pathlib import Path

import pytest

from emu import emu
from helpers import assert_matches_reference, rebuild_problem
from network_flow_solver_b200 import SolverOptions, _capi, solve_min_cost_flow
from network_flow_solver_b200.solver import prepare
from oracle import oracle

DOC = json.loads(gzip.open(Path(__file__).resolve().parent / "golden" / "next" / "bounds_and_undirected.json.gz", "rb").read().decode())
CASES = {c["name"]: c for c in DOC["cases"]}
RUNS = [(c["name"], i) for c in DOC["cases"] for i in range(len(c["runs"]))]


def setup(name, i):
    case, run = CASES[name], CASES[name]["runs"][i]
    cp, plan, options = prepare(rebuild_problem(case["problem"]), SolverOptions(**run["options"]), trace_capacity=1 << 16)
    return case, run, cp, plan, options


def test_fixture_has_shifted_arcs_negative_flows_and_the_known_objectives():
    assert any(any(a[4] > 0 for a in c["problem"]["arcs"]) for c in CASES.values())
    assert sum(not c["problem"]["directed"] for c in CASES.values()) >= 6
    assert CASES["undirected_chain_3"]["runs"][0]["objective"] == 50.0
    assert CASES["undirected_chain_150"]["runs"][0]["objective"] == 534600.0
    assert any(v < 0 for _, _, v in CASES["undirected_star_mixed"]["runs"][0]["flows"])


@pytest.mark.parametrize("name", sorted(CASES))
def test_canonical_arrays_carry_the_shift(name):
    case, run, cp, plan, options = setup(name, 0)
    spec = case["problem"]
    if spec["directed"]:
        assert sorted(cp.shift.tolist()) == sorted(a[4] for a in spec["arcs"])
    else:  # lower = -capacity, residual capacity 2 * capacity
        assert sorted(cp.shift.tolist()) == sorted(-a[2] for a in spec["arcs"])
        assert sorted(cp.upper.tolist()) == sorted(2 * a[2] for a in spec["arcs"])
    assert abs(float(cp.supply.sum())) < 1e-9


@pytest.mark.parametrize("name,i", RUNS)
def test_oracle_matches_reference(name, i):
    case, run, cp, plan, options = setup(name, i)
    assert_matches_reference(run, cp, oracle.solve_canonical(cp, plan.engine), options)


@pytest.mark.parametrize("name,i", RUNS)
def test_emulated_device_core_matches_reference(name, i):
    case, run, cp, plan, options = setup(name, i)
    assert_matches_reference(run, cp, emu.solve_canonical(cp, plan.engine), options)


@pytest.mark.gpu
@pytest.mark.parametrize("name,i", RUNS)
def test_engine_matches_reference(name, i):
    case, run, cp, plan, options = setup(name, i)
    assert_matches_reference(run, cp, _capi.solve_canonical(cp, plan.engine), options)


@pytest.mark.gpu
def test_public_api_on_an_undirected_problem(capsys):
    case = CASES["undirected_chain_150"]
    run = case["runs"][1]
    result = solve_min_cost_flow(rebuild_problem(case["problem"]), SolverOptions(**run["options"]))
    assert (result.status, result.iterations, result.objective) == (run["status"], run["iterations"], run["objective"])
    assert result.flows == {(a, b): v for a, b, v in run["flows"]}
