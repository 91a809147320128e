"""SolverOptions.use_vectorized_pricing=False: the reference's loop-based Devex block scan (simplex_pricing.py:205-269).
Not a SURVEY.md section 8 row by itself (8/a3 is the vectorised search) but part of the option surface of the drop-in call;
it differs from the vectorised search in the cost it prices with (tree cost of the phase), the comparison (a later arc
needs a merit larger by more than the tolerance), the zero-candidate pick, the block pointer after a zero pick and the
missing exclusion of the last degenerate arc.  Vectors recorded from the unmodified reference
(tests/golden/make_devex_loop_golden.py), including the runs where the reference itself ends "infeasible"."""

import gzip
import json
from pathlib import Path

import pytest

from emu import emu
from helpers import assert_matches_reference, rebuild_problem
from network_flow_solver_b200 import SolverOptions, _capi
from network_flow_solver_b200.solver import prepare
from oracle import oracle

DOC = json.loads(gzip.open(Path(__file__).resolve().parent / "golden" / "next" / "devex_loop.json.gz", "rb").read().decode())
CASES = {c["name"]: c for c in DOC["cases"]}
RUNS = [(c["name"], i) for c in DOC["cases"] for i in range(len(c["runs"]))]


def setup(name, i):
    case, run = CASES[name], CASES[name]["runs"][i]
    cp, plan, options = prepare(rebuild_problem(case["problem"]), SolverOptions(**run["options"]), run.get("max_iterations"),
                                trace_capacity=1 << 16)
    assert plan.engine.pricing == _capi.PRICING_DEVEX_LOOP
    return run, cp, plan, options


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
