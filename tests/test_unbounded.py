"""The unbounded exit of the ratio test (theta = +inf => UnboundedProblemError(entering_arc, reduced_cost),
simplex.py:1231-1246): instances with a negative-cost cycle of uncapacitated arcs, recorded from the unmodified reference
(tests/golden/make_unbounded_golden.py).  Pinned: every pivot before the failing one, the arc named in the exception and
the reduced cost reported (bit for bit - it depends on the phase costs and the potentials at that moment)."""

import gzip
import json
from pathlib import Path

import pytest

from emu import emu
from helpers import rebuild_problem
from network_flow_solver_b200 import SolverOptions, UnboundedProblemError, _capi, solve_min_cost_flow
from network_flow_solver_b200 import solver as solver_module
from network_flow_solver_b200.solver import finish, prepare
from oracle import oracle

DOC = json.loads(gzip.open(Path(__file__).resolve().parent / "golden" / "next" / "unbounded.json.gz", "rb").read().decode())
CASES = {c["name"]: c for c in DOC["cases"]}
RUNS = [(c["name"], i) for c in DOC["cases"] for i in range(len(c["runs"]))]


def check(name, i, solve):
    case, run = CASES[name], CASES[name]["runs"][i]
    assert run["status"] == "unbounded"
    cp, plan, options = prepare(rebuild_problem(case["problem"]), SolverOptions(**run["options"]), trace_capacity=1 << 16)
    raw = solve(cp, plan.engine)
    assert raw.status == _capi.STATUS_UNBOUNDED
    assert raw.trace.tolist() == run["trace"]  # the failing pivot is the last entry
    assert list(cp.arc_keys[raw.unbounded_arc]) == run["unbounded_arc"]
    assert raw.unbounded_rc == run["reduced_cost"]
    with pytest.raises(UnboundedProblemError) as err:
        finish(cp, raw, options)
    assert list(err.value.entering_arc) == run["unbounded_arc"] and err.value.reduced_cost == run["reduced_cost"]
    assert str(err.value) == run["message"]


@pytest.mark.parametrize("name,i", RUNS)
def test_oracle_matches_reference(name, i):
    check(name, i, lambda cp, eng: oracle.solve_canonical(cp, eng))


@pytest.mark.parametrize("name,i", RUNS)
def test_emulated_device_core_matches_reference(name, i):
    check(name, i, lambda cp, eng: emu.solve_canonical(cp, eng))


@pytest.mark.gpu
@pytest.mark.parametrize("name,i", RUNS)
def test_engine_matches_reference(name, i):
    check(name, i, lambda cp, eng: _capi.solve_canonical(cp, eng))


@pytest.mark.parametrize("backend", ["oracle", pytest.param("engine", marks=pytest.mark.gpu)])
def test_public_api_raises_like_the_reference(backend, monkeypatch):
    if backend == "oracle":  # test-only stand-in for the C-ABI call, to run the host half of the call on CPU
        monkeypatch.setattr(solver_module._capi, "solve_canonical",
                            lambda cp, opts, out=None, warm=None: oracle.solve_canonical(cp, opts, warm=warm))
    case = CASES["planted_cycle_48"]
    run = case["runs"][3]
    with pytest.raises(UnboundedProblemError) as err:
        solve_min_cost_flow(rebuild_problem(case["problem"]), SolverOptions(**run["options"]))
    assert list(err.value.entering_arc) == run["unbounded_arc"] and err.value.reduced_cost == run["reduced_cost"]
