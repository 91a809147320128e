"""Public-API warm-start scenarios, the ones the reference's own suite walks through (tests/unit/test_warm_start.py:22-690:
None / identical / capacity up / foreign basis / empty basis / capacity cut / supply change / cost change / arc added /
arc removed / chains of re-solves / basis contents), written against `solve_min_cost_flow(..., warm_start_basis=...)` of
the drop-in.  Backend "oracle" runs the whole host path of the call on CPU with the pinned restatement standing in for
the C-ABI call (test-only substitution); backend "engine" is the real thing on the GPU."""

import pytest

from network_flow_solver_b200 import Basis, SolverOptions, _capi, build_problem, solve_min_cost_flow
from network_flow_solver_b200 import solver as solver_module
from oracle import oracle

BACKENDS = [
    "oracle",
    pytest.param("engine", marks=[pytest.mark.gpu, pytest.mark.timeout(300, method="thread")]),
]


@pytest.fixture(params=BACKENDS)
def solve(request, monkeypatch, capsys):
    if request.param == "oracle":
        monkeypatch.setattr(solver_module._capi, "solve_canonical",
                            lambda cp, opts, out=None, warm=None: oracle.solve_canonical(cp, opts, warm=warm))
    return solve_min_cost_flow


def problem(nodes, arcs):
    return build_problem(nodes=[{"id": k, "supply": float(v)} for k, v in nodes.items()],
                         arcs=[{"tail": a, "head": b, "capacity": None if cap is None else float(cap), "cost": float(c)}
                               for a, b, cap, c in arcs], directed=True, tolerance=1e-6)


DIAMOND = [("s", "a", 60, 2), ("s", "b", 60, 3), ("a", "t", 60, 2), ("b", "t", 60, 1), ("a", "b", 30, 0.5)]


def test_none_basis_is_a_cold_start(solve):
    r = solve(problem({"s": 10, "t": -10}, [("s", "t", 20, 1)]), warm_start_basis=None)
    assert (r.status, r.objective) == ("optimal", 10.0) and r.basis is not None


def test_identical_problem_resolves_at_once(solve):
    p = problem({"s": 40, "a": 0, "b": 0, "t": -40}, DIAMOND)
    first = solve(p)
    again = solve(p, warm_start_basis=first.basis)
    assert again.status == "optimal" and again.objective == first.objective and again.flows == first.flows
    assert again.iterations <= first.iterations


def test_capacity_increase(solve):
    nodes = {"s": 100, "m": 0, "t": -100}
    first = solve(problem(nodes, [("s", "m", 100, 1), ("m", "t", 100, 1)]))
    again = solve(problem(nodes, [("s", "m", 200, 1), ("m", "t", 200, 1)]), warm_start_basis=first.basis)
    assert (again.status, again.objective) == ("optimal", first.objective)


def test_basis_from_a_different_network_falls_back_to_cold_start(solve, caplog):
    first = solve(problem({"s": 10, "m": 0, "t": -10}, [("s", "m", 20, 1), ("m", "t", 20, 1)]))
    with caplog.at_level("WARNING"):
        other = solve(problem({"s": 10, "t": -10}, [("s", "t", 20, 1)]), warm_start_basis=first.basis)
    assert (other.status, other.objective) == ("optimal", 10.0)
    assert any("not in current problem" in m for m in caplog.messages)


def test_empty_basis_falls_back_to_cold_start(solve, caplog):
    with caplog.at_level("WARNING"):
        r = solve(problem({"s": 10, "t": -10}, [("s", "t", 20, 1)]), warm_start_basis=Basis(tree_arcs=set(), arc_flows={}))
    assert r.status == "optimal" and any("empty" in m.lower() for m in caplog.messages)


def test_capacity_cut_that_makes_the_instance_infeasible(solve):
    nodes = {"s": 100, "m": 0, "t": -100}
    first = solve(problem(nodes, [("s", "m", 100, 1), ("m", "t", 100, 1)]))
    assert first.status == "optimal"
    again = solve(problem(nodes, [("s", "m", 50, 1), ("m", "t", 100, 1)]), warm_start_basis=first.basis)
    assert again.status == "infeasible"


def test_supply_change(solve):
    arcs = [("s", "t", 100, 1)]
    first = solve(problem({"s": 50, "t": -50}, arcs))
    again = solve(problem({"s": 75, "t": -75}, arcs), warm_start_basis=first.basis)
    assert (first.objective, again.status, again.objective) == (50.0, "optimal", 75.0)


def test_cost_change_moves_the_flow(solve):
    nodes = {"s": 40, "a": 0, "b": 0, "t": -40}
    first = solve(problem(nodes, DIAMOND))
    dearer = [(a, b, cap, 9 if (a, b) == ("b", "t") else c) for a, b, cap, c in DIAMOND]
    again = solve(problem(nodes, dearer), warm_start_basis=first.basis)
    cold = solve(problem(nodes, dearer))
    assert again.status == "optimal" and again.objective == cold.objective != first.objective


def test_added_arc_can_only_help(solve):
    nodes = {"s": 40, "a": 0, "b": 0, "t": -40}
    first = solve(problem(nodes, DIAMOND))
    again = solve(problem(nodes, DIAMOND + [("s", "t", 15, 1)]), warm_start_basis=first.basis)
    assert again.status == "optimal" and again.objective < first.objective


def test_removed_arc(solve):
    nodes = {"s": 40, "a": 0, "b": 0, "t": -40}
    first = solve(problem(nodes, DIAMOND))
    fewer = [x for x in DIAMOND if (x[0], x[1]) != ("a", "b")]
    again = solve(problem(nodes, fewer), warm_start_basis=first.basis)  # accepted or rejected, the optimum is the cold one
    assert again.status == "optimal" and again.objective == solve(problem(nodes, fewer)).objective


def test_chain_of_resolves_with_growing_demand(solve):
    arcs = [("s", "m", None, 1), ("m", "t", None, 2), ("s", "t", None, 4)]
    basis, last = None, 0.0
    for demand in (10, 20, 35, 50):
        r = solve(problem({"s": demand, "m": 0, "t": -demand}, arcs), warm_start_basis=basis)
        assert r.status == "optimal" and r.objective == 3.0 * demand > last
        basis, last = r.basis, r.objective


def test_basis_contents(solve):
    r = solve(problem({"s": 40, "a": 0, "b": 0, "t": -40}, DIAMOND),
              SolverOptions(pricing_strategy="dantzig", explicit_pricing_strategy=True, auto_scale=False))
    assert isinstance(r.basis.tree_arcs, set) and isinstance(r.basis.arc_flows, dict)
    assert len(r.basis.tree_arcs) == 3 and set(r.basis.arc_flows) == r.basis.tree_arcs  # n - 1 arcs, no artificial ones left
    for key, f in r.basis.arc_flows.items():
        assert f == pytest.approx(r.flows.get(key, 0.0))


def test_progress_callback_is_invoked_once_with_the_totals(solve):
    """The reference calls back every `progress_interval` pivots (simplex.py:1143-1154); the device-resident loop has no
    host round trip per pivot, so the callback fires once, after the solve, with the totals - and says so in a warning."""
    p = problem({"s": 40, "a": 0, "b": 0, "t": -40}, DIAMOND)
    seen = []
    with pytest.warns(RuntimeWarning, match="progress_callback is invoked once"):
        res = solve(p, progress_callback=seen.append, progress_interval=1)
    assert len(seen) == 1
    info = seen[0]
    assert info.iteration == res.iterations and info.phase == 2 and info.objective_estimate == res.objective
    assert 0 <= info.phase_iterations <= info.iteration <= info.max_iterations and info.elapsed_time >= 0.0
