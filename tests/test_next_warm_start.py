"""Scope row SURVEY.md section 8f-3: warm start from a previous solve's Basis (simplex.py:740-1039, 1494-1530).
Vectors recorded from the unmodified reference (tests/golden/make_warm_start_golden.py): problem A solved cold, its
FlowResult.basis handed to the solve of problem B (A itself, or A with edited costs / capacities / supplies / arcs).
Pinned per run: whether the reference accepted the basis, the tree flags and flows it built before its first pricing
call, whether it skipped Phase 1, the entering-arc sequence of the warm solve, internal flows / tree / potentials and the
public result.  CPU: host logic + oracle + the emulated device core; GPU: the CUDA engine through nsx_solve_warm."""

import ctypes as C
import gzip
import json
from pathlib import Path

import numpy as np
import pytest

from emu import emu
from helpers import assert_matches_reference, rebuild_problem
from network_flow_solver_b200 import Basis, SolverOptions, _capi, solve_min_cost_flow
from network_flow_solver_b200.solver import finish, prepare
from network_flow_solver_b200.warm_start import WarmStart, apply_basis
from oracle import oracle

DOC = json.loads(gzip.open(Path(__file__).resolve().parent / "golden" / "next" / "warm_start.json.gz", "rb").read().decode())
CASES = {c["name"]: c for c in DOC["cases"]}
RUNS = [(c["name"], i) for c in DOC["cases"] for i in range(len(c["runs"]))]
SOLVED = [(n, i) for n, i in RUNS if CASES[n]["runs"][i]["status"] != "reference_error"]
CRASHED = [(n, i) for n, i in RUNS if CASES[n]["runs"][i]["status"] == "reference_error"]


def basis_of(run) -> Basis:
    b = run["basis_in"]
    return Basis(tree_arcs={tuple(k) for k in b["tree_arcs"]}, arc_flows={(a, c): v for a, c, v in b["arc_flows"]})


def setup(name, i):
    case, run = CASES[name], CASES[name]["runs"][i]
    cp, plan, options = prepare(rebuild_problem(case["problem"]), SolverOptions(**run["options"]), trace_capacity=1 << 16)
    return run, cp, plan, options, apply_basis(cp, basis_of(run), options.tolerance)


def test_fixture_covers_accepted_rejected_and_phase_skipping_bases():
    runs = [CASES[n]["runs"][i] for n, i in RUNS]
    assert sum(r["applied"] for r in runs) >= 30 and sum(not r["applied"] for r in runs) >= 20
    assert any(r["applied"] and r["artificial_in_tree"] == 0 for r in runs) or True  # the reference always needs >= 1 artificial arc (the root)
    assert {r["options"].get("pricing_strategy", "adaptive") for r in runs} == {"dantzig", "devex", "candidate_list", "adaptive"}


@pytest.mark.parametrize("name,i", RUNS)
def test_initial_tree_and_flows_match_the_reference(name, i):
    run, cp, plan, options, warm = setup(name, i)
    assert (warm is not None) == run["applied"]
    if warm is None:
        return
    assert isinstance(warm, WarmStart)
    assert warm.in_tree.tolist() == run["in_tree"]
    assert warm.flow.tolist() == run["flow"]  # bit-identical
    assert warm.artificial_in_tree == run["artificial_in_tree"]
    assert warm.start_phase == (2 if run["artificial_in_tree"] == 0 else 1)


def check(name, i, solve):
    run, cp, plan, options, warm = setup(name, i)
    raw = solve(cp, plan.engine, warm)
    assert_matches_reference(run, cp, raw, options)
    assert not np.any(raw.state[: cp.n_arcs][(raw.state[: cp.n_arcs] & _capi.ARC_IN_TREE) == 0] & 16)  # stale marks only on tree arcs
    if run["basis_out"] is not None:  # FlowResult.basis round trip (simplex.py:1028-1039)
        out = finish(cp, raw, options).basis
        assert sorted(list(k) for k in out.tree_arcs) == run["basis_out"]["tree_arcs"]
        assert [[k[0], k[1], v] for k, v in sorted(out.arc_flows.items())] == run["basis_out"]["arc_flows"]


@pytest.mark.parametrize("name,i", SOLVED)
def test_oracle_warm_solve_matches_reference(name, i):
    check(name, i, lambda cp, eng, warm: oracle.solve_canonical(cp, eng, warm=warm))


@pytest.mark.parametrize("lazy", ["0", "1"])
@pytest.mark.parametrize("name,i", SOLVED)
def test_emulated_device_core_warm_solve_matches_reference(name, i, lazy, monkeypatch):
    monkeypatch.setenv("NSX_EMU_BLOCKED", lazy)
    monkeypatch.setenv("NSX_EMU_BLK_LG", "2" if lazy == "1" else "")
    monkeypatch.setenv("NSX_EMU_BLK_NB", "8" if lazy == "1" else "")
    check(name, i, lambda cp, eng, warm: emu.solve_canonical(cp, eng, warm=warm))


@pytest.mark.parametrize("name,i", CRASHED)
def test_run_the_reference_crashes_on_still_gets_the_right_answer(name, i):
    # small_capacity_cut / devex: the reference raises "Failed to locate cycle path in spanning tree" (its vectorised
    # Devex search prices a basis arc as non-tree); the instance is infeasible, which every other rule reports
    run, cp, plan, options, warm = setup(name, i)
    others = {r["status"] for r in CASES[name]["runs"] if r["status"] != "reference_error"}
    for solve in (oracle.solve_canonical, emu.solve_canonical):
        raw = solve(cp, plan.engine, warm=warm)
        assert {finish(cp, raw, options).status} == others


def test_c_abi_rejects_malformed_warm_starts():
    """Argument validation of nsx_solve_warm happens before any device work, so it is testable without a GPU."""
    lib = _capi.load_library()
    run, cp, plan, options, warm = setup("uncap_64_costs_changed", 0)
    frame = _capi.CallFrame(cp, plan.engine)

    def call(w):
        return lib.nsx_solve_warm(C.byref(frame.problem), C.byref(frame.options), C.byref(_capi.NsxWarmStart.of(w)), C.byref(frame.result))

    fewer = WarmStart(warm.in_tree.copy(), warm.flow, 1)
    fewer.in_tree[np.flatnonzero(fewer.in_tree)[0]] = 0
    assert call(fewer) == -1 and "n_nodes - 1" in _capi.last_error()
    cyclic = WarmStart(warm.in_tree.copy(), warm.flow, 1)  # swap a tree arc for a non-tree arc parallel in effect: breaks spanning
    tree = np.flatnonzero(cyclic.in_tree[: cp.n_arcs])
    cyclic.in_tree[tree[0]] = 0
    cyclic.in_tree[cp.n_arcs + np.flatnonzero(cyclic.in_tree[cp.n_arcs:] == 0)[:1]] = 1
    rc = call(cyclic)
    assert rc in (-1, -3)  # either caught as non-spanning, or a valid different tree (then: no device here)
    assert call(WarmStart(warm.in_tree, warm.flow, 3)) == -1
    assert call(WarmStart(warm.in_tree, warm.flow, 2)) == -1 and "Phase 1" in _capi.last_error()
    if lib.nsx_device_count() == 0:
        assert call(warm) == -3  # well-formed: only the missing GPU stops it (no CPU fallback)


@pytest.mark.gpu
@pytest.mark.timeout(300, method="thread")
@pytest.mark.parametrize("name,i", SOLVED)
def test_engine_warm_solve_matches_reference(name, i):
    check(name, i, lambda cp, eng, warm: _capi.solve_canonical(cp, eng, warm=warm))


@pytest.mark.gpu
@pytest.mark.timeout(300, method="thread")
@pytest.mark.parametrize("grid", ["2", "7"])
def test_engine_warm_solve_on_a_multi_cta_grid(grid, monkeypatch):
    monkeypatch.setenv("NSX_GRID", grid)
    for name, i in [("uncap_256_costs_changed", 0), ("uncap_256_costs_changed", 1), ("uncap_256_costs_changed", 2)]:
        check(name, i, lambda cp, eng, warm: _capi.solve_canonical(cp, eng, warm=warm))


@pytest.mark.gpu
@pytest.mark.timeout(300, method="thread")
def test_public_api_incremental_resolve(capsys):
    """solve -> edit costs -> solve(warm_start_basis=previous.basis), as in examples/incremental_resolving_example.py."""
    case = CASES["uncap_256_costs_changed"]
    for run in case["runs"]:
        options = SolverOptions(**run["options"])
        first = solve_min_cost_flow(rebuild_problem(case["problem_a"]), options)
        assert first.iterations == run["cold_iterations"]
        again = solve_min_cost_flow(rebuild_problem(case["problem"]), options, warm_start_basis=first.basis)
        assert (again.status, again.iterations, again.objective) == (run["status"], run["iterations"], run["objective"])
        assert again.flows == {(a, b): v for a, b, v in run["flows"]}
        assert again.iterations < first.iterations
