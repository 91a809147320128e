"""Scope row SURVEY.md section 8f-2: the reference's DEFAULT pricing - "adaptive", which in practice is its
candidate-list rule (top-100 arcs by |rc|, refreshed every 10 major iterations, 3 minor iterations per candidate
scan, reset with the Devex cadence).  The oracle restates it and is pinned here to runs recorded from the unmodified
reference (tests/golden/make_candidate_golden.py); the device driver is checked through the host emulation and the
CUDA engine (candidate scan in the pivot CTA, grid-wide top-100 refresh sweep) through the C ABI."""

import gzip
import json
from pathlib import Path

import pytest

from helpers import assert_matches_reference, rebuild_problem
from network_flow_solver_b200 import SolverOptions, _capi
from network_flow_solver_b200.solver import prepare
from oracle import oracle

DOC = json.loads(gzip.open(Path(__file__).resolve().parent / "golden" / "next" / "candidate_list.json.gz", "rb").read().decode())
RUNS = [(c["name"], i) for c in DOC["cases"] for i in range(len(c["runs"]))]
CASES = {c["name"]: c for c in DOC["cases"]}


@pytest.mark.parametrize("name,i", RUNS)
def test_oracle_reproduces_reference_candidate_list_runs(name, i):
    case, run = CASES[name], CASES[name]["runs"][i]
    problem = rebuild_problem(case["problem"])
    options = SolverOptions(**run["options"])
    cp, plan, options = prepare(problem, options, run.get("max_iterations"), trace_capacity=1 << 16,
)
    want = {"CandidateListPricing": _capi.PRICING_CANDIDATE_LIST, "AdaptivePricing": _capi.PRICING_CANDIDATE_LIST,
            "DantzigPricing": _capi.PRICING_DANTZIG}[run["strategy"]]
    assert plan.engine.pricing == want
    assert_matches_reference(run, cp, oracle.solve_canonical(cp, plan.engine), options)


@pytest.mark.parametrize("name,i", RUNS)
def test_emulated_device_core_reproduces_reference_candidate_list_runs(name, i):
    """The device driver / candidate scan / reset cadence (nsx_core.cuh compiled for the host, serial refresh)."""
    from emu import emu

    case, run = CASES[name], CASES[name]["runs"][i]
    problem = rebuild_problem(case["problem"])
    cp, plan, options = prepare(problem, SolverOptions(**run["options"]), run.get("max_iterations"),
                                trace_capacity=1 << 16)
    assert_matches_reference(run, cp, emu.solve_canonical(cp, plan.engine), options)


def test_adaptive_and_candidate_list_coincide_in_the_reference():
    for case in DOC["cases"]:
        by = {r["strategy"]: r for r in case["runs"] if r.get("max_iterations") is None}
        if "AdaptivePricing" in by and "CandidateListPricing" in by:
            assert by["AdaptivePricing"]["trace"] == by["CandidateListPricing"]["trace"], case["name"]


def test_default_options_resolve_to_the_candidate_list_rule():
    problem = rebuild_problem(DOC["cases"][0]["problem"])
    _, plan, _ = prepare(problem, SolverOptions(auto_scale=False))  # reference defaults = adaptive
    assert plan.engine.pricing == _capi.PRICING_CANDIDATE_LIST
    with pytest.raises(Exception):
        SolverOptions(pricing_strategy="steepest_edge")


@pytest.mark.gpu
def test_public_api_with_default_options_matches_the_reference():
    from network_flow_solver_b200 import solve_min_cost_flow

    for case in DOC["cases"]:
        run = next((r for r in case["runs"] if r["options"] == {"auto_scale": False}), None)
        if run is None:
            continue
        result = solve_min_cost_flow(rebuild_problem(case["problem"]), SolverOptions(auto_scale=False))
        assert (result.status, result.iterations, result.objective) == (run["status"], run["iterations"], run["objective"])
        assert result.flows == {(a, b): v for a, b, v in run["flows"]}


@pytest.mark.gpu
@pytest.mark.parametrize("name,i", RUNS)
def test_engine_reproduces_reference_candidate_list_runs(name, i):
    """CUDA engine (single-CTA path: these instances are small) against the recorded reference runs."""
    case, run = CASES[name], CASES[name]["runs"][i]
    problem = rebuild_problem(case["problem"])
    cp, plan, options = prepare(problem, SolverOptions(**run["options"]), run.get("max_iterations"),
                                trace_capacity=1 << 16)
    assert_matches_reference(run, cp, _capi.solve_canonical(cp, plan.engine), options)


@pytest.mark.gpu
@pytest.mark.parametrize("grid", [0, 5])
def test_engine_candidate_list_multi_cta_matches_oracle(grid, monkeypatch):
    """Grid-wide top-100 refresh (per-CTA buffers, bitonic compaction, merge in the pivot CTA) against the oracle."""
    import numpy as np
    from network_flow_solver_b200 import generators as gen
    from network_flow_solver_b200.canonical import initial_block_size

    if grid:
        monkeypatch.setenv("NSX_GRID", str(grid))
    cp = gen.netgen_like(1 << 13, 1 << 17, n_sources=32, n_sinks=32, cost_max=50, seed=77).canonical()
    m = cp.n_arcs
    opts = _capi.EngineOptions(pricing=_capi.PRICING_CANDIDATE_LIST, row_scan_first=False, block_size=initial_block_size(m),
                               auto_block=True, ft_update_limit=64, max_iterations=max(100, 20 * (m + cp.n_nodes - 1)),
                               tolerance=1e-6, trace_capacity=1 << 20)
    a, b = _capi.solve_canonical(cp, opts), oracle.solve_canonical(cp, opts, threads=4)
    assert (a.status, a.iterations) == (b.status, b.iterations)
    np.testing.assert_array_equal(a.trace, b.trace)
    np.testing.assert_array_equal(a.flow, b.flow)
    np.testing.assert_array_equal(a.potential, b.potential)


@pytest.mark.gpu
def test_batch_kernel_with_candidate_list_matches_oracle():
    import numpy as np
    from network_flow_solver_b200 import generators as gen
    from network_flow_solver_b200.canonical import initial_block_size

    cps = [gen.netgen_like(256, 2048, n_sources=4, n_sinks=4, seed=300 + k).canonical() for k in range(6)]
    m = cps[0].n_arcs
    opts = _capi.EngineOptions(pricing=_capi.PRICING_CANDIDATE_LIST, row_scan_first=False, block_size=initial_block_size(m),
                               auto_block=True, ft_update_limit=64, max_iterations=10**7, tolerance=1e-6,
                               trace_capacity=1 << 16)
    for cp, got in zip(cps, _capi.solve_batch_canonical(cps, opts)):
        want = oracle.solve_canonical(cp, opts)
        assert (got.status, got.iterations) == (want.status, want.iterations)
        np.testing.assert_array_equal(got.trace, want.trace)
        np.testing.assert_array_equal(got.flow, want.flow)
