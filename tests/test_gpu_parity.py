"""Parity tests proper: the CUDA engine, called through the C ABI, against (i) every recorded
reference run and (ii) the oracle on larger generated instances.  Bit-exact: entering-arc
sequence, per-arc flows, tree flags, node potentials, statuses, iteration counts."""

import numpy as np
import pytest

from helpers import assert_matches_reference, golden_cases, load_golden, prepare_run
from network_flow_solver_b200 import _capi
from network_flow_solver_b200 import generators as gen
from network_flow_solver_b200.canonical import initial_block_size
from oracle import oracle

pytestmark = pytest.mark.gpu


def engine_options(cp, pricing, **kw):
    m = cp.n_arcs
    base = dict(
        pricing=pricing,
        row_scan_first=cp.network_type == "transportation",
        block_size=initial_block_size(m),
        auto_block=True,
        ft_update_limit=64,
        max_iterations=max(100, 20 * (m + cp.n_nodes - 1)),
        tolerance=1e-6,
        trace_capacity=1 << 22,
    )
    base.update(kw)
    return _capi.EngineOptions(**base)


def assert_same_solution(a, b):
    assert a.status == b.status
    assert a.iterations == b.iterations
    if not np.array_equal(a.trace, b.trace):
        k = int(np.argmax(a.trace[: len(b.trace)] != b.trace[: len(a.trace)])) if len(a.trace) and len(b.trace) else 0
        raise AssertionError(f"entering sequence diverges at pivot {k}: {a.trace[k:k+3]} vs {b.trace[k:k+3]}")
    np.testing.assert_array_equal(a.flow, b.flow)
    np.testing.assert_array_equal(a.potential, b.potential)
    np.testing.assert_array_equal(a.state, b.state)
    assert a.degenerate_pivots == b.degenerate_pivots
    assert a.final_block_size == b.final_block_size
    assert a.tree_updates == b.tree_updates
    assert a.weight_resets == b.weight_resets
    assert a.phase1_iterations == b.phase1_iterations


@pytest.mark.parametrize("name,idx", golden_cases())
def test_engine_reproduces_reference(name, idx):
    doc = load_golden(name)
    run = doc["runs"][idx]
    _, cp, plan, options = prepare_run(doc, run)
    raw = _capi.solve_canonical(cp, plan.engine)
    assert_matches_reference(run, cp, raw, options)


CASES = [
    ("netgen", lambda: gen.netgen_like(2048, 16384, n_sources=8, n_sinks=8, seed=5), 0),
    ("netgen", lambda: gen.netgen_like(2048, 16384, n_sources=8, n_sinks=8, seed=5), 1),
    ("netgen_deep", lambda: gen.netgen_like(4096, 8192, n_sources=16, n_sinks=16, seed=6), 0),
    ("netgen_deep", lambda: gen.netgen_like(4096, 8192, n_sources=16, n_sinks=16, seed=6), 1),
    ("netgen_2^14", lambda: gen.netgen_like(1 << 14, 1 << 18, n_sources=64, n_sinks=64, seed=1601), 0),
    ("netgen_2^14", lambda: gen.netgen_like(1 << 14, 1 << 18, n_sources=64, n_sinks=64, seed=1601), 1),
    ("transport", lambda: gen.transportation(96, 128, cost_max=100, supply_each=64, seed=7), 0),
    ("transport", lambda: gen.transportation(96, 128, cost_max=100, supply_each=64, seed=7), 1),
    ("transport_512", lambda: gen.transportation(512, 512, cost_max=1000, seed=4096), 0),
    ("goto", lambda: gen.goto_like(32, seed=8), 0),
    ("goto_64", lambda: gen.goto_like(64, seed=0), 0),
    ("gridgen", lambda: gen.gridgen_like(32, 8200, seed=9), 1),
]


@pytest.mark.parametrize("family,make,pricing", CASES)
def test_engine_matches_oracle(family, make, pricing):
    cp = make().canonical()
    opts = engine_options(cp, pricing)
    assert_same_solution(_capi.solve_canonical(cp, opts), oracle.solve_canonical(cp, opts, threads=4))


@pytest.mark.parametrize("grid", [1, 2, 7, 148])
def test_engine_result_independent_of_grid_size(grid, monkeypatch):
    monkeypatch.setenv("NSX_GRID", str(grid))
    cp = gen.netgen_like(1024, 8192, n_sources=8, n_sinks=8, seed=15).canonical()
    for pricing in (0, 1):
        opts = engine_options(cp, pricing)
        assert_same_solution(_capi.solve_canonical(cp, opts), oracle.solve_canonical(cp, opts))


LAYOUT_CASES = [
    # (label, generator, eps_base, pricing, expected bytes per arc, NSX_LAYOUT)
    ("u16_i16", lambda: gen.transportation(96, 128, cost_max=100, supply_each=64, seed=7), 0.0, 0, 7, None),
    ("u16_i16_devex", lambda: gen.netgen_like(1024, 8192, n_sources=8, n_sinks=8, seed=15), 0.0, 1, 7, None),
    ("u16_i32", lambda: gen.netgen_like(1024, 8192, n_sources=8, n_sinks=8, cost_max=100000, seed=17), 0.0, 0, 9, None),
    ("u16_f64", lambda: gen.netgen_like(1024, 8192, n_sources=8, n_sinks=8, seed=15), 1e-10, 0, 13, None),
    ("i32_f64_forced", lambda: gen.netgen_like(1024, 8192, n_sources=8, n_sinks=8, seed=15), 1e-10, 1, 17, "wide"),
    ("i32_i32_forced", lambda: gen.transportation(300, 300, cost_max=1000, seed=3), 0.0, 0, 13, "i32"),
    ("multi_cta_u16_i16", lambda: gen.transportation(384, 512, cost_max=1000, seed=4), 0.0, 0, 7, None),
    # (n - 1 <= 8192 with int16 costs is stored pre-scaled - NSX_NODE_U16X8 / NSX_COST_I16M1; "plain16" forces the plain columns;
    # capacitated arcs keep the sweep on the path that reads the state bytes, the uncapacitated cases above take the state-free one)
    ("u16_i16_plain_forced", lambda: gen.transportation(96, 128, cost_max=100, supply_each=64, seed=7), 0.0, 0, 7, "plain16"),
    ("multi_cta_u16_i16_plain_forced", lambda: gen.transportation(384, 512, cost_max=1000, seed=4), 0.0, 0, 7, "plain16"),
    ("multi_cta_u16x8_capacitated", lambda: gen.transportation(256, 384, cost_max=1000, capacity=8.0, seed=5), 0.0, 0, 7, None),
    ("multi_cta_u16x8_devex", lambda: gen.transportation(384, 512, cost_max=1000, seed=6), 0.0, 1, 7, None),
    ("u16 above 8192 nodes", lambda: gen.netgen_like(9000, 1 << 16, n_sources=16, n_sinks=16, seed=23), 0.0, 0, 7, None),
]


@pytest.mark.parametrize("label,make,eps,pricing,bpa,force", LAYOUT_CASES)
def test_engine_packed_store_layouts(label, make, eps, pricing, bpa, force, monkeypatch):
    """Every encoding of the pricing store (uint16 / int32 node ids, int16 / int32 / float64 costs)
    gives the oracle's pivots bit for bit."""
    if force:
        monkeypatch.setenv("NSX_LAYOUT", force)
    cp = make().canonical(eps_base=eps)
    opts = engine_options(cp, pricing)
    got = _capi.solve_canonical(cp, opts)
    assert got.stats["bytes_per_arc"] == bpa
    assert_same_solution(got, oracle.solve_canonical(cp, opts, threads=4))


@pytest.mark.parametrize("stages", [2, 3])
def test_engine_shallow_tile_ring(stages, monkeypatch):
    monkeypatch.setenv("NSX_STAGES", str(stages))
    cp = gen.netgen_like(1 << 12, 1 << 17, n_sources=16, n_sinks=16, seed=21).canonical()
    for pricing in (0, 1):
        opts = engine_options(cp, pricing)
        assert_same_solution(_capi.solve_canonical(cp, opts), oracle.solve_canonical(cp, opts, threads=4))


def test_sweep_probe_prices_every_arc():
    import torch

    cp = gen.transportation(512, 512, cost_max=1000, seed=4096).canonical(eps_base=0.0)
    opts = engine_options(cp, 0)
    dev = [torch.from_numpy(getattr(cp, k)).cuda() for k in ("tail", "head", "pert_cost", "upper")]
    r = _capi.sweep_probe(cp, opts, [t.data_ptr() for t in dev], 10)
    assert r.status == 0 and r.arcs_priced == 10 * cp.n_arcs and r.stats["sweeps"] == 10


@pytest.mark.parametrize("par16", [0, 1])
def test_engine_blocked_preorder_array(par16, monkeypatch):
    """Trees that live in HBM (blocked preorder array, directory in shared memory) give the oracle's pivots."""
    monkeypatch.setenv("NSX_PAR16", str(par16))
    monkeypatch.setenv("NSX_GRID", "8")  # multi-CTA: the pivot CTA does not sweep and may mirror the parents
    monkeypatch.setenv("NSX_RESIDENT", "0")  # keep the tree out of shared memory
    cp = gen.netgen_like(4096, 32768, n_sources=16, n_sinks=16, seed=31).canonical()
    for pricing in (0, 1):
        opts = engine_options(cp, pricing)
        assert_same_solution(_capi.solve_canonical(cp, opts), oracle.solve_canonical(cp, opts, threads=4))


def test_engine_without_shared_memory_potentials(monkeypatch):
    monkeypatch.setenv("NSX_STAGE_PI", "0")
    cp = gen.netgen_like(1024, 8192, n_sources=8, n_sinks=8, seed=15).canonical()
    opts = engine_options(cp, 0)
    assert_same_solution(_capi.solve_canonical(cp, opts), oracle.solve_canonical(cp, opts))


def test_engine_fixed_small_blocks_and_short_reset_cadence():
    cp = gen.netgen_like(1024, 8192, n_sources=8, n_sinks=8, seed=15).canonical()
    for bs, ft in ((10, 64), (37, 3), (8192, 1), (100000, 64)):
        opts = engine_options(cp, 1, block_size=bs, auto_block=False, ft_update_limit=ft)
        assert_same_solution(_capi.solve_canonical(cp, opts), oracle.solve_canonical(cp, opts))


def test_engine_iteration_limits():
    cp = gen.netgen_like(512, 4096, n_sources=8, n_sinks=8, seed=16).canonical()
    full = oracle.solve_canonical(cp, engine_options(cp, 0))
    for limit in (1, 7, full.phase1_iterations, full.phase1_iterations + 1, full.iterations, full.iterations + 5):
        opts = engine_options(cp, 0, max_iterations=limit)
        assert_same_solution(_capi.solve_canonical(cp, opts), oracle.solve_canonical(cp, opts))


def test_batch_matches_single_solves():
    cps = [gen.goto_like(16, seed=100 + k).canonical() for k in range(40)]
    cps += [gen.netgen_like(256, 2048, n_sources=4, n_sinks=4, seed=200 + k).canonical() for k in range(8)]
    opts = engine_options(cps[0], 0, max_iterations=10**7)
    outs = _capi.solve_batch_canonical(cps, opts)
    for cp, out in zip(cps, outs):
        assert_same_solution(out, oracle.solve_canonical(cp, opts))


def test_public_api_matches_reference_fixture():
    from network_flow_solver_b200 import SolverOptions, solve_min_cost_flow
    from helpers import rebuild_problem

    doc = load_golden("gridgen_257")
    run = doc["runs"][0]
    result = solve_min_cost_flow(rebuild_problem(doc["problem"]), SolverOptions(**run["options"]))
    assert result.status == run["status"] and result.iterations == run["iterations"]
    assert result.objective == run["objective"]
    assert result.flows == {(a, b): v for a, b, v in run["flows"]}
    assert result.duals == dict((k, v) for k, v in run["duals"])
