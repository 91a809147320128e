"""CPU check of the device pivot code: nsx_core.cuh compiled as a serial host emulation
(tests/emu) must reproduce every recorded reference run bit for bit, and agree with the oracle
on larger generated instances of each family."""

import numpy as np
import pytest

from helpers import assert_matches_reference, golden_cases, load_golden, prepare_run
from network_flow_solver_b200 import _capi
from network_flow_solver_b200 import generators as gen
from network_flow_solver_b200.canonical import initial_block_size
from oracle import oracle
from emu import emu


@pytest.mark.parametrize("name,idx", golden_cases())
def test_emulated_core_reproduces_reference(name, idx):
    doc = load_golden(name)
    run = doc["runs"][idx]
    _, cp, plan, options = prepare_run(doc, run)
    raw = emu.solve_canonical(cp, plan.engine)
    assert_matches_reference(run, cp, raw, options)


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


def assert_same_solution(a: _capi.RawSolution, b: _capi.RawSolution):
    assert a.status == b.status
    assert a.iterations == b.iterations
    np.testing.assert_array_equal(a.trace, b.trace)
    np.testing.assert_array_equal(a.flow, b.flow)
    np.testing.assert_array_equal(a.potential, b.potential)
    np.testing.assert_array_equal(a.state, b.state)
    assert a.degenerate_pivots == b.degenerate_pivots
    assert a.final_block_size == b.final_block_size
    assert a.tree_updates == b.tree_updates
    assert a.weight_resets == b.weight_resets


CASES = [
    ("netgen", lambda: gen.netgen_like(2048, 16384, n_sources=8, n_sinks=8, seed=5), 0),
    ("netgen", lambda: gen.netgen_like(2048, 16384, n_sources=8, n_sinks=8, seed=5), 1),
    ("netgen_deep", lambda: gen.netgen_like(4096, 8192, n_sources=16, n_sinks=16, seed=6), 0),
    ("netgen_deep", lambda: gen.netgen_like(4096, 8192, n_sources=16, n_sinks=16, seed=6), 1),
    ("transport", lambda: gen.transportation(96, 128, cost_max=100, supply_each=64, seed=7), 0),
    ("transport", lambda: gen.transportation(96, 128, cost_max=100, supply_each=64, seed=7), 1),
    ("goto", lambda: gen.goto_like(32, seed=8), 0),
    ("goto", lambda: gen.goto_like(32, seed=8), 1),
    ("gridgen", lambda: gen.gridgen_like(32, 8200, seed=9), 1),
]


@pytest.mark.parametrize("family,make,pricing", CASES)
def test_emulated_core_matches_oracle(family, make, pricing):
    cp = make().canonical()
    opts = engine_options(cp, pricing)
    assert_same_solution(emu.solve_canonical(cp, opts), oracle.solve_canonical(cp, opts))


def test_emulated_core_fixed_small_blocks_and_short_reset_cadence():
    cp = gen.netgen_like(1024, 8192, n_sources=8, n_sinks=8, seed=15).canonical()
    for bs, ft in ((10, 64), (37, 3), (8192, 1), (100000, 64)):
        opts = engine_options(cp, 1, block_size=bs, auto_block=False, ft_update_limit=ft)
        assert_same_solution(emu.solve_canonical(cp, opts), oracle.solve_canonical(cp, opts))


def test_emulated_core_iteration_limits():
    cp = gen.netgen_like(512, 4096, n_sources=8, n_sinks=8, seed=16).canonical()
    full = oracle.solve_canonical(cp, engine_options(cp, 0))
    for limit in (1, 7, full.phase1_iterations, full.phase1_iterations + 1, full.iterations, full.iterations + 5):
        opts = engine_options(cp, 0, max_iterations=limit)
        assert_same_solution(emu.solve_canonical(cp, opts), oracle.solve_canonical(cp, opts))


@pytest.mark.parametrize("lg,spare", [("", ""), ("1", "6"), ("2", "4"), ("3", "40"), ("5", "4")])
@pytest.mark.parametrize("family,make,pricing", CASES[:6] + CASES[-2:])
def test_emulated_core_blocked_preorder(family, make, pricing, lg, spare, monkeypatch):
    """Blocked preorder array (+ depth-synchronised cycle walk), the variant the engine uses for trees that live
    in HBM, gives the same pivots as the oracle - with the production block size and with blocks of 2 / 4 / 8 / 32
    slots and few spare blocks, where every pivot splits, empties or merges blocks and the array is laid out afresh
    again and again."""
    monkeypatch.setenv("NSX_EMU_BLOCKED", "1")
    monkeypatch.setenv("NSX_EMU_BLK_LG", lg)
    monkeypatch.setenv("NSX_EMU_BLK_NB", spare)
    cp = make().canonical()
    opts = engine_options(cp, pricing)
    got = emu.solve_canonical(cp, opts)
    assert_same_solution(got, oracle.solve_canonical(cp, opts))
    if lg in ("1", "2") and got.tree_updates > 50:
        assert got.timing["pivot_ms"] > 0  # (emulation: number of re-layouts) the rebuild path ran


STAR_CASES = [
    ("netgen", lambda: gen.netgen_like(2048, 16384, n_sources=8, n_sinks=8, seed=5), 0, 0),
    ("netgen_deep", lambda: gen.netgen_like(4096, 8192, n_sources=16, n_sinks=16, seed=6), 0, 0),
    ("netgen_caps", lambda: gen.netgen_like(512, 8192, n_sources=32, n_sinks=32, supply_each=3000, cap_max=50, seed=12), 0, 0),
    ("transport", lambda: gen.transportation(96, 128, cost_max=100, supply_each=64, seed=7), 0, 1),
    ("transport_ties", lambda: gen.transportation(64, 64, cost_max=3, seed=17), 0, 1),
    ("transport_devex_fallthrough", lambda: gen.transportation(48, 80, cost_max=50, seed=27), 1, 1),
    ("goto", lambda: gen.goto_like(32, seed=8), 0, 0),
    ("gridgen", lambda: gen.gridgen_like(32, 8200, seed=9), 0, 0),
]


@pytest.mark.parametrize("blocked", ["0", "1"])
@pytest.mark.parametrize("family,make,pricing,eps0", STAR_CASES)
def test_emulated_core_star_pricing(family, make, pricing, eps0, blocked, monkeypatch):
    """Star pricing (row cache kept up to date from the re-hung subtree of each pivot instead of sweeping all arcs)
    enters exactly the arcs the full Dantzig sweep enters - with and without cost perturbation (ties), with capacities
    (backward candidates, bound flips), through Phase 1 zero-candidate passes and the phase switch."""
    monkeypatch.setenv("NSX_EMU_STAR", "1")
    monkeypatch.setenv("NSX_EMU_BLOCKED", blocked)
    cp = make().canonical(eps_base=0.0) if eps0 else make().canonical()
    opts = engine_options(cp, pricing)
    got = emu.solve_canonical(cp, opts)
    assert_same_solution(got, oracle.solve_canonical(cp, opts))
    assert got.timing["sync_ms"] > 0  # (emulation: number of star updates) the incremental path ran
    assert got.arcs_priced < 0.7 * got.iterations * cp.n_arcs


DEVEX_STAR_CASES = [
    ("netgen", lambda: gen.netgen_like(2048, 16384, n_sources=8, n_sinks=8, seed=5), False),
    ("netgen_deep", lambda: gen.netgen_like(4096, 8192, n_sources=16, n_sinks=16, seed=6), False),
    ("netgen_caps", lambda: gen.netgen_like(512, 8192, n_sources=32, n_sinks=32, supply_each=3000, cap_max=50, seed=12), False),
    ("netgen_caps_ties", lambda: gen.netgen_like(256, 4096, n_sources=32, n_sinks=32, supply_each=2000, cap_max=20, cost_max=4, seed=13), True),
    ("goto", lambda: gen.goto_like(32, seed=8), False),
    ("gridgen", lambda: gen.gridgen_like(32, 8200, seed=9), False),
]


@pytest.mark.parametrize("blocks", ["one_block", "adaptive", "short_cadence"])
@pytest.mark.parametrize("family,make,eps0", DEVEX_STAR_CASES)
def test_emulated_core_star_pricing_devex(family, make, eps0, blocks, monkeypatch):
    """Devex block pricing while its block covers all arcs (what the block adaptation grows it to on degenerate runs) is an
    arg-max over everything: kept in a forward and a backward row cache, with the last degenerate arc left out and put back,
    the caches rebuilt at every weight reset.  Same pivots as the oracle's block sweeps."""
    monkeypatch.setenv("NSX_EMU_STAR", "1")
    cp = make().canonical(eps_base=0.0) if eps0 else make().canonical()
    kw = {"one_block": dict(block_size=cp.n_arcs, auto_block=False), "adaptive": {},
          "short_cadence": dict(block_size=cp.n_arcs, auto_block=False, ft_update_limit=3)}[blocks]
    opts = engine_options(cp, 1, **kw)
    got = emu.solve_canonical(cp, opts)
    assert_same_solution(got, oracle.solve_canonical(cp, opts))
    if blocks != "adaptive":
        assert got.timing["sync_ms"] > 0 and got.timing["exchange_ms"] >= 1  # (emulation: star updates / rebuilds)


@pytest.mark.parametrize("blocked", ["0", "1"])
@pytest.mark.parametrize("family,make,pricing", CASES[:6])
def test_emulated_core_deferred_bookkeeping_can_wait_for_the_next_pivot(family, make, pricing, blocked, monkeypatch):
    """The part of a tree update that pricing does not need (closing the gap the re-hung subtree left in the preorder array)
    is deferred: the engine's pivot CTA does it while the sweep workers price.  Here it is left until the next pivot needs the
    array - the latest possible moment - and the pivots must not change."""
    monkeypatch.setenv("NSX_EMU_DEFER", "late")
    monkeypatch.setenv("NSX_EMU_BLOCKED", blocked)
    cp = make().canonical()
    opts = engine_options(cp, pricing)
    assert_same_solution(emu.solve_canonical(cp, opts), oracle.solve_canonical(cp, opts))
