"""Star pricing on the GPU (nsx_engine.cu: NSX_CMD_STAR / NSX_CMD_STAR_BUILD on the sweep workers, the row cache updated
with 128-bit compare-and-swap): a multi-CTA solve under the Dantzig rule must enter exactly the arcs the oracle's full
sweeps enter, bit for bit, and the solve with NSX_STAR=0 (full sweeps on the same grid) must agree too."""

import numpy as np
import pytest

from network_flow_solver_b200 import _capi
from network_flow_solver_b200 import generators as gen
from oracle import oracle
from test_gpu_parity import assert_same_solution, engine_options

pytestmark = pytest.mark.gpu

CASES = [
    # family, instance, pricing, eps = 0, extra environment
    ("netgen", lambda: gen.netgen_like(4096, 1 << 17, n_sources=32, n_sinks=32, seed=7), 0, False, {}),
    ("netgen_tree_in_hbm", lambda: gen.netgen_like(4096, 1 << 17, n_sources=32, n_sinks=32, seed=7), 0, False, {"NSX_RESIDENT": "0"}),
    ("netgen_caps", lambda: gen.netgen_like(1024, 1 << 16, n_sources=64, n_sinks=64, supply_each=3000, cap_max=50, seed=12), 0, False, {"NSX_GRID": "16"}),
    ("netgen_few_workers", lambda: gen.netgen_like(2048, 16384, n_sources=8, n_sinks=8, seed=5), 0, False, {"NSX_GRID": "3"}),
    ("transport", lambda: gen.transportation(320, 320, cost_max=100, seed=11), 0, True, {}),
    ("transport_ties", lambda: gen.transportation(256, 300, cost_max=3, seed=17), 0, True, {}),
    ("transport_then_devex", lambda: gen.transportation(200, 340, cost_max=50, seed=27), 1, True, {}),
    ("transport_perturbed", lambda: gen.transportation(96, 128, cost_max=100, supply_each=64, seed=7), 0, False, {"NSX_GRID": "8"}),
    ("gridgen", lambda: gen.gridgen_like(48, 18000, seed=9), 0, False, {"NSX_GRID": "12"}),
]


@pytest.mark.parametrize("family,make,pricing,eps0,env", CASES)
def test_star_pricing_enters_the_arcs_of_the_full_sweeps(family, make, pricing, eps0, env, monkeypatch):
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    cp = make().canonical(eps_base=0.0) if eps0 else make().canonical()
    opts = engine_options(cp, pricing)
    want = oracle.solve_canonical(cp, opts, threads=4)
    monkeypatch.setenv("NSX_STAR", "1")  # (the default turns it on for sparse instances only)
    got = _capi.solve_canonical(cp, opts)
    assert got.stats["grid"] > 1 and got.stats["star_pricing"] == 1  # (1: Dantzig rule, 2: Devex)
    assert got.stats["star_updates"] > 0 and got.stats["star_builds"] >= 1
    assert_same_solution(got, want)
    assert got.arcs_priced < 0.7 * got.iterations * cp.n_arcs  # far fewer arcs examined than full sweeps would
    monkeypatch.setenv("NSX_STAR", "0")
    full = _capi.solve_canonical(cp, opts)
    assert full.stats["star_pricing"] == 0 and full.stats["star_updates"] == 0
    assert_same_solution(full, want)


def test_star_pricing_iteration_limits_and_repeatability(monkeypatch):
    monkeypatch.setenv("NSX_STAR", "1")
    cp = gen.netgen_like(4096, 1 << 17, n_sources=32, n_sinks=32, seed=21).canonical()
    full = oracle.solve_canonical(cp, engine_options(cp, 0), threads=4)
    for limit in (1, 5, full.phase1_iterations, full.phase1_iterations + 1, full.iterations - 1):
        opts = engine_options(cp, 0, max_iterations=limit)
        assert_same_solution(_capi.solve_canonical(cp, opts), oracle.solve_canonical(cp, opts, threads=4))
    opts = engine_options(cp, 0)
    a, b = _capi.solve_canonical(cp, opts), _capi.solve_canonical(cp, opts)
    assert np.array_equal(a.trace, b.trace) and np.array_equal(a.flow, b.flow) and np.array_equal(a.potential, b.potential)


DEVEX_CASES = [
    ("netgen", lambda: gen.netgen_like(4096, 1 << 17, n_sources=32, n_sinks=32, seed=7), False, {}),
    ("netgen_tree_in_hbm", lambda: gen.netgen_like(4096, 1 << 17, n_sources=32, n_sinks=32, seed=7), False, {"NSX_RESIDENT": "0"}),
    ("netgen_caps", lambda: gen.netgen_like(1024, 1 << 16, n_sources=64, n_sinks=64, supply_each=3000, cap_max=50, seed=12), False, {"NSX_GRID": "16"}),
    ("netgen_caps_ties", lambda: gen.netgen_like(512, 1 << 16, n_sources=32, n_sinks=32, supply_each=2000, cap_max=20, cost_max=4, seed=13), True, {"NSX_GRID": "9"}),
    ("gridgen", lambda: gen.gridgen_like(48, 18000, seed=9), False, {"NSX_GRID": "12"}),
]


@pytest.mark.parametrize("blocks", ["one_block", "adaptive", "short_cadence"])
@pytest.mark.parametrize("family,make,eps0,env", DEVEX_CASES)
def test_star_pricing_under_devex_with_a_single_block(family, make, eps0, env, blocks, monkeypatch):
    """Devex block pricing while the block covers all arcs: forward / backward row caches, the last degenerate arc left out
    and put back, caches rebuilt at every weight reset - the oracle's pivots, bit for bit."""
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    monkeypatch.setenv("NSX_STAR", "1")
    cp = make().canonical(eps_base=0.0) if eps0 else make().canonical()
    kw = {"one_block": dict(block_size=cp.n_arcs, auto_block=False), "adaptive": {},
          "short_cadence": dict(block_size=cp.n_arcs, auto_block=False, ft_update_limit=3)}[blocks]
    opts = engine_options(cp, 1, **kw)
    want = oracle.solve_canonical(cp, opts, threads=4)
    got = _capi.solve_canonical(cp, opts)
    assert got.stats["grid"] > 1 and got.stats["star_pricing"] == 2
    if blocks != "adaptive":
        assert got.stats["star_updates"] > 0 and got.stats["star_builds"] >= 1
    assert_same_solution(got, want)
