"""The section-8f rows at sizes beyond what the pure-Python reference records in seconds: oracle (CPU restatement, pinned to
the reference by the golden tests of each row) against the device pivot source - emulated on the host here, the CUDA
engine on the GPU.  Entering-arc sequence, flows, potentials, arc states and counters must be identical.
Covers what the small recordings cannot: structure-rule / loop-Devex scans spanning many 1024-arc chunks, warm starts
whose initial tree is thousands of nodes deep, lazy preorder positions on a warm tree."""

import numpy as np
import pytest

from emu import emu
from network_flow_solver_b200 import _capi
from network_flow_solver_b200 import generators as gen
from network_flow_solver_b200.canonical import initial_block_size
from network_flow_solver_b200.solver import special_rule
from network_flow_solver_b200.warm_start import apply_tree_arcs
from oracle import oracle
from test_emulated_device_core import assert_same_solution


def options_for(cp, pricing, **kw):
    rule, mask = special_rule(cp, 1e-6)
    base = dict(pricing=pricing, row_scan_first=rule, node_mask=mask, block_size=initial_block_size(cp.n_arcs), auto_block=True,
                ft_update_limit=64, max_iterations=max(100, 20 * (cp.n_arcs + cp.n_nodes - 1)), tolerance=1e-6,
                trace_capacity=1 << 22)
    base.update(kw)
    return _capi.EngineOptions(**base)


SCANS = [
    ("assignment_64", lambda: gen.assignment(64, seed=1), _capi.PRICING_DANTZIG),
    ("assignment_64", lambda: gen.assignment(64, seed=1), _capi.PRICING_CANDIDATE_LIST),
    ("assignment_48_ties", lambda: gen.assignment(48, cost_max=3, seed=2), _capi.PRICING_DEVEX),
    ("shortest_path_2048", lambda: gen.shortest_path(2048, 12000, seed=3), _capi.PRICING_DANTZIG),
    ("shortest_path_2048", lambda: gen.shortest_path(2048, 12000, seed=3), _capi.PRICING_DEVEX_LOOP),
    ("max_flow_1024", lambda: gen.max_flow(1024, 6000, seed=4), _capi.PRICING_DANTZIG),
    ("loop_devex_netgen_2048", lambda: gen.netgen_like(2048, 16384, n_sources=8, n_sinks=8, seed=5), _capi.PRICING_DEVEX_LOOP),
    ("loop_devex_transport", lambda: gen.transportation(64, 96, cost_max=100, supply_each=48, seed=7), _capi.PRICING_DEVEX_LOOP),
]


def run_scan_case(make, pricing, solve):
    cp = make().canonical()
    opts = options_for(cp, pricing)
    ref = oracle.solve_canonical(cp, opts)
    assert ref.iterations > 50
    assert_same_solution(solve(cp, opts), ref)
    return cp, opts, ref


@pytest.mark.parametrize("name,make,pricing", SCANS)
def test_emulated_core_agrees_with_oracle_on_scan_rules(name, make, pricing):
    cp, opts, ref = run_scan_case(make, pricing, emu.solve_canonical)
    if name.startswith(("assignment", "shortest", "max_flow")):
        assert opts.row_scan_first >= _capi.SPECIAL_ASSIGNMENT and cp.n_arcs > 2048  # several chunks per scan
    if name.startswith("shortest"):
        assert 0 < int(opts.node_mask[1:].sum()) < cp.n_nodes - 1


def warm_case(n, m, seed, pricing, drop_every):
    """Uncapacitated NETGEN instance solved cold; some costs change; the old tree (minus a few idle arcs, so that several
    components hang off the root by artificial arcs) warm-starts the new solve."""
    a = gen.netgen_like(n, m, n_sources=8, n_sinks=8, cost_max=1000, seed=seed)
    a.capacity[:] = np.inf
    cp0 = a.canonical()
    first = oracle.solve_canonical(cp0, options_for(cp0, pricing))
    assert first.status == _capi.STATUS_OPTIMAL
    rng = np.random.default_rng(seed + 100)
    a.cost = a.cost.copy()
    changed = rng.choice(a.n_arcs, size=a.n_arcs // 20, replace=False)
    a.cost[changed] = rng.integers(1, 1001, size=changed.size).astype(np.float64)
    cp = a.canonical()
    tree = np.flatnonzero(first.state[: cp.n_arcs] & _capi.ARC_IN_TREE)
    if drop_every:  # only arcs without flow can go: a split-off component with net supply would overload the artificial
        idle = np.flatnonzero(first.flow[tree] == 0.0)  # arc of a single node and the reference rejects such a basis
        tree = np.delete(tree, idle[::drop_every])
    warm = apply_tree_arcs(cp, tree.tolist(), {int(i): float(first.flow[i]) for i in tree}, 1e-6)
    return cp, options_for(cp, pricing), warm, first


WARM = [(4096, 16384, 21, _capi.PRICING_DANTZIG, 0), (4096, 16384, 21, _capi.PRICING_DEVEX, 7),
        (2048, 8192, 22, _capi.PRICING_CANDIDATE_LIST, 5), (2048, 8192, 23, _capi.PRICING_DEVEX_LOOP, 0)]


@pytest.mark.parametrize("lazy", ["0", "1"])
@pytest.mark.parametrize("n,m,seed,pricing,drop_every", WARM)
def test_emulated_core_agrees_with_oracle_on_warm_starts(n, m, seed, pricing, drop_every, lazy, monkeypatch):
    monkeypatch.setenv("NSX_EMU_BLOCKED", lazy)
    cp, opts, warm, first = warm_case(n, m, seed, pricing, drop_every)
    assert warm is not None and warm.artificial_in_tree >= 1
    ref = oracle.solve_canonical(cp, opts, warm=warm)
    assert ref.status == _capi.STATUS_OPTIMAL and 0 < ref.iterations < first.iterations  # the old tree helps
    cold = oracle.solve_canonical(cp, opts)
    assert cold.status == _capi.STATUS_OPTIMAL
    m_ = cp.n_arcs
    assert float(np.dot(ref.flow[:m_], cp.orig_cost)) == float(np.dot(cold.flow[:m_], cp.orig_cost))  # same optimum either way
    assert_same_solution(emu.solve_canonical(cp, opts, warm=warm), ref)


@pytest.mark.gpu
@pytest.mark.timeout(600, method="thread")
@pytest.mark.parametrize("name,make,pricing", SCANS)
def test_engine_agrees_with_oracle_on_scan_rules(name, make, pricing):
    run_scan_case(make, pricing, _capi.solve_canonical)


@pytest.mark.gpu
@pytest.mark.timeout(600, method="thread")
@pytest.mark.parametrize("n,m,seed,pricing,drop_every", WARM)
def test_engine_agrees_with_oracle_on_warm_starts(n, m, seed, pricing, drop_every):
    cp, opts, warm, first = warm_case(n, m, seed, pricing, drop_every)
    assert_same_solution(_capi.solve_canonical(cp, opts, warm=warm), oracle.solve_canonical(cp, opts, warm=warm))
