"""Drop-in entry point: ``solve_min_cost_flow`` backed by the device-resident pivot loop.

Same signature, option meaning, statuses and error behaviour as the reference façade
(reference: src/network_solver/solver.py:13-104 -> simplex.py:99-265,1446-1765).  The host does
exactly three things: resolve the options the way ``NetworkSimplex.__init__`` does, canonicalise
the problem into the engine's structure-of-arrays, and turn the raw arrays that come back over
the C ABI into a ``FlowResult`` with the reference's rounding (simplex.py:1703-1765).  Pricing,
ratio test, flow / tree / potential updates all run inside libnsx_b200.so on the GPU.
"""

from __future__ import annotations

import logging
import time
import warnings
from dataclasses import dataclass

import numpy as np

from . import _capi
from .canonical import (
    NET_ASSIGNMENT,
    NET_GENERAL,
    NET_MAX_FLOW,
    NET_SHORTEST_PATH,
    NET_TRANSPORTATION,
    PERTURB_EPS_BASE,
    CanonicalProblem,
    canonicalize,
    initial_block_size,
    is_likely_goto,
)
from .data import Basis, FlowResult, NetworkProblem, ProgressCallback, ProgressInfo, SolverOptions
from .exceptions import SolverConfigurationError, UnboundedProblemError

_log = logging.getLogger(__name__)

_STATUS_TEXT = {
    _capi.STATUS_OPTIMAL: "optimal",
    _capi.STATUS_INFEASIBLE: "infeasible",
    _capi.STATUS_ITERATION_LIMIT: "iteration_limit",
    _capi.STATUS_ITERATION_LIMIT_P1: "iteration_limit",
}


@dataclass
class ResolvedPlan:
    """Outcome of option resolution: what the engine is asked to do."""

    strategy: str
    engine: _capi.EngineOptions
    scaling: object = None  # scaling.ScalingFactors when the instance was rescaled before the solve


def reachable_from(n_nodes: int, tail: np.ndarray, head: np.ndarray, source: int) -> np.ndarray:
    """uint8[n_nodes]: 1 for every node reachable from `source` along tail -> head (frontier-at-a-time BFS)."""
    order = np.argsort(tail, kind="stable")
    heads = np.asarray(head)[order]
    start = np.searchsorted(np.asarray(tail)[order], np.arange(n_nodes + 1))
    mask = np.zeros(n_nodes, dtype=np.uint8)
    mask[source] = 1
    frontier = np.asarray([source], dtype=np.int64)
    while frontier.size:
        lo, cnt = start[frontier], start[frontier + 1] - start[frontier]
        total = int(cnt.sum())
        if total == 0:
            break
        offs = np.arange(total) - np.repeat(np.cumsum(cnt) - cnt, cnt)
        nxt = heads[np.repeat(lo, cnt) + offs]
        nxt = np.unique(nxt[mask[nxt] == 0])
        mask[nxt] = 1
        frontier = nxt.astype(np.int64)
    return mask


def special_rule(cp: CanonicalProblem, tol: float) -> tuple[int, np.ndarray | None]:
    """Mirror of select_pivot_strategy (specialized_pivots.py:452-527): which structure-specific entering rule runs before
    the configured one, and the node mask the shortest-path rule needs.  Supplies are the shifted ones (node_supply after
    _build_arcs), the tolerance is the solver's."""
    kind = cp.network_type
    supply = np.asarray(cp.supply, dtype=np.float64)
    if kind == NET_GENERAL:
        return _capi.SPECIAL_NONE, None
    if kind == NET_TRANSPORTATION:
        return _capi.SPECIAL_ROW_SCAN, None
    if kind == NET_ASSIGNMENT:
        return _capi.SPECIAL_ASSIGNMENT, None
    if kind == NET_MAX_FLOW:  # first supply node, then first demand node, in index order (:487-499)
        src = np.flatnonzero(supply[1:] > tol)
        snk = np.flatnonzero(supply[1:] < -tol)
        return (_capi.SPECIAL_MAX_FLOW if src.size and snk.size else _capi.SPECIAL_NONE), None
    if kind == NET_SHORTEST_PATH:  # first node with supply +1, then first OTHER node with supply -1 (:501-516)
        src = np.flatnonzero(np.abs(supply[1:] - 1.0) <= tol)
        if src.size:
            s = int(src[0]) + 1
            snk = [v for v in (np.flatnonzero(np.abs(supply[1:] + 1.0) <= tol) + 1).tolist() if v != s]
            if snk:
                return _capi.SPECIAL_SHORTEST_PATH, reachable_from(cp.n_nodes, cp.tail, cp.head, s)
        return _capi.SPECIAL_NONE, None
    # NET_BIPARTITE_MATCHING: the reference's rule for it (BipartiteMatchingPivotStrategy, specialized_pivots.py:212-273)
    # walks Python sets of node indices built from sets of node-id strings, so the arc it picks depends on PYTHONHASHSEED -
    # there is no pivot sequence to reproduce.  The instance is solved with the configured pricing rule instead: same
    # optimal objective, pivot-for-pivot parity with one particular reference run is not defined for this type.
    _log.warning(
        "network type %r: the reference's matching pivot rule is hash-seed dependent; solving with the configured "
        "pricing rule (optimal objective guaranteed, pivot sequence not comparable)", kind)
    return _capi.SPECIAL_NONE, None


def resolve_plan(
    cp: CanonicalProblem,
    options: SolverOptions,
    max_iterations: int | None,
    *,
    goto_like: bool = False,
    trace_capacity: int = 0,
    device: int = 0,
) -> ResolvedPlan:
    """Mirror of the decisions taken in NetworkSimplex.__init__ / solve().

    * pricing rule: explicit choice, else Dantzig on grid-on-torus structure
      (simplex.py:312-377);
    * structure override: transportation instances are priced by the row-scan rule first
      (simplex.py:259-261,1060-1064; specialized_pivots.py:80-120);
    * Devex block size: fixed int, or the static heuristic plus runtime adaptation
      (simplex.py:195-211; simplex_adaptive.py:70-151);
    * iteration budget: argument, else options, else max(100, 20*(arcs incl. artificial))
      (simplex.py:1465-1470).
    """
    strategy = options.pricing_strategy
    if not options.explicit_pricing_strategy and goto_like:
        strategy = "dantzig"
    if strategy == "dantzig":
        pricing = _capi.PRICING_DANTZIG
    elif strategy == "devex":
        # use_vectorized_pricing=False selects the reference's loop-based block scan (simplex_pricing.py:205-269)
        pricing = _capi.PRICING_DEVEX if options.use_vectorized_pricing else _capi.PRICING_DEVEX_LOOP
    elif strategy in ("candidate_list", "adaptive"):
        # "adaptive" (the reference's default) only leaves its candidate-list stage after 5 consecutive searches
        # that return None, and a search that returns None ends the phase: it IS the candidate-list rule
        # (simplex_pricing.py:545-639; tests/test_next_candidate_list.py pins this against the reference)
        pricing = _capi.PRICING_CANDIDATE_LIST
    else:
        raise SolverConfigurationError(
            f"Unknown pricing strategy '{strategy}'. Valid options: 'devex', 'dantzig', 'candidate_list', 'adaptive'."
        )
    row_scan, node_mask = special_rule(cp, options.tolerance)
    auto = options.block_size is None or isinstance(options.block_size, str)
    block = initial_block_size(cp.n_arcs) if auto else int(options.block_size)
    if max_iterations is None:
        if options.max_iterations is not None:
            max_iterations = options.max_iterations
        else:
            max_iterations = max(100, 20 * (cp.n_arcs + cp.n_nodes - 1))
    eng = _capi.EngineOptions(
        pricing=pricing,
        row_scan_first=row_scan,
        block_size=block,
        auto_block=auto,
        ft_update_limit=options.ft_update_limit,
        max_iterations=int(max_iterations),
        tolerance=options.tolerance,
        trace_capacity=trace_capacity,
        device=device,
        node_mask=node_mask,
    )
    return ResolvedPlan(strategy=strategy, engine=eng)


def _np_round12(x: float) -> float:
    """NumPy's scale-rint-unscale rounding, which the reference applies to np.float64 values
    (6749969302.0 comes back as 6749969301.999999)."""
    return float(round(np.float64(x), 12))


def objective_value(cp: CanonicalProblem, raw: _capi.RawSolution, divisor: float | None = None) -> float:
    """Sum over real arcs in index order of (flow+shift)*original cost, rounded like the
    reference (simplex.py:1703-1714,1759); `divisor` = cost_scale * supply_scale when the instance was
    rescaled (the division happens before the rounding, simplex.py:1753-1759)."""
    m = cp.n_arcs
    if m == 0:
        return 0.0
    values = raw.flow[:m] + cp.shift
    total = float(np.cumsum(values * cp.orig_cost)[-1])  # cumsum folds left to right
    if divisor is not None:
        total = total / divisor
    touched = bool(np.any(raw.state[:m] & _capi.ARC_TOUCHED))
    return _np_round12(total) if touched else float(round(total, 12))


def finish(
    cp: CanonicalProblem, raw: _capi.RawSolution, options: SolverOptions, scaling=None
) -> FlowResult:
    """Raw engine arrays -> FlowResult (simplex.py:1600-1624,1703-1765); `scaling` = factors applied by prepare()."""
    if raw.status == _capi.STATUS_UNBOUNDED:
        key = cp.arc_keys[raw.unbounded_arc] if cp.arc_keys is not None else None
        raise UnboundedProblemError(
            "Unbounded problem detected: entering arc can increase indefinitely without "
            "hitting any capacity constraint. This indicates a negative-cost cycle with "
            "infinite capacity.",
            entering_arc=key,
            reduced_cost=raw.unbounded_rc,
        )
    status = _STATUS_TEXT[raw.status]
    if raw.status in (_capi.STATUS_INFEASIBLE, _capi.STATUS_ITERATION_LIMIT_P1):
        return FlowResult(
            objective=0.0, flows={}, status=status, iterations=raw.iterations, duals={}
        )
    tol = options.tolerance
    m = cp.n_arcs
    values = raw.flow[:m] + cp.shift
    touched = (raw.state[:m] & _capi.ARC_TOUCHED) != 0
    flows: dict[tuple[str, str], float] = {}
    if cp.arc_keys is not None:
        np_typed: dict[tuple[str, str], bool] = {}
        keys = cp.arc_keys
        for i in np.flatnonzero((values != 0.0) | touched).tolist():
            k = keys[i]
            if k in flows:
                flows[k] += float(values[i])
                np_typed[k] = np_typed[k] or bool(touched[i])
            else:
                flows[k] = float(values[i])
                np_typed[k] = bool(touched[i])
        for k, v in list(flows.items()):
            if abs(v) <= tol:
                del flows[k]
            else:
                flows[k] = _np_round12(v) if np_typed[k] else float(round(v, 12))
    duals: dict[str, float] = {}
    if cp.node_ids is not None:
        pot = raw.potential.tolist()
        for i in range(1, cp.n_nodes):
            duals[cp.node_ids[i]] = float(round(pot[i], 12))
    basis = None
    if cp.arc_keys is not None:
        in_tree = np.flatnonzero(raw.state[:m] & _capi.ARC_IN_TREE).tolist()
        basis = Basis(
            tree_arcs={cp.arc_keys[i] for i in in_tree},
            arc_flows={cp.arc_keys[i]: float(raw.flow[i]) for i in in_tree},
        )
    divisor = None
    if scaling is not None and scaling.enabled:  # simplex.py:1753-1756: flows after their rounding, objective before
        flows = {k: v / scaling.supply_scale for k, v in flows.items()}
        divisor = scaling.cost_scale * scaling.supply_scale
    return FlowResult(
        objective=objective_value(cp, raw, divisor),
        flows=flows,
        status=status,
        iterations=raw.iterations,
        duals=duals,
        basis=basis,
    )


def prepare(
    problem: NetworkProblem,
    options: SolverOptions | None = None,
    max_iterations: int | None = None,
    *,
    trace_capacity: int = 0,
    device: int = 0,
    eps_base: float = PERTURB_EPS_BASE,
) -> tuple[CanonicalProblem, ResolvedPlan, SolverOptions]:
    """Host-side half of the call: options + canonical arrays, no device work."""
    options = options if options is not None else SolverOptions()
    factors = None
    if options.auto_scale:  # simplex.py:103-114
        from .scaling import compute_scaling_factors, scale_problem, should_scale_problem

        if should_scale_problem(problem):
            factors = compute_scaling_factors(problem)
            problem = scale_problem(problem, factors)
    cp = canonicalize(problem, options.tolerance, eps_base)
    goto = is_likely_goto(problem, cp.n_arcs, options.tolerance)
    plan = resolve_plan(
        cp,
        options,
        max_iterations,
        goto_like=goto,
        trace_capacity=trace_capacity,
        device=device,
    )
    plan.scaling = factors
    return cp, plan, options


def solve_min_cost_flow(
    problem: NetworkProblem,
    options: SolverOptions | None = None,
    max_iterations: int | None = None,
    progress_callback: ProgressCallback | None = None,
    progress_interval: int = 100,
    warm_start_basis: Basis | None = None,
    *,
    device: int = 0,
) -> FlowResult:
    """Solve a minimum-cost flow problem on the GPU; drop-in for the reference call
    (solver.py:13-104).  Raises DeviceEngineError when the CUDA engine is unavailable."""
    t_start = time.time()
    if progress_callback is not None:
        # The reference calls back every `progress_interval` pivots (simplex.py:1143-1154).  Here the pivot loop is one
        # device-resident kernel with no host round trip per pivot: the callback is invoked ONCE, after the solve, with the
        # totals - enough for logging, not for cancellation.
        warnings.warn(
            "progress_callback is invoked once, after the solve (the pivot loop is device-resident; "
            f"progress_interval={progress_interval} is not honoured)", RuntimeWarning, stacklevel=2)
    cp, plan, options = prepare(problem, options, max_iterations, device=device)
    warm = None
    if warm_start_basis is not None:  # simplex.py:1494-1530; None = basis rejected, cold start
        from .warm_start import apply_basis

        _log.info("Attempting to apply warm-start basis")
        warm = apply_basis(cp, warm_start_basis, options.tolerance)
        if warm is None:
            _log.info("Warm-start failed, performing cold start")
    raw = _capi.solve_canonical(cp, plan.engine, warm=warm)
    result = finish(cp, raw, options, plan.scaling)
    if progress_callback is not None:
        in_phase_two = raw.iterations > raw.phase1_iterations or raw.status == _capi.STATUS_OPTIMAL
        progress_callback(ProgressInfo(
            iteration=int(raw.iterations),
            max_iterations=int(plan.engine.max_iterations),
            phase=2 if in_phase_two else 1,
            phase_iterations=int(raw.iterations - raw.phase1_iterations) if in_phase_two else int(raw.iterations),
            objective_estimate=float(result.objective),
            elapsed_time=time.time() - t_start,
        ))
    if raw.status in (_capi.STATUS_OPTIMAL, _capi.STATUS_ITERATION_LIMIT):
        rate = (raw.degenerate_pivots / raw.iterations * 100) if raw.iterations > 0 else 0.0
        # the reference prints this line unconditionally (simplex.py:1672-1674)
        print(f"  > Degeneracy: {raw.degenerate_pivots}/{raw.iterations} pivots ({rate:.1f}%)")
    return result
