"""The BASELINE.json configurations as named workloads (generator parameters + solver options).

Used by bench.py, the parity tests at full size and __graft_entry__.  Generator seeds are the ones
SURVEY.md section 8(d) fixes; every instance was screened with the oracle to end `optimal`.

Perturbation: configs 1, 2 and 4 use the reference's literal cost perturbation
(1e-10 * 1.00001^i).  At 16.7M arcs (config 3) that factor reaches ~1e62 and would swamp the costs,
so config 3 runs the same pivot rule with PERTURB_EPS_BASE = 0.
"""

from __future__ import annotations

from dataclasses import dataclass

from . import _capi
from . import generators as gen
from .canonical import PERTURB_EPS_BASE, CanonicalProblem, initial_block_size


@dataclass
class Workload:
    name: str
    description: str
    pricing: int  # _capi.PRICING_*
    eps_base: float
    make: object  # (seed_offset:int) -> ArcArrays

    def arrays(self, seed_offset: int = 0):
        return self.make(seed_offset)

    def canonical(self, seed_offset: int = 0) -> CanonicalProblem:
        return self.arrays(seed_offset).canonical(eps_base=self.eps_base)

    def engine_options(self, cp: CanonicalProblem, **overrides) -> _capi.EngineOptions:
        from .solver import special_rule  # which structure-specific rule the reference would run first, + its node mask

        m = cp.n_arcs
        rule, mask = special_rule(cp, 1e-6)
        opts = dict(
            pricing=self.pricing,
            row_scan_first=rule,
            node_mask=mask,
            block_size=initial_block_size(m),
            auto_block=True,
            ft_update_limit=64,
            max_iterations=max(100, 20 * (m + cp.n_nodes - 1)),
            tolerance=1e-6,
            trace_capacity=0,
        )
        opts.update(overrides)
        return _capi.EngineOptions(**opts)


def _transport(size):
    return lambda off: gen.transportation(size, size, cost_max=1000, seed=4096 + off)


WORKLOADS = {
    # config 1 stand-in (gridgen_8_08a.min is absent from the reference tree)
    "gridgen_8_08a_like": Workload(
        "gridgen_8_08a_like", "GRIDGEN-style 257 nodes / 2056 arcs, Devex (auto block)",
        _capi.PRICING_DEVEX, PERTURB_EPS_BASE, lambda off: gen.gridgen_like(seed=808 + off)),
    # config 2
    "netgen_2e16_dantzig": Workload(
        "netgen_2e16_dantzig", "NETGEN-style 2^16 nodes / 2^20 arcs, Dantzig",
        _capi.PRICING_DANTZIG, PERTURB_EPS_BASE,
        lambda off: gen.netgen_like(1 << 16, 1 << 20, n_sources=256, n_sinks=256, seed=1601 + off)),
    "netgen_2e16_devex": Workload(
        "netgen_2e16_devex", "NETGEN-style 2^16 nodes / 2^20 arcs, Devex (auto block)",
        _capi.PRICING_DEVEX, PERTURB_EPS_BASE,
        lambda off: gen.netgen_like(1 << 16, 1 << 20, n_sources=256, n_sinks=256, seed=1601 + off)),
    "netgen_2e16_candidate": Workload(
        "netgen_2e16_candidate", "NETGEN-style 2^16 nodes / 2^20 arcs, candidate-list pricing (the reference's default)",
        _capi.PRICING_CANDIDATE_LIST, PERTURB_EPS_BASE,
        lambda off: gen.netgen_like(1 << 16, 1 << 20, n_sources=256, n_sinks=256, seed=1601 + off)),
    # config 3 - the pricing-bandwidth-bound case
    "transport_4096": Workload(
        "transport_4096", "dense transportation 4096x4096 (16.7M arcs), row-scan pricing, eps=0",
        _capi.PRICING_DANTZIG, 0.0, _transport(4096)),
    "transport_2048": Workload(
        "transport_2048", "dense transportation 2048x2048 (4.2M arcs), row-scan pricing, eps=0",
        _capi.PRICING_DANTZIG, 0.0, _transport(2048)),
    "transport_1024": Workload(
        "transport_1024", "dense transportation 1024x1024 (1M arcs), row-scan pricing, eps=0",
        _capi.PRICING_DANTZIG, 0.0, _transport(1024)),
    # config 5 - the single largest instance (arc-sharded pricing across GPUs); eps = 0 as for config 3
    "netgen_2e20_devex": Workload(
        "netgen_2e20_devex", "NETGEN-style 2^20 nodes / 2^26 arcs, Devex block pricing (block = M/16), eps=0",
        _capi.PRICING_DEVEX, 0.0,
        lambda off: gen.netgen_like(1 << 20, 1 << 26, n_sources=4096, n_sinks=4096, seed=2026 + off)),
    "netgen_2e20_dantzig": Workload(
        "netgen_2e20_dantzig", "NETGEN-style 2^20 nodes / 2^26 arcs, Dantzig full sweeps, eps=0",
        _capi.PRICING_DANTZIG, 0.0,
        lambda off: gen.netgen_like(1 << 20, 1 << 26, n_sources=4096, n_sinks=4096, seed=2026 + off)),
    # quarter-scale stand-ins of config 5 (same family and ratio m/n = 64)
    "netgen_2e18_devex": Workload(
        "netgen_2e18_devex", "NETGEN-style 2^18 nodes / 2^24 arcs, Devex block pricing, eps=0",
        _capi.PRICING_DEVEX, 0.0,
        lambda off: gen.netgen_like(1 << 18, 1 << 24, n_sources=1024, n_sinks=1024, seed=2026 + off)),
    "netgen_2e18_dantzig": Workload(
        "netgen_2e18_dantzig", "NETGEN-style 2^18 nodes / 2^24 arcs, Dantzig full sweeps, eps=0",
        _capi.PRICING_DANTZIG, 0.0,
        lambda off: gen.netgen_like(1 << 18, 1 << 24, n_sources=1024, n_sinks=1024, seed=2026 + off)),
    # instances on which the reference switches to a structure-specific pivot rule
    "assignment_192": Workload(
        "assignment_192", "dense assignment 192x192 (36.9K arcs), assignment rule first, then Dantzig (at 256x256 the "
        "reference's rule stalls in degenerate pivots until the iteration limit - screened with the oracle)",
        _capi.PRICING_DANTZIG, PERTURB_EPS_BASE, lambda off: gen.assignment(192, cost_max=1000, seed=256 + off)),
    "shortest_path_2e16": Workload(
        "shortest_path_2e16", "one unit over a NETGEN-style graph 2^16 nodes / ~2^20 arcs, shortest-path rule first, then Dantzig",
        _capi.PRICING_DANTZIG, PERTURB_EPS_BASE, lambda off: gen.shortest_path(1 << 16, 1 << 20, seed=1701 + off)),
    "max_flow_2e12": Workload(
        "max_flow_2e12", "unit-cost max-flow form, 2^12 nodes / 2^16 arcs, max-flow rule first, then Dantzig (37 055 pivots; "
        "at 2^14 nodes the reference's rule needs 713 802 almost all degenerate pivots - screened with the oracle)",
        _capi.PRICING_DANTZIG, PERTURB_EPS_BASE, lambda off: gen.max_flow(1 << 12, 1 << 16, flow=64, seed=1801 + off)),
    "netgen_2e13_devex_loop": Workload(
        "netgen_2e13_devex_loop", "NETGEN-style 2^13 nodes / 2^17 arcs, loop-based Devex (use_vectorized_pricing=False); 16 623 "
        "pivots (at 2^16 nodes the rule, which has no anti-cycling exclusion, did not finish in the oracle within 10 minutes)",
        _capi.PRICING_DEVEX_LOOP, PERTURB_EPS_BASE,
        lambda off: gen.netgen_like(1 << 13, 1 << 17, n_sources=64, n_sinks=64, seed=1601 + off)),
    # config 4 - one instance of the batch (the batch itself is built by bench.py)
    "goto_64": Workload(
        "goto_64", "GOTO-style grid-on-torus 64x64 (4096 nodes / ~32.7K arcs), Dantzig",
        _capi.PRICING_DANTZIG, PERTURB_EPS_BASE, lambda off: gen.goto_like(64, seed=off)),
}
