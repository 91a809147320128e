#!/usr/bin/env python
"""Float-cost golden family (north_star: "on float-cost inputs the objective must agree within 1e-9"): the synthetic
families of make_golden.py with costs drawn from U(0, 1) * 10^3 - nothing about them is integral, every reduced cost and
potential carries rounding - recorded from the UNMODIFIED reference in the build container.

    NUMBA_CACHE_DIR=/tmp/numba_cache python tests/golden/make_float_costs_golden.py

The fixtures have the format of make_golden.py, so the oracle, the emulated device core and the CUDA engine replay them in
the ordinary golden tests - bit for bit (entering arcs, flows, potentials, objective), which implies the 1e-9 bound;
tests/test_float_costs.py states the bound itself."""
from __future__ import annotations

import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent))
import make_golden as mg  # noqa: E402  (imports the reference)

gen = mg.gen


def float_costs(arrays, seed):
    rng = np.random.default_rng(seed)
    arrays.cost = rng.random(arrays.cost.shape[0]) * 1e3
    arrays.family = arrays.family + "_float_costs"
    return arrays


def family(name, arrays, variants):
    problem = gen.to_network_problem(arrays)
    rp = mg.ref_build([{"id": n.id, "supply": n.supply} for n in problem.nodes.values()],
                      [{"tail": a.tail, "head": a.head, "capacity": a.capacity, "cost": a.cost} for a in problem.arcs],
                      directed=True, tolerance=1e-3)
    mg.dump(name, rp, [mg.run_reference(rp, v) for v in variants], note=f"{arrays.family} seed={arrays.seed}, costs U(0,1)*1e3")


if __name__ == "__main__":
    DZ, DX = mg.DZ, mg.DX
    CL = {"pricing_strategy": "candidate_list", "explicit_pricing_strategy": True, "auto_scale": False}
    family("float_netgen_64", float_costs(gen.netgen_like(64, 512, n_sources=4, n_sinks=4, seed=41), 141), [DZ, DX, CL, {}])
    family("float_netgen_256", float_costs(gen.netgen_like(256, 2048, n_sources=8, n_sinks=8, seed=42), 142), [DZ, DX, {**DX, "block_size": 64}])
    family("float_netgen_512", float_costs(gen.netgen_like(512, 4096, n_sources=8, n_sinks=8, seed=43), 143), [DZ, DX])
    family("float_transport_32", float_costs(gen.transportation(32, 32, cost_max=100, seed=44), 144), [DZ, DX])
    family("float_gridgen_257", float_costs(gen.gridgen_like(), 145), [DX, DZ])
