"""Sweep probe: time K pricing sweeps of a workload's initial state (no pivots). Usage:
   python scripts/probe.py WORKLOAD [SWEEPS]   (env NSX_* knobs apply)"""
import sys
sys.path.insert(0, '.')
import torch
from network_flow_solver_b200 import _capi
from network_flow_solver_b200.workloads import WORKLOADS
name = sys.argv[1]; sweeps = int(sys.argv[2]) if len(sys.argv) > 2 else 200
wl = WORKLOADS[name]; cp = wl.canonical(0)
opts = wl.engine_options(cp)
dev = [torch.from_numpy(getattr(cp, k)).cuda() for k in ("tail", "head", "pert_cost", "upper")]
ptrs = [t.data_ptr() for t in dev]
_capi.sweep_probe(cp, opts, ptrs, 20)
r = _capi.sweep_probe(cp, opts, ptrs, sweeps)
bpa = r.stats["bytes_per_arc"] + (4 if wl.pricing == 1 else 0)
us = r.timing["solve_ms"] * 1e3 / sweeps
print(f"{name}: {sweeps} sweeps, {us:.2f} us/sweep, {r.arcs_priced / sweeps:.0f} arcs/sweep, {bpa} B/arc -> "
      f"{r.arcs_priced * bpa / (r.timing['solve_ms'] * 1e-3) / 1e9:.0f} GB/s; grid {r.stats['grid']} stages {r.stats['ring_stages']} "
      f"sync_wait {r.timing['sync_ms'] * 1e3 / sweeps:.2f} us/sweep")
print("  handshake us/sweep [seen, enter, pi, tiles, reduced, arrived | all_arrived, merged]:",
      [round(x / 1e3 / sweeps, 2) for x in r.stats["handshake_ns"]])
