"""Star pricing on the GPU, first P pivots of a workload: per-pivot pricing / pivot time, the workers' timeline and the
row-cache statistics, with star pricing on and off.   python scripts/star_probe.py transport_4096 20000"""
import json
import os
import sys

sys.path.insert(0, ".")
from network_flow_solver_b200 import _capi  # noqa: E402
from network_flow_solver_b200.workloads import WORKLOADS  # noqa: E402

name, pivots = sys.argv[1], int(sys.argv[2])
wl = WORKLOADS[name]
cp = wl.canonical(0)
os.environ["NSX_TIMELINE"] = "1"
for star in ("1", "0"):
    os.environ["NSX_STAR"] = star
    opts = wl.engine_options(cp, max_iterations=pivots, trace_capacity=0)
    _capi.solve_canonical(cp, opts)
    r = _capi.solve_canonical(cp, opts)
    it = max(r.iterations, 1)
    sw = max(r.stats["sweeps"], 1)
    print(json.dumps({
        "workload": name, "star": star, "pivots": r.iterations, "status": r.status, "solve_ms": r.timing["solve_ms"],
        "us_per_pivot": 1e3 * r.timing["solve_ms"] / it, "pricing_us": 1e3 * r.timing["pricing_ms"] / it,
        "pivot_us": 1e3 * r.timing["pivot_ms"] / it, "arcs_per_sweep": r.arcs_priced / sw,
        "star": {k: r.stats[k] for k in ("star_pricing", "star_updates", "star_builds", "star_rescans")},
        "timeline_us_per_sweep": [round(x / 1e3 / sw, 2) for x in r.stats["handshake_ns"]],
        "phase_us": [round(v / 1.9e3 / it, 2) for v in r.stats["phase_cycles"]],
    }))
