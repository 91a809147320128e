"""Solve a named workload on one GPU (to optimality unless MAX_PIVOTS is given) and keep a record of the outcome:
status, pivots, objective, SHA-256 of the entering-arc trace / flows / potentials / arc states, timing, star-pricing
statistics.    python scripts/solve_record.py WORKLOAD OUT.json [MAX_PIVOTS]"""
import hashlib
import json
import sys
import time

sys.path.insert(0, ".")
import numpy as np  # noqa: E402

from network_flow_solver_b200 import _capi  # noqa: E402
from network_flow_solver_b200.workloads import WORKLOADS  # noqa: E402

name, out = sys.argv[1], sys.argv[2]
wl = WORKLOADS[name]
t0 = time.time()
cp = wl.canonical(0)
kw = {"max_iterations": int(sys.argv[3])} if len(sys.argv) > 3 else {}
opts = wl.engine_options(cp, trace_capacity=0, **kw)
t1 = time.time()
r = _capi.solve_canonical(cp, opts)
m = cp.n_arcs
sha = lambda a: hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()  # noqa: E731
rec = {
    "workload": name, "description": wl.description, "nodes": int(cp.n_nodes), "arcs": int(m), "max_pivots": kw.get("max_iterations"),
    "status": int(r.status), "pivots": int(r.iterations), "phase1_pivots": int(r.phase1_iterations),
    "degenerate_pivots": int(r.degenerate_pivots), "tree_updates": int(r.tree_updates), "weight_resets": int(r.weight_resets),
    "final_block_size": int(r.final_block_size),
    "objective": float(np.dot(np.asarray(cp.orig_cost, dtype=np.float64), r.flow[:m])),
    "artificial_flow": float(r.flow[m:].sum()),
    "flow_sha": sha(r.flow), "pi_sha": sha(r.potential), "state_sha": sha(r.state),
    "solve_ms": r.timing["solve_ms"], "pivots_per_s": r.iterations / (r.timing["solve_ms"] * 1e-3),
    "us_per_pivot": {"total": 1e3 * r.timing["solve_ms"] / max(r.iterations, 1), "pricing": 1e3 * r.timing["pricing_ms"] / max(r.iterations, 1),
                     "pivot_and_tree": 1e3 * r.timing["pivot_ms"] / max(r.iterations, 1)},
    "arcs_priced_per_pivot": r.arcs_priced / max(r.iterations, 1), "sweeps": r.stats["sweeps"],
    "star": {k: r.stats[k] for k in ("star_pricing", "star_updates", "star_builds", "star_rescans", "blk_rebuilds")},
    "avg_rehung_subtree": r.stats["sum_subtree"] / max(r.tree_updates, 1), "grid": r.stats["grid"],
    "pivot_phase_us": {k: round(v / 1.9e3 / max(r.iterations, 1), 3) for k, v in zip(
        ["walk", "residuals", "ratio", "flow", "bookkeeping", "snapshot", "permute", "stem", "potentials", "cadence"], r.stats["phase_cycles"])},
    "instance_build_s": round(t1 - t0, 1), "wall_s": round(time.time() - t1, 1),
}
json.dump(rec, open(out, "w"), indent=1)
print(json.dumps(rec))
