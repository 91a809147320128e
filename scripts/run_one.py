"""Run one resident solve of a named workload (used under ncu). Usage: run_one.py WORKLOAD [REPS] [MAX_ITERATIONS]"""
import sys
sys.path.insert(0, '.')
import torch
from network_flow_solver_b200 import _capi
from network_flow_solver_b200.workloads import WORKLOADS
name = sys.argv[1]; reps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
wl = WORKLOADS[name]; cp = wl.canonical(0)
kw = {"max_iterations": int(sys.argv[3])} if len(sys.argv) > 3 else {}
opts = wl.engine_options(cp, **kw)
dev = [torch.from_numpy(getattr(cp, k)).cuda() for k in ("tail", "head", "pert_cost", "upper")]
for _ in range(reps):
    r = _capi.solve_resident(cp, opts, [t.data_ptr() for t in dev])
print(name, "status", r.status, "pivots", r.iterations, "solve_ms", round(r.timing["solve_ms"], 2), "grid", r.stats["grid"],
      "pricing_ms", round(r.timing["pricing_ms"], 2), "pivot_ms", round(r.timing["pivot_ms"], 2))
