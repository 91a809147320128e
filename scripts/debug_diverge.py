"""Debug helper: find the first pivot count at which the engine state differs from the oracle."""
import sys
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np
from helpers import load_golden, prepare_run
from network_flow_solver_b200 import _capi
from oracle import oracle
import dataclasses

name, idx = sys.argv[1], int(sys.argv[2])
doc = load_golden(name); run = doc["runs"][idx]
_, cp, plan, options = prepare_run(doc, run)
def cmp(limit):
    eo = dataclasses.replace(plan.engine, max_iterations=limit)
    a = _capi.solve_canonical(cp, eo); b = oracle.solve_canonical(cp, eo)
    same = (np.array_equal(a.trace, b.trace) and np.array_equal(a.flow, b.flow) and np.array_equal(a.potential, b.potential)
            and np.array_equal(a.state, b.state) and a.status == b.status and a.artificial_with_flow == b.artificial_with_flow)
    return same, a, b
lo, hi = 0, len(run["trace"]) + 5
while lo < hi:
    mid = (lo + hi) // 2
    ok, a, b = cmp(mid)
    if ok: lo = mid + 1
    else: hi = mid
print("first differing limit:", lo)
ok, a, b = cmp(lo)
print("status", a.status, b.status, "iters", a.iterations, b.iterations, "p1", a.phase1_iterations, b.phase1_iterations, "art", a.artificial_with_flow, b.artificial_with_flow)
print("trace eq", np.array_equal(a.trace, b.trace), a.trace[-3:], b.trace[-3:])
for nm in ("flow", "potential", "state"):
    x, y = getattr(a, nm), getattr(b, nm)
    d = np.flatnonzero(x != y)
    print(nm, "diff count", d.size, d[:10], x[d[:10]], y[d[:10]])
print("m", cp.n_arcs, "n", cp.n_nodes)
