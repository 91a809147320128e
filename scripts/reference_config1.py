"""Time the UNMODIFIED reference (pure Python) on the config-1 stand-in (GRIDGEN-style 257 nodes / 2056 arcs, Devex) in
the build container, and check that it agrees with the oracle.  Needs /root/reference (not available on the GPU box)."""
import io, os, sys, time
from contextlib import redirect_stdout
sys.path.insert(0, '.'); sys.path.insert(0, '/root/reference/src')
os.environ.setdefault("NUMBA_CACHE_DIR", "/tmp/numba_cache")
from network_solver import SolverOptions as RefOptions, build_problem as ref_build, solve_min_cost_flow as ref_solve
from network_flow_solver_b200 import generators as gen
from network_flow_solver_b200.workloads import WORKLOADS
from network_flow_solver_b200.solver import finish, prepare
from network_flow_solver_b200 import SolverOptions
from oracle import oracle

arrays = WORKLOADS["gridgen_8_08a_like"].arrays(0)
p = gen.to_network_problem(arrays)
nodes = [{"id": n.id, "supply": n.supply} for n in p.nodes.values()]
arcs = [{"tail": a.tail, "head": a.head, "capacity": a.capacity, "cost": a.cost, "lower": a.lower} for a in p.arcs]
rp = ref_build(nodes, arcs, directed=True, tolerance=p.tolerance)
kw = dict(pricing_strategy="devex", explicit_pricing_strategy=True, auto_scale=False)
best = None
for _ in range(3):
    t0 = time.perf_counter()
    with redirect_stdout(io.StringIO()):
        r = ref_solve(rp, RefOptions(**kw))
    dt = time.perf_counter() - t0
    best = dt if best is None else min(best, dt)
cp, plan, opts = prepare(p, SolverOptions(**kw))
o = finish(cp, oracle.solve_canonical(cp, plan.engine), opts)
assert (o.status, o.iterations, o.objective) == (r.status, r.iterations, r.objective), (o, r)
print(f"reference (pure Python, 1 core): status {r.status}, {r.iterations} pivots, objective {r.objective}, best of 3 = "
      f"{best:.3f} s = {r.iterations / best:.0f} pivots/s; oracle agrees on status / pivots / objective")
