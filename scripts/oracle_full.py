"""Run the oracle on a full-size workload and save a compact record (status, pivots, objective,
hash of flows / trace) under tests/golden/full/ for the at-scale GPU parity test."""
import sys, time, json, hashlib
sys.path.insert(0, '.')
import numpy as np
from network_flow_solver_b200.workloads import WORKLOADS
from network_flow_solver_b200.solver import objective_value
from oracle import oracle
name = sys.argv[1]; threads = int(sys.argv[2]) if len(sys.argv) > 2 else oracle.max_threads()
wl = WORKLOADS[name]
t=time.time(); cp = wl.canonical(0); print("built", cp.n_nodes, cp.n_arcs, f"{time.time()-t:.1f}s", flush=True)
opts = wl.engine_options(cp, trace_capacity=1 << 24)
t=time.time(); r = oracle.solve_canonical(cp, opts, threads=threads); dt=time.time()-t
rec = dict(workload=name, status=r.status, iterations=r.iterations, phase1=r.phase1_iterations,
           degenerate=r.degenerate_pivots, tree_updates=r.tree_updates, objective=objective_value(cp, r),
           trace_sha=hashlib.sha256(r.trace.tobytes()).hexdigest(), flow_sha=hashlib.sha256(r.flow.tobytes()).hexdigest(),
           pi_sha=hashlib.sha256(r.potential.tobytes()).hexdigest(), state_sha=hashlib.sha256(r.state.tobytes()).hexdigest(),
           oracle_seconds=dt, oracle_threads=threads, final_block_size=r.final_block_size)
print(json.dumps(rec), flush=True)
import os
out = sys.argv[3] if len(sys.argv) > 3 else "tests/golden/full"  # tests/golden/full_next: workloads of rows not yet run on hardware
os.makedirs(out, exist_ok=True)
json.dump(rec, open(f"{out}/{name}.json", "w"), indent=1)
