"""Run the oracle on the first P pivots of a full-size workload and save a compact prefix record under
tests/golden/full/<workload>_prefix<P>.json: status, pivots, hashes of the entering-arc trace at several prefix
lengths and of flows / potentials / arc states after exactly P pivots (max_iterations = P).

    python scripts/oracle_prefix.py netgen_2e20_devex 50000 [threads]

Used for BASELINE config 5 (2^20 nodes / 2^26 arcs), where a full oracle solve takes hours: bench.py's sharded leg
and tests/test_gpu_full_size.py run the CUDA engine with the same max_iterations and compare the hashes."""
import hashlib
import json
import os
import sys
import time

sys.path.insert(0, ".")
from network_flow_solver_b200.workloads import WORKLOADS  # noqa: E402
from oracle import oracle  # noqa: E402

name = sys.argv[1]
pivots = int(sys.argv[2])
threads = int(sys.argv[3]) if len(sys.argv) > 3 else len(os.sched_getaffinity(0))
wl = WORKLOADS[name]
t = time.time()
cp = wl.canonical(0)
print("built", cp.n_nodes, cp.n_arcs, f"{time.time() - t:.1f}s", flush=True)
opts = wl.engine_options(cp, trace_capacity=pivots, max_iterations=pivots)
t = time.time()
r = oracle.solve_canonical(cp, opts, threads=threads)
dt = time.time() - t
marks = sorted({k for k in (100, 1000, 2000, 5000, 10000, 20000, 50000, 100000, pivots) if k <= len(r.trace)})
rec = dict(
    workload=name, max_iterations=pivots, status=r.status, iterations=r.iterations, phase1=r.phase1_iterations,
    degenerate=r.degenerate_pivots, tree_updates=r.tree_updates, weight_resets=r.weight_resets,
    final_block_size=r.final_block_size,
    trace_sha_at={str(k): hashlib.sha256(r.trace[:k].tobytes()).hexdigest() for k in marks},
    trace_sha=hashlib.sha256(r.trace.tobytes()).hexdigest(), flow_sha=hashlib.sha256(r.flow.tobytes()).hexdigest(),
    pi_sha=hashlib.sha256(r.potential.tobytes()).hexdigest(), state_sha=hashlib.sha256(r.state.tobytes()).hexdigest(),
    oracle_seconds=dt, oracle_threads=threads,
)
print(json.dumps(rec), flush=True)
out = "tests/golden/full"
os.makedirs(out, exist_ok=True)
json.dump(rec, open(f"{out}/{name}_prefix{pivots}.json", "w"), indent=1)
