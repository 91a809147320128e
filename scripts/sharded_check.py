"""torchrun --nproc-per-node G scripts/sharded_check.py WORKLOAD [MAX_ITERATIONS]
Arc-sharded solve on G GPUs; rank 0 also solves alone and checks that the results are identical."""
import os, sys, hashlib, time
sys.path.insert(0, '.')
import numpy as np
import torch
import torch.distributed as dist
from network_flow_solver_b200 import _capi
from network_flow_solver_b200.sharded import MailboxRing, solve_canonical_sharded
from network_flow_solver_b200.workloads import WORKLOADS

rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
name = sys.argv[1]
wl = WORKLOADS[name]; cp = wl.canonical(0)
kw = {"max_iterations": int(sys.argv[2])} if len(sys.argv) > 2 else {}
opts = wl.engine_options(cp, device=local, trace_capacity=1 << 22, **kw)
ring = MailboxRing(local, dist)
r = solve_canonical_sharded(cp, opts, ring)   # warm-up
t0 = time.perf_counter()
r = solve_canonical_sharded(cp, opts, ring)
wall = time.perf_counter() - t0
sig = hashlib.sha256(r.trace.tobytes() + r.flow.tobytes() + r.potential.tobytes() + r.state.tobytes()).hexdigest()
sigs = [None] * world
dist.all_gather_object(sigs, (r.status, r.iterations, sig))
if rank == 0:
    assert len(set(sigs)) == 1, f"ranks disagree: {sigs}"
    single = _capi.solve_canonical(cp, opts)
    sig1 = hashlib.sha256(single.trace.tobytes() + single.flow.tobytes() + single.potential.tobytes() + single.state.tobytes()).hexdigest()
    ok = (single.status, single.iterations, sig1) == sigs[0]
    print(f"{name} world={world}: status {r.status} pivots {r.iterations} sweeps {r.stats['sweeps']} solve_ms {r.timing['solve_ms']:.1f} "
          f"(single GPU {single.timing['solve_ms']:.1f}) pricing_ms {r.timing['pricing_ms']:.1f} (single {single.timing['pricing_ms']:.1f}) "
          f"exchange_ms {r.timing['exchange_ms']:.1f} = {1e3 * r.timing['exchange_ms'] / max(r.stats['sweeps'], 1):.2f} us/sweep; "
          f"identical to the single-GPU solve: {ok}")
    assert ok
ring.close()
dist.destroy_process_group()
