import sys
sys.path.insert(0, '.')
import torch
from network_flow_solver_b200 import _capi
from network_flow_solver_b200.workloads import WORKLOADS
wl = WORKLOADS["transport_4096"]; cp = wl.canonical(0)
opts = wl.engine_options(cp)
dev = [torch.from_numpy(getattr(cp, k)).cuda() for k in ("tail", "head", "pert_cost", "upper")]
ptrs = [t.data_ptr() for t in dev]
_capi.sweep_probe(cp, opts, ptrs, 20)
for n in (50, 400):
    r = _capi.sweep_probe(cp, opts, ptrs, n)
    print(n, {k: round(v, 3) for k, v in r.timing.items()}, [round(x / 1e3 / n, 2) for x in r.stats["handshake_ns"]])
