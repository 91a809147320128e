"""Last check of a built library on the GPU box (about 15 s): __graft_entry__.smoke() and one resident solve of config 2
(netgen 2^16 / 2^20, Dantzig) with its pivot count, time and whether the row cache priced it.    python scripts/final_check.py"""
import sys, io, contextlib
sys.path.insert(0, '.')
import __graft_entry__ as g
buf = io.StringIO()
with contextlib.redirect_stdout(buf):
    g.smoke()
print(buf.getvalue().strip().splitlines()[-1])
import torch
from network_flow_solver_b200 import _capi
from network_flow_solver_b200.workloads import WORKLOADS
wl = WORKLOADS["netgen_2e16_dantzig"]; cp = wl.canonical(0)
opts = wl.engine_options(cp)
dev = [torch.from_numpy(getattr(cp, k)).cuda() for k in ("tail", "head", "pert_cost", "upper")]
r = _capi.solve_resident(cp, opts, [t.data_ptr() for t in dev])
print("config2", r.status, r.iterations, round(r.timing["solve_ms"], 1), "ms", round(r.iterations / r.timing["solve_ms"] * 1e3), "pivots/s star", r.stats["star_pricing"])
