import os, sys, time
sys.path.insert(0, ".")
from network_flow_solver_b200 import _capi
from network_flow_solver_b200.workloads import WORKLOADS
name = sys.argv[1]
wl = WORKLOADS[name]
cp = wl.canonical(0)
for star in ("0", "1"):
    os.environ["NSX_STAR"] = star
    opts = wl.engine_options(cp, trace_capacity=0, spin_timeout_ms=3000)
    t = time.time()
    try:
        r = _capi.solve_canonical(cp, opts)
        print(name, "star", star, "status", r.status, "pivots", r.iterations, "ms", round(r.timing["solve_ms"], 1), {k: r.stats[k] for k in ("star_pricing", "star_updates", "star_builds", "star_rescans", "sweeps")}, flush=True)
    except Exception as e:
        print(name, "star", star, "FAILED", str(e)[:200], round(time.time() - t, 1), "s", flush=True)
