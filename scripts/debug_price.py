import sys
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np, dataclasses
from helpers import load_golden, prepare_run
from network_flow_solver_b200 import _capi
from oracle import oracle
name, idx, limit = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
arcs = [int(x) for x in sys.argv[4:]]
doc = load_golden(name); run = doc["runs"][idx]
_, cp, plan, options = prepare_run(doc, run)
eo = dataclasses.replace(plan.engine, max_iterations=limit)
a = _capi.solve_canonical(cp, eo)
m = cp.n_arcs
pi = a.potential
i = np.arange(m)
cost1 = (cp.pert_cost - 1.0) - 1e-6 * i
rc = (cost1 + pi[cp.tail]) - pi[cp.head]
st = a.state[:m]
elig_f = ((st & 1) == 0) & ((st & 2) != 0) & (rc < -1e-6)
elig_b = ((st & 1) == 0) & ((st & 4) != 0) & (rc > 1e-6)
key = np.where(elig_f, rc, np.where(elig_b, -rc, np.inf))
best = int(np.argmin(key))
print("host best from GPU state:", best, key[best], "n improving", int(np.sum(np.isfinite(key))))
zero = np.flatnonzero(((st & 1) == 0) & (np.abs(rc) <= 1e-6) & ((st & 6) != 0))
print("zero candidates:", zero[:10])
for x in arcs:
    print("arc", x, "tail", cp.tail[x], "head", cp.head[x], "pert", repr(cp.pert_cost[x]), "cost1", repr(cost1[x]), "pi_t", repr(pi[cp.tail[x]]), "pi_h", repr(pi[cp.head[x]]), "rc", repr(rc[x]), "state", st[x], "flow", a.flow[x], "upper", cp.upper[x])
order = np.argsort(key)[:12]
print("top keys:", [(int(x), float(key[x]), int(x)//4, (int(x)//4)//32) for x in order])
