import sys
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np, dataclasses
from helpers import load_golden, prepare_run
from network_flow_solver_b200 import _capi
name, idx, limit = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
doc = load_golden(name); run = doc["runs"][idx]
_, cp, plan, options = prepare_run(doc, run)
b = _capi.solve_canonical(cp, dataclasses.replace(plan.engine, max_iterations=limit))
print("trace tail", b.trace[-3:], "status", b.status)
