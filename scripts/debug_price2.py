import sys, os
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
os.environ["NSX_DEBUG"] = "1"
import numpy as np, dataclasses
from helpers import load_golden, prepare_run
from network_flow_solver_b200 import _capi
name, idx, limit = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
doc = load_golden(name); run = doc["runs"][idx]
_, cp, plan, options = prepare_run(doc, run)
m = cp.n_arcs
a = _capi.solve_canonical(cp, dataclasses.replace(plan.engine, max_iterations=limit - 1))
pi = a.potential; i = np.arange(m)
rc_host = ((cp.pert_cost - 1.0) - 1e-6 * i + pi[cp.tail]) - pi[cp.head]
b = _capi.solve_canonical(cp, dataclasses.replace(plan.engine, max_iterations=limit))
raw = open("/tmp/nsx_dbg.bin", "rb").read()
rc_dev = np.frombuffer(raw[: m * 8], dtype=np.float64); st_dev = np.frombuffer(raw[m * 8 :], dtype=np.uint32)
priced = st_dev != 0xffffffff
print("arcs priced in last sweep:", int(priced.sum()), "of", m, " unpriced with state elig:", np.flatnonzero(~priced & ((a.state[:m] & 1) == 0) & ((a.state[:m] & 6) != 0))[:20])
bad = np.flatnonzero(priced & (rc_dev != rc_host))
print("rc mismatches:", bad.size, bad[:20])
for x in bad[:6]:
    print(x, repr(rc_dev[x]), repr(rc_host[x]), "tail", cp.tail[x], "head", cp.head[x])
sb = np.flatnonzero(priced & (st_dev != a.state[:m]))
print("state mismatches", sb.size, sb[:10], st_dev[sb[:10]], a.state[sb[:10]])
cand = np.frombuffer(open("/tmp/nsx_dbg_cand.bin", "rb").read(), dtype=np.dtype([("key", "f8"), ("arc2", "i4"), ("zero2", "i4")]))
per = cand[:1024]
valid = np.flatnonzero(per["arc2"] >= 0)
print("threads with candidates:", valid.size, "best per-thread:", per[valid[np.argmin(per["key"][valid])]], "thread", valid[np.argmin(per["key"][valid])])
print("thread 394:", per[394], " thread 27:", per[27])
print("warp buf:", [(int(c["arc2"]), float(c["key"])) for c in cand[1024:1056]])
print("after first reduce:", cand[1056], " final:", cand[1057])
wr = cand[2048:2048+1024]
print("after warp reduce, lane0 of warps:", [(int(wr[w*32]["arc2"])) for w in range(32)])
print("warp 12 lanes after reduce:", [int(x) for x in wr[384:416]["arc2"]])
print("warp 12 lanes before:", [int(x) for x in per[384:416]["arc2"]])
di = np.frombuffer(open("/tmp/nsx_dbg_i.bin","rb").read(), dtype=np.int32)
print("k.arc2 stored by lane0 of warps:", di[0:32].tolist())
print("buf[warp].arc2 read back:", di[32:64].tolist())
print("buf addr:", [hex(x & 0xffffffff) for x in di[64:68].tolist()])
print("warp0 stage2 read arc2:", di[96:128].tolist())
print("warp0 stage2 addr:", [hex(x & 0xffffffff) for x in di[128:132].tolist()])
