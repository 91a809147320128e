"""SASS evidence of the shipped library (runs without a GPU): counts of the TMA / mbarrier instructions per kernel and the
listing of the hot loop of the config-3 pricing sweep (Phase-1 Dantzig rule, two tiles per step, potentials in shared
memory) inside nsx_resident_kernel.    python scripts/sass_excerpt.py > profiles/rNN/sass_hot_loop.md"""
import re
import subprocess
import sys
from collections import Counter

LIB = "network_flow_solver_b200/csrc/libnsx_b200.so"
sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
kern, per = None, {}
for line in sass.splitlines():
    m = re.match(r"\s+Function : (\S+)", line)
    if m:
        kern = m.group(1); per[kern] = Counter(); continue
    m = re.match(r"\s+/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and kern:
        op = m.group(1)
        per[kern][op.split(".")[0]] += 1
        if op.startswith("SYNCS") or op.startswith("UBLKCP"): per[kern][op] += 1
print("# SASS of `libnsx_b200.so` (sm_100a cubin; `cuobjdump -sass`)\n")
print("| kernel | instructions | UBLKCP (TMA bulk copy) | SYNCS.* (mbarrier) | DADD/DMUL/DSETP | LDS |\n|---|---|---|---|---|---|")
for k, c in per.items():
    tot = sum(v for o, v in c.items() if "." not in o)
    print(f"| `{k}` | {tot} | {c['UBLKCP']} | {c['SYNCS']} | {c['DADD'] + c['DMUL'] + c['DSETP']} | {c['LDS']} |")
print("\nNo `UTMALDG` (tensor-map TMA): the tile store is a 1-D stream, moved with `cp.async.bulk.shared.global` = `UBLKCP.S.G`.\n")

# hot loop of the worker CTAs: third inlined copy of the Phase-1 Dantzig ring with potentials in shared memory
ins = []
on = False
for line in sass.splitlines():
    m = re.match(r"\s+Function : (\S+)", line)
    if m: on = m.group(1) == "nsx_resident_kernel"; continue
    if not on: continue
    m = re.match(r"\s+/\*([0-9a-f]{4,6})\*/\s+(.*?);", line)
    if m: ins.append((int(m.group(1), 16), m.group(2).strip()))
idx = [i for i, (a, t) in enumerate(ins) if re.search(r"DMUL R\d+, R\d+, UR", t)]
groups = []
for i in idx:
    if not groups or i - groups[-1][-1] > 400: groups.append([i])
    else: groups[-1].append(i)
def loop_of(g):
    s = g[0]
    while not ins[s][1].startswith("SYNCS.PHASECHK"): s -= 1
    for j in range(s - 1, s - 14, -1):
        if ins[j][1].startswith("SYNCS.PHASECHK"): s = j
    e = g[-1]
    while "SYNCS.ARRIVE" not in ins[e][1]: e += 1
    return s, e


# the instantiation for the pre-scaled store (LAYOUT = 1, two tiles per step, potentials in shared memory): int16 costs are
# converted directly (I2F.F64.S16), potentials are gathered without an address multiply, nothing is loaded from global memory;
# the last such loop in the kernel is the copy inlined into the worker CTAs
picked = None
for g in groups:
    if len(g) < 7: continue
    s, e = loop_of(g)
    body = [t for _, t in ins[s:e]]
    if any("LDG" in t for t in body) or not any(t.startswith("I2F.F64.S16") for t in body): continue
    if any(re.match(r"IMAD R\d+, R\d+, 0x8, R\d+", t) for t in body): continue
    picked = (s, e)
s, e = picked
print(f"## Pricing loop of the sweep workers, one step = 2 tiles x 4 arcs per thread (0x{ins[s][0]:x} .. 0x{ins[e + 1][0]:x})\n")
print("Executed on the state-free fast path (`plain`): the loop head, the uint16 node-id loads, the int16 cost loads, the\n"
      "arithmetic block (8 x [I2F.F64, DMUL 1e-6*idx, 3 x DADD, 2 x LDS.64 potential gather, DSETP.LE.OR]), the arrive.\n```")
for a, t in ins[s:e + 2]: print(f"/*{a:06x}*/ {t}")
print("```")
