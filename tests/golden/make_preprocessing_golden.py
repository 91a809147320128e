#!/usr/bin/env python
"""Golden vectors for the preprocessing wrappers (preprocessing.py), recorded from the
UNMODIFIED reference:
    NUMBA_CACHE_DIR=/tmp/numba_cache python tests/golden/make_preprocessing_golden.py
For every case: the original problem, what the reference's preprocess_problem() made of it (node / arc order, the
statistics, both maps), the reference's solve of the REDUCED problem (entering-arc trace, result) and what
translate_result() turned that into for the original problem."""

from __future__ import annotations

import gzip
import json
import random
import sys
from pathlib import Path

REPO = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(REPO / "tests" / "golden"))
import make_golden as mg  # noqa: E402

from network_solver.preprocessing import preprocess_problem, translate_result  # noqa: E402
from network_solver.data import FlowResult as RefFlowResult  # noqa: E402

from network_flow_solver_b200 import generators as gen  # noqa: E402


def decorated(n, m, seed, dup=12, chains=10, pendants=6, unbounded=0.1):
    """netgen-like instance + parallel duplicates + arcs split into chains through new zero-supply nodes +
    dangling zero-supply nodes."""
    rng = random.Random(seed)
    p = gen.to_network_problem(gen.netgen_like(n, m, n_sources=4, n_sinks=4, cost_max=50, cap_max=40, seed=seed))
    nodes = [{"id": x.id, "supply": x.supply} for x in p.nodes.values()]
    arcs = [{"tail": x.tail, "head": x.head, "capacity": x.capacity, "cost": x.cost, "lower": 0.0} for x in p.arcs]
    for a in arcs:
        if rng.random() < unbounded:
            a["capacity"] = None
    for _ in range(dup):  # exact duplicates (merged) and same-key different-cost arcs (kept)
        a = dict(rng.choice(arcs))
        if rng.random() < 0.3:
            a["cost"] = a["cost"] + 1.0
        if rng.random() < 0.3:
            a["capacity"] = None
        arcs.insert(rng.randrange(len(arcs) + 1), a)
    fresh = 0
    for _ in range(chains):
        i = rng.randrange(len(arcs))
        a = arcs.pop(i)
        hops = rng.choice([1, 1, 2, 3])
        prev = a["tail"]
        for h in range(hops):
            nid = f"x{fresh:03d}"
            fresh += 1
            nodes.insert(rng.randrange(len(nodes) + 1), {"id": nid, "supply": 0.0})
            arcs.insert(rng.randrange(len(arcs) + 1), {"tail": prev, "head": nid, "capacity": a["capacity"] if h % 2 == 0 else None,
                                                       "cost": float(rng.randrange(0, 9)), "lower": 0.0})
            prev = nid
        arcs.insert(rng.randrange(len(arcs) + 1), {"tail": prev, "head": a["head"], "capacity": a["capacity"], "cost": a["cost"], "lower": 0.0})
    for k in range(pendants):
        nid = f"y{k:03d}"
        nodes.append({"id": nid, "supply": 0.0})
        other = rng.choice(nodes[:n])["id"]
        arc = {"tail": nid, "head": other, "capacity": 5.0, "cost": 1.0, "lower": 0.0}
        if k % 2:
            arc["tail"], arc["head"] = arc["head"], arc["tail"]
        arcs.append(arc)
    return mg.ref_build(nodes, arcs, directed=True, tolerance=1e-6)


def handmade():
    out = {}
    out["parallel_mixed_capacity"] = mg.ref_build(
        [{"id": "A", "supply": 90.0}, {"id": "B", "supply": -90.0}],
        [{"tail": "A", "head": "B", "capacity": 50.0, "cost": 2.0}, {"tail": "A", "head": "B", "capacity": 30.0, "cost": 2.0},
         {"tail": "A", "head": "B", "capacity": 20.0, "cost": 2.0}, {"tail": "A", "head": "B", "capacity": 40.0, "cost": 5.0}],
        directed=True, tolerance=1e-6)
    out["parallel_with_unbounded"] = mg.ref_build(
        [{"id": "A", "supply": 12.0}, {"id": "B", "supply": -12.0}],
        [{"tail": "A", "head": "B", "capacity": 5.0, "cost": 1.0}, {"tail": "A", "head": "B", "capacity": None, "cost": 1.0},
         {"tail": "A", "head": "B", "capacity": None, "cost": 1.0}],
        directed=True, tolerance=1e-6)
    out["parallel_all_unbounded"] = mg.ref_build(
        [{"id": "A", "supply": 9.0}, {"id": "B", "supply": -9.0}],
        [{"tail": "A", "head": "B", "capacity": None, "cost": 1.0}, {"tail": "A", "head": "B", "capacity": None, "cost": 1.0},
         {"tail": "A", "head": "B", "capacity": None, "cost": 1.0}],
        directed=True, tolerance=1e-6)
    out["chain_of_three"] = mg.ref_build(
        [{"id": "s", "supply": 7.0}, {"id": "m1", "supply": 0.0}, {"id": "m2", "supply": 0.0}, {"id": "m3", "supply": 0.0},
         {"id": "t", "supply": -7.0}],
        [{"tail": "s", "head": "m1", "capacity": 10.0, "cost": 1.0}, {"tail": "m1", "head": "m2", "capacity": 8.0, "cost": 2.0},
         {"tail": "m2", "head": "m3", "capacity": None, "cost": 3.0}, {"tail": "m3", "head": "t", "capacity": 9.0, "cost": 4.0},
         {"tail": "s", "head": "t", "capacity": 3.0, "cost": 20.0}],
        directed=True, tolerance=1e-6)
    out["two_cycle_not_contracted"] = mg.ref_build(
        [{"id": "a", "supply": 4.0}, {"id": "b", "supply": 0.0}, {"id": "c", "supply": -4.0}],
        [{"tail": "a", "head": "b", "capacity": 9.0, "cost": 1.0}, {"tail": "b", "head": "a", "capacity": 9.0, "cost": 1.0},
         {"tail": "a", "head": "c", "capacity": 9.0, "cost": 3.0}],
        directed=True, tolerance=1e-6)
    out["pendants_and_isolated_pair"] = mg.ref_build(
        [{"id": "a", "supply": 5.0}, {"id": "b", "supply": -5.0}, {"id": "p", "supply": 0.0}, {"id": "q", "supply": 0.0},
         {"id": "r", "supply": 0.0}, {"id": "lonely", "supply": 0.0}],
        [{"tail": "a", "head": "b", "capacity": 9.0, "cost": 2.0}, {"tail": "a", "head": "p", "capacity": 9.0, "cost": 1.0},
         {"tail": "q", "head": "r", "capacity": 9.0, "cost": 1.0}],
        directed=True, tolerance=1e-6)
    out["nothing_to_do"] = mg.ref_build(
        [{"id": "a", "supply": 5.0}, {"id": "b", "supply": 0.0}, {"id": "c", "supply": -5.0}],
        [{"tail": "a", "head": "b", "capacity": 9.0, "cost": 2.0}, {"tail": "a", "head": "c", "capacity": 2.0, "cost": 1.0},
         {"tail": "b", "head": "c", "capacity": 9.0, "cost": 1.0}, {"tail": "c", "head": "b", "capacity": 9.0, "cost": 1.0}],
        directed=True, tolerance=1e-6)
    return out


def record(name, problem):
    pre = preprocess_problem(problem)
    changed = pre.removed_arcs > 0 or pre.removed_nodes > 0 or pre.merged_arcs > 0
    runs = []
    for opts in ({**mg.DZ}, {**mg.DX}, {}):
        inner = mg.run_reference(pre.problem, dict(opts))
        fr = RefFlowResult(objective=inner["objective"], flows={(a, b): v for a, b, v in inner["flows"]},
                           status=inner["status"], iterations=inner["iterations"], duals=dict(inner["duals"]))
        final = translate_result(fr, pre, problem) if changed else fr
        runs.append({"inner": inner,
                     "final": {"objective": final.objective, "status": final.status, "iterations": final.iterations,
                               "flows": [[k[0], k[1], v] for k, v in final.flows.items()],
                               "duals": [[k, v] for k, v in final.duals.items()]}})
    doc = {"name": name, "problem": mg.problem_to_spec(problem), "reduced": mg.problem_to_spec(pre.problem), "changed": changed,
           "stats": {"removed_arcs": pre.removed_arcs, "removed_nodes": pre.removed_nodes, "merged_arcs": pre.merged_arcs,
                     "redundant_arcs": pre.redundant_arcs, "disconnected_components": pre.disconnected_components,
                     "optimizations": pre.optimizations},
           "arc_mapping": [[i, None if k is None else list(k)] for i, k in pre.arc_mapping.items()],
           "node_mapping": [[k, v] for k, v in pre.node_mapping.items()], "runs": runs}
    print(name, doc["stats"], [(r["final"]["status"], r["final"]["iterations"], r["final"]["objective"]) for r in runs], flush=True)
    return doc


def main() -> None:
    cases = [record(k, v) for k, v in handmade().items()]
    cases.append(record("decorated_48", decorated(48, 300, 41)))
    cases.append(record("decorated_96", decorated(96, 700, 42, dup=30, chains=25, pendants=10)))
    cases.append(record("decorated_64_dense_dups", decorated(64, 400, 43, dup=80, chains=5, pendants=3, unbounded=0.3)))
    path = REPO / "tests" / "golden" / "next" / "preprocessing.json.gz"
    with gzip.GzipFile(path, "wb", mtime=0) as fh:
        fh.write(json.dumps({"cases": cases}, separators=(",", ":")).encode())
    print(f"wrote {path} ({path.stat().st_size / 1024:.1f} KiB)")


if __name__ == "__main__":
    main()
