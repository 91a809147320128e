"""Fuzz the warm-start row against the UNMODIFIED reference (needs /root/reference; build container only):
random small capacitated instances, solved cold by the reference, edited (costs / capacities / supplies), re-solved by the
reference with the previous basis - and by this repo's host logic + oracle.  Reports every disagreement in
accepted/rejected, status, iteration count, entering-arc sequence, objective or flows.
    NUMBA_CACHE_DIR=/tmp/numba_cache python scripts/fuzz_warm_start_vs_reference.py [trials] [seed]"""
import io, logging, os, random, sys
from contextlib import redirect_stdout
sys.path.insert(0, '.'); sys.path.insert(0, '/root/reference/src')
os.environ.setdefault("NUMBA_CACHE_DIR", "/tmp/numba_cache")
logging.disable(logging.CRITICAL)
from network_solver import SolverOptions as RefOptions, build_problem as ref_build
from network_solver.simplex import NetworkSimplex
from network_flow_solver_b200 import Basis, SolverOptions, build_problem
from network_flow_solver_b200.solver import finish, prepare
from network_flow_solver_b200.warm_start import apply_basis
from oracle import oracle
sys.path.insert(0, 'tests')
from emu import emu

SCALE = int(os.environ.get('FUZZ_SCALE', '1'))  # node-count multiplier
trials = int(sys.argv[1]) if len(sys.argv) > 1 else 200
rng = random.Random(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
STRATEGIES = ["dantzig", "devex", "candidate_list", "adaptive"]
bad = applied = rejected = errors = 0
for trial in range(trials):
    n = rng.randint(4, 9) * SCALE
    ids = [f"v{i}" for i in range(n)]
    total = rng.randint(2, 12)
    supply = {v: 0 for v in ids}
    for _ in range(total):
        supply[rng.choice(ids[: n // 2])] += 1
        supply[rng.choice(ids[n // 2:])] -= 1
    arcs, seen = [], set()
    for i in range(n - 1):  # a path with room for everything keeps it feasible in one direction ...
        arcs.append([ids[i], ids[i + 1], float(total), float(rng.randint(1, 9))]); seen.add((i, i + 1))
    for _ in range(rng.randint(n, 3 * n)):
        a, b = rng.randrange(n), rng.randrange(n)
        if a != b and (a, b) not in seen:
            seen.add((a, b)); arcs.append([ids[a], ids[b], float(rng.randint(1, total)), float(rng.randint(0, 9))])
    rng.shuffle(arcs)
    edited = [list(x) for x in arcs]
    for _ in range(rng.randint(1, 4) * SCALE):
        k = rng.randrange(len(edited))
        what = rng.random()
        if what < 0.5: edited[k][3] = float(rng.randint(0, 9))
        elif what < 0.8: edited[k][2] = float(max(1, edited[k][2] + rng.randint(-2, 3)))
        else:
            u, w = rng.sample(ids, 2); 
            if supply[u] != 0 or supply[w] != 0: pass
    strategy = rng.choice(STRATEGIES)
    kw = dict(pricing_strategy=strategy, explicit_pricing_strategy=strategy != "adaptive", auto_scale=False)
    mk = lambda build, arcs_: build([{"id": v, "supply": float(supply[v])} for v in ids],
                                    [{"tail": a, "head": b, "capacity": c, "cost": w} for a, b, c, w in arcs_], directed=True, tolerance=1e-6)
    try:
        with redirect_stdout(io.StringIO()):
            first = NetworkSimplex(mk(ref_build, arcs), RefOptions(**kw)).solve()
            if first.status != "optimal" or first.basis is None:
                continue
            solver = NetworkSimplex(mk(ref_build, edited), RefOptions(**kw))
            trace, pivot, seen_apply = [], solver._pivot, {}
            apply = solver._apply_warm_start_basis
            solver._pivot = lambda a, d: (trace.append(int(a) * 2 + (1 if d < 0 else 0)), pivot(a, d))[1]
            solver._apply_warm_start_basis = lambda b: seen_apply.setdefault("ok", apply(b))
            ref = solver.solve(warm_start_basis=first.basis)
    except RuntimeError:
        errors += 1
        continue
    basis = Basis(tree_arcs=set(first.basis.tree_arcs), arc_flows=dict(first.basis.arc_flows))
    cp, plan, options = prepare(mk(build_problem, edited), SolverOptions(**kw), trace_capacity=1 << 14)
    if plan.engine.row_scan_first >= 2:
        continue
    warm = apply_basis(cp, basis, options.tolerance)
    applied += warm is not None; rejected += warm is None
    raw = oracle.solve_canonical(cp, plan.engine, warm=warm)
    mine = finish(cp, raw, options)
    dev = emu.solve_canonical(cp, plan.engine, warm=warm)  # the device pivot source, emulated on the host
    if dev.status != raw.status or dev.trace.tolist() != raw.trace.tolist() or dev.flow.tolist() != raw.flow.tolist():
        bad += 1
        print(f"trial {trial} [{strategy}]: emulated device core and oracle differ: status {dev.status}/{raw.status}")
    same = ((warm is not None) == bool(seen_apply["ok"]) and mine.status == ref.status and mine.iterations == ref.iterations
            and raw.trace.tolist() == trace and mine.objective == ref.objective and mine.flows == ref.flows)
    if not same:
        bad += 1
        print(f"trial {trial} [{strategy}]: applied {warm is not None}/{seen_apply['ok']} status {mine.status}/{ref.status} "
              f"its {mine.iterations}/{ref.iterations} objective {mine.objective}/{ref.objective} trace_equal {raw.trace.tolist() == trace}")
        print("   nodes", supply, "\n   arcs", arcs, "\n   edited", edited)
print(f"{trials} trials: {applied} accepted, {rejected} rejected, {errors} reference crashes, {bad} disagreements")
