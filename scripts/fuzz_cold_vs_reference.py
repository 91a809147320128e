"""Fuzz the cold solve, the option surface and the preprocessing pass against the UNMODIFIED reference (needs
/root/reference; build container only).  Random small instances - directed / undirected, lower bounds, uncapacitated arcs,
negative costs (unbounded exits), assignment / max-flow / shortest-path structures - under random options (all pricing
strategies, loop-based Devex, fixed block sizes, iteration limits).  The reference solves each instance; this repo's host
logic + oracle + emulated device core must give the same status, iteration count, objective, flows and duals (or the same
exception), and preprocess_problem() the same reduced problem and maps.
    NUMBA_CACHE_DIR=/tmp/numba_cache [FUZZ_SCALE=4] python scripts/fuzz_cold_vs_reference.py [trials] [seed]"""
import io, logging, os, random, sys
from contextlib import redirect_stdout
sys.path.insert(0, '.'); sys.path.insert(0, 'tests'); sys.path.insert(0, '/root/reference/src')
os.environ.setdefault("NUMBA_CACHE_DIR", "/tmp/numba_cache")
logging.disable(logging.CRITICAL)
from network_solver import SolverOptions as RefOptions, build_problem as ref_build, solve_min_cost_flow as ref_solve
from network_solver.exceptions import UnboundedProblemError as RefUnbounded
from network_solver.preprocessing import preprocess_problem as ref_preprocess, preprocess_and_solve as ref_pas
from network_flow_solver_b200 import SolverConfigurationError, SolverOptions, UnboundedProblemError, build_problem
from network_flow_solver_b200.preprocessing import preprocess_problem, preprocess_and_solve
from network_flow_solver_b200 import solver as solver_module
from network_flow_solver_b200.solver import finish, prepare
from oracle import oracle
from emu import emu

# the public call's C-ABI step is served by the oracle here (no GPU in the build container); everything around it is the product's host code
solver_module._capi.solve_canonical = lambda cp, opts, out=None, warm=None: oracle.solve_canonical(cp, opts, warm=warm)
SCALE = int(os.environ.get('FUZZ_SCALE', '1'))  # node-count multiplier: >1 reaches the 50-pivot block adaptation and the 65-pivot reset cadence
trials = int(sys.argv[1]) if len(sys.argv) > 1 else 300
rng = random.Random(int(sys.argv[2]) if len(sys.argv) > 2 else 1)


def instance():
    kind = rng.random()
    n = rng.randint(3, 10) * SCALE
    ids = [f"v{i}" for i in range(n)]
    directed = True
    if kind < 0.12:  # assignment
        k = rng.randint(2, 5) * max(1, SCALE // 2)
        nodes = {f"w{i}": 1.0 for i in range(k)} | {f"j{i}": -1.0 for i in range(k)}
        arcs = [(f"w{i}", f"j{j}", 1.0, float(rng.randint(1, 9)), 0.0) for i in range(k) for j in range(k) if i == j or rng.random() < 0.7]
        return nodes, arcs, True
    nodes = {v: 0.0 for v in ids}
    if kind < 0.24:  # one unit from a source to a sink: shortest path
        s, t = rng.sample(ids, 2); nodes[s], nodes[t] = 1.0, -1.0
    elif kind < 0.34:  # one source, one sink
        s, t = rng.sample(ids, 2); q = float(rng.randint(2, 9)); nodes[s], nodes[t] = q, -q
    else:
        for _ in range(rng.randint(1, 10)):
            nodes[rng.choice(ids)] += 1.0; nodes[rng.choice(ids)] -= 1.0
        directed = rng.random() < 0.85
    total = sum(v for v in nodes.values() if v > 0) or 1.0
    uniform = rng.choice([None, 0.0, 1.0]) if kind < 0.34 and kind >= 0.24 else None
    arcs, seen = [], set()
    order = ids[:]; rng.shuffle(order)
    for a, b in zip(order, order[1:] + order[:1]):  # a ring with room for everything: feasible whatever the supplies
        arcs.append((a, b, total + 1.0, float(rng.randint(1, 9)) if uniform is None else uniform, 0.0)); seen.add((a, b))
    for _ in range(rng.randint(0, 2 * n) * (2 if SCALE > 1 else 1)):
        a, b = rng.sample(ids, 2)
        if (a, b) in seen or (not directed and (b, a) in seen):
            continue
        seen.add((a, b))
        cap = None if directed and rng.random() < 0.2 else float(rng.randint(1, int(total) + 2))
        cost = float(rng.randint(-2 if rng.random() < 0.15 else 0, 9)) if uniform is None else uniform
        lower = float(rng.randint(1, int(cap))) if directed and cap is not None and rng.random() < 0.12 and cap >= 1 else 0.0
        arcs.append((a, b, cap, cost, lower))
    for a, b, cap, cost, lower in arcs:  # keep the instance balanced after the lower-bound shift
        pass
    rng.shuffle(arcs)
    return nodes, arcs, directed


def options():
    strategy = rng.choice(["dantzig", "devex", "devex", "candidate_list", "adaptive"])
    kw = dict(pricing_strategy=strategy, explicit_pricing_strategy=rng.random() < 0.8, auto_scale=False)
    if strategy == "devex":
        if rng.random() < 0.4: kw["use_vectorized_pricing"] = False
        if rng.random() < 0.5: kw["block_size"] = rng.choice([1, 2, 3, 5, 8, 50])
    if rng.random() < 0.15: kw["ft_update_limit"] = rng.choice([1, 2, 3])
    return kw, (rng.choice([1, 2, 3, 5, 8]) if rng.random() < 0.12 else None)


def spec(p):
    return ([(k, n.supply) for k, n in p.nodes.items()], [(a.tail, a.head, a.capacity, a.cost, a.lower) for a in p.arcs])


bad = compared = skipped = 0
import collections
seen_kinds = collections.Counter()
for trial in range(trials):
    nodes, arcs, directed = instance()
    kw, limit = options()
    if rng.random() < 0.15:  # badly scaled values + the default auto_scale=True (scaling.py)
        cm, qm = rng.choice([1e-4, 1e-3, 1e3, 1e5]), rng.choice([1.0, 1e3, 1e4])
        nodes = {k: v * qm for k, v in nodes.items()}
        arcs = [(a, b, None if c is None else c * qm, w * cm, lo * qm) for a, b, c, w, lo in arcs]
        kw["auto_scale"] = True
    via_preprocessing = rng.random() < 0.15  # preprocess_and_solve(): reduce, solve, translate back
    mk = lambda build: build([{"id": k, "supply": v} for k, v in nodes.items()],
                             [{"tail": a, "head": b, "capacity": c, "cost": w, "lower": lo} for a, b, c, w, lo in arcs], directed=directed, tolerance=1e-6)
    try:
        rp, mp = mk(ref_build), mk(build_problem)
    except Exception:
        skipped += 1
        continue
    # preprocessing pass
    a, b = ref_preprocess(rp), preprocess_problem(mp)
    if (spec(a.problem), a.arc_mapping, a.node_mapping, a.removed_arcs, a.removed_nodes, a.merged_arcs, a.redundant_arcs, a.disconnected_components) != \
       (spec(b.problem), b.arc_mapping, b.node_mapping, b.removed_arcs, b.removed_nodes, b.merged_arcs, b.redundant_arcs, b.disconnected_components):
        bad += 1
        print(f"trial {trial}: preprocess_problem differs\n   nodes {nodes}\n   arcs {arcs} directed {directed}")
    # solve
    if via_preprocessing:
        try:
            with redirect_stdout(io.StringIO()):
                _, ref = ref_pas(rp, options=RefOptions(**kw), max_iterations=limit)
            want = (ref.status, ref.iterations, ref.objective, ref.flows, ref.duals)
        except RefUnbounded as exc:
            want = ("unbounded", tuple(exc.entering_arc), exc.reduced_cost)
        except Exception as exc:
            want = ("error", type(exc).__name__)
        try:
            with redirect_stdout(io.StringIO()):
                _, r = preprocess_and_solve(mp, options=SolverOptions(**kw), max_iterations=limit)  # C-ABI call -> oracle, see below
            got = (r.status, r.iterations, r.objective, r.flows, r.duals)
        except UnboundedProblemError as exc:
            got = ("unbounded", tuple(exc.entering_arc), exc.reduced_cost)
        except SolverConfigurationError:
            skipped += 1
            continue
        except Exception as exc:
            got = ("error", type(exc).__name__)
        compared += 1
        seen_kinds[("preprocess_and_solve", str(want[0]), 'directed' if directed else 'undirected')] += 1
        if got != want:
            bad += 1
            print(f"trial {trial}: preprocess_and_solve {kw} limit {limit}\n   reference {want[:3]}\n   mine      {got[:3]}\n   nodes {nodes}\n   arcs {arcs} directed {directed}")
        continue
    try:
        with redirect_stdout(io.StringIO()):
            ref = ref_solve(rp, RefOptions(**kw), max_iterations=limit)
        want = (ref.status, ref.iterations, ref.objective, ref.flows, ref.duals)
    except RefUnbounded as exc:
        want = ("unbounded", tuple(exc.entering_arc), exc.reduced_cost)
    except Exception as exc:  # InvalidProblemError etc.: the drop-in must raise the same kind
        want = ("error", type(exc).__name__)
    try:
        cp, plan, opts = prepare(mp, SolverOptions(**kw), limit, trace_capacity=1 << 14)
    except SolverConfigurationError:
        skipped += 1  # bipartite-matching structure: refused on purpose
        continue
    except Exception as exc:
        got = gote = ("error", type(exc).__name__)
    else:
        def outcome(raw):
            try:
                r = finish(cp, raw, opts, plan.scaling)
                return (r.status, r.iterations, r.objective, r.flows, r.duals)
            except UnboundedProblemError as exc:
                return ("unbounded", tuple(exc.entering_arc), exc.reduced_cost)
        got, gote = outcome(oracle.solve_canonical(cp, plan.engine)), outcome(emu.solve_canonical(cp, plan.engine))
    compared += 1
    seen_kinds[(getattr(cp, 'network_type', 'n/a') if 'cp' in dir() else 'n/a', str(want[0]), 'directed' if directed else 'undirected')] += 1
    if got != want or gote != want:
        bad += 1
        print(f"trial {trial}: {kw} limit {limit} type {getattr(cp, 'network_type', '?')}\n   reference {want[:3]}\n   oracle    {got[:3]}\n   emulated  {gote[:3]}"
              f"\n   nodes {nodes}\n   arcs {arcs} directed {directed}")
print(f"{trials} trials: {compared} compared, {skipped} skipped, {bad} disagreements")
for k, v in sorted(seen_kinds.items()):
    print("  ", v, *k)
