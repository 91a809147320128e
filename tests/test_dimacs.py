"""DIMACS `.min` ingest against vectors recorded from the unmodified reference
(tests/golden/make_dimacs_golden.py): parser output, error behaviour, the array-native loader, and - through the
oracle on CPU and the CUDA engine on GPU - the reference's pivots and optima (111 / 95 known optima included)."""

import gzip
import json
from pathlib import Path

import numpy as np
import pytest

from helpers import assert_matches_reference
from network_flow_solver_b200 import SolverOptions, _capi
from network_flow_solver_b200.canonical import canonicalize
from network_flow_solver_b200.dimacs import (lexicographic_ranks, load_dimacs_canonical, parse_dimacs_file,
                                             parse_dimacs_string, write_dimacs)
from network_flow_solver_b200.exceptions import InvalidProblemError
from network_flow_solver_b200.solver import prepare
from oracle import oracle

DOC = json.loads(gzip.open(Path(__file__).resolve().parent / "golden" / "dimacs" / "dimacs.json.gz", "rb").read().decode())
CASES = {c["name"]: c for c in DOC["cases"]}


@pytest.mark.parametrize("name", sorted(CASES))
def test_parser_reproduces_the_reference_parser(name):
    case = CASES[name]
    problem = parse_dimacs_string(case["text"])
    spec = case["problem"]
    assert problem.directed == spec["directed"] and problem.tolerance == spec["tolerance"]
    assert [[n.id, n.supply] for n in problem.nodes.values()] == spec["nodes"]
    assert [[a.tail, a.head, a.capacity, a.cost, a.lower] for a in problem.arcs] == spec["arcs"]


@pytest.mark.parametrize("k", range(len(DOC["errors"])))
def test_parser_rejects_what_the_reference_rejects(k):
    rec = DOC["errors"][k]
    if rec["error"] is None:
        parse_dimacs_string(rec["text"])
    else:
        with pytest.raises(InvalidProblemError) as info:
            parse_dimacs_string(rec["text"])
        assert str(info.value) == rec["error"]


def test_lexicographic_rank_is_python_string_order():
    for n in (9, 10, 11, 99, 100, 101, 1234):
        ids = sorted(str(i) for i in range(1, n + 1))
        rank = lexicographic_ranks(n)
        assert [ids[rank[i]] for i in range(1, n + 1)] == [str(i) for i in range(1, n + 1)]


@pytest.mark.parametrize("name", sorted(CASES))
def test_array_loader_equals_object_path(name, tmp_path):
    path = tmp_path / f"{name}.min"
    path.write_text(CASES[name]["text"])
    a = canonicalize(parse_dimacs_file(path), 1e-6)
    b = load_dimacs_canonical(path)
    for field in ("tail", "head", "orig_cost", "pert_cost", "upper", "shift", "supply"):
        np.testing.assert_array_equal(getattr(a, field), getattr(b, field), err_msg=field)
    assert (a.n_nodes, a.penalty, a.node_ids, a.arc_keys) == (b.n_nodes, b.penalty, b.node_ids, b.arc_keys)


def test_write_then_parse_round_trip(tmp_path):
    arcs = [(1, 2, 0, 10, 3), (2, 3, 1, None, 2), (1, 3, 0, 4, 9)]
    write_dimacs(tmp_path / "x.min", 3, {1: 5, 3: -5}, arcs, comment="round trip")
    p = parse_dimacs_file(tmp_path / "x.min")
    assert [(a.tail, a.head, a.lower, a.capacity, a.cost) for a in p.arcs] == [
        ("1", "2", 0.0, 10.0, 3.0), ("2", "3", 1.0, None, 2.0), ("1", "3", 0.0, 4.0, 9.0)]
    assert [n.supply for n in p.nodes.values()] == [5.0, 0.0, -5.0]


def solvable_runs():
    out = []
    for name, case in sorted(CASES.items()):
        for i, run in enumerate(case["runs"]):
            out.append((name, i))
    return out


STRUCTURE_RULES = ("assignment", "max_flow", "shortest_path")  # the reference runs a structure-specific rule there


def run_case(name, i, solve):
    case, run = CASES[name], CASES[name]["runs"][i]
    problem = parse_dimacs_string(case["text"])
    options = SolverOptions(**run["options"])
    cp, plan, options = prepare(problem, options, trace_capacity=1 << 16)
    assert (plan.engine.row_scan_first >= _capi.SPECIAL_ASSIGNMENT) == (run["network_type"] in STRUCTURE_RULES)
    assert_matches_reference(run, cp, solve(cp, plan.engine), options)


@pytest.mark.parametrize("name,i", solvable_runs())
def test_oracle_reproduces_reference_on_dimacs_instances(name, i):
    run_case(name, i, lambda cp, eng: oracle.solve_canonical(cp, eng))


@pytest.mark.parametrize("name,i", solvable_runs())
def test_emulated_device_core_reproduces_reference_on_dimacs_instances(name, i):
    from emu import emu

    run_case(name, i, lambda cp, eng: emu.solve_canonical(cp, eng))


@pytest.mark.gpu
@pytest.mark.parametrize("name,i", [r for r in solvable_runs() if CASES[r[0]]["runs"][r[1]]["network_type"] not in STRUCTURE_RULES])
def test_engine_reproduces_reference_on_dimacs_instances(name, i):
    run_case(name, i, lambda cp, eng: _capi.solve_canonical(cp, eng))


@pytest.mark.gpu
@pytest.mark.timeout(300, method="thread")
@pytest.mark.parametrize("name,i", [r for r in solvable_runs() if CASES[r[0]]["runs"][r[1]]["network_type"] in STRUCTURE_RULES])
def test_engine_reproduces_reference_on_dimacs_instances_with_structure_rules(name, i):
    run_case(name, i, lambda cp, eng: _capi.solve_canonical(cp, eng))


@pytest.mark.gpu
def test_engine_solves_known_optima_from_dimacs_files(tmp_path):
    from network_flow_solver_b200 import solve_min_cost_flow

    for name, optimum in (("ref_tiny_transportation", 111.0), ("ref_small_transshipment", 95.0)):
        path = tmp_path / f"{name}.min"
        path.write_text(CASES[name]["text"])
        result = solve_min_cost_flow(parse_dimacs_file(path), SolverOptions(
            pricing_strategy="devex", explicit_pricing_strategy=True, auto_scale=False))
        assert result.status == "optimal" and result.objective == optimum
