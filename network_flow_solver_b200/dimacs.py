"""DIMACS minimum-cost-flow (``.min``) ingest.

Two entry points:

* ``parse_dimacs_file`` / ``parse_dimacs_string`` -> ``NetworkProblem``: same accepted syntax, defaults
  and errors as the reference's parser (benchmarks/parsers/dimacs.py:45-286): ``p min N M``,
  ``n id supply`` (unlisted nodes are transshipment nodes), ``a tail head [lower] capacity cost``,
  capacity ``-1`` / ``inf`` / >= 1e15 meaning uncapacitated, ids kept as the strings "1".."N",
  tolerance 1e-6.
* ``load_dimacs_canonical`` -> ``CanonicalProblem`` straight from the text with NumPy, for files too
  large to turn into ``Arc`` objects.  It reproduces the reference's internal index space for DIMACS
  ids: nodes in LEXICOGRAPHIC order of the id strings ("1", "10", "100", ..., "2": simplex.py:149),
  arcs stably sorted by (tail id string, head id string) (simplex.py:392-395), lower bounds shifted
  out of capacities and supplies (simplex.py:398-432), perturbed costs (simplex.py:1431-1440).
  ``tests/test_dimacs.py`` checks it array for array against ``canonicalize(parse_dimacs_file(...))``.
"""

from __future__ import annotations

from pathlib import Path
from typing import Any

import numpy as np

from .canonical import (
    NET_GENERAL,
    PERTURB_EPS_BASE,
    CanonicalProblem,
    _penalty,
    canonicalize,
    perturbed_costs,
)
from .data import NetworkProblem, build_problem
from .exceptions import InvalidProblemError

_INF_CAPACITY = 1e15  # benchmarks/parsers/dimacs.py:222-224


def parse_dimacs_file(path: str | Path) -> NetworkProblem:
    path = Path(path)
    if not path.exists():
        raise FileNotFoundError(f"DIMACS file not found: {path}")
    with open(path, encoding="utf-8") as fh:
        return _parse_lines([line.rstrip("\n\r") for line in fh])


def parse_dimacs_string(content: str) -> NetworkProblem:
    return _parse_lines(content.splitlines())


def _capacity(token: str) -> float | None:
    if token == "-1" or token.lower() == "inf":
        return None
    value = float(token)
    return None if value >= _INF_CAPACITY else value


def _scan(lines: list[str]):
    """Shared tokenizer: returns (num_nodes, supplies {id: value}, arcs [(tail, head, lower, cap, cost)])."""
    num_nodes = num_arcs = None
    seen_p = False
    supplies: dict[str, float] = {}
    arcs: list[tuple[str, str, float, float | None, float]] = []
    for line_num, raw in enumerate(lines, start=1):
        line = raw.strip()
        if not line or line.startswith("c"):
            continue
        tokens = line.split()
        kind = tokens[0]
        try:
            if kind == "p":
                if seen_p:
                    raise InvalidProblemError(
                        f"Line {line_num}: Multiple problem descriptor lines found. Only one 'p min' line is allowed.")
                if len(tokens) != 4:
                    raise InvalidProblemError(
                        f"Line {line_num}: Invalid problem descriptor format. Expected 'p min <nodes> <arcs>', got: {line}")
                if tokens[1] != "min":
                    raise InvalidProblemError(
                        f"Line {line_num}: Only 'min' (minimum cost flow) problems supported. Got: {tokens[1]}")
                num_nodes, num_arcs = int(tokens[2]), int(tokens[3])
                seen_p = True
                if num_nodes <= 0:
                    raise InvalidProblemError(f"Line {line_num}: Number of nodes must be positive, got {num_nodes}")
                if num_arcs < 0:
                    raise InvalidProblemError(f"Line {line_num}: Number of arcs cannot be negative, got {num_arcs}")
            elif kind == "n":
                if not seen_p:
                    raise InvalidProblemError(
                        f"Line {line_num}: Node descriptor before problem descriptor. The 'p min' line must come first.")
                if len(tokens) != 3:
                    raise InvalidProblemError(
                        f"Line {line_num}: Invalid node descriptor format. Expected 'n <node_id> <supply>', got: {line}")
                supplies[tokens[1]] = float(tokens[2])
            elif kind == "a":
                if not seen_p:
                    raise InvalidProblemError(
                        f"Line {line_num}: Arc descriptor before problem descriptor. The 'p min' line must come first.")
                if len(tokens) == 6:
                    arcs.append((tokens[1], tokens[2], float(tokens[3]), _capacity(tokens[4]), float(tokens[5])))
                elif len(tokens) == 5:  # variant without a lower bound
                    arcs.append((tokens[1], tokens[2], 0.0, _capacity(tokens[3]), float(tokens[4])))
                else:
                    raise InvalidProblemError(
                        f"Line {line_num}: Invalid arc descriptor format. Expected 'a <tail> <head> <lower> <capacity> "
                        f"<cost>' or 'a <tail> <head> <capacity> <cost>', got: {line}")
            else:
                raise InvalidProblemError(
                    f"Line {line_num}: Unknown line type '{kind}'. Expected 'c' (comment), 'p' (problem), 'n' (node), "
                    "or 'a' (arc).")
        except (ValueError, IndexError) as exc:
            raise InvalidProblemError(f"Line {line_num}: Failed to parse line: {line}. Error: {exc}") from exc
    if not seen_p:
        raise InvalidProblemError(
            "No problem descriptor found. DIMACS file must contain a 'p min <nodes> <arcs>' line.")
    if num_arcs != len(arcs):
        raise InvalidProblemError(
            f"Arc count mismatch: problem descriptor specifies {num_arcs} arcs, but {len(arcs)} arc descriptors found.")
    valid = {str(i) for i in range(1, num_nodes + 1)}
    unexpected = {t for t, _, _, _, _ in arcs} | {h for _, h, _, _, _ in arcs}
    unexpected -= valid
    if unexpected:
        raise InvalidProblemError(
            f"Arc references node IDs outside the expected range [1, {num_nodes}]: {sorted(unexpected)}")
    return num_nodes, supplies, arcs


def _parse_lines(lines: list[str]) -> NetworkProblem:
    num_nodes, supplies, arcs = _scan(lines)
    nodes = [{"id": str(i), "supply": supplies.get(str(i), 0.0)} for i in range(1, num_nodes + 1)]
    arc_dicts: list[dict[str, Any]] = [
        {"tail": t, "head": h, "lower": lo, "capacity": cap, "cost": c} for t, h, lo, cap, c in arcs
    ]
    return build_problem(nodes=nodes, arcs=arc_dicts, directed=True, tolerance=1e-6)


def lexicographic_ranks(num_nodes: int) -> np.ndarray:
    """rank[k] = position of the id string str(k) among "1".."N" sorted as strings (rank[0] unused)."""
    ids = np.array([str(i) for i in range(1, num_nodes + 1)])
    order = np.argsort(ids, kind="stable")  # NumPy compares unicode arrays code point by code point, like Python
    rank = np.empty(num_nodes + 1, dtype=np.int64)
    rank[order + 1] = np.arange(num_nodes)
    rank[0] = -1
    return rank


def load_dimacs_canonical(path: str | Path, *, eps_base: float = PERTURB_EPS_BASE, tolerance: float = 1e-6,
                          keep_names: bool = True) -> CanonicalProblem:
    """Array-native ingest (see module docstring).  ``keep_names=False`` drops the id / arc-key lists
    (instances with tens of millions of arcs)."""
    with open(path, encoding="utf-8") as fh:
        lines = [line.rstrip("\n\r") for line in fh]
    num_nodes, supplies, arcs = _scan(lines)
    m = len(arcs)
    rank = lexicographic_ranks(num_nodes)
    t_id = np.fromiter((int(a[0]) for a in arcs), dtype=np.int64, count=m)
    h_id = np.fromiter((int(a[1]) for a in arcs), dtype=np.int64, count=m)
    lower = np.fromiter((a[2] for a in arcs), dtype=np.float64, count=m)
    cap = np.fromiter((np.inf if a[3] is None else a[3] for a in arcs), dtype=np.float64, count=m)
    cost = np.fromiter((a[4] for a in arcs), dtype=np.float64, count=m)
    bad = np.flatnonzero(~np.isinf(cap) & (cap < lower))
    if bad.size:
        i = int(bad[0])
        raise InvalidProblemError(f"Arc ({arcs[i][0]}, {arcs[i][1]}) has capacity {cap[i]} less than lower bound {lower[i]}")
    tail, head = rank[t_id], rank[h_id]
    order = np.lexsort((head, tail))  # stable sort by (tail id string, head id string)
    tail, head, lower, cap, cost = tail[order], head[order], lower[order], cap[order], cost[order]
    sup = np.zeros(num_nodes + 1, dtype=np.float64)
    for node, value in supplies.items():
        k = int(node)
        if 1 <= k <= num_nodes:
            sup[rank[k] + 1] = value
    # lower-bound shift, arc by arc in the reference's (sorted) order: simplex.py:416-428
    np.subtract.at(sup, tail + 1, lower)
    np.add.at(sup, head + 1, lower)
    total = float(np.sum(sup))
    if abs(total) > tolerance:
        raise InvalidProblemError(
            f"Supplies do not balance after lower-bound adjustment: total supply {total:.6f} exceeds tolerance {tolerance}.")
    n_nodes = num_nodes + 1
    ids_sorted = None
    keys = None
    if keep_names:
        inv = np.empty(num_nodes, dtype=np.int64)
        inv[rank[1:]] = np.arange(1, num_nodes + 1)
        ids_sorted = ["__network_simplex_root__"] + [str(int(i)) for i in inv]
        keys = [(ids_sorted[int(t) + 1], ids_sorted[int(h) + 1]) for t, h in zip(tail.tolist(), head.tolist())]
    return CanonicalProblem(
        n_nodes=n_nodes,
        tail=np.ascontiguousarray(tail + 1, dtype=np.int32),
        head=np.ascontiguousarray(head + 1, dtype=np.int32),
        orig_cost=np.ascontiguousarray(cost),
        pert_cost=perturbed_costs(cost, eps_base),
        upper=np.where(np.isinf(cap), np.inf, cap - lower),
        shift=np.ascontiguousarray(lower),
        supply=sup,
        penalty=_penalty(cost, n_nodes),
        network_type=NET_GENERAL,
        node_ids=ids_sorted,
        arc_keys=keys,
        n_supply_nodes=int(np.sum(sup > tolerance)),
        n_demand_nodes=int(np.sum(sup < -tolerance)),
    )


def write_dimacs(path: str | Path, num_nodes: int, supplies: dict[int, float], arcs, comment: str = "") -> None:
    """Write a ``.min`` file; ``arcs`` = iterable of (tail, head, lower, capacity or None, cost), 1-based ids."""
    with open(path, "w", encoding="utf-8") as fh:
        if comment:
            fh.write(f"c {comment}\n")
        arcs = list(arcs)
        fh.write(f"p min {num_nodes} {len(arcs)}\n")
        for node in sorted(supplies):
            if supplies[node] != 0:
                fh.write(f"n {node} {supplies[node]:g}\n")
        for t, h, lo, cap, c in arcs:
            fh.write(f"a {t} {h} {lo:g} {'-1' if cap is None else format(cap, 'g')} {c:g}\n")


__all__ = ["parse_dimacs_file", "parse_dimacs_string", "load_dimacs_canonical", "lexicographic_ranks", "write_dimacs",
           "canonicalize"]
