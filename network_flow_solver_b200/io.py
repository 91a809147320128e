"""JSON problem / result files in the reference's schema (reference: src/network_solver/io.py:33-69).

Problem document: ``{"directed": bool, "tolerance": float, "nodes": [{"id", "supply"}...],
"edges" | "arcs": [{"tail", "head", "capacity"?, "cost"?, "lower"?}...]}``; result document:
status, objective, iterations, flows as a list of ``{"tail", "head", "flow"}`` sorted by key, duals sorted by node.
"""

from __future__ import annotations

import json
from pathlib import Path
from typing import Any

from .data import FlowResult, NetworkProblem, build_problem
from .exceptions import InvalidProblemError

_ARC_DEFAULTS = {"capacity": None, "cost": 0.0, "lower": 0.0}


def _arc_record(edge: dict[str, Any]) -> dict[str, Any]:
    missing = [field for field in ("tail", "head") if field not in edge]
    if missing:
        raise InvalidProblemError(
            f"Invalid edge specification: {edge}. Each edge must have 'tail' and 'head' fields."
        )
    record = {"tail": edge["tail"], "head": edge["head"]}
    record.update({field: edge.get(field, default) for field, default in _ARC_DEFAULTS.items()})
    return record


def load_problem(path) -> NetworkProblem:
    document = json.loads(Path(path).read_text(encoding="utf-8"))
    nodes, edges = document.get("nodes"), document.get("edges") or document.get("arcs")
    if not (isinstance(nodes, list) and isinstance(edges, list)):
        raise InvalidProblemError(
            "Invalid problem format: JSON must include 'nodes' and 'edges' (or 'arcs') arrays."
        )
    return build_problem(
        nodes=nodes,
        arcs=[_arc_record(edge) for edge in edges],
        directed=bool(document.get("directed", True)),
        tolerance=float(document.get("tolerance", 1e-3)),
    )


def save_result(path, result: FlowResult) -> None:
    flows = [{"tail": key[0], "head": key[1], "flow": value} for key, value in sorted(result.flows.items())]
    document = {
        "status": result.status,
        "objective": result.objective,
        "iterations": result.iterations,
        "flows": flows,
        "duals": {node: result.duals[node] for node in sorted(result.duals)},
    }
    Path(path).write_text(json.dumps(document, indent=2, sort_keys=False), encoding="utf-8")
