"""JSON problem / result files, same schema as the reference (io.py:33-69)."""

from __future__ import annotations

import json
from pathlib import Path

from .data import FlowResult, NetworkProblem, build_problem
from .exceptions import InvalidProblemError


def load_problem(path) -> NetworkProblem:
    with Path(path).open("r", encoding="utf-8") as fh:
        doc = json.load(fh)
    nodes = doc.get("nodes")
    edges = doc.get("edges") or doc.get("arcs")
    if not isinstance(nodes, list) or not isinstance(edges, list):
        raise InvalidProblemError(
            "Invalid problem format: JSON must include 'nodes' and 'edges' (or 'arcs') arrays."
        )
    arcs = []
    for e in edges:
        if "tail" not in e or "head" not in e:
            raise InvalidProblemError(
                f"Invalid edge specification: {e}. Each edge must have 'tail' and 'head' fields."
            )
        arcs.append(
            {
                "tail": e["tail"],
                "head": e["head"],
                "capacity": e.get("capacity"),
                "cost": e.get("cost", 0.0),
                "lower": e.get("lower", 0.0),
            }
        )
    return build_problem(
        nodes=nodes,
        arcs=arcs,
        directed=bool(doc.get("directed", True)),
        tolerance=float(doc.get("tolerance", 1e-3)),
    )


def save_result(path, result: FlowResult) -> None:
    doc = {
        "status": result.status,
        "objective": result.objective,
        "iterations": result.iterations,
        "flows": [
            {"tail": t, "head": h, "flow": f} for (t, h), f in sorted(result.flows.items())
        ],
        "duals": dict(sorted(result.duals.items())),
    }
    with Path(path).open("w", encoding="utf-8") as fh:
        json.dump(doc, fh, indent=2, sort_keys=False)
