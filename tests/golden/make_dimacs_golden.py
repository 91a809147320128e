#!/usr/bin/env python
"""Golden vectors for the DIMACS ingest, produced by the UNMODIFIED
reference in the build container:

    cd /tmp/refcopy && NUMBA_CACHE_DIR=/tmp/numba_cache python /root/repo/tests/golden/make_dimacs_golden.py

(the reference's parser imports `src.network_solver`, so the working directory must be a copy of the
reference root).  For every case the DIMACS text, what the reference's parser made of it (nodes, arcs)
and - where the instance is inside the accelerated scope - a recorded reference solve are stored in
tests/golden/dimacs/dimacs.json.gz.  Cases: the reference's three hand-made `.min` fixtures
(benchmarks/problems/generated, known optima 111 / 95 / 11: benchmarks/metadata/known_solutions.json:29-55)
and generated files whose ids reach three digits (lexicographic id order matters), with lower bounds,
uncapacitated arcs in all three spellings and the 4-field arc variant.
"""

from __future__ import annotations

import gzip
import json
import os
import sys
import tempfile
from pathlib import Path

REPO = Path(__file__).resolve().parents[2]
REF = Path(os.environ.get("NSX_REFERENCE_COPY", os.getcwd()))
sys.path.insert(0, str(REPO))
sys.path.insert(0, str(REPO / "tests" / "golden"))
sys.path.insert(0, str(REF))
sys.path.insert(0, str(REF / "src"))
os.environ.setdefault("NUMBA_CACHE_DIR", "/tmp/numba_cache")

import numpy as np  # noqa: E402
from benchmarks.parsers.dimacs import parse_dimacs_string as ref_parse  # noqa: E402
from src.network_solver.exceptions import InvalidProblemError as RefInvalid  # noqa: E402

import make_golden as mg  # noqa: E402  (run_reference, problem_to_spec)


def synthetic(n: int, m: int, seed: int, four_field: bool) -> str:
    rng = np.random.default_rng(seed)
    lines = [f"c generated instance n={n} m={m} seed={seed}", f"p min {n} {m}"]
    src, dst = 1, n
    lines += [f"n {src} 40", f"n {dst} -40"]
    arcs = [(i, i + 1) for i in range(1, n)]  # a chain keeps it feasible
    while len(arcs) < m:
        t, h = int(rng.integers(1, n + 1)), int(rng.integers(1, n + 1))
        if t != h and (t, h) not in arcs:
            arcs.append((t, h))
    for k, (t, h) in enumerate(arcs):
        cost = int(rng.integers(1, 60))
        if k < n - 1:
            cap = ["-1", "inf", "1e15", "2000000000000000"][k % 4]  # the uncapacitated spellings
        else:
            cap = str(int(rng.integers(5, 60)))
        lower = 2 if (k % 7 == 3 and k >= n - 1) else 0
        lines.append(f"a {t} {h} {cap} {cost}" if (four_field and lower == 0) else f"a {t} {h} {lower} {cap} {cost}")
    return "\n".join(lines) + "\n"


def main() -> None:
    cases = []
    for stem in ("tiny_transportation", "small_transshipment", "simple_assignment"):
        cases.append((f"ref_{stem}", (REF / "benchmarks" / "problems" / "generated" / f"{stem}.min").read_text()))
    cases.append(("gen_n30", synthetic(30, 90, 1, False)))
    cases.append(("gen_n120_4field", synthetic(120, 420, 2, True)))
    out = []
    for name, text in cases:
        problem = ref_parse(text)
        rec = {"name": name, "text": text, "problem": mg.problem_to_spec(problem), "runs": []}
        for opts in (mg.DX, mg.DZ):
            try:
                rec["runs"].append(mg.run_reference(problem, dict(opts)))
            except Exception as exc:  # e.g. nothing; keep the generator robust
                rec["runs"].append({"options": dict(opts), "error": repr(exc)})
        out.append(rec)
        print(name, [r.get("status") for r in rec["runs"]], [r.get("objective") for r in rec["runs"]],
              [r.get("network_type") for r in rec["runs"]])
    bad = ["p max 2 1\na 1 2 0 1 1\n", "p min 2 2\na 1 2 0 1 1\n", "a 1 2 0 1 1\np min 2 1\n", "p min 2 1\na 1 3 0 1 1\n",
           "p min 2 1\nx 1 2\na 1 2 0 1 1\n", "p min 2 1\na 1 2 0 one 1\n"]
    errors = []
    for text in bad:
        try:
            ref_parse(text)
            errors.append({"text": text, "error": None})
        except RefInvalid as exc:
            errors.append({"text": text, "error": str(exc)})
    raw = json.dumps({"cases": out, "errors": errors}, separators=(",", ":")).encode()
    path = REPO / "tests" / "golden" / "dimacs" / "dimacs.json.gz"
    with gzip.GzipFile(path, "wb", mtime=0) as fh:
        fh.write(raw)
    print(f"wrote {path} ({path.stat().st_size / 1024:.1f} KiB)")


if __name__ == "__main__":
    main()
