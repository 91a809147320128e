"""Parity at BASELINE.json's full sizes.  The reference itself cannot run these (dense (N-1)^2 bases,
SURVEY.md 8c); the oracle - pinned to the reference on small instances - was run once per workload
(scripts/oracle_full.py, 3 min .. 44 min of CPU each) and its status / pivot count / objective and the
SHA-256 of its entering-arc trace, flows, potentials and arc states are committed under
tests/golden/full/.  The engine must reproduce them bit for bit, and its solution must satisfy the
size-independent optimality conditions of min-cost flow."""

import hashlib
import json
from pathlib import Path

import numpy as np
import pytest

from network_flow_solver_b200 import _capi
from network_flow_solver_b200.solver import objective_value
from network_flow_solver_b200.workloads import WORKLOADS

pytestmark = pytest.mark.gpu
FULL = Path(__file__).resolve().parent / "golden" / "full"


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def check_optimality(cp, r, tol=1e-6):
    m, n = cp.n_arcs, cp.n_nodes
    f = r.flow[:m]
    assert np.all(f >= 0) and np.all(f <= cp.upper + tol)
    net = np.bincount(cp.tail, weights=f, minlength=n) - np.bincount(cp.head, weights=f, minlength=n)
    np.testing.assert_allclose(net[1:], cp.supply[1:], atol=1e-6)  # conservation with zero artificial flow
    assert np.all(r.flow[m:] <= tol)
    rc = (cp.pert_cost + r.potential[cp.tail]) - r.potential[cp.head]
    in_tree = (r.state[:m] & _capi.ARC_IN_TREE) != 0
    can_fwd = (r.state[:m] & _capi.ARC_CAN_FWD) != 0
    can_bwd = (r.state[:m] & _capi.ARC_CAN_BWD) != 0
    assert not np.any(~in_tree & can_fwd & (rc < -tol)), "an arc with residual capacity still has negative reduced cost"
    assert not np.any(~in_tree & can_bwd & (rc > tol)), "an arc carrying flow still has positive reduced cost"
    assert int(in_tree.sum()) + int((r.state[m:] & _capi.ARC_IN_TREE).astype(np.int64).sum()) == n - 1  # a spanning tree


@pytest.mark.parametrize("name", sorted(p.stem for p in FULL.glob("*.json") if "_prefix" not in p.stem))
def test_full_size_workload_matches_oracle_record(name):
    want = json.loads((FULL / f"{name}.json").read_text())
    wl = WORKLOADS[name]
    cp = wl.canonical(0)
    r = _capi.solve_canonical(cp, wl.engine_options(cp, trace_capacity=1 << 24))
    assert r.status == want["status"] and r.iterations == want["iterations"]
    assert r.phase1_iterations == want["phase1"] and r.degenerate_pivots == want["degenerate"]
    assert sha(r.trace) == want["trace_sha"], "entering-arc sequence differs from the oracle's"
    assert sha(r.flow) == want["flow_sha"] and sha(r.potential) == want["pi_sha"] and sha(r.state) == want["state_sha"]
    assert objective_value(cp, r) == want["objective"]
    check_optimality(cp, r)


_CONFIG5 = {}


def _config5_instance(name):
    """netgen_2e20_dantzig / _devex are the same arcs under two pricing rules: build the 2^26-arc instance once."""
    family = name.rsplit("_", 1)[0]
    if family not in _CONFIG5:
        _CONFIG5.clear()
        _CONFIG5[family] = WORKLOADS[name].canonical(0)
    return _CONFIG5[family]


@pytest.mark.parametrize("record", sorted(p.stem for p in FULL.glob("*_prefix*.json")))
def test_config5_prefix_matches_oracle_record(record):
    """BASELINE config 5 (2^20 nodes / 2^26 arcs): a full oracle solve takes days, so the oracle ran the first P pivots
    (scripts/oracle_prefix.py, max_iterations = P) and the engine - ONE GPU, star pricing - must reach the same state:
    entering-arc sequence, flows, potentials and arc states after exactly P pivots, bit for bit."""
    want = json.loads((FULL / f"{record}.json").read_text())
    name, pivots = want["workload"], want["max_iterations"]
    wl = WORKLOADS[name]
    cp = _config5_instance(name)
    r = _capi.solve_canonical(cp, wl.engine_options(cp, trace_capacity=pivots, max_iterations=pivots))
    assert r.status == want["status"] and r.iterations == want["iterations"]
    assert r.phase1_iterations == want["phase1"] and r.degenerate_pivots == want["degenerate"]
    assert r.tree_updates == want["tree_updates"] and r.weight_resets == want["weight_resets"]
    assert r.final_block_size == want["final_block_size"]
    for k, h in want["trace_sha_at"].items():
        assert sha(r.trace[: int(k)]) == h, f"entering-arc sequence differs from the oracle's within the first {k} pivots"
    assert sha(r.flow) == want["flow_sha"] and sha(r.potential) == want["pi_sha"] and sha(r.state) == want["state_sha"]
