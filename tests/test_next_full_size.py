"""Full-size workloads of the rows added after the hardware budget of round 1 (structure-specific rules, loop-based Devex):
the oracle was run once per workload (scripts/oracle_full.py <name> 8 tests/golden/full_next) and its status, pivot counts,
objective and the SHA-256 of its entering-arc trace / flows / potentials / arc states are committed.  CPU: the emulated
device core reproduces the smaller records; GPU: the CUDA engine must reproduce every record bit for bit and end in a
state that satisfies the optimality conditions."""

import json
from pathlib import Path

import pytest

from emu import emu
from network_flow_solver_b200 import _capi
from network_flow_solver_b200.solver import objective_value
from network_flow_solver_b200.workloads import WORKLOADS
from test_gpu_full_size import check_optimality, sha

FULL = Path(__file__).resolve().parent / "golden" / "full_next"
NAMES = sorted(p.stem for p in FULL.glob("*.json"))


def check(name, solve):
    want = json.loads((FULL / f"{name}.json").read_text())
    wl = WORKLOADS[name]
    cp = wl.canonical(0)
    r = solve(cp, wl.engine_options(cp, trace_capacity=1 << 24))
    assert r.status == want["status"] and r.iterations == want["iterations"]
    assert r.phase1_iterations == want["phase1"] and r.degenerate_pivots == want["degenerate"]
    assert sha(r.trace) == want["trace_sha"], "entering-arc sequence differs from the oracle's"
    assert sha(r.flow) == want["flow_sha"] and sha(r.potential) == want["pi_sha"] and sha(r.state) == want["state_sha"]
    assert objective_value(cp, r) == want["objective"]
    if want["status"] == _capi.STATUS_OPTIMAL:
        check_optimality(cp, r)


@pytest.mark.slow
@pytest.mark.parametrize("name", [n for n in NAMES if n in ("assignment_192", "max_flow_2e12", "netgen_2e13_devex_loop")])
def test_emulated_device_core_reproduces_the_record(name):
    check(name, emu.solve_canonical)


@pytest.mark.gpu
@pytest.mark.timeout(1800, method="thread")
@pytest.mark.parametrize("name", NAMES)
def test_engine_reproduces_the_record(name):
    check(name, _capi.solve_canonical)
