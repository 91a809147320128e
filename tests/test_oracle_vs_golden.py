"""Pins the CPU restatement (oracle/nsx_oracle.c) to the reference: every golden run recorded by
tests/golden/make_golden.py must be reproduced bit for bit - entering-arc sequence, per-arc flows,
tree flags, node potentials, and the public FlowResult."""

import pytest

from helpers import assert_matches_reference, golden_cases, load_golden, prepare_run
from oracle import oracle


@pytest.mark.parametrize("name,idx", golden_cases())
def test_oracle_reproduces_reference(name, idx):
    doc = load_golden(name)
    run = doc["runs"][idx]
    _, cp, plan, options = prepare_run(doc, run)
    raw = oracle.solve_canonical(cp, plan.engine)
    assert_matches_reference(run, cp, raw, options)


@pytest.mark.parametrize("name,idx", golden_cases()[::3])
def test_oracle_threaded_sweep_is_order_preserving(name, idx):
    doc = load_golden(name)
    run = doc["runs"][idx]
    _, cp, plan, options = prepare_run(doc, run)
    raw = oracle.solve_canonical(cp, plan.engine, threads=4)
    assert_matches_reference(run, cp, raw, options)
