"""Device code of the arc-sharded path (nsx_solve_sharded: tile interleaving over all ranks' sweep CTAs, the mailbox
exchange inside the resident kernel, the identical pivot on every rank) on ONE GPU: the ranks are host threads whose
resident kernels run side by side on disjoint SMs (NSX_GRID bounds each grid) and exchange through mailboxes in the same
device's memory - the same loads / stores / scopes as over NVLink, only the wire differs.  Every rank must return what a
single-GPU solve returns, bit for bit ("every GPU applies the identical pivot").  Also covered: the
deadline and the abort word that end a wait for a peer that never comes."""

import hashlib
import threading
import time

import numpy as np
import pytest

from network_flow_solver_b200 import DeviceEngineError, _capi
from network_flow_solver_b200 import generators as gen
from network_flow_solver_b200.canonical import initial_block_size

pytestmark = pytest.mark.gpu


def sig(r):
    return (r.status, r.iterations, hashlib.sha256(r.trace.tobytes() + r.flow.tobytes() + r.potential.tobytes() + r.state.tobytes()).hexdigest())


def instance(kind):
    if kind == "transport":
        return gen.transportation(320, 320, cost_max=100, seed=11).canonical(eps_base=0.0)  # 102 400 arcs: grid-wide sweeps
    return gen.netgen_like(4096, 1 << 17, n_sources=32, n_sinks=32, seed=7).canonical()


def options(cp, pricing, **kw):
    m = cp.n_arcs
    o = dict(pricing=pricing, row_scan_first=_capi.SPECIAL_ROW_SCAN if cp.network_type == "transportation" else 0,
             block_size=initial_block_size(m), auto_block=True, ft_update_limit=64, max_iterations=20 * (m + cp.n_nodes),
             tolerance=1e-6, trace_capacity=1 << 20, spin_timeout_ms=20000)
    o.update(kw)
    return _capi.EngineOptions(**o)


class Boxes:
    """`world` mailboxes on device 0, reset by the calling thread before the ranks start."""

    def __init__(self, world):
        self.ptrs = [_capi.mailbox_create(0)[0] for _ in range(world)]
        self.reset()

    def reset(self):
        for p in self.ptrs:
            _capi.mailbox_reset(0, p)

    def close(self):
        for p in self.ptrs:
            _capi.mailbox_close(0, p, True)


def run_ranks(cp, opts, world, boxes):
    out, err = [None] * world, [None] * world

    def work(r):
        try:
            out[r] = _capi.solve_sharded(cp, opts, r, world, boxes.ptrs)
        except Exception as exc:  # noqa: BLE001
            err[r] = exc

    threads = [threading.Thread(target=work, args=(r,)) for r in range(world)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=180)
    assert not any(t.is_alive() for t in threads), "a rank did not return"
    return out, err


@pytest.mark.parametrize("kind,pricing,world", [("netgen", _capi.PRICING_DANTZIG, 2), ("netgen", _capi.PRICING_DEVEX, 2),
                                                ("netgen", _capi.PRICING_DEVEX, 3), ("transport", _capi.PRICING_DANTZIG, 2),
                                                ("netgen", _capi.PRICING_CANDIDATE_LIST, 2)])
def test_every_rank_returns_the_single_gpu_solve(monkeypatch, kind, pricing, world):
    monkeypatch.setenv("NSX_GRID", "24")  # 1 pivot CTA + 23 sweep CTAs per rank: `world` kernels fit the GPU side by side
    monkeypatch.setenv("NSX_LAUNCH_PLAIN", str(world))  # ordinary launches, all ranks set up before any of them launches
    cp = instance(kind)
    opts = options(cp, pricing)
    single = _capi.solve_canonical(cp, opts)
    assert single.status == _capi.STATUS_OPTIMAL
    boxes = Boxes(world)
    try:
        out, err = run_ranks(cp, opts, world, boxes)
        assert err == [None] * world, err
        for r in range(world):
            assert sig(out[r]) == sig(single), f"rank {r} differs from the single-GPU solve"
            assert out[r].stats["sweeps"] == single.stats["sweeps"]
        assert out[0].timing["exchange_ms"] > 0.0
    finally:
        boxes.close()


def test_world_of_one_is_the_single_gpu_solve():
    cp = instance("netgen")
    opts = options(cp, _capi.PRICING_DEVEX)
    boxes = Boxes(1)
    try:
        assert sig(_capi.solve_sharded(cp, opts, 0, 1, boxes.ptrs)) == sig(_capi.solve_canonical(cp, opts))
    finally:
        boxes.close()


def test_a_peer_that_never_comes_ends_in_an_error_not_a_hang(monkeypatch):
    monkeypatch.setenv("NSX_GRID", "24"); monkeypatch.setenv("NSX_LAUNCH_PLAIN", "1")
    cp = instance("netgen")
    opts = options(cp, _capi.PRICING_DANTZIG, spin_timeout_ms=400)
    boxes = Boxes(2)
    try:
        t0 = time.perf_counter()
        with pytest.raises(DeviceEngineError, match="peer GPU did not deliver"):
            _capi.solve_sharded(cp, opts, 0, 2, boxes.ptrs)  # rank 1 is never started
        assert time.perf_counter() - t0 < 30.0
        # the engine is usable afterwards
        assert _capi.solve_canonical(cp, options(cp, _capi.PRICING_DANTZIG)).status == _capi.STATUS_OPTIMAL
    finally:
        boxes.close()


def test_abort_word_releases_a_waiting_rank(monkeypatch):
    monkeypatch.setenv("NSX_GRID", "24"); monkeypatch.setenv("NSX_LAUNCH_PLAIN", "1")
    cp = instance("netgen")
    opts = options(cp, _capi.PRICING_DANTZIG, spin_timeout_ms=60000)
    boxes = Boxes(2)
    try:
        err = []

        def work():
            try:
                _capi.solve_sharded(cp, opts, 0, 2, boxes.ptrs)
            except Exception as exc:  # noqa: BLE001
                err.append(exc)

        t = threading.Thread(target=work)
        t0 = time.perf_counter()
        t.start()
        time.sleep(1.0)
        _capi.mailbox_abort(0, boxes.ptrs[0])  # what a failing peer (kernel or host) does to rank 0's mailbox
        t.join(timeout=50)
        assert not t.is_alive() and time.perf_counter() - t0 < 45.0
        assert len(err) == 1 and isinstance(err[0], DeviceEngineError) and "abort word" in str(err[0])
    finally:
        boxes.close()


def test_bad_node_ids_are_refused():
    cp = instance("netgen")
    bad = np.array(cp.tail, copy=True)
    bad[5] = cp.n_nodes + 3
    cp.tail = bad
    with pytest.raises(DeviceEngineError, match="arc endpoint outside"):
        _capi.solve_canonical(cp, options(cp, _capi.PRICING_DANTZIG))
    cp.tail = np.where(np.arange(cp.n_arcs) == 9, 0, instance("netgen").tail).astype(np.int32)  # the root is not a real endpoint
    with pytest.raises(DeviceEngineError, match="arc endpoint outside"):
        _capi.solve_canonical(cp, options(cp, _capi.PRICING_DANTZIG))
