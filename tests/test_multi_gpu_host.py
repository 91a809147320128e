"""CPU tests (gloo, world_size 2) of the multi-GPU host plumbing: mailbox handle exchange, the
round-robin batch assignment and the tile -> (rank, worker) map of arc-sharded sweeps.  The device
side of the exchange is exercised on real GPUs by scripts/sharded_check.py."""

import multiprocessing as mp
import os
import socket

import pytest

from network_flow_solver_b200.sharded import MailboxRing, assign_round_robin, solve_batch_round_robin, sweeper_of_tile


class FakeApi:
    """Stands in for _capi: mailboxes are (rank-tagged) integers, handles carry the owner's rank."""

    def __init__(self, rank):
        self.rank = rank
        self.opened = []
        self.resets = 0

    def mailbox_create(self, device):
        return 1000 + self.rank, bytes([self.rank]) * 64

    def mailbox_open(self, device, handle):
        assert len(handle) == 64
        self.opened.append(handle[0])
        return 2000 + handle[0]

    def mailbox_reset(self, device, ptr):
        assert ptr == 1000 + self.rank
        self.resets += 1

    def mailbox_close(self, device, ptr, is_local):
        assert (ptr == 1000 + self.rank) == bool(is_local)

    def mailbox_abort(self, device, ptr):
        self.aborted = getattr(self, "aborted", []) + [ptr]

    def solve_sharded(self, cp, opts, rank, world, pointers, out=None, probe_sweeps=0, device_arrays=None):
        if opts == "rank1 fails" and rank == 1:
            raise RuntimeError("launch failed on rank 1")
        return ("solved", rank, world, tuple(pointers))

    def solve_batch_canonical(self, cps, opts):
        return [("batch", c) for c in cps]


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    import torch.distributed as dist

    from network_flow_solver_b200.sharded import solve_canonical_sharded

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    api = FakeApi(rank)
    ring = MailboxRing(0, dist, api=api)
    res = solve_canonical_sharded(None, None, ring)
    batch = solve_batch_round_robin(list(range(7)), None, rank, world, api=api)
    # a rank that fails raises the peers' abort words; every rank of the call ends with an exception, none hangs
    try:
        solve_canonical_sharded(None, "rank1 fails", ring)
        failed = None
    except Exception as exc:
        failed = (type(exc).__name__, getattr(api, "aborted", []))
    assert failed == (("RuntimeError", [2000]) if rank == 1 else ("DeviceEngineError", [])), failed
    ring.close()
    q.put((rank, ring.pointers, sorted(api.opened), api.resets - 1, res, sorted(batch)))
    dist.destroy_process_group()


def test_mailbox_ring_and_batch_assignment_over_gloo():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, ptr0, open0, resets0, res0, b0), (r1, ptr1, open1, resets1, res1, b1) = got
    assert ptr0 == [1000, 2001] and ptr1 == [2000, 1001]  # own mailbox local, the peer's opened from its handle
    assert open0 == [1] and open1 == [0] and resets0 == resets1 == 1
    assert res0 == ("solved", 0, 2, (1000, 2001)) and res1 == ("solved", 1, 2, (2000, 1001))
    assert b0 == [0, 2, 4, 6] and b1 == [1, 3, 5]


def test_round_robin_covers_every_instance_once():
    for world in (1, 2, 4, 8):
        seen = sorted(i for r in range(world) for i in assign_round_robin(8192, r, world))
        assert seen == list(range(8192))


@pytest.mark.parametrize("world,w", [(1, 147), (2, 147), (8, 128)])
def test_tile_ownership_is_a_partition(world, w):
    owners = [sweeper_of_tile(t, world, w) for t in range(5 * world * w + 3)]
    assert all(0 <= r < world and 0 <= k < w for r, k in owners)
    # consecutive tiles go to consecutive sweepers: every rank gets the same number of tiles (+-1) of any range
    for lo in (0, 17, 1000):
        counts = [0] * world
        for t in range(lo, lo + 3 * world * w):
            counts[owners[t % len(owners)][0] if t < len(owners) else sweeper_of_tile(t, world, w)[0]] += 1
        assert max(counts) - min(counts) <= w
