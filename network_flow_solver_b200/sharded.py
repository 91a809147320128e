"""Multi-GPU host plumbing (one process per GPU, torch.distributed for the rendezvous only).

* ``solve_canonical_sharded``: arc-sharded pricing of ONE large instance (BASELINE config 5).  Every
  rank holds the full instance; the ranks exchange the 64-byte IPC handles of their mailboxes once,
  then the device kernels exchange one candidate record per rank and sweep over NVLink.  There is no
  NCCL collective on the data path - the exchange is peer stores issued by the pivot CTA.
* ``assign_round_robin`` / ``solve_batch_round_robin``: independent instances, instance i -> rank
  i mod world (BASELINE config 4), no communication during the solve.

The reference has no multi-process path; results are identical to the
single-GPU engine by construction and are tested as such.
"""

from __future__ import annotations

from . import _capi
from .canonical import CanonicalProblem


def assign_round_robin(count: int, rank: int, world: int) -> list[int]:
    """Indices of the instances rank `rank` solves: i mod world == rank."""
    return [i for i in range(count) if i % world == rank]


def sweeper_of_tile(tile: int, world: int, workers_per_rank: int) -> tuple[int, int]:
    """(rank, worker) that prices absolute tile `tile` in an arc-sharded sweep: sweeper
    s = tile mod (world * W), rank = s // W, worker = s mod W (mirrors nsx_sweep_ring)."""
    s = tile % (world * workers_per_rank)
    return s // workers_per_rank, s % workers_per_rank


class MailboxRing:
    """This rank's mailbox plus the peer mailboxes mapped into its address space."""

    def __init__(self, device: int, dist=None, api=_capi):
        self.api = api
        self.device = device
        self.dist = dist
        self.rank = dist.get_rank() if dist is not None else 0
        self.world = dist.get_world_size() if dist is not None else 1
        self.local, handle = api.mailbox_create(device)
        handles = [None] * self.world
        if dist is not None and self.world > 1:
            dist.all_gather_object(handles, handle)
        else:
            handles[0] = handle
        self.pointers = [
            self.local if r == self.rank else api.mailbox_open(device, handles[r]) for r in range(self.world)
        ]

    def reset(self):
        self.api.mailbox_reset(self.device, self.local)
        if self.dist is not None and self.world > 1:
            self.dist.barrier()

    def abort_peers(self):
        """This rank failed between the barrier and the end of its kernel: raise the abort word in every peer's
        mailbox so that their resident kernels leave at once (fault 3) instead of waiting out the spin deadline."""
        for r, p in enumerate(self.pointers):
            if r != self.rank:
                try:
                    self.api.mailbox_abort(self.device, p)
                except Exception:  # best effort: the peers still have their own deadline
                    pass

    def agree(self, ok: bool) -> list[bool]:
        """Every rank learns which ranks completed their part (host collective, after the kernels have ended)."""
        if self.dist is None or self.world == 1:
            return [bool(ok)]
        flags = [None] * self.world
        self.dist.all_gather_object(flags, bool(ok))
        return [bool(f) for f in flags]

    def close(self):
        if self.dist is not None and self.world > 1:
            self.dist.barrier()
        for r, p in enumerate(self.pointers):
            if r != self.rank:
                self.api.mailbox_close(self.device, p, False)
        self.api.mailbox_close(self.device, self.local, True)


def solve_canonical_sharded(cp: CanonicalProblem, opts: _capi.EngineOptions, ring: MailboxRing, out=None,
                            probe_sweeps: int = 0, device_arrays=None) -> _capi.RawSolution:
    ring.reset()  # zero this rank's mailbox, then a host barrier: nobody writes before everybody is clean
    error = None
    sol = None
    try:
        sol = ring.api.solve_sharded(cp, opts, ring.rank, ring.world, ring.pointers, out=out,
                                     probe_sweeps=probe_sweeps, device_arrays=device_arrays)
    except Exception as exc:  # allocation / launch failure, a spin deadline, a peer's abort word
        error = exc
        ring.abort_peers()
    done = ring.agree(error is None)  # nobody starts the next sharded call before all ranks have left this one
    if error is not None:
        raise error
    if not all(done):
        from .exceptions import DeviceEngineError

        raise DeviceEngineError(f"arc-sharded solve: ranks {[r for r, ok in enumerate(done) if not ok]} failed their part")
    return sol


def solve_batch_round_robin(cps: list[CanonicalProblem], opts: _capi.EngineOptions, rank: int, world: int,
                            api=_capi) -> dict[int, _capi.RawSolution]:
    """Solve this rank's share of a batch; returns {global index: solution}."""
    mine = assign_round_robin(len(cps), rank, world)
    outs = api.solve_batch_canonical([cps[i] for i in mine], opts) if mine else []
    return dict(zip(mine, outs))
