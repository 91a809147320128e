"""ctypes binding of the C ABI in include/nsx_b200.h (libnsx_b200.so).

This is the only door from Python to the CUDA engine.  If the shared library is missing
or no B200 is visible the call raises DeviceEngineError - there is no CPU fallback.
"""

from __future__ import annotations

import ctypes as C
import os
import threading
from dataclasses import dataclass, field
from pathlib import Path

import numpy as np

from .canonical import CanonicalProblem
from .exceptions import DeviceEngineError

ABI_VERSION = 4  # NSX_ABI_VERSION of include/nsx_b200.h these declarations mirror
PRICING_DANTZIG = 0
PRICING_DEVEX = 1
PRICING_CANDIDATE_LIST = 2  # also what pricing_strategy="adaptive" (the reference's default) amounts to
PRICING_DEVEX_LOOP = 3  # use_vectorized_pricing=False: the sequential Devex block scan (simplex_pricing.py:205-269)
# EngineOptions.row_scan_first: structure-specific rule tried before the configured one (NSX_SPECIAL_*)
SPECIAL_NONE, SPECIAL_ROW_SCAN, SPECIAL_ASSIGNMENT, SPECIAL_MAX_FLOW, SPECIAL_SHORTEST_PATH = 0, 1, 2, 3, 4

STATUS_OPTIMAL = 0
STATUS_INFEASIBLE = 1
STATUS_ITERATION_LIMIT = 2
STATUS_UNBOUNDED = 3
STATUS_ITERATION_LIMIT_P1 = 4

ARC_IN_TREE = 1
ARC_CAN_FWD = 2
ARC_CAN_BWD = 4
ARC_TOUCHED = 8

_p_i32 = C.POINTER(C.c_int32)
_p_f64 = C.POINTER(C.c_double)
_p_u8 = C.POINTER(C.c_uint8)


class NsxProblem(C.Structure):
    _fields_ = [
        ("n_nodes", C.c_int32),
        ("n_arcs", C.c_int64),
        ("tail", _p_i32),
        ("head", _p_i32),
        ("pert_cost", _p_f64),
        ("upper", _p_f64),
        ("supply", _p_f64),
        ("penalty", C.c_double),
    ]


class NsxOptions(C.Structure):
    _fields_ = [
        ("pricing", C.c_int32),
        ("row_scan_first", C.c_int32),
        ("block_size", C.c_int64),
        ("auto_block", C.c_int32),
        ("ft_update_limit", C.c_int32),
        ("max_iterations", C.c_int64),
        ("tolerance", C.c_double),
        ("trace_capacity", C.c_int64),
        ("device", C.c_int32),
        ("flags", C.c_uint32),
        ("node_mask", C.POINTER(C.c_uint8)),
        ("spin_timeout_ms", C.c_int32),
    ]


class NsxShard(C.Structure):
    _fields_ = [("rank", C.c_int32), ("world", C.c_int32), ("mailboxes", C.POINTER(C.c_void_p))]


class NsxWarmStart(C.Structure):
    """nsx_warm_start (include/nsx_b200.h): initial spanning tree + flows built by warm_start.apply_basis."""

    _fields_ = [("in_tree", C.POINTER(C.c_uint8)), ("flow", C.POINTER(C.c_double)), ("start_phase", C.c_int32)]

    @classmethod
    def of(cls, warm) -> "NsxWarmStart":
        """`warm` = warm_start.WarmStart; the arrays must stay alive for the duration of the call."""
        return cls(_ptr(warm.in_tree, _p_u8), _ptr(warm.flow, _p_f64), int(warm.start_phase))


class NsxResult(C.Structure):
    _fields_ = [
        ("flow", _p_f64),
        ("potential", _p_f64),
        ("state", _p_u8),
        ("entering_trace", _p_i32),
        ("trace_len", C.c_int64),
        ("iterations", C.c_int64),
        ("phase1_iterations", C.c_int64),
        ("degenerate_pivots", C.c_int64),
        ("artificial_with_flow", C.c_int64),
        ("tree_updates", C.c_int64),
        ("weight_resets", C.c_int64),
        ("final_block_size", C.c_int64),
        ("arcs_priced", C.c_int64),
        ("sweeps", C.c_int64),
        ("unbounded_arc", C.c_int64),
        ("unbounded_rc", C.c_double),
        ("status", C.c_int32),
        ("grid_ctas", C.c_int32),
        ("bytes_per_arc", C.c_int32),
        ("ring_stages", C.c_int32),
        ("resident_mode", C.c_int32),
        ("store_layout", C.c_int32),
        ("solve_ms", C.c_double),
        ("h2d_ms", C.c_double),
        ("d2h_ms", C.c_double),
        ("pricing_ms", C.c_double),
        ("pivot_ms", C.c_double),
        ("sync_ms", C.c_double),
        ("exchange_ms", C.c_double),
        ("sum_cycle_len", C.c_int64),
        ("sum_subtree", C.c_int64),
        ("max_subtree", C.c_int64),
        ("sum_rounds", C.c_int64),
        ("sum_window", C.c_int64),
        ("phase_cycles", C.c_int64 * 12),
        ("handshake_ns", C.c_int64 * 8),
        ("fault", C.c_int32),
        ("star_pricing", C.c_int32),
        ("star_updates", C.c_int64),
        ("star_builds", C.c_int64),
        ("star_rescans", C.c_int64),
        ("blk_rebuilds", C.c_int64),
    ]


@dataclass
class EngineOptions:
    """Flat option record of the C ABI (nsx_options)."""

    pricing: int = PRICING_DANTZIG
    row_scan_first: bool = False
    block_size: int = 1
    auto_block: bool = False
    ft_update_limit: int = 64
    max_iterations: int = 100
    tolerance: float = 1e-6
    trace_capacity: int = 0
    device: int = 0
    flags: int = 0  # reserved by the C ABI, must stay 0
    node_mask: object = None  # uint8[n_nodes], SPECIAL_SHORTEST_PATH only (kept alive by this record)
    spin_timeout_ms: int = 0  # deadline of device-side waits, 0 = the library default (30 s)

    def to_c(self) -> NsxOptions:
        mask = None
        if self.node_mask is not None:
            self.node_mask = np.ascontiguousarray(self.node_mask, dtype=np.uint8)
            mask = _ptr(self.node_mask, _p_u8)
        return NsxOptions(
            int(self.pricing),
            int(self.row_scan_first),  # SPECIAL_*: False/True = none / transportation row scan
            int(self.block_size),
            int(bool(self.auto_block)),
            int(self.ft_update_limit),
            int(self.max_iterations),
            float(self.tolerance),
            int(self.trace_capacity),
            int(self.device),
            int(self.flags),
            mask,
            int(self.spin_timeout_ms),
        )


@dataclass
class RawSolution:
    """What comes back over the C ABI, in the engine's index space."""

    status: int
    iterations: int
    phase1_iterations: int
    flow: np.ndarray  # float64[M + N]
    potential: np.ndarray  # float64[n_nodes]
    state: np.ndarray  # uint8[M + N]
    trace: np.ndarray  # int32[trace_len] arc*2 + (dir<0)
    degenerate_pivots: int = 0
    artificial_with_flow: int = 0
    tree_updates: int = 0
    weight_resets: int = 0
    final_block_size: int = 0
    arcs_priced: int = 0
    unbounded_arc: int = -1
    unbounded_rc: float = 0.0
    timing: dict = field(default_factory=dict)
    stats: dict = field(default_factory=dict)


def _ptr(a: np.ndarray, typ):
    return a.ctypes.data_as(typ)


class CallFrame:
    """Owns the numpy buffers behind one nsx_problem / nsx_result pair."""

    def __init__(self, cp: CanonicalProblem, opts: EngineOptions, device_arrays=None, out=None):
        """device_arrays: optional (tail, head, pert_cost, upper) DEVICE addresses (ints) for the
        resident entry point; out: optional dict of preallocated (e.g. pinned) flow / potential /
        state arrays."""
        self.cp = cp
        m = cp.n_arcs
        ma = m + cp.n_nodes - 1
        self.supply = np.ascontiguousarray(cp.supply, dtype=np.float64)
        if device_arrays is None:
            self.tail = np.ascontiguousarray(cp.tail, dtype=np.int32)
            self.head = np.ascontiguousarray(cp.head, dtype=np.int32)
            self.pert = np.ascontiguousarray(cp.pert_cost, dtype=np.float64)
            self.upper = np.ascontiguousarray(cp.upper, dtype=np.float64)
            ptrs = (
                _ptr(self.tail, _p_i32),
                _ptr(self.head, _p_i32),
                _ptr(self.pert, _p_f64),
                _ptr(self.upper, _p_f64),
            )
        else:
            t, h, c, u = (int(x) for x in device_arrays)
            ptrs = (C.cast(t, _p_i32), C.cast(h, _p_i32), C.cast(c, _p_f64), C.cast(u, _p_f64))
        self.problem = NsxProblem(
            int(cp.n_nodes), int(m), *ptrs, _ptr(self.supply, _p_f64), float(cp.penalty)
        )
        self.options = opts.to_c()
        out = out or {}
        for key, dtype, size in (("flow", np.float64, ma), ("potential", np.float64, cp.n_nodes), ("state", np.uint8, ma)):
            buf = out.get(key)
            if buf is not None and not (isinstance(buf, np.ndarray) and buf.dtype == dtype and buf.ndim == 1
                                        and buf.shape[0] == size and buf.flags.c_contiguous and buf.flags.writeable):
                raise ValueError(f"out[{key!r}] must be a writable C-contiguous {np.dtype(dtype).name}[{size}] array "
                                 f"(the engine writes it through a raw pointer)")
        self.flow = out.get("flow") if out.get("flow") is not None else np.zeros(ma, dtype=np.float64)
        self.potential = (
            out.get("potential") if out.get("potential") is not None else np.zeros(cp.n_nodes, dtype=np.float64)
        )
        self.state = out.get("state") if out.get("state") is not None else np.zeros(ma, dtype=np.uint8)
        cap = max(int(opts.trace_capacity), 0)
        self.trace = np.zeros(max(cap, 1), dtype=np.int32)
        self.result = NsxResult()
        self.result.flow = _ptr(self.flow, _p_f64)
        self.result.potential = _ptr(self.potential, _p_f64)
        self.result.state = _ptr(self.state, _p_u8)
        self.result.entering_trace = _ptr(self.trace, _p_i32) if cap > 0 else None
        self.trace_capacity = cap

    def harvest(self) -> RawSolution:
        r = self.result
        n_tr = int(min(r.trace_len, self.trace_capacity))
        return RawSolution(
            status=int(r.status),
            iterations=int(r.iterations),
            phase1_iterations=int(r.phase1_iterations),
            flow=self.flow,
            potential=self.potential,
            state=self.state,
            trace=self.trace[:n_tr].copy(),
            degenerate_pivots=int(r.degenerate_pivots),
            artificial_with_flow=int(r.artificial_with_flow),
            tree_updates=int(r.tree_updates),
            weight_resets=int(r.weight_resets),
            final_block_size=int(r.final_block_size),
            arcs_priced=int(r.arcs_priced),
            unbounded_arc=int(r.unbounded_arc),
            unbounded_rc=float(r.unbounded_rc),
            timing={
                "solve_ms": float(r.solve_ms),
                "h2d_ms": float(r.h2d_ms),
                "d2h_ms": float(r.d2h_ms),
                "pricing_ms": float(r.pricing_ms),
                "pivot_ms": float(r.pivot_ms),
                "sync_ms": float(r.sync_ms),
                "exchange_ms": float(r.exchange_ms),
            },
            stats={
                "grid": int(r.grid_ctas),
                "bytes_per_arc": int(r.bytes_per_arc),
                "ring_stages": int(r.ring_stages),
                "resident_mode": int(r.resident_mode),
                "node_kind": int(r.store_layout) & 0xff,
                "cost_kind": (int(r.store_layout) >> 8) & 0xff,
                "sweeps": int(r.sweeps),
                "sum_cycle_len": int(r.sum_cycle_len),
                "sum_subtree": int(r.sum_subtree),
                "max_subtree": int(r.max_subtree),
                "sum_rounds": int(r.sum_rounds),
                "sum_window": int(r.sum_window),
                "phase_cycles": [int(x) for x in r.phase_cycles],
                "star_pricing": int(r.star_pricing), "star_updates": int(r.star_updates), "star_builds": int(r.star_builds),
                "star_rescans": int(r.star_rescans), "blk_rebuilds": int(r.blk_rebuilds),
                "handshake_ns": [int(x) for x in r.handshake_ns],
            },
        )


# ----------------------------------------------------------------------------------------------
# library loading
# ----------------------------------------------------------------------------------------------
_LIB_NAME = "libnsx_b200.so"
_lib = None
_lib_lock = threading.Lock()


def library_path() -> Path:
    return Path(__file__).resolve().parent / "csrc" / _LIB_NAME


def load_library():
    """dlopen libnsx_b200.so (built in-tree by __graft_entry__.build()); raise if absent."""
    global _lib
    with _lib_lock:
        if _lib is not None:
            return _lib
        path = Path(os.environ.get("NSX_B200_LIB", str(library_path())))
        if not path.exists():
            raise DeviceEngineError(
                f"CUDA engine {path} is not built. Run `python -c 'import __graft_entry__ as g; "
                f"g.build()'` from the repo root. There is no CPU fallback."
            )
        try:
            lib = C.CDLL(str(path))
        except OSError as exc:  # missing libcudart etc.
            raise DeviceEngineError(f"cannot load {path}: {exc}") from exc
        for name in ("nsx_solve", "nsx_solve_resident"):
            fn = getattr(lib, name)
            fn.argtypes = [C.POINTER(NsxProblem), C.POINTER(NsxOptions), C.POINTER(NsxResult)]
            fn.restype = C.c_int
        lib.nsx_solve_warm.argtypes = [
            C.POINTER(NsxProblem), C.POINTER(NsxOptions), C.POINTER(NsxWarmStart), C.POINTER(NsxResult)]
        lib.nsx_solve_warm.restype = C.c_int
        lib.nsx_sweep_probe.argtypes = [
            C.POINTER(NsxProblem), C.POINTER(NsxOptions), C.c_int32, C.POINTER(NsxResult)]
        lib.nsx_sweep_probe.restype = C.c_int
        lib.nsx_solve_sharded.argtypes = [
            C.POINTER(NsxProblem), C.POINTER(NsxOptions), C.POINTER(NsxResult), C.POINTER(NsxShard)]
        lib.nsx_solve_sharded.restype = C.c_int
        lib.nsx_solve_sharded_resident.argtypes = lib.nsx_solve_sharded.argtypes
        lib.nsx_solve_sharded_resident.restype = C.c_int
        lib.nsx_sweep_probe_sharded.argtypes = [
            C.POINTER(NsxProblem), C.POINTER(NsxOptions), C.c_int32, C.POINTER(NsxResult), C.POINTER(NsxShard)]
        lib.nsx_sweep_probe_sharded.restype = C.c_int
        lib.nsx_mailbox_bytes.restype = C.c_int64
        lib.nsx_mailbox_create.argtypes = [C.c_int32, C.POINTER(C.c_void_p), C.c_char_p]
        lib.nsx_mailbox_open.argtypes = [C.c_int32, C.c_char_p, C.POINTER(C.c_void_p)]
        lib.nsx_mailbox_reset.argtypes = [C.c_int32, C.c_void_p]
        lib.nsx_mailbox_close.argtypes = [C.c_int32, C.c_void_p, C.c_int32]
        lib.nsx_mailbox_abort.argtypes = [C.c_int32, C.c_void_p]
        for name in ("nsx_mailbox_create", "nsx_mailbox_open", "nsx_mailbox_reset", "nsx_mailbox_close", "nsx_mailbox_abort"):
            getattr(lib, name).restype = C.c_int
        lib.nsx_solve_batch.argtypes = [
            C.c_int64,
            C.POINTER(NsxProblem),
            C.POINTER(NsxOptions),
            C.POINTER(NsxResult),
        ]
        lib.nsx_solve_batch.restype = C.c_int
        lib.nsx_last_error.restype = C.c_char_p
        lib.nsx_version.argtypes = [_p_i32, _p_i32]
        abi = C.c_int32(-1)
        lib.nsx_version(C.byref(abi), None)
        if abi.value != ABI_VERSION:  # a stale build would read the structs with another layout
            raise DeviceEngineError(
                f"{path} implements ABI version {abi.value}, these bindings need {ABI_VERSION}: rebuild it "
                f"(python -c 'import __graft_entry__ as g; g.build()')."
            )
        lib.nsx_device_count.restype = C.c_int
        _lib = lib
        return lib


def last_error() -> str:
    msg = load_library().nsx_last_error()
    return msg.decode("utf-8", "replace") if msg else ""


def _check(rc: int, what: str) -> None:
    if rc != 0:
        raise DeviceEngineError(f"{what} failed with code {rc}: {last_error()}")


def solve_canonical(cp: CanonicalProblem, opts: EngineOptions, out=None, warm=None) -> RawSolution:
    """One instance, host buffers in, host buffers out (nsx_solve; nsx_solve_warm when `warm`, a
    warm_start.WarmStart, gives the initial tree). Releases the GIL."""
    lib = load_library()
    frame = CallFrame(cp, opts, out=out)
    if warm is not None:
        w = NsxWarmStart.of(warm)
        rc = lib.nsx_solve_warm(C.byref(frame.problem), C.byref(frame.options), C.byref(w), C.byref(frame.result))
        _check(rc, "nsx_solve_warm")
        return frame.harvest()
    rc = lib.nsx_solve(C.byref(frame.problem), C.byref(frame.options), C.byref(frame.result))
    _check(rc, "nsx_solve")
    return frame.harvest()


def solve_resident(cp: CanonicalProblem, opts: EngineOptions, device_arrays, out=None) -> RawSolution:
    """Arc arrays already resident in HBM: device_arrays = (tail, head, pert_cost, upper) device
    addresses, 16-byte aligned (nsx_solve_resident)."""
    lib = load_library()
    frame = CallFrame(cp, opts, device_arrays=device_arrays, out=out)
    rc = lib.nsx_solve_resident(
        C.byref(frame.problem), C.byref(frame.options), C.byref(frame.result)
    )
    _check(rc, "nsx_solve_resident")
    return frame.harvest()


def sweep_probe(cp: CanonicalProblem, opts: EngineOptions, device_arrays, sweeps: int) -> RawSolution:
    """Measurement aid: `sweeps` pricing sweeps of the initial state, no pivots (nsx_sweep_probe)."""
    lib = load_library()
    frame = CallFrame(cp, opts, device_arrays=device_arrays)
    rc = lib.nsx_sweep_probe(
        C.byref(frame.problem), C.byref(frame.options), int(sweeps), C.byref(frame.result)
    )
    _check(rc, "nsx_sweep_probe")
    return frame.harvest()


def mailbox_create(device: int) -> tuple[int, bytes]:
    """Allocate this rank's mailbox; returns (device pointer, 64-byte IPC handle)."""
    lib = load_library()
    ptr = C.c_void_p()
    handle = C.create_string_buffer(64)
    _check(lib.nsx_mailbox_create(int(device), C.byref(ptr), handle), "nsx_mailbox_create")
    return int(ptr.value), bytes(handle.raw)


def mailbox_open(device: int, handle: bytes) -> int:
    lib = load_library()
    ptr = C.c_void_p()
    _check(lib.nsx_mailbox_open(int(device), C.create_string_buffer(handle, 64), C.byref(ptr)), "nsx_mailbox_open")
    return int(ptr.value)


def mailbox_reset(device: int, ptr: int) -> None:
    _check(load_library().nsx_mailbox_reset(int(device), C.c_void_p(ptr)), "nsx_mailbox_reset")


def mailbox_abort(device: int, ptr: int) -> None:
    """Raise the abort word of a (local or peer-mapped) mailbox: the kernel polling it leaves with fault 3."""
    _check(load_library().nsx_mailbox_abort(int(device), C.c_void_p(ptr)), "nsx_mailbox_abort")


def mailbox_close(device: int, ptr: int, is_local: bool) -> None:
    _check(load_library().nsx_mailbox_close(int(device), C.c_void_p(ptr), int(is_local)), "nsx_mailbox_close")


def solve_sharded(cp: CanonicalProblem, opts: EngineOptions, rank: int, world: int, mailboxes: list[int],
                  out=None, probe_sweeps: int = 0, device_arrays=None) -> RawSolution:
    """This rank's part of an arc-sharded solve (nsx_solve_sharded); every rank gets the full result."""
    lib = load_library()
    frame = CallFrame(cp, opts, out=out, device_arrays=device_arrays)
    boxes = (C.c_void_p * world)(*[C.c_void_p(p) for p in mailboxes])
    shard = NsxShard(int(rank), int(world), C.cast(boxes, C.POINTER(C.c_void_p)))
    if probe_sweeps > 0:
        rc = lib.nsx_sweep_probe_sharded(C.byref(frame.problem), C.byref(frame.options), int(probe_sweeps),
                                         C.byref(frame.result), C.byref(shard))
    else:  # device_arrays: the arc arrays are already in this rank's HBM
        fn = lib.nsx_solve_sharded_resident if device_arrays is not None else lib.nsx_solve_sharded
        rc = fn(C.byref(frame.problem), C.byref(frame.options), C.byref(frame.result), C.byref(shard))
    _check(rc, "nsx_solve_sharded")
    return frame.harvest()


def solve_batch_canonical(cps: list[CanonicalProblem], opts: EngineOptions) -> list[RawSolution]:
    """Independent instances, one CTA each (nsx_solve_batch)."""
    lib = load_library()
    frames = [CallFrame(cp, opts) for cp in cps]
    n = len(frames)
    probs = (NsxProblem * n)(*[f.problem for f in frames])
    ress = (NsxResult * n)(*[f.result for f in frames])
    o = opts.to_c()
    rc = lib.nsx_solve_batch(n, probs, C.byref(o), ress)
    _check(rc, "nsx_solve_batch")
    for f, r in zip(frames, ress):
        f.result = r
    return [f.harvest() for f in frames]
