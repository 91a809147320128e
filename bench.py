#!/usr/bin/env python
"""bench.py - throughput of the device-resident network-simplex pivot loop on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl reference]

A "step" is one complete solve of the workload instance (one pass of the hot path: every pricing
sweep, ratio test, tree and potential update until optimality).  Default workload: BASELINE.json
config 3, the dense 4096x4096 transportation instance (16.7M arcs) - the pricing-bandwidth-bound
case the metric's GB/s figure and the >=100x target are quoted on.  Every timed step goes through
the host-buffer C-ABI call (pinned inputs in, results out): its wall clock is the `e2e` figure, the
CUDA events the library records around the resident kernel on its own stream are `value`.

Multi-GPU.  The pivot loop of ONE instance of configs 1-3 does not shard
without a per-pivot exchange, so the headline at --gpus N is N replicas (seed + rank), weak
scaling, no collective on the data path.  The two paths that DO shard are measured in the same
run, for every N including 1, and reported under `detail.multi_gpu`:
  * `sharded_cfg5` / `sharded_cfg5_dantzig`: BASELINE config 5 (2^20 nodes / 2^26 arcs), ONE
    instance, arc-sharded pricing with the NVLink candidate exchange fused into the resident kernel
    (Devex block pricing, block = M/16, and full Dantzig sweeps), a bounded prefix of the solve;
    every rank hashes trace / flows / potentials / arc states, the hashes are compared across
    ranks, with the oracle's committed prefix record and (N > 1) with a single-GPU solve;
  * `batch_cfg4`: BASELINE config 4, 8192 independent GOTO instances, instance i -> rank i mod N,
    sampled instances compared with the oracle.
`--workload goto_batch` and `--mode sharded` still run those paths as the headline of a run.

Prints ONE JSON line (rank 0).  Timing is on the device: CUDA events recorded by the C-ABI
library on the stream its kernels run on, max over ranks.  The oracle (oracle/) is executed only
as a checker of sampled results, for the cpu_baseline leg and for --impl reference (as the thing
being timed as a baseline) - never as part of the GPU path.
"""

from __future__ import annotations

import argparse
import hashlib
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

SHARDED_PREFIX = {"netgen_2e20_devex": 50000, "netgen_2e20_dantzig": 4000,   # pivots of config 5 the sharded legs run
                  "netgen_2e18_devex": 50000, "netgen_2e18_dantzig": 4000}   # (= the oracle's committed prefix records)


# ALGORITHMIC bytes per arc examined: tail 4 + head 4 + cost + state 1 (+ 4 Devex weight),
# cost = 4 when the instance's costs are exact integers (int32 column), 8 when they are float64 (perturbed costs).
# The engine may store the columns narrower (uint16 ids, int16 costs: nsx_result.bytes_per_arc is the physical
# figure); narrower storage and L2 residency show up as DRAM traffic below the algorithmic bytes, not as a
# smaller numerator.
def algorithmic_bytes_per_arc(stats: dict, devex: bool) -> int:
    return 8 + (8 if stats["cost_kind"] == 0 else 4) + 1 + (4 if devex else 0)


def env_int(name: str, default: int) -> int:
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


def host_threads() -> int:
    """Cores this process may run on - NOT omp_get_max_threads(), which torchrun pins to 1 through OMP_NUM_THREADS."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def measured_peak() -> tuple[float, str]:
    path = ROOT / "MEASURED_PEAKS.json"
    if path.exists():
        try:
            return float(json.loads(path.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """Samples SM clocks / throttle reasons with nvidia-smi while the timed region runs."""

    QUERY = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, device: int):
        self.device = device
        self.rows: list[list[str]] = []
        self.proc = None
        self.thread = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.device}", f"--query-gpu={self.QUERY}",
                 "--format=csv,noheader,nounits", "-lms", "200"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return
        def pump():
            for line in self.proc.stdout:
                self.rows.append([x.strip() for x in line.split(",")])
        self.thread = threading.Thread(target=pump, daemon=True)
        self.thread.start()

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for name, val in zip(names, r[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {
            "sm_mhz": float(np.median(sm)) if sm else None,
            "sm_max_mhz": max(mx) if mx else None,
            "samples": len(sm),
            "reasons": sorted(reasons),
        }


def pinned_copy(arr: np.ndarray):
    import torch

    t = torch.from_numpy(np.ascontiguousarray(arr)).pin_memory()
    return t, t.numpy()


def solution_hashes(r) -> dict:
    return {"trace": hashlib.sha256(np.ascontiguousarray(r.trace).tobytes()).hexdigest(),
            "flow": hashlib.sha256(np.ascontiguousarray(r.flow).tobytes()).hexdigest(),
            "pi": hashlib.sha256(np.ascontiguousarray(r.potential).tobytes()).hexdigest(),
            "state": hashlib.sha256(np.ascontiguousarray(r.state).tobytes()).hexdigest()}


# ------------------------------------------------------------------------------------------------------------------
# instance cache: canonical arrays of the large workloads are built once per box (rank 0) and shared through /dev/shm
# ------------------------------------------------------------------------------------------------------------------
def cache_dir() -> Path:
    base = Path(os.environ.get("NSX_BENCH_CACHE", "/dev/shm" if os.path.isdir("/dev/shm") else "/tmp"))
    d = base / "nsx_bench_cache"
    d.mkdir(parents=True, exist_ok=True)
    return d


def cached_canonical(wl, name: str, rank: int, barrier):
    """CanonicalProblem of `name` (seed offset 0): rank 0 builds and stores it, the others map the stored arrays."""
    from network_flow_solver_b200.canonical import CanonicalProblem

    d = cache_dir() / name
    done = d / "done.json"
    if rank == 0 and not done.exists():
        cp = wl.canonical(0)
        d.mkdir(parents=True, exist_ok=True)
        for key in ("tail", "head", "orig_cost", "pert_cost", "upper", "supply"):
            np.save(d / f"{key}.npy", np.ascontiguousarray(getattr(cp, key)))
        tmp = d / "done.json.tmp"
        tmp.write_text(json.dumps({"n_nodes": cp.n_nodes, "penalty": cp.penalty, "network_type": cp.network_type,
                                   "n_supply_nodes": cp.n_supply_nodes, "n_demand_nodes": cp.n_demand_nodes}))
        tmp.rename(done)
    barrier()
    meta = json.loads(done.read_text())
    arr = {key: np.load(d / f"{key}.npy", mmap_mode="r") for key in ("tail", "head", "orig_cost", "pert_cost", "upper", "supply")}
    return CanonicalProblem(n_nodes=meta["n_nodes"], tail=arr["tail"], head=arr["head"], orig_cost=arr["orig_cost"],
                            pert_cost=arr["pert_cost"], upper=arr["upper"], shift=np.zeros(0), supply=np.asarray(arr["supply"]),
                            penalty=meta["penalty"], network_type=meta["network_type"],
                            n_supply_nodes=meta["n_supply_nodes"], n_demand_nodes=meta["n_demand_nodes"])


# ------------------------------------------------------------------------------------------------------------------
# CPU baselines
# ------------------------------------------------------------------------------------------------------------------
def oracle_sample(cp, opts_factory, threads: int, target_seconds: float, warm=None):
    """Time the CPU restatement on a bounded sample: the first P pivots of the same instance (or, with `warm`, P pivots
    from a mid-solve tree)."""
    from oracle import oracle

    probe = 20
    t0 = time.perf_counter()
    r = oracle.solve_canonical(cp, opts_factory(max_iterations=probe), threads=threads, warm=warm)
    dt = time.perf_counter() - t0
    per = dt / max(r.iterations, 1)
    pivots = int(max(probe, min(target_seconds / max(per, 1e-9), 5_000_000)))
    t0 = time.perf_counter()
    r = oracle.solve_canonical(cp, opts_factory(max_iterations=pivots), threads=threads, warm=warm)
    dt = time.perf_counter() - t0
    return {"pivots": r.iterations, "seconds": dt, "pivots_per_s": r.iterations / dt,
            "finished": r.status in (0, 1, 3)}


def python_reference_leg(budget_s: float = 40.0) -> dict:
    """the UNMODIFIED reference (baseline/_ref/network_solver, installed by
    __graft_entry__.build(); pure Python, one core): (a) config 1 end to end through solve_min_cost_flow,
    (b) its per-pivot cost on a 512x512 dense transportation instance, driven through _find_entering_arc + _pivot
    (simplex.py:1058-1075,1176-1425), projected per arc to config 3 (the reference cannot hold 16.7M ArcState objects
    and its Python pricing loop is O(M) per pivot)."""
    ref_dir = ROOT / "baseline" / "_ref"
    if not (ref_dir / "network_solver").is_dir():
        return {"unavailable": "baseline/_ref/network_solver is not installed (run __graft_entry__.build() where /root/reference exists)"}
    code = r'''
import json, sys, time, io, contextlib
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, sys.argv[2])
from network_solver import SolverOptions, solve_min_cost_flow
from network_solver.simplex import NetworkSimplex
from network_flow_solver_b200 import generators as gen
out = {}
a = gen.gridgen_like(seed=808)
p = gen.to_network_problem(a, tolerance=1e-6)
o = SolverOptions(pricing_strategy="devex", explicit_pricing_strategy=True, auto_scale=False)
best = None
for _ in range(3):
    t = time.perf_counter()
    with contextlib.redirect_stdout(io.StringIO()):
        r = solve_min_cost_flow(p, o)
    dt = time.perf_counter() - t
    best = dt if best is None or dt < best else best
out["config1"] = {"workload": "gridgen_8_08a_like", "status": r.status, "pivots": r.iterations, "objective": r.objective,
                  "seconds_best_of_3": best, "pivots_per_s": r.iterations / best}
size = 512
a = gen.transportation(size, size, cost_max=1000, seed=4096)
p = gen.to_network_problem(a, tolerance=1e-6)
s = NetworkSimplex(p, SolverOptions(pricing_strategy="dantzig", explicit_pricing_strategy=True, auto_scale=False))
s._apply_phase_costs(1); s._rebuild_tree_structure()
K, tp, tv = 8, 0.0, 0.0
for k in range(K):
    t = time.perf_counter(); e = s._find_entering_arc(True); tp += time.perf_counter() - t
    t = time.perf_counter(); s._pivot(*e); tv += time.perf_counter() - t
m3, n3 = 4096 * 4096, 8192
pr, pv = tp / K / a.n_arcs, tv / K / (a.n_arcs + a.n_nodes)
out["config3_projected"] = {"measured_on": "transportation 512x512 (262144 arcs), first %d Phase-1 pivots" % K,
    "pricing_ns_per_arc": pr * 1e9, "pivot_ns_per_arc_plus_node": pv * 1e9,
    "seconds_per_pivot_at_16.7M_arcs": pr * m3 + pv * (m3 + n3),
    "pivots_per_s": 1.0 / (pr * m3 + pv * (m3 + n3)),
    "note": "projection, per-arc costs are linear in M (Python loops over all arcs, simplex.py:1081-1107, specialized_pivots.py:80-120); the dense (N-1)^2 basis rebuild every 65 pivots is not included"}
print(json.dumps(out))
'''
    env = dict(os.environ, NUMBA_CACHE_DIR="/tmp/nsx_numba_cache", OMP_NUM_THREADS="1")
    try:
        proc = subprocess.run([sys.executable, "-c", code, str(ref_dir), str(ROOT)], capture_output=True, text=True,
                              timeout=budget_s * 4, env=env)
        if proc.returncode != 0:
            return {"unavailable": "reference run failed: " + proc.stderr.strip().splitlines()[-1][:200]}
        out = json.loads(proc.stdout.strip().splitlines()[-1])
        out["cores"] = 1
        out["kind"] = "reference"
        return out
    except Exception as exc:  # missing numba/scipy on the box, timeout, ...
        return {"unavailable": f"{type(exc).__name__}: {exc}"[:200]}


# ------------------------------------------------------------------------------------------------------------------
# multi-GPU legs (run for every N, reported under detail.multi_gpu)
# ------------------------------------------------------------------------------------------------------------------
def sharded_leg(name: str, rank: int, world: int, device: int, dist, barrier, peak: float, log) -> dict:
    """One bounded prefix of BASELINE config 5, ONE instance, pricing sharded over all ranks (nsx_solve_sharded)."""
    import torch

    from network_flow_solver_b200 import _capi
    from network_flow_solver_b200.sharded import MailboxRing, solve_canonical_sharded
    from network_flow_solver_b200.workloads import WORKLOADS

    wl = WORKLOADS[name]
    pivots = SHARDED_PREFIX[name]
    t0 = time.perf_counter()
    cp = cached_canonical(wl, name.rsplit("_", 1)[0], rank, barrier)  # (the Devex and the Dantzig leg price the same instance)
    t_build = time.perf_counter() - t0
    m = cp.n_arcs
    opts = wl.engine_options(cp, device=device, max_iterations=pivots, trace_capacity=pivots)
    devex = wl.pricing == _capi.PRICING_DEVEX
    t0 = time.perf_counter()
    dev = [torch.from_numpy(np.ascontiguousarray(getattr(cp, k))).to(f"cuda:{device}") for k in ("tail", "head", "pert_cost", "upper")]
    ptrs = [t.data_ptr() for t in dev]
    torch.cuda.synchronize()
    t_upload = time.perf_counter() - t0
    ring = MailboxRing(device, dist if world > 1 else None)
    rec: dict = {"workload": name, "description": wl.description, "max_pivots": pivots, "arcs": int(m), "nodes": int(cp.n_nodes),
                 "world": world, "instance_build_s": round(t_build, 1), "upload_s": round(t_upload, 2)}
    try:
        runs = []
        for _ in range(2):  # first run: warm-up + parity; second: timed (both are full runs of the same prefix)
            barrier()
            runs.append(solve_canonical_sharded(cp, opts, ring, device_arrays=ptrs))
        r = runs[-1]
        mine = solution_hashes(r)
        first = solution_hashes(runs[0])
        sigs = [None] * world
        if world > 1:
            dist.all_gather_object(sigs, (r.status, r.iterations, mine, first == mine))
        else:
            sigs[0] = (r.status, r.iterations, mine, first == mine)
        ranks_agree = all(s[:3] == sigs[0][:3] for s in sigs)
        repeatable = all(s[3] for s in sigs)
        # oracle prefix records (scripts/oracle_prefix.py): the record of exactly this prefix pins trace, flows, potentials and
        # arc states; a record of a shorter prefix pins the entering-arc sequence up to its length (hashes of trace[:k])
        golden_ok = None
        golden_note = None
        for gpath in sorted((ROOT / "tests" / "golden" / "full").glob(f"{name}_prefix*.json")):
            g = json.loads(gpath.read_text())
            if g["max_iterations"] == pivots:
                ok = bool(g["status"] == r.status and g["iterations"] == r.iterations and g["trace_sha"] == mine["trace"]
                          and g["flow_sha"] == mine["flow"] and g["pi_sha"] == mine["pi"] and g["state_sha"] == mine["state"])
                golden_note = (golden_note or "") + f"{gpath.name}: trace/flow/pi/state {'equal' if ok else 'DIFFER'}; "
            else:
                marks = {int(k): v for k, v in g["trace_sha_at"].items() if int(k) <= min(len(r.trace), g["iterations"])}
                ok = bool(marks) and all(hashlib.sha256(np.ascontiguousarray(r.trace[:k]).tobytes()).hexdigest() == v for k, v in marks.items())
                golden_note = (golden_note or "") + f"{gpath.name}: entering arcs of the first {max(marks) if marks else 0} pivots {'equal' if ok else 'DIFFER'}; "
            golden_ok = ok if golden_ok is None else (golden_ok and ok)
        single_ok = None
        single_ms = None
        single_rec = None
        # the same prefix as an ordinary single-GPU solve (rank 0; the others wait at the barrier): the parity anchor of the
        # sharded run and - under the Dantzig rule, where one GPU prices from the row cache (star pricing) instead of sweeping
        # all arcs - the figure the sharded sweeps have to be read against
        if True:
            if rank == 0:
                single = _capi.solve_resident(cp, opts, ptrs)
                single_ok = bool((single.status, single.iterations, solution_hashes(single)) == sigs[0][:3])
                single_ms = single.timing["solve_ms"]
                single_rec = {"solve_ms": single_ms, "pivots_per_s": single.iterations / (single_ms * 1e-3),
                              "us_per_pivot": {"total": 1e3 * single_ms / max(single.iterations, 1),
                                               "pricing": 1e3 * single.timing["pricing_ms"] / max(single.iterations, 1),
                                               "pivot_and_tree": 1e3 * single.timing["pivot_ms"] / max(single.iterations, 1)},
                              "arcs_priced_per_pivot": single.arcs_priced / max(single.iterations, 1),
                              "star_pricing": {k: single.stats[k] for k in ("star_pricing", "star_updates", "star_builds", "star_rescans")}}
            barrier()
        t = torch.tensor([r.timing["solve_ms"], r.timing["pricing_ms"], r.timing["exchange_ms"], r.timing["sync_ms"],
                          r.timing["pivot_ms"]], dtype=torch.float64, device=f"cuda:{device}")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        solve_ms, pricing_ms, xchg_ms, sync_ms, pivot_ms = (float(x) for x in t.tolist())
        sweeps = max(r.stats["sweeps"], 1)
        bpa = algorithmic_bytes_per_arc(r.stats, devex)
        my_arcs = r.arcs_priced / world
        rec.update({
            "status": r.status, "pivots": r.iterations, "sweeps": r.stats["sweeps"],
            "solve_ms": solve_ms, "pivots_per_s": r.iterations / (solve_ms * 1e-3),
            "us_per_pivot": {"total": 1e3 * solve_ms / max(r.iterations, 1),
                             "pricing_incl_exchange": 1e3 * pricing_ms / max(r.iterations, 1),
                             "exchange": 1e3 * xchg_ms / max(r.iterations, 1),
                             "pivot_and_tree": 1e3 * pivot_ms / max(r.iterations, 1)},
            "exchange_us_per_sweep": 1e3 * xchg_ms / sweeps,
            "arcs_priced": int(r.arcs_priced), "arcs_priced_per_sweep": r.arcs_priced / sweeps,
            "per_rank_roofline": {"bound": "hbm", "bytes_per_arc": bpa, "stored_bytes_per_arc": r.stats["bytes_per_arc"] + (4 if devex else 0),
                                  "whole_kernel_GBps": my_arcs * bpa / (solve_ms * 1e-3) / 1e9,
                                  "whole_kernel_frac": my_arcs * bpa / (solve_ms * 1e-3) / 1e9 / peak,
                                  "sweeps_only_GBps": my_arcs * bpa / max(pricing_ms * 1e-3, 1e-12) / 1e9,
                                  "sweeps_only_frac": my_arcs * bpa / max(pricing_ms * 1e-3, 1e-12) / 1e9 / peak, "peak": peak},
            "pivot_phase_us": {k: round(v / 1.9e3 / max(r.iterations, 1), 3) for k, v in zip(
                ["walk", "residuals", "ratio", "flow", "bookkeeping", "snapshot", "window", "copy_stem", "potentials", "cadence"],
                r.stats["phase_cycles"])},
            "avg_rehung_subtree": r.stats["sum_subtree"] / max(r.tree_updates, 1),
            "hashes": mine, "ranks_agree": ranks_agree, "repeatable": repeatable, "matches_oracle_prefix_record": golden_ok,
            "oracle_prefix_records": golden_note,
            "matches_single_gpu": single_ok, "single_gpu_solve_ms": single_ms, "single_gpu": single_rec,
            "parity_ok": bool(ranks_agree and repeatable and golden_ok is not False and single_ok is not False
                              and (golden_ok is True or single_ok is True)),
        })
    finally:
        ring.close()
        del dev
        torch.cuda.empty_cache()
    return rec


def batch_leg(count: int, rank: int, world: int, device: int, dist, barrier, peak: float, sample: int = 2) -> dict:
    """BASELINE config 4: `count` independent GOTO instances, instance i -> rank i mod world (nsx_solve_batch)."""
    import torch

    from network_flow_solver_b200 import _capi
    from network_flow_solver_b200.sharded import assign_round_robin
    from network_flow_solver_b200.workloads import WORKLOADS
    from collections import Counter

    wl = WORKLOADS["goto_64"]
    mine = assign_round_robin(count, rank, world)
    t0 = time.perf_counter()
    cps = [wl.canonical(i) for i in mine]
    t_build = time.perf_counter() - t0
    opts = wl.engine_options(cps[0], device=device, max_iterations=10**8, trace_capacity=0)
    _capi.solve_batch_canonical(cps[: max(1, len(cps) // 16)], opts)  # warm-up (module load, clocks)
    barrier()
    t0 = time.perf_counter()
    outs = _capi.solve_batch_canonical(cps, opts)
    wall = time.perf_counter() - t0
    barrier()
    tm = outs[0].timing
    pivots = sum(o.iterations for o in outs)
    arcs = sum(o.arcs_priced for o in outs)
    status = Counter(int(o.status) for o in outs)
    # sampled parity: a few instances of this rank against the oracle (bit-exact flows / potentials / states / pivots)
    from oracle import oracle

    checked, bad = 0, 0
    for k in range(0, len(cps), max(1, len(cps) // max(sample, 1)))[:sample] if cps else []:
        ref = oracle.solve_canonical(cps[k], wl.engine_options(cps[k], max_iterations=10**8))
        g = outs[k]
        same = (ref.status == g.status and ref.iterations == g.iterations and np.array_equal(ref.flow, g.flow)
                and np.array_equal(ref.potential, g.potential) and np.array_equal(ref.state, g.state))
        checked += 1
        bad += 0 if same else 1
    vec = torch.tensor([tm["solve_ms"], tm["h2d_ms"] + tm["solve_ms"] + tm["d2h_ms"], 1e3 * wall], dtype=torch.float64, device=f"cuda:{device}")
    tot = torch.tensor([float(pivots), float(arcs), float(checked), float(bad), float(len(cps)), float(status.get(0, 0))],
                       dtype=torch.float64, device=f"cuda:{device}")
    if world > 1:
        dist.all_reduce(vec, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    solve_ms, dev_e2e_ms, wall_ms = (float(x) for x in vec.tolist())
    pivots_all, arcs_all, checked_all, bad_all, n_all, optimal_all = (float(x) for x in tot.tolist())
    bpa = algorithmic_bytes_per_arc(outs[0].stats, False)
    h2d = sum(cp.n_arcs * 24 + cp.n_nodes * 8 for cp in cps)
    return {
        "workload": "goto_batch", "instances": int(n_all), "instances_this_rank": len(cps), "world": world,
        "instance_build_s": round(t_build, 1),
        "solve_s": solve_ms * 1e-3, "pivots": int(pivots_all), "pivots_per_s": pivots_all / (solve_ms * 1e-3),
        "e2e_s_device_events": dev_e2e_ms * 1e-3, "e2e_s_wall": wall_ms * 1e-3, "e2e_pivots_per_s": pivots_all / (wall_ms * 1e-3),
        "h2d_bytes_this_rank": int(h2d),
        "status_counts_this_rank": {str(k): v for k, v in status.items()}, "optimal": int(optimal_all),
        "grid_ctas": outs[0].stats.get("grid"), "resident_mode": outs[0].stats["resident_mode"],
        "per_rank_roofline": {"bound": "hbm", "bytes_per_arc": bpa, "GBps": arcs * bpa / (tm["solve_ms"] * 1e-3) / 1e9,
                              "frac": arcs * bpa / (tm["solve_ms"] * 1e-3) / 1e9 / peak, "peak": peak},
        "sampled_vs_oracle": {"checked": int(checked_all), "mismatches": int(bad_all)},
        "parity_ok": bool(checked_all > 0 and bad_all == 0 and optimal_all == n_all),
    }


def main() -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="transport_4096")
    ap.add_argument("--batch", type=int, default=8192, help="instances in the goto_batch workload / the batch_cfg4 leg")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="CPU-baseline sample budget")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-probe", action="store_true")
    ap.add_argument("--mode", default="replicas", choices=["replicas", "sharded"],
                    help="N>1: independent replicas (weak scaling) or ONE instance with arc-sharded pricing (strong)")
    ap.add_argument("--max-pivots", type=int, default=0, help="bound the solve to this many pivots (0 = to optimality)")
    ap.add_argument("--probe-sweeps", type=int, default=200)
    ap.add_argument("--legs", default="auto",
                    help="multi-GPU legs reported under detail.multi_gpu: comma list of sharded,sharded_dantzig,batch; "
                         "'auto' = all three when the headline is the default workload, 'none' = skip")
    ap.add_argument("--sharded-workload", default="netgen_2e20", help="family of the sharded legs (netgen_2e20 = config 5, netgen_2e18 = quarter scale)")
    args = ap.parse_args()

    rank = env_int("RANK", 0)
    world = env_int("WORLD_SIZE", 1)
    local_rank = env_int("LOCAL_RANK", 0)

    from network_flow_solver_b200 import _capi
    from network_flow_solver_b200.workloads import WORKLOADS

    batch_mode = args.workload == "goto_batch"
    wl = WORKLOADS["goto_64" if batch_mode else args.workload]
    bpa = None  # known after the first solve (the engine reports the layout it chose)
    config = {
        "workload": args.workload,
        "description": wl.description,
        "pricing": ("row_scan" if "transport" in wl.name else {0: "dantzig", 1: "devex", 2: "candidate_list", 3: "devex_loop"}[wl.pricing]),
        "perturbation_eps": wl.eps_base,
        "parallelism": (f"batch round-robin x{world}" if batch_mode else
                        f"arc-sharded pricing x{world} (NVLink candidate exchange)" if args.mode == "sharded" else f"replicas x{world}"),
        "max_pivots": args.max_pivots or None,
        "l2_policy": "256 MB memset (> 126 MB L2) between timed steps; within a step the arc store is re-streamed once per pivot",
    }

    # ------------------------------------------------------------------ reference arm (CPU)
    if args.impl == "reference":
        if rank != 0:
            return 0
        from concurrent.futures import ThreadPoolExecutor

        cores = host_threads()
        replicas = max(1, args.gpus)  # the GPU arm solves one instance per GPU: the CPU arm solves as many at once,
        per = max(1, cores // replicas)  # sharing the host's cores (the fixed resource) between them
        cps = [wl.canonical(i) for i in range(replicas)]
        factories = [(lambda cpi: (lambda **kw: wl.engine_options(cpi, **kw)))(c) for c in cps]
        per_step_budget = max(2.0, min(args.cpu_seconds, 100.0 / max(args.steps + args.warmup, 1)))

        def one_step(seconds):
            t0 = time.perf_counter()
            with ThreadPoolExecutor(max_workers=replicas) as pool:
                outs = list(pool.map(lambda i: oracle_sample(cps[i], factories[i], per, seconds), range(replicas)))
            return sum(o["pivots"] for o in outs), time.perf_counter() - t0, sum(o["pivots_per_s"] for o in outs)

        for _ in range(args.warmup):
            one_step(per_step_budget / 4)
        piv, rate = 0, 0.0
        sec = 0.0
        for _ in range(args.steps):
            p, s, r = one_step(per_step_budget)
            piv += p; sec += s; rate += r
        value = rate / max(args.steps, 1)  # aggregate over the concurrent replicas (each timed on its own clock)
        one = oracle_sample(cps[0], factories[0], 1, min(per_step_budget, 6.0))
        sample = (f"first {piv // max(args.steps * replicas, 1)} pivots of the same instance per step and replica (Phase 1 prefix); "
                  f"{replicas} concurrent replica(s) x {per} thread(s) of {cores} host cores (sched_getaffinity)")
        line = {
            "impl": "reference", "metric": "pivots_per_second", "value": value, "unit": "pivots/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * sec / max(args.steps, 1), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
            "cpu_baseline": {"value": value, "unit": "pivots/s", "cores": per * replicas, "kind": "port", "sample": sample,
                             "one_thread": {"value": one["pivots_per_s"], "cores": 1,
                                            "sample": f"first {one['pivots']} pivots, 1 thread (the reference itself is single-threaded)"},
                             "python_reference": python_reference_leg()},
            "e2e": {"value": value, "unit": "pivots/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
        }
        print(json.dumps(line))
        return 0

    # ------------------------------------------------------------------ our arm (GPU)
    # NCCL / torch print banners on stdout; the contract is ONE JSON line there, so everything
    # before the final print goes to stderr
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    import torch

    def log(msg):
        if rank == 0:
            print(f"[bench {time.strftime('%H:%M:%S')}] {msg}", file=sys.stderr, flush=True)

    dist = None
    if world > 1:
        import torch.distributed as dist  # plumbing only: barrier + max over ranks

        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    device = local_rank
    torch.cuda.set_device(device)

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(device)
    peak, peak_src = measured_peak()
    mid_state = None

    if batch_mode:
        mine = [i for i in range(args.batch) if i % world == rank]
        cps = [wl.canonical(i) for i in mine]
        opts = wl.engine_options(cps[0], device=device, max_iterations=10**8)
        for _ in range(args.warmup):
            _capi.solve_batch_canonical(cps[: max(1, len(cps) // 8)], opts)
        barrier(); sampler.start()
        dev_ms = e2e_ms = 0.0; pivots = 0; arcs = 0
        for _ in range(args.steps):
            outs = _capi.solve_batch_canonical(cps, opts)
            t = outs[0].timing
            dev_ms += t["solve_ms"]; e2e_ms += t["h2d_ms"] + t["solve_ms"] + t["d2h_ms"]
            pivots += sum(o.iterations for o in outs); arcs += sum(o.arcs_priced for o in outs)
        barrier(); clocks = sampler.stop()
        bpa = algorithmic_bytes_per_arc(outs[0].stats, wl.pricing == 1)
        bpa_phys = outs[0].stats["bytes_per_arc"] + (4 if wl.pricing == 1 else 0)
        h2d = sum(cp.n_arcs * 24 + cp.n_nodes * 8 for cp in cps)
        d2h = sum((cp.n_arcs + cp.n_nodes) * 9 + cp.n_nodes * 8 for cp in cps)
        launches = 1
        from collections import Counter
        stats = {"instances_this_rank": len(cps), "status_counts": dict(Counter(int(o.status) for o in outs)),
                 "pivots_per_instance_mean": pivots / args.steps / max(len(cps), 1),
                 "grid_ctas": outs[0].stats.get("grid"), "bytes_per_arc": outs[0].stats["bytes_per_arc"],
                 "resident_mode": outs[0].stats["resident_mode"], "ring_stages": outs[0].stats["ring_stages"]}
        cp0 = cps[0]
        ring = None
    else:
        sharded = args.mode == "sharded"
        cp0 = wl.canonical(0 if sharded else rank)
        m = cp0.n_arcs
        okw = {"max_iterations": args.max_pivots} if args.max_pivots > 0 else {}
        opts = wl.engine_options(cp0, device=device, **okw)
        # pinned host copies (e2e path) and resident device copies (warm-up, probes)
        keep, host = [], {}
        for name in ("tail", "head", "pert_cost", "upper"):
            t, a = pinned_copy(getattr(cp0, name))
            keep.append(t); host[name] = a
        cp0.tail, cp0.head, cp0.pert_cost, cp0.upper = host["tail"], host["head"], host["pert_cost"], host["upper"]
        dev = [torch.from_numpy(host[k]).to(f"cuda:{device}") for k in ("tail", "head", "pert_cost", "upper")]
        ptrs = [t.data_ptr() for t in dev]
        ma = m + cp0.n_nodes - 1
        out_t = {"flow": torch.empty(ma, dtype=torch.float64).pin_memory(),
                 "potential": torch.empty(cp0.n_nodes, dtype=torch.float64).pin_memory(),
                 "state": torch.empty(ma, dtype=torch.uint8).pin_memory()}
        out = {k: v.numpy() for k, v in out_t.items()}
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=f"cuda:{device}")  # > 126 MB L2
        ring = None
        if sharded:
            from network_flow_solver_b200.sharded import MailboxRing, solve_canonical_sharded
            ring = MailboxRing(device, dist)

        def solve_dev():   # arc arrays resident in HBM
            if sharded:
                return solve_canonical_sharded(cp0, opts, ring, out=out)  # (host-buffer entry; device-timed part is solve_ms)
            return _capi.solve_resident(cp0, opts, ptrs, out=out)

        def solve_host():  # pinned host arrays in, pinned host results out
            if sharded:
                return solve_canonical_sharded(cp0, opts, ring, out=out)
            return _capi.solve_canonical(cp0, opts, out=out)

        log(f"headline {args.workload}: {args.warmup} warm-up + {args.steps} timed solves")
        for _ in range(args.warmup):
            last = solve_dev()
        barrier(); sampler.start()
        # Every timed step is ONE host-buffer call: its wall clock (allocation + H2D of the pinned inputs + pack + solve +
        # D2H of flows / potentials / states) is the end-to-end figure; the CUDA events the library records around the
        # resident kernel on its own stream give the device figure of the same step.
        dev_ms = 0.0; pivots = 0; arcs = 0; pricing_ms = pivot_ms = sync_ms = 0.0
        e2e_ms = 0.0; e2e_dev_ms = 0.0
        for _ in range(args.steps):
            flush.zero_(); torch.cuda.synchronize()
            t0 = time.perf_counter()
            last = solve_host()
            e2e_ms += 1e3 * (time.perf_counter() - t0)
            e2e_dev_ms += last.timing["h2d_ms"] + last.timing["solve_ms"] + last.timing["d2h_ms"]
            dev_ms += last.timing["solve_ms"]; pivots += last.iterations; arcs += last.arcs_priced
            pricing_ms += last.timing["pricing_ms"]; pivot_ms += last.timing["pivot_ms"]; sync_ms += last.timing["sync_ms"]
        barrier(); clocks = sampler.stop()
        bpa = algorithmic_bytes_per_arc(last.stats, wl.pricing == 1)
        bpa_phys = last.stats["bytes_per_arc"] + (4 if wl.pricing == 1 else 0)
        # the sweep kernel alone: K sweeps of the initial state through the real command / arrival protocol
        probe = {}
        if rank == 0 and not args.no_probe and not sharded:
            for label, env in (("engine_layout", None), ("wide_layout", "wide")):
                old = os.environ.get("NSX_LAYOUT")
                if env: os.environ["NSX_LAYOUT"] = env
                try:
                    _capi.sweep_probe(cp0, opts, ptrs, 20)
                    pr_short = _capi.sweep_probe(cp0, opts, ptrs, max(args.probe_sweeps // 4, 1))
                    pr = _capi.sweep_probe(cp0, opts, ptrs, args.probe_sweeps)
                finally:
                    if env:
                        if old is None: os.environ.pop("NSX_LAYOUT", None)
                        else: os.environ["NSX_LAYOUT"] = old
                b = algorithmic_bytes_per_arc(pr.stats, wl.pricing == 1)
                # per-sweep time = slope between a short and a long probe: the launch, the copy of the node state into shared
                # memory and the first full recompute of the potentials are paid once per launch, not per sweep
                n_short = max(args.probe_sweeps // 4, 1)
                if args.probe_sweeps > n_short:
                    us_sweep = 1e3 * (pr.timing["solve_ms"] - pr_short.timing["solve_ms"]) / (args.probe_sweeps - n_short)
                else:
                    us_sweep = 1e3 * pr.timing["solve_ms"] / args.probe_sweeps
                arcs_sweep = pr.arcs_priced / args.probe_sweeps
                gbs = arcs_sweep * b / (us_sweep * 1e-6) / 1e9
                probe[label] = {"algorithmic_bytes_per_arc": b,
                                "stored_bytes_per_arc": pr.stats["bytes_per_arc"] + (4 if wl.pricing == 1 else 0), "us_per_sweep": us_sweep,
                                "us_per_sweep_incl_launch": 1e3 * pr.timing["solve_ms"] / args.probe_sweeps,
                                "method": f"slope between {n_short} and {args.probe_sweeps} sweeps (one launch each)",
                                "arcs_per_sweep": arcs_sweep, "GBps": gbs, "frac_of_peak": gbs / peak,
                                "handshake_us": [round(x / 1e3 / args.probe_sweeps, 2) for x in pr.stats["handshake_ns"]]}
        # a mid-solve tree for the CPU baseline's second sample (our arm only; the GPU result is just the oracle's start state)
        if rank == 0 and world == 1 and not args.no_cpu_baseline and not sharded and last.status == 0 and args.max_pivots == 0:
            try:
                half = max(1, last.iterations // 2)
                mid = _capi.solve_resident(cp0, wl.engine_options(cp0, device=device, max_iterations=half), ptrs)
                mid_state = (half, (mid.state & _capi.ARC_IN_TREE).astype(np.uint8), mid.flow.copy())
            except Exception as exc:  # the sample is optional
                log(f"mid-solve state unavailable: {exc}")
        h2d = m * 24 + cp0.n_nodes * 8
        d2h = ma * 9 + cp0.n_nodes * 8
        launches = 4  # nsx_classify_costs_kernel, nsx_pack_kernel, nsx_init_kernel, nsx_resident_kernel per step
        stats = {
            "status": last.status, "pivots_per_solve": last.iterations,
            "phase1_pivots": last.phase1_iterations, "degenerate_pivots": last.degenerate_pivots,
            "grid_ctas": last.stats.get("grid"),
            "avg_cycle_len": last.stats["sum_cycle_len"] / max(last.iterations, 1),
            "avg_rehung_subtree": last.stats["sum_subtree"] / max(last.tree_updates, 1),
            "avg_potential_levels": last.stats["sum_rounds"] / max(last.tree_updates, 1),
            "avg_preorder_window": last.stats["sum_window"] / max(last.tree_updates, 1),
            "bytes_per_arc": last.stats["bytes_per_arc"], "ring_stages": last.stats["ring_stages"],
            "resident_mode": last.stats["resident_mode"], "sweeps_per_solve": last.stats["sweeps"],
            "phase_ms_per_step": {"pricing": pricing_ms / args.steps, "pivot_and_tree": pivot_ms / args.steps,
                                  "of_which_grid_wait": sync_ms / args.steps},
            "sweep_only_GBps": arcs * bpa / max(pricing_ms, 1e-9) / 1e6,
            "us_per_pivot": {"total": 1e3 * dev_ms / max(pivots, 1), "pricing": 1e3 * pricing_ms / max(pivots, 1),
                             "pivot_and_tree": 1e3 * pivot_ms / max(pivots, 1)},
            "pivot_phase_us": {k: round(v / 1.9e3 / max(last.iterations, 1), 3) for k, v in zip(
                ["walk", "residuals", "ratio", "flow", "bookkeeping", "snapshot", "window", "copy_stem",
                 "potentials", "cadence"], last.stats["phase_cycles"])},
            "star_pricing": {k: last.stats[k] for k in ("star_pricing", "star_updates", "star_builds", "star_rescans")},
            "blk_rebuilds": last.stats["blk_rebuilds"],
            "e2e_device_events_ms_per_step": e2e_dev_ms / args.steps,
            "exchange_ms_per_step": last.timing.get("exchange_ms", 0.0),
            "exchange_us_per_sweep": 1e3 * last.timing.get("exchange_ms", 0.0) / max(last.stats["sweeps"], 1),
            "sweep_probe": probe,
        }
        del dev, flush
        torch.cuda.empty_cache()

    # max over ranks of the device time, sum of the work
    t_dev = torch.tensor([dev_ms, e2e_ms], dtype=torch.float64, device=f"cuda:{device}")
    work = torch.tensor([float(pivots), float(arcs)], dtype=torch.float64, device=f"cuda:{device}")
    if dist is not None:
        dist.all_reduce(t_dev, op=dist.ReduceOp.MAX)
        dist.all_reduce(work, op=dist.ReduceOp.SUM)
    dev_ms_max, e2e_ms_max = (float(x) for x in t_dev.tolist())
    pivots_all, arcs_all = (float(x) for x in work.tolist())
    if args.mode == "sharded" and not batch_mode:
        pivots_all /= world  # every rank applies the same pivots: one solve, counted once
    value = pivots_all / (dev_ms_max * 1e-3)
    e2e_value = pivots_all / (e2e_ms_max * 1e-3)
    my_arcs = arcs / world if (args.mode == "sharded" and not batch_mode) else arcs  # arcs THIS rank streamed
    achieved = (my_arcs / args.steps) * bpa / (dev_ms / args.steps * 1e-3) / 1e9  # this rank's kernel
    if ring is not None:  # (ring only exists on the single-instance sharded path)
        ring.close()

    # ---------------------------------------------------------------- the multi-GPU paths that shard (every N)
    legs = args.legs
    if legs == "auto":
        legs = "sharded,sharded_dantzig,batch" if (args.workload == "transport_4096" and args.mode == "replicas"
                                                    and args.max_pivots == 0) else "none"
    multi = {}
    for leg in [x for x in legs.split(",") if x and x != "none"]:
        t0 = time.perf_counter()
        try:
            if leg == "sharded":
                multi["sharded_cfg5"] = sharded_leg(args.sharded_workload + "_devex", rank, world, device, dist, barrier, peak, log)
            elif leg == "sharded_dantzig":
                multi["sharded_cfg5_dantzig"] = sharded_leg(args.sharded_workload + "_dantzig", rank, world, device, dist, barrier, peak, log)
            elif leg == "batch":
                multi["batch_cfg4"] = batch_leg(args.batch, rank, world, device, dist, barrier, peak)
        except Exception as exc:  # a failed leg is reported, it does not take the headline down with it
            multi[leg] = {"error": f"{type(exc).__name__}: {exc}"[:400], "parity_ok": False}
            if dist is not None:
                log(f"leg {leg} failed on this rank: {exc}")
        log(f"leg {leg}: {time.perf_counter() - t0:.1f} s")

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import oracle  # noqa: F401  (the thing being timed as the CPU baseline)

        threads = host_threads()
        cpb = wl.canonical(0)
        factory = lambda **kw: wl.engine_options(cpb, **kw)
        s = oracle_sample(cpb, factory, threads, args.cpu_seconds)
        one = oracle_sample(cpb, factory, 1, min(args.cpu_seconds, 6.0))
        cpu = {"value": s["pivots_per_s"], "unit": "pivots/s", "cores": threads, "kind": "port",
               "sample": f"first {s['pivots']} pivots of the same instance ({s['seconds']:.1f} s), {threads} threads (sched_getaffinity)",
               "one_thread": {"value": one["pivots_per_s"], "cores": 1, "sample": f"first {one['pivots']} pivots, 1 thread"}}
        if mid_state is not None:
            from network_flow_solver_b200.warm_start import WarmStart

            half, in_tree, flow = mid_state
            try:
                warm = WarmStart(in_tree=in_tree, flow=flow, start_phase=1)
                sm = oracle_sample(cpb, factory, threads, args.cpu_seconds, warm=warm)
                cpu["mid_solve"] = {"value": sm["pivots_per_s"], "cores": threads,
                                    "sample": f"{sm['pivots']} pivots from the tree the GPU solve had after {half} pivots "
                                              f"(warm-started oracle, {sm['seconds']:.1f} s)"}
            except Exception as exc:
                cpu["mid_solve"] = {"unavailable": f"{type(exc).__name__}: {exc}"[:200]}
        cpu["python_reference"] = python_reference_leg()

    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return 0
    traffic = None
    tpath = ROOT / "profiles" / "traffic.json"
    traffic_note = None
    if tpath.exists():
        try:
            tj = json.loads(tpath.read_text()).get(args.workload, {})
            traffic = tj.get("dram_bytes_per_launch")
            traffic_note = ("not measured in this run: dram__bytes_read.sum + dram__bytes_write.sum of one ncu capture of the same "
                            "kernel, " + str(tj.get("captured", "see profiles/traffic.json")))
        except Exception:
            traffic = None
    if multi:
        stats["multi_gpu"] = multi
    line = {
        "metric": "pivots_per_second", "value": value, "unit": "pivots/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms_max / args.steps,
        "higher_is_better": True, "scaling": "strong" if (args.mode == "sharded" or batch_mode) else "weak",  # fixed total work: one sharded instance / one batch
        "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": config,
        "e2e": {"value": e2e_value, "unit": "pivots/s", "h2d_bytes_per_step": int(h2d),
                "d2h_bytes_per_step": int(d2h), "ms_per_step": e2e_ms_max / args.steps},
        "gpu_launches": launches * args.steps,
        "clocks": clocks,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": traffic, "traffic_note": traffic_note, "peak_source": peak_src,
                     "kernel": ("nsx_batch_kernel" if batch_mode else "nsx_resident_kernel") + " (whole resident pivot loop: sweeps + pivots)",
                     "bytes_per_arc": bpa, "stored_bytes_per_arc": bpa_phys, "arcs_priced_per_launch": arcs / args.steps,
                     "algorithmic_bytes_per_launch": arcs / args.steps * bpa},
        "cpu_baseline": cpu,
        "detail": stats,
    }
    sys.stdout.flush()
    os.dup2(saved_stdout, 1)
    print(json.dumps(line))
    return 0


if __name__ == "__main__":
    sys.exit(main())
