#!/usr/bin/env python
"""bench.py - throughput of the device-resident network-simplex pivot loop on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl reference]

A "step" is one complete solve of the workload instance (one pass of the hot path: every pricing
sweep, ratio test, tree and potential update until optimality).  Default workload: BASELINE.json
config 3, the dense 4096x4096 transportation instance (16.7M arcs) - the pricing-bandwidth-bound
case the metric's GB/s figure and the >=100x target are quoted on.  With --gpus N > 1 every rank
solves an independent instance of the same family (seed + rank): the pivot loop of one instance
does not shard without a per-pivot exchange, so this is "replicas", weak scaling, no collective
on the data path (multi-GPU).  --workload goto_batch runs BASELINE config 4 (a
batch of independent GOTO instances, sharded round-robin over the ranks).

Prints ONE JSON line (rank 0).  Timing is on the device: CUDA events recorded by the C-ABI
library on the stream its kernels run on, max over ranks.  The oracle (oracle/) is executed only
for the cpu_baseline leg and for --impl reference, as the thing being timed as a baseline - never
as part of the GPU path.
"""

from __future__ import annotations

import argparse
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


def oracle_sample(cp, opts_factory, threads: int, target_seconds: float):
    """Time the CPU restatement on a bounded sample: the first P pivots of the same instance."""
    from oracle import oracle

    probe = 20
    t0 = time.perf_counter()
    r = oracle.solve_canonical(cp, opts_factory(max_iterations=probe), threads=threads)
    dt = time.perf_counter() - t0
    per = dt / max(r.iterations, 1)
    pivots = int(max(probe, min(target_seconds / max(per, 1e-9), 5_000_000)))
    t0 = time.perf_counter()
    r = oracle.solve_canonical(cp, opts_factory(max_iterations=pivots), threads=threads)
    dt = time.perf_counter() - t0
    return {"pivots": r.iterations, "seconds": dt, "pivots_per_s": r.iterations / dt,
            "finished": r.status in (0, 1, 3)}


def main() -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="transport_4096")
    ap.add_argument("--batch", type=int, default=8192, help="instances in the goto_batch workload")
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="CPU-baseline sample budget")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-probe", action="store_true")
    ap.add_argument("--mode", default="replicas", choices=["replicas", "sharded"],
                    help="N>1: independent replicas (weak scaling) or ONE instance with arc-sharded pricing (strong)")
    ap.add_argument("--max-pivots", type=int, default=0, help="bound the solve to this many pivots (0 = to optimality)")
    ap.add_argument("--probe-sweeps", type=int, default=200)
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
        "pricing": ("row_scan" if "transport" in wl.name else {0: "dantzig", 1: "devex", 2: "candidate_list"}[wl.pricing]),
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
        from oracle import oracle

        threads = oracle.max_threads()
        cp = wl.canonical(0)
        factory = lambda **kw: wl.engine_options(cp, **kw)
        per_step_budget = max(2.0, min(args.cpu_seconds, 120.0 / max(args.steps + args.warmup, 1)))
        for _ in range(args.warmup):
            oracle_sample(cp, factory, threads, per_step_budget / 4)
        piv, sec = 0, 0.0
        for _ in range(args.steps):
            s = oracle_sample(cp, factory, threads, per_step_budget)
            piv += s["pivots"]; sec += s["seconds"]
        value = piv / sec
        sample = f"first {piv // max(args.steps,1)} pivots of the same instance per step (Phase 1 prefix)"
        line = {
            "impl": "reference", "metric": "pivots_per_second", "value": value, "unit": "pivots/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * sec / max(args.steps, 1), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
            "cpu_baseline": {"value": value, "unit": "pivots/s", "cores": threads, "kind": "port",
                             "sample": sample},
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
    else:
        sharded = args.mode == "sharded"
        cp0 = wl.canonical(0 if sharded else rank)
        m = cp0.n_arcs
        okw = {"max_iterations": args.max_pivots} if args.max_pivots > 0 else {}
        opts = wl.engine_options(cp0, device=device, **okw)
        # pinned host copies (e2e path) and resident device copies (kernel-only path)
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

        for _ in range(args.warmup):
            last = solve_dev()
        barrier(); sampler.start()
        dev_ms = 0.0; pivots = 0; arcs = 0; pricing_ms = pivot_ms = sync_ms = 0.0
        for _ in range(args.steps):
            flush.zero_(); torch.cuda.synchronize()
            last = solve_dev()
            dev_ms += last.timing["solve_ms"]; pivots += last.iterations; arcs += last.arcs_priced
            pricing_ms += last.timing["pricing_ms"]; pivot_ms += last.timing["pivot_ms"]; sync_ms += last.timing["sync_ms"]
        barrier()
        bpa = algorithmic_bytes_per_arc(last.stats, wl.pricing == 1)
        bpa_phys = last.stats["bytes_per_arc"] + (4 if wl.pricing == 1 else 0)
        # end to end through the host-buffer entry point (pinned inputs, results read back)
        e2e_ms = 0.0; e2e_dev_ms = 0.0
        for _ in range(args.steps):
            flush.zero_(); torch.cuda.synchronize()
            t0 = time.perf_counter()
            r = solve_host()
            e2e_ms += 1e3 * (time.perf_counter() - t0)       # wall clock of the C-ABI call: alloc + H2D + solve + D2H
            e2e_dev_ms += r.timing["h2d_ms"] + r.timing["solve_ms"] + r.timing["d2h_ms"]
        barrier(); clocks = sampler.stop()
        # the sweep kernel alone: K sweeps of the initial state through the real command / arrival protocol
        probe = {}
        if rank == 0 and not args.no_probe and not sharded:
            for label, env in (("engine_layout", None), ("wide_layout", "wide")):
                old = os.environ.get("NSX_LAYOUT")
                if env: os.environ["NSX_LAYOUT"] = env
                try:
                    _capi.sweep_probe(cp0, opts, ptrs, 20)
                    pr = _capi.sweep_probe(cp0, opts, ptrs, args.probe_sweeps)
                finally:
                    if env:
                        if old is None: os.environ.pop("NSX_LAYOUT", None)
                        else: os.environ["NSX_LAYOUT"] = old
                b = algorithmic_bytes_per_arc(pr.stats, wl.pricing == 1)
                gbs = pr.arcs_priced * b / (pr.timing["solve_ms"] * 1e-3) / 1e9
                probe[label] = {"algorithmic_bytes_per_arc": b,
                                "stored_bytes_per_arc": pr.stats["bytes_per_arc"] + (4 if wl.pricing == 1 else 0), "us_per_sweep": 1e3 * pr.timing["solve_ms"] / args.probe_sweeps,
                                "arcs_per_sweep": pr.arcs_priced / args.probe_sweeps, "GBps": gbs, "frac_of_peak": gbs / peak,
                                "handshake_us": [round(x / 1e3 / args.probe_sweeps, 2) for x in pr.stats["handshake_ns"]]}
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
            "pivot_phase_us": {k: round(v / 1.9e3 / max(last.iterations, 1), 3) for k, v in zip(
                ["walk", "residuals", "ratio", "flow", "bookkeeping", "snapshot", "window", "copy_stem",
                 "potentials", "cadence"], last.stats["phase_cycles"])},
            "e2e_device_events_ms_per_step": e2e_dev_ms / args.steps,
            "exchange_ms_per_step": last.timing.get("exchange_ms", 0.0),
            "exchange_us_per_sweep": 1e3 * last.timing.get("exchange_ms", 0.0) / max(last.stats["sweeps"], 1),
            "sweep_probe": probe,
        }

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

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import oracle

        threads = oracle.max_threads()
        cpb = wl.canonical(0)
        s = oracle_sample(cpb, lambda **kw: wl.engine_options(cpb, **kw), threads, args.cpu_seconds)
        cpu = {"value": s["pivots_per_s"], "unit": "pivots/s", "cores": threads, "kind": "port",
               "sample": f"first {s['pivots']} pivots of the same instance ({s['seconds']:.1f} s)"}

    if not batch_mode and ring is not None:  # (ring only exists on the single-instance path)
        ring.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return 0
    traffic = None
    tpath = ROOT / "profiles" / "traffic.json"
    if tpath.exists():
        try:
            traffic = json.loads(tpath.read_text()).get(args.workload, {}).get("dram_bytes_per_launch")
        except Exception:
            traffic = None
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
                     "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
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
