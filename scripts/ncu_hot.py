"""Summarise an ncu report: headline metrics + hottest SASS regions. Usage: ncu_hot.py REPORT [print_from print_to]"""
import csv, subprocess, sys, io
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "launch__registers_per_thread", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "sm__warps_active.avg.pct_of_peak_sustained_active"]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print(d.get("Kernel Name"))
    for k in want:
        if k in d: print("  ", k, d[k], units[hdr.index(k)])
    for k in hdr:
        if "issue_stalled" in k and k.endswith("per_issue_active.ratio"): print("  ", k.split("issue_stalled_")[1].split("_per_issue")[0], d[k])
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]
isrc, iex, ismp = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
data = []
for r in rows[2:]:
    try: data.append((r[isrc], int(r[iex]), int(r[ismp])))
    except Exception: pass
tot = sum(d[1] for d in data); tots = sum(d[2] for d in data)
print("total inst", tot, "samples", tots)
thr = max(d[1] for d in data) * 0.02
cur = None
for i, d in enumerate(data + [("", 0, 0)]):
    if d[1] >= thr:
        if cur is None: cur = [i, i, 0, 0]
        cur[1] = i; cur[2] += d[1]; cur[3] += d[2]
    elif cur:
        print(f"  region {cur[0]}-{cur[1]} n={cur[1]-cur[0]+1} inst={cur[2]/1e6:.0f}M share={cur[2]/tot:.3f} samples={cur[3]/max(tots,1):.3f}")
        cur = None
if len(sys.argv) > 3:
    for i in range(int(sys.argv[2]), int(sys.argv[3])):
        d = data[i]; print(i, f"{d[1]/1e6:.1f}", d[2], d[0][:110])
