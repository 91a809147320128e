"""profiles/traffic.json from an ncu CSV of the whole resident solve (scripts/r02_profiles.sh ... c3solve):
   python scripts/make_traffic_json.py gpurun_out/prof_TAG/c3solve_traffic.csv transport_4096 "note"
bench.py copies dram_bytes_per_launch into roofline.traffic (it is NOT measured in the bench run itself)."""
import csv
import json
import sys

path, workload = sys.argv[1], sys.argv[2]
note = sys.argv[3] if len(sys.argv) > 3 else ""
vals = {}
kernel = None
for row in csv.reader(open(path)):
    if len(row) > 14 and row[0].isdigit():
        kernel = row[4]
        vals[row[12]] = float(row[14])
rec = {
    "kernel": kernel,
    "dram_bytes_read": int(vals["dram__bytes_read.sum"]),
    "dram_bytes_write": int(vals["dram__bytes_write.sum"]),
    "dram_bytes_per_launch": int(vals["dram__bytes_read.sum"] + vals["dram__bytes_write.sum"]),
    "l2_bytes_per_launch": int(vals["lts__t_bytes.sum"]),
    "duration_ns": int(vals["gpu__time_duration.sum"]),
    "source": path.replace("gpurun_out/prof_", "profiles/r02/ncu/ <- gpurun_out/prof_"),
    "note": note,
}
out = "profiles/traffic.json"
doc = json.load(open(out))
doc["_comment"] = ("DRAM traffic of the dominant kernel per launch, from `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,"
                   "lts__t_bytes.sum,gpu__time_duration.sum` on `python scripts/run_one.py <workload> 1` (scripts/r02_profiles.sh c3solve). "
                   "bench.py copies dram_bytes_per_launch into roofline.traffic.")
doc[workload] = rec
json.dump(doc, open(out, "w"), indent=1)
print(json.dumps(rec, indent=1))
