#!/bin/bash
# ncu evidence of one build (run on the GPU box through gpurun; every capture follows a plain run of the same command
# that exited 0).  Outputs land in gpurun_out/prof_<tag>/ and are summarised into profiles/<round>/ by hand.
#   bash scripts/r02_profiles.sh TAG [what ...]     what: c3sweep c3wide c3solve c5sweep c2pivot batch launches
set -u
TAG=${1:-r02}; shift || true
WHAT=${*:-c3sweep c3wide c3solve c5sweep c2pivot batch launches}
OUT=gpurun_out/prof_$TAG
mkdir -p "$OUT"
NCU="ncu --clock-control none --import-source on"
FULL="$NCU --set full"
has() { [[ " $WHAT " == *" $1 "* ]]; }
run_plain() { echo "== $*" >> "$OUT/plain.log"; timeout 900 "$@" >> "$OUT/plain.log" 2>&1; }

if has c3sweep; then   # config-3 pricing sweep alone, engine layout (7 B/arc stored, L2-resident)
  run_plain python scripts/probe.py transport_4096 100 &&
  timeout 900 $FULL -k regex:nsx_resident -s 1 -c 1 -o "$OUT/c3sweep" -f python scripts/probe.py transport_4096 100 > "$OUT/c3sweep.log" 2>&1
  python scripts/ncu_hot.py "$OUT/c3sweep.ncu-rep" > "$OUT/c3sweep.txt" 2>&1
fi
if has c3wide; then    # the same sweep on the 17 B/arc layout (285 MB per sweep > L2: streams HBM)
  NSX_LAYOUT=wide run_plain python scripts/probe.py transport_4096 100 &&
  NSX_LAYOUT=wide timeout 900 $FULL -k regex:nsx_resident -s 1 -c 1 -o "$OUT/c3wide" -f python scripts/probe.py transport_4096 100 > "$OUT/c3wide.log" 2>&1
  python scripts/ncu_hot.py "$OUT/c3wide.ncu-rep" > "$OUT/c3wide.txt" 2>&1
fi
if has c3solve; then   # DRAM / L2 traffic of the whole resident solve of config 3 (roofline.traffic)
  run_plain python scripts/run_one.py transport_4096 1 &&
  timeout 1200 $NCU --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,gpu__time_duration.sum -k regex:nsx_resident --csv \
      --log-file "$OUT/c3solve_traffic.csv" python scripts/run_one.py transport_4096 1 > "$OUT/c3solve.log" 2>&1
fi
if has c5sweep; then   # config-5 full sweeps (potentials gathered through L2): sector efficiency of the gather
  run_plain python scripts/probe.py netgen_2e20_dantzig 20 &&
  timeout 1200 $FULL -k regex:nsx_resident -s 1 -c 1 -o "$OUT/c5sweep" -f python scripts/probe.py netgen_2e20_dantzig 20 > "$OUT/c5sweep.log" 2>&1
  python scripts/ncu_hot.py "$OUT/c5sweep.ncu-rep" > "$OUT/c5sweep.txt" 2>&1
fi
if has c2pivot; then   # pivot-dominated solve: config 2 (tree in HBM, star pricing), first 3000 pivots
  run_plain python scripts/run_one.py netgen_2e16_dantzig 1 3000 &&
  timeout 900 $FULL -k regex:nsx_resident -c 1 -o "$OUT/c2pivot" -f python scripts/run_one.py netgen_2e16_dantzig 1 3000 > "$OUT/c2pivot.log" 2>&1
  python scripts/ncu_hot.py "$OUT/c2pivot.ncu-rep" > "$OUT/c2pivot.txt" 2>&1
fi
if has batch; then     # config 4: one CTA per instance (nsx_batch_kernel), 296 instances
  run_plain python bench.py --workload goto_batch --batch 296 --steps 1 --warmup 1 --no-cpu-baseline --no-probe --legs none &&
  timeout 900 $FULL -k regex:nsx_batch -c 1 -o "$OUT/batch" -f python bench.py --workload goto_batch --batch 296 --steps 1 --warmup 0 --no-cpu-baseline --no-probe --legs none > "$OUT/batch.log" 2>&1
  python scripts/ncu_hot.py "$OUT/batch.ncu-rep" > "$OUT/batch.txt" 2>&1
fi
if has launches; then  # launch list of the default bench command's headline (legs off: they are separate solves)
  run_plain python bench.py --steps 1 --warmup 1 --no-cpu-baseline --legs none &&
  timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file "$OUT/launches.csv" \
      python bench.py --steps 1 --warmup 1 --no-cpu-baseline --legs none > "$OUT/launches.log" 2>&1
fi
# keep the pulled directory small: reports stay only when they fit
du -sh "$OUT"; ls -la "$OUT"
find "$OUT" -name '*.ncu-rep' -size +20M -delete
exit 0
