#!/bin/bash
# Per-layer durations of one 4096-pair chunk pair with the conv kernels' halves switched off (ASR_CONV_DEBUG 0 / 1 / 2).
for d in 0 1 2; do
  ASR_CONV_DEBUG=$d ncu --metrics gpu__time_duration.sum --clock-control none -s 34 -c 17 --csv --log-file gpurun_out/r2_dbg$d.csv \
    python bench.py --pairs 8192 --max-batch 4096 --steps 1 --warmup 1 --skip-extras > /dev/null 2>&1
done
python - <<'PY'
import csv
cols = []
for d in (0, 1, 2):
    rows = [l for l in open("gpurun_out/r2_dbg%d.csv" % d) if l.startswith('"')]
    cols.append([(r["Kernel Name"][:28], float(r["Metric Value"].replace(",", "")) / 1e3) for r in csv.DictReader(rows)])
print("%-30s %10s %14s %10s" % ("kernel", "full us", "no epilogue", "no MMA"))
for a, b, c in zip(*cols):
    print("%-30s %10.1f %14.1f %10.1f" % (a[0], a[1], b[1], c[1]))
PY
