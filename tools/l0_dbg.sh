#!/bin/bash
for d in ${DBGS:-0 1 2 4 3 5 6 7}; do
  ASR_L0_DEBUG=$d timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 36 -c 18 --csv \
    --log-file gpurun_out/l0_dbg$d.csv python bench.py --pairs 2048 --max-batch 1024 --steps 1 --warmup 1 --skip-extras > /dev/null 2>&1
  echo "== ASR_L0_DEBUG=$d"; python tools/launch_table.py gpurun_out/l0_dbg$d.csv | grep "^l0" | awk '{printf "%s ", $NF} END {print ""}'
done
