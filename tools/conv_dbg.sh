#!/bin/bash
# per-layer durations of one chunk pair under the conv kernel's diagnostic switches (ASR_CONV_DEBUG)
for d in ${DBGS:-0 1 2}; do
  ASR_CONV_DEBUG=$d timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 36 -c 18 --csv \
    --log-file gpurun_out/conv_dbg$d.csv python bench.py --pairs 2048 --max-batch 1024 --steps 1 --warmup 1 --skip-extras > /dev/null 2>&1
  echo "== ASR_CONV_DEBUG=$d ASR_CONV_SLOTS=$ASR_CONV_SLOTS"; python tools/launch_table.py gpurun_out/conv_dbg$d.csv | grep -v "^l0\|^head" | awk '{printf "%s ", $NF} END {print ""}'
done
