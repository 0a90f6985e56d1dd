#!/bin/bash
# Per-launch durations of one 4096-pair chunk pair (16 launches) for the environment given on the command line:
#   bash tools/prof_layers.sh tag [VAR=value ...]
TAG=$1; shift
env "$@" ncu --metrics gpu__time_duration.sum --clock-control none -s 32 -c 16 --csv --log-file gpurun_out/layers_$TAG.csv \
  python bench.py --pairs 8192 --max-batch 4096 --steps 1 --warmup 1 --skip-extras > /dev/null 2>&1
python tools/launch_table.py gpurun_out/layers_$TAG.csv | cut -c1-70
