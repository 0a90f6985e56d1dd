#!/bin/bash
# final launch list at the bench's default chunk size (one chunk pair = 18 launches)
MB=${MB:-4096}
ncu --metrics gpu__time_duration.sum,sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_elapsed,dram__bytes_read.sum,dram__bytes_write.sum \
  --clock-control none -s 36 -c 18 --csv --log-file gpurun_out/r1_launches_mb$MB.csv \
  python bench.py --pairs $((2*MB)) --max-batch $MB --steps 1 --warmup 1 --skip-extras > /dev/null 2>&1
python tools/launch_table.py gpurun_out/r1_launches_mb$MB.csv | tail -20
