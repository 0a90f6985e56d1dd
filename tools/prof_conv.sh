#!/bin/bash
# launch list (one chunk pair = 18 launches) with pipe metrics, then --set full captures of the two heaviest kernels
ncu --metrics gpu__time_duration.sum,sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_elapsed,dram__bytes_read.sum,dram__bytes_write.sum \
  --clock-control none -s 36 -c 18 --csv --log-file gpurun_out/r1_launches_final.csv \
  python bench.py --pairs 2048 --max-batch 1024 --steps 1 --warmup 1 --skip-extras > /dev/null 2>&1
ncu --set full --import-source on --clock-control none -k regex:conv3x3_rows_kernel -s 6 -c 1 -o gpurun_out/r1_rows_l1 -f \
  python bench.py --pairs 2048 --max-batch 1024 --steps 1 --warmup 1 --skip-extras > /dev/null 2>&1
ncu --set full --import-source on --clock-control none -k regex:l0_tc_kernel -s 5 -c 1 -o gpurun_out/r1_l0tc -f \
  python bench.py --pairs 2048 --max-batch 1024 --steps 1 --warmup 1 --skip-extras > /dev/null 2>&1
ls -la gpurun_out/
