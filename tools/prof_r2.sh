#!/bin/bash
# Round-2 launch list at the bench's default chunk size: one chunk pair of the steady state
# (fused sheet branch: 7 launches -- layers 0+1, layers 2+3, layers 4-7, head --, spectrogram branch: 9 launches = 16 per chunk pair).
MB=${MB:-4096}
N=${N:-16}
ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_elapsed,dram__bytes_read.sum,dram__bytes_write.sum \
  --clock-control none -s $((2*N)) -c $N --csv --log-file gpurun_out/r2_launches_mb$MB.csv \
  python bench.py --pairs $((2*MB)) --max-batch $MB --steps 1 --warmup 1 --skip-extras > /dev/null 2>&1
python tools/launch_table.py gpurun_out/r2_launches_mb$MB.csv --traffic-json profiles/r2_traffic.json --mb $MB \
  --command "tools/prof_r2.sh: ncu ... -s $((2*N)) -c $N python bench.py --pairs $((2*MB)) --max-batch $MB --steps 1 --warmup 1 --skip-extras"
cp profiles/r2_traffic.json gpurun_out/r2_traffic.json
