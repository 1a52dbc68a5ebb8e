#!/bin/bash
# launch list of the batched workload (config 5: 1024 x N = 4096 on the native fp64 path)
mkdir -p gpurun_out; O=gpurun_out
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $O/r2c20_launches_batched.csv \
    python bench.py --workload batched --steps 1 --warmup 0 > $O/r2c20_launches_batched.log 2>&1
python tools/ncu_summary.py $O/r2c20_launches_batched.csv > $O/r2c20_launches_batched_summary.txt 2>&1; head -20 $O/r2c20_launches_batched_summary.txt
tail -2 $O/r2c20_launches_batched.log | cut -c1-400
gzip -f $O/r2c20_launches_batched.csv
