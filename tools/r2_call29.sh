#!/bin/bash
# ncu --set full of the two point-wise quasiseparable kernels (fold + replay) at N = 1e7 for the C4 record's traffic figure
mkdir -p gpurun_out; O=gpurun_out
timeout 400 ncu --set full --clock-control none -k regex:qsf_ -s 4 -c 2 -o $O/r2c29_qsf -f python bench.py --workload quasisep --steps 1 --warmup 1 > $O/r2c29_ncu_qsf.log 2>&1
ncu -i $O/r2c29_qsf.ncu-rep --page raw --csv > $O/r2c29_qsf_raw.csv 2>/dev/null; rm -f $O/r2c29_qsf.ncu-rep
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/r2c29_qsf_raw.csv')))
H=rows[0]; U=rows[1]; idx={h:i for i,h in enumerate(H)}
for r in rows[2:]:
    print({k:(r[idx[k]],U[idx[k]]) for k in ['Kernel Name','gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active','smsp__inst_executed.sum','launch__registers_per_thread','sm__warps_active.avg.pct_of_peak_sustained_active'] if k in idx})
PY
