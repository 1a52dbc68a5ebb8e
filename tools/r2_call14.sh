#!/bin/bash
# whole GPU suite on the new build (single-leaf build kernels, templated digit cutting, tail split-K, QSM, CARMA), bench,
# ncu of K1 and of the shipped int8 update, launch list of one dense step
mkdir -p gpurun_out; O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q -p no:cacheprovider --durations=10 > $O/r2c14_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/r2c14_pytest_gpu.log
tail -16 $O/r2c14_pytest_gpu.log
timeout 300 python bench.py --quick --steps 3 --warmup 2 > $O/r2c14_quick.json 2> $O/r2c14_quick.err
grep -h -o '"value": [0-9.]*\|"rel_err": [0-9.e-]*\|"frac": [0-9.]*\|"kernel_ms_per_step": {[^}]*}' $O/r2c14_quick.json | head -4
timeout 300 ncu --set full --clock-control none -k regex:build_rect -c 2 -o $O/r2c14_build -f python tools/k1_build.py > $O/r2c14_ncu_build.log 2>&1
ncu -i $O/r2c14_build.ncu-rep --page raw --csv > $O/r2c14_build_raw.csv 2>/dev/null; rm -f $O/r2c14_build.ncu-rep
tail -3 $O/r2c14_ncu_build.log
timeout 500 ncu --set full --clock-control none --import-source on -k regex:i8_update_kernel_2sm -s 45 -c 1 -o $O/r2c14_i8 -f \
    python bench.py --quick --steps 1 --warmup 0 > $O/r2c14_ncu_i8.log 2>&1
ncu -i $O/r2c14_i8.ncu-rep --page raw --csv > $O/r2c14_i8_raw.csv 2>/dev/null; rm -f $O/r2c14_i8.ncu-rep
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file $O/r2c14_launches_dense.csv \
    python bench.py --quick --steps 1 --warmup 0 > $O/r2c14_launches.log 2>&1
python tools/ncu_summary.py $O/r2c14_launches_dense.csv > $O/r2c14_launches_dense_summary.txt 2>&1; head -14 $O/r2c14_launches_dense_summary.txt
gzip -f $O/r2c14_launches_dense.csv
