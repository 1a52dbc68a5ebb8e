#!/bin/bash
# suite on the new defaults (sub-panels, normal-form build, warp tree) + panel / qs sweeps + ncu of the shipped kernels + C3 golden
mkdir -p gpurun_out; O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q -p no:cacheprovider > $O/r2c5_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/r2c5_pytest_gpu.log
tail -4 $O/r2c5_pytest_gpu.log
run() { tag=$1; shift; timeout 240 python bench.py --quick --steps 3 --warmup 2 "$@" > $O/r2c5_$tag.json 2> $O/r2c5_$tag.err; }
run base
run sub0 --opt ozaki_subpanel=0
run sub512 --opt ozaki_subpanel=512
run povl --opt panel_overlap=1
run povl_sub0 --opt panel_overlap=1 --opt ozaki_subpanel=0
run bf0 --opt build_fast=0
run nb2048 --nb 2048
grep -h -o '"value": [0-9.]*\|"options": \[[^]]*\]\|"rel_err": [0-9.e-]*\|"frac": [0-9.]*\|"kernel_ms_per_step": {[^}]*}' $O/r2c5_*.json | paste - - - - - > $O/r2c5_sweep_summary.txt
cat $O/r2c5_sweep_summary.txt
for v in "qs_occupancy=1" "qs_occupancy=0" "qs_occupancy=1 qs_chunk=96" ; do
  tag=$(echo "$v" | tr ' =' '__'); args=""; for o in $v; do args="$args --opt $o"; done
  timeout 200 python bench.py --workload quasisep --steps 5 --warmup 3 $args > $O/r2c5_qs_$tag.json 2> $O/r2c5_qs_$tag.err
done
grep -h -o '"value": [0-9.]*\|"frac": [0-9.]*\|"kernel_ms_per_step": {[^}]*}\|"rel_err": [-0-9.e]*' $O/r2c5_qs_*.json | paste - - - - -
# ncu: shipped int8 update kernel at bench size (one mid-factorisation launch), K1 build, quasisep launch list
timeout 500 ncu --set full --clock-control none --import-source on -k regex:i8_update_kernel_2sm -s 45 -c 1 -o $O/r2c5_i8 -f \
    python bench.py --quick --steps 1 --warmup 0 --opt ozaki_subpanel=0 > $O/r2c5_ncu_i8.log 2>&1
ncu -i $O/r2c5_i8.ncu-rep --page raw --csv > $O/r2c5_i8_raw.csv 2>/dev/null
ncu -i $O/r2c5_i8.ncu-rep --page source --csv 2>/dev/null | gzip > $O/r2c5_i8_source.csv.gz; rm -f $O/r2c5_i8.ncu-rep
timeout 300 ncu --set full --clock-control none -k regex:build_rect -c 2 -o $O/r2c5_build -f python tools/k1_build.py > $O/r2c5_ncu_build.log 2>&1
ncu -i $O/r2c5_build.ncu-rep --page raw --csv > $O/r2c5_build_raw.csv 2>/dev/null; rm -f $O/r2c5_build.ncu-rep
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $O/r2c5_launches_dense.csv \
    python bench.py --quick --steps 1 --warmup 0 > $O/r2c5_launches_dense.log 2>&1
gzip -f $O/r2c5_launches_dense.csv
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2c5_launches_qs.csv \
    python bench.py --workload quasisep --steps 1 --warmup 1 > $O/r2c5_launches_qs.log 2>&1
# C3 known answer at full size through the native fp64 path (one-off)
timeout 400 python tools/make_c3_golden_gpu.py > $O/r2c5_c3_golden.log 2>&1; tail -2 $O/r2c5_c3_golden.log
du -sh $O
