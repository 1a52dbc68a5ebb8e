#!/bin/bash
# full GPU suite on the new defaults + quasisep timings (fast path / trees) + dense default bench
mkdir -p gpurun_out; O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q -p no:cacheprovider > $O/r2c4_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/r2c4_pytest_gpu.log
tail -5 $O/r2c4_pytest_gpu.log
for v in "qs_kernel=1" "qs_kernel=1 qs_tree=1" "qs_kernel=0" "qs_kernel=0 qs_tree=1" "qs_kernel=1 qs_tree=1 qs_chunk=32" "qs_kernel=1 qs_tree=1 qs_chunk=128"; do
  tag=$(echo "$v" | tr ' =' '__'); args=""; for o in $v; do args="$args --opt $o"; done
  timeout 200 python bench.py --workload quasisep --steps 5 --warmup 3 $args > $O/r2c4_qs_$tag.json 2> $O/r2c4_qs_$tag.err
done
grep -h -o '"value": [0-9.]*\|"frac": [0-9.]*\|"kernel_ms_per_step": {[^}]*}\|"logp": [-0-9.e]*' $O/r2c4_qs_*.json | paste - - - - 
timeout 300 python bench.py --steps 3 --warmup 3 > $O/r2c4_bench_dense.json 2> $O/r2c4_bench_dense.err
tail -c 600 $O/r2c4_bench_dense.json
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'qsf_' -c 3 -o $O/r2c4_qsf_full -f \
    python bench.py --workload quasisep --steps 1 --warmup 0 > $O/r2c4_ncu_qs.log 2>&1
ncu -i $O/r2c4_qsf_full.ncu-rep --page raw --csv > $O/r2c4_qsf_full_raw.csv 2>/dev/null
ncu -i $O/r2c4_qsf_full.ncu-rep --page source --csv > $O/r2c4_qsf_full_source.csv 2>/dev/null
gzip -f $O/r2c4_qsf_full_source.csv; rm -f $O/r2c4_qsf_full.ncu-rep
timeout 300 ncu --set full --clock-control none -k regex:build_rect -c 2 -o $O/r2c4_build_full -f \
    python tools/k1_build.py > $O/r2c4_ncu_build.log 2>&1
ncu -i $O/r2c4_build_full.ncu-rep --page raw --csv > $O/r2c4_build_full_raw.csv 2>/dev/null
rm -f $O/r2c4_build_full.ncu-rep
du -sh $O
