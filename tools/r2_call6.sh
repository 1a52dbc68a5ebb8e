#!/bin/bash
mkdir -p gpurun_out; O=gpurun_out
timeout 900 python -m pytest tests/test_zz_first_run_gpu.py tests/test_quasisep_gpu.py tests/test_zzy_quasisep_reference_gpu.py -m gpu -x -q -p no:cacheprovider > $O/r2c6_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2c6_pytest.log
tail -3 $O/r2c6_pytest.log
run() { tag=$1; shift; timeout 240 python bench.py --quick --steps 3 --warmup 2 "$@" > $O/r2c6_$tag.json 2> $O/r2c6_$tag.err; }
run sub0_la --opt ozaki_subpanel=0 --opt panel_overlap=2
run sub256_la --opt panel_overlap=2
run sub512_la --opt ozaki_subpanel=512 --opt panel_overlap=2
run sub0_la_nb2048 --opt ozaki_subpanel=0 --opt panel_overlap=2 --nb 2048
run sub0_la_bahead --opt ozaki_subpanel=0 --opt panel_overlap=2 --opt build_ahead=1
grep -h -o '"value": [0-9.]*\|"options": \[[^]]*\]\|"rel_err": [0-9.e-]*\|"frac": [0-9.]*\|"kernel_ms_per_step": {[^}]*}' $O/r2c6_*.json | paste - - - - - > $O/r2c6_sweep_summary.txt
cat $O/r2c6_sweep_summary.txt
for v in "qs_chunk=0" "qs_chunk=64" "qs_chunk=88"; do
  tag=$(echo "$v" | tr ' =' '__'); args=""; for o in $v; do args="$args --opt $o"; done
  timeout 200 python bench.py --workload quasisep --steps 5 --warmup 3 $args > $O/r2c6_qs_$tag.json 2> $O/r2c6_qs_$tag.err
done
grep -h -o '"value": [0-9.]*\|"frac": [0-9.]*\|"kernel_ms_per_step": {[^}]*}\|"rel_err": [-0-9.e]*' $O/r2c6_qs_*.json | paste - - - - -
