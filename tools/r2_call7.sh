#!/bin/bash
mkdir -p gpurun_out; O=gpurun_out
timeout 600 python -m pytest tests/test_zz_first_run_gpu.py -m gpu -x -q -p no:cacheprovider -k "overlap" > $O/r2c7_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2c7_pytest.log
tail -3 $O/r2c7_pytest.log
run() { tag=$1; shift; timeout 240 python bench.py --quick --steps 3 --warmup 2 "$@" > $O/r2c7_$tag.json 2> $O/r2c7_$tag.err; }
run sub0_la --opt ozaki_subpanel=0 --opt panel_overlap=2
run sub256_la --opt panel_overlap=2
run sub512_la --opt ozaki_subpanel=512 --opt panel_overlap=2
run sub0 --opt ozaki_subpanel=0
grep -h -o '"value": [0-9.]*\|"options": \[[^]]*\]\|"rel_err": [0-9.e-]*\|"frac": [0-9.]*\|"kernel_ms_per_step": {[^}]*}' $O/r2c7_*.json | paste - - - - - > $O/r2c7_sweep_summary.txt
cat $O/r2c7_sweep_summary.txt
