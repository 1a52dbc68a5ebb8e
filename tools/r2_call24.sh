#!/bin/bash
# right-looking chain inside the diagonal block of the look-ahead panel: parity + A/B on one box
mkdir -p gpurun_out; O=gpurun_out
timeout 600 python -m pytest tests/test_zzz_int8_variants_gpu.py -m gpu -x -q -p no:cacheprovider -k "right_looking" > $O/r2c24_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2c24_pytest.log
tail -3 $O/r2c24_pytest.log
run() { tag=$1; shift; timeout 240 python bench.py --quick --steps 4 --warmup 2 "$@" > $O/r2c24_$tag.json 2> $O/r2c24_$tag.err; }
run pc0
run pc1 --opt panel_chain=1
run pc0b
run pc1b --opt panel_chain=1
grep -h -o '"value": [0-9.]*\|"options": \[[^]]*\]\|"rel_err": [0-9.e-]*\|"frac": [0-9.]*\|"kernel_ms_per_step": {[^}]*}' $O/r2c24_pc*.json | paste - - - - - > $O/r2c24_sweep_summary.txt
cat $O/r2c24_sweep_summary.txt
