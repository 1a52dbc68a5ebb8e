#!/bin/bash
# forward substitution hidden under the factorisation (A/B on one box), dense + golden tests on it, quasisep default chunk
mkdir -p gpurun_out; O=gpurun_out
timeout 900 python -m pytest tests/test_dense_gpu.py tests/test_golden.py tests/test_ozaki_gpu.py -m gpu -x -q -p no:cacheprovider -k "not factor_parity or 6144" > $O/r2c16_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2c16_pytest.log
tail -4 $O/r2c16_pytest.log
run() { tag=$1; shift; timeout 240 python bench.py --quick --steps 4 --warmup 2 "$@" > $O/r2c16_$tag.json 2> $O/r2c16_$tag.err; }
run ov1
run ov0 --opt solve_overlap=0
run ov1b
run ov0b --opt solve_overlap=0
grep -h -o '"value": [0-9.]*\|"options": \[[^]]*\]\|"rel_err": [0-9.e-]*\|"frac": [0-9.]*\|"kernel_ms_per_step": {[^}]*}' $O/r2c16_ov*.json | paste - - - - - > $O/r2c16_sweep_summary.txt
cat $O/r2c16_sweep_summary.txt
timeout 200 python bench.py --workload quasisep --steps 5 --warmup 3 > $O/r2c16_qs.json 2> $O/r2c16_qs.err
grep -h -o '"value": [0-9.]*\|"frac": [0-9.]*' $O/r2c16_qs.json | head -3
