#!/bin/bash
# split-K of the CTA-pair int8 update: exactness, then the policy constant swept on the headline problem
mkdir -p gpurun_out; O=gpurun_out
timeout 600 python -m pytest tests/test_zzz_int8_variants_gpu.py -m gpu -x -q -p no:cacheprovider -k "split_k or cta_pair" > $O/r2c11_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2c11_pytest.log
tail -6 $O/r2c11_pytest.log
run() { tag=$1; shift; timeout 240 python bench.py --quick --steps 3 --warmup 2 "$@" > $O/r2c11_$tag.json 2> $O/r2c11_$tag.err; }
run sk0 --opt ozaki_splitk=0
run sk512 --opt ozaki_splitk=512
run sk1024
run sk2048 --opt ozaki_splitk=2048
run sk4096 --opt ozaki_splitk=4096
grep -h -o '"value": [0-9.]*\|"options": \[[^]]*\]\|"rel_err": [0-9.e-]*\|"frac": [0-9.]*\|"kernel_ms_per_step": {[^}]*}' $O/r2c11_sk*.json | paste - - - - - > $O/r2c11_sweep_summary.txt
cat $O/r2c11_sweep_summary.txt
tail -2 $O/r2c11_sk1024.err
