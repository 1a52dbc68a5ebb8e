#!/bin/bash
# split-K of the CTA-pair int8 update (exactness + policy sweep) and the first GPU run of the QSM algebra
mkdir -p gpurun_out; O=gpurun_out
timeout 900 python -m pytest tests/test_qsm_gpu.py -m gpu -x -q -p no:cacheprovider --durations=8 > $O/r2c12_pytest_qsm.log 2>&1; echo "pytest rc=$?" >> $O/r2c12_pytest_qsm.log
tail -14 $O/r2c12_pytest_qsm.log
timeout 600 python -m pytest tests/test_zzz_int8_variants_gpu.py -m gpu -x -q -p no:cacheprovider -k "split_k or cta_pair" > $O/r2c12_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2c12_pytest.log
tail -6 $O/r2c12_pytest.log
run() { tag=$1; shift; timeout 240 python bench.py --quick --steps 3 --warmup 2 "$@" > $O/r2c12_$tag.json 2> $O/r2c12_$tag.err; }
run sk0 --opt ozaki_splitk=0
run sk512 --opt ozaki_splitk=512
run sk1024
run sk2048 --opt ozaki_splitk=2048
run sk4096 --opt ozaki_splitk=4096
grep -h -o '"value": [0-9.]*\|"options": \[[^]]*\]\|"rel_err": [0-9.e-]*\|"frac": [0-9.]*\|"kernel_ms_per_step": {[^}]*}' $O/r2c12_sk*.json | paste - - - - - > $O/r2c12_sweep_summary.txt
cat $O/r2c12_sweep_summary.txt
tail -2 $O/r2c12_sk1024.err
