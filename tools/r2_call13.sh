#!/bin/bash
# tail split-K (exactness + block-column width sweep), QSM algebra incl. the order-J conditioned covariance, CARMA on the device
mkdir -p gpurun_out; O=gpurun_out
timeout 900 python -m pytest tests/test_qsm_gpu.py tests/test_zx_reference_tests_gpu.py tests/test_zzy_quasisep_reference_gpu.py tests/test_quasisep_gpu.py -m gpu -x -q -p no:cacheprovider --durations=6 > $O/r2c13_pytest_qs.log 2>&1; echo "pytest rc=$?" >> $O/r2c13_pytest_qs.log
tail -12 $O/r2c13_pytest_qs.log
timeout 600 python -m pytest tests/test_zzz_int8_variants_gpu.py tests/test_ozaki_gpu.py -m gpu -x -q -p no:cacheprovider -k "split_k or cta_pair or factor_parity" > $O/r2c13_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2c13_pytest.log
tail -5 $O/r2c13_pytest.log
run() { tag=$1; shift; timeout 240 python bench.py --quick --steps 3 --warmup 2 "$@" > $O/r2c13_$tag.json 2> $O/r2c13_$tag.err; }
run nb1024
run nb1024_sk0 --opt ozaki_splitk=0
run nb512 --opt nb=512
run nb512_sk0 --opt nb=512 --opt ozaki_splitk=0
run nb256 --opt nb=256
run nb512_e2048 --opt nb=512 --opt ozaki_splitk=2048
grep -h -o '"value": [0-9.]*\|"options": \[[^]]*\]\|"rel_err": [0-9.e-]*\|"frac": [0-9.]*\|"kernel_ms_per_step": {[^}]*}' $O/r2c13_nb*.json | paste - - - - - > $O/r2c13_sweep_summary.txt
cat $O/r2c13_sweep_summary.txt
tail -2 $O/r2c13_nb512.err
