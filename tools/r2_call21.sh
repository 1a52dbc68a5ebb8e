#!/bin/bash
# batched workload: lower-triangle build with per-problem constants in the single-leaf kernel
mkdir -p gpurun_out; O=gpurun_out
timeout 600 python -m pytest tests/test_dense_gpu.py -m gpu -x -q -p no:cacheprovider > $O/r2c21_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2c21_pytest.log
tail -4 $O/r2c21_pytest.log
timeout 400 python bench.py --workload batched --steps 2 --warmup 1 > $O/r2c21_batched.json 2> $O/r2c21_batched.err
grep -h -o '"value": [0-9.]*\|"ms_per_step": [0-9.]*\|"tflops_n3_over_3": [0-9.]*\|"max_rel_err_vs_oracle_on_grid_corners": [0-9.e-]*' $O/r2c21_batched.json | tr '\n' ' '; echo
tail -2 $O/r2c21_batched.err
