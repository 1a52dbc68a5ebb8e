#!/bin/bash
# digit cutting in halves, quasisep chunk-length sweep, pinned e2e; quick dense line
mkdir -p gpurun_out; O=gpurun_out
timeout 900 python -m pytest tests/test_ozaki_gpu.py tests/test_golden.py tests/test_quasisep_gpu.py tests/test_zzy_quasisep_reference_gpu.py -m gpu -x -q -p no:cacheprovider > $O/r2c15_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2c15_pytest.log
tail -4 $O/r2c15_pytest.log
for cm in 128 180 264 400; do
  timeout 200 python bench.py --workload quasisep --steps 5 --warmup 3 --opt qs_chunk_max=$cm > $O/r2c15_qs_$cm.json 2> $O/r2c15_qs_$cm.err
  echo "qs_chunk_max=$cm $(grep -h -o '"value": [0-9.]*\|"frac": [0-9.]*\|"rel_err": [0-9.e-]*' $O/r2c15_qs_$cm.json | head -4 | tr '\n' ' ')"
done
timeout 300 python bench.py --quick --steps 3 --warmup 2 > $O/r2c15_quick.json 2> $O/r2c15_quick.err
grep -h -o '"value": [0-9.]*\|"rel_err": [0-9.e-]*\|"frac": [0-9.]*\|"kernel_ms_per_step": {[^}]*}' $O/r2c15_quick.json | head -4
