#!/bin/bash
# restated reference QSM tests on the GPU, quasisep line after the host-side fixes (zero-mean, pinned), quasisep launch list
mkdir -p gpurun_out; O=gpurun_out
timeout 900 python -m pytest tests/test_qsm_gpu.py -m gpu -x -q -p no:cacheprovider --durations=5 > $O/r2c19_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2c19_pytest.log
tail -10 $O/r2c19_pytest.log
timeout 200 python bench.py --workload quasisep --steps 5 --warmup 3 > $O/r2c19_qs.json 2> $O/r2c19_qs.err
python - <<'PY'
import json
q=json.loads(open('gpurun_out/r2c19_qs.json').read().strip().splitlines()[-1])
print(q['value'], q['ms_per_step'], q['roofline']['frac'], q['e2e'], q['parity'], q['clocks'])
PY
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/r2c19_launches_qs.csv \
    python bench.py --workload quasisep --steps 1 --warmup 1 > $O/r2c19_launches_qs.log 2>&1
python tools/ncu_summary.py $O/r2c19_launches_qs.csv > $O/r2c19_launches_qs_summary.txt 2>&1; head -16 $O/r2c19_launches_qs_summary.txt
