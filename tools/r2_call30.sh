#!/bin/bash
# last check of the final tree: QSM GPU tests (device log-sum), smoke, the driver-form bench line
mkdir -p gpurun_out; O=gpurun_out
timeout 600 python -m pytest tests/test_qsm_gpu.py tests/test_zzy_quasisep_reference_gpu.py -m gpu -x -q -p no:cacheprovider > $O/r2c30_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2c30_pytest.log
tail -3 $O/r2c30_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/r2c30_smoke.log 2>&1; echo "smoke rc=$?" >> $O/r2c30_smoke.log; tail -2 $O/r2c30_smoke.log
timeout 600 python bench.py > $O/r2c30_bench.json 2> $O/r2c30_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2c30_bench.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ['value','ms_per_step','e2e','gpu_launches']}, d['roofline']['frac'])
for k,v in d.get('configs',{}).items(): print(k, v['value'], v.get('e2e',{}).get('value') if v.get('e2e') else None, v.get('roofline',{}).get('traffic'))
PY
tail -2 $O/r2c30_bench.err
