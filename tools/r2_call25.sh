#!/bin/bash
# final candidate on one GPU: whole GPU suite, smoke, driver-form bench line (with sub-records), reference arm
mkdir -p gpurun_out; O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q -p no:cacheprovider > $O/r2c25_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/r2c25_pytest_gpu.log
tail -4 $O/r2c25_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/r2c25_smoke.log 2>&1; echo "smoke rc=$?" >> $O/r2c25_smoke.log; tail -5 $O/r2c25_smoke.log
timeout 900 python bench.py --gpus 1 --steps 10 --warmup 3 > $O/r2c25_bench.json 2> $O/r2c25_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2c25_bench.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ['value','ms_per_step','per_step_ms','kernel_ms_per_step','clocks','e2e','gpu_launches']})
print(d['roofline']['frac'], d['roofline']['traffic'], d['golden'], d['cpu_baseline']['value'])
for k,v in d.get('configs',{}).items(): print(k, v['value'], v.get('e2e'), v.get('roofline',{}).get('frac'))
PY
tail -3 $O/r2c25_bench.err
timeout 400 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > $O/r2c25_bench_ref.json 2> $O/r2c25_bench_ref.err; echo "ref rc=$?"
cut -c1-400 $O/r2c25_bench_ref.json
