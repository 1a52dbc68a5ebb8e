#!/bin/bash
# the driver's 2-GPU form: torchrun bench.py --gpus 2 (replicas of the headline + C4/C5 sub-records + the sharded C3 record)
mkdir -p gpurun_out; O=gpurun_out
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 5 --warmup 3 > $O/r2c26_bench_gpus2.json 2> $O/r2c26_bench_gpus2.err; echo "rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2c26_bench_gpus2.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ['value','n_gpus','ms_per_step','e2e','clocks']})
s=d.get('sharded'); print(s and {k:s[k] for k in ['ms_per_step','logp','golden','kernel_ms_per_step_rank0','tflops_n3_over_3']})
for k,v in d.get('configs',{}).items(): print(k, v['value'])
PY
tail -3 $O/r2c26_bench_gpus2.err
