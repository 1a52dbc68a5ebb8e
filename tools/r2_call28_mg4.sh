#!/bin/bash
# the driver's 4-GPU form: torchrun bench.py --gpus 4 (replicas + the sharded C3 record over 4 ranks)
mkdir -p gpurun_out; O=gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29545 bench.py --gpus 4 --steps 3 --warmup 3 > $O/r2c27_bench_gpus4.json 2> $O/r2c27_bench_gpus4.err; echo "rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2c27_bench_gpus4.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ['value','n_gpus','ms_per_step','e2e','clocks']})
s=d.get('sharded'); print(s and {k:s.get(k) for k in ['ms_per_step','logp','golden','kernel_ms_per_step_rank0','tflops_n3_over_3','exchange_bytes_per_step_per_rank']})
PY
tail -4 $O/r2c27_bench_gpus4.err
