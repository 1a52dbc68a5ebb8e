#!/bin/bash
# re-entry check of HEAD on one GPU: whole GPU suite, smoke, driver-form bench line
mkdir -p gpurun_out; O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q -p no:cacheprovider --durations=15 > $O/r2c10_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/r2c10_pytest_gpu.log
tail -25 $O/r2c10_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/r2c10_smoke.log 2>&1; echo "smoke rc=$?" >> $O/r2c10_smoke.log; tail -4 $O/r2c10_smoke.log
timeout 900 python bench.py --gpus 1 --steps 5 --warmup 3 > $O/r2c10_bench.json 2> $O/r2c10_bench.err; echo "bench rc=$?"
tail -c 1500 $O/r2c10_bench.json; tail -3 $O/r2c10_bench.err
