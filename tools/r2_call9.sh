#!/bin/bash
# final-candidate check on one GPU: whole GPU suite, smoke, the driver's default bench line (with sub-records), reference arm
mkdir -p gpurun_out; O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q -p no:cacheprovider > $O/r2c9_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/r2c9_pytest_gpu.log
tail -4 $O/r2c9_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/r2c9_smoke.log 2>&1; echo "smoke rc=$?" >> $O/r2c9_smoke.log; tail -4 $O/r2c9_smoke.log
timeout 900 python bench.py --gpus 1 --steps 10 --warmup 3 > $O/r2c9_bench.json 2> $O/r2c9_bench.err; echo "bench rc=$?"
tail -c 3000 $O/r2c9_bench.json; tail -3 $O/r2c9_bench.err
timeout 400 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > $O/r2c9_bench_ref.json 2> $O/r2c9_bench_ref.err; echo "ref rc=$?"
cat $O/r2c9_bench_ref.json | cut -c1-1500
