#!/bin/bash
mkdir -p gpurun_out; O=gpurun_out
timeout 900 python -m pytest tests/test_zzzz_multigpu_gpu.py tests/test_quasisep_gpu.py tests/test_zzy_quasisep_reference_gpu.py -m gpu -x -q -p no:cacheprovider > $O/r2mg2b_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2mg2b_pytest.log
tail -5 $O/r2mg2b_pytest.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512"
timeout 600 $TR bench.py --workload sharded --steps 2 --warmup 1 > $O/r2mg2b_sharded.json 2> $O/r2mg2b_sharded.err
tail -c 1200 $O/r2mg2b_sharded.json; tail -3 $O/r2mg2b_sharded.err
