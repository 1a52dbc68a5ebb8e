#!/bin/bash
# 2-GPU call: NCCL test of the sharded path + sharded bench + the driver's own launch form
mkdir -p gpurun_out; O=gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > $O/r2mg2_smi.txt
timeout 900 python -m pytest tests/test_zzzz_multigpu_gpu.py -m gpu -x -q -p no:cacheprovider > $O/r2mg2_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2mg2_pytest.log
tail -5 $O/r2mg2_pytest.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR bench.py --workload sharded --size 65536 --steps 2 --warmup 1 > $O/r2mg2_sharded_64k.json 2> $O/r2mg2_sharded_64k.err
tail -c 1500 $O/r2mg2_sharded_64k.json
timeout 900 $TR bench.py --gpus 2 --steps 3 --warmup 2 > $O/r2mg2_bench_gpus2.json 2> $O/r2mg2_bench_gpus2.err
tail -c 2500 $O/r2mg2_bench_gpus2.json
tail -5 $O/r2mg2_bench_gpus2.err
