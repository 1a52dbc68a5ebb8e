#!/bin/bash
# two GPUs: sharded-path tests (tail split-K on / off) and the config-3 line with and without the split on every rank's rows
mkdir -p gpurun_out; O=gpurun_out
timeout 900 python -m pytest tests/test_zzzz_multigpu_gpu.py -m gpu -x -q -p no:cacheprovider > $O/r2mg_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2mg_pytest.log
tail -5 $O/r2mg_pytest.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512"
timeout 600 $TR bench.py --workload sharded --steps 2 --warmup 1 > $O/r2mg_sharded_split1.json 2> $O/r2mg_sharded_split1.err
timeout 600 $TR bench.py --workload sharded --steps 2 --warmup 1 --opt mg_splitk=0 > $O/r2mg_sharded_split0.json 2> $O/r2mg_sharded_split0.err
for f in $O/r2mg_sharded_split1.json $O/r2mg_sharded_split0.json; do grep -h -o '"ms_per_step": [0-9.]*\|"rel_err": [0-9.e-]*\|"kernel_ms_per_step_rank0": {[^}]*}\|"logp": [-0-9.e]*' $f | tr '\n' ' '; echo; done
tail -2 $O/r2mg_sharded_split1.err
