#!/bin/bash
# after the batched-epilogue fix: micro-benchmarks (compact and factorisation-like strides) + full-size quick bench of the candidates
mkdir -p gpurun_out; O=gpurun_out
python tools/i8_microbench.py 32768 1024 32768 > $O/r2c3_micro_32k.jsonl 2> $O/r2c3_micro_32k.err
python tools/i8_microbench.py 16384 1024 32768 65536 65536 > $O/r2c3_micro_ld64k.jsonl 2> $O/r2c3_micro_ld64k.err
run() { tag=$1; shift; timeout 240 python bench.py --quick --steps 2 --warmup 1 --slices 7 "$@" > $O/r2c3_$tag.json 2> $O/r2c3_$tag.err; }
run base
run pair1 --opt ozaki_pairing=1
run cl2_pair1 --opt ozaki_cluster=2 --opt ozaki_pairing=1
run cl2 --opt ozaki_cluster=2
run cl1 --opt ozaki_cluster=1
grep -h -o '"value": [0-9.]*\|"options": \[[^]]*\]\|"rel_err": [0-9.e-]*\|"frac": [0-9.]*\|"kernel_ms_per_step": {[^}]*}' $O/r2c3_*.json | paste - - - - - > $O/r2c3_sweep_summary.txt
cat $O/r2c3_sweep_summary.txt
