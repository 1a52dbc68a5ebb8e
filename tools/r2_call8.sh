#!/bin/bash
mkdir -p gpurun_out; O=gpurun_out
run() { tag=$1; shift; timeout 240 python bench.py --quick --steps 3 --warmup 2 "$@" > $O/r2c8_$tag.json 2> $O/r2c8_$tag.err; }
run sub0_la --opt ozaki_subpanel=0 --opt panel_overlap=2
run sub256 
run sub256_la --opt panel_overlap=2
run sub512_la --opt ozaki_subpanel=512 --opt panel_overlap=2
run sub512 --opt ozaki_subpanel=512
grep -h -o '"value": [0-9.]*\|"options": \[[^]]*\]\|"rel_err": [0-9.e-]*\|"frac": [0-9.]*\|"kernel_ms_per_step": {[^}]*}' $O/r2c8_*.json | paste - - - - - > $O/r2c8_sweep_summary.txt
cat $O/r2c8_sweep_summary.txt
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:potf2 -c 40 --csv --log-file $O/r2c8_potf2.csv \
    python bench.py --quick --steps 1 --warmup 0 --size 16384 > $O/r2c8_potf2.log 2>&1
tail -3 $O/r2c8_potf2.csv
