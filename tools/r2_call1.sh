#!/bin/bash
# Round-2 first GPU call: variant sweep of the int8 update, ncu captures (i8 update, CTA-pair kernel, K1 build, quasisep)
set -x
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/r2c1_smi.txt
run() { tag=$1; shift; timeout 240 python bench.py --quick --steps 2 --warmup 2 --slices 7 "$@" > $O/r2c1_$tag.json 2> $O/r2c1_$tag.err; }
run base
run pair1 --opt ozaki_pairing=1
run pair1_lay1 --opt ozaki_pairing=1 --opt ozaki_layout=1
run lay1 --opt ozaki_layout=1
run cl2 --opt ozaki_cluster=2
run cl2_pair1 --opt ozaki_cluster=2 --opt ozaki_pairing=1
run cl1 --opt ozaki_cluster=1
run cl11 --opt ozaki_cluster=11
run pair1_cl11 --opt ozaki_pairing=1 --opt ozaki_cluster=11
run povl --opt panel_overlap=1
run bahead --opt build_ahead=1
run nb512 --nb 512
run nb2048 --nb 2048
grep -h -o '"value": [0-9.]*\|"options": \[[^]]*\]\|"rel_err": [0-9.e-]*\|"frac": [0-9.]*\|"kernel_ms_per_step": {[^}]*}' $O/r2c1_*.json | paste - - - - - | tee $O/r2c1_sweep_summary.txt
timeout 500 ncu --set full --clock-control none --import-source on -k regex:i8_update -s 40 -c 2 -o $O/r2c1_i8_full -f \
    python bench.py --quick --steps 1 --warmup 1 --slices 7 > $O/r2c1_ncu.log 2>&1
ncu -i $O/r2c1_i8_full.ncu-rep --page raw --csv > $O/r2c1_i8_full_raw.csv 2>/dev/null
timeout 400 ncu --set full --clock-control none --import-source on -k regex:i8_update_kernel_2sm -s 20 -c 2 -o $O/r2c1_i8_2sm_full -f \
    python bench.py --quick --steps 1 --warmup 1 --slices 7 --size 16384 --opt ozaki_cluster=2 > $O/r2c1_ncu_2sm.log 2>&1
ncu -i $O/r2c1_i8_2sm_full.ncu-rep --page raw --csv > $O/r2c1_i8_2sm_full_raw.csv 2>/dev/null
timeout 300 ncu --set full --clock-control none --import-source on -k regex:build_rect -c 2 -o $O/r2c1_build_full -f \
    python tools/k1_build.py > $O/r2c1_ncu_build.log 2>&1
ncu -i $O/r2c1_build_full.ncu-rep --page raw --csv > $O/r2c1_build_full_raw.csv 2>/dev/null
timeout 200 python bench.py --workload quasisep --steps 5 --warmup 3 > $O/r2c1_qs.json 2> $O/r2c1_qs.err
timeout 200 python bench.py --workload quasisep --steps 5 --warmup 3 --opt qs_tree=1 > $O/r2c1_qs_tree1.json 2> $O/r2c1_qs_tree1.err
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'chol_chunk|chol_replay|tree_' -c 8 -o $O/r2c1_qs_full -f \
    python bench.py --workload quasisep --steps 1 --warmup 0 > $O/r2c1_ncu_qs.log 2>&1
ncu -i $O/r2c1_qs_full.ncu-rep --page raw --csv > $O/r2c1_qs_full_raw.csv 2>/dev/null
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $O/r2c1_launches_dense.csv \
    python bench.py --quick --steps 1 --warmup 0 --slices 7 > $O/r2c1_launches_dense.log 2>&1
gzip -f $O/r2c1_launches_dense.csv
ls -la $O
