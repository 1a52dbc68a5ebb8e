#!/bin/bash
# First gpurun call of the next round (round 1 ran out of GPU minutes before these could be measured):
#   gpurun --timeout 3000 -- 'bash tools/r2_first_call.sh'        (~35-45 GPU-minutes: suite ~12, 15 quick bench runs ~12, two ncu captures ~10)
# 1. the whole GPU suite (new this round: tests/test_reference_golden.py -m gpu, b200gp_qs_condition)
# 2. digit-plane count vs the full-size LAPACK golden (N = 65536): is S = 6 / 7 inside the 1e-8 tolerance, and how fast
# 3. operand-traffic experiments on the int8 update (pairing / chunk-major layout / cluster shape); the kernel is bound
#    by the L2 -> SM stream (LTS cap ~6300 B/clk chip-wide = 43 B/clk/SM against the 96 B/clk the unpaired loop needs)
# 4. one ncu --set full capture of the best variant (dram bytes -> roofline.traffic)
set -x
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -p no:cacheprovider > $O/r2_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/r2_pytest_gpu.log
tail -3 $O/r2_pytest_gpu.log
for S in 8 7 6 5; do
  timeout 200 python bench.py --quick --steps 2 --warmup 2 --slices $S > $O/r2_bench_S$S.json 2> $O/r2_bench_S$S.err
done
for OPTS in "ozaki_pairing=1" "ozaki_pairing=1 ozaki_layout=1" "ozaki_layout=1" "ozaki_pairing=1 ozaki_cluster=11" "ozaki_pairing=1 ozaki_cluster=22" "ozaki_cluster=2" "ozaki_cluster=2 ozaki_pairing=1" "ozaki_cluster=2 ozaki_pairing=2" "ozaki_cluster=1" "panel_overlap=1" "build_ahead=1" "ozaki_pairing=1 panel_overlap=1 build_ahead=1"; do
  tag=$(echo "$OPTS" | tr ' =' '__')
  args=""; for o in $OPTS; do args="$args --opt $o"; done
  timeout 200 python bench.py --quick --steps 2 --warmup 2 --slices 7 $args > $O/r2_bench_S7_$tag.json 2> $O/r2_bench_S7_$tag.err
done
grep -h -o '"value": [0-9.]*\|"options": \[[^]]*\]\|"rel_err": [0-9.e-]*\|"digit_planes": [0-9]*\|"frac": [0-9.]*' $O/r2_bench_*.json | paste - - - - - | tee $O/r2_sweep_summary.txt
timeout 500 ncu --set full --clock-control none --import-source on -k regex:i8_update -s 40 -c 2 -o $O/r2_i8_full -f \
    python bench.py --steps 1 --warmup 1 --slices 7 --size 32768 > $O/r2_ncu.log 2>&1
ncu -i $O/r2_i8_full.ncu-rep --page raw --csv > $O/r2_i8_full_raw.csv 2>/dev/null
# 5. why is the CTA-pair kernel slow?  full capture of two of its launches (warp-state and memory tables)
timeout 400 ncu --set full --clock-control none --import-source on -k regex:i8_update_kernel_2sm -s 20 -c 2 -o $O/r2_i8_2sm_full -f \
    python bench.py --steps 1 --warmup 1 --slices 7 --size 16384 --opt ozaki_cluster=2 > $O/r2_ncu_2sm.log 2>&1
ncu -i $O/r2_i8_2sm_full.ncu-rep --page raw --csv > $O/r2_i8_2sm_full_raw.csv 2>/dev/null
# 6. quasisep (BASELINE config 4): default tree vs the warp-shuffle scan over the chunk composites
timeout 200 python bench.py --workload quasisep --steps 5 --warmup 3 > $O/r2_bench_qs.json 2> $O/r2_bench_qs.err
timeout 200 python bench.py --workload quasisep --steps 5 --warmup 3 --opt qs_tree=1 > $O/r2_bench_qs_tree1.json 2> $O/r2_bench_qs_tree1.err
