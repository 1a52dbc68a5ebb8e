#!/bin/bash
mkdir -p gpurun_out; O=gpurun_out
for nb in 2048 4096; do
  timeout 300 python bench.py --workload batched --steps 2 --warmup 1 --opt nb_batched=$nb > $O/r2c23_batched_$nb.json 2> $O/r2c23_batched_$nb.err
  echo "nb_batched=$nb $(grep -h -o '"value": [0-9.]*\|"ms_per_step": [0-9.]*\|"max_rel_err_vs_oracle_on_grid_corners": [0-9.e-]*' $O/r2c23_batched_$nb.json | tr '\n' ' ')"
done
