#!/bin/bash
# ncu --set full capture of the first fused-sweep launch (8 warm-up attempts) at the bench's clone count.
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:bisect_sweep_fused -c 1 -o gpurun_out/prof_sweep -f python bench.py --steps 1 --warmup 1 --cpu-evals 0 --attempts 8 --pipeline 1 "$@" > gpurun_out/prof_sweep.log 2>&1
tail -3 gpurun_out/prof_sweep.log
