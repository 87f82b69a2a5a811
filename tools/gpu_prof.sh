#!/bin/bash
# usage: tools/gpu_prof.sh <kernel regex> <out name> [extra bench args]
# ncu --set full capture of one launch of the named kernel inside a short bench run (256 clones).
K=$1; O=$2; shift 2
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$K -s 1 -c 1 -o gpurun_out/$O -f python bench.py --steps 1 --warmup 1 --cpu-evals 0 --attempts 1 --clones 256 "$@" > gpurun_out/$O.log 2>&1
tail -3 gpurun_out/$O.log
