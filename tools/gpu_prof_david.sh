#!/bin/bash
# ncu --set full capture of one launch of the fast David whole-path kernel (C3 shape, 256 clones)
mkdir -p gpurun_out
TIME_DAVID_FIRST_ONLY=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:david_full_fast -s 1 -c 1 -o gpurun_out/prof_david -f python tools/time_david.py > gpurun_out/prof_david.log 2>&1
tail -3 gpurun_out/prof_david.log
