#!/bin/bash
# ncu --set full captures of the displace pair kernel (bench.py, 256 clones) and of the Bare whole-path / fast Potential kernels
mkdir -p gpurun_out
bash tools/gpu_prof.sh displace_pair prof_displace
timeout 900 ncu --set full --clock-control none --import-source on -k regex:bare_full_fast -s 1 -c 1 -o gpurun_out/prof_bare -f python tools/time_david.py > gpurun_out/prof_bare.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:potential_fast -s 1 -c 1 -o gpurun_out/prof_potential -f python tools/time_david.py > gpurun_out/prof_potential.log 2>&1
ls -la gpurun_out/*.ncu-rep
