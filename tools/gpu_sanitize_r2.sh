#!/bin/bash
# round 2: compute-sanitizer over the kernels that are new or changed this round (kinetic / FreeSpline tables, image
# branches of both sweep kernels, disjoint windows, the cp.async partner ring and fused decision of the displace path)
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 99 --print-limit 20 python -m pytest tests/test_gpu_kinetic.py tests/test_gpu_sweep.py -x -q -k "kinetic or images or multi_window or displace or device_sweep or sharded_sweeps" > gpurun_out/memcheck_r2.log 2>&1
echo "memcheck rc=$?"
grep -E "ERROR SUMMARY|Invalid|out of bounds|passed|failed" gpurun_out/memcheck_r2.log | head -10
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 99 --print-limit 10 python -m pytest tests/test_gpu_kinetic.py tests/test_gpu_sweep.py -x -q -k "images or multi_window or displace or device_sweep" > gpurun_out/racecheck_r2.log 2>&1
echo "racecheck rc=$?"
grep -E "RACECHECK SUMMARY|hazard|passed|failed|Error" gpurun_out/racecheck_r2.log | head -10
