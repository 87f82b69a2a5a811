#!/bin/bash
# compute-sanitizer racecheck (shared-memory hazards) over the kernels with team barriers and staged tiles
mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 99 --print-limit 10 python -m pytest tests/test_gpu_sweep.py tests/test_gpu_parity.py -x -q -k "device_sweep or displace or full_path_values or tiled_gofr or large_path or fast_and_general or david" > gpurun_out/racecheck.log 2>&1
echo "racecheck rc=$?"
grep -E "RACECHECK SUMMARY|hazard|passed|failed|Error" gpurun_out/racecheck.log | head -20
