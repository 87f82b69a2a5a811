#!/bin/bash
# compute-sanitizer memcheck over a subset of the GPU tests (small systems; every kernel family)
mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 99 --print-limit 20 python -m pytest tests/test_gpu_parity.py tests/test_gpu_sweep.py -x -q -k "full_path_values or move_windows or estimators or tiled_gofr or gradient or perm_table or several_listed or device_sweep or displace or sharded_sweeps or large_path or fast_and_general or david or sharded_contexts" > gpurun_out/memcheck.log 2>&1
echo "memcheck rc=$?"
grep -E "ERROR SUMMARY|Invalid|out of bounds|passed|failed" gpurun_out/memcheck.log | head -20
