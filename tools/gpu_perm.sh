#!/bin/bash
# device-resident permuting bisection: its own tests first, then every other GPU test (the proposal slots grew to 16)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_perm.py -q --durations=5 > gpurun_out/pytest_perm.log 2>&1; echo "perm rc=$?" >> gpurun_out/pytest_perm.log
grep -E "passed|failed|rc=|^E  |Error" gpurun_out/pytest_perm.log | head -60
timeout 900 python -m pytest tests -m gpu -q --deselect tests/test_gpu_perm.py --durations=5 > gpurun_out/pytest_rest.log 2>&1; echo "rest rc=$?" >> gpurun_out/pytest_rest.log
grep -E "passed|failed|rc=|^E  |Error|^FAILED" gpurun_out/pytest_rest.log | head -40
