#!/bin/bash
# GPU pass for the David fast kernel: parity tests + timings
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -25 gpurun_out/pytest_gpu.log
timeout 300 python tools/time_david.py > gpurun_out/time_david.log 2>&1; echo "rc=$?" >> gpurun_out/time_david.log
cat gpurun_out/time_david.log
