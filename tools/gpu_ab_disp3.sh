#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_sweep.py tests/test_gpu_kinetic.py -m gpu -q -x 2>&1 | tail -3
for f in ab_libs/lib_d0.so simpimc_b200/csrc/libsimpimc_b200.so; do
  SIMPIMC_B200_LIB=$PWD/$f timeout 120 python tools/time_displace.py 2>&1 | tail -1
done
