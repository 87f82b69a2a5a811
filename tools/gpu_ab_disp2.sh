#!/bin/bash
# displace: stream-exact tests of the cp.async-staged kernel, then A/B of the staging depth
timeout 600 python -m pytest tests/test_gpu_sweep.py -m gpu -q -k "displace" 2>&1 | tail -3
for f in ab_libs/lib_d0.so ab_libs/lib_d2.so simpimc_b200/csrc/libsimpimc_b200.so ab_libs/lib_d4.so; do
  SIMPIMC_B200_LIB=$PWD/$f timeout 120 python tools/time_displace.py 2>&1 | tail -1
done
