#!/bin/bash
# A/B of the fused sweep's partner-bead layout and of the image branches (ab_libs/lib_m{0,1,2}i{0,1}.so)
for f in ab_libs/lib_m0i0.so ab_libs/lib_m0i1.so ab_libs/lib_m1i0.so ab_libs/lib_m2i1.so simpimc_b200/csrc/libsimpimc_b200.so; do
  SIMPIMC_B200_LIB=$PWD/$f timeout 120 python tools/time_sweep.py 2>&1 | tail -1
done
