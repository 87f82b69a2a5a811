#!/bin/bash
# A/B timing of library builds in ab_libs/*.so on one box (K1 time from a short bench run each), then the parity tests on the default build.
for lib in ab_libs/*.so; do
  SIMPIMC_B200_LIB=$PWD/$lib timeout 300 python bench.py --steps 5 --warmup 3 --cpu-evals 0 --attempts 8 --pipeline 1 2>&1 | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('$lib', 'K1 %.2f ms' % d['roofline']['kernel_ms']['K1_pair_full'], 'step %.2f' % d['ms_per_step'], 'sweeps/s %.0f' % d['mc_sweeps_per_s'])
    elif 'rror' in l: print('$lib', l.strip()[:200])
"
done
SIMPIMC_B200_LIB=$PWD/ab_libs/lib_exp0.so timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
