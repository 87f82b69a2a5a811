#!/bin/bash
# A/B timing of the fused sweep for library builds in ab_libs/*.so on one box, then the sweep tests per build.
for lib in ab_libs/lib_teams*.so; do
  SIMPIMC_B200_LIB=$PWD/$lib timeout 300 python bench.py --steps 1 --warmup 3 --cpu-evals 0 --attempts 256 --pipeline 1 2>&1 | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('$lib', 'sweeps/s %.0f' % d['mc_sweeps_per_s'], 'ms/attempt %.4f' % d['mc']['ms_per_attempt'])
    elif 'rror' in l: print('$lib', l.strip()[:200])
"
  SIMPIMC_B200_LIB=$PWD/$lib timeout 600 python -m pytest tests/test_gpu_sweep.py -x -q 2>&1 | tail -1
done
