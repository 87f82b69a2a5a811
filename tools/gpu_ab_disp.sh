#!/bin/bash
# A/B timing of the displace kernel for library builds in ab_libs/lib_disp*.so on one box, then the displace tests per build.
for lib in ab_libs/lib_disp*.so; do
  SIMPIMC_B200_LIB=$PWD/$lib timeout 300 python bench.py --steps 1 --warmup 3 --cpu-evals 0 --attempts 16 --pipeline 1 2>&1 | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('$lib', 'displace ms/attempt %.4f' % d['mc']['displace']['ms_per_attempt'], 'accept %.4f' % d['mc']['displace']['accept_ratio'])
    elif 'rror' in l: print('$lib', l.strip()[:200])
"
  SIMPIMC_B200_LIB=$PWD/$lib timeout 600 python -m pytest tests/test_gpu_sweep.py -x -q -k displace 2>&1 | tail -1
done
