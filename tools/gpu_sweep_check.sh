#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_sweep.py tests/test_gpu_parity.py -x -q 2>&1 | tail -5
timeout 600 python bench.py --steps 3 --warmup 3 --cpu-evals 0 > gpurun_out/bench_mc.log 2>&1
python - <<'PY'
import json
for l in open('gpurun_out/bench_mc.log'):
    if l.startswith('{'):
        d=json.loads(l); print('value %.4g  sweeps/s %.1f' % (d['value'], d['mc_sweeps_per_s'])); print(d['mc'])
    elif 'rror' in l: print(l.strip())
PY
