#!/bin/bash
# sweep tests + a short bench (MC figures) on the GPU box
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_sweep.py -x -q > gpurun_out/pytest_sweep.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_sweep.log
tail -25 gpurun_out/pytest_sweep.log
timeout 600 python bench.py --steps 3 --warmup 3 --cpu-evals 0 --attempts 256 > gpurun_out/bench_sweep.log 2>&1; echo "bench rc=$?" >> gpurun_out/bench_sweep.log
python - <<'PY'
import json
for l in open('gpurun_out/bench_sweep.log'):
    if l.startswith('{'):
        d=json.loads(l); print('value %.4g e2e %.4g ms/step %.2f sweeps/s %.1f' % (d['value'], d['e2e']['value'], d['ms_per_step'], d['mc_sweeps_per_s'])); print(d['mc'])
    elif 'rc=' in l or 'rror' in l: print(l.strip()[:300])
PY
