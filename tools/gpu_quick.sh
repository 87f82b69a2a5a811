#!/bin/bash
# quick GPU pass: parity tests + bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 5 --warmup 3 --cpu-evals 0 --attempts 4 > gpurun_out/bench_quick.log 2>&1; echo "bench rc=$?" >> gpurun_out/bench_quick.log
python - <<'PY'
import json
for l in open('gpurun_out/bench_quick.log'):
    if l.startswith('{'):
        d=json.loads(l); print('value %.4g e2e %.4g ms/step %.2f' % (d['value'], d['e2e']['value'], d['ms_per_step'])); print(d['roofline']['kernel_ms'], 'frac', d['roofline']['frac'], 'peak', d['roofline']['peak'])
    elif 'rc=' in l or 'Error' in l or 'error' in l: print(l.strip())
PY
