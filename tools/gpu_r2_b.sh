#!/bin/bash
# round 2: new tests (kinetic images, stat parity, David drop-in, asym tables, k-space growth, sweeps on the mirror) + quick bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kinetic.py tests/test_gpu_sweep.py tests/test_gpu_dropin.py tests/test_gpu_golden.py tests/test_gpu_stat_parity.py -m gpu -q --durations=8 > gpurun_out/pytest_b.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_b.log
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "asym or growth or full_path or per_pair" >> gpurun_out/pytest_b.log 2>&1; echo "pytest2 rc=$?" >> gpurun_out/pytest_b.log
grep -E "passed|failed|rc=|^E  |Error|slowest|s call" gpurun_out/pytest_b.log | head -60
timeout 600 python bench.py --steps 5 --warmup 3 --cpu-evals 0 --no-sharded > gpurun_out/bench_b.log 2>&1; echo "bench rc=$?"
python - <<'PY'
import json
for l in open('gpurun_out/bench_b.log'):
    if l.startswith('{'):
        d=json.loads(l); print('value %.4g e2e %.4g ms/step %.2f' % (d['value'], d['e2e']['value'], d['ms_per_step'])); print(d['roofline']['kernel_ms'], 'frac', d['roofline']['frac'])
        print('mc', {k:d['mc'][k] for k in ('ms_per_attempt','accept_ratio','launches')}, 'sweeps/s', d['mc_sweeps_per_s'], 'displace', d['mc']['displace']['ms_per_attempt'])
    elif 'rror' in l: print(l.strip()[:300])
PY
