#!/bin/bash
# last pass of round 2: every GPU test, smoke and the default bench line of the final build
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q --durations=6 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
grep -E "passed|failed|rc=|^E  |Error" gpurun_out/pytest_gpu.log | head -30
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log; tail -2 gpurun_out/smoke.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.log 2>&1; echo "bench rc=$?" >> gpurun_out/bench.log
python - <<'PY'
import json
for l in open('gpurun_out/bench.log'):
    if l.startswith('{'):
        d=json.loads(l); s=d.get('sharded',{}) or {}
        print('value %.4g e2e %.4g ms/step %.3f' % (d['value'], d['e2e']['value'], d['ms_per_step']), 'frac', d['roofline']['frac'], 'clocks', d['clocks'])
        print('mc', d['mc_sweeps_per_s'], d['mc']['ms_per_attempt'], 'displace', d['mc']['displace']['ms_per_attempt'], 'perm', d['mc']['perm_bisect']['ms_per_attempt'], 'parity', d.get('parity'))
        print('sharded', {k:s.get(k) for k in ('value','ms_per_step','step_mode','graph_ms_per_step','eager_ms_per_step','energies_match','error')})
    elif 'rror' in l or 'rc=' in l: print(l.strip()[:300])
PY
