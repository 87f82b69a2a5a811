#!/bin/bash
# full GPU pass: every GPU test, smoke, the default bench line (cpu baseline, parity block, sharded leg)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q --durations=6 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
grep -E "passed|failed|rc=|^E  |Error|s call" gpurun_out/pytest_gpu.log | head -40
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log; tail -2 gpurun_out/smoke.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.log 2>&1; echo "bench rc=$?" >> gpurun_out/bench.log
python - <<'PY'
import json
for l in open('gpurun_out/bench.log'):
    if l.startswith('{'):
        d=json.loads(l); s=d.get('sharded',{})
        print('value %.4g e2e %.4g ms/step %.2f frac %.3f' % (d['value'], d['e2e']['value'], d['ms_per_step'], d['roofline']['frac']))
        print('mc', d['mc_sweeps_per_s'], d['mc']['ms_per_attempt'], 'displace', d['mc']['displace']['ms_per_attempt'], 'parity', d.get('parity'))
        print('sharded', {k:s.get(k) for k in ('value','ms_per_step','eager_ms_per_step','energies_match','graph_replay_max_rel_diff_vs_eager','mc','error')})
    elif 'rror' in l or 'rc=' in l: print(l.strip()[:300])
PY
