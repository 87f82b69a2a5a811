#!/bin/bash
# round 2, final build: every GPU test, smoke, compute-sanitizer over the permuting-bisection kernels, the default bench
# line (cpu baseline, parity block, sharded leg, mc block), the reference arm, the ncu launch list and one --set full
# capture of K1 (bucket-centred long-range tables)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q --durations=6 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
grep -E "passed|failed|rc=|^E  |Error" gpurun_out/pytest_gpu.log | head -30
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log; tail -2 gpurun_out/smoke.log
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 99 --print-limit 20 python -m pytest tests/test_gpu_perm.py -x -q -k "mirror or permuted" > gpurun_out/memcheck_perm.log 2>&1
echo "memcheck rc=$?"; grep -E "ERROR SUMMARY|Invalid|out of bounds|passed|failed" gpurun_out/memcheck_perm.log | head -10
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 99 --print-limit 10 python -m pytest tests/test_gpu_perm.py -x -q -k "mirror" > gpurun_out/racecheck_perm.log 2>&1
echo "racecheck rc=$?"; grep -E "RACECHECK SUMMARY|hazard|passed|failed|Error" gpurun_out/racecheck_perm.log | head -10
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.log 2>&1; echo "bench rc=$?" >> gpurun_out/bench.log
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.log 2>&1; echo "bench ref rc=$?" >> gpurun_out/bench_ref.log
python - <<'PY'
import json
for fn in ('gpurun_out/bench.log', 'gpurun_out/bench_ref.log'):
    for l in open(fn):
        if l.startswith('{'):
            d=json.loads(l); s=d.get('sharded',{}) or {}
            print(fn, 'value %.4g e2e %.4g ms/step %.3f' % (d['value'], d['e2e']['value'], d['ms_per_step']), 'frac', (d.get('roofline') or {}).get('frac'))
            if 'mc' in d:
                print('mc', d['mc_sweeps_per_s'], d['mc']['ms_per_attempt'], 'displace', d['mc']['displace']['ms_per_attempt'], 'perm', d['mc'].get('perm_bisect',{}).get('ms_per_attempt'), 'parity', d.get('parity'))
                print('sharded', {k:s.get(k) for k in ('value','ms_per_step','eager_ms_per_step','energies_match','error')})
        elif 'rror' in l or 'rc=' in l: print(l.strip()[:300])
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_final_launches.csv python bench.py --steps 2 --warmup 1 --cpu-evals 0 --attempts 4 --no-sharded > gpurun_out/bench_ncu.log 2>&1; echo "ncu list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:pair_full_fast -s 1 -c 1 -o gpurun_out/r2_prof_k1_v8 -f python bench.py --steps 1 --warmup 1 --cpu-evals 0 --attempts 1 --clones 256 --no-sharded > gpurun_out/prof_k1_v8.log 2>&1; echo "ncu k1 rc=$?"
ls -la gpurun_out | tail -6
