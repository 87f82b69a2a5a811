#!/bin/bash
# 4-GPU confirmation of the line the driver's scaling run prints (clone replicas + the slice-sharded leg over the library's
# NCCL collectives); the permuting-bisection tests first (one GPU) for the last kernel change
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_perm.py -q > gpurun_out/pytest_perm.log 2>&1; echo "perm rc=$?"; tail -2 gpurun_out/pytest_perm.log
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 4 --steps 5 --warmup 3 > gpurun_out/bench_4.log 2>&1; echo "bench rc=$?"
grep -E 'rror|Traceback' gpurun_out/bench_4.log | head -5
python - <<'PY'
import json
for l in open('gpurun_out/bench_4.log'):
    if l.startswith('{'):
        d=json.loads(l); s=d.get('sharded',{})
        print('C3 n_gpus', d['n_gpus'], 'value %.4g e2e %.4g' % (d['value'], d['e2e']['value']), 'clocks', d['clocks'])
        print('sharded', {k:s.get(k) for k in ('value','ms_per_step','eager_ms_per_step','halo_bytes','energies_match','graph_replay_max_rel_diff_vs_eager','clocks','mc','error')})
        print('sharded e2e', s.get('e2e'), (s.get('roofline') or {}).get('kernel_ms_per_step'))
PY
