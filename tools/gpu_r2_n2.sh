#!/bin/bash
# 2-GPU confirmation of the final build's bench line (side-by-side actions + NCCL all-reduce inside the captured step)
mkdir -p gpurun_out
timeout 65 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 5 --warmup 3 --sharded-timeout 40 > gpurun_out/bench_2.log 2>&1; echo "bench rc=$?"
grep -E 'rror|Traceback' gpurun_out/bench_2.log | head -5
python - <<'PY'
import json
for l in open('gpurun_out/bench_2.log'):
    if l.startswith('{'):
        d=json.loads(l); s=d.get('sharded',{})
        print('C3 n_gpus', d['n_gpus'], 'value %.4g e2e %.4g' % (d['value'], d['e2e']['value']))
        print('sharded', {k:s.get(k) for k in ('value','ms_per_step','step_mode','graph_ms_per_step','eager_ms_per_step','energies_match','graph_replay_max_rel_diff_vs_eager','error')})
PY
