#!/bin/bash
# multi-GPU pass: usage tools/gpu_r2_multi.sh N  (sharded check over the library's NCCL collectives, then the bench line at N GPUs)
N=$1
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/sharded_check.py > gpurun_out/sharded_check_$N.log 2>&1; echo "check rc=$?"
grep -E '^\{|rror|Traceback' gpurun_out/sharded_check_$N.log | cut -c1-1500; tail -3 gpurun_out/sharded_check_$N.log | cut -c1-300
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_$N.log 2>&1; echo "bench rc=$?"
grep -E 'rror|Traceback' gpurun_out/bench_$N.log | head -5
python - <<PY
import json
for l in open('gpurun_out/bench_$N.log'):
    if l.startswith('{'):
        d=json.loads(l); s=d.get('sharded',{})
        print('C3 value %.4g e2e %.4g' % (d['value'], d['e2e']['value']))
        print('sharded', {k:s.get(k) for k in ('value','ms_per_step','eager_ms_per_step','halo_bytes','energies_match','graph_replay_max_rel_diff_vs_eager','clocks','mc')})
        print('sharded e2e', s.get('e2e'), s.get('roofline',{}).get('kernel_ms_per_step'))
PY
