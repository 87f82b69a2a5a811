#!/bin/bash
# side-by-side evaluation of a context's actions (pimc_internal_evaluate_many): its parity test, then A/B of the C5 evaluation
# step at the size of an 8-GPU shard (M = 64 on one GPU) and at full size, sequential library (ab_libs/lib_seq.so) vs this one
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_sharded_evaluate.py -q > gpurun_out/pytest_many.log 2>&1; echo "many rc=$?"; tail -3 gpurun_out/pytest_many.log
for m in 64 512; do
for lib in ab_libs/lib_seq.so simpimc_b200/csrc/libsimpimc_b200.so; do
  SIMPIMC_B200_LIB=$PWD/$lib timeout 200 python bench.py --workload c5 --c5-m $m --steps 5 --warmup 3 --attempts 0 2>&1 | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l)
        print('M=$m', '$lib', 'graph %.4f ms eager %.4f ms' % (d.get('graph_ms_per_step', d['ms_per_step']), d['eager_ms_per_step']), 'match', d['energies_match'], 'graph_vs_eager', d['graph_replay_max_rel_diff_vs_eager'], d['roofline']['kernel_ms_per_step'])
    elif 'rror' in l: print('$lib', l.strip()[:300])
"
done
done
