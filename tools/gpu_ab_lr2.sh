#!/bin/bash
# A/B on one box: the variant builds in ab_libs/*.so (e.g. -DPIMC_LR2=0: 1-D splines in interval form; -DPIMC_XY16=0: x / y
# intervals through the knot-pair compare) against the default build; then every GPU test on the default build
mkdir -p gpurun_out
for lib in ab_libs/*.so simpimc_b200/csrc/libsimpimc_b200.so; do
  SIMPIMC_B200_LIB=$PWD/$lib timeout 300 python bench.py --steps 5 --warmup 3 --cpu-evals 0 --attempts 64 --no-sharded --pipeline 1 2>&1 | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); m=d['mc']
        print('$lib', 'K1 %.3f ms' % d['roofline']['kernel_ms']['K1_pair_full'], 'step %.3f' % d['ms_per_step'], 'frac %.4f' % d['roofline']['frac'], 'sweep ms/attempt %.5f' % m['ms_per_attempt'], 'displace %.4f' % m['displace']['ms_per_attempt'], 'perm', m.get('perm_bisect'))
    elif 'rror' in l: print('$lib', l.strip()[:300])
"
done
timeout 900 python -m pytest tests -m gpu -q --durations=5 > gpurun_out/pytest_lr2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_lr2.log
grep -E "passed|failed|rc=|^E  |Error|^FAILED" gpurun_out/pytest_lr2.log | head -40
