#!/bin/bash
# config C5 (slice-sharded plasma) on N GPUs: usage tools/gpu_c5.sh N [extra bench args]
N=$1; shift
mkdir -p gpurun_out
if [ "$N" = "1" ]; then
  timeout 600 python bench.py --workload c5 --steps 5 --warmup 3 "$@" > gpurun_out/bench_c5_$N.log 2>&1
else
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --workload c5 --gpus $N --steps 5 --warmup 3 "$@" > gpurun_out/bench_c5_$N.log 2>&1
fi
echo "rc=$?"; grep -E '^\{|rror|Traceback' gpurun_out/bench_c5_$N.log | cut -c1-1800; tail -5 gpurun_out/bench_c5_$N.log | cut -c1-300
