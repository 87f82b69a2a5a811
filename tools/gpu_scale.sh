#!/bin/bash
# headline workload (C3, independent clones) at N GPUs given as arguments, back to back on one box
mkdir -p gpurun_out
for N in "$@"; do
  if [ "$N" = "1" ]; then
    timeout 600 python bench.py --steps 5 --warmup 3 --cpu-evals 0 > gpurun_out/scale_$N.log 2>&1
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2955$N bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/scale_$N.log 2>&1
  fi
  echo "N=$N rc=$?"; grep -E 'rror|Traceback' gpurun_out/scale_$N.log | head -3
done
