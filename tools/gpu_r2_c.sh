#!/bin/bash
# round 2: the tests the -x stop of the full pass did not reach, the adapter timing, and the ncu evidence of the final build
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_sweep.py "tests/test_gpu_stat_parity.py::test_sampled_egas_run_matches_the_reference_program[33-800-80]" -m gpu -q --durations=5 > gpurun_out/pytest_c.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_c.log
grep -E "passed|failed|rc=|^E  |Error|s call" gpurun_out/pytest_c.log | head -40
timeout 300 python tools/time_dropin.py > gpurun_out/time_dropin.log 2>&1; tail -3 gpurun_out/time_dropin.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 1 --cpu-evals 0 --attempts 4 > gpurun_out/bench_ncu.log 2>&1; echo "ncu list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:pair_full_fast -s 1 -c 1 -o gpurun_out/r2_prof_k1 -f python bench.py --steps 1 --warmup 1 --cpu-evals 0 --attempts 1 --clones 256 --no-sharded > gpurun_out/prof_k1.log 2>&1; echo "ncu k1 rc=$?"
ls -la gpurun_out | tail -8
