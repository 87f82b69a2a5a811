#!/bin/bash
# DisplaceParticle with the rho_k deltas recomputed in the commit (no drho round trip): its stream-exact tests, the timing,
# then every other GPU test
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_sweep.py -q -k "displace" > gpurun_out/pytest_disp.log 2>&1; echo "displace tests rc=$?"; tail -2 gpurun_out/pytest_disp.log
timeout 120 python tools/time_displace.py 2>&1 | tail -1
timeout 600 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_stat_parity.py > gpurun_out/pytest_rest.log 2>&1; echo "rest rc=$?"; tail -2 gpurun_out/pytest_rest.log
