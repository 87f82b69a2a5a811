#!/usr/bin/env python
"""ms per bisection attempt of the device-resident sweep at the headline shape (UEG N=256, M=128, 1024 clones, n_level=3):
   SIMPIMC_B200_LIB=<variant .so> python tools/time_sweep.py [clones] [attempts]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from simpimc_b200 import host, system as S  # noqa: E402

C = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
n_att = int(sys.argv[2]) if len(sys.argv) > 2 else 256
cfg = S.ueg_config(N=256, M=128)
path = host.Path(cfg, n_clones=C)
R0 = S.synthetic_paths(cfg, 0, 0)
rng = np.random.default_rng(1)
R = np.stack([np.roll(R0, int(rng.integers(0, 128)), axis=1)[rng.permutation(256)] + rng.uniform(-1, 1, 3) for _ in range(C)])
path.SetPositions(0, R)
path.BisectSweep(0, 3, 16, 99, attempt0=0)
path.Sync()
best = 1e9
acc = None
for rep in range(3):
    t0 = time.perf_counter()
    acc = path.BisectSweep(0, 3, n_att, 99, attempt0=16 + rep * n_att)
    path.Sync()
    best = min(best, (time.perf_counter() - t0) * 1e3 / n_att)
print("%-40s ms/attempt %.4f  clone-sweeps/s %.0f  accept %.3f" % (os.path.basename(os.environ.get("SIMPIMC_B200_LIB", "default")), best,
                                                                     C / (256 * 128 / 8) / (best * 1e-3), acc.sum() / (C * n_att)), flush=True)
path.close()
