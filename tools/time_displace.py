#!/usr/bin/env python
"""ms per DisplaceParticle attempt of the device-resident sweep at the headline shape (UEG N=256, M=128, 1024 clones):
   SIMPIMC_B200_LIB=<variant .so> python tools/time_displace.py"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from simpimc_b200 import host, system as S  # noqa: E402

C, n_att = 1024, 16
cfg = S.ueg_config(N=256, M=128)
path = host.Path(cfg, n_clones=C)
R0 = S.synthetic_paths(cfg, 0, 0)
rng = np.random.default_rng(1)
R = np.stack([np.roll(R0, int(rng.integers(0, 128)), axis=1)[rng.permutation(256)] + rng.uniform(-1, 1, 3) for _ in range(C)])
path.SetPositions(0, R)
path.DisplaceSweep(0, cfg.L / 10.0, 2, 7, attempt0=0)
path.Sync()
best, acc = 1e9, None
for rep in range(3):
    t0 = time.perf_counter()
    acc = path.DisplaceSweep(0, cfg.L / 10.0, n_att, 7, attempt0=2 + rep * n_att)
    path.Sync()
    best = min(best, (time.perf_counter() - t0) * 1e3 / n_att)
print("%-28s displace ms/attempt %.4f  pair evals/s %.3e  accept %.3f" % (os.path.basename(os.environ.get("SIMPIMC_B200_LIB", "default")), best,
                                                                          C * 2 * 255 * 128 / (best * 1e-3), acc.sum() / (C * n_att)), flush=True)
path.close()
