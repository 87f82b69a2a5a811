"""Times device-resident bisection and displacement sweeps of a David-action system (UEG N=256, M=128,
256 clones; kernel-per-phase path) with the pp-form fast tables (default) and with the general
B-spline evaluation (ForceGeneral)."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
from simpimc_b200 import host, system as S
cfg = S.ueg_config(N=256, M=128, action="DavidPairAction", use_long_range=True)
C = 256
path = host.Path(cfg, n_clones=C)
R = np.stack([S.synthetic_paths(cfg, 0, c, 3) for c in range(C)])
for general in (False, True):
    path.SetPositions(0, R)
    path.ForceGeneral(general)
    path.BisectSweep(0, 3, 8, 5, attempt0=0)
    path.Sync()
    t0 = time.perf_counter()
    n = 64
    acc = path.BisectSweep(0, 3, n, 5, attempt0=8)
    path.Sync()
    t1 = time.perf_counter()
    path.DisplaceSweep(0, cfg.L / 10, 1, 7, attempt0=0)
    path.Sync()
    t2 = time.perf_counter()
    nd = 8
    accd = path.DisplaceSweep(0, cfg.L / 10, nd, 7, attempt0=1)
    path.Sync()
    t3 = time.perf_counter()
    print("David UEG N=256 M=128, %d clones, %s: bisect %.4f ms per attempt (accept %.3f), displace %.3f ms per attempt (accept %.3f)" % (
        C, "general B-spline evaluation" if general else "pp-form fast tables", 1e3 * (t1 - t0) / n, acc.sum() / (C * n),
        1e3 * (t3 - t2) / nd, accd.sum() / (C * nd)))
path.close()
