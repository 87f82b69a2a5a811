"""Times the kernel-per-phase sweep path (any action mix) on the GPU: UEG C3 with the one-launch kernel
disabled is not possible without ForceGeneral (which also disables the fast tables), so a two-species
plasma of the same size class is used: 128 e + 128 p, M = 128, 1024 clones."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
from simpimc_b200 import host, system as S
cfg = S.plasma_config(Ne=128, Np=128, M=128, n_xy=100, n_r_long=1000, pp_action="IlkkaPairAction")
C = 1024
path = host.Path(cfg, n_clones=C)
for sp in range(2):
    path.SetPositions(sp, np.stack([S.synthetic_paths(cfg, sp, c, 3) for c in range(C)]))
path.BisectSweep(0, 3, 8, 5, attempt0=0)
path.Sync()
path.SetTiming(True)
t0 = time.perf_counter()
n = 64
acc = path.BisectSweep(0, 3, n, 5, attempt0=8)
path.Sync()
t1 = time.perf_counter()
k4, n4 = path.KernelTime(4)
k3, n3 = path.KernelTime(3)
print("plasma 128+128, M=128, 1024 clones: %.4f ms per attempt (species e: 2 pair actions), window kernels %.4f ms, lr %.4f ms, accept %.3f" % (1e3 * (t1 - t0) / n, k4 / n, k3 / n, acc.sum() / (C * n)))
