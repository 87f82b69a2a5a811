import time, numpy as np, sys
sys.path.insert(0, '.')
from simpimc_b200 import host, system as S
cfg = S.ueg_config(N=256, M=128)
C = 256
path = host.Path(cfg, n_clones=C)
R = np.stack([S.synthetic_paths(cfg, 0, c) for c in range(C)])
path.SetPositions(0, R)
act = path.actions[0]
for name, f in (("DActionDBeta", act.DActionDBeta), ("Potential", act.Potential), ("TotalAction", act.TotalAction)):
    f(); path.Sync()
    path.SetTiming(True)
    t0 = time.perf_counter(); f(); path.Sync(); t1 = time.perf_counter()
    k1, n = path.KernelTime(1)
    path.SetTiming(False)
    print(name, "wall %.2f ms" % (1e3*(t1-t0)), "K1 kernel %.2f ms (x4 for 1024 clones: %.1f)" % (k1/max(n,1), 4*k1/max(n,1)))
