"""Times the whole-path evaluations of the other two action families at the C3 shape (256 clones),
through the shared-memory fast kernels and (David) through the general kernel."""
import os, sys, time
import numpy as np
sys.path.insert(0, ".")
from simpimc_b200 import host, system as S
for action, lr, kw in (("DavidPairAction", False, {}), ("DavidPairAction", False, {"david_grid": "LINEAR", "david_n_grid": 400}),
                       ("BarePairAction", True, {}))[:1 if os.environ.get("TIME_DAVID_FIRST_ONLY") else 3]:
    cfg = S.ueg_config(N=256, M=128, action=action, use_long_range=lr, **kw)
    C = 256
    path = host.Path(cfg, n_clones=C)
    path.SetPositions(0, np.stack([S.synthetic_paths(cfg, 0, c) for c in range(C)]))
    act = path.actions[0]
    for general in ((False, True) if action == "DavidPairAction" else (False,)):
        path.ForceGeneral(general)
        for name, f in (("DActionDBeta", act.DActionDBeta), ("Potential", act.Potential), ("TotalAction", act.TotalAction)):
            f(); path.Sync()
            path.SetTiming(True)
            f(); f(); path.Sync()
            k1, n = path.KernelTime(1)
            path.SetTiming(False)
            print("%s %s %s %s: K1 kernel %.2f ms per 256 clones (x4 for 1024: %.1f ms)" % (
                action, kw, "general" if general else "fast", name, k1 / max(n, 1), 4 * k1 / max(n, 1)))
    path.close()
