#!/usr/bin/env python
"""Moves per second of the REFERENCE'S OWN Bisect::DoEvent (compiled in place) on its own IlkkaPairAction versus on
GpuPairAction (include/simpimc_b200_action.hpp: one walker, one synchronous C-ABI call per GetAction) -- the number
INTEGRATION.md quotes for a maintainer who only swaps the action class."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from simpimc_b200 import system as S  # noqa: E402
from oracle import refsim  # noqa: E402

for N, M in ((33, 64), (256, 128)):
    cfg = S.ueg_config(N=N, M=M, with_kinetic=True)
    cfg.moves = [{"name": "BisectE", "type": "Bisect", "species": "e", "n_level": 3}]
    cfg.observables = []
    out = {}
    for name, kw in (("reference", dict(fast=refsim.available(fast=True))), ("adapter", dict(dropin=True))):
        sim = refsim.RefSim(cfg, seed=3, **kw)
        sim.set_positions(0, S.synthetic_paths(cfg, 0, 0, 5))
        sim.move_do(0, 20)
        n = 300
        t0 = time.perf_counter()
        sim.move_do(0, n)
        out[name] = n / (time.perf_counter() - t0)
        t0 = time.perf_counter()
        for _ in range(5):
            sim.dbeta(1)
        out[name + "_dbeta_ms"] = (time.perf_counter() - t0) / 5 * 1e3
        sim.close()
    print("N=%d M=%d  Bisect::DoEvent per s: reference %.0f, GpuPairAction %.0f;  DActionDBeta ms: reference %.2f, GpuPairAction %.2f"
          % (N, M, out["reference"], out["adapter"], out["reference_dbeta_ms"], out["adapter_dbeta_ms"]), flush=True)
