#!/usr/bin/env python
"""Multi-GPU check of the slice-sharded path (run under torchrun, one rank per GPU):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/sharded_check.py
Builds a two-species plasma (config C5 shape at reduced size), shards its slices over the
ranks, exchanges halos over NCCL, all-reduces the shard partial sums and compares with the
unsharded single-GPU evaluation and the CPU oracle.  Prints one JSON line on rank 0."""
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from simpimc_b200 import host, sharded, system as S  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    Ne = int(os.environ.get("PLASMA_N", "128"))
    M = int(os.environ.get("PLASMA_M", "64"))
    C = int(os.environ.get("PLASMA_CLONES", "8"))
    cfg = S.plasma_config(Ne=Ne, Np=Ne, M=M, n_xy=100, n_r_long=1000)
    Rs = [np.stack([S.synthetic_paths(cfg, sp, c, 777) for c in range(C)]) for sp in range(2)]
    sp_path = sharded.ShardedPath(cfg, C, local, rank, world)
    for sp in range(2):
        # upload WITHOUT the right halo (zeros), then let the NCCL ring fill it
        Rsh = sp_path.sh.shard_positions(Rs[sp])
        if world > 1:
            Rsh = Rsh.copy()
            Rsh[:, :, -1, :] = 0.0
        sp_path.path.SetPositions(sp, Rsh)
        sp_path.ExchangeHalo(sp)
        assert np.array_equal(sp_path.path.GetPositions(sp), sp_path.sh.shard_positions(Rs[sp])), "halo exchange"
    dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    du = [sp_path.DActionDBeta(a) for a in range(3)]
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    v = [sp_path.Potential(a) for a in range(3)]
    counts = sp_path.PairCorrelationCounts(0, 1, 0.0, cfg.L / 2, 100)
    ok = True
    err = 0.0
    if rank == 0:
        whole = host.Path(cfg, n_clones=C, device=local)
        for sp in range(2):
            whole.SetPositions(sp, Rs[sp])
        for a in range(3):
            ref_du, ref_v = whole.actions[a].DActionDBeta(), whole.actions[a].Potential()
            err = max(err, float(np.max(np.abs(du[a] - ref_du) / np.abs(ref_du))), float(np.max(np.abs(v[a] - ref_v) / np.abs(ref_v))))
        ref_counts = host.PairCorrelation(whole, 0, 1, 0.0, cfg.L / 2, 100).Counts().astype(np.int64)
        ok = err <= 1e-10 and np.array_equal(counts, ref_counts)
        whole.close()
        n_pairs = Ne * (Ne - 1) + Ne * Ne
        print(json.dumps({"check": "slice-sharded plasma", "n_gpus": world, "N": 2 * Ne, "M": M, "clones": C, "max_rel_err_vs_unsharded": err,
                          "gofr_bins_equal": bool(np.array_equal(counts, ref_counts)), "ok": bool(ok),
                          "dbeta_3_actions_ms": 1e3 * (t1 - t0), "pair_slice_evals": n_pairs * M * C}), flush=True)
    sp_path.close()
    dist.barrier()
    dist.destroy_process_group()
    if not ok:
        sys.exit(1)


if __name__ == "__main__":
    main()
