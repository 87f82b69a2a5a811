#!/usr/bin/env python
"""Multi-GPU check of the slice-sharded path (run under torchrun, one rank per GPU):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/sharded_check.py
Builds a two-species plasma (config C5 shape at reduced size), shards its slices over the
ranks, exchanges halos over NCCL, all-reduces the shard partial sums and compares with the
unsharded single-GPU evaluation.  Every collective of the data path is the library's own
(pimc_halo_exchange / pimc_allreduce_sum / pimc_rotate / pimc_sharded_evaluate behind the C ABI);
torch.distributed only hands rank 0's NCCL unique id to the other ranks and provides the barriers
around the timed regions.  Also replays the evaluation step from a captured CUDA graph and checks it
against the eager result.  Prints JSON lines on rank 0."""
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
    # the evaluation step from a captured CUDA graph (kernels + the NCCL all-reduce in one launch)
    out_g = torch.zeros((3, C), dtype=torch.float64, device=sp_path.device)
    replay = sp_path.CaptureStep(out_g)
    out_g.zero_()
    torch.cuda.synchronize()
    replay()
    sp_path.path.Sync()
    # equal to rounding: NCCL may order the captured all-reduce differently from the eager one (bit-identical at 2 ranks)
    graph_ok = bool(np.max(np.abs(out_g.cpu().numpy() - np.stack(du)) / np.abs(np.stack(du))) <= 1e-13)
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
        ok = err <= 1e-10 and np.array_equal(counts, ref_counts) and graph_ok
        whole.close()
        n_pairs = Ne * (Ne - 1) + Ne * Ne
        print(json.dumps({"check": "slice-sharded plasma", "n_gpus": world, "N": 2 * Ne, "M": M, "clones": C, "max_rel_err_vs_unsharded": err,
                          "gofr_bins_equal": bool(np.array_equal(counts, ref_counts)), "graph_replay_matches_eager": graph_ok,
                          "graph_nodes": replay.n_nodes, "comm_bytes_sent_rank0": sp_path.BytesSent(), "ok": bool(ok),
                          "dbeta_3_actions_ms": 1e3 * (t1 - t0), "pair_slice_evals": n_pairs * M * C}), flush=True)
    # ---- moves on the sharded path: shard-interior bisection windows, ring rotation over NCCL ----
    n_level, n_att = 2, 24
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    n_acc = 0
    for rnd in range(3):
        for sp in range(2):
            n_acc += int(sp_path.BisectSweep(sp, n_level, n_att, 1000 + rank, attempt0=rnd * n_att).sum())
        sp_path.Rotate(5)
    torch.cuda.synchronize()
    t3 = time.perf_counter()
    du_mc = [sp_path.DActionDBeta(a) for a in range(3)]
    # gather the shards and evaluate the whole (rotated) path unsharded on rank 0
    mc_err, moved = 0.0, False
    gathered = []
    for sp in range(2):
        # every rank writes its own slices into a zeroed whole-path array; the library's own all-reduce assembles it
        full = np.zeros_like(Rs[sp])
        full[:, :, sp_path.sh.lo:sp_path.sh.hi, :] = sp_path.path.GetPositions(sp)[:, :, :sp_path.sh.n_local, :]
        gathered.append(sp_path._allreduce_host(full))
    if rank == 0:
        whole = host.Path(cfg, n_clones=C, device=local)
        for sp in range(2):
            whole.SetPositions(sp, gathered[sp])
        for a in range(3):
            ref = whole.actions[a].DActionDBeta()
            mc_err = max(mc_err, float(np.max(np.abs(du_mc[a] - ref) / np.abs(ref))))
            moved = moved or bool(np.any(du_mc[a] != du[a]))
        whole.close()
        ok_mc = mc_err <= 1e-10 and moved and n_acc > 0
        print(json.dumps({"check": "slice-sharded moves", "n_gpus": world, "attempts_per_rank": 3 * 2 * n_att * C, "accepted_on_rank0": n_acc,
                          "rotations": 3, "max_rel_err_vs_unsharded_after_moves": mc_err, "ok": bool(ok_mc),
                          "ms_per_attempt_batch": 1e3 * (t3 - t2) / (3 * 2 * n_att)}), flush=True)
        ok = ok and ok_mc
    sp_path.close()
    dist.barrier()
    dist.destroy_process_group()
    if not ok:
        sys.exit(1)


if __name__ == "__main__":
    main()
