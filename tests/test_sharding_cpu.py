"""CPU tests of the multi-GPU host logic (world_size 2 and 3, gloo): slice-shard arithmetic, the
ring halo exchange, the ring rotation of the slices and the all-reduce of shard partial sums.  The compute stand-in is the CPU
oracle evaluating the action over each rank's slice window -- PairAction::GetAction(b0, b1, all
particles, 0) sums links b0..b1-1 (pair_action_class.h:282-290), which is exactly a shard's
partial sum -- so the reduced value must equal the whole-path action."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from simpimc_b200 import sharded, system as S  # noqa: E402


def test_slice_sharding_arithmetic():
    for M, G in ((128, 8), (512, 8), (10, 3), (7, 7), (16, 1)):
        seen = []
        for g in range(G):
            sh = sharded.SliceSharding(M, G, g)
            seen += list(range(sh.lo, sh.hi))
            st = sh.stored_slices()
            assert st[:sh.n_local] == list(range(sh.lo, sh.hi))
            if G > 1:
                assert st[-1] == sh.hi % M and sh.owner(st[-1]) == sh.next_rank
                assert sharded.SliceSharding(M, G, sh.prev_rank).hi % M == sh.lo
            else:
                assert len(st) == M
        assert seen == list(range(M))       # contiguous, disjoint, complete
    with pytest.raises(ValueError):
        sharded.SliceSharding(4, 8, 0)


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import oracle as O
        cfg = S.plasma_config(Ne=5, Np=4, M=12)
        sh = sharded.SliceSharding(cfg.n_bead, world, rank)
        o = O.Oracle(cfg)
        Rs = [S.synthetic_paths(cfg, sp, 0, 31) for sp in range(2)]
        for sp in range(2):
            o.set_positions(sp, Rs[sp])
        parts = [(s, p) for s in range(2) for p in range(cfg.species[s].n_part)]
        # shard partial sums of the three actions, all-reduced
        t = torch.tensor([o.get_action(a, 0, sh.lo, sh.hi, parts, 0) for a in range(3)], dtype=torch.float64)
        sharded.allreduce_sum(t)
        whole = np.array([o.get_action(a, 0, 0, cfg.n_bead, parts, 0) for a in range(3)])
        ok_sum = bool(np.all(np.abs(t.numpy() - whole) <= 1e-12 * np.abs(whole)))
        # ring halo: every rank sends its first slice back, receives the slice after its block
        ok_halo = True
        for sp in range(2):
            mine = sh.shard_positions(Rs[sp][None])          # [1][N][n_local + 1][3]
            send = torch.from_numpy(np.ascontiguousarray(mine[:, :, 0, :]))
            recv = torch.zeros_like(send)
            sharded.ring_halo(send, recv, sh)
            ok_halo = ok_halo and np.array_equal(recv.numpy(), Rs[sp][None][:, :, sh.hi % cfg.n_bead, :])
            ok_halo = ok_halo and np.array_equal(recv.numpy(), mine[:, :, -1, :])
        # ring rotation (ShardedPath.Rotate): every rank hands its first `shift` owned slices to the
        # previous rank and appends what the next rank sent; reassembled, the path is the original
        # rolled by `shift` slices, and its action -- the oracle on the rolled path -- is unchanged
        shift = 3
        ok_rot = True
        rolled = []
        for sp in range(2):
            own = np.ascontiguousarray(Rs[sp][None][:, :, sh.lo:sh.hi, :])           # [1][N][n_local][3]
            send = torch.from_numpy(np.ascontiguousarray(np.moveaxis(own[:, :, :shift, :], 2, 3)))   # [1][N][3][shift] as pimc_rotate_pack lays it out
            recv = torch.zeros_like(send)
            sharded.ring_halo(send, recv, sh)
            new_own = np.concatenate([own[:, :, shift:, :], np.moveaxis(recv.numpy(), 3, 2)], axis=2)
            parts_t = [torch.zeros_like(torch.from_numpy(new_own)) for _ in range(world)]   # n_bead = 12 divides evenly over 2 and 3 ranks
            dist.all_gather(parts_t, torch.from_numpy(np.ascontiguousarray(new_own)))
            full = torch.cat(parts_t, dim=2).numpy()[0]
            ok_rot = ok_rot and np.array_equal(full, np.roll(Rs[sp], -shift, axis=1))
            rolled.append(full)
        o2 = O.Oracle(cfg)
        for sp in range(2):
            o2.set_positions(sp, rolled[sp])
        again = np.array([o2.get_action(a, 0, 0, cfg.n_bead, parts, 0) for a in range(3)])
        ok_rot = ok_rot and bool(np.all(np.abs(again - whole) <= 1e-11 * np.abs(whole)))
        o2.close()
        o.close()
        ret[rank] = (ok_sum, ok_halo and ok_rot)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_shard_partials_and_halo_ring_over_gloo(world, oracle_mod):
    port = 29500 + (os.getpid() % 2000) + world
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    procs = [ctx.Process(target=_worker, args=(r, world, port, ret)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert len(ret) == world and all(v == (True, True) for v in ret.values()), dict(ret)
