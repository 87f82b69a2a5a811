"""Imaginary-time (slice) sharding of one large path across the GPUs of a node.

The reference never splits a path (its MPI ranks are independent walkers,
src/framework/framework_class.h:44-53); this is the build's own multi-GPU strategy for systems
like BASELINE config C5: rank g owns the contiguous slices
[g M / G, (g+1) M / G) of every particle of every species plus ONE halo slice of positions (the
first slice of the next shard, beta-periodic ring, same particle labels -- App. A-4), because the
pair action at level 0 couples slice b only with b + 1 (pair_action_class.h:282-288) while
rho_k(b), the k-sums and the estimators are slice-local.  Collectives (torch.distributed; NCCL
over NVLink on GPUs, gloo in the CPU tests):

* halo      one slice of positions per species to the PREVIOUS rank after positions change
            (ring send/recv, N * n_d * 8 bytes per clone and species);
* reduce    all-reduce(SUM) of one double per clone and action (energies, actions) and of the
            g(r) / S(k) accumulators.

`SliceSharding` is the rank arithmetic, `ring_halo` / `allreduce_sum` are the two collectives on
plain tensors (CPU-testable), `ShardedPath` binds them to a slice-sharded CUDA context.
"""
import numpy as np


class SliceSharding:
    """Contiguous slice blocks of a path of n_bead slices over `world` ranks."""

    def __init__(self, n_bead, world, rank):
        if world < 1 or not (0 <= rank < world):
            raise ValueError("bad rank / world size")
        if world > n_bead:
            raise ValueError("more ranks than time slices")
        self.n_bead, self.world, self.rank = n_bead, world, rank
        self.lo = rank * n_bead // world
        self.hi = (rank + 1) * n_bead // world
        self.next_rank = (rank + 1) % world   # owns slice hi (mod n_bead): the sender of our halo
        self.prev_rank = (rank - 1) % world   # needs our first slice as its halo

    @property
    def n_local(self):
        return self.hi - self.lo

    @property
    def sharded(self):
        return self.world > 1

    def owner(self, b):
        """Rank that owns (global) slice b."""
        b %= self.n_bead
        for g in range(self.world):
            if g * self.n_bead // self.world <= b < (g + 1) * self.n_bead // self.world:
                return g
        raise AssertionError

    def stored_slices(self):
        """Global indices of the slices a rank stores: its block, then the halo."""
        idx = list(range(self.lo, self.hi))
        if self.sharded:
            idx.append(self.hi % self.n_bead)
        return idx

    def shard_positions(self, R):
        """R[..., bead, dim] of the whole path -> the stored slices of this rank."""
        return np.ascontiguousarray(np.take(R, self.stored_slices(), axis=-2))


def allreduce_sum(t, group=None):
    """In-place SUM over the ranks; returns t.  No-op without an initialised process group."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t


def ring_halo(send, recv, sharding, group=None):
    """Send `send` (this rank's first slice) to the previous rank and receive the next rank's
    first slice into `recv`.  Tensors of identical shape on every rank."""
    import torch.distributed as dist
    if sharding.world == 1:
        recv.copy_(send)
        return recv
    ops = [dist.P2POp(dist.isend, send, sharding.prev_rank, group), dist.P2POp(dist.irecv, recv, sharding.next_rank, group)]
    for req in dist.batch_isend_irecv(ops):
        req.wait()
    return recv


class ShardedPath:
    """A slice-sharded host.Path: every rank constructs it with the same cfg; whole-path
    quantities come back all-reduced (identical on every rank)."""

    def __init__(self, cfg, n_clones, device, rank, world, group=None):
        import torch
        from . import host
        self.torch = torch
        self.host = host
        self.cfg = cfg
        self.group = group
        self.sh = SliceSharding(cfg.n_bead, world, rank)
        self.path = host.Path(cfg, n_clones=n_clones, device=device, slice_lo=self.sh.lo, slice_hi=self.sh.hi)
        self.n_clones = n_clones
        self.device = torch.device("cuda", device)
        self.stream = torch.cuda.ExternalStream(self.path.L.pimc_ctx_stream(self.path.h), device=self.device)
        self.actions = self.path.actions

    def close(self):
        self.path.close()

    def SetPositions(self, species, R_full):
        """R_full[clone][particle][bead][dim] of the WHOLE path (every rank passes the same array)."""
        self.path.SetPositions(species, self.sh.shard_positions(R_full))

    def ExchangeHalo(self, species):
        """After this rank's positions changed: refresh the neighbours' halo slices (NCCL ring)."""
        from . import capi
        if not self.sh.sharded:
            return
        N = self.cfg.species[species].n_part
        torch = self.torch
        with torch.cuda.stream(self.stream):
            send = torch.empty((self.n_clones, N, 3), dtype=torch.float64, device=self.device)
            recv = torch.empty_like(send)
            capi.check(self.path.L.pimc_halo_pack(self.path.h, species, send.data_ptr()))
            ring_halo(send, recv, self.sh, self.group)
            capi.check(self.path.L.pimc_halo_unpack(self.path.h, species, recv.data_ptr()))
        self.stream.synchronize()

    def Rotate(self, shift):
        """Rotate the ring of slices by `shift`: every rank hands its first `shift` owned slices to
        the previous rank (same ring as the halo), then halos and rho_k are refreshed.  Global slice
        labels rotate, shard ranges stay; actions and estimators are invariant (imaginary time is a
        ring), and former shard-boundary slices become interior, where BisectSweep can move them."""
        from . import capi
        if not self.sh.sharded:
            return
        torch = self.torch
        for species in range(len(self.cfg.species)):
            N = self.cfg.species[species].n_part
            with torch.cuda.stream(self.stream):
                send = torch.empty((self.n_clones, N, 3, shift), dtype=torch.float64, device=self.device)
                recv = torch.empty_like(send)
                capi.check(self.path.L.pimc_rotate_pack(self.path.h, species, shift, send.data_ptr()))
                ring_halo(send, recv, self.sh, self.group)
                capi.check(self.path.L.pimc_rotate_apply(self.path.h, species, shift, recv.data_ptr()))
            self.stream.synchronize()
            self.ExchangeHalo(species)
        self.RebuildRhoK()

    def BisectSweep(self, species, n_level, n_attempts, seed, attempt0=0, with_kinetic=True):
        """n_attempts device-resident bisection attempts per clone on THIS rank's shard-interior
        windows (no communication: the shard's first slice and its halo stay fixed; call Rotate
        between batches).  Seeds should differ between ranks.  Returns the accept counts."""
        return self.path.BisectSweep(species, n_level, n_attempts, seed, attempt0, with_kinetic)

    def _reduced(self, fn, act):
        from . import capi
        torch = self.torch
        with torch.cuda.stream(self.stream):
            out = torch.zeros(self.n_clones, dtype=torch.float64, device=self.device)
            capi.check(fn(act.h, out.data_ptr()))
            allreduce_sum(out, self.group)
            res = out.cpu().numpy()
        return res

    def RebuildRhoK(self):
        """Species::InitRhoK of every species with rho_k (slice-local: no communication)."""
        from . import capi
        if self.path.n_k:
            for sp in range(len(self.cfg.species)):
                capi.check(self.path.L.pimc_rhok_rebuild(self.path.h, sp))

    def DActionDBetaAllDevice(self, out):
        """Every pair action's DActionDBeta into the device tensor out[action][clone], shard partial
        sums combined by ONE all-reduce; asynchronous on the context's stream."""
        from . import capi
        torch = self.torch
        with torch.cuda.stream(self.stream):
            row = 0
            for act in self.actions:
                if act is None:
                    continue
                capi.check(self.path.L.pimc_action_dbeta_device(act.h, out[row].data_ptr()))
                row += 1
            allreduce_sum(out, self.group)
        return out

    def DActionDBeta(self, ai):
        return self._reduced(self.path.L.pimc_action_dbeta_device, self.actions[ai])

    def Potential(self, ai):
        return self._reduced(self.path.L.pimc_action_potential_device, self.actions[ai])

    def TotalAction(self, ai):
        return self._reduced(self.path.L.pimc_action_total_device, self.actions[ai])

    def PairCorrelationCounts(self, sa, sb, r_min, r_max, n_r):
        torch = self.torch
        counts = self.host.PairCorrelation(self.path, sa, sb, r_min, r_max, n_r).Counts().astype(np.int64)
        t = torch.from_numpy(counts).to(self.device)
        allreduce_sum(t, self.group)
        return t.cpu().numpy()

    def StructureFactor(self, sa, sb, k_cut):
        torch = self.torch
        sk = self.host.StructureFactor(self.path, sa, sb, k_cut)
        sk.Accumulate()
        t = torch.from_numpy(sk.sk).to(self.device)
        allreduce_sum(t, self.group)
        return t.cpu().numpy()
