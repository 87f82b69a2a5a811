"""Imaginary-time (slice) sharding of one large path across the GPUs of a node.

The reference never splits a path (its MPI ranks are independent walkers,
src/framework/framework_class.h:44-53); this is the build's own multi-GPU strategy for systems
like BASELINE config C5: rank g owns the contiguous slices
[g M / G, (g+1) M / G) of every particle of every species plus ONE halo slice of positions (the
first slice of the next shard, beta-periodic ring, same particle labels -- App. A-4), because the
pair action at level 0 couples slice b only with b + 1 (pair_action_class.h:282-288) while
rho_k(b), the k-sums and the estimators are slice-local.  Collectives -- on GPUs the library's own
NCCL calls behind the C ABI (pimc_halo_exchange, pimc_allreduce_sum, pimc_rotate,
pimc_sharded_evaluate; csrc/comm.cc), issued on the context's stream; torch.distributed is used only
to hand rank 0's NCCL unique id to the other ranks (and by the gloo CPU tests of the rank
arithmetic, through `ring_halo` / `allreduce_sum` below):

* halo      one slice of positions per species to the PREVIOUS rank after positions change
            (ring send/recv, N * n_d * 8 bytes per clone and species);
* reduce    all-reduce(SUM) of one double per clone and action (energies, actions) and of the
            g(r) / S(k) accumulators.

`SliceSharding` is the rank arithmetic, `ring_halo` / `allreduce_sum` are the two collectives on
plain tensors (CPU tests of the host logic), `ShardedPath` is a slice-sharded CUDA context plus
its `pimc_comm`.
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


def bootstrap_unique_id(lib, rank, world, group=None):
    """The 128-byte NCCL unique id of rank 0 on every rank.  The only use of torch.distributed on
    the GPU path: a C++ host would MPI_Bcast the same bytes."""
    import ctypes as C
    import torch.distributed as dist
    from . import capi
    buf = (C.c_ubyte * 128)()
    if world == 1:
        return buf
    if rank == 0:
        capi.check(lib.pimc_comm_unique_id(buf))
    box = [bytes(buf)]
    dist.broadcast_object_list(box, src=0, group=group)
    return (C.c_ubyte * 128).from_buffer_copy(box[0])


class ShardedPath:
    """A slice-sharded host.Path: every rank constructs it with the same cfg; whole-path
    quantities come back all-reduced (identical on every rank).  unique_id: the bytes of
    pimc_comm_unique_id from rank 0 (default: broadcast through torch.distributed)."""

    def __init__(self, cfg, n_clones, device, rank, world, group=None, unique_id=None):
        import ctypes as C
        import torch
        from . import host, capi
        self.torch = torch
        self.host = host
        self.cfg = cfg
        self.sh = SliceSharding(cfg.n_bead, world, rank)
        self.path = host.Path(cfg, n_clones=n_clones, device=device, slice_lo=self.sh.lo, slice_hi=self.sh.hi)
        self.lib = self.path.L
        self.n_clones = n_clones
        self.device = torch.device("cuda", device)
        self.stream = torch.cuda.ExternalStream(self.lib.pimc_ctx_stream(self.path.h), device=self.device)
        self.actions = self.path.actions
        self.pair_actions = [a for a in self.actions if a.type != "Kinetic"]
        if unique_id is None:
            unique_id = bootstrap_unique_id(self.lib, rank, world, group)
        elif isinstance(unique_id, (bytes, bytearray)):
            unique_id = (C.c_ubyte * 128).from_buffer_copy(bytes(unique_id))
        comm = C.c_void_p()
        capi.check(self.lib.pimc_comm_init(self.path.h, unique_id, rank, world, C.byref(comm)))
        self.comm = comm
        self._act_handles = (C.c_void_p * len(self.pair_actions))(*[a.h for a in self.pair_actions])
        self._graphs = []

    def close(self):
        for g in self._graphs:
            self.lib.pimc_graph_destroy(g)
        self._graphs = []
        if self.comm:
            self.lib.pimc_comm_destroy(self.comm)
            self.comm = None
        self.path.close()

    def SetPositions(self, species, R_full):
        """R_full[clone][particle][bead][dim] of the WHOLE path (every rank passes the same array)."""
        self.path.SetPositions(species, self.sh.shard_positions(R_full))

    def ExchangeHalo(self, species):
        """After this rank's positions changed: refresh the neighbours' halo slices (NCCL ring on
        the context's stream; asynchronous)."""
        from . import capi
        capi.check(self.lib.pimc_halo_exchange(self.path.h, self.comm, species))

    def Rotate(self, shift):
        """Rotate the ring of slices by `shift`: every rank hands its first `shift` owned slices to
        the previous rank (same ring as the halo), then halos and rho_k are refreshed.  Global slice
        labels rotate, shard ranges stay; actions and estimators are invariant (imaginary time is a
        ring), and former shard-boundary slices become interior, where BisectSweep can move them."""
        from . import capi
        capi.check(self.lib.pimc_rotate(self.path.h, self.comm, shift))

    def BisectSweep(self, species, n_level, n_attempts, seed, attempt0=0, with_kinetic=True):
        """n_attempts device-resident bisection attempts per clone on THIS rank's shard-interior
        windows (no communication: the shard's first slice and its halo stay fixed; call Rotate
        between batches).  The Philox key is the caller's seed mixed with the shard's first slice,
        so that the natural SPMD call -- the same seed on every rank -- still draws independent
        particle picks, window offsets and Levy displacements per shard."""
        return self.path.BisectSweep(species, n_level, n_attempts, shard_seed(seed, self.sh.lo), attempt0, with_kinetic)

    def BisectSweepWindows(self, species, n_level, n_rounds, seed, attempt0=0, with_kinetic=True):
        """Rounds of the move on every disjoint window of THIS rank's shard at once (no communication; Rotate
        between batches moves the windows' fixed end points).  Returns (accepts per clone, windows per round)."""
        return self.path.BisectSweepWindows(species, n_level, n_rounds, shard_seed(seed, self.sh.lo), attempt0, with_kinetic)

    def _reduced(self, which, ai):
        import ctypes as C
        from . import capi
        torch = self.torch
        out = torch.zeros(self.n_clones, dtype=torch.float64, device=self.device)
        h = (C.c_void_p * 1)(self.actions[ai].h)
        capi.check(self.lib.pimc_sharded_evaluate(self.path.h, self.comm, which, h, 1, out.data_ptr()))
        self.path.Sync()
        return out.cpu().numpy()

    def RebuildRhoK(self):
        """Species::InitRhoK of every species with rho_k (slice-local: no communication)."""
        from . import capi
        if self.path.n_k:
            for sp in range(len(self.cfg.species)):
                capi.check(self.lib.pimc_rhok_rebuild(self.path.h, sp))

    def DActionDBetaAllDevice(self, out):
        """Every pair action's DActionDBeta into the device tensor out[action][clone], shard partial
        sums combined by ONE all-reduce; asynchronous on the context's stream."""
        from . import capi
        capi.check(self.lib.pimc_sharded_evaluate(self.path.h, self.comm, 1, self._act_handles, len(self.pair_actions), out.data_ptr()))
        return out

    def CaptureStep(self, out):
        """The evaluation step -- rho_k rebuild of every species, DActionDBeta of every pair action,
        one all-reduce -- captured into a CUDA graph (run once un-captured first so every scratch
        buffer exists).  Returns a callable that replays it with a single launch."""
        import ctypes as C
        from . import capi
        self.RebuildRhoK()
        self.DActionDBetaAllDevice(out)
        self.path.Sync()
        capi.check(self.lib.pimc_capture_begin(self.path.h))
        try:
            self.RebuildRhoK()
            self.DActionDBetaAllDevice(out)
        finally:
            g = C.c_void_p()
            rc = self.lib.pimc_capture_end(self.path.h, C.byref(g))
        capi.check(rc)
        self._graphs.append(g)
        lib = self.lib

        def launch():
            capi.check(lib.pimc_graph_launch(g))
        launch.n_nodes = int(lib.pimc_graph_nodes(g))
        return launch

    def BytesSent(self):
        return int(self.lib.pimc_comm_bytes_sent(self.comm))

    def DActionDBeta(self, ai):
        return self._reduced(1, ai)

    def Potential(self, ai):
        return self._reduced(2, ai)

    def TotalAction(self, ai):
        return self._reduced(0, ai)

    def _allreduce_host(self, a):
        """SUM over the ranks of a host array of doubles (estimator rows) through the library's all-reduce."""
        from . import capi
        torch = self.torch
        t = torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).to(self.device)
        torch.cuda.current_stream(self.device).synchronize()
        capi.check(self.lib.pimc_allreduce_sum(self.path.h, self.comm, t.data_ptr(), t.numel()))
        self.path.Sync()
        return t.cpu().numpy()

    def PairCorrelationCounts(self, sa, sb, r_min, r_max, n_r):
        counts = self.host.PairCorrelation(self.path, sa, sb, r_min, r_max, n_r).Counts()
        # counts are far below 2^53: their sum in doubles is exact
        return np.rint(self._allreduce_host(counts.astype(np.float64))).astype(np.int64)

    def StructureFactor(self, sa, sb, k_cut):
        sk = self.host.StructureFactor(self.path, sa, sb, k_cut)
        sk.Accumulate()
        return self._allreduce_host(sk.sk)


def shard_seed(seed, slice_lo):
    """Philox key of a shard: the caller's seed mixed with the shard's first slice (golden-ratio
    multiplier, 64-bit wrap-around) -- identical seeds on every rank give independent streams."""
    return (int(seed) ^ (int(slice_lo) * 0x9E3779B97F4A7C15)) & 0xFFFFFFFFFFFFFFFF
