"""Device-resident bisection moves (pimc_bisect_sweep, csrc/mc.cuh).

(1) Exact stream parity: the host mirror simpimc_b200.moves.bisect_attempt_philox draws the
    same Philox numbers and takes its pair-action values from the CPU oracle; after every
    attempt the device walkers must sit at the same positions (1e-12) with the same
    accept/reject history.
(2) Statistical parity with the REFERENCE program (oracle/_ref: the reference's own Bisect,
    Kinetic and IlkkaPairAction, std::mt19937): thermal energy within combined error bars.
"""
import numpy as np
import pytest

from simpimc_b200 import system as S

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name,general", [("ueg", False), ("ueg", True), ("plasma", False), ("nolr", False), ("nolr", True), ("ueg4", False),
                                          ("carbon", False), ("david", False), ("david", True), ("plasma_david", False)])
def test_device_sweep_follows_the_host_mirror_of_its_stream(name, general):
    """general = False: the single-launch sweep (csrc/sweep_fused.cuh) where it applies (ueg, nolr,
    ueg4: one same-species Ilkka action), the kernel-per-phase path otherwise (plasma);
    general = True forces the kernel-per-phase path."""
    from simpimc_b200 import host, moves
    from oracle import oracle as O
    if name == "ueg":
        cfg, n_level = S.ueg_config(N=9, M=16), 3
    elif name == "ueg4":
        cfg, n_level = S.ueg_config(N=37, M=32), 4
    elif name == "nolr":
        cfg, n_level = S.ueg_config(N=6, M=8, use_long_range=False), 2
    elif name == "david":      # general = False: the pp-form tables of the fast David kernel, read from global memory
        cfg, n_level = S.ueg_config(N=7, M=16, action="DavidPairAction", use_long_range=True), 3
    elif name == "plasma_david":
        cfg, n_level = S.plasma_config(Ne=5, Np=4, M=8, pp_action="DavidPairAction", ep_action="DavidPairAction"), 2
    elif name == "carbon":
        cfg, n_level = S.carbon_config(), 2     # BASELINE config C4: 4 species, 9 Ilkka + 1 Bare action, all long-range
    else:
        cfg, n_level = S.plasma_config(Ne=5, Np=4, M=8), 2
    C = 3
    path = host.Path(cfg, n_clones=C)
    path.ForceGeneral(general)
    oracles = []
    for c in range(C):
        o = O.Oracle(cfg)
        oracles.append(o)
    for sp in range(len(cfg.species)):
        R = np.stack([S.synthetic_paths(cfg, sp, c, 99) for c in range(C)])
        path.SetPositions(sp, R)
        for c in range(C):
            oracles[c].set_positions(sp, R[c])
    seed = 0x1234567800000042
    M = cfg.n_bead
    n_acc_dev = np.zeros(C, dtype=np.int64)
    n_acc_host = np.zeros(C, dtype=np.int64)
    for attempt in range(40):
        sp = attempt % len(cfg.species)
        acts = [ai for ai, a in enumerate(cfg.actions) if cfg.species[sp].name in (a.species_a, a.species_b)]

        def get_beads(c, p, b0, n):
            return oracles[c].get_positions(sp, 0)[p, (b0 + np.arange(n)) % M]

        def action_old_new(c, p, b0, nb, new):
            oracles[c].propose(sp, p, (b0 + 1) % M, new)
            old = sum(oracles[c].get_action(ai, 0, b0, b0 + nb, [(sp, p)], 0) for ai in acts)
            nw = sum(oracles[c].get_action(ai, 1, b0, b0 + nb, [(sp, p)], 0) for ai in acts)
            return old, nw

        def finish(c, p, b0, nb, accept, new):
            if new is not None:
                oracles[c].finish_move(sp, p, b0, b0 + nb, bool(accept))

        _, _, acc = moves.bisect_attempt_philox(cfg, sp, n_level, seed, attempt, C, get_beads, action_old_new, finish)
        n_acc_host += acc
        n_acc_dev += path.BisectSweep(sp, n_level, 1, seed, attempt0=attempt)
        got = path.GetPositions(sp)
        for c in range(C):
            ref = oracles[c].get_positions(sp, 0)
            assert np.max(np.abs(got[c] - ref)) <= 1e-12 * max(1.0, np.max(np.abs(ref))), (name, attempt, c)
        assert np.array_equal(n_acc_dev, n_acc_host), (name, attempt, n_acc_dev, n_acc_host)
    assert n_acc_dev.sum() > 0
    # rho_k was carried along incrementally on the device: compare with a rebuild by the oracle
    if path._n_k():
        for sp in range(len(cfg.species)):
            for c in range(C):
                assert np.max(np.abs(path.GetRhoK(sp, c, host.OLD_MODE) - oracles[c].rhok(sp, 0))) <= 1e-10 * cfg.species[sp].n_part
    for ai, act in enumerate(path.actions):
        du = act.DActionDBeta()
        for c in range(C):
            ref = oracles[c].dbeta(ai)
            assert abs(du[c] - ref) <= 1e-10 * abs(ref)
    path.close()


@pytest.mark.parametrize("C,N,M,n_level,lr", [(310, 9, 16, 3, True), (1337, 5, 8, 2, True), (150, 7, 8, 1, False)])
def test_single_launch_sweep_equals_the_kernel_per_phase_path(C, N, M, n_level, lr):
    """Many clones per CTA (C > 148), several batches of 8 clones per CTA (C > 8 x 148) and many
    attempts inside ONE launch: same accept history and positions as the general path, which
    the test above pins to the oracle attempt by attempt."""
    from simpimc_b200 import host
    cfg = S.ueg_config(N=N, M=M, use_long_range=lr)
    R = np.stack([S.synthetic_paths(cfg, 0, c, 3) for c in range(C)])
    out = []
    for general in (False, True):
        path = host.Path(cfg, n_clones=C)
        path.ForceGeneral(general)
        path.SetPositions(0, R)
        launches0 = path.LaunchCount()
        acc = path.BisectSweep(0, n_level, 24, 77, attempt0=5)
        acc = acc + path.BisectSweep(0, n_level, 9, 77, attempt0=29)
        n_launch = path.LaunchCount() - launches0
        out.append((acc, path.GetPositions(0), path.GetRhoK(0, C - 1, host.OLD_MODE) if lr else None, path.actions[0].DActionDBeta(), n_launch))
        path.close()
    (a0, r0, k0, d0, l0), (a1, r1, k1, d1, l1) = out
    assert l0 == 2 and l1 > 2 * 24          # one launch per call vs several per attempt
    assert np.array_equal(a0, a1) and a0.sum() > 0
    assert np.max(np.abs(r0 - r1)) <= 1e-12 * cfg.L
    if lr:
        assert np.max(np.abs(k0 - k1)) <= 1e-11 * N
    assert np.max(np.abs(d0 - d1) / np.abs(d1)) <= 1e-10


def _mean_err(x):
    """scripts/Stats.cpp:42-87: mean, autocorrelation time kappa = 1 + 2 sum_{t: C(t) > 0} C(t), error."""
    x = np.asarray(x, dtype=np.float64)
    n = len(x)
    m, var = x.mean(), x.var()
    if var == 0:
        return m, 0.0
    kappa = 1.0
    for t in range(1, n // 2):
        ct = np.mean((x[:n - t] - m) * (x[t:] - m)) / var
        if ct <= 0:
            break
        kappa += 2.0 * ct
    return m, np.sqrt(var * kappa / n)


def test_sampled_gofr_and_sofk_match_the_reference_program():
    """north_star: sampled runs reproduce the reference's g(r) and S(k) within statistical error
    bars.  Reference: its own Bisect + PairCorrelation + StructureFactor on one walker
    (std::mt19937), blocked series.  Device: 256 independent walkers (Philox), the estimators of
    the CUDA path; error bars from the walker-to-walker spread."""
    from oracle import refsim
    if not refsim.available():
        pytest.skip("oracle/_ref not built")
    from simpimc_b200 import host
    N, M, n_level, n_r = 4, 8, 2, 8
    cfg = S.ueg_config(N=N, M=M, with_kinetic=True, n_xy=60, n_r_long=400)
    cfg.moves = [{"name": "BisectE", "type": "Bisect", "species": "e", "n_level": n_level}]
    cfg.observables = [{"name": "gr", "type": "PairCorrelation", "species_a": "e", "species_b": "e", "r_min": 0.0, "r_max": cfg.L / 2.0, "n_r": n_r},
                       {"name": "sk", "type": "StructureFactor", "species_a": "e", "species_b": "e", "k_cut": cfg.k_cut}]
    attempts_per_sweep = N * M // (1 << n_level)
    sim = refsim.RefSim(cfg, seed=23)
    sim.set_positions(0, S.synthetic_paths(cfg, 0, 0, 5))
    sim.move_do(0, 400 * attempts_per_sweep)
    n_blocks, per_block = 40, 60
    g_blocks, s_blocks = [], []
    g_prev, s_prev = sim.gofr_counts(0, n_r), sim.sofk_sums(1)
    for _ in range(n_blocks):
        for _ in range(per_block):
            sim.move_do(0, attempts_per_sweep)
            sim.observable_accumulate(0)
            sim.observable_accumulate(1)
        g_now, s_now = sim.gofr_counts(0, n_r), sim.sofk_sums(1)
        g_blocks.append((g_now - g_prev) / per_block)
        s_blocks.append((s_now - s_prev) / per_block)
        g_prev, s_prev = g_now, s_now
    sim.close()
    g_ref, s_ref = np.array(g_blocks), np.array(s_blocks)
    # device
    C = 256
    gcfg = S.ueg_config(N=N, M=M, n_xy=60, n_r_long=400)
    path = host.Path(gcfg, n_clones=C)
    path.SetPositions(0, np.stack([S.synthetic_paths(gcfg, 0, c, 5) for c in range(C)]))
    gr = host.PairCorrelation(path, 0, 0, 0.0, gcfg.L / 2.0, n_r)
    sk = host.StructureFactor(path, 0, 0, gcfg.k_cut)
    att = 200 * attempts_per_sweep
    path.BisectSweep(0, n_level, att, 31, attempt0=0)
    n_meas = 40
    for _ in range(n_meas):
        path.BisectSweep(0, n_level, 3 * attempts_per_sweep, 31, attempt0=att)
        att += 3 * attempts_per_sweep
        gr.Accumulate()
        sk.Accumulate()
    g_dev, s_dev = gr.y / n_meas, sk.sk / n_meas      # per walker: counts per measurement, sum_b |rho_k|^2
    path.close()
    n_sig = 4.5
    for name, dev, ref in (("g(r)", g_dev, g_ref), ("S(k)", s_dev, s_ref)):
        d_mean, d_err = dev.mean(axis=0), dev.std(axis=0, ddof=1) / np.sqrt(C)
        r_mean, r_err = ref.mean(axis=0), ref.std(axis=0, ddof=1) / np.sqrt(len(ref))
        scale = np.max(np.abs(r_mean))
        sig = np.hypot(d_err, r_err) + 1e-12 * scale
        assert np.all(np.abs(d_mean - r_mean) <= n_sig * sig), (name, d_mean, r_mean, sig)
        # the comparison has teeth: the error bars are small against the signal where it is large
        big = np.abs(r_mean) > 0.2 * scale
        assert np.all(sig[big] < 0.1 * np.abs(r_mean[big])), (name, sig, r_mean)
    # same normalisation on both sides: pairs counted per measurement
    assert abs(g_dev.sum(axis=1).mean() - g_ref.sum(axis=1).mean()) < 0.1 * N * (N - 1) / 2 * M


def test_sampled_energy_matches_the_reference_program():
    from oracle import refsim
    if not refsim.available():
        pytest.skip("oracle/_ref not built")
    from simpimc_b200 import host
    cfg = S.ueg_config(N=4, M=8, with_kinetic=True, n_xy=60, n_r_long=400)
    n_level = 2
    cfg.moves = [{"name": "BisectE", "type": "Bisect", "species": "e", "n_level": n_level}]
    cfg.observables = []
    attempts_per_sweep = 4 * 8 // (1 << n_level)
    # reference: one walker, its own Bisect + Kinetic + IlkkaPairAction, std::mt19937
    sim = refsim.RefSim(cfg, seed=11)
    R0 = S.synthetic_paths(cfg, 0, 0, 5)
    sim.set_positions(0, R0)
    sim.move_do(0, 400 * attempts_per_sweep)
    ref_series = []
    for _ in range(3000):
        sim.move_do(0, attempts_per_sweep)
        ref_series.append(sim.dbeta(1) / cfg.n_bead)   # pair-action part of the thermal energy
    sim.close()
    # device: 256 walkers, Philox
    C = 256
    gcfg = S.ueg_config(N=4, M=8, n_xy=60, n_r_long=400)
    path = host.Path(gcfg, n_clones=C)
    path.SetPositions(0, np.stack([S.synthetic_paths(gcfg, 0, c, 5) for c in range(C)]))
    act = path.actions[0]
    att = 0
    path.BisectSweep(0, n_level, 200 * attempts_per_sweep, 7, attempt0=att)
    att += 200 * attempts_per_sweep
    blocks = []
    for _ in range(40):
        path.BisectSweep(0, n_level, 5 * attempts_per_sweep, 7, attempt0=att)
        att += 5 * attempts_per_sweep
        blocks.append(act.DActionDBeta() / gcfg.n_bead)
    path.close()
    per_clone = np.mean(np.array(blocks), axis=0)          # independent walkers
    g_mean, g_err = per_clone.mean(), per_clone.std(ddof=1) / np.sqrt(C)
    r_mean, r_err = _mean_err(ref_series)
    assert abs(g_mean - r_mean) <= 4.0 * np.hypot(g_err, r_err), (g_mean, g_err, r_mean, r_err)
    assert g_err < 0.05 * abs(r_mean) and r_err < 0.05 * abs(r_mean)


@pytest.mark.parametrize("name,n_shards", [("plasma", 2), ("ueg", 4)])
def test_slice_sharded_sweeps_follow_the_host_mirror_and_rotate(name, n_shards):
    """Moves on a slice-sharded path: every shard bisects windows that lie inside its
    stored slices (no communication), then the ring of slices is rotated (pimc_rotate_pack /
    _apply + halo refresh + rho_k rebuild) so that the former boundary slices become interior.
    Shards live on one GPU here and the ring is moved by hand; in production it is NCCL
    (sharded.ShardedPath.Rotate).  The host mirror draws each shard's Philox stream with the same
    window range and takes its action values from the CPU oracle of the WHOLE path: same positions
    after every attempt, same accept history; the rotated path has the same energy."""
    import ctypes as C
    import torch
    from simpimc_b200 import host, moves, sharded, capi
    from oracle import oracle as O
    if name == "plasma":
        cfg, n_level = S.plasma_config(Ne=5, Np=4, M=16), 2
    else:
        cfg, n_level = S.ueg_config(N=7, M=32), 3      # single same-species Ilkka action: the one-launch sweep kernel
    M, nb, n_clones, ns = cfg.n_bead, 1 << n_level, 2, len(cfg.species)
    Rs = [np.stack([S.synthetic_paths(cfg, sp, c, 4242) for c in range(n_clones)]) for sp in range(ns)]
    oracles = [O.Oracle(cfg) for _ in range(n_clones)]
    for c, o in enumerate(oracles):
        for sp in range(ns):
            o.set_positions(sp, Rs[sp][c])
    e0 = [[o.dbeta(ai) for ai in range(len(cfg.actions))] for o in oracles]
    shards = []
    for g in range(n_shards):
        sh = sharded.SliceSharding(M, n_shards, g)
        p = host.Path(cfg, n_clones=n_clones, slice_lo=sh.lo, slice_hi=sh.hi)
        for sp in range(ns):
            p.SetPositions(sp, sh.shard_positions(Rs[sp]))
        shards.append((sh, p))
    attempt, n_acc_dev, n_acc_host = 0, 0, 0
    for rnd in range(3):
        for g, (sh, p) in enumerate(shards):
            seed = 0xABC0000 + 17 * g
            for _ in range(6):
                sp = attempt % ns
                acts = [ai for ai, a in enumerate(cfg.actions) if cfg.species[sp].name in (a.species_a, a.species_b)]

                def get_beads(c, q, b0, n):
                    return oracles[c].get_positions(sp, 0)[q, (b0 + np.arange(n)) % M]

                def action_old_new(c, q, b0, nb_, new):
                    oracles[c].propose(sp, q, (b0 + 1) % M, new)
                    old = sum(oracles[c].get_action(ai, 0, b0, b0 + nb_, [(sp, q)], 0) for ai in acts)
                    nw = sum(oracles[c].get_action(ai, 1, b0, b0 + nb_, [(sp, q)], 0) for ai in acts)
                    return old, nw

                def finish(c, q, b0, nb_, accept, new):
                    if new is not None:
                        oracles[c].finish_move(sp, q, b0, b0 + nb_, bool(accept))

                _, b0s, acc = moves.bisect_attempt_philox(cfg, sp, n_level, seed, attempt, n_clones, get_beads, action_old_new, finish,
                                                          b0_range=(sh.lo, sh.n_local - nb + 1))
                assert all(sh.lo <= b and b + nb <= sh.hi for b in b0s)
                n_acc_host += int(np.sum(acc))
                n_acc_dev += int(p.BisectSweep(sp, n_level, 1, seed, attempt0=attempt).sum())
                attempt += 1
                for s2 in range(ns):
                    got = p.GetPositions(s2)
                    for c in range(n_clones):
                        ref = oracles[c].get_positions(s2, 0)[:, np.array(sh.stored_slices()), :]
                        assert np.max(np.abs(got[c] - ref)) <= 1e-12 * cfg.L, (name, rnd, g, attempt, s2, c)
                assert n_acc_dev == n_acc_host
        # rotate the ring by `shift` slices: pack everywhere first, then apply, then halos, then rho_k
        shift = 3
        for sp in range(ns):
            N = cfg.species[sp].n_part
            bufs = []
            for sh, p in shards:
                t = torch.zeros((n_clones, N, 3, shift), dtype=torch.float64, device="cuda")
                capi.check(p.L.pimc_rotate_pack(p.h, sp, shift, C.c_void_p(t.data_ptr())))
                p.Sync()
                bufs.append(t)
            for sh, p in shards:
                capi.check(p.L.pimc_rotate_apply(p.h, sp, shift, C.c_void_p(bufs[sh.next_rank].data_ptr())))
                p.Sync()
            halos = []
            for sh, p in shards:
                t = torch.zeros((n_clones, N, 3), dtype=torch.float64, device="cuda")
                capi.check(p.L.pimc_halo_pack(p.h, sp, C.c_void_p(t.data_ptr())))
                p.Sync()
                halos.append(t)
            for sh, p in shards:
                capi.check(p.L.pimc_halo_unpack(p.h, sp, C.c_void_p(halos[sh.next_rank].data_ptr())))
                capi.check(p.L.pimc_rhok_rebuild(p.h, sp))
                p.Sync()
        # the same relabelling on the oracle: slice b now holds what slice b + shift held
        for c, o in enumerate(oracles):
            for sp in range(ns):
                o.set_positions(sp, np.roll(o.get_positions(sp, 0), -shift, axis=1))
        for sp in range(ns):
            for sh, p in shards:
                got = p.GetPositions(sp)
                for c in range(n_clones):
                    ref = oracles[c].get_positions(sp, 0)[:, np.array(sh.stored_slices()), :]
                    assert np.max(np.abs(got[c] - ref)) <= 1e-12 * cfg.L, (name, "rotation", sp, c)
    assert n_acc_dev > 0
    for ai in range(len(cfg.actions)):
        du = sum(p.actions[ai].DActionDBeta() for _, p in shards)
        for c, o in enumerate(oracles):
            ref = o.dbeta(ai)
            assert abs(du[c] - ref) <= 1e-10 * abs(ref)
            assert ref != e0[c][ai]           # the walkers did move
    # rho_k carried through sweeps and rotations equals a rebuild by the oracle
    for sh, p in shards:
        for sp in range(ns):
            for c in range(n_clones):
                got = p.GetRhoK(sp, c, host.OLD_MODE)
                assert np.max(np.abs(got - oracles[c].rhok(sp, 0)[sh.lo:sh.hi])) <= 1e-10 * cfg.species[sp].n_part
    for _, p in shards:
        p.close()
    for o in oracles:
        o.close()


@pytest.mark.parametrize("name,general", [("ueg", False), ("ueg", True), ("plasma", False), ("bare", False), ("nolr", False),
                                          ("david", False), ("david", True), ("plasma_david", False), ("ueg_wide", False)])
def test_device_displace_follows_the_host_mirror_of_its_stream(name, general):
    """pimc_displace_sweep (csrc/displace.cuh): DisplaceParticle::DoEvent on the device.  The host
    mirror draws the same Philox numbers and takes the whole-path OLD / NEW actions from the CPU
    oracle: same positions (1e-12) and accept history after every attempt, rho_k carried along."""
    from simpimc_b200 import host, moves
    from oracle import oracle as O
    if name == "ueg":
        cfg, step = S.ueg_config(N=9, M=40), 0.9      # M = 40: a partial second chunk of 32 links
    elif name == "ueg_wide":   # > 16 partner steps per warp: the parked long-range r' of lane 31 are flushed inside the partner loop
        cfg, step = S.ueg_config(N=530, M=40, n_xy=60, n_r_long=400), 0.9
    elif name == "bare":
        cfg, step = S.ueg_config(N=6, M=8, action="BarePairAction"), 1.2
    elif name == "nolr":
        cfg, step = S.ueg_config(N=6, M=8, use_long_range=False), 1.0
    elif name == "david":      # general = False: the pp-form tables of the fast David kernel, read from global memory
        cfg, step = S.ueg_config(N=6, M=8, action="DavidPairAction", use_long_range=True), 1.0
    elif name == "plasma_david":
        cfg, step = S.plasma_config(Ne=5, Np=4, M=8, pp_action="DavidPairAction", ep_action="DavidPairAction"), 0.8
    else:
        cfg, step = S.plasma_config(Ne=5, Np=4, M=8), 0.8
    C = 3
    path = host.Path(cfg, n_clones=C)
    path.ForceGeneral(general)
    oracles = [O.Oracle(cfg) for _ in range(C)]
    for sp in range(len(cfg.species)):
        R = np.stack([S.synthetic_paths(cfg, sp, c, 99) for c in range(C)])
        path.SetPositions(sp, R)
        for c in range(C):
            oracles[c].set_positions(sp, R[c])
    seed, M = 0x5EED00000007, cfg.n_bead
    n_dev = np.zeros(C, dtype=np.int64)
    n_host = np.zeros(C, dtype=np.int64)
    for attempt in range(24):
        sp = attempt % len(cfg.species)
        acts = [ai for ai, a in enumerate(cfg.actions) if cfg.species[sp].name in (a.species_a, a.species_b)]

        def get_beads(c, p, b0, n):
            return oracles[c].get_positions(sp, 0)[p, (b0 + np.arange(n)) % M]

        def action_old_new(c, p, new):
            oracles[c].propose(sp, p, 0, new)
            old = sum(oracles[c].get_action(ai, 0, 0, M, [(sp, p)], 0) for ai in acts)
            nw = sum(oracles[c].get_action(ai, 1, 0, M, [(sp, p)], 0) for ai in acts)
            return old, nw

        def finish(c, p, accept):
            oracles[c].finish_move(sp, p, 0, M, bool(accept))

        _, acc = moves.displace_attempt_philox(cfg, sp, step, seed, attempt, C, get_beads, action_old_new, finish)
        n_host += acc
        n_dev += path.DisplaceSweep(sp, step, 1, seed, attempt0=attempt)
        got = path.GetPositions(sp)
        for c in range(C):
            ref = oracles[c].get_positions(sp, 0)
            assert np.max(np.abs(got[c] - ref)) <= 1e-12 * max(1.0, np.max(np.abs(ref))), (name, attempt, c)
        assert np.array_equal(n_dev, n_host), (name, attempt, n_dev, n_host)
    assert 0 < n_dev.sum() < 24 * C          # both outcomes occurred
    if path._n_k():
        for sp in range(len(cfg.species)):
            for c in range(C):
                assert np.max(np.abs(path.GetRhoK(sp, c, host.OLD_MODE) - oracles[c].rhok(sp, 0))) <= 1e-10 * cfg.species[sp].n_part
    for ai, act in enumerate(path.actions):
        du = act.DActionDBeta()
        for c in range(C):
            assert abs(du[c] - oracles[c].dbeta(ai)) <= 1e-10 * abs(oracles[c].dbeta(ai))
    # many attempts in one call == one by one (same stream)
    a = path.DisplaceSweep(0, step, 5, seed, attempt0=100)
    assert a.shape == (C,)
    path.close()
    for o in oracles:
        o.close()


def test_path_dump_and_restart_resume_the_same_markov_chain(tmp_path):
    """PathDump::Write + init_type="Restart" (path_dump_class.h:27-68, species_class.h:336-378):
    a run that dumps, is torn down, restarts from the file and continues with the attempt counter
    where it stopped ends bit-identical to the uninterrupted run (the Philox stream is a function
    of (seed, attempt, clone), so no generator state needs saving)."""
    from simpimc_b200 import host
    cfg = S.plasma_config(Ne=5, Np=4, M=16)
    C, n_level, seed = 4, 2, 0xD00D
    R0 = [np.stack([S.synthetic_paths(cfg, sp, c, 8) for c in range(C)]) for sp in range(2)]

    def run(path, a0, a1):
        for att in range(a0, a1):
            path.BisectSweep(att % 2, n_level, 3, seed, attempt0=3 * att)

    whole = host.Path(cfg, n_clones=C)
    for sp in range(2):
        whole.SetPositions(sp, R0[sp])
    run(whole, 0, 12)
    ref = [whole.GetPositions(sp) for sp in range(2)]
    ref_e = [a.DActionDBeta() for a in whole.actions]
    whole.close()
    first = host.Path(cfg, n_clones=C)
    for sp in range(2):
        first.SetPositions(sp, R0[sp])
    dump = host.PathDump(first, skip=2)
    out = host.IO(str(tmp_path / "run"), C)       # the reference's HDF5 layout, one file per walker (io_hdf5.h)
    energy = host.Energy(first, measure_potential=True)
    run(first, 0, 4)
    dump.Write(out)
    energy.Accumulate()
    energy.Write(out)
    run(first, 4, 7)
    dump.Write(out)            # skipped (skip = 2)
    dump.Write(out)
    energy.Accumulate()
    e_last, v_last = energy.Write(out)
    assert dump.n_dump == 2
    fn = str(tmp_path / "run.0.npz")
    dump.Save(fn)
    files = out.Save()
    first.close()
    # the HDF5 files carry the reference's dataset names and shapes (path_dump_class.h:54-62, energy_class.h:236-258)
    from simpimc_b200 import h5lite
    d0 = h5lite.read(files[0])
    assert d0["Observables/path_dump/e/positions"].shape == (2, 5, 16, 3) and d0["Observables/path_dump/p/permutation"].shape == (2, 4, 2)
    assert int(d0["Observables/path_dump/e/n_dump"]) == 2 and d0["Observables/Energy/data_type"] == "scalar"
    assert d0["Observables/Energy/total/x"].shape == (2,) and d0["Observables/Energy/v_CoulombEP/x"].shape == (2,)
    assert d0["Observables/Energy/total/x"][1] == e_last.sum(axis=0)[0] and d0["Observables/Energy/v_total/x"][1] == v_last.sum(axis=0)[0]
    third = host.Path(cfg, n_clones=C)
    host.PathDump.Restart(third, files)            # restart from the HDF5 files
    run(third, 7, 12)
    for sp in range(2):
        assert np.array_equal(third.GetPositions(sp), ref[sp])
    third.close()
    second = host.Path(cfg, n_clones=C)
    host.PathDump.Restart(second, fn)
    run(second, 7, 12)
    for sp in range(2):
        assert np.array_equal(second.GetPositions(sp), ref[sp])
    # rho_k was carried incrementally through the uninterrupted run and rebuilt at the restart
    # (Species::InitRhoK, as the reference does): the k-sums agree to rounding
    for a, e in zip(second.actions, ref_e):
        assert np.max(np.abs(a.DActionDBeta() - e) / np.abs(e)) <= 1e-12
    second.close()


@pytest.mark.parametrize("name", ["ueg", "plasma", "ueg_slack", "sharded"])
def test_multi_window_sweeps_follow_the_host_mirror(name):
    """pimc_bisect_sweep_windows: every disjoint window of every walker attempted in the same launches.  The host
    mirror walks the windows one after the other on the CPU oracle with the same Philox stream (keyed by walker * W +
    window): the windows share only fixed end-point slices, so the sequential walk and the simultaneous device
    update must leave the same positions (1e-12) and the same accept counts after every round."""
    from simpimc_b200 import host, moves, sharded
    from oracle import oracle as O
    n_shards = 1
    if name == "plasma":
        cfg, n_level = S.plasma_config(Ne=5, Np=4, M=16), 2          # W = 4, windows tile the ring: 4 offsets
    elif name == "ueg_slack":
        cfg, n_level = S.ueg_config(N=7, M=22), 2                    # W = 5, 2 slices of slack: 3 offsets
    elif name == "sharded":
        cfg, n_level, n_shards = S.ueg_config(N=7, M=20), 2, 2       # shards of 10 slices: W = 2, 3 offsets
    else:
        cfg, n_level = S.ueg_config(N=9, M=16), 3                    # W = 2, 8 offsets
    M, nb, n_clones, ns = cfg.n_bead, 1 << n_level, 2, len(cfg.species)
    Rs = [np.stack([S.synthetic_paths(cfg, sp, c, 777) for c in range(n_clones)]) for sp in range(ns)]
    oracles = [O.Oracle(cfg) for _ in range(n_clones)]
    for c, o in enumerate(oracles):
        for sp in range(ns):
            o.set_positions(sp, Rs[sp][c])
    shards = []
    for g in range(n_shards):
        sh = sharded.SliceSharding(M, n_shards, g)
        p = host.Path(cfg, n_clones=n_clones, slice_lo=sh.lo, slice_hi=sh.hi) if n_shards > 1 else host.Path(cfg, n_clones=n_clones)
        for sp in range(ns):
            p.SetPositions(sp, sh.shard_positions(Rs[sp]) if n_shards > 1 else Rs[sp])
        shards.append((sh, p))
    n_dev = np.zeros(n_clones, dtype=np.int64)
    n_host = np.zeros(n_clones, dtype=np.int64)
    seed = 0x77AA5500
    for rnd in range(10):
        sp = rnd % ns
        acts = [ai for ai, a in enumerate(cfg.actions) if cfg.species[sp].name in (a.species_a, a.species_b)]
        for g, (sh, p) in enumerate(shards):
            span = sh.n_local if n_shards > 1 else M
            W = span // nb
            n_off = nb if (n_shards == 1 and W * nb == span) else span - W * nb + 1

            def get_beads(vc, q, b0, n):
                return oracles[vc // W].get_positions(sp, 0)[q, (b0 + np.arange(n)) % M]

            def action_old_new(vc, q, b0, nb_, new):
                o = oracles[vc // W]
                o.propose(sp, q, (b0 + 1) % M, new)
                old = sum(o.get_action(ai, 0, b0, b0 + nb_, [(sp, q)], 0) for ai in acts)
                nw = sum(o.get_action(ai, 1, b0, b0 + nb_, [(sp, q)], 0) for ai in acts)
                return old, nw

            def finish(vc, q, b0, nb_, accept, new):
                if new is not None:
                    oracles[vc // W].finish_move(sp, q, b0, b0 + nb_, bool(accept))

            s_seed = seed + 31 * g
            _, b0s, acc = moves.bisect_attempt_philox(cfg, sp, n_level, s_seed, rnd, n_clones * W, get_beads, action_old_new, finish,
                                                      b0_range=(sh.lo if n_shards > 1 else 0, n_off), windows=W)
            b0s = np.asarray(b0s).reshape(n_clones, W)
            assert np.all(np.diff(b0s, axis=1) == nb)                       # disjoint windows sharing end points
            if n_shards > 1:
                assert np.all(b0s >= sh.lo) and np.all(b0s + nb <= sh.hi)
            n_host += np.asarray(acc).reshape(n_clones, W).sum(axis=1)
            got_acc, got_w = p.BisectSweepWindows(sp, n_level, 1, s_seed, attempt0=rnd)
            assert got_w == W
            n_dev += got_acc
            for s2 in range(ns):
                got = p.GetPositions(s2)
                for c in range(n_clones):
                    ref = oracles[c].get_positions(s2, 0)
                    if n_shards > 1:
                        ref = ref[:, np.array(sh.stored_slices()), :]
                    assert np.max(np.abs(got[c] - ref)) <= 1e-12 * cfg.L, (name, rnd, g, s2, c)
            assert np.array_equal(n_dev, n_host), (name, rnd, g, n_dev, n_host)
    assert n_dev.sum() > 0
    for ai in range(len(cfg.actions)):
        du = sum(p.actions[ai].DActionDBeta() for _, p in shards)
        for c, o in enumerate(oracles):
            ref = o.dbeta(ai)
            assert abs(du[c] - ref) <= 1e-10 * abs(ref), (name, ai, c)
    for sh, p in shards:
        for sp in range(ns):
            for c in range(n_clones):
                ref = oracles[c].rhok(sp, 0)
                if n_shards > 1:
                    ref = ref[sh.lo:sh.hi]
                assert np.max(np.abs(p.GetRhoK(sp, c, host.OLD_MODE) - ref)) <= 1e-10 * cfg.species[sp].n_part
    for _, p in shards:
        p.close()
    for o in oracles:
        o.close()
