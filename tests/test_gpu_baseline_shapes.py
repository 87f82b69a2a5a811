"""The BASELINE.json configurations as parity cases (configs[0], [1], [3] against the oracle at
their own sizes) and, at the full size of the headline configuration (UEG N=256, M=128), the
size-independent properties the domain offers: lattice-translation, uniform-translation,
imaginary-time-rotation and relabelling invariance, identical clones, histogram checksums,
incremental-versus-rebuilt rho_k after device sweeps."""
import numpy as np
import pytest

from simpimc_b200 import system as S

pytestmark = pytest.mark.gpu
RTOL = 1e-10


def rel_ok(got, ref, rtol=RTOL):
    got, ref = np.asarray(got, dtype=np.float64), np.asarray(ref, dtype=np.float64)
    return bool(np.all(np.abs(got - ref) <= rtol * np.maximum(np.abs(ref), 1e-300)))


def _against_oracle(cfg, Rs, windows, gofr=None):
    from simpimc_b200 import host
    from oracle import oracle as O
    C = len(Rs[0])
    path = host.Path(cfg, n_clones=C)
    oracles = []
    for sp in range(len(cfg.species)):
        path.SetPositions(sp, Rs[sp])
    for c in range(C):
        o = O.Oracle(cfg)
        for sp in range(len(cfg.species)):
            o.set_positions(sp, Rs[sp][c])
        oracles.append(o)
    parts = [(s, p) for s in range(len(cfg.species)) for p in range(cfg.species[s].n_part)]
    for ai, act in enumerate(path.actions):
        du, v, u = act.DActionDBeta(), act.Potential(), act.TotalAction()
        sa = cfg.species[act.species_a]
        # for a constant action (one particle, or lambda = 0, same species) the
        # reference caches whichever of DActionDBeta / Potential runs first and returns an
        # uninitialised member for the other; only the first call is comparable
        constant = act.species_a == act.species_b and (sa.n_part == 1 or sa.lam == 0.0)
        for c, o in enumerate(oracles):
            assert rel_ok(du[c], o.dbeta(ai)), (ai, c, du[c], o.dbeta(ai))
            if constant:
                continue
            assert rel_ok(v[c], o.potential(ai)), (ai, c, v[c], o.potential(ai))
            assert rel_ok(u[c], o.get_action(ai, 0, 0, cfg.n_bead, parts, 0)), (ai, c)
    rng = np.random.default_rng(3)
    M = cfg.n_bead
    for sp, nb in windows:
        N = cfg.species[sp].n_part
        part = rng.integers(0, N, C)
        b0 = rng.integers(0, M, C)
        first = (b0 + 1) % M
        old_pos = np.stack([Rs[sp][c][part[c], (first[c] + np.arange(nb - 1)) % M] for c in range(C)])
        newR = old_pos + 0.05 * rng.standard_normal(old_pos.shape)
        path.Propose(sp, part, first, newR)
        for c, o in enumerate(oracles):
            o.propose(sp, int(part[c]), int(first[c]), newR[c])
        for ai, act in enumerate(path.actions):
            if sp not in (act.species_a, act.species_b):
                continue
            path.SetMode(host.OLD_MODE)
            old = act.GetAction(b0, b0 + nb, [(sp, part)], 0)
            path.SetMode(host.NEW_MODE)
            new = act.GetAction(b0, b0 + nb, [(sp, part)], 0)
            for c, o in enumerate(oracles):
                ro = o.get_action(ai, 0, int(b0[c]), int(b0[c]) + nb, [(sp, int(part[c]))], 0)
                rn = o.get_action(ai, 1, int(b0[c]), int(b0[c]) + nb, [(sp, int(part[c]))], 0)
                assert rel_ok(old[c], ro) and rel_ok(new[c], rn), (sp, nb, ai, c)
                assert abs((new[c] - old[c]) - (rn - ro)) <= RTOL * max(abs(rn - ro), 1e-4 * (abs(rn) + abs(ro)))
        path.Commit(0)
        for c, o in enumerate(oracles):
            o.finish_move(sp, int(part[c]), int(b0[c]), int(b0[c]) + nb, False)
    if gofr:
        sa, sb, r_min, r_max, n_r = gofr
        counts = host.PairCorrelation(path, sa, sb, r_min, r_max, n_r).Counts()
        for c, o in enumerate(oracles):
            assert np.array_equal(counts[c], o.gofr(sa, sb, r_min, r_max, n_r)[1])
    for o in oracles:
        o.close()
    path.close()


@pytest.mark.parametrize("N", [7, 33])
def test_config_c1_egas(N):
    cfg = S.egas_config(N=N, M=1280)
    Rs = [np.stack([S.synthetic_paths(cfg, 0, c, 8) for c in range(2)])]
    _against_oracle(cfg, Rs, [(0, 8), (0, 32)], gofr=(0, 0, 0.0, cfg.L / 2, 100))


def test_config_c2_hydrogen_atom_open_boundary():
    cfg = S.hatom_config()
    per_clone = [S.hatom_paths(cfg, c) for c in range(3)]
    Rs = [np.stack([pc[sp] for pc in per_clone]) for sp in range(2)]
    _against_oracle(cfg, Rs, [(0, 32), (0, 8)], gofr=(1, 0, 0.0, 5.0, 100))


def test_config_c4_warm_dense_carbon():
    cfg = S.carbon_config()
    Rs = [np.stack([S.synthetic_paths(cfg, sp, c, 21) for c in range(2)]) for sp in range(4)]
    _against_oracle(cfg, Rs, [(0, 8), (2, 4), (1, 8)], gofr=(0, 2, 0.0, cfg.L / 2, 100))


def test_headline_config_against_oracle():
    """The configuration the metric is quoted on -- UEG N=256, M=128, IlkkaPairAction with long
    range -- compared DIRECTLY with the CPU oracle on two clones (pair_action_class.h:241-264,
    ilkka_pair_action_class.h:125-169): DActionDBeta, Potential, the whole-path action, an
    n_level = 3 bisection window and a 32-slice window in OLD / NEW mode with their differences,
    a whole-path DisplaceParticle difference, rho_k, the g(r) bin counts bit for bit and S(k)."""
    from simpimc_b200 import host
    from oracle import oracle as O
    cfg = S.ueg_config(N=256, M=128)
    C, N, M = 2, 256, 128
    R = np.stack([S.synthetic_paths(cfg, 0, c) for c in range(C)])
    _against_oracle(cfg, [R], [(0, 8), (0, 32)], gofr=(0, 0, 0.0, cfg.L / 2, 100))
    path = host.Path(cfg, n_clones=C)
    path.SetPositions(0, R)
    act = path.actions[0]
    sk = host.StructureFactor(path, 0, 0, cfg.k_cut)
    sk.Accumulate()
    rng = np.random.default_rng(11)
    part = rng.integers(0, N, C)
    shift = 0.1 * cfg.L * rng.standard_normal((C, 1, 3))
    newR = np.stack([R[c, part[c]] for c in range(C)]) + shift      # DisplaceParticle: every bead of one particle
    path.Propose(0, part, np.zeros(C, dtype=np.int32), newR)
    path.SetMode(host.OLD_MODE)
    old = act.GetAction(0, M, [(0, part)], 0)
    path.SetMode(host.NEW_MODE)
    new = act.GetAction(0, M, [(0, part)], 0)
    for c in range(C):
        o = O.Oracle(cfg)
        o.set_positions(0, R[c])
        assert np.max(np.abs(path.GetRhoK(0, c, host.OLD_MODE) - o.rhok(0, 0))) <= 1e-12 * N
        ref_sk = o.sofk(0, 0, cfg.k_cut)
        assert np.max(np.abs(sk.sk[c] - ref_sk)) <= 1e-10 * max(1.0, np.max(np.abs(ref_sk)))
        o.propose(0, int(part[c]), 0, newR[c])
        ro = o.get_action(0, 0, 0, M, [(0, int(part[c]))], 0)
        rn = o.get_action(0, 1, 0, M, [(0, int(part[c]))], 0)
        assert rel_ok(old[c], ro) and rel_ok(new[c], rn), (c, old[c], ro, new[c], rn)
        assert abs((new[c] - old[c]) - (rn - ro)) <= RTOL * max(abs(rn - ro), 1e-4 * (abs(rn) + abs(ro)))
        o.close()
    path.Commit(0)
    path.close()


def test_config_c5_shaped_shards_against_oracle():
    """BASELINE config C5's shape at a slice count the oracle finishes in seconds: 1024 e + 1024 p,
    M = 8, three Ilkka actions with long range.  The whole path and the sum of two slice shards
    (4 + 4 slices, halo included) against the oracle's DActionDBeta / Potential of the whole path
    (pair_action_class.h:282-288 couples slice b with b + 1 only; rho_k is slice-local,
    species_class.h:391-395)."""
    from simpimc_b200 import host, sharded
    from oracle import oracle as O
    cfg = S.plasma_config(Ne=1024, Np=1024, M=8, n_xy=100, n_r_long=1000, pp_action="IlkkaPairAction")
    Rs = [S.synthetic_paths(cfg, sp, 0, 778)[None] for sp in range(2)]
    o = O.Oracle(cfg)
    for sp in range(2):
        o.set_positions(sp, Rs[sp][0])
    ref = np.array([[o.dbeta(a), o.potential(a)] for a in range(3)])
    whole = host.Path(cfg, n_clones=1)
    for sp in range(2):
        whole.SetPositions(sp, Rs[sp])
        assert np.max(np.abs(whole.GetRhoK(sp, 0, host.OLD_MODE) - o.rhok(sp, 0))) <= 1e-12 * 1024
    got = np.array([[a.DActionDBeta()[0], a.Potential()[0]] for a in whole.actions])
    assert rel_ok(got, ref), (got, ref)
    counts = host.PairCorrelation(whole, 0, 1, 0.0, cfg.L / 2, 100).Counts()[0]
    assert np.array_equal(counts, o.gofr(0, 1, 0.0, cfg.L / 2, 100)[1])
    whole.close()
    parts = np.zeros_like(ref)
    for g in range(2):
        sh = sharded.SliceSharding(cfg.n_bead, 2, g)
        p = host.Path(cfg, n_clones=1, slice_lo=sh.lo, slice_hi=sh.hi)
        for sp in range(2):
            p.SetPositions(sp, sh.shard_positions(Rs[sp]))
        parts += np.array([[a.DActionDBeta()[0], a.Potential()[0]] for a in p.actions])
        p.close()
    assert rel_ok(parts, ref), (parts, ref)
    o.close()


def test_headline_size_invariances():
    """UEG N=256, M=128 (the bench configuration), 8 clones."""
    from simpimc_b200 import host
    cfg = S.ueg_config(N=256, M=128)
    C, N, M, L = 8, 256, 128, cfg.L
    rng = np.random.default_rng(17)
    R = np.stack([S.synthetic_paths(cfg, 0, c % 4) for c in range(C)])     # clones 4..7 repeat 0..3
    path = host.Path(cfg, n_clones=C)
    act = path.actions[0]

    def evaluate(Rx):
        path.SetPositions(0, Rx)
        return act.DActionDBeta(), act.TotalAction(), act.Potential()

    base = evaluate(R)
    for q in base:
        assert np.array_equal(q[:4], q[4:]), "identical clones must give identical bits"
    # histogram checksum: with r_max beyond the largest minimum-image distance every pair-slice is counted
    counts = host.PairCorrelation(path, 0, 0, 0.0, 0.9 * L, 200).Counts()
    assert np.all(counts.sum(axis=1) == N * (N - 1) // 2 * M)
    sk = host.StructureFactor(path, 0, 0, cfg.k_cut)
    sk.Accumulate()
    sk0 = sk.sk.copy()
    assert np.all(sk0 >= 0.0)
    # every particle's whole path moved by its own lattice vector: minimum-image quantities unchanged
    shift = rng.integers(-2, 3, size=(C, N, 1, 3)) * L
    for got, ref in zip(evaluate(R + shift), base):
        assert rel_ok(got, ref), np.max(np.abs(got - ref) / np.abs(ref))
    # rigid translation of everything, rotation of the imaginary-time origin, relabelling of particles
    for Rx in (R + np.array([0.123, -4.5, 2.718]), np.roll(R, 37, axis=2), R[:, rng.permutation(N)]):
        for got, ref in zip(evaluate(Rx), base):
            assert rel_ok(got, ref), np.max(np.abs(got - ref) / np.abs(ref))
    sk = host.StructureFactor(path, 0, 0, cfg.k_cut)
    sk.Accumulate()   # S(k) of the relabelled configuration
    assert np.max(np.abs(sk.sk - sk0)) <= 1e-9 * np.max(sk0)
    path.close()


@pytest.mark.parametrize("family,lr", [("BarePairAction", True), ("DavidPairAction", False)])
def test_headline_size_other_families(family, lr):
    """The Bare and David whole-path kernels at the bench shape (N=256, M=128: eight particle groups, four slice
    chunks, four staged partner windows, 32 parked r' per ring flush): the shared-memory fast kernels against the
    general kernels of the same library (independent code: pp form + interval tables vs B-spline form / bit LUT),
    identical clones bit-identical, and the size-independent invariances."""
    from simpimc_b200 import host
    cfg = S.ueg_config(N=256, M=128, action=family, use_long_range=lr)
    C, N, L = 4, 256, cfg.L
    rng = np.random.default_rng(23)
    R = np.stack([S.synthetic_paths(cfg, 0, c % 2) for c in range(C)])     # clones 2, 3 repeat 0, 1
    path = host.Path(cfg, n_clones=C)
    act = path.actions[0]

    def evaluate(Rx):
        path.SetPositions(0, Rx)
        return act.DActionDBeta(), act.TotalAction(), act.Potential()

    base = evaluate(R)
    for q in base:
        assert np.array_equal(q[:2], q[2:]), "identical clones must give identical bits"
    path.ForceGeneral(True)
    for got, ref in zip(evaluate(R), base):
        assert rel_ok(ref, got), (family, np.max(np.abs(got - ref) / np.abs(ref)))
    path.ForceGeneral(False)
    shift = rng.integers(-2, 3, size=(C, N, 1, 3)) * L
    for Rx in (R + shift, R + np.array([0.123, -4.5, 2.718]), np.roll(R, 37, axis=2), R[:, rng.permutation(N)]):
        for got, ref in zip(evaluate(Rx), base):
            assert rel_ok(got, ref), (family, np.max(np.abs(got - ref) / np.abs(ref)))
    path.close()


def test_headline_size_sweeps_keep_rhok_and_energy_consistent():
    """After device-resident sweeps at full size, the incrementally updated rho_k and the
    energies equal those of a fresh context built from the downloaded positions."""
    from simpimc_b200 import host
    cfg = S.ueg_config(N=256, M=128)
    C = 4
    path = host.Path(cfg, n_clones=C)
    path.SetPositions(0, np.stack([S.synthetic_paths(cfg, 0, c) for c in range(C)]))
    before = path.actions[0].DActionDBeta()
    n_acc = path.BisectSweep(0, 3, 96, seed=5)
    assert np.all(n_acc > 0) and np.all(n_acc <= 96)
    after = path.actions[0].DActionDBeta()
    assert not np.array_equal(before, after)
    R = path.GetPositions(0)
    rho = np.stack([path.GetRhoK(0, c, host.OLD_MODE) for c in range(C)])
    fresh = host.Path(cfg, n_clones=C)
    fresh.SetPositions(0, R)
    assert rel_ok(after, fresh.actions[0].DActionDBeta())
    rho_f = np.stack([fresh.GetRhoK(0, c, host.OLD_MODE) for c in range(C)])
    assert np.max(np.abs(rho - rho_f)) <= 1e-10 * 256
    assert np.array_equal(path.BisectSweep(0, 3, 0, seed=5), np.zeros(C, dtype=np.int64))
    assert np.array_equal(path.GetPositions(0), R)
    path.close()
    fresh.close()


def test_config_c5_full_size_shards_sum_to_the_whole_path():
    """BASELINE config C5 at full size (1024 e + 1024 p, M = 512, three Ilkka actions with long
    range): the oracle would take minutes, so the checks are size-independent properties -- eight
    slice shards (the 8-GPU layout, here on one device) sum to the unsharded DActionDBeta /
    Potential / g(r) counts; a rigid translation, a lattice shift of individual paths and a
    rotation of the imaginary-time origin leave the energies unchanged; every pair-slice lands in
    the histogram when r_max exceeds the largest minimum-image distance."""
    from simpimc_b200 import host, sharded
    cfg = S.plasma_config(Ne=1024, Np=1024, M=512, n_xy=100, n_r_long=1000, pp_action="IlkkaPairAction")
    Rs = [S.synthetic_paths(cfg, sp, 0, 777)[None] for sp in range(2)]
    whole = host.Path(cfg, n_clones=1)

    def evaluate(Rx):
        for sp in range(2):
            whole.SetPositions(sp, Rx[sp])
        return np.array([[a.DActionDBeta()[0], a.Potential()[0]] for a in whole.actions])

    base = evaluate(Rs)
    counts = host.PairCorrelation(whole, 0, 1, 0.0, 0.9 * cfg.L, 200).Counts()[0]
    assert counts.sum() == 1024 * 1024 * 512
    parts = np.zeros_like(base)
    counts_sh = np.zeros_like(counts)
    for g in range(8):
        sh = sharded.SliceSharding(cfg.n_bead, 8, g)
        p = host.Path(cfg, n_clones=1, slice_lo=sh.lo, slice_hi=sh.hi)
        for sp in range(2):
            p.SetPositions(sp, sh.shard_positions(Rs[sp]))
        parts += np.array([[a.DActionDBeta()[0], a.Potential()[0]] for a in p.actions])
        counts_sh += host.PairCorrelation(p, 0, 1, 0.0, 0.9 * cfg.L, 200).Counts()[0]
        p.close()
    assert rel_ok(parts, base, rtol=1e-11), np.max(np.abs(parts - base) / np.abs(base))
    assert np.array_equal(counts_sh, counts)
    rng = np.random.default_rng(3)
    shift = [rng.integers(-2, 3, size=(1, 1024, 1, 3)) * cfg.L for _ in range(2)]
    for Rx in ([R + np.array([0.31, -7.0, 1.5]) for R in Rs], [R + s for R, s in zip(Rs, shift)], [np.roll(R, 101, axis=2) for R in Rs]):
        got = evaluate(Rx)
        assert rel_ok(got, base), np.max(np.abs(got - base) / np.abs(base))
    whole.close()
