"""GPU parity tests proper: the CUDA path, called through the C ABI, against the CPU oracle on
the same seeded inputs.  Tolerance: |gpu - oracle| <= 1e-10 * |oracle| for floating-point
results (BASELINE.json north_star), bit-exact for k vectors and histogram bins."""
import numpy as np
import pytest

from simpimc_b200 import system as S

pytestmark = pytest.mark.gpu

RTOL = 1e-10


def rel_ok(got, ref, scale=None, rtol=RTOL):
    """|got - ref| <= rtol * max(|ref|, scale).  `scale` is the size of the terms a value is
    a (possibly cancelling) combination of; without it the bound is purely relative."""
    got, ref = np.asarray(got, dtype=np.float64), np.asarray(ref, dtype=np.float64)
    sc = np.abs(ref) if scale is None else np.maximum(np.abs(ref), scale)
    return np.all(np.abs(got - ref) <= rtol * np.maximum(sc, 1e-300))


def make_pair(cfg, n_clones, seed=12345):
    from simpimc_b200 import host
    from oracle import oracle as O
    path = host.Path(cfg, n_clones=n_clones)
    oracles = []
    Rs = []
    for sp in range(len(cfg.species)):
        R = np.stack([S.synthetic_paths(cfg, sp, c, seed) for c in range(n_clones)])
        path.SetPositions(sp, R)
        Rs.append(R)
    for c in range(n_clones):
        o = O.Oracle(cfg)
        for sp in range(len(cfg.species)):
            o.set_positions(sp, Rs[sp][c])
        oracles.append(o)
    return path, oracles, Rs


CONFIGS = {
    "ilkka_lr_n7": lambda: S.ueg_config(N=7, M=8),
    "ilkka_lr_n33": lambda: S.ueg_config(N=33, M=16),
    "ilkka_nolr_n8": lambda: S.ueg_config(N=8, M=8, use_long_range=False),
    "bare_lr_n7": lambda: S.ueg_config(N=7, M=8, action="BarePairAction"),
    "david_n7": lambda: S.ueg_config(N=7, M=8, action="DavidPairAction", use_long_range=False),
    "david_lr_n7": lambda: S.ueg_config(N=7, M=8, action="DavidPairAction", use_long_range=True),
    "plasma": lambda: S.plasma_config(Ne=6, Np=5, M=8),
    # off-diagonal table with n_x != n_y, different x and y grids and no x <-> y symmetry (a transposition would show)
    "ilkka_asym": lambda: S.ueg_config(N=9, M=8, n_xy=80, xy_asym=(64, 60.0, 0.3)),
    "n2": lambda: S.ueg_config(N=2, M=4),
    # David tables on a linear grid (uniform interval table), M > 32 (two slice chunks, lane-31 ring), even N
    "david_lin_n34": lambda: S.ueg_config(N=34, M=40, action="DavidPairAction", use_long_range=False, david_grid="LINEAR",
                                          david_n_grid=120),
    "david_log_n33": lambda: S.ueg_config(N=33, M=40, action="DavidPairAction", use_long_range=True),
    # Ilkka e-e, David e-p (different species) and David p-p on a linear grid; M > 32
    "plasma_david": lambda: S.plasma_config(Ne=6, Np=5, M=36, pp_action="DavidPairAction", ep_action="DavidPairAction"),
    # more than 32 partner offsets per particle: several staged windows, partner loop split over CTAs, three particle groups
    "david_log_n70": lambda: S.ueg_config(N=70, M=40, action="DavidPairAction", use_long_range=False),
    "david_o1_n9": lambda: S.ueg_config(N=9, M=8, action="DavidPairAction", use_long_range=False, david_n_order=1),
    "david_o3_n9": lambda: S.ueg_config(N=9, M=8, action="DavidPairAction", use_long_range=False, david_n_order=3,
                                        david_grid="LINEAR", david_n_grid=90),
}


@pytest.mark.parametrize("name", sorted(CONFIGS))
def test_full_path_values(name):
    cfg = CONFIGS[name]()
    path, oracles, _ = make_pair(cfg, 3)
    for ai, act in enumerate(path.actions):
        if act.type == "Kinetic":
            continue
        du, u = act.DActionDBeta(), act.TotalAction()
        # GetAction(0, M, every particle, 0) in OLD mode is the whole-path action
        parts = [(s, p) for s in range(len(cfg.species)) for p in range(cfg.species[s].n_part)]
        for c, o in enumerate(oracles):
            assert rel_ok(du[c], o.dbeta(ai)), (name, ai, c, du[c], o.dbeta(ai))
            ref_u = o.get_action(ai, 0, 0, cfg.n_bead, parts, 0)
            assert rel_ok(u[c], ref_u), (name, ai, c, u[c], ref_u)
        if not (cfg.actions[ai].type == "DavidPairAction" and cfg.actions[ai].use_long_range):
            v = act.Potential()
            for c, o in enumerate(oracles):
                assert rel_ok(v[c], o.potential(ai)), (name, ai, c, v[c], o.potential(ai))
    path.close()


@pytest.mark.parametrize("name", ["ilkka_lr_n7", "bare_lr_n7", "david_n7", "plasma", "ilkka_asym"])
def test_per_pair_kernels(name):
    cfg = CONFIGS[name]()
    path, oracles, _ = make_pair(cfg, 1)
    rng = np.random.default_rng(3)
    n = 4000
    rmax = np.sqrt(3) * cfg.L / 2
    r = rng.uniform(1e-3, rmax, n)
    rp = np.clip(r + rng.normal(0, 0.1, n), 1e-4, rmax)
    s = np.abs(r - rp) + np.abs(rng.normal(0, 0.05, n))
    # edge cases: zero separation change, grid ends, far beyond the long-range grid
    r[:4] = [1e-5, rmax, 0.3, 2.0]
    rp[:4] = [1e-5, rmax * 1.5, 0.3, 2.0]
    s[:4] = [0.0, 0.1, 0.0, 0.0]
    for ai, act in enumerate(path.actions):
        if act.type == "Kinetic":
            continue
        for which in (0, 1, 2):
            got = act.CalcPair(which, r, rp, s)
            ref = oracles[0].calc_pair(ai, which, r, rp, s)
            # a single pair value is short-range minus long-range spline terms of size
            # ~max|ref| each, so it can be a small residue of them: bound relative to that size
            assert rel_ok(got, ref, scale=1e-3 * np.max(np.abs(ref))), (name, ai, which, np.max(np.abs(got - ref)))
    path.close()


@pytest.mark.parametrize("name", ["ilkka_lr_n7", "ilkka_lr_n33", "plasma"])
def test_kspace_and_rhok(name):
    cfg = CONFIGS[name]()
    path, oracles, _ = make_pair(cfg, 2)
    idx, mags = path.KSpace()
    oidx, omags = oracles[0].kspace()
    assert np.array_equal(idx, oidx)          # integer work: bit-exact, reference order
    assert np.array_equal(mags, omags)
    for sp in range(len(cfg.species)):
        for c, o in enumerate(oracles):
            got, ref = path.GetRhoK(sp, c), o.rhok(sp)
            assert np.max(np.abs(got - ref)) <= 1e-12 * cfg.species[sp].n_part
    path.close()


@pytest.mark.parametrize("name", ["ilkka_lr_n7", "ilkka_lr_n33", "bare_lr_n7", "david_n7", "plasma", "ilkka_nolr_n8", "ilkka_asym"])
def test_move_windows_old_new_commit(name):
    """Bisect-style windows (incl. wrap-around b1 > n_bead) and a displace-style whole-path
    move: OLD and NEW GetAction, their difference, and the state after accept / reject."""
    cfg = CONFIGS[name]()
    n_clones = 4
    path, oracles, Rs = make_pair(cfg, n_clones)
    from simpimc_b200 import host
    rng = np.random.default_rng(11)
    M = cfg.n_bead
    for trial in range(6):
        sp = trial % len(cfg.species)
        N = cfg.species[sp].n_part
        nb = [4, 2, M, 4, 1, M][trial]          # links in the window; M = displace
        part = rng.integers(0, N, n_clones)
        b0 = rng.integers(0, M, n_clones)
        if trial == 3:
            b0[:] = M - 2                       # forces wrap-around
        n_beads = nb - 1 if nb < M else M
        first = (b0 + 1) % M if nb < M else np.zeros(n_clones, dtype=int)
        if nb == M:
            b0[:] = 0
        if n_beads == 0:
            continue
        old_pos = np.stack([oracles[c].get_positions(sp, 0)[part[c], (first[c] + np.arange(n_beads)) % M] for c in range(n_clones)])
        newR = old_pos + 0.08 * rng.standard_normal(old_pos.shape)
        path.Propose(sp, part, first, newR)
        accept = rng.integers(0, 2, n_clones)
        for c, o in enumerate(oracles):
            o.propose(sp, int(part[c]), int(first[c]), newR[c])
        sums = {}
        for ai, act in enumerate(path.actions):
            if act.type == "Kinetic" or sp not in (act.species_a, act.species_b):
                continue
            path.SetMode(host.OLD_MODE)
            old = act.GetAction(b0, b0 + nb, [(sp, part)], 0)
            path.SetMode(host.NEW_MODE)
            new = act.GetAction(b0, b0 + nb, [(sp, part)], 0)
            for c, o in enumerate(oracles):
                ro = o.get_action(ai, 0, int(b0[c]), int(b0[c]) + nb, [(sp, int(part[c]))], 0)
                rn = o.get_action(ai, 1, int(b0[c]), int(b0[c]) + nb, [(sp, int(part[c]))], 0)
                assert rel_ok(old[c], ro), (name, trial, ai, c, old[c], ro)
                assert rel_ok(new[c], rn), (name, trial, ai, c, new[c], rn)
                # the difference a move tests, relative to the size of the sums it comes from
                assert abs((new[c] - old[c]) - (rn - ro)) <= RTOL * max(abs(rn - ro), 1e-4 * (abs(rn) + abs(ro)))
        path.Commit(accept)
        for c, o in enumerate(oracles):
            lo = int(b0[c])
            # bisect commits beads bead0..bead1 and slices [bead0, bead1) (bisect_class.h:24-36);
            # displace commits every bead and slice (displace_particle_class.h:13-25)
            o.finish_move(sp, int(part[c]), lo, lo + nb, bool(accept[c]))
        got = path.GetPositions(sp)
        for c, o in enumerate(oracles):
            assert np.array_equal(got[c], o.get_positions(sp, 0)), (name, trial, c)
            if path._n_k():
                assert np.max(np.abs(path.GetRhoK(sp, c, host.OLD_MODE) - o.rhok(sp, 0))) <= 1e-11 * N
    # after the moves the full-path values still agree
    for ai, act in enumerate(path.actions):
        if act.type == "Kinetic":
            continue
        du = act.DActionDBeta()
        for c, o in enumerate(oracles):
            assert rel_ok(du[c], o.dbeta(ai)), (name, "post", ai, c)
    path.close()


@pytest.mark.parametrize("name", ["ilkka_lr_n7", "ilkka_lr_n33", "plasma"])
def test_estimators(name):
    from simpimc_b200 import host
    cfg = CONFIGS[name]()
    path, oracles, _ = make_pair(cfg, 3)
    ns = len(cfg.species)
    for sa in range(ns):
        for sb in range(sa, ns):
            gr = host.PairCorrelation(path, sa, sb, 0.0, cfg.L / 2, 100)
            counts = gr.Counts()
            gr.Accumulate(cofactor=np.array([1.0, -1.0, 1.0]))
            sk = host.StructureFactor(path, sa, sb, cfg.k_cut)
            sk.Accumulate()
            sk.Accumulate(cofactor=np.array([1.0, -1.0, 1.0]))
            for c, o in enumerate(oracles):
                y, oc = o.gofr(sa, sb, 0.0, cfg.L / 2, 100, cofactor=[1.0, -1.0, 1.0][c])
                assert np.array_equal(counts[c], oc)              # bins: bit-exact
                assert np.array_equal(gr.y[c], y)
                ref = o.sofk(sa, sb, cfg.k_cut) * (1.0 + [1.0, -1.0, 1.0][c])
                assert np.max(np.abs(sk.sk[c] - ref)) <= 1e-10 * max(1.0, np.max(np.abs(o.sofk(sa, sb, cfg.k_cut))))
    path.close()


def test_energy_observable_matches_oracle():
    from simpimc_b200 import host
    cfg = S.plasma_config(Ne=6, Np=5, M=8)
    path, oracles, _ = make_pair(cfg, 2)
    en = host.Energy(path, measure_potential=True)
    en.Accumulate()
    e, v = en.Write()
    for c, o in enumerate(oracles):
        for i in range(len(cfg.actions)):
            assert rel_ok(e[i, c] * cfg.n_bead, o.dbeta(i))
            assert rel_ok(v[i, c] * cfg.n_bead, o.potential(i))
    path.close()


def test_errors_are_loud():
    from simpimc_b200 import host
    cfg = S.ueg_config(N=7, M=8)
    cfg.actions[0].max_level = 1
    with pytest.raises(RuntimeError):
        host.Path(cfg, n_clones=1)
    cfg = S.ueg_config(N=7, M=8)
    cfg.n_d = 2
    with pytest.raises(RuntimeError):
        host.Path(cfg, n_clones=1)
    # misuse of the move / derivative / sharding entry points: an error code and a message, never a silent result
    cfg = S.ueg_config(N=7, M=8)
    path = host.Path(cfg, n_clones=2)
    R = np.stack([S.synthetic_paths(cfg, 0, c) for c in range(2)])
    path.SetPositions(0, R)
    act = path.actions[0]
    path.Propose(0, 1, 2, R[:, 1, 2:4, :] + 0.01)
    for bad in (lambda: path.Propose(0, 1, 2, R[:, 1, 2:4, :]),              # the same particle proposed twice
                lambda: path.Propose(0, 3, 2, R[:, 3, 2:5, :]),              # another bead count on the same species
                lambda: act.GetActionGradient(1, 3, [(0, 1)], 0),            # derivatives while a proposal is pending
                lambda: path.BisectSweep(0, 2, 1, 1),                        # a device sweep while a proposal is pending
                lambda: path.DisplaceSweep(0, 0.1, 1, 1)):
        with pytest.raises(RuntimeError):
            bad()
    path.Commit(0)
    for bad in (lambda: path.DisplaceSweep(0, -1.0, 1, 1),                   # non-positive step
                lambda: path.BisectSweep(0, 4, 1, 1),                        # window longer than the path (2^4 > 8)
                lambda: path.PermTable(0, 0, 9),                             # n_bisect_beads > n_bead
                lambda: path.PermTable(0, 8, 2),                             # window start out of range
                lambda: act.GetAction(0, 2, [(0, 7)], 0),                    # particle out of range
                lambda: act.GetAction(0, 2, [(0, 1), (0, 1)], 0)):           # particle listed twice
        with pytest.raises(RuntimeError):
            bad()
    path.close()
    shard = host.Path(cfg, n_clones=1, slice_lo=0, slice_hi=4)
    shard.SetPositions(0, R[:1, :, [0, 1, 2, 3, 4], :])
    for bad in (lambda: shard.DisplaceSweep(0, 0.1, 1, 1),                   # whole-path move on a slice shard
                lambda: shard.BisectSweep(0, 3, 1, 1),                       # window longer than the shard
                lambda: shard.actions[0].GetAction(0, 2, [(0, 1)], 0)):      # host-driven windows on a shard
        with pytest.raises(RuntimeError):
            bad()
    shard.close()


@pytest.mark.parametrize("name", ["ilkka_lr_n7", "ilkka_lr_n33", "ilkka_nolr_n8", "plasma", "n2", "ilkka_asym", "david_n7", "david_lr_n7",
                                  "david_lin_n34", "david_log_n33", "david_o1_n9", "david_o3_n9", "plasma_david", "david_log_n70"])
def test_fast_and_general_kernels_agree_with_oracle(name):
    """The whole-path Ilkka, Bare and David evaluations run through the shared-memory fast kernels
    (pair_fast.cuh) by default; the general kernel stays for tables the fast layouts cannot hold.
    Both must match the oracle."""
    cfg = CONFIGS[name]()
    path, oracles, _ = make_pair(cfg, 3)
    for general in (False, True):
        path.ForceGeneral(general)
        for ai, act in enumerate(path.actions):
            du, u = act.DActionDBeta(), act.TotalAction()
            parts = [(s, p) for s in range(len(cfg.species)) for p in range(cfg.species[s].n_part)]
            for c, o in enumerate(oracles):
                assert rel_ok(du[c], o.dbeta(ai)), (name, general, ai, c, du[c], o.dbeta(ai))
                assert rel_ok(u[c], o.get_action(ai, 0, 0, cfg.n_bead, parts, 0)), (name, general, ai, c)
            if not (cfg.actions[ai].type == "DavidPairAction" and cfg.actions[ai].use_long_range):
                v = act.Potential()   # fast: one distance per pair and slice (independent images), every family
                for c, o in enumerate(oracles):
                    assert rel_ok(v[c], o.potential(ai)), (name, general, ai, c, v[c], o.potential(ai))
    path.close()


def test_fast_per_pair_matches_oracle_and_general():
    cfg = CONFIGS["ilkka_lr_n33"]()
    path, oracles, _ = make_pair(cfg, 1)
    rng = np.random.default_rng(5)
    n = 20000
    rmax = np.sqrt(3) * cfg.L / 2
    r = rng.uniform(1e-5, rmax, n)
    rp = np.clip(r + rng.normal(0, 0.1, n), 1e-5, 1.2 * rmax)
    s = np.abs(r - rp) + np.abs(rng.normal(0, 0.05, n))
    # knots themselves, grid ends, zero separation, far outside the staged block of cells
    tab = cfg.actions[0].table
    knots = np.asarray(tab["u/diag/r_long"])[:200]
    r[:200] = knots
    rp[:200] = knots
    s[:200] = 0.0
    xk = np.asarray(tab["u/off_diag/x"])[1:40]
    r[200:239] = xk
    rp[200:239] = xk
    s[200:239] = 0.0
    r[300:310] = np.linspace(20.0, 99.0, 10)
    rp[300:310] = r[300:310] + 0.3
    s[300:310] = 0.5
    act = path.actions[0]
    for which in (0, 1):
        fast = act.CalcPairFast(which, r, rp, s)
        gen = act.CalcPair(which, r, rp, s)
        ref = oracles[0].calc_pair(0, which, r, rp, s)
        scale = 1e-3 * np.max(np.abs(ref))
        assert rel_ok(fast, ref, scale=scale), (which, np.max(np.abs(fast - ref)))
        assert rel_ok(fast, gen, scale=scale)
    path.close()


@pytest.mark.parametrize("name", ["david_n7", "david_lin_n34", "david_o1_n9", "david_o3_n9"])
def test_david_fast_per_pair_matches_oracle_and_general(name):
    """DavidPairAction::CalcU / CalcdUdBeta through the fast kernel's tables (pp-form multi-spline,
    bit-pattern or uniform interval table) on triples that include the knots themselves, both grid
    ends, distances outside the grid on either side (SetLimits), s = 0 (no off-diagonal term) and
    q >= r_max (david_pair_action_class.h:80)."""
    cfg = CONFIGS[name]()
    path, oracles, _ = make_pair(cfg, 1)
    rng = np.random.default_rng(6)
    n = 20000
    tab = cfg.actions[0].table
    grp = "u_kj_%d" % cfg.actions[0].n_order
    r0, r1, ng = float(tab[grp + "/grid/start"]), float(tab[grp + "/grid/end"]), int(tab[grp + "/grid/n_grid_points"])
    if tab[grp + "/grid/type"] == "LOG":
        knots = r0 * np.exp(np.arange(ng) * (np.log(r1 / r0) / (ng - 1)))
    else:
        knots = np.linspace(r0, r1, ng)
    r = rng.uniform(0.2 * r0, 1.1 * r1, n)
    rp = np.clip(r + rng.normal(0, 0.1, n), 0.1 * r0, 1.2 * r1)
    s = np.abs(r - rp) + np.abs(rng.normal(0, 0.05, n))
    r[:ng] = knots
    rp[:ng] = knots
    s[:ng] = 0.0
    r[ng:2 * ng] = knots
    rp[ng:2 * ng] = np.roll(knots, 1)
    s[ng:2 * ng] = 0.3
    r[2 * ng:2 * ng + 4] = [r1, r1, 0.5 * r0, 1.5 * r1]
    rp[2 * ng:2 * ng + 4] = [r1, 0.999 * r1, 0.5 * r0, 1.5 * r1]
    s[2 * ng:2 * ng + 4] = [0.1, 0.0, 1e-5, 0.2]
    act = path.actions[0]
    for which in (0, 1):
        fast = act.CalcPairFast(which, r, rp, s)      # templated evaluation of the whole-path kernel
        path.ForceGeneral(False)
        glob = act.CalcPair(which, r, rp, s)          # run-time twin every other kernel uses (DavidFastGlobal)
        path.ForceGeneral(True)
        gen = act.CalcPair(which, r, rp, s)           # B-spline form (Cox-de Boor), the reference's own formulation
        path.ForceGeneral(False)
        ref = oracles[0].calc_pair(0, which, r, rp, s)
        scale = 1e-3 * np.max(np.abs(ref))
        assert rel_ok(fast, ref, scale=scale), (which, np.max(np.abs(fast - ref)))
        assert rel_ok(gen, ref, scale=scale), (which, np.max(np.abs(gen - ref)))
        assert rel_ok(fast, glob, scale=scale, rtol=1e-13), (which, np.max(np.abs(fast - glob)))   # same tables, same arithmetic
    path.close()


def test_fast_sqrt_within_one_ulp():
    cfg = CONFIGS["n2"]()
    path, _, _ = make_pair(cfg, 1)
    rng = np.random.default_rng(9)
    x = np.concatenate([[0.0, 1e-300, 1e-30, 1.0, 2.0, 4.0, 1e10], rng.uniform(0, 100, 50000), 10.0 ** rng.uniform(-20, 8, 50000)])
    got = path.FastSqrt(x)
    ref = np.sqrt(x)
    assert got[0] == 0.0
    ulp = np.spacing(ref)
    assert np.all(np.abs(got - ref) <= ulp), np.max(np.abs(got - ref) / ulp)
    path.close()


@pytest.mark.parametrize("david", [False, True])
@pytest.mark.parametrize("n_shards", [2, 4])
def test_slice_sharded_contexts_sum_to_the_whole_path(n_shards, david):
    """Slice-sharded contexts (one per GPU in production; here all on cuda:0): each shard's
    partial DActionDBeta / Potential / action, g(r) counts and S(k) sum to the unsharded values
    and match the oracle; halo pack/unpack moves the right slice."""
    import ctypes as C
    import torch
    from simpimc_b200 import host, sharded, capi
    from oracle import oracle as O
    cfg = S.plasma_config(Ne=6, Np=5, M=16, **({"pp_action": "DavidPairAction", "ep_action": "DavidPairAction"} if david else {}))
    n_clones = 2
    Rs = [np.stack([S.synthetic_paths(cfg, sp, c, 4242) for c in range(n_clones)]) for sp in range(2)]
    oracles = []
    for c in range(n_clones):
        o = O.Oracle(cfg)
        for sp in range(2):
            o.set_positions(sp, Rs[sp][c])
        oracles.append(o)
    shards = []
    for g in range(n_shards):
        sh = sharded.SliceSharding(cfg.n_bead, n_shards, g)
        p = host.Path(cfg, n_clones=n_clones, slice_lo=sh.lo, slice_hi=sh.hi)
        for sp in range(2):
            p.SetPositions(sp, sh.shard_positions(Rs[sp]))
        shards.append((sh, p))
    for ai in range(len(cfg.actions)):
        du = sum(p.actions[ai].DActionDBeta() for _, p in shards)
        v = sum(p.actions[ai].Potential() for _, p in shards)
        u = sum(p.actions[ai].TotalAction() for _, p in shards)
        parts = [(s, q) for s in range(2) for q in range(cfg.species[s].n_part)]
        for c, o in enumerate(oracles):
            assert rel_ok(du[c], o.dbeta(ai)) and rel_ok(v[c], o.potential(ai))
            assert rel_ok(u[c], o.get_action(ai, 0, 0, cfg.n_bead, parts, 0))
    counts = sum(host.PairCorrelation(p, 0, 1, 0.0, cfg.L / 2, 100).Counts() for _, p in shards)
    for c, o in enumerate(oracles):
        assert np.array_equal(counts[c], o.gofr(0, 1, 0.0, cfg.L / 2, 100)[1])
    # halo: move slice data around the ring by hand and check the action is unchanged
    for sp in range(2):
        N = cfg.species[sp].n_part
        bufs = []
        for sh, p in shards:
            t = torch.zeros((n_clones, N, 3), dtype=torch.float64, device="cuda")
            capi.check(p.L.pimc_halo_pack(p.h, sp, C.c_void_p(t.data_ptr())))
            p.Sync()
            assert np.array_equal(t.cpu().numpy(), Rs[sp][:, :, sh.lo, :])
            bufs.append(t)
        for g, (sh, p) in enumerate(shards):
            capi.check(p.L.pimc_halo_unpack(p.h, sp, C.c_void_p(bufs[sh.next_rank].data_ptr())))
            p.Sync()
            assert np.array_equal(p.GetPositions(sp), sh.shard_positions(Rs[sp]))
    for _, p in shards:
        p.close()


@pytest.mark.parametrize("kind", ["ueg", "plasma"])
def test_one_large_path_fills_the_gpu_by_splitting(kind):
    """One path with many particles and few slices (what a slice shard of BASELINE config C5
    looks like): K1 splits an item's partner loop by windows and K2 splits the particle loop
    (partial rho_k + fixed-order reduction) so the launch still fills the SMs.  Values against
    the oracle; two launches give identical bits (no atomics)."""
    from simpimc_b200 import host
    if kind == "ueg":
        cfg = S.ueg_config(N=300, M=8, n_xy=60, n_r_long=400)
    else:
        cfg = S.plasma_config(Ne=260, Np=131, M=8, pp_action="IlkkaPairAction")
    path, oracles, _ = make_pair(cfg, 1, seed=31)
    o = oracles[0]
    for sp in range(len(cfg.species)):
        got, ref = path.GetRhoK(sp, 0, host.OLD_MODE), o.rhok(sp, 0)
        assert np.max(np.abs(got - ref)) <= 1e-12 * cfg.species[sp].n_part
    for ai, act in enumerate(path.actions):
        du, du2 = act.DActionDBeta(), act.DActionDBeta()
        assert du[0] == du2[0]
        assert rel_ok(du[0], o.dbeta(ai))
        assert rel_ok(act.Potential()[0], o.potential(ai))
    path.close()
    o.close()


@pytest.mark.parametrize("name", ["ilkka_lr_n7", "ilkka_nolr_n8", "bare_lr_n7", "david_n7", "plasma"])
def test_action_gradient_and_laplacian(name):
    """Action::GetActionGradient / GetActionLaplacian (pair_action_class.h:305-366) against the
    oracle.  Ilkka gradients are analytic: 1e-10 relative to the sum of |link terms|.  The Bare /
    David gradients and every Laplacian are the reference's own central differences with
    eps = 1e-4, whose rounding noise is |U| * 2^-52 / eps (gradient) and / eps^2 (Laplacian): the
    bound is 1e-10 of the size of the differenced terms, |U_window| / eps resp. / eps^2."""
    cfg = CONFIGS[name]()
    C_ = 3
    path, oracles, _ = make_pair(cfg, C_, seed=77)
    M = cfg.n_bead
    rng = np.random.default_rng(5)
    eps = 1e-4
    for trial in range(4):
        sp = trial % len(cfg.species)
        parts = [(sp, rng.integers(0, cfg.species[sp].n_part, size=C_))]
        if trial == 3 and len(cfg.species) > 1:   # one particle of each species listed
            parts.append((1 - sp, rng.integers(0, cfg.species[1 - sp].n_part, size=C_)))
        b0 = rng.integers(0, M - 3, size=C_) if trial < 3 else np.full(C_, M - 2)
        n_w = 3 if trial < 3 else 2
        for ai, act in enumerate(path.actions):
            g = act.GetActionGradient(b0, b0 + n_w, parts, 0)
            lap = act.GetActionLaplacian(b0, b0 + n_w, parts, 0)
            for c, o in enumerate(oracles):
                pl = [(s, int(p[c])) for s, p in parts]
                g_ref = o.action_gradient(ai, 1, int(b0[c]), int(b0[c]) + n_w, pl, 0)
                l_ref = o.action_laplacian(ai, 1, int(b0[c]), int(b0[c]) + n_w, pl, 0)
                # size of the terms: |U| of the window's pairs, forward and backward links
                u_scale = 2.0 * abs(o.get_action(ai, 1, int(b0[c]), int(b0[c]) + n_w, pl, 0)) + 1e-3
                analytic = cfg.actions[ai].type == "IlkkaPairAction"
                g_tol = RTOL * max(np.max(np.abs(g_ref)), u_scale if analytic else u_scale / eps)
                assert np.max(np.abs(g[c] - g_ref)) <= g_tol, (name, trial, ai, c, g[c], g_ref)
                assert abs(lap[c] - l_ref) <= RTOL * max(abs(l_ref), u_scale / eps ** 2), (name, trial, ai, c, lap[c], l_ref)
    # level above max_level: zero, as the reference returns
    assert np.all(path.actions[0].GetActionGradient(0, 2, [(0, 0)], 1) == 0.0)
    assert np.all(path.actions[0].GetActionLaplacian(0, 2, [(0, 0)], 1) == 0.0)
    path.close()
    for o in oracles:
        o.close()


@pytest.mark.parametrize("kind,n_r", [("ueg", 100), ("ueg", 3000), ("plasma", 100), ("plasma", 700)])
def test_tiled_gofr_bins_match_oracle_and_general_kernel(kind, n_r):
    """K5 v2 (gofr_tiled_kernel): several particle tiles (N > 128), a partial last 32-slice chunk,
    both species layouts, per-warp and per-CTA histograms (large n_r): counts bit-identical to the
    oracle and to the thread-per-pair kernel."""
    from simpimc_b200 import host
    if kind == "ueg":
        cfg, pairs = S.ueg_config(N=150, M=40, n_xy=60, n_r_long=400), [(0, 0)]
    else:
        cfg, pairs = S.plasma_config(Ne=140, Np=37, M=40), [(0, 1), (1, 1), (0, 0)]
    path, oracles, _ = make_pair(cfg, 2, seed=9)
    for sa, sb in pairs:
        got = host.PairCorrelation(path, sa, sb, 0.0, cfg.L / 2.0, n_r).Counts()
        path.ForceGeneral(True)
        gen = host.PairCorrelation(path, sa, sb, 0.0, cfg.L / 2.0, n_r).Counts()
        path.ForceGeneral(False)
        assert np.array_equal(got, gen)
        for c, o in enumerate(oracles):
            assert np.array_equal(got[c], o.gofr(sa, sb, 0.0, cfg.L / 2.0, n_r)[1]), (kind, sa, sb, c)
        n_pairs = cfg.species[sa].n_part * (cfg.species[sa].n_part - 1) // 2 if sa == sb else cfg.species[sa].n_part * cfg.species[sb].n_part
        assert got.sum(axis=1).max() <= n_pairs * cfg.n_bead
    path.close()
    for o in oracles:
        o.close()


@pytest.mark.parametrize("name", ["ilkka_lr_n7", "ilkka_nolr_n8", "bare_lr_n7", "david_n7", "plasma"])
def test_get_action_with_several_listed_particles(name):
    """Action::GetAction with the particle lists permutation moves pass
    (perm_bisect_iterative_class.h:200-204: the 2-4 particles of a cycle, same species) and lists
    that touch both species of an action: GenerateParticlePairs' (listed, other) + (listed,
    listed) pairs, several pending proposals per species, rho_k refreshed for all of them,
    commit / rollback.  Checked against the oracle: OLD and NEW actions, the Metropolis
    difference, positions and rho_k after the commit."""
    from simpimc_b200 import host
    cfg = CONFIGS[name]()
    C_ = 3
    path, oracles, _ = make_pair(cfg, C_, seed=21)
    M = cfg.n_bead
    rng = np.random.default_rng(17)
    ns = len(cfg.species)
    for trial in range(6):
        nb = [4, 2, 4, M, 4, 2][trial]
        b0 = rng.integers(0, M, size=C_) if nb < M else np.zeros(C_, dtype=np.int64)
        first, n_beads = ((b0 + 1) % M, nb - 1) if nb < M else (b0, M)
        # the list: 2-3 particles of species 0, on odd trials also one of the other species
        k0 = 2 + trial % 2
        plist = []
        chosen = np.stack([rng.permutation(cfg.species[0].n_part)[:k0] for _ in range(C_)])   # [clone][k0], distinct per clone
        for i in range(k0):
            plist.append((0, chosen[:, i].astype(np.int32)))
        if ns > 1 and trial % 2 == 1:
            plist.append((1, rng.integers(0, cfg.species[1].n_part, size=C_).astype(np.int32)))
        news = []
        for sp, parts in plist:
            cur = path.GetBeads(sp, parts, first, n_beads)
            new = cur + 0.07 * rng.standard_normal(cur.shape)
            path.Propose(sp, parts, first, new)
            news.append(new)
            for c, o in enumerate(oracles):
                o.propose(sp, int(parts[c]), int(first[c]), new[c])
        accept = rng.integers(0, 2, size=C_)
        for ai, act in enumerate(path.actions):
            path.SetMode(host.OLD_MODE)
            old = act.GetAction(b0, b0 + nb, plist, 0)
            path.SetMode(host.NEW_MODE)
            new = act.GetAction(b0, b0 + nb, plist, 0)
            for c, o in enumerate(oracles):
                pl = [(sp, int(parts[c])) for sp, parts in plist]
                ro = o.get_action(ai, 0, int(b0[c]), int(b0[c]) + nb, pl, 0)
                rn = o.get_action(ai, 1, int(b0[c]), int(b0[c]) + nb, pl, 0)
                assert rel_ok(old[c], ro) and rel_ok(new[c], rn), (name, trial, ai, c, old[c], ro, new[c], rn)
                assert abs((new[c] - old[c]) - (rn - ro)) <= RTOL * max(abs(rn - ro), 1e-4 * (abs(rn) + abs(ro))), (name, trial, ai, c)
        path.Commit(accept)
        for c, o in enumerate(oracles):
            for sp, parts in plist:
                o.finish_move(sp, int(parts[c]), int(b0[c]), int(b0[c]) + nb, bool(accept[c]))
        for sp in range(ns):
            got = path.GetPositions(sp)
            for c, o in enumerate(oracles):
                assert np.array_equal(got[c], o.get_positions(sp, 0)), (name, trial, "positions after commit", sp, c)
                if path._n_k():
                    assert np.max(np.abs(path.GetRhoK(sp, c, host.OLD_MODE) - o.rhok(sp, 0))) <= 1e-11 * cfg.species[sp].n_part
    # seventeen listed particles of one species (more than the proposal slots): refused loudly
    with pytest.raises(RuntimeError):
        path.actions[0].GetAction(0, 2, [(0, i) for i in range(17)], 0)
    path.close()
    for o in oracles:
        o.close()


@pytest.mark.parametrize("relative", [False, True])
def test_perm_table_matches_oracle(relative):
    """pimc_perm_table: PermBisectIterative::UpdatePermTable / the PermBisectTable variant for every
    clone at once; windows that wrap past n_bead; the epsilon cut is applied (exact zeros)."""
    cfg = S.ueg_config(N=33, M=16)
    C_ = 3
    path, oracles, _ = make_pair(cfg, C_, seed=4)
    b0 = np.array([2, 13, 15], dtype=np.int32)
    for eps in (1e-100, 1e-3):
        t = path.PermTable(0, b0, 4, epsilon=eps, relative=relative)
        for c, o in enumerate(oracles):
            ref = o.perm_table(0, int(b0[c]), 4, epsilon=eps, relative=relative)
            assert np.array_equal(t[c] == 0.0, ref == 0.0) or np.max(np.abs(t[c] - ref)) <= 1e-12   # a value within rounding of the cut may fall on either side
            assert np.max(np.abs(t[c] - ref)) <= 1e-12 * max(1.0, np.max(ref))
    assert np.any(path.PermTable(0, b0, 4, epsilon=1e-3) == 0.0)
    path.close()
    for o in oracles:
        o.close()


@pytest.mark.parametrize("name", ["ilkka_lr_n33", "bare_lr_n7", "david_lr_n7", "plasma"])
def test_kspace_growth_after_actions_exist(name):
    """KSpace::Setup is grow-only (k_space_class.h:34-41): a StructureFactor with a larger k_cut
    (structure_factor_class.h:49-51) rebuilds the k-vector list after the long-range actions matched
    their weights to the old one.  The actions' values, window differences and device sweeps must
    not change: the weights are re-matched to the new list (vectors beyond the table's shells get 0)."""
    from simpimc_b200 import host
    cfg = CONFIGS[name]()
    C = 2
    path, oracles, Rs = make_pair(cfg, C)
    n_k0 = path._n_k()
    before = [(a.DActionDBeta(), a.TotalAction(), a.Potential() if a.type != "DavidPairAction" else None) for a in path.actions]
    sk = host.StructureFactor(path, 0, 0, 1.6 * cfg.k_cut)
    assert path._n_k() > n_k0
    sk.Accumulate()
    for ai, a in enumerate(path.actions):
        du, u = a.DActionDBeta(), a.TotalAction()
        assert rel_ok(du, before[ai][0], rtol=1e-12) and rel_ok(u, before[ai][1], rtol=1e-12), (name, ai)
        if before[ai][2] is not None:
            assert rel_ok(a.Potential(), before[ai][2], rtol=1e-12)
        for c, o in enumerate(oracles):
            assert rel_ok(du[c], o.dbeta(ai)), (name, ai, c)
    # a move window in OLD / NEW mode against the oracle (whose k set did not grow)
    rng = np.random.default_rng(5)
    N, M, nb = cfg.species[0].n_part, cfg.n_bead, 4
    part, b0 = rng.integers(0, N, C), rng.integers(0, M, C)
    first = (b0 + 1) % M
    newR = np.stack([Rs[0][c][part[c], (first[c] + np.arange(nb - 1)) % M] for c in range(C)]) + 0.05 * rng.standard_normal((C, nb - 1, 3))
    path.Propose(0, part, first, newR)
    for c, o in enumerate(oracles):
        o.propose(0, int(part[c]), int(first[c]), newR[c])
    for ai, a in enumerate(path.actions):
        if 0 not in (a.species_a, a.species_b):
            continue
        path.SetMode(host.OLD_MODE)
        old = a.GetAction(b0, b0 + nb, [(0, part)], 0)
        path.SetMode(host.NEW_MODE)
        new = a.GetAction(b0, b0 + nb, [(0, part)], 0)
        for c, o in enumerate(oracles):
            ro = o.get_action(ai, 0, int(b0[c]), int(b0[c]) + nb, [(0, int(part[c]))], 0)
            rn = o.get_action(ai, 1, int(b0[c]), int(b0[c]) + nb, [(0, int(part[c]))], 0)
            assert rel_ok(old[c], ro) and rel_ok(new[c], rn), (name, ai, c)
    path.Commit(1)
    for c, o in enumerate(oracles):
        o.finish_move(0, int(part[c]), int(b0[c]), int(b0[c]) + nb, True)
    for ai, a in enumerate(path.actions):
        du = a.DActionDBeta()
        for c, o in enumerate(oracles):
            assert rel_ok(du[c], o.dbeta(ai)), (name, "after commit", ai, c)
    n_acc = path.BisectSweep(0, 2, 6, seed=3)       # device sweep with the grown set: rho_k stays consistent
    R_after = path.GetPositions(0)
    fresh = host.Path(cfg, n_clones=C)
    for sp in range(len(cfg.species)):
        fresh.SetPositions(sp, R_after if sp == 0 else Rs[sp])
    for ai, a in enumerate(path.actions):
        assert rel_ok(a.DActionDBeta(), fresh.actions[ai].DActionDBeta(), rtol=1e-11), (name, "after sweep", ai)
    assert n_acc.shape == (C,)
    fresh.close()
    for o in oracles:
        o.close()
    path.close()
