"""Permuting bisection on the device (pimc_perm_bisect_sweep, csrc/perm.cuh):
PermBisectIterative::Attempt / Accept / Reject (perm_bisect_iterative_class.h:10-222, perm_bisect_class.h:32-82).

(1) Exact stream parity: the host mirror simpimc_b200.perm_moves.perm_bisect_attempt -- pinned to the REFERENCE's own
    PermBisectIterative on injected random numbers in tests/test_stream_ref_cpu.py -- draws the same Philox numbers and
    takes its pair-action values from the CPU oracle; after every attempt the device walkers show the same cycle
    (window, members, selection steps: integer state, bit-exact), the same accept flag, the same permutation at the
    beta seam and the same positions by label (1e-12).
(2) Whole-path evaluations on a permuted path: Kinetic follows the links over the seam; the entry points that read one
    particle's path by label refuse loudly; PathDump carries the permutation through a restart.
(3) Statistical parity of a sampled run with the reference program (oracle/_ref: its PermBisectIterative, Kinetic,
    IlkkaPairAction, std::mt19937): energies within combined error bars, the same distribution of attempted and
    accepted cycle lengths.
"""
import copy

import numpy as np
import pytest

from simpimc_b200 import system as S

pytestmark = pytest.mark.gpu


def _with_kinetic(pair_cfg, n_images):
    cfg = copy.copy(pair_cfg)
    cfg.actions = [S.ActionConfig("Kinetic%s" % s.name, "Kinetic", s.name, n_images=n_images) for s in pair_cfg.species if s.lam > 0] + list(pair_cfg.actions)
    return cfg


@pytest.mark.parametrize("name,n_images_kin,n_level", [("egas", 0, 2), ("egas", 1, 3), ("plasma", 0, 2)])
def test_device_perm_sweep_follows_the_host_mirror_of_its_stream(name, n_images_kin, n_level):
    from simpimc_b200 import host, perm_moves as PM
    from oracle import oracle as O
    if name == "egas":
        pair_cfg = S.egas_config(N=5, M=16, n_xy=40, n_r_long=200)        # theta = 0.1: exchange is frequent
    else:
        pair_cfg = S.plasma_config(Ne=5, Np=4, M=8, theta=0.25, pp_action="IlkkaPairAction")   # e-p: the moved species is species b of an action
    cfg = _with_kinetic(pair_cfg, n_images_kin)
    n_sp = len(cfg.species)
    M = cfg.n_bead
    C = 3
    path = host.Path(cfg, n_clones=C)
    oracles = [O.Oracle(pair_cfg) for _ in range(C)]
    R = [np.stack([S.synthetic_paths(cfg, sp, c, 5) for c in range(C)]) for sp in range(n_sp)]
    nxt = [np.tile(np.arange(cfg.species[sp].n_part, dtype=np.int32), (C, 1)) for sp in range(n_sp)]
    for sp in range(n_sp):
        path.SetPositions(sp, R[sp])
        for c in range(C):
            oracles[c].set_positions(sp, R[sp][c])
    seed = 0x9E3700000000C1C1
    n_att = 150
    n_acc_host = np.zeros(C, dtype=np.int64)
    n_acc_dev = np.zeros(C, dtype=np.int64)
    att_dev = np.zeros((C, 8), dtype=np.int64)
    acc_dev = np.zeros((C, 8), dtype=np.int64)
    att_host = np.zeros((C, 8), dtype=np.int64)
    acc_host = np.zeros((C, 8), dtype=np.int64)
    wrapped_on_permuted = 0
    for attempt in range(n_att):
        sp = attempt % n_sp
        N = cfg.species[sp].n_part
        acts = [ai for ai, a in enumerate(pair_cfg.actions) if cfg.species[sp].name in (a.species_a, a.species_b)]
        results = []
        for c in range(C):
            o = oracles[c]

            def action_old_new(labels, bead0, nb, windows, o=o):
                for l in labels:
                    o.propose(sp, l, bead0, windows[l])
                parts = [(sp, l) for l in labels]
                old = sum(o.get_action(ai, 0, bead0, bead0 + nb, parts, 0) for ai in acts)
                new = sum(o.get_action(ai, 1, bead0, bead0 + nb, parts, 0) for ai in acts)
                for l in labels:
                    o.finish_move(sp, l, bead0, bead0 + nb, False)
                return old, new

            was_permuted = not np.array_equal(nxt[sp][c], np.arange(N))
            res = PM.perm_bisect_attempt(pair_cfg, sp, n_level, seed, attempt, c, R[sp][c], nxt[sp][c], action_old_new,
                                         n_images_kin=n_images_kin)
            results.append(res)
            if res["n_perm"] > 0:
                att_host[c, res["n_perm"] - 1] += 1
                if res["accept"]:
                    acc_host[c, res["n_perm"] - 1] += 1
                    n_acc_host[c] += 1
                    o.set_positions(sp, R[sp][c])
                    if res["bead0"] + (1 << n_level) > M - 1 and was_permuted:
                        wrapped_on_permuted += 1
        a, t, k = path.PermBisectSweep(sp, n_level, 1, seed, attempt0=attempt)
        n_acc_dev += a
        att_dev += t
        acc_dev += k
        last = path.PermLastCycle()
        for c, res in enumerate(results):
            what = (name, attempt, c, res, {q: v[c].tolist() for q, v in last.items()})
            assert last["b0"][c] == res["bead0"], what
            assert last["n_perm"][c] == res["n_perm"], what
            assert last["n_steps"][c] == res["steps"], what
            assert last["particles"][c][:max(res["n_perm"], 0)].tolist() == list(res["particles"]), what
            assert bool(last["accept"][c]) == bool(res["accept"]), what
        assert np.array_equal(n_acc_dev, n_acc_host) and np.array_equal(att_dev, att_host) and np.array_equal(acc_dev, acc_host), (name, attempt)
        assert np.array_equal(path.GetPermutation(sp), nxt[sp]), (name, attempt, path.GetPermutation(sp), nxt[sp])
        got = path.GetPositions(sp)
        assert np.max(np.abs(got - R[sp])) <= 1e-12 * max(1.0, np.max(np.abs(R[sp]))), (name, attempt)
    if name == "egas":
        assert acc_dev[:, 1:].sum() >= 3, acc_dev        # real permutations (two or more particles) were accepted
        assert wrapped_on_permuted >= 1                  # a window rolled over the seam of an already permuted path
    assert 0 < n_acc_dev.sum() < C * n_att
    # rho_k was carried along incrementally (a relabelling does not change it): compare with a rebuild by the oracle
    for sp in range(n_sp):
        for c in range(C):
            assert np.max(np.abs(path.GetRhoK(sp, c, host.OLD_MODE) - oracles[c].rhok(sp, 0))) <= 1e-10 * cfg.species[sp].n_part
    # pair actions pair labels at equal slices: unchanged code path, still the oracle's numbers on the final configuration
    n_kin = len(cfg.actions) - len(pair_cfg.actions)
    for ai in range(len(pair_cfg.actions)):
        du = path.actions[n_kin + ai].DActionDBeta()
        for c in range(C):
            ref = oracles[c].dbeta(ai)
            assert abs(du[c] - ref) <= 1e-10 * abs(ref)
    path.close()
    for o in oracles:
        o.close()


def _kinetic_dbeta_along_links(cfg, sp, R, nxt, n_images):
    """Kinetic::DActionDBeta (kinetic_class.h:35-45) of one walker with the links followed over the beta seam."""
    from simpimc_b200.free_spline import FreeSpline
    s = cfg.species[sp]
    N, M = s.n_part, cfg.n_bead
    fs = FreeSpline(cfg.L if cfg.pbc else 0.0, n_images, s.lam, cfg.tau, use_tau_derivative=True)
    tot = 0.0
    for p in range(N):
        for b in range(M):
            r1 = R[p, b + 1] if b + 1 < M else R[nxt[p], 0]
            d = R[p, b] - r1
            if cfg.pbc:
                d = d - np.rint(d / cfg.L) * cfg.L
            tot += float(fs.GetDLogRhoFreeDTau(d))
    return N * M * cfg.n_d / (2.0 * cfg.tau) + tot


@pytest.mark.parametrize("n_images", [0, 1])
def test_permuted_paths_kinetic_follows_the_links_and_label_moves_refuse(n_images, tmp_path):
    from simpimc_b200 import host
    pair_cfg = S.egas_config(N=5, M=16, n_xy=40, n_r_long=200)
    cfg = _with_kinetic(pair_cfg, n_images)
    C = 4
    path = host.Path(cfg, n_clones=C)
    R = np.stack([S.synthetic_paths(cfg, 0, c, 11) for c in range(C)])
    path.SetPositions(0, R)
    kin = path.actions[0]
    du_identity = kin.DActionDBeta()
    for c in range(C):
        ref = _kinetic_dbeta_along_links(cfg, 0, R[c], np.arange(5), n_images)
        assert abs(du_identity[c] - ref) <= 1e-10 * abs(ref), (c, du_identity[c], ref)
    nxt = np.array([[1, 2, 0, 3, 4], [0, 1, 2, 3, 4], [4, 3, 2, 1, 0], [1, 0, 3, 4, 2]], dtype=np.int32)
    path.SetPermutation(0, nxt)
    assert np.array_equal(path.GetPermutation(0), nxt)
    du = kin.DActionDBeta()
    for c in range(C):
        ref = _kinetic_dbeta_along_links(cfg, 0, R[c], nxt[c], n_images)
        assert abs(du[c] - ref) <= 1e-10 * abs(ref), (c, du[c], ref)
    assert du[1] == du_identity[1] and du[0] != du_identity[0]
    # pair actions, g(r), S(k) do not depend on the links
    pair = path.actions[1]
    before = pair.DActionDBeta()
    # moves that read one particle's path by label across the seam refuse while the path is permuted
    with pytest.raises(RuntimeError):
        path.BisectSweep(0, 2, 1, 5)
    with pytest.raises(RuntimeError):
        path.DisplaceSweep(0, 0.1, 1, 5)
    with pytest.raises(RuntimeError):
        kin.GetAction(np.zeros(C, dtype=np.int32), 4, [(0, np.zeros(C, dtype=np.int32))], 0)
    with pytest.raises(RuntimeError):
        path.SetPermutation(0, np.array([0, 0, 1, 2, 3], dtype=np.int32))      # not a permutation
    assert np.array_equal(pair.DActionDBeta(), before)
    # PathDump writes the reference's permutation table (previous of bead 0, next of the last bead) and a restart restores it
    dump = host.PathDump(path)
    dump.Write()
    fn = str(tmp_path / "dump.npz")
    dump.Save(fn)
    perm = np.load(fn)["Observables/path_dump/e/permutation"][-1]
    for c in range(C):
        assert np.array_equal(perm[c, :, 1], nxt[c])
        assert np.array_equal(perm[c, nxt[c], 0], np.arange(5))
    path2 = host.Path(cfg, n_clones=C)
    host.PathDump.Restart(path2, fn)
    assert np.array_equal(path2.GetPermutation(0), nxt)
    assert np.array_equal(path2.GetPositions(0), R)
    assert np.array_equal(path2.actions[0].DActionDBeta(), du)
    path2.close()
    # back on the identity every move is available again
    path.SetPermutation(0, np.arange(5, dtype=np.int32))
    assert np.array_equal(kin.DActionDBeta(), du_identity)
    path.BisectSweep(0, 2, 1, 5)
    path.close()


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


def test_sampled_permuting_run_matches_the_reference_program():
    """Bosonic sampling (every permutation sector) of N = 5 particles at theta = 0.1 by PermBisectIterative alone:
    reference program (one walker, blocked series) vs 256 device walkers."""
    from oracle import refsim
    if not refsim.available():
        pytest.skip("oracle/_ref not built")
    from simpimc_b200 import host
    N, M, n_level = 5, 16, 2
    pair_cfg = S.egas_config(N=N, M=M, n_xy=40, n_r_long=200)
    cfg = _with_kinetic(pair_cfg, 1)
    cfg.moves = [{"name": "PermE", "type": "PermBisectIterative", "species": "e", "n_level": n_level, "n_images": 0}]
    cfg.observables = []
    sim = refsim.RefSim(cfg, seed=23, fast=refsim.available(fast=True))
    if not hasattr(sim.lib, "ref_permutation"):
        pytest.skip("oracle/_ref predates the permutation hooks")
    sim.set_positions(0, S.synthetic_paths(cfg, 0, 0, 5))
    per_sweep = N * M // (1 << n_level)
    sim.move_do(0, 150 * per_sweep)
    att0, acc0 = sim.perm_counts(0)
    e_kin, e_pair, permuted = [], [], []
    n_ref_sweeps = 10000         # ~40 us per attempt of the reference program on the GPU box's host
    for i in range(n_ref_sweeps):
        sim.move_do(0, per_sweep)
        e_kin.append(sim.dbeta(0) / M)
        e_pair.append(sim.dbeta(1) / M)
        permuted.append(float(np.sum(sim.permutation(0)[1] == np.arange(N))))      # particles whose path closes on itself
    att1, acc1 = sim.perm_counts(0)
    n_moves_ref, n_acc_ref = sim.move_counts(0)
    sim.close()
    att_ref, acc_ref = (att1 - att0)[:N].astype(np.float64), (acc1 - acc0)[:N].astype(np.float64)
    n_ref = n_ref_sweeps * per_sweep
    # ---- device
    C = 256
    path = host.Path(cfg, n_clones=C)
    path.SetPositions(0, np.stack([S.synthetic_paths(cfg, 0, c, 5) for c in range(C)]))
    kin, pair = path.actions
    att = 300 * per_sweep
    path.PermBisectSweep(0, n_level, att, 4242, attempt0=0)
    dk, dp, dperm = [], [], []
    att_dev = np.zeros((C, 8), dtype=np.int64)
    acc_dev = np.zeros((C, 8), dtype=np.int64)
    n_dev = 0
    for _ in range(80):
        _, t, k = path.PermBisectSweep(0, n_level, 3 * per_sweep, 4242, attempt0=att)
        att += 3 * per_sweep
        n_dev += 3 * per_sweep * C
        att_dev += t
        acc_dev += k
        dk.append(kin.DActionDBeta() / M)
        dp.append(pair.DActionDBeta() / M)
        dperm.append(np.sum(path.GetPermutation(0) == np.arange(N), axis=1).astype(np.float64))
    path.close()
    # cycle statistics: attempted and accepted cycles per move, by length (binomial error bars, 5 sigma + a floor)
    for what, ref, dev in (("attempted", att_ref, att_dev.sum(axis=0)[:N]), ("accepted", acc_ref, acc_dev.sum(axis=0)[:N])):
        f_ref, f_dev = ref / n_ref, dev / n_dev
        # successive moves of one walker are correlated: allow a generous autocorrelation factor on the reference leg
        sig = np.sqrt(10.0 * f_ref * (1 - f_ref) / n_ref + f_dev * (1 - f_dev) / n_dev) + 1e-4
        assert np.all(np.abs(f_ref - f_dev) <= 5.0 * sig), (what, f_ref, f_dev, sig)
    assert acc_dev.sum(axis=0)[1] > 0 and acc_ref[1] > 0          # exchanges happen in both runs
    # energies and the weight of the permuted sectors
    e_kin, e_pair, permuted = np.array(e_kin), np.array(e_pair), np.array(permuted)
    for what, dev, ref in (("kinetic", np.mean(dk, axis=0), e_kin), ("pair", np.mean(dp, axis=0), e_pair),
                           ("total", np.mean(dk, axis=0) + np.mean(dp, axis=0), e_kin + e_pair),
                           ("particles on closed single-particle paths", np.mean(dperm, axis=0), permuted)):
        d_mean, d_err = dev.mean(), dev.std(ddof=1) / np.sqrt(C)
        r_mean, r_err = _mean_err(ref)
        assert abs(d_mean - r_mean) <= 4.0 * np.hypot(d_err, r_err), (what, d_mean, d_err, r_mean, r_err)
        if what == "pair":       # teeth: the error bar of the large component is below 5 % of it
            assert np.hypot(d_err, r_err) < 0.05 * abs(r_mean), (what, d_err, r_err, r_mean)
    assert 0.05 < np.mean(dperm) < N - 0.05, np.mean(dperm)      # both kinds of path occur
