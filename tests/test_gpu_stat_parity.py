"""Statistical parity of SAMPLED runs at the shape of the reference's own e-gas input
(inputs/e-gas/e-gas.xml: polarized electrons at r_s = 1, theta = 0.1, beta = 3.42075, Kinetic with
n_images = 100, the bisection's sampling splines with n_images = 1, IlkkaPairAction with long
range), N = 7 (L = 3.08363, the shipped size) and N = 33 (BASELINE config C1), M = 64 slices.

Reference: the reference program itself (oracle/_ref: its Bisect + Kinetic + IlkkaPairAction +
PairCorrelation + StructureFactor, std::mt19937), one walker, blocked series.  Device: 256
independent walkers of pimc_bisect_sweep (Philox), Kinetic and pair DActionDBeta, g(r) and S(k)
through the C ABI.  north_star: energy, g(r) and S(k) within statistical error bars -- here
|difference| <= 4 sigma (combined) with error bars below 5 % of the signal where it is large."""
import copy

import numpy as np
import pytest

from simpimc_b200 import system as S

pytestmark = pytest.mark.gpu


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


@pytest.mark.parametrize("N,ref_sweeps,dev_meas", [(7, 10000, 60), (33, 800, 80)])
def test_sampled_egas_run_matches_the_reference_program(N, ref_sweeps, dev_meas):
    from oracle import refsim
    if not refsim.available():
        pytest.skip("oracle/_ref not built")
    from simpimc_b200 import host
    M, n_level, n_r = 64, 3, 20
    pair_cfg = S.egas_config(N=N, M=M, n_xy=60, n_r_long=400)
    cfg = copy.copy(pair_cfg)
    cfg.actions = [S.ActionConfig("Kinetic", "Kinetic", "e", n_images=100)] + list(pair_cfg.actions)
    cfg.moves = [{"name": "BisectE", "type": "Bisect", "species": "e", "n_level": n_level, "n_images": 1}]
    cfg.observables = [{"name": "gr", "type": "PairCorrelation", "species_a": "e", "species_b": "e", "r_min": 0.0, "r_max": cfg.L / 2.0, "n_r": n_r},
                       {"name": "sk", "type": "StructureFactor", "species_a": "e", "species_b": "e", "k_cut": cfg.k_cut}]
    attempts_per_sweep = N * M // (1 << n_level)
    # ---- reference: one walker
    sim = refsim.RefSim(cfg, seed=17, fast=refsim.available(fast=True))
    sim.set_positions(0, S.synthetic_paths(cfg, 0, 0, 5))
    sim.move_do(0, 300 * attempts_per_sweep)
    e_kin, e_pair = [], []
    per_block = 50
    g_blocks, s_blocks = [], []
    g_prev, s_prev = sim.gofr_counts(0, n_r), sim.sofk_sums(1)
    for i in range(ref_sweeps):
        sim.move_do(0, attempts_per_sweep)
        e_kin.append(sim.dbeta(0) / M)
        e_pair.append(sim.dbeta(1) / M)
        sim.observable_accumulate(0)
        sim.observable_accumulate(1)
        if (i + 1) % per_block == 0:
            g_now, s_now = sim.gofr_counts(0, n_r), sim.sofk_sums(1)
            g_blocks.append((g_now - g_prev) / per_block)
            s_blocks.append((s_now - s_prev) / per_block)
            g_prev, s_prev = g_now, s_now
    acc_ref = sim.move_counts(0)
    sim.close()
    e_kin, e_pair = np.array(e_kin), np.array(e_pair)
    # ---- device: 256 walkers
    C = 256
    path = host.Path(cfg, n_clones=C)
    path.SetMoveImages(0, 1)
    path.SetPositions(0, np.stack([S.synthetic_paths(cfg, 0, c, 5) for c in range(C)]))
    kin, pair = path.actions
    gr = host.PairCorrelation(path, 0, 0, 0.0, cfg.L / 2.0, n_r)
    sk = host.StructureFactor(path, 0, 0, cfg.k_cut)
    att = 150 * attempts_per_sweep
    n_acc = path.BisectSweep(0, n_level, att, 1717, attempt0=0)
    dk, dp = [], []
    for _ in range(dev_meas):
        n_acc = n_acc + path.BisectSweep(0, n_level, 3 * attempts_per_sweep, 1717, attempt0=att)
        att += 3 * attempts_per_sweep
        dk.append(kin.DActionDBeta() / M)
        dp.append(pair.DActionDBeta() / M)
        gr.Accumulate()
        sk.Accumulate()
    path.close()
    # acceptance ratios agree (a coarse check that the same move is being made)
    r_ref, r_dev = acc_ref[1] / acc_ref[0], n_acc.sum() / (C * att)
    assert abs(r_ref - r_dev) < 0.03, (r_ref, r_dev)
    n_sig = 4.0
    # ---- energies: kinetic, pair and total thermal estimator per clone vs the blocked reference series
    for name, dev, ref in (("kinetic", np.mean(dk, axis=0), e_kin), ("pair", np.mean(dp, axis=0), e_pair),
                           ("total", np.mean(dk, axis=0) + np.mean(dp, axis=0), e_kin + e_pair)):
        d_mean, d_err = dev.mean(), dev.std(ddof=1) / np.sqrt(C)
        r_mean, r_err = _mean_err(ref)
        assert abs(d_mean - r_mean) <= n_sig * np.hypot(d_err, r_err), (N, name, d_mean, d_err, r_mean, r_err)
        # teeth: error bars below 5 % of the signal -- the pair energy, the large component (the thermal kinetic
        # estimator N M n_d / 2 tau - sum r^2 / 4 lambda tau^2 averages to ~0 here with a variance of ~50 per sample)
        assert np.hypot(d_err, r_err) < 0.05 * abs(np.mean(e_pair)), (N, name, d_err, r_err, np.mean(e_pair))
    # ---- g(r) and S(k)
    g_dev, s_dev = gr.y / dev_meas, sk.sk / dev_meas
    g_ref, s_ref = np.array(g_blocks), np.array(s_blocks)
    for name, dev, ref in (("g(r)", g_dev, g_ref), ("S(k)", s_dev, s_ref)):
        d_mean, d_err = dev.mean(axis=0), dev.std(axis=0, ddof=1) / np.sqrt(C)
        r_mean, r_err = ref.mean(axis=0), ref.std(axis=0, ddof=1) / np.sqrt(len(ref))
        scale = np.max(np.abs(r_mean))
        sig = np.hypot(d_err, r_err) + 1e-12 * scale
        assert np.all(np.abs(d_mean - r_mean) <= 4.5 * sig), (N, name, np.max(np.abs(d_mean - r_mean) / sig))
        # teeth: where the signal is large the error bars are below 5 % of it for the bulk of the bins / k vectors (median;
        # the few slowest modes of S(k) decorrelate over more sweeps than the reference leg can afford here) and below 10 % everywhere
        big = np.abs(r_mean) > 0.5 * scale
        rel = sig[big] / np.abs(r_mean[big])
        assert np.median(rel) < 0.05 and np.max(rel) < 0.10, (N, name, np.median(rel), np.max(rel))
