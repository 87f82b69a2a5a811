"""TEST INFRASTRUCTURE ONLY -- writes tests/golden/*.npz from the REFERENCE ITSELF.

Runs in the build container only (needs /root/reference compiled in place into
oracle/_ref/libsimpimc_ref.so by `make -C oracle ref`): the reference's own Path, Species,
KSpace, Ilkka/Bare/DavidPairAction, PairCorrelation and StructureFactor classes are driven
through oracle/ref_driver.cc on seeded synthetic inputs, and every number they return is stored
next to the inputs' recipe.  The fixtures travel to the GPU box; the reference does not.

    python oracle/make_golden.py            # rewrites tests/golden/*.npz

What a fixture holds (all float64 unless noted), for one named configuration:
    kspace_index int32 [n_k][3], kspace_mag [n_k]          KSpace::Setup order (bit-exact)
    rhok_<s> complex [M][n_k]                                Species::GetRhoK after InitRhoK
    dbeta [A], potential [A], total [A]                      DActionDBeta / Potential / whole-path GetAction
    pair_r, pair_rp, pair_s [n], pair_<a>_<which> [n]        CalcU / CalcdUdBeta / CalcV samples
    win_* ...                                                move windows: particle, b0, n links, proposed
                                                             positions, OLD and NEW GetAction per action,
                                                             accept flag, rho_k and positions after commit
    gofr_<sa><sb> [n_r] (counts), gofr_bins uint32           PairCorrelation::Accumulate / ReverseMap
    sofk_<sa><sb> [n_k]                                      StructureFactor::Accumulate
    grad_meta int64 [n][4] (species, particle, b0, n links), grad_val [n][A][3], lap_val [n][A]
                                                             GetActionGradient / GetActionLaplacian
The inputs are regenerated in the tests from (config name, seed) by simpimc_b200.system.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from simpimc_b200 import system as S  # noqa: E402
from oracle import refsim  # noqa: E402

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")
SEED = 20261017

CONFIGS = {
    "ilkka_lr_n7": lambda: S.ueg_config(N=7, M=8),
    "ilkka_lr_n33": lambda: S.ueg_config(N=33, M=16),
    "ilkka_nolr_n8": lambda: S.ueg_config(N=8, M=8, use_long_range=False),
    "bare_lr_n7": lambda: S.ueg_config(N=7, M=8, action="BarePairAction"),
    "david_n7": lambda: S.ueg_config(N=7, M=8, action="DavidPairAction", use_long_range=False),
    "plasma": lambda: S.plasma_config(Ne=6, Np=5, M=8),
    # off-diagonal table with n_x != n_y, different x and y grids and no x <-> y symmetry (a transposition would show)
    "ilkka_asym": lambda: S.ueg_config(N=9, M=8, n_xy=80, xy_asym=(64, 60.0, 0.3)),
}


def observables_for(cfg, n_r=100):
    obs = []
    ns = len(cfg.species)
    for sa in range(ns):
        for sb in range(sa, ns):
            obs.append({"name": "gr_%d%d" % (sa, sb), "type": "PairCorrelation", "species_a": cfg.species[sa].name,
                        "species_b": cfg.species[sb].name, "r_min": 0.0, "r_max": cfg.L / 2.0, "n_r": n_r})
            if cfg.k_cut is not None:
                obs.append({"name": "sk_%d%d" % (sa, sb), "type": "StructureFactor", "species_a": cfg.species[sa].name,
                            "species_b": cfg.species[sb].name, "k_cut": cfg.k_cut})
    return obs


def window_cases(cfg, rng, n_cases=8):
    """(species, particle, b0, n_links, n_beads, first, newR offsets) -- bisection-style windows,
    one wrapping past n_bead, one whole-path displacement."""
    M = cfg.n_bead
    cases = []
    for t in range(n_cases):
        sp = t % len(cfg.species)
        N = cfg.species[sp].n_part
        nb = [4, 2, M, 4, 8 if M >= 8 else 2, M, 2, 4][t % 8]
        nb = min(nb, M)
        p = int(rng.integers(0, N))
        b0 = int(rng.integers(0, M))
        if t % 8 == 3:
            b0 = M - 2
        if nb == M:
            b0, first, n_beads = 0, 0, M
        else:
            first, n_beads = (b0 + 1) % M, nb - 1
        if n_beads == 0:
            continue
        cases.append((sp, p, b0, nb, n_beads, first, 0.08 * rng.standard_normal((n_beads, cfg.n_d)), int(rng.integers(0, 2))))
    return cases


def make_one(name):
    cfg = CONFIGS[name]()
    cfg.observables = observables_for(cfg)
    sim = refsim.RefSim(cfg, seed=SEED)
    out = {"seed": np.int64(SEED)}
    Rs = []
    for sp in range(len(cfg.species)):
        R = S.synthetic_paths(cfg, sp, 0, SEED)
        sim.set_positions(sp, R)
        Rs.append(R)
    n_act = len(cfg.actions)
    if sim.n_k() > 0:
        idx, _, mags, mx = sim.kspace()
        # KSpace stores table positions max_index + lattice index (k_space_class.h:62-70)
        out["kspace_index"], out["kspace_mag"] = (idx - mx[None, :]).astype(np.int32), mags
        for sp in range(len(cfg.species)):
            out["rhok_%d" % sp] = sim.rhok(sp)
    out["dbeta"] = np.array([sim.dbeta(a) for a in range(n_act)])
    parts = [(s, p) for s in range(len(cfg.species)) for p in range(cfg.species[s].n_part)]
    sim.set_mode(0)
    out["total"] = np.array([sim.get_action(a, 0, 0, cfg.n_bead, parts, 0) for a in range(n_act)])
    sim.set_mode(1)
    david_lr = [a.type == "DavidPairAction" and a.use_long_range for a in cfg.actions]
    out["potential"] = np.array([np.nan if david_lr[a] else sim.potential(a) for a in range(n_act)])
    # per-pair kernels
    rng = np.random.default_rng(SEED + 1)
    n = 2000
    rmax = np.sqrt(3) * cfg.L / 2
    r = rng.uniform(1e-3, rmax, n)
    rp = np.clip(r + rng.normal(0, 0.1, n), 1e-4, rmax)
    s = np.abs(r - rp) + np.abs(rng.normal(0, 0.05, n))
    r[:4] = [1e-5, rmax, 0.3, 2.0]
    rp[:4] = [1e-5, rmax * 1.5, 0.3, 2.0]
    s[:4] = [0.0, 0.1, 0.0, 0.0]
    out["pair_r"], out["pair_rp"], out["pair_s"] = r, rp, s
    for a in range(n_act):
        for which in (0, 1, 2):
            out["pair_%d_%d" % (a, which)] = sim.calc_pair(a, which, r, rp, s)
    # estimators on the initial configuration
    oi = 0
    ns = len(cfg.species)
    probe_r = np.concatenate([np.linspace(0, cfg.L, 4001), rng.uniform(0, cfg.L, 2000)])
    for sa in range(ns):
        for sb in range(sa, ns):
            sim.observable_accumulate(oi)
            out["gofr_%d%d" % (sa, sb)] = sim.gofr_counts(oi, 100)
            if sa == 0 and sb == 0:
                out["gofr_probe_r"] = probe_r
                out["gofr_probe_bins"] = sim.gofr_bins(oi, probe_r)
            oi += 1
            if cfg.k_cut is not None:
                sim.observable_accumulate(oi)
                out["sofk_%d%d" % (sa, sb)] = sim.sofk_sums(oi)
                oi += 1
    # spatial derivatives on the initial configuration (the last case wraps past n_bead)
    grng = np.random.default_rng(SEED + 3)
    metas, gvals, lvals = [], [], []
    for t in range(6):
        sp = t % ns
        p = int(grng.integers(0, cfg.species[sp].n_part))
        n_w = [1, 3, 2, cfg.n_bead, 2, 2][t]
        b0 = int(grng.integers(0, cfg.n_bead - n_w + 1)) if t < 5 else cfg.n_bead - 1
        metas.append([sp, p, b0, n_w])
        gvals.append([sim.action_gradient(a, 1, b0, b0 + n_w, [(sp, p)], 0) for a in range(n_act)])
        lvals.append([sim.action_laplacian(a, 1, b0, b0 + n_w, [(sp, p)], 0) for a in range(n_act)])
    out["grad_meta"], out["grad_val"], out["lap_val"] = np.array(metas, dtype=np.int64), np.array(gvals), np.array(lvals)
    # move windows: OLD / NEW action of every action touching the species, then commit
    cases = window_cases(cfg, np.random.default_rng(SEED + 2))
    out["win_n"] = np.int64(len(cases))
    for t, (sp, p, b0, nb, n_beads, first, dR, accept) in enumerate(cases):
        M = cfg.n_bead
        cur = sim.get_positions(sp, 0)
        newR = cur[p, (first + np.arange(n_beads)) % M] + dR
        sim.propose(sp, p, first, newR)
        old = np.full(n_act, np.nan)
        new = np.full(n_act, np.nan)
        for a, acfg in enumerate(cfg.actions):
            if cfg.species[sp].name not in (acfg.species_a, acfg.species_b):
                continue
            old[a] = sim.get_action(a, 0, b0, b0 + nb, [(sp, p)], 0)
            new[a] = sim.get_action(a, 1, b0, b0 + nb, [(sp, p)], 0)
        sim.finish_move(sp, p, b0, b0 + nb, bool(accept))
        out["win_%d_meta" % t] = np.array([sp, p, b0, nb, n_beads, first, accept], dtype=np.int64)
        out["win_%d_newR" % t] = newR
        out["win_%d_old" % t], out["win_%d_new" % t] = old, new
        out["win_%d_pos" % t] = sim.get_positions(sp, 0)
        if sim.n_k() > 0:
            out["win_%d_rhok" % t] = sim.rhok(sp, 0)
    out["dbeta_after"] = np.array([sim.dbeta(a) for a in range(n_act)])
    sim.close()
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    np.savez_compressed(os.path.join(GOLDEN_DIR, name + ".npz"), **out)
    return out


def make_perm_table():
    """PermBisectIterative::UpdatePermTable of the reference on seeded paths -> tests/golden/perm_table_n7.npz."""
    cfg = S.ueg_config(N=7, M=16, with_kinetic=True)
    cfg.moves = [{"name": "PermE", "type": "PermBisectIterative", "species": "e", "n_level": 2, "n_images": 0}]
    sim = refsim.RefSim(cfg, seed=SEED)
    sim.set_positions(0, S.synthetic_paths(cfg, 0, 0, SEED))
    b0s = np.array([0, 3, 11, 14, 15], dtype=np.int64)      # the last three windows wrap past n_bead
    out = {"seed": np.int64(SEED), "n_bisect_beads": np.int64(4), "b0": b0s, "t": np.stack([sim.perm_table(0, int(b), 7) for b in b0s])}
    sim.close()
    np.savez_compressed(os.path.join(GOLDEN_DIR, "perm_table_n7.npz"), **out)
    return out


def main():
    if not refsim.available():
        raise SystemExit("oracle/_ref/libsimpimc_ref.so missing: run `make -C oracle ref` where /root/reference exists")
    for name in sorted(CONFIGS):
        out = make_one(name)
        print("%-16s dbeta %s" % (name, out["dbeta"]))
    print("perm_table_n7    max t %.6g" % make_perm_table()["t"].max())


if __name__ == "__main__":
    main()
