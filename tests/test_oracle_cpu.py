"""CPU tests (no GPU): the oracle -- oracle/pimc_oracle.cc, the CPU restatement the GPU parity
tests check against -- is itself pinned here:

* against tests/golden/*.npz, numbers produced by the REFERENCE'S OWN classes compiled in place
  (oracle/make_golden.py; the reference tree does not travel, the fixtures do);
* against the reference library directly where oracle/_ref exists (this container);
* against the known answers the reference tree offers for this path: the k-vector list of
  KSpace::Setup (182 vectors / 17 shells for every shipped k_cut = 14/(L/2), first and last
  triples) and the Madelung constant of the StandardEwald breakup (scripts/pagen/Ewald.py).
"""
import os

import numpy as np
import pytest

import golden_util as G
from simpimc_b200 import system as S, tables as T


@pytest.mark.parametrize("name", sorted(G.CONFIGS))
def test_oracle_matches_reference_golden(name, oracle_mod):
    G.check_backend(name, G.OracleBackend)


def test_kspace_known_answers(oracle_mod):
    # 182 half-space vectors in 17 shells; first kept triples (1,-4,-1), (1,-1,-4),
    # (0,1,-4); last (3,2,2) -- for every shipped k_cut = 14/(L/2)
    for N in (7, 33, 256):
        cfg = S.ueg_config(N=N, M=4, n_xy=20, n_r_long=50)
        o = oracle_mod.Oracle(cfg)
        idx, mags = o.kspace()
        assert len(idx) == 182
        assert len(np.unique(np.round(mags / (2 * np.pi / cfg.L), 9))) == 17
        assert [tuple(v) for v in idx[:3]] == [(1, -4, -1), (1, -1, -4), (0, 1, -4)]
        assert tuple(idx[-1]) == (3, 2, 2)
        o.close()
    # strict '<' at the default cutoff 2 pi / L leaves no vector (App. A-14)
    cfg = S.ueg_config(N=7, M=4, n_xy=20, n_r_long=50)
    cfg.k_cut = 2 * np.pi / cfg.L
    cfg.actions[0].k_cut = cfg.k_cut
    o = oracle_mod.Oracle(cfg)
    assert len(o.kspace()[0]) == 0
    o.close()


def test_madelung_constant_from_ewald_breakup(oracle_mod):
    """NaCl cell (4 Na+ and 4 Cl- on a cube of side 2, nearest-neighbour distance 1), one time
    slice: BarePairAction::Potential with the StandardEwald tables of scripts/pagen/Ewald.py
    (v = Z1 Z2 / r short range, long-range r and k parts, constants) sums to N_pairs * Madelung.
    Exact value -1.7475645946331822 (Ewald.py:743-748); the breakup at this cutoff with the
    minimum-image short-range sum reproduces it to ~4e-8."""
    L = 2.0
    k_cut = 30.0 / (L / 2.0)  # alpha * L/2 = sqrt(15): the minimum-image short-range sum is converged to ~1e-7
    cfg = S.SystemConfig(n_d=3, n_bead=1, beta=1.0, L=L, pbc=True, k_cut=k_cut)
    cfg.species.append(S.SpeciesConfig("Na", 4, 0.5))
    cfg.species.append(S.SpeciesConfig("Cl", 4, 0.5))
    for nm, a, b, z in (("NaNa", "Na", "Na", 1.0), ("NaCl", "Na", "Cl", -1.0), ("ClCl", "Cl", "Cl", 1.0)):
        # v(r) = Z1 Z2 / r as a spline table (the analytic is_coulomb branch carries no charges)
        tab = T.make_bare_table(z, L, k_cut, n_r=4000, r_max=4.0, n_r_long=2000)
        cfg.actions.append(S.ActionConfig(nm, "BarePairAction", a, b, table=tab, max_level=0, use_long_range=True, k_cut=k_cut))
    o = oracle_mod.Oracle(cfg)
    na = np.array([[0, 0, 0], [1, 1, 0], [1, 0, 1], [0, 1, 1]], dtype=float) - 0.5
    cl = np.array([[1, 0, 0], [0, 1, 0], [0, 0, 1], [1, 1, 1]], dtype=float) - 0.5
    o.set_positions(0, na.reshape(4, 1, 3))
    o.set_positions(1, cl.reshape(4, 1, 3))
    v = sum(o.potential(a) for a in range(3))
    o.close()
    madelung = v / 4.0  # energy per ion pair, nearest-neighbour distance 1
    assert abs(madelung - (-1.7475645946331822)) < 2e-7, madelung


def test_gofr_bin_map_matches_reference(oracle_mod):
    g = G.load("ilkka_lr_n7")
    cfg = G.CONFIGS["ilkka_lr_n7"]()
    bins = oracle_mod.gofr_bins(0.0, cfg.L / 2.0, 100, g["gofr_probe_r"])
    ref = g["gofr_probe_bins"]
    # negative arguments wrap to huge unsigned values in the reference (App. A-11); both sides
    # agree on every index that passes the i < n_r test and on which samples are dropped
    keep = ref < 100
    assert np.array_equal(bins < 100, keep)
    assert np.array_equal(bins[keep], ref[keep])


@pytest.mark.ref
@pytest.mark.parametrize("seed", [1, 2, 3])
def test_oracle_matches_live_reference(seed, oracle_mod):
    """Fresh seeds against the reference library itself (only where oracle/_ref was built)."""
    from oracle import refsim
    if not refsim.available():
        pytest.skip("oracle/_ref not built (no /root/reference on this machine)")
    cfg = S.plasma_config(Ne=5, Np=4, M=8)
    sim = refsim.RefSim(cfg, seed=seed)
    o = oracle_mod.Oracle(cfg)
    for sp in range(2):
        R = S.synthetic_paths(cfg, sp, 0, 1000 + seed)
        sim.set_positions(sp, R)
        o.set_positions(sp, R)
    for a in range(3):
        assert G.rel_ok(o.dbeta(a), sim.dbeta(a))
        assert G.rel_ok(o.potential(a), sim.potential(a))
    for sp in range(2):
        assert np.max(np.abs(o.rhok(sp) - sim.rhok(sp))) <= 1e-12 * 5
    sim.close()
    o.close()


DAVID_CASES = {
    "david_lin": lambda: S.ueg_config(N=9, M=8, action="DavidPairAction", use_long_range=False, david_grid="LINEAR", david_n_grid=90),
    "david_o1": lambda: S.ueg_config(N=9, M=8, action="DavidPairAction", use_long_range=False, david_n_order=1),
    "david_o3": lambda: S.ueg_config(N=9, M=8, action="DavidPairAction", use_long_range=False, david_n_order=3, david_grid="LINEAR",
                                     david_n_grid=90),
    "plasma_david": lambda: S.plasma_config(Ne=5, Np=4, M=8, pp_action="DavidPairAction", ep_action="DavidPairAction"),
}


@pytest.mark.ref
@pytest.mark.parametrize("name", sorted(DAVID_CASES))
def test_oracle_matches_live_reference_david_variants(name, oracle_mod):
    """DavidPairAction on a linear grid, with n_order 1 and 3, and between different species: the table shapes the GPU
    parity tests use beyond the golden fixture, pinned to the reference library itself (where oracle/_ref was built)."""
    from oracle import refsim
    if not refsim.available():
        pytest.skip("oracle/_ref not built (no /root/reference on this machine)")
    cfg = DAVID_CASES[name]()
    sim = refsim.RefSim(cfg, seed=1)
    o = oracle_mod.Oracle(cfg)
    for sp in range(len(cfg.species)):
        R = S.synthetic_paths(cfg, sp, 0, 1001)
        sim.set_positions(sp, R)
        o.set_positions(sp, R)
    for a in range(len(cfg.actions)):
        assert G.rel_ok(o.dbeta(a), sim.dbeta(a))
        assert G.rel_ok(o.potential(a), sim.potential(a))
    sim.close()
    o.close()


def test_spline_definition_matches_scipy_natural_cubic_spline(oracle_mod):
    """The spline library the reference links (etano/meinspline, unpinned) is absent; the oracle
    restates einspline's published algorithm.  Independent check of that restatement: an
    interpolating cubic spline with NATURAL boundaries (zero second derivative at both ends) is
    unique, so it must coincide with scipy.interpolate.CubicSpline(bc_type='natural') -- in 1-D on
    the reference's non-uniform OPTIMIZED grid, and in 2-D as the tensor product (natural along
    x for every y column, then natural along y), which is how create_NUBspline_2d_d solves."""
    from scipy.interpolate import CubicSpline
    from simpimc_b200 import tables as T
    O = oracle_mod
    rng = np.random.default_rng(3)
    g = T.gen_grid("OPTIMIZED", 1.0e-4, 9.0, 200)
    f = np.exp(-0.3 * g) * np.cos(1.7 * g) / (0.2 + g)
    x = np.concatenate([rng.uniform(g[0], g[-1], 4000), g, [g[0], g[-1]]])
    ref = CubicSpline(g, f, bc_type="natural")(x)
    got = O.spline1d_eval(g, f, x)
    assert np.max(np.abs(got - ref)) <= 1e-12 * np.max(np.abs(f))
    # 2-D tensor product on the (x, y) grid of an off-diagonal table
    gx = T.gen_grid("OPTIMIZED", 0.0, 12.0, 40)
    gy = T.gen_grid("OPTIMIZED", 0.0, 12.0, 35)
    X, Y = np.meshgrid(gx, gy, indexing="ij")
    F = np.exp(-0.2 * (X + Y)) * (1.0 + 0.3 * np.sin(X - Y)) / (0.5 + X * Y / 10.0)
    xs, ys = rng.uniform(gx[0], gx[-1], 500), rng.uniform(gy[0], gy[-1], 500)
    got2 = O.spline2d_eval(gx, gy, F, xs, ys)
    along_y = CubicSpline(gy, F, axis=1, bc_type="natural")          # for every x row: natural spline in y
    ref2 = np.array([CubicSpline(gx, along_y(yy), bc_type="natural")(xx) for xx, yy in zip(xs, ys)])
    assert np.max(np.abs(got2 - ref2)) <= 1e-12 * np.max(np.abs(F))


def test_perm_table_matches_reference_golden(oracle_mod):
    """PermBisectIterative::UpdatePermTable (perm_bisect_iterative_class.h:10-30): the restatement
    against the reference's own table (tests/golden/perm_table_n7.npz, oracle/make_golden.py)."""
    from simpimc_b200 import system as S
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "perm_table_n7.npz"))
    cfg = S.ueg_config(N=7, M=16)
    o = oracle_mod.Oracle(cfg)
    o.set_positions(0, S.synthetic_paths(cfg, 0, 0, int(g["seed"])))
    for b0, t_ref in zip(g["b0"], g["t"]):
        assert np.array_equal(o.perm_table(0, int(b0), int(g["n_bisect_beads"])), t_ref)
    # the PermBisectTable variant differs by the row factor exp(+|dr_ii|^2 / (4 lambda tau n))
    t_rel = o.perm_table(0, 3, 4, relative=True)
    t_abs = o.perm_table(0, 3, 4)
    assert np.allclose(np.diag(t_rel), 1.0) and np.allclose(t_rel * np.diag(t_abs)[:, None], t_abs, rtol=1e-12)
    o.close()
