"""Shared checker: one backend (the CPU oracle, or the CUDA path through the C ABI) against the
fixtures tests/golden/*.npz that oracle/make_golden.py wrote from the reference itself.

Tolerances (BASELINE.json north_star): floating point |x - ref| <= 1e-10 * |ref| (for sums that
are differences of larger terms, relative to the size of those terms); k vectors, histogram
bins and positions after commit bit-exact."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from simpimc_b200 import system as S  # noqa: E402

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")
RTOL = 1e-10

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


def load(name):
    return np.load(os.path.join(GOLDEN_DIR, name + ".npz"))


def rel_ok(got, ref, scale=None, rtol=RTOL):
    got, ref = np.asarray(got, dtype=np.float64), np.asarray(ref, dtype=np.float64)
    sc = np.abs(ref) if scale is None else np.maximum(np.abs(ref), scale)
    return bool(np.all(np.abs(got - ref) <= rtol * np.maximum(sc, 1e-300)))


class OracleBackend:
    """oracle.Oracle (one walker)."""

    def __init__(self, cfg, seed):
        from oracle import oracle as O
        self.cfg = cfg
        self.o = O.Oracle(cfg)
        for sp in range(len(cfg.species)):
            self.o.set_positions(sp, S.synthetic_paths(cfg, sp, 0, seed))

    def close(self):
        self.o.close()

    def kspace(self):
        return self.o.kspace()

    def rhok(self, sp):
        return self.o.rhok(sp, 0)

    def dbeta(self, a):
        return self.o.dbeta(a)

    def potential(self, a):
        return self.o.potential(a)

    def total(self, a):
        cfg = self.cfg
        parts = [(s, p) for s in range(len(cfg.species)) for p in range(cfg.species[s].n_part)]
        return self.o.get_action(a, 0, 0, cfg.n_bead, parts, 0)

    def calc_pair(self, a, which, r, rp, s):
        return self.o.calc_pair(a, which, r, rp, s)

    def gofr_counts(self, sa, sb, n_r=100):
        return self.o.gofr(sa, sb, 0.0, self.cfg.L / 2.0, n_r)[1]

    def sofk(self, sa, sb):
        return self.o.sofk(sa, sb, self.cfg.k_cut)

    def gradient(self, a, b0, b1, sp, p):
        return self.o.action_gradient(a, 1, b0, b1, [(sp, p)], 0)

    def laplacian(self, a, b0, b1, sp, p):
        return self.o.action_laplacian(a, 1, b0, b1, [(sp, p)], 0)

    def propose(self, sp, p, first, newR):
        self.o.propose(sp, p, first, newR)

    def get_action(self, a, mode, b0, b1, sp, p):
        return self.o.get_action(a, mode, b0, b1, [(sp, p)], 0)

    def finish_move(self, sp, p, b0, b1, accept):
        self.o.finish_move(sp, p, b0, b1, accept)

    def positions(self, sp):
        return self.o.get_positions(sp, 0)


class GpuBackend:
    """simpimc_b200.host.Path with two identical clones (both are checked)."""

    def __init__(self, cfg, seed, general=False):
        from simpimc_b200 import host
        self.host = host
        self.cfg = cfg
        self.path = host.Path(cfg, n_clones=2)
        self.path.ForceGeneral(general)
        for sp in range(len(cfg.species)):
            R = S.synthetic_paths(cfg, sp, 0, seed)
            self.path.SetPositions(sp, np.stack([R, R]))

    def close(self):
        self.path.close()

    def _both(self, v):
        assert v[0] == v[1], "identical clones gave different results"
        return v[0]

    def kspace(self):
        return self.path.KSpace()

    def rhok(self, sp):
        a, b = self.path.GetRhoK(sp, 0, self.host.OLD_MODE), self.path.GetRhoK(sp, 1, self.host.OLD_MODE)
        assert np.array_equal(a, b)
        return a

    def dbeta(self, a):
        return self._both(self.path.actions[a].DActionDBeta())

    def potential(self, a):
        return self._both(self.path.actions[a].Potential())

    def total(self, a):
        return self._both(self.path.actions[a].TotalAction())

    def calc_pair(self, a, which, r, rp, s):
        return self.path.actions[a].CalcPair(which, r, rp, s)

    def gofr_counts(self, sa, sb, n_r=100):
        c = self.host.PairCorrelation(self.path, sa, sb, 0.0, self.cfg.L / 2.0, n_r).Counts()
        assert np.array_equal(c[0], c[1])
        return c[0]

    def sofk(self, sa, sb):
        sk = self.host.StructureFactor(self.path, sa, sb, self.cfg.k_cut)
        sk.Accumulate()
        assert np.array_equal(sk.sk[0], sk.sk[1])
        return sk.sk[0]

    def gradient(self, a, b0, b1, sp, p):
        g = self.path.actions[a].GetActionGradient(b0, b1, [(sp, p)], 0)
        assert np.array_equal(g[0], g[1])
        return g[0]

    def laplacian(self, a, b0, b1, sp, p):
        return self._both(self.path.actions[a].GetActionLaplacian(b0, b1, [(sp, p)], 0))

    def propose(self, sp, p, first, newR):
        self.path.Propose(sp, p, first, np.stack([newR, newR]))

    def get_action(self, a, mode, b0, b1, sp, p):
        self.path.SetMode(mode)
        return self._both(self.path.actions[a].GetAction(b0, b1, [(sp, p)], 0))

    def finish_move(self, sp, p, b0, b1, accept):
        self.path.Commit(1 if accept else 0)

    def positions(self, sp):
        R = self.path.GetPositions(sp)
        assert np.array_equal(R[0], R[1])
        return R[0]


def check_backend(name, make_backend):
    """Every quantity of the fixture `name` against one backend."""
    g = load(name)
    cfg = CONFIGS[name]()
    be = make_backend(cfg, int(g["seed"]))
    n_act = len(cfg.actions)
    ns = len(cfg.species)
    has_k = "kspace_index" in g
    if has_k:
        idx, mags = be.kspace()
        assert np.array_equal(idx, g["kspace_index"]), "k-vector list / order"
        assert np.array_equal(mags, g["kspace_mag"])
        for sp in range(ns):
            assert np.max(np.abs(be.rhok(sp) - g["rhok_%d" % sp])) <= 1e-12 * cfg.species[sp].n_part
    for a in range(n_act):
        assert rel_ok(be.dbeta(a), g["dbeta"][a]), (name, "dbeta", a, be.dbeta(a), g["dbeta"][a])
        assert rel_ok(be.total(a), g["total"][a]), (name, "total", a, be.total(a), g["total"][a])
        if not np.isnan(g["potential"][a]):
            assert rel_ok(be.potential(a), g["potential"][a]), (name, "potential", a)
        for which in (0, 1, 2):
            ref = g["pair_%d_%d" % (a, which)]
            got = be.calc_pair(a, which, g["pair_r"], g["pair_rp"], g["pair_s"])
            assert rel_ok(got, ref, scale=1e-3 * np.max(np.abs(ref))), (name, "pair", a, which, np.max(np.abs(got - ref)))
    for sa in range(ns):
        for sb in range(sa, ns):
            assert np.array_equal(np.asarray(be.gofr_counts(sa, sb), dtype=np.float64), g["gofr_%d%d" % (sa, sb)]), "g(r) bins"
            if has_k:
                ref = g["sofk_%d%d" % (sa, sb)]
                assert np.max(np.abs(be.sofk(sa, sb) - ref)) <= 1e-10 * max(1.0, np.max(np.abs(ref)))
    # spatial derivatives (pair_action_class.h:305-366).  Analytic Ilkka gradients: RTOL of the
    # largest component; central differences (eps = 1e-4) amplify the rounding of U by 1 / eps
    # (gradient) and 1 / eps^2 (Laplacian): RTOL of |U of the window's pairs| / eps resp. / eps^2
    if "grad_meta" in g:
        eps = 1e-4
        for t, (sp, p, b0, n_w) in enumerate(g["grad_meta"].tolist()):
            for a in range(n_act):
                g_ref, l_ref = g["grad_val"][t][a], g["lap_val"][t][a]
                u_scale = 2.0 * abs(be.get_action(a, 0, b0, b0 + n_w, sp, p)) + 1e-3   # OLD mode: leaves the rho_k refresh flag alone
                analytic = cfg.actions[a].type == "IlkkaPairAction"
                got = be.gradient(a, b0, b0 + n_w, sp, p)
                tol = RTOL * max(np.max(np.abs(g_ref)), u_scale if analytic else u_scale / eps)
                assert np.max(np.abs(got - g_ref)) <= tol, (name, "gradient", t, a, got, g_ref)
                got = be.laplacian(a, b0, b0 + n_w, sp, p)
                assert abs(got - l_ref) <= RTOL * max(abs(l_ref), u_scale / eps ** 2), (name, "laplacian", t, a, got, l_ref)
    for t in range(int(g["win_n"])):
        sp, p, b0, nb, n_beads, first, accept = [int(v) for v in g["win_%d_meta" % t]]
        be.propose(sp, p, first, g["win_%d_newR" % t])
        for a in range(n_act):
            ro, rn = g["win_%d_old" % t][a], g["win_%d_new" % t][a]
            if np.isnan(ro):
                continue
            go = be.get_action(a, 0, b0, b0 + nb, sp, p)
            gn = be.get_action(a, 1, b0, b0 + nb, sp, p)
            assert rel_ok(go, ro) and rel_ok(gn, rn), (name, t, a, go, ro, gn, rn)
            assert abs((gn - go) - (rn - ro)) <= RTOL * max(abs(rn - ro), 1e-4 * (abs(rn) + abs(ro))), (name, t, a, "delta")
        be.finish_move(sp, p, b0, b0 + nb, bool(accept))
        assert np.array_equal(be.positions(sp), g["win_%d_pos" % t]), (name, t, "positions after commit")
        if has_k:
            assert np.max(np.abs(be.rhok(sp) - g["win_%d_rhok" % t])) <= 1e-11 * cfg.species[sp].n_part
    for a in range(n_act):
        assert rel_ok(be.dbeta(a), g["dbeta_after"][a]), (name, "dbeta after moves", a)
    be.close()
