"""TEST INFRASTRUCTURE ONLY -- ctypes front end of oracle/liboracle.so (the CPU restatement
in oracle/pimc_oracle.cc).  Imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs only; never by the product package.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from simpimc_b200 import capi

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "liboracle.so")

_lib = None


def build(force=False):
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < max(
            os.path.getmtime(os.path.join(HERE, f)) for f in ("pimc_oracle.cc", "spline_oracle.h")):
        subprocess.check_call(["make", "-C", HERE, "liboracle.so"], stdout=subprocess.DEVNULL)


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB)
        vp, i32, dbl = C.c_void_p, C.c_int32, C.c_double
        L.orc_create.restype = vp
        L.orc_create.argtypes = [C.POINTER(capi.Config)]
        L.orc_destroy.argtypes = [vp]
        L.orc_kspace_setup.restype = i32
        L.orc_kspace_setup.argtypes = [vp, dbl]
        L.orc_kspace_get.argtypes = [vp, vp, vp]
        L.orc_set_positions.argtypes = [vp, i32, vp]
        L.orc_get_positions.argtypes = [vp, i32, i32, vp]
        L.orc_rhok.argtypes = [vp, i32, i32, vp]
        L.orc_set_mode.argtypes = [vp, i32]
        L.orc_action_create_ilkka.restype = i32
        L.orc_action_create_ilkka.argtypes = [vp, i32, i32, C.POINTER(capi.IlkkaTables), i32, i32, dbl]
        L.orc_action_create_bare.restype = i32
        L.orc_action_create_bare.argtypes = [vp, i32, i32, C.POINTER(capi.BareTables), i32, i32, dbl]
        L.orc_action_create_david.restype = i32
        L.orc_action_create_david.argtypes = [vp, i32, i32, C.POINTER(capi.DavidTables), i32, i32]
        L.orc_dbeta.restype = dbl
        L.orc_dbeta.argtypes = [vp, i32]
        L.orc_potential.restype = dbl
        L.orc_potential.argtypes = [vp, i32]
        L.orc_get_action.restype = dbl
        L.orc_get_action.argtypes = [vp, i32, i32, i32, i32, i32, vp, vp, i32]
        L.orc_perm_table.argtypes = [vp, i32, i32, i32, C.c_double, i32, vp]
        L.orc_action_gradient.argtypes = [vp, i32, i32, i32, i32, i32, vp, vp, i32, vp]
        L.orc_action_laplacian.restype = dbl
        L.orc_action_laplacian.argtypes = [vp, i32, i32, i32, i32, i32, vp, vp, i32]
        L.orc_calc_pair.argtypes = [vp, i32, i32, i32, vp, vp, vp, i32, vp]
        L.orc_calc_long.restype = dbl
        L.orc_calc_long.argtypes = [vp, i32, i32, i32, i32, i32]
        L.orc_dr_drp_drrp.argtypes = [vp, i32, i32, i32, i32, i32, i32, vp]
        L.orc_propose.argtypes = [vp, i32, i32, i32, i32, vp]
        L.orc_finish_move.argtypes = [vp, i32, i32, i32, i32, i32]
        L.orc_gofr_bins.argtypes = [dbl, dbl, i32, i32, vp, vp]
        L.orc_gofr.argtypes = [vp, i32, i32, dbl, dbl, i32, dbl, vp, vp]
        L.orc_sofk.argtypes = [vp, i32, i32, dbl, dbl, vp]
        L.orc_spline1d_eval.argtypes = [i32, vp, vp, i32, vp, vp]
        L.orc_spline1d_coefs.argtypes = [i32, vp, vp, vp]
        L.orc_spline2d_eval.argtypes = [i32, i32, vp, vp, vp, i32, vp, vp, vp]
        L.orc_spline2d_coefs.argtypes = [i32, i32, vp, vp, vp, vp]
        L.orc_grid_reverse_map.argtypes = [i32, vp, i32, vp, vp]
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _d(a):
    return np.ascontiguousarray(a, dtype=np.float64)


class Oracle:
    """One walker of a SystemConfig evaluated by the CPU restatement."""

    def __init__(self, cfg):
        self.L = lib()
        self.cfg = cfg
        c, self._keep = capi.make_config(cfg)
        self.h = self.L.orc_create(C.byref(c))
        self.n_k = 0
        if cfg.k_cut is not None and cfg.pbc:
            self.n_k = self.L.orc_kspace_setup(self.h, cfg.k_cut)
        self.actions = []
        self._tables = []
        for a in cfg.actions:
            if a.type == "Kinetic":
                self.actions.append(None)
                continue
            sa, sb = cfg.species_index(a.species_a), cfg.species_index(a.species_b)
            kc = a.k_cut if a.k_cut is not None else (cfg.k_cut or 0.0)
            if a.type == "IlkkaPairAction":
                t, keep = capi.pack_ilkka(a.table, a.use_long_range)
                idx = self.L.orc_action_create_ilkka(self.h, sa, sb, C.byref(t), a.max_level, int(a.use_long_range), kc)
            elif a.type == "BarePairAction":
                t, keep = capi.pack_bare(a.table, a.use_long_range, a.is_coulomb)
                idx = self.L.orc_action_create_bare(self.h, sa, sb, C.byref(t), a.max_level, int(a.use_long_range), kc)
            elif a.type == "DavidPairAction":
                t, keep = capi.pack_david(a.table, a.n_order, a.use_long_range)
                idx = self.L.orc_action_create_david(self.h, sa, sb, C.byref(t), a.max_level, int(a.use_long_range))
            else:
                raise ValueError(a.type)
            if idx < 0:
                raise RuntimeError("oracle: action %s rejected" % a.name)
            self._tables.append((t, keep))
            self.actions.append(idx)
        if any(a.use_long_range for a in cfg.actions):
            idx, _ = self.kspace()
            self.n_k = len(idx)

    def close(self):
        if self.h:
            self.L.orc_destroy(self.h)
            self.h = None

    def kspace(self):
        # size query by growing buffers: n_k is small
        idx = np.zeros((100000, self.cfg.n_d), dtype=np.int32)
        mags = np.full(100000, -1.0)
        self.L.orc_kspace_get(self.h, _p(idx), _p(mags))
        n = int(np.sum(mags >= 0))
        return idx[:n].copy(), mags[:n].copy()

    def set_positions(self, sp, R):
        R = _d(R)
        self.L.orc_set_positions(self.h, sp, _p(R))

    def get_positions(self, sp, mode=1):
        s = self.cfg.species[sp]
        R = np.zeros((s.n_part, self.cfg.n_bead, self.cfg.n_d))
        self.L.orc_get_positions(self.h, sp, mode, _p(R))
        return R

    def rhok(self, sp, mode=1):
        out = np.zeros((self.cfg.n_bead, self.n_k, 2))
        self.L.orc_rhok(self.h, sp, mode, _p(out))
        return out[..., 0] + 1j * out[..., 1]

    def dbeta(self, a):
        return self.L.orc_dbeta(self.h, self.actions[a])

    def potential(self, a):
        return self.L.orc_potential(self.h, self.actions[a])

    def get_action(self, a, mode, b0, b1, particles, level):
        sp = np.array([p[0] for p in particles], dtype=np.int32)
        pi = np.array([p[1] for p in particles], dtype=np.int32)
        return self.L.orc_get_action(self.h, self.actions[a], mode, b0, b1, len(particles), _p(sp), _p(pi), level)

    def action_gradient(self, a, mode, b0, b1, particles, level):
        sp = np.array([p[0] for p in particles], dtype=np.int32)
        pi = np.array([p[1] for p in particles], dtype=np.int32)
        out = np.zeros(3)
        self.L.orc_action_gradient(self.h, self.actions[a], mode, b0, b1, len(particles), _p(sp), _p(pi), level, _p(out))
        return out

    def action_laplacian(self, a, mode, b0, b1, particles, level):
        sp = np.array([p[0] for p in particles], dtype=np.int32)
        pi = np.array([p[1] for p in particles], dtype=np.int32)
        return self.L.orc_action_laplacian(self.h, self.actions[a], mode, b0, b1, len(particles), _p(sp), _p(pi), level)

    def perm_table(self, sp, bead0, n_bisect_beads, epsilon=1e-100, relative=False):
        N = self.cfg.species[sp].n_part
        t = np.zeros((N, N))
        self.L.orc_perm_table(self.h, sp, bead0, n_bisect_beads, C.c_double(epsilon), 1 if relative else 0, _p(t))
        return t

    def calc_pair(self, a, which, r, rp, s, level=0):
        r, rp, s = _d(r), _d(rp), _d(s)
        out = np.zeros_like(r)
        self.L.orc_calc_pair(self.h, self.actions[a], which, len(r), _p(r), _p(rp), _p(s), level, _p(out))
        return out

    def calc_long(self, a, which, b0=0, b1=0, level=0):
        return self.L.orc_calc_long(self.h, self.actions[a], which, b0, b1, level)

    def dr_drp_drrp(self, b0, b1, sa, sb, p0, p1):
        out = np.zeros(3)
        self.L.orc_dr_drp_drrp(self.h, b0, b1, sa, sb, p0, p1, _p(out))
        return out

    def set_mode(self, mode):
        self.L.orc_set_mode(self.h, mode)

    def propose(self, sp, p, b_first, newR):
        newR = _d(newR)
        self.L.orc_propose(self.h, sp, p, b_first, newR.shape[0], _p(newR))

    def finish_move(self, sp, p, b0, b1, accept):
        self.L.orc_finish_move(self.h, sp, p, b0, b1, 1 if accept else 0)

    def gofr(self, sa, sb, r_min, r_max, n_r, cofactor=1.0):
        y = np.zeros(n_r)
        counts = np.zeros(n_r, dtype=np.uint64)
        self.L.orc_gofr(self.h, sa, sb, r_min, r_max, n_r, cofactor, _p(y), _p(counts))
        return y, counts

    def sofk(self, sa, sb, k_cut, cofactor=1.0):
        sk = np.zeros(self.n_k)
        self.L.orc_sofk(self.h, sa, sb, k_cut, cofactor, _p(sk))
        return sk


def gofr_bins(r_min, r_max, n_r, r):
    r = _d(r)
    bins = np.zeros(len(r), dtype=np.uint32)
    lib().orc_gofr_bins(r_min, r_max, n_r, len(r), _p(r), _p(bins))
    return bins


def spline1d_eval(grid, data, x):
    grid, data, x = _d(grid), _d(data), _d(x)
    out = np.zeros_like(x)
    lib().orc_spline1d_eval(len(grid), _p(grid), _p(data), len(x), _p(x), _p(out))
    return out


def spline1d_coefs(grid, data):
    grid, data = _d(grid), _d(data)
    out = np.zeros(len(grid) + 2)
    lib().orc_spline1d_coefs(len(grid), _p(grid), _p(data), _p(out))
    return out


def spline2d_eval(gx, gy, data, x, y):
    gx, gy, data, x, y = _d(gx), _d(gy), _d(data), _d(x), _d(y)
    out = np.zeros_like(x)
    lib().orc_spline2d_eval(len(gx), len(gy), _p(gx), _p(gy), _p(data), len(x), _p(x), _p(y), _p(out))
    return out


def spline2d_coefs(gx, gy, data):
    gx, gy, data = _d(gx), _d(gy), _d(data)
    out = np.zeros((len(gx) + 2, len(gy) + 2))
    lib().orc_spline2d_coefs(len(gx), len(gy), _p(gx), _p(gy), _p(data), _p(out))
    return out


def grid_reverse_map(grid, x):
    grid, x = _d(grid), _d(x)
    out = np.zeros(len(x), dtype=np.int32)
    lib().orc_grid_reverse_map(len(grid), _p(grid), len(x), _p(x), _p(out))
    return out
