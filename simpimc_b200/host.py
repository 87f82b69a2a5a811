"""Host-side mirror of the reference's operator interface for the action-evaluation path,
bound to the CUDA library through the C ABI (include/simpimc_b200.h).

Names and argument meaning follow the reference so tests read like calls into it:

* `Path`        -- src/data_structures/path_class.h (box, tau, species, KSpace, OLD/NEW mode)
                   holding `n_clones` walkers instead of one;
* `PairAction`  -- src/actions/action_class.h:37-73 as implemented by
                   src/actions/pair_action/{ilkka,bare,david}_pair_action_class.h:
                   `DActionDBeta()`, `Potential()`, `GetAction(b0, b1, particles, level)`,
                   `Accept()`, `Reject()`;
* `PairCorrelation`, `StructureFactor`, `Energy` -- src/events/observables/*_class.h
                   (`Accumulate()`, and the normalisation of `Write()`).

Every method returns one value per clone (numpy array of length n_clones).
There is no CPU fallback: constructing a `Path` without the built CUDA library or without a
GPU raises.
"""
import ctypes as C
import math

import numpy as np

from . import capi

OLD_MODE, NEW_MODE = capi.PIMC_OLD, capi.PIMC_NEW


def _vp(a):
    return a.ctypes.data_as(C.c_void_p)


class Path:
    """Path + Species[] + KSpace of `n_clones` independent walkers on one GPU."""

    def __init__(self, cfg, n_clones=1, device=0, slice_lo=0, slice_hi=None):
        self.L = capi.lib()
        self.cfg = cfg
        self.n_clones = n_clones
        self.n_d = cfg.n_d
        self.n_bead = cfg.n_bead
        self.slice_lo = slice_lo
        self.slice_hi = cfg.n_bead if slice_hi is None else slice_hi
        self.sharded = (self.slice_hi - self.slice_lo) != cfg.n_bead
        self.n_store = (self.slice_hi - self.slice_lo) + (1 if self.sharded else 0)
        c, self._keep = capi.make_config(cfg, n_clones, device, slice_lo, self.slice_hi)
        h = C.c_void_p()
        capi.check(self.L.pimc_ctx_create(C.byref(c), C.byref(h)))
        self.h = h
        self.mode = NEW_MODE
        self.n_k = 0
        if cfg.pbc and cfg.k_cut is not None:
            self.SetupKSpace(cfg.k_cut)
        self.actions = []
        for a in cfg.actions:
            self.actions.append(Kinetic(self, a) if a.type == "Kinetic" else PairAction(self, a))
        if any(a.use_long_range for a in self.actions):
            self.n_k = self._n_k()

    def close(self):
        if self.h:
            self.L.pimc_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- KSpace (k_space_class.h) -------------------------------------------------------
    def _n_k(self):
        if not self.cfg.pbc:
            return 0
        n = C.c_int32()
        capi.check(self.L.pimc_kspace_setup(self.h, 0.0, C.byref(n)))
        return n.value

    def SetupKSpace(self, k_cut):
        n = C.c_int32()
        capi.check(self.L.pimc_kspace_setup(self.h, float(k_cut), C.byref(n)))
        self.n_k = n.value
        return self.n_k

    def KSpace(self):
        n_k = self._n_k()
        idx = np.zeros((n_k, self.n_d), dtype=np.int32)
        mags = np.zeros(n_k)
        capi.check(self.L.pimc_kspace_get(self.h, _vp(idx), _vp(mags)))
        return idx, mags

    # -- mode flag (path_class.h:95) -----------------------------------------------------
    def SetMode(self, mode):
        self.mode = mode

    def GetMode(self):
        return self.mode

    def GetTau(self):
        return self.cfg.tau

    # -- positions ----------------------------------------------------------------------
    def SetPositions(self, species, R, clone_lo=0, clone_hi=None):
        """R[clone][particle][bead][dim] (host); sets r and r_c and rebuilds rho_k."""
        clone_hi = self.n_clones if clone_hi is None else clone_hi
        R = np.ascontiguousarray(R, dtype=np.float64)
        s = self.cfg.species[species]
        assert R.shape == (clone_hi - clone_lo, s.n_part, self.n_store, self.n_d), R.shape
        capi.check(self.L.pimc_positions_upload(self.h, species, clone_lo, clone_hi, _vp(R)))
        if not (clone_lo == 0 and clone_hi == self.n_clones):
            capi.check(self.L.pimc_rhok_rebuild(self.h, species))

    def GetPositions(self, species, clone_lo=0, clone_hi=None):
        clone_hi = self.n_clones if clone_hi is None else clone_hi
        s = self.cfg.species[species]
        R = np.zeros((clone_hi - clone_lo, s.n_part, self.n_store, self.n_d))
        capi.check(self.L.pimc_positions_download(self.h, species, OLD_MODE, clone_lo, clone_hi, _vp(R)))
        return R

    def GetRhoK(self, species, clone=0, mode=None):
        mode = self.mode if mode is None else mode
        n_k = self._n_k()
        out = np.zeros((self.slice_hi - self.slice_lo, n_k, 2))
        capi.check(self.L.pimc_rhok_download(self.h, species, mode, clone, _vp(out)))
        return out[..., 0] + 1j * out[..., 1]

    # -- the moves' side of the contract --------------------------------------------------
    def Propose(self, species, particle, b_first, newR):
        """NEW-mode SetR of beads b_first[c]..+n_beads-1 of particle[c]; newR[c][i][dim]."""
        particle = np.ascontiguousarray(np.broadcast_to(particle, (self.n_clones,)), dtype=np.int32)
        b_first = np.ascontiguousarray(np.broadcast_to(b_first, (self.n_clones,)), dtype=np.int32)
        newR = np.ascontiguousarray(newR, dtype=np.float64)
        assert newR.shape[0] == self.n_clones and newR.shape[2] == self.n_d
        capi.check(self.L.pimc_propose(self.h, species, _vp(particle), _vp(b_first), newR.shape[1], _vp(newR)))

    def GetBeads(self, species, particle, b_first, n_beads):
        """Committed positions [clone][i][dim] of beads b_first[c]+i of particle[c]."""
        particle = np.ascontiguousarray(np.broadcast_to(particle, (self.n_clones,)), dtype=np.int32)
        b_first = np.ascontiguousarray(np.broadcast_to(b_first, (self.n_clones,)), dtype=np.int32)
        out = np.zeros((self.n_clones, n_beads, self.n_d))
        capi.check(self.L.pimc_beads_download(self.h, species, _vp(particle), _vp(b_first), n_beads, _vp(out)))
        return out

    def Commit(self, accept):
        accept = np.ascontiguousarray(np.broadcast_to(accept, (self.n_clones,)), dtype=np.int32)
        capi.check(self.L.pimc_commit(self.h, _vp(accept)))

    def SetMoveImages(self, species, n_images):
        """Bisect's n_images attribute (bisect_class.h:173) for later BisectSweep calls on `species`."""
        capi.check(self.L.pimc_move_set_images(self.h, species, int(n_images)))

    def BisectSweep(self, species, n_level, n_attempts, seed, attempt0=0, with_kinetic=True):
        """n_attempts device-resident Bisect::DoEvent calls per clone; returns accepts per clone."""
        n_accept = np.zeros(self.n_clones, dtype=np.int64)
        capi.check(self.L.pimc_bisect_sweep(self.h, species, n_level, n_attempts, seed, attempt0, 1 if with_kinetic else 0,
                                            _vp(n_accept)))
        return n_accept

    def BisectSweepWindows(self, species, n_level, n_rounds, seed, attempt0=0, with_kinetic=True):
        """n_rounds rounds of the bisection move on every disjoint window of every walker at once
        (pimc_bisect_sweep_windows); returns (accepts per clone, windows per walker and round)."""
        n_accept = np.zeros(self.n_clones, dtype=np.int64)
        n_win = C.c_int32()
        capi.check(self.L.pimc_bisect_sweep_windows(self.h, species, n_level, n_rounds, seed, attempt0, 1 if with_kinetic else 0,
                                                    _vp(n_accept), C.byref(n_win)))
        return n_accept, n_win.value

    def DisplaceSweep(self, species, step_size, n_attempts, seed, attempt0=0):
        """n_attempts device-resident DisplaceParticle::DoEvent calls per clone; returns accepts per clone."""
        n_accept = np.zeros(self.n_clones, dtype=np.int64)
        capi.check(self.L.pimc_displace_sweep(self.h, species, float(step_size), n_attempts, seed, attempt0, _vp(n_accept)))
        return n_accept

    def PermTable(self, species, b0, n_bisect_beads, epsilon=1e-100, relative=False):
        """PermBisectIterative::UpdatePermTable (relative: the PermBisectTable variant): t[clone][i][j]."""
        b0 = np.ascontiguousarray(np.broadcast_to(b0, (self.n_clones,)), dtype=np.int32)
        N = self.cfg.species[species].n_part
        t = np.zeros((self.n_clones, N, N))
        capi.check(self.L.pimc_perm_table(self.h, species, _vp(b0), n_bisect_beads, float(epsilon), 1 if relative else 0, _vp(t)))
        return t

    def PermBisectSweep(self, species, n_level, n_attempts, seed, attempt0=0, with_kinetic=True, epsilon=1e-100):
        """n_attempts device-resident PermBisectIterative::DoEvent calls per clone (perm_bisect_iterative_class.h:113-222).
        Returns (accepts per clone, perm_attempt[clone][8], perm_accept[clone][8]) -- the last two by cycle length - 1."""
        n_accept = np.zeros(self.n_clones, dtype=np.int64)
        att = np.zeros((self.n_clones, 8), dtype=np.int64)
        acc = np.zeros((self.n_clones, 8), dtype=np.int64)
        capi.check(self.L.pimc_perm_bisect_sweep(self.h, species, n_level, n_attempts, seed, attempt0, 1 if with_kinetic else 0,
                                                 float(epsilon), _vp(n_accept), _vp(att), _vp(acc)))
        return n_accept, att, acc

    def GetPermutation(self, species):
        """next[clone][p]: label of the bead that follows (p, n_bead - 1) -- the permutation at the beta seam."""
        nxt = np.zeros((self.n_clones, self.cfg.species[species].n_part), dtype=np.int32)
        capi.check(self.L.pimc_permutation_get(self.h, species, _vp(nxt)))
        return nxt

    def SetPermutation(self, species, nxt):
        nxt = np.ascontiguousarray(np.broadcast_to(nxt, (self.n_clones, self.cfg.species[species].n_part)), dtype=np.int32)
        capi.check(self.L.pimc_permutation_set(self.h, species, _vp(nxt)))

    def PermLastCycle(self):
        """Cycle of the last PermBisectSweep attempt per clone: dict(b0, n_perm, particles[clone][8], n_steps, accept)."""
        Cn = self.n_clones
        out = {k: np.zeros(Cn, dtype=np.int32) for k in ("b0", "n_perm", "n_steps", "accept")}
        out["particles"] = np.zeros((Cn, 8), dtype=np.int32)
        capi.check(self.L.pimc_perm_last_cycle(self.h, _vp(out["b0"]), _vp(out["n_perm"]), _vp(out["particles"]), _vp(out["n_steps"]),
                                               _vp(out["accept"])))
        return out

    def LaunchCount(self):
        return int(self.L.pimc_ctx_launch_count(self.h))

    def Sync(self):
        capi.check(self.L.pimc_ctx_sync(self.h))

    def SetTiming(self, enable=True):
        capi.check(self.L.pimc_ctx_set_timing(self.h, 1 if enable else 0))

    def KernelTime(self, kernel_id):
        """(total ms, launches) of one kernel family since timing was enabled."""
        ms, n = C.c_double(), C.c_int64()
        capi.check(self.L.pimc_ctx_kernel_time(self.h, kernel_id, C.byref(ms), C.byref(n)))
        return ms.value, n.value

    def ForceGeneral(self, enable=True):
        """Tests: evaluate with the general kernels even where the fast Ilkka path applies."""
        capi.check(self.L.pimc_ctx_force_general(self.h, 1 if enable else 0))

    def FastSqrt(self, x):
        x = np.ascontiguousarray(x, dtype=np.float64)
        out = np.zeros_like(x)
        capi.check(self.L.pimc_debug_fast_sqrt(self.h, len(x), _vp(x), _vp(out)))
        return out

    def Fp64Peak(self):
        t = C.c_double()
        capi.check(self.L.pimc_fp64_peak(self.h, C.byref(t)))
        return t.value


class Kinetic:
    """src/actions/single_action/kinetic_class.h on the device: the free-particle action of one
    species with `n_images` periodic images (FreeSpline, free_spline_class.h:25-84)."""

    def __init__(self, path, acfg):
        self.path = path
        self.L = path.L
        self.name = acfg.name
        self.type = "Kinetic"
        self.use_long_range = False
        self.max_level = acfg.max_level
        self.n_images = int(acfg.n_images)
        self.is_importance_weight = False
        self.species_a = self.species_b = path.cfg.species_index(acfg.species_a)
        h = C.c_void_p()
        capi.check(self.L.pimc_action_create_kinetic(path.h, self.species_a, self.n_images, C.byref(h)))
        self.h = h

    def _full(self, fn):
        out = np.zeros(self.path.n_clones)
        capi.check(fn(self.h, _vp(out)))
        return out

    def DActionDBeta(self):
        """kinetic_class.h:35-45: N M n_d / (2 tau) + sum over links of dlog rho_free / dtau."""
        return self._full(self.L.pimc_action_dbeta)

    def Potential(self):
        return self._full(self.L.pimc_action_potential)    # Action's default: 0 (action_class.h:43)

    def TotalAction(self):
        return self._full(self.L.pimc_action_total)

    def GetAction(self, b0, b1, particles, level):
        """kinetic_class.h:105-122: -sum over the listed particles of this species and the links of stride
        2^level in [b0, b1) of log rho_free; arguments as PairAction.GetAction."""
        C_ = self.path.n_clones
        b0a = np.ascontiguousarray(np.broadcast_to(b0, (C_,)), dtype=np.int32)
        n_window = int(np.broadcast_to(b1, (C_,))[0] - b0a[0])
        sp = np.ascontiguousarray([p[0] for p in particles], dtype=np.int32)
        pi = np.zeros((C_, len(particles)), dtype=np.int32)
        for i, p in enumerate(particles):
            pi[:, i] = np.broadcast_to(p[1], (C_,))
        out = np.zeros(C_)
        capi.check(self.L.pimc_action_get(self.h, self.path.mode, _vp(b0a), n_window, len(particles), _vp(sp), _vp(pi), level,
                                          _vp(out)))
        return out

    def VirialEnergy(self, virial_window_size=1):
        raise NotImplementedError("Kinetic::VirialEnergy (kinetic_class.h:48-102) is outside the device path")

    def ImportanceWeight(self):
        return np.ones(self.path.n_clones)

    def Accept(self):
        pass

    def Reject(self):
        pass


class PairAction:
    """One IlkkaPairAction / BarePairAction / DavidPairAction on the device."""

    def __init__(self, path, acfg):
        self.path = path
        self.L = path.L
        self.name = acfg.name
        self.type = acfg.type
        self.use_long_range = acfg.use_long_range
        self.max_level = acfg.max_level
        self.is_importance_weight = bool(getattr(acfg, "is_importance_weight", False))  # action_class.h:31
        cfg = path.cfg
        sa, sb = cfg.species_index(acfg.species_a), cfg.species_index(acfg.species_b)
        self.species_a, self.species_b = sa, sb
        kc = acfg.k_cut if acfg.k_cut is not None else (cfg.k_cut or 0.0)
        h = C.c_void_p()
        if acfg.type == "IlkkaPairAction":
            t, self._keep = capi.pack_ilkka(acfg.table, acfg.use_long_range)
            capi.check(self.L.pimc_action_create_ilkka(path.h, sa, sb, C.byref(t), acfg.max_level, int(acfg.use_long_range), kc,
                                                      C.byref(h)))
        elif acfg.type == "BarePairAction":
            t, self._keep = capi.pack_bare(acfg.table, acfg.use_long_range, acfg.is_coulomb)
            capi.check(self.L.pimc_action_create_bare(path.h, sa, sb, C.byref(t), acfg.max_level, int(acfg.use_long_range), kc,
                                                     C.byref(h)))
        elif acfg.type == "DavidPairAction":
            t, self._keep = capi.pack_david(acfg.table, acfg.n_order, acfg.use_long_range)
            capi.check(self.L.pimc_action_create_david(path.h, sa, sb, C.byref(t), acfg.max_level, int(acfg.use_long_range),
                                                      C.byref(h)))
        else:
            raise ValueError("ERROR: Unrecognized Action, %s" % acfg.type)  # actions.h:32
        self.h = h

    def _full(self, fn):
        out = np.zeros(self.path.n_clones)
        capi.check(fn(self.h, _vp(out)))
        return out

    def DActionDBeta(self):
        return self._full(self.L.pimc_action_dbeta)

    def Potential(self):
        return self._full(self.L.pimc_action_potential)

    def TotalAction(self):
        return self._full(self.L.pimc_action_total)

    def GetAction(self, b0, b1, particles, level):
        """particles: list of (species index, particle) with particle an int or an array of
        one index per clone; b0 an int or per-clone array; b1 - b0 the common window."""
        C_ = self.path.n_clones
        b0a = np.ascontiguousarray(np.broadcast_to(b0, (C_,)), dtype=np.int32)
        n_window = int(np.broadcast_to(b1, (C_,))[0] - b0a[0])
        sp = np.ascontiguousarray([p[0] for p in particles], dtype=np.int32)
        pi = np.zeros((C_, len(particles)), dtype=np.int32)
        for i, p in enumerate(particles):
            pi[:, i] = np.broadcast_to(p[1], (C_,))
        out = np.zeros(C_)
        capi.check(self.L.pimc_action_get(self.h, self.path.mode, _vp(b0a), n_window, len(particles), _vp(sp), _vp(pi), level,
                                          _vp(out)))
        return out

    def _window_args(self, b0, b1, particles):
        C_ = self.path.n_clones
        b0a = np.ascontiguousarray(np.broadcast_to(b0, (C_,)), dtype=np.int32)
        n_window = int(np.broadcast_to(b1, (C_,))[0] - b0a[0])
        sp = np.ascontiguousarray([p[0] for p in particles], dtype=np.int32)
        pi = np.zeros((C_, len(particles)), dtype=np.int32)
        for i, p in enumerate(particles):
            pi[:, i] = np.broadcast_to(p[1], (C_,))
        return b0a, n_window, sp, pi

    def GetActionGradient(self, b0, b1, particles, level):
        """Action::GetActionGradient (action_class.h:46): [clone][3]; arguments as GetAction."""
        b0a, n_window, sp, pi = self._window_args(b0, b1, particles)
        out = np.zeros((self.path.n_clones, 3))
        capi.check(self.L.pimc_action_gradient(self.h, _vp(b0a), n_window, len(particles), _vp(sp), _vp(pi), level, _vp(out)))
        return out

    def GetActionLaplacian(self, b0, b1, particles, level):
        """Action::GetActionLaplacian (action_class.h:49): one value per clone."""
        b0a, n_window, sp, pi = self._window_args(b0, b1, particles)
        out = np.zeros(self.path.n_clones)
        capi.check(self.L.pimc_action_laplacian(self.h, _vp(b0a), n_window, len(particles), _vp(sp), _vp(pi), level, _vp(out)))
        return out

    def VirialEnergy(self, virial_window_size=1):
        """Action::VirialEnergy (action_class.h:40): DActionDBeta for every pair action."""
        return self.DActionDBeta()

    def ImportanceWeight(self):
        """PairAction::ImportanceWeight (pair_action_class.h:398-400)."""
        if not self.is_importance_weight:
            return np.ones(self.path.n_clones)
        return np.exp(self.DActionDBeta() / self.path.n_bead)

    def Accept(self):
        capi.check(self.L.pimc_action_accept(self.h))

    def Reject(self):
        capi.check(self.L.pimc_action_reject(self.h))

    def CalcPair(self, which, r, r_p, s, level=0):
        r = np.ascontiguousarray(r, dtype=np.float64)
        r_p = np.ascontiguousarray(r_p, dtype=np.float64)
        s = np.ascontiguousarray(s, dtype=np.float64)
        out = np.zeros_like(r)
        capi.check(self.L.pimc_action_calc_pair(self.h, which, len(r), _vp(r), _vp(r_p), _vp(s), level, _vp(out)))
        return out


    def CalcPairFast(self, which, r, r_p, s):
        r = np.ascontiguousarray(r, dtype=np.float64)
        r_p = np.ascontiguousarray(r_p, dtype=np.float64)
        s = np.ascontiguousarray(s, dtype=np.float64)
        out = np.zeros_like(r)
        capi.check(self.L.pimc_action_calc_pair_fast(self.h, which, len(r), _vp(r), _vp(r_p), _vp(s), _vp(out)))
        return out


class PairCorrelation:
    """src/events/observables/pair_correlation_class.h."""

    def __init__(self, path, species_a, species_b, r_min=0.0, r_max=None, n_r=100):
        self.path = path
        self.sa, self.sb = species_a, species_b
        self.r_min = r_min
        self.r_max = path.cfg.L / 2.0 if r_max is None else r_max
        self.n_r = int(n_r)
        self.Reset()

    def Reset(self):
        self.n_measure = 0
        self.y = np.zeros((self.path.n_clones, self.n_r))

    def Accumulate(self, cofactor=None):
        cf = None if cofactor is None else np.ascontiguousarray(cofactor, dtype=np.float64)
        capi.check(self.path.L.pimc_est_gofr(self.path.h, self.sa, self.sb, self.r_min, self.r_max, self.n_r,
                                             None if cf is None else _vp(cf), _vp(self.y)))
        self.n_measure += 1

    def Counts(self):
        counts = np.zeros((self.path.n_clones, self.n_r), dtype=np.uint64)
        capi.check(self.path.L.pimc_est_gofr_counts(self.path.h, self.sa, self.sb, self.r_min, self.r_max, self.n_r, _vp(counts)))
        return counts

    def Write(self):
        """Normalised g(r) per clone (pair_correlation_class.h:89-123), then reset."""
        cfg = self.path.cfg
        Na, Nb = cfg.species[self.sa].n_part, cfg.species[self.sb].n_part
        vol = cfg.L ** cfg.n_d if cfg.pbc else 1.0
        if self.sa == self.sb:
            norm = 0.5 * self.n_measure * Na * (Nb - 1) * cfg.n_bead / vol
        else:
            norm = self.n_measure * Na * Nb * cfg.n_bead / vol
        dr = (self.r_max - self.r_min) / (self.n_r - 1.0)
        x = self.r_min + np.arange(self.n_r) * dr
        r1 = x
        r2 = np.concatenate([x[1:], [2.0 * x[-1] - x[-2]]])
        bin_vol = 4.0 * math.pi / 3.0 * (r2 ** 3 - r1 ** 3)
        g = self.y / (bin_vol * norm)
        self.Reset()
        return g


class StructureFactor:
    """src/events/observables/structure_factor_class.h."""

    def __init__(self, path, species_a, species_b, k_cut=None):
        self.path = path
        self.sa, self.sb = species_a, species_b
        self.k_cut = path.cfg.k_cut if k_cut is None else k_cut
        path.SetupKSpace(self.k_cut)
        self.Reset()

    def Reset(self):
        self.n_measure = 0
        self.sk = np.zeros((self.path.n_clones, self.path._n_k()))

    def Accumulate(self, cofactor=None):
        cf = None if cofactor is None else np.ascontiguousarray(cofactor, dtype=np.float64)
        capi.check(self.path.L.pimc_est_sofk(self.path.h, self.sa, self.sb, self.k_cut, None if cf is None else _vp(cf),
                                             _vp(self.sk)))
        self.n_measure += 1

    def Write(self):
        cfg = self.path.cfg
        norm = self.n_measure * cfg.n_bead * cfg.species[self.sa].n_part * cfg.species[self.sb].n_part
        out = self.sk / norm
        self.Reset()
        return out


class IO:
    """scaffold::io::IO (include/scaffold/io/io_hdf5.h:12-212) for the walkers of one context: Write / Rewrite /
    CreateExtendableDataSet / AppendDataSet with the reference's dataset names, one HDF5 file per walker
    (`<prefix>.<clone>.h5`, as the reference's ranks write `<output_prefix>.<rank>.h5`).  An extendable dataset is
    the reference's (n_records, shape(data)...) array -- appended records along a new leading axis -- written
    contiguously when the file is saved (simpimc_b200.h5lite; no HDF5 library in this build)."""

    def __init__(self, prefix, n_clones):
        self.prefix, self.n_clones = prefix, n_clones
        self.fixed = [dict() for _ in range(n_clones)]
        self.series = [dict() for _ in range(n_clones)]

    @staticmethod
    def _key(prefix, name=""):
        return "/".join(p for p in (prefix + name).split("/") if p)

    def Write(self, name, value):
        """One value for every walker, or an array with a leading clone axis (per_clone=True values come from
        estimators)."""
        for c in range(self.n_clones):
            self.fixed[c][self._key(name)] = value

    Rewrite = Write

    def WritePerClone(self, name, values):
        for c in range(self.n_clones):
            self.fixed[c][self._key(name)] = values[c]

    def CreateExtendableDataSet(self, prefix, name, data):
        for c in range(self.n_clones):
            self.series[c][self._key(prefix, name)] = [np.asarray(data[c])]

    def AppendDataSet(self, prefix, name, data):
        for c in range(self.n_clones):
            self.series[c].setdefault(self._key(prefix, name), []).append(np.asarray(data[c]))

    def FileName(self, clone):
        return "%s.%d.h5" % (self.prefix, clone)

    def Save(self):
        from . import h5lite
        for c in range(self.n_clones):
            out = dict(self.fixed[c])
            for k, recs in self.series[c].items():
                out[k] = np.stack(recs)
            h5lite.write(self.FileName(c), out)
        return [self.FileName(c) for c in range(self.n_clones)]


class Energy:
    """Thermal and potential estimators of src/events/observables/energy_class.h:22-33,119-161."""

    def __init__(self, path, measure_potential=False):
        self.path = path
        self.measure_potential = measure_potential
        self.actions = list(path.actions)   # energy_class.h:22-33: every action, Kinetic included
        self.Reset()

    def Reset(self):
        self.n_measure = 0
        self.energies = np.zeros((len(self.actions), self.path.n_clones))
        self.potentials = np.zeros((len(self.actions), self.path.n_clones))

    def Accumulate(self, cofactor=1.0):
        for i, a in enumerate(self.actions):
            self.energies[i] += cofactor * a.DActionDBeta()
            if self.measure_potential:
                self.potentials[i] += cofactor * a.Potential()
        self.n_measure += 1

    def Write(self, out=None, name="Energy"):
        """energy_class.h:232-276: block means per action and their sum.  With `out` (an IO) the block is appended to
        the reference's datasets Observables/<name>/{total,<action>,v_total,v_<action>}/x."""
        norm = self.path.cfg.n_bead * self.n_measure  # energy_class.h:234
        e, v = self.energies / norm, self.potentials / norm
        if out is not None and self.n_measure > 0:
            prefix = "Observables/%s/" % name
            first = not getattr(self, "_written", False)
            put = out.CreateExtendableDataSet if first else out.AppendDataSet
            groups = [("total", e.sum(axis=0))] + [(a.name, e[i]) for i, a in enumerate(self.actions)]
            if self.measure_potential:
                groups += [("v_total", v.sum(axis=0))] + [("v_" + a.name, v[i]) for i, a in enumerate(self.actions)]
            if first:
                out.Write(prefix + "type", "Energy")
                out.Write(prefix + "data_type", "scalar")
            for g, x in groups:
                put("/" + prefix + g + "/", "x", x)
                if first:
                    out.Write(prefix + g + "/data_type", "scalar")
            self._written = True
        self.Reset()
        return e, v


class PathDump:
    """src/events/observables/path_dump_class.h:27-68 and the "Restart" branch of Species::InitPaths
    (species_class.h:336-378): every Write() appends each species' positions -- the reference's
    cube (n_d, n_bead, n_part) is [n_part][n_bead][n_d] in file order, the order used here, with a
    leading clone axis -- and the permutation table (previous particle of bead 0, next particle
    of the last bead: the context's permutation at the beta seam, Path.GetPermutation).  The container
    is a flat .npz with the reference's dataset names as keys, or the reference's HDF5 layout through IO."""

    def __init__(self, path, name="path_dump", skip=1):
        self.path, self.name, self.skip = path, name, max(1, int(skip))
        self.n_dump, self.n_write_calls = 0, 0
        self.positions = {s.name: [] for s in path.cfg.species}
        self.permutations = {s.name: [] for s in path.cfg.species}

    @staticmethod
    def _perm_table(nxt):
        """[clone][n_part][2]: (particle of the bead before (p, 0), particle of the bead after (p, n_bead - 1))."""
        prev = np.empty_like(nxt)
        for c in range(nxt.shape[0]):
            prev[c, nxt[c]] = np.arange(nxt.shape[1])
        return np.stack([prev, nxt], axis=-1).astype(np.float64)

    def Write(self, out=None):
        """path_dump_class.h:29-68.  With `out` (an IO) the dump is appended to the reference's datasets
        Observables/<name>/<species>/{n_dump, positions (n_dump, n_part, n_bead, n_d), permutation (n_dump, n_part, 2)}."""
        if self.n_write_calls % self.skip == 0:
            self.n_dump += 1
            for si, s in enumerate(self.path.cfg.species):
                R = self.path.GetPositions(si)
                perm = self._perm_table(self.path.GetPermutation(si))
                self.positions[s.name].append(R)
                self.permutations[s.name].append(perm)
                if out is not None:
                    prefix = "Observables/%s/%s/" % (self.name, s.name)
                    put = out.CreateExtendableDataSet if self.n_dump == 1 else out.AppendDataSet
                    out.Write(prefix + "n_dump", np.uint32(self.n_dump))
                    put(prefix, "positions", R)
                    put(prefix, "permutation", perm)
        self.n_write_calls += 1

    def Save(self, file_name):
        out = {}
        for s in self.path.cfg.species:
            key = "Observables/%s/%s/" % (self.name, s.name)
            out[key + "n_dump"] = np.int64(self.n_dump)
            out[key + "positions"] = np.stack(self.positions[s.name]) if self.positions[s.name] else np.zeros((0,))
            out[key + "permutation"] = (np.stack(self.permutations[s.name]) if self.permutations[s.name]
                                        else np.zeros((0,)))                                  # [dump][clone][n_part][2]
        np.savez_compressed(file_name, **out)

    @staticmethod
    def Restart(path, file_name, name="path_dump"):
        """init_type="Restart": the LAST dump of every species becomes the configuration.  file_name: the .npz of
        Save(), or the reference-format HDF5 files of IO.Save() -- one per walker (a list), or one for all."""
        if isinstance(file_name, (list, tuple)) or str(file_name).endswith(".h5"):
            from . import h5lite
            files = list(file_name) if isinstance(file_name, (list, tuple)) else [file_name] * path.n_clones
            if len(files) != path.n_clones:
                raise ValueError("one restart file per walker is needed")
            dumps = [h5lite.read(fn) for fn in files]
            for si, s in enumerate(path.cfg.species):
                key = "Observables/%s/%s/" % (name, s.name)
                path.SetPositions(si, np.stack([d[key + "positions"][-1] for d in dumps]))
                PathDump._restore_permutation(path, si, np.stack([d[key + "permutation"][-1] for d in dumps]))
            return
        f = np.load(file_name)
        for si, s in enumerate(path.cfg.species):
            key = "Observables/%s/%s/" % (name, s.name)
            path.SetPositions(si, f[key + "positions"][-1])
            PathDump._restore_permutation(path, si, f[key + "permutation"][-1])

    @staticmethod
    def _restore_permutation(path, si, perm):
        """species_class.h:368-377: column 1 is the particle that follows each particle's last bead (column 0 its inverse).
        perm: [clone][n_part][2] or, from files written before permutations were tracked, [n_part][2]."""
        perm = np.asarray(perm)
        if perm.ndim == 2:
            perm = np.broadcast_to(perm, (path.n_clones,) + perm.shape)
        nxt = np.rint(perm[..., 1]).astype(np.int32)
        n_part = nxt.shape[-1]
        if np.array_equal(nxt, np.broadcast_to(np.arange(n_part, dtype=np.int32), nxt.shape)):
            return      # unpermuted: leave the context untracked (every move stays available)
        path.SetPermutation(si, nxt)
