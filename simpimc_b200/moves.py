"""Host-side mirror of the two moves that define the action-delta contract, vectorised over
the clones of a `host.Path`:

* `Bisect`            src/events/moves/single_species_move/bisect/bisect_class.h:24-139
* `DisplaceParticle`  src/events/moves/single_species_move/displace_particle_class.h:13-89

Same call sequence as the reference: pick particle and window, Levy-sample the midpoints
level by level (NEW mode), evaluate every action in OLD then NEW mode, Metropolis-test the
difference, then Accept (StoreR / StoreRhoK) or Reject.  The pair actions run on the GPU
through the C ABI; the free-particle (kinetic) action and the Levy sampling probabilities are
evaluated here with the host mirror of FreeSpline (free_spline.py; with n_images = 0 the closed
form -|r|^2 / 4 lambda tau that src/actions/single_action/kinetic_class.h:105-122 reduces to).
Random numbers come from numpy's Generator, not std::mt19937: sampled runs agree with the
reference statistically, not stream for stream.
"""
import math

import numpy as np

from . import host
from .free_spline import FreeSpline


def _put_in_box(d, L, pbc):
    """Path::PutInBox (path_class.h:131-134); nearbyint == numpy rint (half to even)."""
    return d - np.rint(d / L) * L if pbc else d


class _Move:
    def __init__(self, path, rng, species, with_kinetic=True):
        self.path = path
        self.rng = rng
        self.species = species
        self.cfg = path.cfg
        self.lam = self.cfg.species[species].lam
        self.with_kinetic = with_kinetic and self.lam > 0.0
        # move_class.h:27-31: the actions that involve this species (the Kinetic action is evaluated by the
        # host mirror below, with the n_images of the path's Kinetic object when there is one)
        self.action_list = [a for a in path.actions if a.type != "Kinetic" and species in (a.species_a, a.species_b)]
        kin = [a for a in path.actions if a.type == "Kinetic" and a.species_a == species]
        self.kinetic_images = kin[0].n_images if kin else 0
        self.n_attempt = 0
        self.n_accept = np.zeros(path.n_clones, dtype=np.int64)

    def _dr(self, a, b):
        return _put_in_box(a - b, self.cfg.L, self.cfg.pbc)

    def _kinetic_log_rho(self, dr, level_tau, n_images=0):
        # FreeSpline::GetLogRhoFree; with no images: -|r|^2 / (4 lambda tau)
        if n_images:
            return _free_spline(self.cfg, self.lam, n_images, level_tau).GetLogRhoFree(dr)
        return -np.sum(dr * dr, axis=-1) / (4.0 * self.lam * level_tau)

    def accept_ratio(self):
        return self.n_accept.sum() / max(1, self.n_attempt * self.path.n_clones)


class Bisect(_Move):
    def __init__(self, path, rng, species, n_level, with_kinetic=True, n_images=0):
        super().__init__(path, rng, species, with_kinetic)
        self.n_images = int(n_images)   # bisect_class.h:173
        self.n_level = n_level
        self.n_bisect_beads = 1 << n_level

    def DoEvent(self):
        """One Move::DoEvent (move_class.h:61-77) for every clone at once."""
        path, cfg, C = self.path, self.cfg, self.path.n_clones
        N, M, nb = cfg.species[self.species].n_part, cfg.n_bead, self.n_bisect_beads
        tau = cfg.tau
        self.n_attempt += 1
        p_i = self.rng.integers(0, N, C).astype(np.int32)
        bead0 = self.rng.integers(0, M, C).astype(np.int32)
        # beads bead0 .. bead0+nb of the particle (labels)
        old = path.GetBeads(self.species, p_i, bead0, nb + 1)
        new = old.copy()
        alive = np.ones(C, dtype=bool)          # clones not yet rejected at a coarser level
        prev_change = np.zeros(C)
        particles = [(self.species, p_i)]
        for level in range(self.n_level - 1, -1, -1):
            skip = 1 << level
            level_tau = tau * skip
            sigma = math.sqrt(self.lam * level_tau)
            old_lp = np.zeros(C)
            new_lp = np.zeros(C)
            for a in range(0, nb, 2 * skip):
                b, c = a + skip, a + 2 * skip
                # RBar(bead_c, bead_a) = r_a + 0.5 * Dr(r_c, r_a)   (path_class.h:124)
                rbar_old = old[:, a] + 0.5 * self._dr(old[:, c], old[:, a])
                old_lp += self._kinetic_log_rho(self._dr(old[:, b], rbar_old), 0.5 * level_tau, self.n_images)
                rbar_new = new[:, a] + 0.5 * self._dr(new[:, c], new[:, a])
                delta = _put_in_box(self.rng.normal(0.0, sigma, (C, cfg.n_d)), cfg.L, cfg.pbc)
                new[:, b] = rbar_new + delta
                new_lp += self._kinetic_log_rho(delta, 0.5 * level_tau, self.n_images)
            old_action = np.zeros(C)
            new_action = np.zeros(C)
            if self.with_kinetic:  # Kinetic::GetAction (kinetic_class.h:105-122)
                for a in range(0, nb, skip):
                    old_action -= self._kinetic_log_rho(self._dr(old[:, a], old[:, a + skip]), level_tau, self.kinetic_images)
                    new_action -= self._kinetic_log_rho(self._dr(new[:, a], new[:, a + skip]), level_tau, self.kinetic_images)
            if level == 0 and self.action_list:
                # pair actions return 0 above max_level = 0 (pair_action_class.h:269)
                path.Propose(self.species, p_i, (bead0 + 1) % M, new[:, 1:nb])
                for act in self.action_list:
                    path.SetMode(host.OLD_MODE)
                    old_action += act.GetAction(bead0, bead0 + nb, particles, 0)
                    path.SetMode(host.NEW_MODE)
                    new_action += act.GetAction(bead0, bead0 + nb, particles, 0)
            log_sample_ratio = -new_lp + old_lp
            change = new_action - old_action
            log_accept = log_sample_ratio - change + prev_change
            alive &= ~(log_accept < np.log(self.rng.random(C)))
            prev_change = change
        if not self.action_list:
            path.Propose(self.species, p_i, (bead0 + 1) % M, new[:, 1:nb])
        path.Commit(alive.astype(np.int32))
        self.n_accept += alive
        return alive


class DisplaceParticle(_Move):
    def __init__(self, path, rng, species, step_size=None):
        super().__init__(path, rng, species, with_kinetic=False)  # a rigid shift leaves the springs unchanged
        self.step_size = self.cfg.L / 10.0 if step_size is None else step_size

    def DoEvent(self):
        path, cfg, C = self.path, self.cfg, self.path.n_clones
        N, M = cfg.species[self.species].n_part, cfg.n_bead
        self.n_attempt += 1
        p_i = self.rng.integers(0, N, C).astype(np.int32)
        # RNG::UnifRand(vec, l): uniform in the cube, normalised to length l (rng.h:43-56)
        dr = self.rng.uniform(-1.0, 1.0, (C, cfg.n_d))
        dr *= self.step_size / np.linalg.norm(dr, axis=1, keepdims=True)
        old = path.GetBeads(self.species, p_i, np.zeros(C, dtype=np.int32), M)
        path.Propose(self.species, p_i, np.zeros(C, dtype=np.int32), old + dr[:, None, :])
        particles = [(self.species, p_i)]
        old_action = np.zeros(C)
        new_action = np.zeros(C)
        for act in self.action_list:
            path.SetMode(host.OLD_MODE)
            old_action += act.GetAction(0, M, particles, 0)
            path.SetMode(host.NEW_MODE)
            new_action += act.GetAction(0, M, particles, 0)
        accept = ~((old_action - new_action) < np.log(self.rng.random(C)))
        path.Commit(accept.astype(np.int32))
        self.n_accept += accept
        return accept


_FS_CACHE = {}


def _free_spline(cfg, lam, n_images, tau_s):
    key = (cfg.L if cfg.pbc else 0.0, int(n_images), lam, tau_s)
    if key not in _FS_CACHE:
        _FS_CACHE[key] = FreeSpline(key[0], n_images, lam, tau_s)
    return _FS_CACHE[key]


def bisect_attempt_philox(cfg, species, n_level, seed, attempt, n_clones, get_beads, action_old_new, finish, with_kinetic=True,
                          b0_range=None, n_images_move=0, n_images_kin=0, windows=None):
    """Host mirror of ONE device-resident bisection attempt (csrc/mc.cuh: bisect_sample_kernel +
    pair_window_both_kernel + k-sums + bisect_decide_kernel) drawing the same Philox stream.

    get_beads(c, p, bead0, n)            -> committed positions [n][3] of beads bead0.. of particle p
    action_old_new(c, p, bead0, nb, new) -> (old_action, new_action) summed over the pair actions
                                            that involve `species`, `new` = proposed beads 1..nb-1
    finish(c, p, bead0, nb, accept)      -> Move::Accept / Reject
    b0_range = (first, count): window starts uniform in [first, first + count) -- a slice shard's
    interior windows (pimc_bisect_sweep on a sharded context); default: the whole path.
    windows = W: the multi-window mode of pimc_bisect_sweep_windows -- n_clones counts VIRTUAL clones (walker * W +
    window, the index the callbacks receive and the stream is keyed by); window w of a walker starts at first + offset
    + w 2^n_level with the offset drawn from the stream of the walker's window 0 (b0_range = (first, n_offsets)).
    n_images_move / n_images_kin: periodic images of Bisect's sampling splines (bisect_class.h:158-163,173)
    and of the Kinetic action (kinetic_class.h:16-24), evaluated with the FreeSpline host mirror.
    Returns (particle[c], bead0[c], accept[c]).
    """
    from . import philox as PX
    sp = cfg.species[species]
    N, M, lam, tau, L, pbc = sp.n_part, cfg.n_bead, sp.lam, cfg.tau, cfg.L, cfg.pbc
    nb = 1 << n_level
    k0, k1 = seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF
    a_lo, a_hi = attempt & 0xFFFFFFFF, (attempt >> 32) & 0xFFFFFFFF
    out_p, out_b, out_acc = [], [], []

    def pib(d):
        return _put_in_box(d, L, pbc)

    for c in range(n_clones):
        r = PX.philox4x32(a_lo, a_hi, c, 0, k0, k1)
        p_i = min(int(PX.uniform_from_bits(r[0], r[1]) * N), N - 1)
        b_first, b_count = (0, M) if b0_range is None else b0_range
        bead0 = b_first + min(int(PX.uniform_from_bits(r[2], r[3]) * b_count), b_count - 1)
        if windows:
            c0 = (c // windows) * windows
            r_w = PX.philox4x32(a_lo, a_hi, c0, 0, k0, k1)
            off = min(int(PX.uniform_from_bits(r_w[2], r_w[3]) * b_count), b_count - 1)
            bead0 = b_first + off + (c - c0) * nb
            if b0_range is None or bead0 >= M:
                bead0 %= M
        old = np.array(get_beads(c, p_i, bead0, nb + 1), dtype=np.float64)
        new = old.copy()
        slot, alive, prev_change, partial, logu0 = 1, True, 0.0, 0.0, 0.0
        for level in range(n_level - 1, -1, -1):
            skip = 1 << level
            level_tau = tau * skip
            sigma = math.sqrt(lam * level_tau)
            i4s, i4k = 1.0 / (4.0 * lam * (0.5 * level_tau)), 1.0 / (4.0 * lam * level_tau)
            old_lp = new_lp = 0.0
            for ia in range(0, nb, 2 * skip):
                ib, ic = ia + skip, ia + 2 * skip
                r0 = PX.philox4x32(a_lo, a_hi, c, slot, k0, k1)
                r1 = PX.philox4x32(a_lo, a_hi, c, slot + 1, k0, k1)
                slot += 2
                ua, ub = PX.uniform_from_bits(r0[0], r0[1]), PX.uniform_from_bits(r0[2], r0[3])
                uc, ud = PX.uniform_from_bits(r1[0], r1[1]), PX.uniform_from_bits(r1[2], r1[3])
                ra, rc = math.sqrt(-2.0 * math.log(ua)), math.sqrt(-2.0 * math.log(uc))
                nrm = np.array([ra * math.cos(2 * math.pi * ub), ra * math.sin(2 * math.pi * ub), rc * math.cos(2 * math.pi * ud)])
                rbar_old = old[ia] + 0.5 * pib(old[ic] - old[ia])
                del_old = pib(old[ib] - rbar_old)
                rbar_new = new[ia] + 0.5 * pib(new[ic] - new[ia])
                del_new = pib(sigma * nrm)
                new[ib] = rbar_new + del_new
                if n_images_move:
                    fs = _free_spline(cfg, lam, n_images_move, 0.5 * level_tau)
                    old_lp += float(fs.GetLogRhoFree(del_old))
                    new_lp += float(fs.GetLogRhoFree(del_new))
                else:
                    old_lp -= float(np.sum(del_old * del_old)) * i4s
                    new_lp -= float(np.sum(del_new * del_new)) * i4s
            old_kin = new_kin = 0.0
            if with_kinetic:
                fk = _free_spline(cfg, lam, n_images_kin, level_tau) if n_images_kin else None
                for ia in range(0, nb, skip):
                    o, n_ = pib(old[ia] - old[ia + skip]), pib(new[ia] - new[ia + skip])
                    if fk is not None:
                        old_kin -= float(fk.GetLogRhoFree(o))
                        new_kin -= float(fk.GetLogRhoFree(n_))
                    else:
                        old_kin += float(np.sum(o * o)) * i4k
                        new_kin += float(np.sum(n_ * n_)) * i4k
            ru = PX.philox4x32(a_lo, a_hi, c, slot, k0, k1)
            slot += 1
            logu = math.log(PX.uniform_from_bits(ru[0], ru[1]))
            lsr, change = -new_lp + old_lp, new_kin - old_kin
            if level > 0:
                if lsr - change + prev_change < logu:
                    alive = False
                prev_change = change
            else:
                partial, logu0 = lsr - change + prev_change, logu
        accept = False
        if alive:
            old_a, new_a = action_old_new(c, p_i, bead0, nb, new[1:nb])
            accept = not (partial - (new_a - old_a) < logu0)
        finish(c, p_i, bead0, nb, accept, new[1:nb] if alive else None)
        out_p.append(p_i)
        out_b.append(bead0)
        out_acc.append(accept)
    return np.array(out_p), np.array(out_b), np.array(out_acc)


def bisect_philox_numbers(cfg, species, n_level, seed, attempt, clone):
    """The random numbers ONE device-resident bisection attempt of `clone` draws from its Philox stream, in the order
    the reference's Bisect::Attempt consumes its own (bisect_class.h:39-125): uniforms = [particle, first bead, then
    the Metropolis uniform of every level from the top]; normals = the three Levy-displacement normals of every
    midpoint, level by level from the top.  Feeding them to the reference program (oracle/refsim.py: inject_random)
    makes it take the very decisions the device takes."""
    from . import philox as PX
    nb = 1 << n_level
    k0, k1 = seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF
    a_lo, a_hi = attempt & 0xFFFFFFFF, (attempt >> 32) & 0xFFFFFFFF
    r = PX.philox4x32(a_lo, a_hi, clone, 0, k0, k1)
    uniforms = [PX.uniform_from_bits(r[0], r[1]), PX.uniform_from_bits(r[2], r[3])]
    normals = []
    slot = 1
    for level in range(n_level - 1, -1, -1):
        skip = 1 << level
        for _ in range(0, nb, 2 * skip):
            r0 = PX.philox4x32(a_lo, a_hi, clone, slot, k0, k1)
            r1 = PX.philox4x32(a_lo, a_hi, clone, slot + 1, k0, k1)
            slot += 2
            ua, ub = PX.uniform_from_bits(r0[0], r0[1]), PX.uniform_from_bits(r0[2], r0[3])
            uc, ud = PX.uniform_from_bits(r1[0], r1[1]), PX.uniform_from_bits(r1[2], r1[3])
            ra, rc = math.sqrt(-2.0 * math.log(ua)), math.sqrt(-2.0 * math.log(uc))
            normals += [ra * math.cos(2 * math.pi * ub), ra * math.sin(2 * math.pi * ub), rc * math.cos(2 * math.pi * ud)]
        ru = PX.philox4x32(a_lo, a_hi, clone, slot, k0, k1)
        slot += 1
        uniforms.append(PX.uniform_from_bits(ru[0], ru[1]))
    return uniforms, normals


def displace_philox_numbers(seed, attempt, clone):
    """The uniforms ONE device-resident DisplaceParticle attempt draws, in the reference's order of consumption
    (displace_particle_class.h:28-71): particle, three components of the direction, Metropolis uniform."""
    from . import philox as PX
    k0, k1 = seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF
    a_lo, a_hi = attempt & 0xFFFFFFFF, (attempt >> 32) & 0xFFFFFFFF
    r = [PX.philox4x32(a_lo, a_hi, clone, slot, k0, k1) for slot in range(4)]
    return [PX.uniform_from_bits(r[0][0], r[0][1]), PX.uniform_from_bits(r[1][0], r[1][1]), PX.uniform_from_bits(r[1][2], r[1][3]),
            PX.uniform_from_bits(r[2][0], r[2][1]), PX.uniform_from_bits(r[3][0], r[3][1])]


def displace_attempt_philox(cfg, species, step_size, seed, attempt, n_clones, get_beads, action_old_new, finish):
    """Host mirror of ONE device-resident DisplaceParticle attempt (csrc/displace.cuh) drawing the
    same Philox stream: slot 0 particle, slots 1-2 direction, slot 3 Metropolis uniform.

    get_beads(c, p, 0, M) -> committed path [M][3]; action_old_new(c, p, new) -> (old, new) action
    over the whole path summed over the pair actions that involve `species`; finish(c, p, accept).
    Returns (particle[c], accept[c])."""
    from . import philox as PX
    N, M = cfg.species[species].n_part, cfg.n_bead
    k0, k1 = seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF
    a_lo, a_hi = attempt & 0xFFFFFFFF, (attempt >> 32) & 0xFFFFFFFF
    out_p, out_acc = [], []
    for c in range(n_clones):
        r0 = PX.philox4x32(a_lo, a_hi, c, 0, k0, k1)
        r1 = PX.philox4x32(a_lo, a_hi, c, 1, k0, k1)
        r2 = PX.philox4x32(a_lo, a_hi, c, 2, k0, k1)
        r3 = PX.philox4x32(a_lo, a_hi, c, 3, k0, k1)
        p_i = min(int(PX.uniform_from_bits(r0[0], r0[1]) * N), N - 1)
        # RNG::UnifRand(vec, l): (b - a) * u + a per component, normalised, times l (rng.h:31-56)
        v = np.array([2.0 * PX.uniform_from_bits(r1[0], r1[1]) + -1.0, 2.0 * PX.uniform_from_bits(r1[2], r1[3]) + -1.0,
                      2.0 * PX.uniform_from_bits(r2[0], r2[1]) + -1.0])
        m = math.sqrt((v[0] * v[0] + v[2] * v[2]) + v[1] * v[1])
        dr = (v / m) * step_size
        logu = math.log(PX.uniform_from_bits(r3[0], r3[1]))
        old = np.array(get_beads(c, p_i, 0, M), dtype=np.float64)
        new = old + dr[None, :]
        old_action, new_action = action_old_new(c, p_i, new)
        accept = not ((old_action - new_action) < logu)
        finish(c, p_i, accept)
        out_p.append(p_i)
        out_acc.append(1 if accept else 0)
    return np.array(out_p), np.array(out_acc)
