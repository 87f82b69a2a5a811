"""Host-side mirror of the two moves that define the action-delta contract, vectorised over
the clones of a `host.Path`:

* `Bisect`            src/events/moves/single_species_move/bisect/bisect_class.h:24-139
* `DisplaceParticle`  src/events/moves/single_species_move/displace_particle_class.h:13-89

Same call sequence as the reference: pick particle and window, Levy-sample the midpoints
level by level (NEW mode), evaluate every action in OLD then NEW mode, Metropolis-test the
difference, then Accept (StoreR / StoreRhoK) or Reject.  The pair actions run on the GPU
through the C ABI; the free-particle (kinetic) action is evaluated here in closed form,
which is what src/actions/single_action/kinetic_class.h:105-122 computes when n_images = 0
(FreeSpline's image sum is then identically zero, free_spline_class.h:47-62).
Random numbers come from numpy's Generator, not std::mt19937: sampled runs agree with the
reference statistically, not stream for stream.
"""
import math

import numpy as np

from . import host


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
        # move_class.h:27-31: the actions that involve this species
        self.action_list = [a for a in path.actions if a is not None and species in (a.species_a, a.species_b)]
        self.n_attempt = 0
        self.n_accept = np.zeros(path.n_clones, dtype=np.int64)

    def _dr(self, a, b):
        return _put_in_box(a - b, self.cfg.L, self.cfg.pbc)

    def _kinetic_log_rho(self, dr, level_tau):
        # FreeSpline::GetLogRhoFree with no images: -|r|^2 / (4 lambda tau)
        return -np.sum(dr * dr, axis=-1) / (4.0 * self.lam * level_tau)

    def accept_ratio(self):
        return self.n_accept.sum() / max(1, self.n_attempt * self.path.n_clones)


class Bisect(_Move):
    def __init__(self, path, rng, species, n_level, with_kinetic=True):
        super().__init__(path, rng, species, with_kinetic)
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
                old_lp += self._kinetic_log_rho(self._dr(old[:, b], rbar_old), 0.5 * level_tau)
                rbar_new = new[:, a] + 0.5 * self._dr(new[:, c], new[:, a])
                delta = _put_in_box(self.rng.normal(0.0, sigma, (C, cfg.n_d)), cfg.L, cfg.pbc)
                new[:, b] = rbar_new + delta
                new_lp += self._kinetic_log_rho(delta, 0.5 * level_tau)
            old_action = np.zeros(C)
            new_action = np.zeros(C)
            if self.with_kinetic:  # Kinetic::GetAction (kinetic_class.h:105-122)
                for a in range(0, nb, skip):
                    old_action -= self._kinetic_log_rho(self._dr(old[:, a], old[:, a + skip]), level_tau)
                    new_action -= self._kinetic_log_rho(self._dr(new[:, a], new[:, a + skip]), level_tau)
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
