"""Host mirror of the permuting bisection move
(src/events/moves/single_species_move/bisect/perm_bisect/perm_bisect_iterative_class.h:10-222 on
perm_bisect_class.h:32-82) in the dense representation of SURVEY App. A-4: positions by particle
LABEL R[label][slice] plus the permutation at the beta seam, next[p] = label of the bead that follows
(p, n_bead - 1).  Inside the path a chain keeps its label; across the seam it continues as next[label].

    select_cycle        SelectCycleIterative (:33-111): row p of the table t_ij = exp(-|Dr(r_i(b0), r_j(b1))|^2 /
                        4 lambda tau n_bisect_beads) (UpdatePermTable, :10-30, chains followed across the seam),
                        "continue?" and "which next particle?" from two uniforms per step, cycle weight
    perm_bisect_attempt Attempt (:113-222): PermuteBeads (the last link of member i now leads to the old end
                        point of member i + 1), Levy construction of every member level by level, Kinetic along
                        the links, pair actions by label (the reference's GetBead(p, b) ignores the links,
                        App. A-4) over the particles list (+ the end points' labels when the window rolls over),
                        Metropolis per level starting from -log(weight)
    apply_cycle         Accept: AssignParticleLabels (perm_bisect_class.h:47-56): from the window's last moved
                        slice to the end of the path the members' labels rotate, and so does the seam permutation

Random numbers are the Philox stream of the device (counter = attempt, clone, slot):
    slot 0            first bead (words 0-1), first particle of the cycle (words 2-3)
    slot 1 + k        step k of the cycle selection: continue? (words 0-1), next particle (words 2-3)
    slot 16 + 128 i + s Levy displacement slots s (as in the single-particle move, s >= 1) of cycle member i
    slot 1040 + level   Metropolis uniform of the level
`perm_philox_numbers` lists them in the order the reference consumes its own, for oracle/refsim.py's injection:
tests/test_stream_ref_cpu.py runs the reference's PermBisectIterative on them and finds this mirror on the same
cycles, the same accept history, the same labels and the same seam permutation after every attempt.
"""
import math

import numpy as np

from . import philox as PX
from .free_spline import FreeSpline

PERM_MAX_LEN = 8
SLOT_CYCLE0, SLOT_LEVY0, SLOT_LEVY_STRIDE = 1, 16, 128
SLOT_METRO0 = SLOT_LEVY0 + PERM_MAX_LEN * SLOT_LEVY_STRIDE


def _pib(d, L, pbc):
    return d - np.rint(d / L) * L if pbc else d


def _dot3(d):
    return (d[..., 0] * d[..., 0] + d[..., 2] * d[..., 2]) + d[..., 1] * d[..., 1]      # arma::dot: two accumulators


class _Stream:
    def __init__(self, seed, attempt, clone):
        self.k0, self.k1 = seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF
        self.a_lo, self.a_hi = attempt & 0xFFFFFFFF, (attempt >> 32) & 0xFFFFFFFF
        self.clone = clone

    def pair(self, slot):
        r = PX.philox4x32(self.a_lo, self.a_hi, self.clone, slot, self.k0, self.k1)
        return PX.uniform_from_bits(r[0], r[1]), PX.uniform_from_bits(r[2], r[3])

    def normals3(self, slot):
        ua, ub = self.pair(slot)
        uc, ud = self.pair(slot + 1)
        ra, rc = math.sqrt(-2.0 * math.log(ua)), math.sqrt(-2.0 * math.log(uc))
        return [ra * math.cos(2 * math.pi * ub), ra * math.sin(2 * math.pi * ub), rc * math.cos(2 * math.pi * ud)]


def chain_label(label, bead0, k, M, perm_next):
    """(label, slice) of the bead k links after (label, bead0): labels change only across the seam."""
    b = bead0 + k
    return (label, b) if b < M else (int(perm_next[label]), b - M)


def perm_table_row(R, perm_next, p, bead0, nb, cfg, lam, log_eps):
    """Row p of UpdatePermTable (:18-28): t(p, j) for every j."""
    M, N = cfg.n_bead, R.shape[0]
    end = np.stack([R[chain_label(j, bead0, nb, M, perm_next)] for j in range(N)])
    dr = _pib(R[p, bead0] - end, cfg.L, cfg.pbc)
    expo = (-_dot3(dr)) * ((1.0 / (4.0 * lam * cfg.tau)) / nb)
    return np.where(expo > log_eps, np.exp(expo), 0.0)


def select_cycle(R, perm_next, bead0, nb, cfg, lam, stream, epsilon=1e-100, max_len=PERM_MAX_LEN):
    """SelectCycleIterative.  Returns (particles or None when the selection stops, weight, steps taken)."""
    N = R.shape[0]
    log_eps = math.log(epsilon)
    _, u_p0 = stream.pair(0)
    p0 = min(int(u_p0 * N), N - 1)
    p, ps, weight_terms, step = p0, [], [], 0
    while True:
        ps.append(p)
        if len(ps) > max_len:
            return "overflow", 0.0, step
        t_row = perm_table_row(R, perm_next, p, bead0, nb, cfg, lam, log_eps)
        t_c = t_row.copy()
        for q in ps:
            t_c[q] = 0.0
        t_c[p0] = t_row[p0]
        Q_p = Q_p_c = 0.0
        for i in range(N):
            Q_p += t_row[i]
            Q_p_c += t_c[i]
        u_cont, u_sel = stream.pair(SLOT_CYCLE0 + step)
        step += 1
        if Q_p_c / Q_p < u_cont:
            return None, 0.0, step
        t_Q, nxt = 0.0, p
        for i in range(N):
            t_Q += t_c[i] / Q_p_c
            if t_Q > u_sel:
                nxt = i
                break
        weight_terms.append((t_row[nxt], t_row[p]))
        p = nxt
        if p == p0:
            break
    weight = 1.0
    for t_next, t_self in weight_terms:
        weight *= t_next / t_self
    return ps, weight, step


def apply_cycle(R, perm_next, ps, bead0, nb, M):
    """AssignParticleLabels after an accepted cycle: in place on R (positions by label) and perm_next."""
    n = len(ps)
    if n < 2:
        return
    e = (bead0 + nb - 1) % M
    labels = [chain_label(a, bead0, nb - 1, M, perm_next)[0] for a in ps]
    old_rows = [R[l, e + 1:].copy() for l in labels]
    old_next = [int(perm_next[l]) for l in labels]
    for i, l in enumerate(labels):
        R[l, e + 1:] = old_rows[(i + 1) % n]
        perm_next[l] = old_next[(i + 1) % n]


def perm_bisect_attempt(cfg, species, n_level, seed, attempt, clone, R, perm_next, action_old_new, with_kinetic=True,
                        n_images_move=0, n_images_kin=0, epsilon=1e-100, max_len=PERM_MAX_LEN):
    """One PermBisectIterative::Attempt + Accept / Reject of one walker.

    R[label][slice][3], perm_next[label]: committed state, UPDATED IN PLACE on acceptance.
    action_old_new(labels, bead0, nb, new_windows) -> (old, new): the pair actions of the species summed, for the sorted
    list of listed labels; new_windows[label] = [nb + 1][3] NEW positions of that label at slices bead0 .. bead0 + nb.
    Returns dict(n_perm (0: the cycle selection stopped, no bisection attempted), particles, accept, bead0)."""
    sp = cfg.species[species]
    N, M, lam, tau = sp.n_part, cfg.n_bead, sp.lam, cfg.tau
    nb = 1 << n_level
    stream = _Stream(seed, attempt, clone)
    u_b0, _ = stream.pair(0)
    bead0 = min(int(u_b0 * M), M - 1)
    ps, weight, steps = select_cycle(R, perm_next, bead0, nb, cfg, lam, stream, epsilon, max_len)
    if ps is None or ps == "overflow":
        return {"n_perm": 0 if ps is None else -1, "particles": [], "accept": False, "bead0": bead0, "steps": steps}
    n = len(ps)
    pib = lambda d: _pib(d, cfg.L, cfg.pbc)
    # chains of the members: OLD follows the committed links, NEW ends on the next member's old end point
    old = [np.stack([R[chain_label(a, bead0, k, M, perm_next)] for k in range(nb + 1)]) for a in ps]
    new = [o.copy() for o in old]
    for i in range(n):
        new[i][nb] = old[(i + 1) % n][nb]
    roll_over = bead0 + nb > M - 1
    labels = set(ps)
    if roll_over:
        labels |= {chain_label(a, bead0, nb, M, perm_next)[0] for a in ps}        # bead_f(i)->GetP(), OLD links
    labels = sorted(labels)
    prev_change = -math.log(weight)
    alive = True
    for level in range(n_level - 1, -1, -1):
        skip = 1 << level
        level_tau = tau * skip
        sigma = math.sqrt(lam * level_tau)
        i4s, i4k = 1.0 / (4.0 * lam * (0.5 * level_tau)), 1.0 / (4.0 * lam * level_tau)
        fs = FreeSpline(cfg.L if cfg.pbc else 0.0, n_images_move, lam, 0.5 * level_tau) if n_images_move else None
        fk = FreeSpline(cfg.L if cfg.pbc else 0.0, n_images_kin, lam, level_tau) if n_images_kin else None
        old_lp = new_lp = 0.0
        for i in range(n):
            idx = 0
            for ia in range(0, nb, 2 * skip):
                ib, ic = ia + skip, ia + 2 * skip
                s = _levy_slot(level, n_level, nb, idx)
                idx += 1
                nrm = np.array(stream.normals3(SLOT_LEVY0 + SLOT_LEVY_STRIDE * i + s))
                rbar_old = old[i][ia] + 0.5 * pib(old[i][ic] - old[i][ia])
                del_old = pib(old[i][ib] - rbar_old)
                rbar_new = new[i][ia] + 0.5 * pib(new[i][ic] - new[i][ia])
                del_new = pib(sigma * nrm)
                new[i][ib] = rbar_new + del_new
                if fs is not None:
                    old_lp += float(fs.GetLogRhoFree(del_old))
                    new_lp += float(fs.GetLogRhoFree(del_new))
                else:
                    old_lp -= float(np.sum(del_old * del_old)) * i4s
                    new_lp -= float(np.sum(del_new * del_new)) * i4s
        old_action = new_action = 0.0
        if with_kinetic:      # Kinetic::GetAction follows the links (GetNextBead): the members' chains
            for i in range(n):
                for ia in range(0, nb, skip):
                    o, n_ = pib(old[i][ia] - old[i][ia + skip]), pib(new[i][ia] - new[i][ia + skip])
                    if fk is not None:
                        old_action -= float(fk.GetLogRhoFree(o))
                        new_action -= float(fk.GetLogRhoFree(n_))
                    else:
                        old_action += float(np.sum(o * o)) * i4k
                        new_action += float(np.sum(n_ * n_)) * i4k
        if level == 0:        # pair actions return 0 above max_level = 0 (pair_action_class.h:269)
            windows = {l: np.stack([R[l, (bead0 + k) % M] for k in range(nb + 1)]) for l in labels}
            for i, a in enumerate(ps):
                for k in range(1, nb):
                    l, b = chain_label(a, bead0, k, M, perm_next)
                    windows[l][k] = new[i][k]
            po, pn = action_old_new(labels, bead0, nb, windows)
            old_action += po
            new_action += pn
        u_metro, _ = stream.pair(SLOT_METRO0 + level)
        change = new_action - old_action
        if (-new_lp + old_lp) - change + prev_change < math.log(u_metro):
            alive = False
            break
        prev_change = change
    if alive:
        for i, a in enumerate(ps):
            for k in range(1, nb):
                l, b = chain_label(a, bead0, k, M, perm_next)
                R[l, b] = new[i][k]
        apply_cycle(R, perm_next, ps, bead0, nb, M)
    return {"n_perm": n, "particles": list(ps), "accept": alive, "bead0": bead0, "labels": labels, "steps": steps}


def _levy_slot(level, n_level, nb, idx):
    """Slot of midpoint idx of a level inside a member's block: the single-particle move's layout (csrc/mc.cuh: SweepSlotStart)."""
    s = 1
    for l in range(n_level - 1, level, -1):
        s += 2 * (nb >> (l + 1)) + 1
    return s + 2 * idx


def perm_philox_numbers(cfg, species, n_level, seed, attempt, clone, n_steps, n_perm):
    """The numbers of one attempt in the reference's order of consumption (perm_bisect_iterative_class.h:114-208):
    uniforms = first bead, first particle, (continue?, next particle) for each of the n_steps selection steps, then --
    only when a cycle of n_perm members closed -- the Metropolis uniform of every level from the top; normals = level
    by level, member by member, midpoint by midpoint."""
    nb = 1 << n_level
    st = _Stream(seed, attempt, clone)
    u_b0, u_p0 = st.pair(0)
    uniforms, normals = [u_b0, u_p0], []
    for k in range(n_steps):
        uniforms += list(st.pair(SLOT_CYCLE0 + k))
    if n_perm > 0:
        for level in range(n_level - 1, -1, -1):
            skip = 1 << level
            for i in range(n_perm):
                for idx in range(nb // (2 * skip)):
                    normals += st.normals3(SLOT_LEVY0 + SLOT_LEVY_STRIDE * i + _levy_slot(level, n_level, nb, idx))
            uniforms.append(st.pair(SLOT_METRO0 + level)[0])
    return uniforms, normals
