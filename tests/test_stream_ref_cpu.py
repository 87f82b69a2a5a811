"""CPU test: the host mirror of the device-resident bisection (simpimc_b200.moves.bisect_attempt_philox, which
tests/test_gpu_sweep.py and tests/test_gpu_kinetic.py pin the CUDA sweeps to attempt by attempt) against the
REFERENCE PROGRAM ITSELF -- its own Bisect::Attempt / Accept / Reject, Kinetic (FreeSpline with periodic images) and
IlkkaPairAction, compiled in place (oracle/_ref) -- made to consume the numbers of the device's Philox stream
(oracle/refsim.py: inject_random).  Same decisions, same bead positions after every attempt: the device sweeps
therefore follow the reference's Markov chain move for move, not just statistically."""
import copy

import numpy as np
import pytest

from simpimc_b200 import moves, system as S

pytestmark = pytest.mark.ref


@pytest.mark.parametrize("name,n_images_kin,n_images_move", [("ueg", 0, 0), ("egas", 100, 1), ("egas", 2, 3), ("nolr", 1, 0)])
def test_host_mirror_follows_the_reference_program(name, n_images_kin, n_images_move, oracle_mod):
    from oracle import refsim
    if not refsim.available():
        pytest.skip("oracle/_ref not built")
    N, M, n_level = 5, 16, 3
    if name == "ueg":
        pair_cfg = S.ueg_config(N=N, M=M, n_xy=40, n_r_long=200)
    elif name == "nolr":
        pair_cfg = S.ueg_config(N=N, M=M, rs=1.0, theta=0.1, use_long_range=False, n_xy=40, n_r_long=200)
    else:
        pair_cfg = S.egas_config(N=N, M=M, n_xy=40, n_r_long=200)      # theta = 0.1: the images matter
    cfg = copy.copy(pair_cfg)
    cfg.actions = [S.ActionConfig("Kinetic", "Kinetic", "e", n_images=n_images_kin)] + list(pair_cfg.actions)
    cfg.moves = [{"name": "BisectE", "type": "Bisect", "species": "e", "n_level": n_level, "n_images": n_images_move}]
    cfg.observables = []
    sim = refsim.RefSim(cfg, seed=1)
    if not hasattr(sim.lib, "ref_inject_random"):
        pytest.skip("oracle/_ref predates the injection hooks")
    R = S.synthetic_paths(cfg, 0, 0, 5)
    sim.set_positions(0, R)
    o = oracle_mod.Oracle(pair_cfg)
    o.set_positions(0, R)
    seed = 0x5EED00000000ABCD

    def get_beads(c, p, b0, n):
        return o.get_positions(0, 0)[p, (b0 + np.arange(n)) % M]

    def action_old_new(c, p, b0, nb, new):
        o.propose(0, p, (b0 + 1) % M, new)
        return o.get_action(0, 0, b0, b0 + nb, [(0, p)], 0), o.get_action(0, 1, b0, b0 + nb, [(0, p)], 0)

    def finish(c, p, b0, nb, accept, new):
        if new is not None:
            o.finish_move(0, p, b0, b0 + nb, bool(accept))

    n_att, n_acc = 120, 0
    for attempt in range(n_att):
        u, n = moves.bisect_philox_numbers(cfg, 0, n_level, seed, attempt, 0)
        sim.inject_random(u, n)
        sim.move_do(0, 1)
        left_u, left_n = sim.inject_pending(clear=True)
        _, _, acc = moves.bisect_attempt_philox(pair_cfg, 0, n_level, seed, attempt, 1, get_beads, action_old_new, finish,
                                                n_images_move=n_images_move, n_images_kin=n_images_kin)
        n_acc += int(acc[0])
        att_ref, acc_ref = sim.move_counts(0)
        assert (att_ref, acc_ref) == (attempt + 1, n_acc), (name, attempt, att_ref, acc_ref, n_acc)
        if acc[0]:                      # an accepted attempt consumed every number
            assert (left_u, left_n) == (0, 0), (attempt, left_u, left_n)
        a, b = sim.get_positions(0, 0), o.get_positions(0, 0)
        assert np.max(np.abs(a - b)) <= 1e-12 * max(1.0, np.max(np.abs(b))), (name, attempt, np.max(np.abs(a - b)))
    assert 0 < n_acc < n_att, n_acc          # both outcomes occurred
    o.close()
    sim.close()


@pytest.mark.parametrize("name", ["ueg", "plasma"])
def test_displace_mirror_follows_the_reference_program(name, oracle_mod):
    """DisplaceParticle::Attempt (displace_particle_class.h:28-71) of the reference on injected Philox numbers against
    moves.displace_attempt_philox, the host mirror the device-resident pimc_displace_sweep is pinned to."""
    from oracle import refsim
    if not refsim.available():
        pytest.skip("oracle/_ref not built")
    if name == "ueg":
        pair_cfg, sp = S.ueg_config(N=6, M=8, n_xy=40, n_r_long=200), 0
    else:
        pair_cfg, sp = S.plasma_config(Ne=4, Np=3, M=8), 1
    cfg = copy.copy(pair_cfg)
    spn = cfg.species[sp].name
    step = 0.25
    cfg.actions = [S.ActionConfig("Kinetic", "Kinetic", spn, n_images=1)] + list(pair_cfg.actions)
    cfg.moves = [{"name": "Displace", "type": "DisplaceParticle", "species": spn, "step_size": step}]
    cfg.observables = []
    sim = refsim.RefSim(cfg, seed=1)
    if not hasattr(sim.lib, "ref_inject_random"):
        pytest.skip("oracle/_ref predates the injection hooks")
    o = oracle_mod.Oracle(pair_cfg)
    for s_i in range(len(cfg.species)):
        R = S.synthetic_paths(cfg, s_i, 0, 5)
        sim.set_positions(s_i, R)
        o.set_positions(s_i, R)
    M = cfg.n_bead
    acts = [ai for ai, a in enumerate(pair_cfg.actions) if spn in (a.species_a, a.species_b)]
    seed = 0x0D15C0DE

    def get_beads(c, p, b0, n):
        return o.get_positions(sp, 0)[p, (b0 + np.arange(n)) % M]

    def action_old_new(c, p, new):
        o.propose(sp, p, 0, new)
        old = sum(o.get_action(ai, 0, 0, M, [(sp, p)], 0) for ai in acts)
        nw = sum(o.get_action(ai, 1, 0, M, [(sp, p)], 0) for ai in acts)
        return old, nw

    def finish(c, p, accept):
        o.finish_move(sp, p, 0, M, bool(accept))

    n_att, n_acc = 60, 0
    for attempt in range(n_att):
        sim.inject_random(moves.displace_philox_numbers(seed, attempt, 0), [])
        sim.move_do(0, 1)
        assert sim.inject_pending(clear=True) == (0, 0)
        _, acc = moves.displace_attempt_philox(pair_cfg, sp, step, seed, attempt, 1, get_beads, action_old_new, finish)
        n_acc += int(acc[0])
        assert sim.move_counts(0) == (attempt + 1, n_acc), (name, attempt)
        a, b = sim.get_positions(sp, 0), o.get_positions(sp, 0)
        assert np.max(np.abs(a - b)) <= 1e-12 * max(1.0, np.max(np.abs(b))), (name, attempt)
    assert 0 < n_acc < n_att, n_acc
    o.close()
    sim.close()


@pytest.mark.parametrize("n_images_kin,n_level,n_images_move", [(0, 2, 0), (1, 3, 0), (2, 3, 1)])
def test_permuting_bisection_mirror_follows_the_reference_program(n_images_kin, n_level, n_images_move, oracle_mod):
    """PermBisectIterative (cycle selection from the permutation table, PermuteBeads, the members' Levy bridges,
    AssignParticleLabels; perm_bisect_iterative_class.h:10-222, perm_bisect_class.h:32-82) of the reference on injected
    Philox numbers against simpimc_b200.perm_moves.perm_bisect_attempt, which keeps positions by label plus the seam
    permutation.  After every attempt: the same cycle lengths attempted and accepted (integer state,
    bit-exact), the same seam permutation, the same positions by label; windows that roll over the seam of an already
    permuted path and cycles of two and more particles occur."""
    from oracle import refsim
    from simpimc_b200 import perm_moves as PM
    if not refsim.available():
        pytest.skip("oracle/_ref not built")
    N, M = 5, 16
    pair_cfg = S.egas_config(N=N, M=M, n_xy=40, n_r_long=200)      # theta = 0.1: exchange is frequent
    cfg = copy.copy(pair_cfg)
    cfg.actions = [S.ActionConfig("Kinetic", "Kinetic", "e", n_images=n_images_kin)] + list(pair_cfg.actions)
    cfg.moves = [{"name": "PermE", "type": "PermBisectIterative", "species": "e", "n_level": n_level, "n_images": n_images_move}]
    cfg.observables = []
    sim = refsim.RefSim(cfg, seed=1)
    if not hasattr(sim.lib, "ref_inject_random") or not hasattr(sim.lib, "ref_permutation"):
        pytest.skip("oracle/_ref predates the injection hooks")
    R = S.synthetic_paths(cfg, 0, 0, 5).copy()
    perm_next = np.arange(N, dtype=np.int32)
    sim.set_positions(0, R)
    o = oracle_mod.Oracle(pair_cfg)
    o.set_positions(0, R)
    seed = 0x9E3700000000C1C1

    def action_old_new(labels, bead0, nb, windows):
        for l in labels:
            o.propose(0, l, bead0, windows[l])
        parts = [(0, l) for l in labels]
        old = o.get_action(0, 0, bead0, bead0 + nb, parts, 0)
        new = o.get_action(0, 1, bead0, bead0 + nb, parts, 0)
        for l in labels:
            o.finish_move(0, l, bead0, bead0 + nb, False)
        return old, new

    n_att = 150
    attempted = np.zeros(8, dtype=np.int64)
    accepted = np.zeros(8, dtype=np.int64)
    n_acc = wrapped_on_permuted = 0
    for attempt in range(n_att):
        R_try, perm_try = R.copy(), perm_next.copy()
        res = PM.perm_bisect_attempt(pair_cfg, 0, n_level, seed, attempt, 0, R_try, perm_try, action_old_new, n_images_kin=n_images_kin,
                                     n_images_move=n_images_move)
        u, n = PM.perm_philox_numbers(pair_cfg, 0, n_level, seed, attempt, 0, res["steps"], max(res["n_perm"], 0))
        sim.inject_random(u, n)
        sim.move_do(0, 1)
        left = sim.inject_pending(clear=True)
        if res["n_perm"] > 0:
            attempted[res["n_perm"] - 1] += 1
            if res["accept"]:
                accepted[res["n_perm"] - 1] += 1
                n_acc += 1
                assert left == (0, 0), (attempt, left)
                if res["bead0"] + (1 << n_level) > M - 1 and not np.array_equal(perm_next, np.arange(N)):
                    wrapped_on_permuted += 1
                R, perm_next = R_try, perm_try
                o.set_positions(0, R)
        att_ref, acc_ref = sim.perm_counts(0)
        assert np.array_equal(att_ref[:N], attempted[:N]) and np.array_equal(acc_ref[:N], accepted[:N]), (attempt, att_ref, attempted, acc_ref, accepted)
        assert sim.move_counts(0) == (attempt + 1, n_acc)
        assert np.array_equal(sim.permutation(0)[1], perm_next), (attempt, sim.permutation(0)[1], perm_next)
        ref_R = sim.get_positions(0, 0)
        assert np.max(np.abs(ref_R - R)) <= 1e-12 * max(1.0, np.max(np.abs(R))), (attempt, res)
    assert accepted[1:].sum() >= 3, accepted          # real permutations (two or more particles) were accepted
    assert wrapped_on_permuted >= 1                   # a window rolled over the seam of an already permuted path
    assert 0 < n_acc < n_att
    o.close()
    sim.close()


def test_permuting_bisection_mirror_follows_the_reference_program_on_two_species(oracle_mod):
    """The same pin for a two-species plasma: permuting moves of the electrons (species a of the e-p action) and of the
    protons (species b of it, the (unlisted a, listed b) branch of GenerateParticlePairs, pair_action_class.h:95-112),
    alternating; three pair actions with long range, every one of them evaluated for the listed labels."""
    from oracle import refsim
    from simpimc_b200 import perm_moves as PM
    if not refsim.available():
        pytest.skip("oracle/_ref not built")
    n_level = 2
    pair_cfg = S.plasma_config(Ne=5, Np=4, M=8, theta=0.25, n_xy=40, n_r_long=200, pp_action="IlkkaPairAction")
    M = pair_cfg.n_bead
    cfg = copy.copy(pair_cfg)
    cfg.actions = [S.ActionConfig("KineticE", "Kinetic", "e", n_images=0), S.ActionConfig("KineticP", "Kinetic", "p", n_images=0)] + list(pair_cfg.actions)
    cfg.moves = [{"name": "PermE", "type": "PermBisectIterative", "species": "e", "n_level": n_level, "n_images": 0},
                 {"name": "PermP", "type": "PermBisectIterative", "species": "p", "n_level": n_level, "n_images": 0}]
    cfg.observables = []
    sim = refsim.RefSim(cfg, seed=1)
    if not hasattr(sim.lib, "ref_inject_random") or not hasattr(sim.lib, "ref_permutation"):
        pytest.skip("oracle/_ref predates the injection hooks")
    o = oracle_mod.Oracle(pair_cfg)
    R, nxt = [], []
    for sp in range(2):
        R.append(S.synthetic_paths(cfg, sp, 0, 5).copy())
        nxt.append(np.arange(cfg.species[sp].n_part, dtype=np.int32))
        sim.set_positions(sp, R[sp])
        o.set_positions(sp, R[sp])
    seed = 0x9E3700000000C1C1
    n_acc = [0, 0]
    n_done = [0, 0]
    exchanged = 0
    for attempt in range(160):
        sp = attempt % 2
        acts = [ai for ai, a in enumerate(pair_cfg.actions) if cfg.species[sp].name in (a.species_a, a.species_b)]

        def action_old_new(labels, bead0, nb, windows):
            for l in labels:
                o.propose(sp, l, bead0, windows[l])
            parts = [(sp, l) for l in labels]
            old = sum(o.get_action(ai, 0, bead0, bead0 + nb, parts, 0) for ai in acts)
            new = sum(o.get_action(ai, 1, bead0, bead0 + nb, parts, 0) for ai in acts)
            for l in labels:
                o.finish_move(sp, l, bead0, bead0 + nb, False)
            return old, new

        R_try, perm_try = R[sp].copy(), nxt[sp].copy()
        res = PM.perm_bisect_attempt(pair_cfg, sp, n_level, seed, attempt, 0, R_try, perm_try, action_old_new)
        u, n = PM.perm_philox_numbers(pair_cfg, sp, n_level, seed, attempt, 0, res["steps"], max(res["n_perm"], 0))
        sim.inject_random(u, n)
        sim.move_do(sp, 1)
        left = sim.inject_pending(clear=True)
        n_done[sp] += 1
        if res["n_perm"] > 0 and res["accept"]:
            n_acc[sp] += 1
            assert left == (0, 0), (attempt, left)
            exchanged += res["n_perm"] >= 2
            R[sp], nxt[sp] = R_try, perm_try
            o.set_positions(sp, R[sp])
        assert sim.move_counts(sp) == (n_done[sp], n_acc[sp]), (attempt, sp, sim.move_counts(sp), n_done, n_acc)
        assert np.array_equal(sim.permutation(sp)[1], nxt[sp]), (attempt, sp)
        ref_R = sim.get_positions(sp, 0)
        assert np.max(np.abs(ref_R - R[sp])) <= 1e-12 * max(1.0, np.max(np.abs(R[sp]))), (attempt, sp, res)
    assert n_acc[0] > 0 and n_acc[1] > 0 and exchanged >= 1, (n_acc, exchanged)
    o.close()
    sim.close()
