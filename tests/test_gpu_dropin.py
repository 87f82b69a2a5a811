"""Drop-in test: the REFERENCE'S OWN moves and estimators (Bisect::Attempt / Accept / Reject,
DisplaceParticle, Energy -- compiled in place from /root/reference/src into
integration/_build/libsimpimc_dropin.so) running on top of the CUDA library through
include/simpimc_b200_action.hpp (class GpuPairAction : public Action), against the same
program with the reference's IlkkaPairAction / BarePairAction (oracle/_ref).  Same seed, same
RNG stream: every Metropolis decision, hence every bead position, must come out identical,
and the energies must agree to 1e-10."""
import numpy as np
import pytest

from simpimc_b200 import system as S

pytestmark = pytest.mark.gpu


def _sims(cfg, seed):
    from oracle import refsim
    if not (refsim.available() and refsim.dropin_available()):
        pytest.skip("oracle/_ref or integration/_build not built (needs /root/reference at build time)")
    a = refsim.RefSim(cfg, seed=seed)
    b = refsim.RefSim(cfg, seed=seed, dropin=True)
    for sp in range(len(cfg.species)):
        R = S.synthetic_paths(cfg, sp, 0, 77 + seed)
        a.set_positions(sp, R)
        b.set_positions(sp, R)
    return a, b


@pytest.mark.parametrize("which", ["ueg", "plasma", "perm", "david", "david_lr"])
def test_reference_moves_on_gpu_actions_reproduce_the_reference_run(which):
    if which == "perm":
        # the reference's own permuting bisection (PermBisectIterative: cycle selection from the
        # permutation table, relabelling of accepted permutations) on top of the CUDA action:
        # GetAction receives the 1-4 particles of the cycle
        cfg = S.ueg_config(N=7, M=16, with_kinetic=True)
        cfg.moves = [{"name": "PermE", "type": "PermBisectIterative", "species": "e", "n_level": 2, "n_images": 0},
                     {"name": "BisectE", "type": "Bisect", "species": "e", "n_level": 3}]
    elif which == "david":
        # DavidPairAction behind the adapter (david_pair_action_class.h:192-360 reads u_kj_<n_order>, du_kj_dbeta_<n_order>,
        # potential): an e-p action between different species on the LOG grid and a p-p action on a LINEAR grid, next
        # to an Ilkka e-e action with long range
        cfg = S.plasma_config(Ne=6, Np=5, M=8, pp_action="DavidPairAction", ep_action="DavidPairAction")
        cfg.actions.insert(0, S.ActionConfig("KineticE", "Kinetic", "e"))
        cfg.actions.insert(1, S.ActionConfig("KineticP", "Kinetic", "p"))
        cfg.moves = [{"name": "BisectE", "type": "Bisect", "species": "e", "n_level": 2},
                     {"name": "BisectP", "type": "Bisect", "species": "p", "n_level": 2},
                     {"name": "DisplaceP", "type": "DisplaceParticle", "species": "p", "step_size": 0.2}]
    elif which == "david_lr":
        # DavidPairAction with use_long_range=1 (long_range/{n_k,k_points,u_k}, squarer/v_image): the k-space part of
        # the action differences and of DActionDBeta.  Potential() is not measured: the reference's CalcVLong indexes a
        # shell-length array by k vector (david_pair_action_class.h:42-60 vs :324)
        cfg = S.ueg_config(N=9, M=16, action="DavidPairAction", use_long_range=True, with_kinetic=True)
        cfg.moves = [{"name": "BisectE", "type": "Bisect", "species": "e", "n_level": 3},
                     {"name": "DisplaceE", "type": "DisplaceParticle", "species": "e", "step_size": 0.3}]
    elif which == "ueg":
        cfg = S.ueg_config(N=14, M=16, with_kinetic=True)
        cfg.moves = [{"name": "BisectE", "type": "Bisect", "species": "e", "n_level": 3},
                     {"name": "DisplaceE", "type": "DisplaceParticle", "species": "e", "step_size": 0.3}]
    else:
        cfg = S.plasma_config(Ne=6, Np=5, M=8)
        cfg.actions.insert(0, S.ActionConfig("KineticE", "Kinetic", "e"))
        cfg.actions.insert(1, S.ActionConfig("KineticP", "Kinetic", "p"))
        cfg.moves = [{"name": "BisectE", "type": "Bisect", "species": "e", "n_level": 2},
                     {"name": "BisectP", "type": "Bisect", "species": "p", "n_level": 2},
                     {"name": "DisplaceP", "type": "DisplaceParticle", "species": "p", "step_size": 0.2}]
    cfg.observables = [{"name": "Energy", "type": "Energy", "measure_potential": 0 if which == "david_lr" else 1}]
    a, b = _sims(cfg, seed=5)
    n_moves = len(cfg.moves)
    for sweep in range(30):
        for m in range(n_moves):
            a.move_do(m, 5)
            b.move_do(m, 5)
        a.observable_accumulate(0)
        b.observable_accumulate(0)
    for m in range(n_moves):
        assert a.move_counts(m) == b.move_counts(m), "accept/reject history differs"
        att, acc = a.move_counts(m)
        assert att == 150 and 0 < acc
    if which == "perm":
        (att_a, acc_a), (att_b, acc_b) = a.perm_counts(0), b.perm_counts(0)
        assert np.array_equal(att_a, att_b) and np.array_equal(acc_a, acc_b)
        assert acc_a[1:].sum() >= 1, "no real permutation (cycle of >= 2 particles) was accepted"
    for sp in range(len(cfg.species)):
        assert np.array_equal(a.get_positions(sp, 0), b.get_positions(sp, 0)), "trajectories diverged"
    # the adapter's GetActionGradient / GetActionLaplacian after the run (pair actions only: Kinetic is the reference's own)
    for ai, acfg in enumerate(cfg.actions):
        if acfg.type == "Kinetic":
            continue
        for sp, p, b0, b1 in ((0, 1, 3, 6), (len(cfg.species) - 1, 2, cfg.n_bead - 1, cfg.n_bead)):
            ga, gb = a.action_gradient(ai, 1, b0, b1, [(sp, p)], 0), b.action_gradient(ai, 1, b0, b1, [(sp, p)], 0)
            la, lb = a.action_laplacian(ai, 1, b0, b1, [(sp, p)], 0), b.action_laplacian(ai, 1, b0, b1, [(sp, p)], 0)
            u = 2.0 * abs(a.get_action(ai, 0, b0, b1, [(sp, p)], 0)) + 1e-3
            analytic = acfg.type == "IlkkaPairAction"
            assert np.max(np.abs(ga - gb)) <= 1e-10 * max(np.max(np.abs(ga)), u if analytic else u / 1e-4), (ai, ga, gb)
            assert abs(la - lb) <= 1e-10 * max(abs(la), u / 1e-8), (ai, la, lb)
    ea, va = a.energy_sums(0)
    eb, vb = b.energy_sums(0)
    assert np.all(np.abs(ea - eb) <= 1e-10 * np.maximum(np.abs(ea), 1e-300)), (ea, eb)
    assert np.all(np.abs(va - vb) <= 1e-10 * np.maximum(np.abs(va), 1e-300)), (va, vb)
    a.close()
    b.close()


def test_adapter_evaluates_the_particle_lists_of_permutation_moves():
    """GpuPairAction::GetAction with several particles of one species (what PermBisect passes,
    perm_bisect_iterative_class.h:200-204) and a mixed-species list: the adapter proposes every
    listed particle's NEW beads and evaluates the (listed, other) + (listed, listed) pairs; same
    values as the reference's own actions, and the same state after the move is stored."""
    cfg = S.plasma_config(Ne=6, Np=5, M=8)
    a, b = _sims(cfg, seed=9)
    rng = np.random.default_rng(2)
    M = cfg.n_bead
    for trial, (plist, b0, nb, accept) in enumerate([([(0, 1), (0, 4)], 2, 4, True), ([(0, 0), (0, 2), (0, 5), (1, 3)], 6, 4, False),
                                                     ([(1, 0), (1, 4), (0, 3)], 0, 2, True)]):
        first, n = (b0 + 1) % M, nb - 1
        for sp, p in plist:
            cur = a.get_positions(sp, 0)[p, (first + np.arange(n)) % M]
            new = cur + 0.06 * rng.standard_normal(cur.shape)
            for sim in (a, b):
                for i in range(n):      # bead by bead: the window may wrap past n_bead
                    sim.propose(sp, p, (first + i) % M, new[i:i + 1])
        for ai in range(len(cfg.actions)):
            oa, ob = a.get_action(ai, 0, b0, b0 + nb, plist, 0), b.get_action(ai, 0, b0, b0 + nb, plist, 0)
            na, nb_ = a.get_action(ai, 1, b0, b0 + nb, plist, 0), b.get_action(ai, 1, b0, b0 + nb, plist, 0)
            assert abs(oa - ob) <= 1e-10 * abs(oa) + 1e-300 and abs(na - nb_) <= 1e-10 * abs(na) + 1e-300, (trial, ai, oa, ob, na, nb_)
            assert abs((na - oa) - (nb_ - ob)) <= 1e-10 * max(abs(na - oa), 1e-4 * (abs(na) + abs(oa)))
        for sp, p in plist:
            for sim in (a, b):
                sim.finish_move(sp, p, b0, b0 + nb, accept)
        for ai in range(len(cfg.actions)):
            ea, eb = a.dbeta(ai), b.dbeta(ai)
            assert abs(ea - eb) <= 1e-10 * abs(ea), (trial, ai, ea, eb)
    a.close()
    b.close()
