"""Kinetic action with periodic images on the device (csrc/kinetic.cuh):
FreeSpline (src/actions/free_spline_class.h:25-84), Kinetic::DActionDBeta / GetAction
(src/actions/single_action/kinetic_class.h:35-45,105-122) and the image-aware Levy sampling of
Bisect (src/events/moves/single_species_move/bisect/bisect_class.h:69-98).

(1) against the REFERENCE'S OWN Kinetic action (oracle/_ref) with 0, 1 and 100 images in a box
    small enough against the thermal wavelength that the images change the leading digits;
(2) against the numpy / scipy mirror of FreeSpline (simpimc_b200/free_spline.py, itself pinned to
    the reference in tests/test_free_spline_cpu.py);
(3) device-resident sweeps with images, attempt by attempt against the host mirror of their Philox
    stream (positions to 1e-12, identical accept history), both sweep implementations.
Tolerance 1e-10 relative on action values (north_star)."""
import numpy as np
import pytest

from simpimc_b200 import system as S

pytestmark = pytest.mark.gpu
RTOL = 1e-10


def _cfg(n_images, N=4, M=16, with_pair=True):
    cfg = S.egas_config(N=N, M=M, n_xy=30, n_r_long=100)       # theta = 0.1: 4 lambda tau ~ L^2 / 15 at M = 16
    cfg.actions.insert(0, S.ActionConfig("Kinetic", "Kinetic", "e", n_images=n_images))
    if not with_pair:
        cfg.actions = cfg.actions[:1]
    return cfg


@pytest.mark.parametrize("n_images", [0, 1, 100])
def test_device_kinetic_matches_the_reference_action(n_images):
    from oracle import refsim
    if not refsim.available():
        pytest.skip("oracle/_ref not built")
    from simpimc_b200 import host
    cfg = _cfg(n_images)
    N, M, L = 4, cfg.n_bead, cfg.L
    C = 3
    rng = np.random.default_rng(41 + n_images)
    R = rng.uniform(-L / 2, L / 2, size=(C, N, M, 3))          # uncorrelated beads: links reach the box edge
    path = host.Path(cfg, n_clones=C)
    path.SetPositions(0, R)
    kin = path.actions[0]
    assert kin.type == "Kinetic" and kin.n_images == n_images
    sims = []
    for c in range(C):
        sim = refsim.RefSim(cfg, seed=3)
        sim.set_positions(0, R[c])
        sims.append(sim)
    du, v = kin.DActionDBeta(), kin.Potential()
    for c, sim in enumerate(sims):
        ref = sim.dbeta(0)
        assert abs(du[c] - ref) <= RTOL * abs(ref), (n_images, c, du[c], ref)
        assert v[c] == 0.0 and sim.potential(0) == 0.0
    # GetAction in OLD and NEW mode at levels 0, 1, 2 with a proposal pending, windows that wrap
    part = np.array([0, 3, 2], dtype=np.int32)
    first = np.array([14, 3, 9], dtype=np.int32)
    n_prop = 7
    newR = rng.uniform(-L / 2, L / 2, size=(C, n_prop, 3))
    path.Propose(0, part, first, newR)
    for c, sim in enumerate(sims):
        sim.propose(0, int(part[c]), int(first[c]), newR[c])
    b0 = (first - 1) % M
    for level in (0, 1, 2):
        for mode in (host.OLD_MODE, host.NEW_MODE):
            path.SetMode(mode)
            got = kin.GetAction(b0, b0 + 8, [(0, part)], level)
            for c, sim in enumerate(sims):
                ref = sim.get_action(0, mode, int(b0[c]), int(b0[c]) + 8, [(0, int(part[c]))], level)
                assert abs(got[c] - ref) <= RTOL * abs(ref), (n_images, level, mode, c, got[c], ref)
    # a window of the whole path closes on itself: the reference's loop body never runs (kinetic_class.h:112-113)
    path.SetMode(host.OLD_MODE)
    assert np.all(kin.GetAction(b0, b0 + M, [(0, part)], 0) == 0.0)
    # particles of another species are ignored (:110)
    path.Commit(0)
    for sim in sims:
        sim.close()
    path.close()


@pytest.mark.parametrize("n_images,L_scale", [(1, 1.0), (100, 1.0), (3, 4.0)])
def test_device_kinetic_matches_the_free_spline_mirror(n_images, L_scale):
    """No reference library needed: the scipy mirror on more points, including the regime where the
    image sum underflows over the centre of the grid (L_scale = 4: lookups in the zero run)."""
    from simpimc_b200 import host
    from simpimc_b200.free_spline import FreeSpline
    cfg = _cfg(n_images, N=6, M=16, with_pair=False)
    cfg.L *= L_scale
    N, M, L, lam, tau = 6, cfg.n_bead, cfg.L, 0.5, cfg.tau
    C = 4
    rng = np.random.default_rng(7)
    R = rng.uniform(-L / 2, L / 2, size=(C, N, M, 3))
    R[1] = R[1, :, :1, :] + 0.05 * rng.standard_normal((N, M, 3))     # short links: the centre of the table
    path = host.Path(cfg, n_clones=C)
    path.SetPositions(0, R)
    kin = path.actions[0]
    pib = lambda d: d - np.rint(d / L) * L
    links = pib(R - np.roll(R, -1, axis=2))
    fs = FreeSpline(L, n_images, lam, tau, use_tau_derivative=True)
    ref = N * M * 3 / (2 * tau) + np.sum(fs.GetDLogRhoFreeDTau(links), axis=(1, 2))
    got = kin.DActionDBeta()
    assert np.all(np.abs(got - ref) <= RTOL * np.abs(ref)), (got, ref)
    tot = kin.TotalAction()           # every link of the path at level 0
    ref = -np.sum(FreeSpline(L, n_images, lam, tau).GetLogRhoFree(links), axis=(1, 2))
    assert np.all(np.abs(tot - ref) <= RTOL * np.abs(ref)), (tot, ref)
    path.SetMode(host.OLD_MODE)
    for level in range(0, 4):
        skip = 1 << level
        fl = FreeSpline(L, n_images, lam, tau * skip)
        b0 = np.array([3, 15, 8, 0], dtype=np.int32)
        got = kin.GetAction(b0, b0 + 8, [(0, 2), (0, 5)], level)
        for c in range(C):
            idx = (b0[c] + np.arange(0, 8, skip)) % M
            ref = -sum(float(np.sum(fl.GetLogRhoFree(pib(R[c, p, idx] - R[c, p, (idx + skip) % M])))) for p in (2, 5))
            assert abs(got[c] - ref) <= RTOL * abs(ref), (level, c, got[c], ref)
    path.close()


@pytest.mark.parametrize("n_images_move,n_images_kin,general,lr", [(1, 1, False, True), (100, 100, True, True), (2, 0, False, False), (0, 3, True, False)])
def test_device_sweeps_with_images_follow_the_host_mirror(n_images_move, n_images_kin, general, lr):
    """pimc_bisect_sweep with image-aware Levy sampling and Kinetic action, attempt by attempt against
    moves.bisect_attempt_philox (FreeSpline mirror + CPU oracle for the pair action): same positions to
    1e-12 and the same accept history; both outcomes must occur."""
    from simpimc_b200 import host, moves
    from oracle import oracle as O
    N, M, n_level = 5, 16, 3
    cfg = S.egas_config(N=N, M=M, n_xy=30, n_r_long=100)
    if not lr:
        cfg = S.ueg_config(N=N, M=M, rs=1.0, theta=0.1, use_long_range=False, n_xy=30, n_r_long=100)
    ocfg = cfg                                                    # the oracle knows the pair action only
    import copy
    gcfg = copy.copy(cfg)
    gcfg.actions = [S.ActionConfig("Kinetic", "Kinetic", "e", n_images=n_images_kin)] + list(cfg.actions)
    C = 3
    path = host.Path(gcfg, n_clones=C)
    path.ForceGeneral(general)
    path.SetMoveImages(0, n_images_move)
    R = np.stack([S.synthetic_paths(cfg, 0, c, 5) for c in range(C)])
    path.SetPositions(0, R)
    oracles = []
    for c in range(C):
        o = O.Oracle(ocfg)
        o.set_positions(0, R[c])
        oracles.append(o)
    seed = 0x0BADC0DE00000007
    n_dev = np.zeros(C, dtype=np.int64)
    n_host = np.zeros(C, dtype=np.int64)

    def get_beads(c, p, b0, n):
        return oracles[c].get_positions(0, 0)[p, (b0 + np.arange(n)) % M]

    def action_old_new(c, p, b0, nb, new):
        oracles[c].propose(0, p, (b0 + 1) % M, new)
        return oracles[c].get_action(0, 0, b0, b0 + nb, [(0, p)], 0), oracles[c].get_action(0, 1, b0, b0 + nb, [(0, p)], 0)

    def finish(c, p, b0, nb, accept, new):
        if new is not None:
            oracles[c].finish_move(0, p, b0, b0 + nb, bool(accept))

    n_att = 40
    for attempt in range(n_att):
        _, _, acc = moves.bisect_attempt_philox(ocfg, 0, n_level, seed, attempt, C, get_beads, action_old_new, finish,
                                                n_images_move=n_images_move, n_images_kin=n_images_kin)
        n_host += acc
        n_dev += path.BisectSweep(0, n_level, 1, seed, attempt0=attempt)
        got = path.GetPositions(0)
        for c in range(C):
            ref = oracles[c].get_positions(0, 0)
            assert np.max(np.abs(got[c] - ref)) <= 1e-12 * max(1.0, np.max(np.abs(ref))), (attempt, c)
        assert np.array_equal(n_dev, n_host), (attempt, n_dev, n_host)
    assert 0 < n_dev.sum() < n_att * C, n_dev
    for o in oracles:
        o.close()
    path.close()
