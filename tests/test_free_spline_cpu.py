"""CPU tests of the FreeSpline host mirror (simpimc_b200/free_spline.py): against closed forms,
and -- where oracle/_ref exists -- against the REFERENCE'S OWN Kinetic action
(src/actions/single_action/kinetic_class.h on src/actions/free_spline_class.h, compiled in place)
with 1 and 100 periodic images, in a box small enough against the thermal wavelength that the
image sum changes the action in its leading digits."""
import numpy as np
import pytest

from simpimc_b200 import system as S
from simpimc_b200.free_spline import FreeSpline


def _kinetic_cfg(n_images, N=4, M=16):
    cfg = S.egas_config(N=N, M=M, n_xy=30, n_r_long=100)       # theta = 0.1: 4 lambda tau ~ L^2 / 15 at M = 16
    cfg.actions.insert(0, S.ActionConfig("Kinetic", "Kinetic", "e", n_images=n_images))
    return cfg


def test_no_images_is_the_closed_form():
    fs = FreeSpline(3.0, 0, 0.5, 0.2, use_tau_derivative=True)
    r = np.array([[0.3, -1.2, 1.5], [0.0, 0.0, 0.0]])
    assert np.allclose(fs.GetLogRhoFree(r), -np.sum(r * r, axis=1) / (4 * 0.5 * 0.2), rtol=1e-15, atol=0)
    assert np.allclose(fs.GetDLogRhoFreeDTau(r), -np.sum(r * r, axis=1) / (4 * 0.5 * 0.2 * 0.2), rtol=1e-15, atol=0)


def test_image_sum_at_the_box_edge_and_centre():
    """At r = -L/2 the first image sits at +L/2 with the same weight: image_action = -log1p(1 + ...)."""
    L, lam, tau = 2.0, 0.5, 0.3
    fs = FreeSpline(L, 3, lam, tau)
    a = 1.0 / (4 * lam * tau)
    exact = lambda r: -np.log1p(sum(np.exp((r * r - (r + i * L) ** 2) * a) + np.exp((r * r - (r - i * L) ** 2) * a) for i in (1, 2, 3)))
    for r in (-L / 2, 0.0, 0.3712, L / 2):
        assert abs(float(fs.image_action(r)) - exact(r)) <= 1e-9, r


@pytest.mark.ref
@pytest.mark.parametrize("n_images", [1, 100])
def test_mirror_matches_the_reference_kinetic_action(n_images):
    from oracle import refsim
    if not refsim.available():
        pytest.skip("oracle/_ref not built")
    cfg = _kinetic_cfg(n_images)
    N, M, lam, tau, L = 4, cfg.n_bead, 0.5, cfg.tau, cfg.L
    sim = refsim.RefSim(cfg, seed=3)
    rng = np.random.default_rng(9)
    R = rng.uniform(-L / 2, L / 2, size=(N, M, 3))       # uncorrelated beads: links reach the box edge
    sim.set_positions(0, R)
    pib = lambda d: d - np.rint(d / L) * L
    # DActionDBeta (kinetic_class.h:35-45)
    fs = FreeSpline(L, n_images, lam, tau, use_tau_derivative=True)
    links = pib(R - np.roll(R, -1, axis=1))
    mine = N * M * 3 / (2 * tau) + float(np.sum(fs.GetDLogRhoFreeDTau(links)))
    ref = sim.dbeta(0)
    assert abs(mine - ref) <= 1e-9 * abs(ref), (mine, ref)
    # and the image part matters at this tau / L
    closed = N * M * 3 / (2 * tau) - float(np.sum(links * links)) / (4 * lam * tau * tau)
    assert abs(closed - ref) > 1e-3 * abs(ref)
    # GetAction at levels 0, 1, 2 (kinetic_class.h:105-122), windows that wrap around
    for level, b0, nb in ((0, 13, 6), (1, 10, 8), (2, 4, 8)):
        skip = 1 << level
        fl = FreeSpline(L, n_images, lam, tau * skip)
        for p in (0, 3):
            idx = (b0 + np.arange(0, nb, skip)) % M
            nxt = (idx + skip) % M
            mine = -float(np.sum(fl.GetLogRhoFree(pib(R[p, idx] - R[p, nxt]))))
            ref = sim.get_action(0, 0, b0, b0 + nb, [(0, p)], level)
            assert abs(mine - ref) <= 1e-9 * abs(ref), (level, p, mine, ref)
    # a window of the whole path: bead_b == bead_a, the loop body never runs (kinetic_class.h:112-113)
    assert sim.get_action(0, 0, 3, 3 + M, [(0, 1)], 0) == 0.0
    sim.close()
