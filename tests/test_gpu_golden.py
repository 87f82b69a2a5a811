"""GPU parity against the reference's own numbers: the CUDA path through the C ABI versus
tests/golden/*.npz (written by oracle/make_golden.py from the reference compiled in place).
Runs every fixture twice: default dispatch (fast Ilkka kernel where it applies) and with the
general kernels forced."""
import pytest

import golden_util as G

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("general", [False, True])
@pytest.mark.parametrize("name", sorted(G.CONFIGS))
def test_gpu_matches_reference_golden(name, general):
    G.check_backend(name, lambda cfg, seed: G.GpuBackend(cfg, seed, general=general))


def test_madelung_constant_on_gpu():
    """Same known answer as tests/test_oracle_cpu.py, evaluated by the CUDA kernels (K1 potential
    variant + K2 rho_k + K3 k-sum with 1871 k vectors)."""
    import numpy as np
    from simpimc_b200 import host, system as S, tables as T
    L = 2.0
    k_cut = 30.0 / (L / 2.0)
    cfg = S.SystemConfig(n_d=3, n_bead=1, beta=1.0, L=L, pbc=True, k_cut=k_cut)
    cfg.species.append(S.SpeciesConfig("Na", 4, 0.5))
    cfg.species.append(S.SpeciesConfig("Cl", 4, 0.5))
    for nm, a, b, z in (("NaNa", "Na", "Na", 1.0), ("NaCl", "Na", "Cl", -1.0), ("ClCl", "Cl", "Cl", 1.0)):
        tab = T.make_bare_table(z, L, k_cut, n_r=4000, r_max=4.0, n_r_long=2000)
        cfg.actions.append(S.ActionConfig(nm, "BarePairAction", a, b, table=tab, max_level=0, use_long_range=True, k_cut=k_cut))
    path = host.Path(cfg, n_clones=2)
    na = np.array([[0, 0, 0], [1, 1, 0], [1, 0, 1], [0, 1, 1]], dtype=float) - 0.5
    cl = np.array([[1, 0, 0], [0, 1, 0], [0, 0, 1], [1, 1, 1]], dtype=float) - 0.5
    path.SetPositions(0, np.stack([na.reshape(4, 1, 3)] * 2))
    path.SetPositions(1, np.stack([cl.reshape(4, 1, 3)] * 2))
    v = sum(act.Potential() for act in path.actions)
    path.close()
    assert np.all(np.abs(v / 4.0 - (-1.7475645946331822)) < 2e-7), v / 4.0
