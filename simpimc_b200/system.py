"""Host-side description of a path-integral system: the fields the reference reads from its
XML input (<System>, <Species>, <Action>; src/data_structures/path_class.h:25-69,
species_class.h:48-60, actions/action_class.h:25-33, pair_action_class.h:208-222) plus the
synthetic configurations SURVEY.md section 8(d) defines for the BASELINE shapes.
"""
import math
from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np

from . import tables as T


@dataclass
class SpeciesConfig:
    name: str
    n_part: int
    lam: float  # hbar^2/2m, the XML attribute "lambda"


@dataclass
class ActionConfig:
    name: str
    type: str  # "IlkkaPairAction" | "BarePairAction" | "DavidPairAction" | "Kinetic"
    species_a: str
    species_b: str = ""
    table: Optional[dict] = None
    max_level: int = 0
    use_long_range: bool = False
    k_cut: Optional[float] = None
    n_order: int = 0
    is_coulomb: bool = False
    n_images: int = 0


@dataclass
class SystemConfig:
    n_d: int
    n_bead: int
    beta: float
    L: float
    pbc: bool = True
    k_cut: Optional[float] = None  # <System k_cut=...>; default 2*pi/L gives no k vectors
    species: List[SpeciesConfig] = field(default_factory=list)
    actions: List[ActionConfig] = field(default_factory=list)
    moves: List[dict] = field(default_factory=list)
    observables: List[dict] = field(default_factory=list)

    @property
    def tau(self):
        return self.beta / (1.0 * self.n_bead)

    def species_index(self, name):
        for i, s in enumerate(self.species):
            if s.name == name:
                return i
        raise KeyError(name)


def ueg_parameters(N, rs=1.0, theta=1.0, polarized=True):
    """inputs/e-gas/gen_e_pa.py:14-31: box, k_cut = 14/(L/2) and beta = 1/(theta T_F)."""
    if polarized:
        TF = 0.5 * (9.0 * math.pi / 2.0) ** (2.0 / 3.0) / (rs ** 2)
    else:
        TF = 0.5 * (9.0 * math.pi / 4.0) ** (2.0 / 3.0) / (rs ** 2)
    beta = 1.0 / (theta * TF)
    L = pow(N * (4.0 / 3.0) * math.pi * rs ** 3, 1.0 / 3.0)
    k_cut = 14.0 / (L / 2.0)
    return L, k_cut, beta


def ueg_config(N=256, M=128, rs=1.0, theta=1.0, action="IlkkaPairAction", use_long_range=True, n_xy=100, n_r_long=1000,
               with_kinetic=False, david_grid="LOG", david_n_grid=200, david_n_order=2, xy_asym=None):
    """Uniform electron gas, one species "e" (SURVEY.md 8(d), config C3 at the defaults)."""
    L, k_cut, beta = ueg_parameters(N, rs, theta)
    tau = beta / M
    cfg = SystemConfig(n_d=3, n_bead=M, beta=beta, L=L, pbc=True, k_cut=k_cut)
    cfg.species.append(SpeciesConfig("e", N, 0.5))
    if with_kinetic:
        cfg.actions.append(ActionConfig("Kinetic", "Kinetic", "e"))
    if action == "IlkkaPairAction":
        # xy_asym = (n_y, y_r_max, asym): an off-diagonal table with its own y grid and no x <-> y symmetry
        extra = {} if xy_asym is None else dict(n_y=xy_asym[0], y_r_max=xy_asym[1], asym=xy_asym[2])
        tab = T.make_ilkka_table(1.0, tau, L, k_cut, use_long_range=use_long_range, n_xy=n_xy, n_r_long=n_r_long, **extra)
    elif action == "BarePairAction":
        tab = T.make_bare_table(1.0, L, k_cut, use_long_range=use_long_range, n_r_long=n_r_long)
    elif action == "DavidPairAction":
        tab = T.make_david_table(1.0, tau, n_order=david_n_order, grid_type=david_grid, n_grid=david_n_grid,
                                 r_end=0.95 * math.sqrt(3.0) * L / 2.0, L=L, k_cut=k_cut, use_long_range=use_long_range)
    else:
        raise ValueError(action)
    cfg.actions.append(ActionConfig("CoulombEE", action, "e", "e", table=tab, max_level=0, use_long_range=use_long_range,
                                    k_cut=k_cut, n_order=david_n_order if action == "DavidPairAction" else 0))
    return cfg


def plasma_config(Ne=8, Np=8, M=16, rs=1.0, theta=1.0, n_xy=60, n_r_long=400, pp_action="BarePairAction", ep_action="IlkkaPairAction"):
    """Two-species hydrogen-like plasma (config C5 shape at reduced size): e-e, e-p Ilkka
    actions and a p-p action -- Bare by default (as inputs/C/c.xml mixes the types), Ilkka with
    pp_action="IlkkaPairAction" -- all with long range."""
    L, k_cut, beta = ueg_parameters(Ne, rs, theta)
    tau = beta / M
    cfg = SystemConfig(n_d=3, n_bead=M, beta=beta, L=L, pbc=True, k_cut=k_cut)
    cfg.species.append(SpeciesConfig("e", Ne, 0.5))
    cfg.species.append(SpeciesConfig("p", Np, 0.5 / 1836.15267))
    cfg.actions.append(ActionConfig("CoulombEE", "IlkkaPairAction", "e", "e", max_level=0, use_long_range=True, k_cut=k_cut,
                                    table=T.make_ilkka_table(1.0, tau, L, k_cut, n_xy=n_xy, n_r_long=n_r_long)))
    r_end = 0.95 * math.sqrt(3.0) * L / 2.0
    if ep_action == "DavidPairAction":   # a David e-p action (different species), no long range
        cfg.actions.append(ActionConfig("CoulombEP", "DavidPairAction", "e", "p", max_level=0, use_long_range=False, k_cut=k_cut, n_order=2,
                                        table=T.make_david_table(-1.0, tau, n_order=2, r_end=r_end, L=L, k_cut=k_cut, sigma=0.4)))
    else:
        cfg.actions.append(ActionConfig("CoulombEP", "IlkkaPairAction", "e", "p", max_level=0, use_long_range=True, k_cut=k_cut,
                                        table=T.make_ilkka_table(-1.0, tau, L, k_cut, n_xy=n_xy, n_r_long=n_r_long, sigma=0.4)))
    if pp_action == "DavidPairAction":
        cfg.actions.append(ActionConfig("CoulombPP", "DavidPairAction", "p", "p", max_level=0, use_long_range=False, k_cut=k_cut, n_order=2,
                                        table=T.make_david_table(1.0, tau, n_order=2, grid_type="LINEAR", n_grid=150, r_end=r_end, L=L,
                                                                 k_cut=k_cut, sigma=0.3)))
        return cfg
    if pp_action == "IlkkaPairAction":
        pp_table = T.make_ilkka_table(1.0, tau, L, k_cut, n_xy=n_xy, n_r_long=n_r_long, sigma=0.3)
    else:
        pp_table = T.make_bare_table(1.0, L, k_cut, n_r_long=n_r_long)
    cfg.actions.append(ActionConfig("CoulombPP", pp_action, "p", "p", max_level=0, use_long_range=True, k_cut=k_cut, table=pp_table))
    return cfg


def synthetic_paths(cfg, sp, clone, seed=12345):
    """Closed Brownian-bridge paths R[particle][bead][dim] for species index `sp` of clone
    `clone`: centres uniform in [-L/2, L/2)^3, bead b = centre + bridge with
    variance 2*lambda*tau per link; generator numpy default_rng(seed + clone) advanced per
    species in order."""
    rng = np.random.default_rng(seed + clone)
    out = None
    for i, s in enumerate(cfg.species):
        N, M, nd = s.n_part, cfg.n_bead, cfg.n_d
        half = cfg.L / 2.0 if cfg.pbc else 1.0
        centres = rng.uniform(-half, half, size=(N, 1, nd))
        steps = rng.normal(0.0, math.sqrt(2.0 * s.lam * cfg.tau), size=(N, M, nd))
        walk = np.cumsum(steps, axis=1)
        frac = (np.arange(1, M + 1) / M).reshape(1, M, 1)
        bridge = walk - frac * walk[:, -1:, :]
        R = centres + bridge
        if i == sp:
            out = np.ascontiguousarray(R)
    return out


def egas_config(N=7, M=1280, n_xy=100, n_r_long=1000):
    """inputs/e-gas/e-gas.xml (BASELINE config C1): N polarized electrons at r_s = 1, theta = 0.1
    (beta = 3.42075, L = 3.08363 and k_cut = 14/(L/2) = 9.08021 for the shipped N = 7), M = 1280,
    one IlkkaPairAction with long range."""
    return ueg_config(N=N, M=M, rs=1.0, theta=0.1, n_xy=n_xy, n_r_long=n_r_long)


def hatom_config(M=320, n_xy=100):
    """inputs/h-atom/h-ilkka.xml (config C2): one electron and one classical proton (lambda = 0),
    no periodic box, beta = 40, an e-p IlkkaPairAction without long range."""
    beta = 40.0
    cfg = SystemConfig(n_d=3, n_bead=M, beta=beta, L=1000.0, pbc=False, k_cut=None)
    cfg.species.append(SpeciesConfig("e", 1, 0.5))
    cfg.species.append(SpeciesConfig("p", 1, 0.0))
    tab = T.make_ilkka_table(-1.0, beta / M, 10.0, 1.0, use_long_range=False, n_xy=n_xy, sigma=0.4)
    cfg.actions.append(ActionConfig("Coulomb", "IlkkaPairAction", "e", "p", table=tab, max_level=0, use_long_range=False))
    return cfg


def hatom_paths(cfg, clone, seed=12345):
    """Proton at rest at the origin on every slice, electron on a closed bridge around it."""
    rng = np.random.default_rng(seed + clone)
    M = cfg.n_bead
    steps = rng.normal(0.0, math.sqrt(2.0 * 0.5 * cfg.tau), size=(1, M, 3))
    walk = np.cumsum(steps, axis=1)
    frac = (np.arange(1, M + 1) / M).reshape(1, M, 1)
    e = walk - frac * walk[:, -1:, :] + rng.normal(0.0, 0.7, size=(1, 1, 3))
    p = np.zeros((1, M, 3))
    return [np.ascontiguousarray(e), p]


def carbon_config(n_xy=60, n_r_long=400):
    """inputs/C/c.xml (config C4): 10 e-up + 10 e-down + 20 p + 1 C in a periodic box, M = 16,
    nine IlkkaPairActions and one BarePairAction (C-C, a one-particle species: constant), all
    with long range."""
    M, beta, L, k_cut = 16, 0.031577464, 1.53327785176, 18.2615303337
    tau = beta / M
    cfg = SystemConfig(n_d=3, n_bead=M, beta=beta, L=L, pbc=True, k_cut=k_cut)
    for name, n, lam in (("eU", 10, 0.5), ("eD", 10, 0.5), ("p", 20, 0.00027216030018605583), ("C", 1, 0.000022857496211250974)):
        cfg.species.append(SpeciesConfig(name, n, lam))
    charge = {"eU": -1.0, "eD": -1.0, "p": 1.0, "C": 6.0}
    pairs = [("eU", "eU"), ("eU", "eD"), ("eD", "eD"), ("eU", "p"), ("eD", "p"), ("eU", "C"), ("eD", "C"), ("p", "p"), ("p", "C")]
    for a, b in pairs:
        tab = T.make_ilkka_table(charge[a] * charge[b], tau, L, k_cut, n_xy=n_xy, xy_r_max=20.0, n_r_long=n_r_long, sigma=0.15)
        cfg.actions.append(ActionConfig("Coulomb" + a + b, "IlkkaPairAction", a, b, table=tab, max_level=0, use_long_range=True, k_cut=k_cut))
    cfg.actions.append(ActionConfig("CoulombCC", "BarePairAction", "C", "C", max_level=0, use_long_range=True, k_cut=k_cut,
                                    table=T.make_bare_table(36.0, L, k_cut, n_r_long=n_r_long)))
    return cfg
