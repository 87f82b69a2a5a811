"""Host-side table plumbing (CPU): the flat table container round-trips, and the offline HDF5
converter's tree walk (tools/h5_to_ptab.py, exercised on an h5py look-alike: no HDF5 library
in this build) reproduces a table dict that the oracle evaluates identically."""
import importlib.util
import os

import numpy as np

from simpimc_b200 import system as S, tables as T

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class FakeDataset:
    def __init__(self, val):
        self.val = val
        self.shape = np.shape(val)

    def __getitem__(self, key):
        assert key == ()
        if isinstance(self.val, str):
            return self.val.encode()          # h5py returns bytes for fixed / variable strings
        return np.asarray(self.val)


class FakeGroup(dict):
    pass


def fake_h5(table):
    root = FakeGroup()
    for path, val in table.items():
        node = root
        parts = path.split("/")
        for p in parts[:-1]:
            node = node.setdefault(p, FakeGroup())
        node[parts[-1]] = FakeDataset(val)
    return root


def load_converter():
    spec = importlib.util.spec_from_file_location("h5_to_ptab", os.path.join(ROOT, "tools", "h5_to_ptab.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def same_table(a, b):
    assert set(a) == set(b)
    for k in a:
        if isinstance(a[k], str):
            assert a[k] == b[k], k
        else:
            assert np.array_equal(np.asarray(a[k], dtype=np.float64).squeeze(), np.asarray(b[k], dtype=np.float64).squeeze()), k  # scalars are stored as 1-element arrays


def test_ptab_round_trip(tmp_path):
    for cfg in (S.ueg_config(N=7, M=8), S.ueg_config(N=7, M=8, action="BarePairAction"),
                S.ueg_config(N=7, M=8, action="DavidPairAction", use_long_range=True)):
        tab = cfg.actions[0].table
        p = str(tmp_path / "t.ptab")
        T.write_ptab(p, tab)
        same_table(tab, T.read_ptab(p))


def test_h5_tree_walk_gives_the_table_the_oracle_reads(tmp_path):
    from oracle import oracle as O
    conv = load_converter()
    for kw in (dict(), dict(action="DavidPairAction", use_long_range=False)):
        cfg = S.ueg_config(N=7, M=8, **kw)
        tab = cfg.actions[0].table
        flat = conv.flatten(fake_h5(tab))
        same_table(tab, flat)
        R = S.synthetic_paths(cfg, 0, 0)
        ref = O.Oracle(cfg)
        ref.set_positions(0, R)
        p = str(tmp_path / "conv.ptab")
        T.write_ptab(p, flat)
        cfg.actions[0].table = T.read_ptab(p)
        got = O.Oracle(cfg)
        got.set_positions(0, R)
        assert got.dbeta(0) == ref.dbeta(0)
