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


def _einspline_interval(grid, x):
    """nubspline general_grid_reverse_map: last i with grid[i] <= x, 0 below the grid, n-1 at or above its end."""
    i = np.searchsorted(grid, x, side="right") - 1
    i = np.where(x <= grid[0], 0, i)
    return np.where(x >= grid[-1], len(grid) - 1, i).astype(np.int32)


def _intervals_from_library(kind, grid, x):
    import ctypes as C
    from simpimc_b200 import capi
    L = capi.lib()
    grid = np.ascontiguousarray(grid, dtype=np.float64)
    x = np.ascontiguousarray(x, dtype=np.float64)
    out = np.zeros(len(x), dtype=np.int32)
    n_keys = C.c_int32(0)
    rc = L.pimc_debug_interval_table(kind, len(grid), grid.ctypes.data, len(x), x.ctypes.data, out.ctypes.data, C.addressof(n_keys))
    return rc, out, n_keys.value


def _probe_points(grid, rng):
    """Random points, the knots themselves and their floating-point neighbours, points outside the grid."""
    lo, hi = grid[0], grid[-1]
    xs = [rng.uniform(0.0, 1.05 * hi, 200000), grid, np.nextafter(grid, 0.0), np.nextafter(grid, np.inf),
          np.array([0.0, 0.5 * lo, hi, 1.5 * hi, 10.0 * hi]), 10.0 ** rng.uniform(np.log10(max(lo, 1e-12)), np.log10(hi), 50000)]
    return np.concatenate(xs)


def test_interval_tables_reproduce_the_einspline_interval_bit_for_bit():
    """Integer work: the fast kernels' interval lookup (uniform buckets for the squarer's OPTIMIZED / linear grids,
    IEEE bit-pattern buckets for David's logarithmic grid) must return einspline's interval index exactly, for
    every x including the knots and their neighbours.  Host twin of the device arithmetic (pimc_debug_interval_table)."""
    rng = np.random.default_rng(17)
    opt = T.gen_grid("OPTIMIZED", 1.0e-4, 8.86, 1000)
    lin = T.gen_grid("LINEAR", 1.0e-3, 8.4, 400)
    xy = T.gen_grid("OPTIMIZED", 0.0, 100.0, 100)
    log = 1.0e-3 * np.exp(np.arange(200) * (np.log(8.4 / 1.0e-3) / 199))
    for kind, grid, must_build in ((0, opt, True), (0, lin, True), (0, xy, True), (1, log, True), (1, lin, True),
                                   (1, T.gen_grid("OPTIMIZED", 1.0e-2, 8.86, 300), True)):
        x = _probe_points(grid, rng)
        rc, got, n_keys = _intervals_from_library(kind, grid, x)
        assert rc == 0, (kind, len(grid))
        assert 0 < n_keys <= 16384
        assert np.array_equal(got, _einspline_interval(grid, x)), (kind, len(grid))
    # a logarithmic grid is far too fine near its start for a uniform table: refused, not approximated
    rc, _, _ = _intervals_from_library(0, 1.0e-6 * np.exp(np.arange(400) * (np.log(1e7) / 399)), np.array([1.0]))
    assert rc != 0
    # a grid starting at 0 has no bit-pattern table (its first bucket would be unbounded below in the exponent)
    rc, _, _ = _intervals_from_library(1, xy, np.array([1.0]))
    assert rc != 0


def test_bucket_centred_splines_equal_the_interval_form():
    """The fast kernels evaluate 1-D splines with a uniform interval table (the long-range r-space parts, v(r)) from
    bucket-centred records: one small load decides the piece, no knot is fetched (csrc/pair_fast.cuh: FastPP1,
    csrc/spline_build.h: BuildLR2).  Host twin of the device arithmetic (pimc_debug_bucket_spline): the same piecewise
    polynomial as the interval form to rounding -- random points, every knot and its floating-point neighbours (where
    either neighbouring piece may be taken: C2 continuity), the grid ends, points outside the grid (clamped)."""
    import ctypes as C
    from simpimc_b200 import capi
    L = capi.lib()
    rng = np.random.default_rng(23)
    for grid in (T.gen_grid("OPTIMIZED", 1.0e-4, 8.86, 1000), T.gen_grid("LINEAR", 1.0e-3, 8.4, 400), T.gen_grid("OPTIMIZED", 0.0, 5.0, 37)):
        grid = np.ascontiguousarray(grid, dtype=np.float64)
        vals = np.ascontiguousarray(np.exp(-0.3 * grid) / (grid + 0.2) + 0.05 * np.sin(3.0 * grid))
        x = np.ascontiguousarray(_probe_points(grid, rng))
        a, b = np.zeros(len(x)), np.zeros(len(x))
        n_keys = C.c_int32(0)
        rc = L.pimc_debug_bucket_spline(len(grid), grid.ctypes.data, vals.ctypes.data, len(x), x.ctypes.data, a.ctypes.data, b.ctypes.data,
                                        C.addressof(n_keys))
        assert rc == 0, capi.last_error() if hasattr(capi, "last_error") else rc
        assert 0 < n_keys.value <= 16384
        scale = np.max(np.abs(vals))
        assert np.max(np.abs(a - b)) <= 4e-16 * scale, (len(grid), np.max(np.abs(a - b)))
        # and the interval form is the spline: it interpolates its data
        at_knots = np.zeros(len(grid))
        rc = L.pimc_debug_bucket_spline(len(grid), grid.ctypes.data, vals.ctypes.data, len(grid), grid.ctypes.data, at_knots.ctypes.data,
                                        b[:len(grid)].ctypes.data, C.addressof(n_keys))
        assert rc == 0
        assert np.max(np.abs(at_knots - vals)) <= 1e-13 * scale
    # a grid too fine near its start for a uniform table is refused
    fine = np.ascontiguousarray(1.0e-6 * np.exp(np.arange(400) * (np.log(1e7) / 399)))
    rc = L.pimc_debug_bucket_spline(len(fine), fine.ctypes.data, fine.ctypes.data, 1, fine.ctypes.data, a.ctypes.data, b.ctypes.data, C.addressof(n_keys))
    assert rc != 0
