"""Synthetic pair-action tables in the layouts the reference's constructors read.

No pair-action table ships with the reference (they come from Fortran squarers,
scripts/pagen/ilkkaSquarer, that cannot be built here), so parity fixtures and the bench
use analytic surrogates written in exactly the dataset layout of

* IlkkaPairAction  (src/actions/pair_action/ilkka_pair_action_class.h:266-418; written by
  scripts/pagen/IlkkaSquarer.py:118-163),
* BarePairAction   (src/actions/pair_action/bare_pair_action_class.h:38-86),
* DavidPairAction  (src/actions/pair_action/david_pair_action_class.h:194-336; written by
  scripts/pagen/DavidParse.py:160-225).

Grids follow scripts/pagen/GenGrid.py:17-24 and the long-range parts follow the
StandardEwald closed forms of scripts/pagen/Ewald.py:281-318 (values rounded through
'%.10E' like the files that script writes).

A table is a plain ``dict`` {hdf5-style dataset name: numpy array / scalar / str}.  The
same dict feeds (i) the product (simpimc_b200.host packs it into the C-ABI structs) and
(ii) the oracle builds (written as a flat "PTAB1" container, see write_ptab).
"""
import math
import struct

import numpy as np
from scipy.special import erf


# ----------------------------------------------------------------------------- grids
def gen_grid(grid_type, r_min, r_max, n_grid):
    """scripts/pagen/GenGrid.py:3-27."""
    if grid_type == "LINEAR":
        return np.linspace(r_min, r_max, num=n_grid, endpoint=True)
    if grid_type == "LOG":
        return np.logspace(math.log10(r_min), math.log10(r_max), num=n_grid, endpoint=True)
    if grid_type == "OPTIMIZED":
        rs = [r_min]
        a = math.exp(-0.58015) * pow(1.0 * (n_grid - 1), 0.506494)
        dr = r_max / ((n_grid - 1) - a)
        for grid_i in range(n_grid - 1):
            fi = 2.0 * (grid_i + 1) * a / (n_grid - 1)
            rs.append(rs[grid_i] + (1.0 - (1.0 / (math.exp(fi) + 1.0))) * dr)
        return np.array(rs)
    raise ValueError("unrecognized grid type %r" % grid_type)


def _r10(x):
    """Round through the '%.10E' text format the pagen scripts write (Ewald.py:289-318)."""
    x = np.asarray(x, dtype=np.float64)
    flat = np.array([float("%.10E" % v) for v in x.ravel()])
    return flat.reshape(x.shape) if x.shape else float(flat[0])


# --------------------------------------------------------------------------- k shells
def k_shell_magnitudes(L, k_cut, n_d=3):
    """Sorted distinct |k| with 0 < |k| < k_cut on the 2*pi/L lattice (Ewald.py:300-318)."""
    kb = 2.0 * math.pi / L
    m = int(math.ceil(1.1 * k_cut / kb))
    rng = np.arange(-m, m + 1)
    grids = np.meshgrid(*([rng] * n_d), indexing="ij")
    k2 = sum((g * kb) ** 2 for g in grids).ravel()
    k2 = k2[(k2 > 0) & (k2 < k_cut * k_cut)]
    mags = np.sort(np.sqrt(k2))
    out = []
    prev = -1.0
    for v in mags:
        if abs(v - prev) > 1.0e-8:
            out.append(v)
        prev = v
    return np.array(out)


def standard_ewald(z1z2, cofactor, L, k_cut, r_grid, n_d=3):
    """StandardEwald breakup (Ewald.py:281-318): returns dict with long_r_0, long_r,
    k (first entry 0), long_k (first entry = long_k_0), long_k_0."""
    assert n_d == 3
    vol = L ** n_d
    r_cut = L / 2.0
    alpha = math.sqrt(k_cut / (2.0 * r_cut))
    long_r_0 = 2.0 * cofactor * z1z2 * alpha / math.sqrt(math.pi)
    long_r = cofactor * z1z2 * erf(alpha * r_grid) / r_grid
    long_k_0 = -4.0 * math.pi * cofactor * z1z2 / (4.0 * alpha * alpha * vol)
    ks = k_shell_magnitudes(L, k_cut, n_d)
    long_k = (4.0 * math.pi * cofactor * z1z2 / (ks * ks * vol)) * np.exp(-ks * ks / (4.0 * alpha * alpha))
    return {
        "long_r_0": _r10(long_r_0),
        "r_long": _r10(r_grid),
        "long_r": _r10(long_r),
        "k": _r10(np.concatenate([[0.0], ks])),
        "long_k": _r10(np.concatenate([[long_k_0], long_k])),
        "long_k_0": _r10(long_k_0),
    }


# --------------------------------------------------------------------- Ilkka / Bare
def _soft_coulomb(r, sigma):
    """erf(r/sigma)/r, finite (2/(sigma sqrt(pi))) at r = 0."""
    r = np.asarray(r, dtype=np.float64)
    out = np.empty_like(r)
    small = r < 1e-12
    out[small] = 2.0 / (sigma * math.sqrt(math.pi))
    out[~small] = erf(r[~small] / sigma) / r[~small]
    return out


def make_ilkka_table(z1z2, tau, L, k_cut, use_long_range=True, n_xy=100, xy_r_max=100.0, n_r=1000, r_min=1.0e-4,
                     r_max=100.0, n_r_long=1000, sigma=0.5, n_y=None, y_r_max=None, asym=0.0):
    """Analytic stand-in for an Ilkka squarer table.

    Grids as inputs/e-gas/gen_e_pa.py:40-69 configures them: OPTIMIZED n=100 up to 100 for
    the off-diagonal u(x,y), du(x,y) (the squarer's own grid starts at 0,
    scripts/pagen/ilkkaSquarer/init.f90:962-968), OPTIMIZED n=1000 on [1e-4,100] for v(r),
    OPTIMIZED n=1000 on [1e-4, sqrt(3) L/2] for the long-range r parts.

    The squarer's table is square, on one radial grid for both axes, and the surrogate above is
    symmetric in (x, y) -- which would hide an x / y transposition or a row- versus column-major
    mix-up anywhere between this dict and the kernels.  n_y / y_r_max give the y axis its own
    grid (n_x != n_y) and asym > 0 multiplies the surfaces by 1 + asym X / (1 + X): a table no
    transposition leaves invariant (parity config "ilkka_asym").
    """
    t = {}
    xs = gen_grid("OPTIMIZED", 0.0, xy_r_max, n_xy)
    ys = xs if n_y is None else gen_grid("OPTIMIZED", 0.0, xy_r_max if y_r_max is None else y_r_max, n_y)
    X, Y = np.meshgrid(xs, ys, indexing="ij")
    D = X - Y
    vs_x, vs_y = _soft_coulomb(X, sigma), _soft_coulomb(Y, sigma)
    # u: endpoint average of a softened Coulomb times a smooth off-diagonal damping
    u_xy = tau * z1z2 * 0.5 * (vs_x + vs_y) * (1.0 - 0.35 * D * D / (4.0 * sigma * sigma + D * D))
    # du/dbeta: a different smooth surface of the same scale
    ws_x = _soft_coulomb(X, 0.8 * sigma) * (1.0 + 0.25 * np.exp(-X * X / (2 * sigma * sigma)))
    ws_y = _soft_coulomb(Y, 0.8 * sigma) * (1.0 + 0.25 * np.exp(-Y * Y / (2 * sigma * sigma)))
    du_xy = z1z2 * 0.5 * (ws_x + ws_y) * (1.0 - 0.2 * D * D / (3.0 * sigma * sigma + D * D))
    if asym:
        u_xy = u_xy * (1.0 + asym * X / (1.0 + X))
        du_xy = du_xy * (1.0 + 0.5 * asym * X / (1.0 + X))
    for name, arr in (("u", u_xy), ("du", du_xy)):
        t[name + "/off_diag/n_x"] = np.uint32(len(xs))
        t[name + "/off_diag/n_y"] = np.uint32(len(ys))
        t[name + "/off_diag/x"] = _r10(xs)
        t[name + "/off_diag/y"] = _r10(ys)
        t[name + "/off_diag/%s_xy" % name] = _r10(arr)
    rv = gen_grid("OPTIMIZED", r_min, r_max, n_r)
    t["v/diag/n_r"] = np.uint32(n_r)
    t["v/diag/r"] = _r10(rv)
    t["v/diag/v_r"] = _r10(z1z2 / rv)
    if use_long_range:
        rl = gen_grid("OPTIMIZED", r_min, math.sqrt(3.0) * L / 2.0, n_r_long)
        for name, cof in (("u", tau), ("du", 1.0), ("v", 1.0)):
            e = standard_ewald(z1z2, cof, L, k_cut, rl)
            t[name + "/diag/n_r_long"] = np.uint32(n_r_long)
            t[name + "/diag/r_long"] = e["r_long"]
            t[name + "/diag/%s_long_r" % name] = e["long_r"]
            t[name + "/diag/%s_long_r_0" % name] = np.float64(e["long_r_0"])
            t[name + "/diag/n_k"] = np.uint32(len(e["k"]))
            t[name + "/diag/k"] = e["k"]
            t[name + "/diag/%s_long_k" % name] = e["long_k"]
            t[name + "/diag/%s_long_k_0" % name] = np.float64(e["long_k_0"])
    return t


def make_bare_table(z1z2, L, k_cut, use_long_range=True, n_r=1000, r_min=1.0e-4, r_max=100.0, n_r_long=1000):
    """The v/diag subset BarePairAction reads (bare_pair_action_class.h:38-86)."""
    t = {}
    rv = gen_grid("OPTIMIZED", r_min, r_max, n_r)
    t["v/diag/n_r"] = np.uint32(n_r)
    t["v/diag/r"] = _r10(rv)
    t["v/diag/v_r"] = _r10(z1z2 / rv)
    if use_long_range:
        rl = gen_grid("OPTIMIZED", r_min, math.sqrt(3.0) * L / 2.0, n_r_long)
        e = standard_ewald(z1z2, 1.0, L, k_cut, rl)
        t["v/diag/n_r_long"] = np.uint32(n_r_long)
        t["v/diag/r_long"] = e["r_long"]
        t["v/diag/v_long_r"] = e["long_r"]
        t["v/diag/v_long_r_0"] = np.float64(e["long_r_0"])
        t["v/diag/n_k"] = np.uint32(len(e["k"]))
        t["v/diag/k"] = e["k"]
        t["v/diag/v_long_k"] = e["long_k"]
        t["v/diag/v_long_k_0"] = np.float64(e["long_k_0"])
    return t


# ------------------------------------------------------------------------- David
def make_david_table(z1z2, tau, n_order=2, grid_type="LOG", r_start=1.0e-3, r_end=12.0, n_grid=200, L=None, k_cut=None,
                     use_long_range=False, sigma=0.5):
    """Analytic stand-in for a David-squarer table (david_pair_action_class.h:194-336).

    data[n_grid][n_val][n_tau] in file (C) order with n_tau = 1 (max_level = 0), which is
    what the reference's raw read into cube(n_val, n_grid, n_tau) expects.
    """
    n_val = 1 + sum(1 + i for i in range(1, n_order + 1))
    r = gen_grid(grid_type, r_start, r_end, n_grid) if grid_type != "LOG" else \
        r_start * np.exp(np.arange(n_grid) * (math.log(r_end / r_start) / (n_grid - 1)))
    vs = _soft_coulomb(r, sigma)
    u = np.zeros((n_grid, n_val, 1))
    du = np.zeros((n_grid, n_val, 1))
    u[:, 0, 0] = tau * z1z2 * vs * np.exp(-r / 6.0)
    du[:, 0, 0] = z1z2 * (_soft_coulomb(r, 0.8 * sigma) * np.exp(-r / 6.0) - z1z2 / r * 0.0)
    idx = 1
    for k in range(1, n_order + 1):
        for j in range(0, k + 1):
            scale = 0.3 / (1.0 + idx)
            u[:, idx, 0] = tau * z1z2 * scale * np.exp(-r * (0.5 + 0.1 * j)) / (1.0 + r) ** (2 * k)
            du[:, idx, 0] = z1z2 * scale * 0.7 * np.exp(-r * (0.6 + 0.1 * j)) / (1.0 + r) ** (2 * k)
            idx += 1
    t = {}
    for grp, arr in (("u_kj_%d" % n_order, u), ("du_kj_dbeta_%d" % n_order, du)):
        t[grp + "/grid/start"] = np.float64(r_start)
        t[grp + "/grid/end"] = np.float64(r_end)
        t[grp + "/grid/n_grid_points"] = np.uint32(n_grid)
        t[grp + "/grid/type"] = grid_type
        t[grp + "/grid/grid_points"] = r.copy()
        t[grp + "/taus"] = np.array([tau])
        t[grp + "/data"] = arr
    t["potential/data"] = z1z2 * vs
    if use_long_range:
        ks = k_shell_magnitudes(L, k_cut)
        alpha = math.sqrt(k_cut / L)
        uk0 = -4.0 * math.pi * z1z2 / (4.0 * alpha * alpha)
        uk = (4.0 * math.pi * z1z2 / (ks * ks)) * np.exp(-ks * ks / (4.0 * alpha * alpha))
        t["long_range/k_cut"] = np.float64(k_cut)
        t["long_range/n_k"] = np.uint32(len(ks) + 1)
        t["long_range/k_points"] = _r10(np.concatenate([[0.0], ks]))
        t["long_range/u_k"] = _r10(np.concatenate([[uk0], uk]))
        t["squarer/v_image"] = np.float64(_r10(2.0 * z1z2 * alpha / math.sqrt(math.pi)))
    return t


# --------------------------------------------------------------------- container
def write_ptab(path, table):
    """Flat container the oracle builds read instead of HDF5 (oracle/shim/scaffold/io)."""
    with open(path, "wb") as f:
        f.write(b"PTAB1\n")
        f.write(struct.pack("<I", len(table)))
        for name, val in table.items():
            nb = name.encode()
            f.write(struct.pack("<I", len(nb)))
            f.write(nb)
            if isinstance(val, str):
                data = val.encode()
                f.write(struct.pack("<BI", 3, 1))
                f.write(struct.pack("<Q", len(data)))
                f.write(data)
                continue
            arr = np.asarray(val)
            if arr.dtype == np.uint32:
                code = 1
            elif arr.dtype == np.int32:
                code = 2
            else:
                code = 0
                arr = arr.astype(np.float64)
            arr = np.ascontiguousarray(arr)
            f.write(struct.pack("<BI", code, arr.ndim))
            for d in arr.shape:
                f.write(struct.pack("<Q", d))
            f.write(arr.tobytes())


def read_ptab(path):
    out = {}
    with open(path, "rb") as f:
        assert f.read(6) == b"PTAB1\n"
        (n,) = struct.unpack("<I", f.read(4))
        for _ in range(n):
            (ln,) = struct.unpack("<I", f.read(4))
            name = f.read(ln).decode()
            code, nd = struct.unpack("<BI", f.read(5))
            dims = [struct.unpack("<Q", f.read(8))[0] for _ in range(nd)]
            count = int(np.prod(dims)) if dims else 1
            if code == 3:
                out[name] = f.read(count).decode()
            else:
                dt = {0: np.float64, 1: np.uint32, 2: np.int32}[code]
                arr = np.frombuffer(f.read(count * np.dtype(dt).itemsize), dtype=dt).reshape(dims)
                out[name] = arr.copy() if dims else arr.reshape(())[()]
    return out


# ------------------------------------------------------------------------ HDF5 table files
def read_h5_table(path):
    """A reference pair-action table file (HDF5, as scripts/pagen/IlkkaSquarer.py:118-163 / DavidParse.py:160-225
    write it) -> the flat dict {dataset path: array | scalar | str} the packers in capi.py take, dataset paths
    verbatim.  Read with simpimc_b200.h5lite (no HDF5 library in this build)."""
    from . import h5lite
    out = {}
    for key, val in h5lite.read(path).items():
        if isinstance(val, np.ndarray) and val.dtype.kind == "f":
            val = np.ascontiguousarray(val, dtype=np.float64)      # row-major, as the reference's raw reads expect
        elif isinstance(val, np.ndarray) and val.dtype == object and val.size == 1:
            val = str(val.reshape(-1)[0])
        elif np.isscalar(val) and isinstance(val, (np.integer,)):
            val = np.uint32(val)
        out[key] = val
    return out


def write_h5_table(path, table):
    """The inverse: a table dict as an HDF5 file with the reference's dataset paths."""
    from . import h5lite
    h5lite.write(path, table)


def load_table(path):
    """A pair-action table from any container this build understands: HDF5 (the reference's own files), the flat
    PTAB1 container, or a .npz with '|' for '/' in the keys."""
    from . import h5lite
    if str(path).endswith(".npz"):
        f = np.load(path, allow_pickle=False)
        return {k.replace("|", "/"): (str(f[k]) if f[k].dtype.kind in "US" else f[k]) for k in f.files}
    if h5lite.is_hdf5(path):
        return read_h5_table(path)
    return read_ptab(path)
