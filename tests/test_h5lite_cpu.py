"""CPU tests of simpimc_b200.h5lite, the package's own reader / writer for the subset of HDF5 the reference's table
and output files use (no HDF5 library exists in this image):

* the reader against the one file here that the HDF5 LIBRARY ITSELF wrote -- scipy's MATLAB-7.3 test file, HDF5 behind
  a 512-byte user block -- with the expected numbers taken from scipy's reading of the same variable in MATLAB's older
  format;
* writer -> reader round trips of every table family (Ilkka, Bare, David with its grid-type string), and the CPU oracle
  evaluating a table that went through the HDF5 file: identical energies;
* the chunked + deflate + shuffle path (how extendable datasets such as the reference's block output are stored) on a
  dataset assembled by hand from the format's primitives;
* format features outside the subset are rejected with a message, not misread."""
import os
import struct
import zlib

import numpy as np
import pytest

from simpimc_b200 import h5lite, system as S, tables as T

MAT = os.path.join(os.path.dirname(__import__("scipy").__file__), "io", "matlab", "tests", "data")


def test_reads_a_file_written_by_the_hdf5_library():
    f = os.path.join(MAT, "testhdf5_7.4_GLNX86.mat")
    if not os.path.exists(f):
        pytest.skip("scipy's MATLAB 7.3 test file is not installed")
    assert h5lite.is_hdf5(f)
    out, attrs = h5lite.read(f, with_attributes=True)
    assert list(out) == ["testdouble"]
    import scipy.io
    ref = scipy.io.loadmat(os.path.join(MAT, "testdouble_7.4_GLNX86.mat"))["testdouble"]     # the same variable, MATLAB v7 format
    assert out["testdouble"].dtype == np.float64
    assert np.array_equal(out["testdouble"].reshape(-1), ref.reshape(-1))
    assert np.allclose(out["testdouble"].reshape(-1), np.arange(9) * np.pi / 4, rtol=1e-15)
    assert attrs["testdouble"]["MATLAB_class"] == "double"


@pytest.mark.parametrize("family", ["ilkka", "bare", "david"])
def test_table_round_trip_through_hdf5(family, tmp_path):
    L, k_cut, tau = 5.0, 5.6, 0.05
    if family == "ilkka":
        tab = T.make_ilkka_table(1.0, tau, L, k_cut, n_xy=30, n_r=200, n_r_long=150, n_y=24, y_r_max=50.0, asym=0.2)
    elif family == "bare":
        tab = T.make_bare_table(-1.0, L, k_cut, n_r=200, n_r_long=150)
    else:
        tab = T.make_david_table(1.0, tau, n_order=2, n_grid=60, L=L, k_cut=k_cut, use_long_range=True)
    f = str(tmp_path / (family + ".h5"))
    T.write_h5_table(f, tab)
    assert h5lite.is_hdf5(f)
    back = T.load_table(f)
    assert sorted(back) == sorted(tab)
    for k, v in tab.items():
        if isinstance(v, str):
            assert back[k] == v, k
        else:
            assert np.array_equal(np.asarray(back[k], dtype=np.float64), np.asarray(v, dtype=np.float64)), k
            assert np.shape(back[k]) == np.shape(v), k


def test_oracle_energy_from_an_hdf5_table(tmp_path, oracle_mod):
    """The table a maintainer hands over as e_e.h5 gives the energies of the in-memory table."""
    cfg = S.ueg_config(N=6, M=8, n_xy=40, n_r_long=200)
    f = str(tmp_path / "e_e.h5")
    T.write_h5_table(f, cfg.actions[0].table)
    cfg2 = S.ueg_config(N=6, M=8, n_xy=40, n_r_long=200)
    cfg2.actions[0].table = T.load_table(f)
    R = S.synthetic_paths(cfg, 0, 0)
    vals = []
    for c in (cfg, cfg2):
        o = oracle_mod.Oracle(c)
        o.set_positions(0, R)
        vals.append((o.dbeta(0), o.potential(0)))
        o.close()
    assert vals[0] == vals[1]


def test_many_links_and_nested_groups(tmp_path):
    data = {"g%02d/sub/x%d" % (i // 3, i): np.arange(i + 1, dtype=np.float64) * 0.5 for i in range(40)}
    data["top/name"] = "IlkkaPairAction"
    data["top/n"] = np.uint32(7)
    data["top/i64"] = np.array([[1, -2], [3, 4]], dtype=np.int64)
    data["top/empty"] = np.zeros((0,), dtype=np.float64)
    f = str(tmp_path / "many.h5")
    h5lite.write(f, data)
    back = h5lite.read(f)
    assert sorted(back) == sorted(data)
    for k, v in data.items():
        if isinstance(v, str):
            assert back[k] == v
        else:
            assert np.array_equal(back[k], v) and np.asarray(back[k]).dtype == np.asarray(v).dtype, k


def test_chunked_deflate_shuffle_dataset(tmp_path):
    """An extendable dataset as the reference's CreateExtendableDataSet / AppendDataSet leave it (io_hdf5.h:106-212):
    chunk 1 x shape(data), here additionally filtered (shuffle + deflate) and with a partial edge chunk."""
    w = h5lite._Writer()
    sb = w.alloc(96)
    full = np.arange(5 * 6, dtype="<f8").reshape(5, 6) * 1.25
    cdims = (2, 4)
    entries = []
    for i0 in range(0, 5, 2):
        for j0 in range(0, 6, 4):
            chunk = np.zeros(cdims, dtype="<f8")
            blk = full[i0:i0 + 2, j0:j0 + 4]
            chunk[:blk.shape[0], :blk.shape[1]] = blk
            raw = np.frombuffer(chunk.tobytes(), dtype=np.uint8).reshape(-1, 8).T.tobytes()    # shuffle
            raw = zlib.compress(raw)
            addr = w.alloc(len(raw))
            w.put(addr, raw)
            entries.append(((i0, j0, 0), len(raw), addr))
    node = b"TREE" + struct.pack("<BBHQQ", 1, 0, len(entries), h5lite.UNDEF, h5lite.UNDEF)
    for offs, size, addr in entries:
        node += struct.pack("<II3Q", size, 0, *offs) + struct.pack("<Q", addr)
    node += struct.pack("<II3Q", 0, 0, 6, 8, 0)
    btree = w.alloc(len(node))
    w.put(btree, node)
    space = struct.pack("<BBBB4x", 1, 2, 1, 0) + struct.pack("<4Q", 5, 6, h5lite.UNDEF, 6)       # max dims present: unlimited x 6
    dt = struct.pack("<BBBBI", 0x11, 0x20, 0x3F, 0x00, 8) + struct.pack("<HHBBBBI", 0, 64, 52, 11, 0, 52, 1023)
    layout = struct.pack("<BBBQ3I", 3, 2, 3, btree, 2, 4, 8)
    pipeline = struct.pack("<BB6x", 1, 2) + struct.pack("<HHHH", 2, 0, 1, 1) + struct.pack("<II", 8, 0) + \
        struct.pack("<HHHH", 1, 0, 1, 1) + struct.pack("<II", 6, 0)
    ds = w.object_header([w.msg(1, space), w.msg(3, dt, 1), w.msg(0xB, pipeline), w.msg(8, layout)])
    # a root group holding the dataset, through the writer's own group code
    w.dataset = lambda value: ds
    root = w.group({"x": 0})
    bt, heap = w._last_group
    s = h5lite.SIGNATURE + struct.pack("<BBBBBBBB", 0, 0, 0, 0, 0, 8, 8, 0) + struct.pack("<HHI", 4, 16, 0)
    s += struct.pack("<QQQQ", 0, h5lite.UNDEF, len(w.buf), h5lite.UNDEF) + struct.pack("<QQII", 0, root, 1, 0) + struct.pack("<QQ", bt, heap)
    w.put(sb, s)
    f = str(tmp_path / "chunked.h5")
    open(f, "wb").write(bytes(w.buf))
    back = h5lite.read(f)
    assert np.array_equal(back["x"], full)


def test_unsupported_formats_fail_loudly(tmp_path):
    f = str(tmp_path / "v2.h5")
    open(f, "wb").write(h5lite.SIGNATURE + bytes([2]) + bytes(200))
    with pytest.raises(h5lite.H5Error, match="superblock version 2"):
        h5lite.read(f)
    g = str(tmp_path / "not.h5")
    open(g, "wb").write(b"PTAB1" + bytes(100))
    assert not h5lite.is_hdf5(g)
    with pytest.raises(h5lite.H5Error, match="not an HDF5 file"):
        h5lite.read(g)
