#!/usr/bin/env python
"""Offline converter for the reference's pair-action table files.

The reference reads its Ilkka / David / Bare tables from HDF5 (ilkka_pair_action_class.h:266-418,
david_pair_action_class.h:194-336, bare_pair_action_class.h:38-95).  Tables enter the product as a flat
dict {dataset path: array} (simpimc_b200.tables: `load_table` reads HDF5 directly through
simpimc_b200.h5lite, the flat PTAB1 container or a .npz).  This script converts a reference `.h5` table into
the flat container the C++ test shim reads -- with h5py where it exists, with h5lite otherwise; the dataset
paths are kept verbatim, so `simpimc_b200.host` packs the result exactly like the synthetic tables.

    python tools/h5_to_ptab.py e_e.h5 e_e.ptab        # or e_e.npz

`flatten(root)` is the whole logic and takes any h5py-like tree (objects with .items(); datasets
expose .shape and [()]), which is how tests/test_tables_cpu.py exercises it without h5py.
"""
import sys

import numpy as np


def _is_dataset(obj):
    return hasattr(obj, "shape") and not hasattr(obj, "items")


def flatten(root, prefix=""):
    """{ 'u/off_diag/x': ndarray, ... } from an h5py.File / Group (or a look-alike)."""
    out = {}
    for name, obj in root.items():
        path = prefix + name
        if _is_dataset(obj):
            val = obj[()]
            if isinstance(val, bytes):
                val = val.decode()
            elif isinstance(val, np.ndarray) and val.dtype.kind in "SO" and val.size == 1:
                v = val.reshape(-1)[0]
                val = v.decode() if isinstance(v, bytes) else str(v)
            elif isinstance(val, np.ndarray) and val.dtype.kind in "iu" and val.ndim == 0:
                val = np.uint32(val)
            elif isinstance(val, np.ndarray) and val.dtype.kind == "f":
                val = np.ascontiguousarray(val, dtype=np.float64)   # row-major, as the reference's raw reads expect
            out[path] = val
        else:
            out.update(flatten(obj, path + "/"))
    return out


def main(argv):
    if len(argv) != 3:
        raise SystemExit(__doc__)
    sys.path.insert(0, ".")
    from simpimc_b200 import tables
    try:
        import h5py
        with h5py.File(argv[1], "r") as f:
            t = flatten(f)
    except ImportError:     # no HDF5 library (the build container): the package's own reader of the subset these files use
        t = tables.read_h5_table(argv[1])
    if argv[2].endswith(".npz"):
        np.savez(argv[2], **{k.replace("/", "|"): v for k, v in t.items()})
    else:
        tables.write_ptab(argv[2], t)
    print("%d datasets -> %s" % (len(t), argv[2]))


if __name__ == "__main__":
    main(sys.argv)
