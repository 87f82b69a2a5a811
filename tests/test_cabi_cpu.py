"""CPU tests of the boundary: the C-ABI library loads and exports every symbol that
include/simpimc_b200.h declares; without a GPU every compute entry point fails loudly (there
is no CPU fallback); the ctypes structs match the header's layout."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "simpimc_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pimc_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from simpimc_b200 import capi
    assert os.path.exists(capi.LIB_PATH), "build the library first: python -c 'import __graft_entry__ as g; g.build()'"
    lib = ctypes.CDLL(capi.LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 40
    for n in names:
        assert hasattr(lib, n), "missing export %s" % n
    # and the Python side binds exactly the declared set
    assert sorted(capi.EXPORTS) == names


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from simpimc_b200 import host, system as S
    cfg = S.ueg_config(N=4, M=4, n_xy=20, n_r_long=50)
    with pytest.raises(RuntimeError, match="no CPU fallback|CUDA"):
        host.Path(cfg, n_clones=1)


def test_product_never_imports_the_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may touch oracle/."""
    pkg = os.path.join(ROOT, "simpimc_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cc", ".hpp")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text and "liboracle" not in text, os.path.join(dirpath, f)


def test_config_struct_layout_matches_header():
    from simpimc_b200 import capi
    # pimc_config: int32 n_d, pbc; double L, beta; int32 n_bead, n_species; ptr n_part, lambda; int32 x4
    assert ctypes.sizeof(capi.Config) == 4 + 4 + 8 + 8 + 4 + 4 + 8 + 8 + 4 * 4
    assert capi.Config.L.offset == 8 and capi.Config.n_part.offset == 32 and capi.Config.slice_lo.offset == 56
    assert ctypes.sizeof(capi.Table1D) == 24 and ctypes.sizeof(capi.Table2D) == 32
    assert ctypes.sizeof(capi.LongRange) == 24 + 8 + 8 + 8 + 8 + 8
