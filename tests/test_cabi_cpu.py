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


def test_philox_known_answers():
    """Random123 kat_vectors for philox4x32-10: the host mirror of the device stream."""
    from simpimc_b200 import philox as P
    kat = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
           ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0), (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for ctr, key, out in kat:
        assert tuple(int(x) for x in P.philox4x32(*ctr, *key)) == out
    u = P.uniform_from_bits(0, 0), P.uniform_from_bits(0xffffffff, 0xffffffff)
    assert 0.0 < u[0] < u[1] <= 1.0  # never 0: log(u) is finite
