"""simpimc_b200 -- B200-native (sm_100a) action-evaluation path of path-integral Monte Carlo
behind simpimc's `Action` operator API.

    csrc/       hand-written CUDA kernels + the C ABI (include/simpimc_b200.h)
    capi.py     ctypes declarations / loader of csrc/libsimpimc_b200.so (no CPU fallback)
    host.py     Path / PairAction / estimators mirroring the reference's interface
    system.py   system description + the synthetic configurations of the BASELINE shapes
    tables.py   synthetic pair-action tables in the reference's layouts
"""
__version__ = "0.1.0"
